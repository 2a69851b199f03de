# Round-2 third batch (one GPU): the one-kernel route for many weighting functions at config 3 -- bench line, launch list,
# ncu --set full of vo_gridgemm_kernel
set -x
B="python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-sub"
$B > gpurun_out/r2r_bench_cfg3.json 2> gpurun_out/r2r_bench_cfg3.err; tail -c 1500 gpurun_out/r2r_bench_cfg3.json
N="$B --no-graph"
$N > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2r_launches_cfg3.csv $N > gpurun_out/r2r_n0.log 2>&1
python profiles/tools/time_cfg3_vo.py > gpurun_out/r2r_time_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vo_gridgemm_kernel --launch-skip 2 -c 1 -f -o gpurun_out/r2r_gridgemm_cfg3 python profiles/tools/time_cfg3_vo.py > gpurun_out/r2r_n1.log 2>&1
cat gpurun_out/r2r_time_cfg3.log
