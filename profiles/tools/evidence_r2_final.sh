# Round-2 closing evidence on one GPU: GPU tests, smoke, default bench (with sub-records), reference arm, cfg 1 line,
# launch list of the bench command and ncu --set full of vo_grid2_kernel (each ncu pass only after the plain command exited 0)
set -x
T=${T:-r2f}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 300 gpurun_out/${T}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 300 python bench.py --workload cfg1 --no-sub > gpurun_out/${T}_bench_cfg1.json 2> gpurun_out/${T}_bench_cfg1.err
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sub --sm-reserve 8"
timeout 300 $B > gpurun_out/${T}_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vo_grid2_kernel --launch-skip 3 -c 1 -f -o gpurun_out/${T}_vo_grid2_cfg2_f64 $B > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
