set -x
T=${T:-r30}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.log 2>&1; tail -1 gpurun_out/${T}_bench.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>&1; tail -1 gpurun_out/${T}_bench_reference.log | cut -c1-300
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > gpurun_out/${T}_bench_cfg3.log 2>&1; tail -1 gpurun_out/${T}_bench_cfg3.log | cut -c1-300
timeout 600 python bench.py --batch 131072 --no-cpu-baseline --steps 10 > gpurun_out/${T}_bench_cfg4_1gpu.log 2>&1; tail -1 gpurun_out/${T}_bench_cfg4_1gpu.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vo_grid2_kernel --launch-skip 3 -c 1 -o gpurun_out/${T}_vo_grid2 python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
