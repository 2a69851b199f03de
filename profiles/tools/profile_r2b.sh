# Round-2 second batch of captures (one GPU): launch lists of the default step and of the config-5 step, ncu --set full of the
# final vo_grid2 kernel and of the windowed ROM adjoint, host-side split of the VO update.
set -x
python profiles/tools/time_vo_update.py > gpurun_out/r2q_vo_update.log 2>&1; tail -1 gpurun_out/r2q_vo_update.log
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sub"
$B > gpurun_out/r2q_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_cfg2.csv $B > gpurun_out/r2q_n0.log 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vo_grid2_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2q_vo_grid2_cfg2_f64 $B > gpurun_out/r2q_n1.log 2>&1
C="python bench.py --workload cfg5 --steps 2"
$C > gpurun_out/r2q_cfg5_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2500 -c 1500 --csv --log-file gpurun_out/r2q_launches_cfg5.csv $C > gpurun_out/r2q_n2.log 2>&1
python profiles/tools/time_rom.py 16384 cfg3 > gpurun_out/r2q_rom_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rom_tpw_adjoint --launch-skip 2 -c 1 -f -o gpurun_out/r2q_rom_tpw_adjoint_cfg3 python profiles/tools/time_rom.py 16384 cfg3 > gpurun_out/r2q_n3.log 2>&1
tail -3 gpurun_out/r2q_rom_cfg3.log
ls -la gpurun_out/r2q_*
