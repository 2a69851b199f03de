# Round-2 ncu captures (one GPU; each only after the same command has exited 0 without ncu).  Output: gpurun_out/r2_*.ncu-rep
set -x
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e --no-sub"
NCU="ncu --set full --clock-control none --import-source on"
$B > gpurun_out/r2p_plain.log 2>&1 && $NCU -k regex:vo_grid2_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_vo_grid2_cfg2_f64 $B > gpurun_out/r2p_ncu1.log 2>&1
$B --dtype f32 > gpurun_out/r2p_plain_f32.log 2>&1 && $NCU -k regex:vo_grid2_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_vo_grid2_cfg2_f32 $B --dtype f32 > gpurun_out/r2p_ncu2.log 2>&1
python profiles/tools/time_rom.py 32768 cfg2 > gpurun_out/r2p_rom_b32768.log 2>&1 && $NCU -k regex:rom_ --launch-skip 6 -c 2 -f -o gpurun_out/r2_rom_tps_b32768_f64 python profiles/tools/time_rom.py 32768 cfg2 > gpurun_out/r2p_ncu3.log 2>&1
python profiles/tools/time_rom.py 4096 cfg3 > gpurun_out/r2p_rom_cfg3.log 2>&1 && $NCU -k regex:rom_ --launch-skip 6 -c 2 -f -o gpurun_out/r2_rom_coop_cfg3_b4096_f64 python profiles/tools/time_rom.py 4096 cfg3 > gpurun_out/r2p_ncu4.log 2>&1
cat gpurun_out/r2p_rom_b32768.log gpurun_out/r2p_rom_cfg3.log
ls -la gpurun_out/*.ncu-rep
