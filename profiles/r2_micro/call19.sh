#!/bin/bash
mkdir -p gpurun_out/c19
GS_TUNE_SH_TC=3 ncu --set full --clock-control none --import-source on -k regex:"blend_sh" -s 2 -c 2 -o gpurun_out/c19/prof_tc_D27 -f python bench.py --colour 27 --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/c19/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c19/ncu.log; ls -la gpurun_out/c19
