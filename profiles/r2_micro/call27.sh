#!/bin/bash
mkdir -p gpurun_out/c27
timeout 120 python profiles/r2_micro/sweep_sh.py 27 S 0,3,7 > gpurun_out/c27/small27.txt 2>&1; echo "small27 rc=$?"; tail -3 gpurun_out/c27/small27.txt
timeout 120 python profiles/r2_micro/sweep_sh.py 48 S 0,7 > gpurun_out/c27/small48.txt 2>&1; echo "small48 rc=$?"; tail -2 gpurun_out/c27/small48.txt
GS_TUNE_SH_TC=7 timeout 400 python -m pytest tests/test_frame_gpu.py tests/test_scale_parity_gpu.py -q -m gpu -k "sh or masked" -x > gpurun_out/c27/pytest_sh.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c27/pytest_sh.log
timeout 200 python profiles/r2_micro/sweep_sh.py 27 C3 3,7 > gpurun_out/c27/c3_27.txt 2>&1; echo "c3_27 rc=$?"; tail -2 gpurun_out/c27/c3_27.txt
timeout 200 python profiles/r2_micro/sweep_sh.py 48 C3 3,7 > gpurun_out/c27/c3_48.txt 2>&1; echo "c3_48 rc=$?"; tail -2 gpurun_out/c27/c3_48.txt
