#!/bin/bash
mkdir -p gpurun_out/c29
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_frame_gpu.py -q -m gpu -k "sh_vs_oracle and tensor" > gpurun_out/c29/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/c29/memcheck.log | tail -3
GS_TUNE_SH_TC=7 timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_scale_parity_gpu.py tests/test_frame_gpu.py -q -m gpu -k "densification_kernel_vs_oracle and 27" > gpurun_out/c29/memcheck_tc2.log 2>&1; echo "memcheck tc2 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/c29/memcheck_tc2.log | tail -3
timeout 400 python -m pytest tests/test_frame_gpu.py tests/test_scale_parity_gpu.py -q -m gpu -k "sh or masked" > gpurun_out/c29/pytest_sh.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c29/pytest_sh.log
timeout 200 python profiles/r2_micro/sweep_sh.py 27 C3 0,3,7 > gpurun_out/c29/c3_27.txt 2>&1; echo "c3_27 rc=$?"; tail -3 gpurun_out/c29/c3_27.txt
