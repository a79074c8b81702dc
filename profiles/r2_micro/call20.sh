#!/bin/bash
mkdir -p gpurun_out/c20
( cd profiles/r2_micro; for t in 8 9; do timeout 60 ./umma_probe_elect $t 0 > ../../gpurun_out/c20/probe_$t.txt 2>&1; echo "probe $t rc=$?"; done )
grep -v "^  row" gpurun_out/c20/probe_8.txt | tail -3; awk '/row/{print $2, $4}' gpurun_out/c20/probe_8.txt | tr '\n' ';' | head -c 700; echo; cat gpurun_out/c20/probe_9.txt
timeout 120 python profiles/r2_micro/sweep_sh.py 27 S 0,3 > gpurun_out/c20/small27.txt 2>&1; echo "small27 rc=$?"; tail -2 gpurun_out/c20/small27.txt
GS_TUNE_SH_TC=3 timeout 300 python -m pytest tests/test_frame_gpu.py tests/test_scale_parity_gpu.py -q -m gpu -k "sh or masked" -x > gpurun_out/c20/pytest_sh.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c20/pytest_sh.log
timeout 200 python profiles/r2_micro/sweep_sh.py 27 C3 0,3 > gpurun_out/c20/c3_27.txt 2>&1; echo "c3_27 rc=$?"; tail -2 gpurun_out/c20/c3_27.txt
timeout 200 python profiles/r2_micro/sweep_sh.py 48 C3 3 > gpurun_out/c20/c3_48.txt 2>&1; echo "c3_48 rc=$?"; tail -1 gpurun_out/c20/c3_48.txt
