#!/bin/bash
mkdir -p gpurun_out/c23
timeout 200 python profiles/r2_micro/sweep_sh.py 27 C3 0,3 > gpurun_out/c23/c3_27.txt 2>&1; echo "c3_27 rc=$?"; tail -2 gpurun_out/c23/c3_27.txt
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c23/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c23/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c23/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c23/smoke.log
