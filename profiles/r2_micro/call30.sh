#!/bin/bash
mkdir -p gpurun_out/c30
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/c30/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c30/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c30/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c30/smoke.log | cut -c1-100
