#!/bin/bash
# tensor-core SH kernels as the default: full GPU suite, SH sweeps, SH bench legs
mkdir -p gpurun_out/c21
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c21/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c21/pytest.log
timeout 200 python profiles/r2_micro/sweep_sh.py 48 C3 0,3 > gpurun_out/c21/c3_48.txt 2>&1; echo "c3_48 rc=$?"; tail -2 gpurun_out/c21/c3_48.txt
timeout 200 python profiles/r2_micro/sweep_sh.py 27 C3 0,3 > gpurun_out/c21/c3_27.txt 2>&1; echo "c3_27 rc=$?"; tail -2 gpurun_out/c21/c3_27.txt
python bench.py --colour 27 --no-cpu-baseline --steps 30 > gpurun_out/c21/bench_sh27.json 2>gpurun_out/c21/bench_sh27.err; echo "bench27 rc=$?"
python bench.py --colour 48 --no-cpu-baseline --steps 30 > gpurun_out/c21/bench_sh48.json 2>gpurun_out/c21/bench_sh48.err; echo "bench48 rc=$?"
python - <<'PY'
import json
for d in (27,48):
    try:
        b=json.loads(open(f'gpurun_out/c21/bench_sh{d}.json').read().strip().splitlines()[-1])
        print(d, b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], b.get('stage_ms'))
    except Exception as e: print(d,'ERR',e)
PY
