#!/bin/bash
mkdir -p gpurun_out/c25
timeout 400 python -m pytest tests/test_frame_gpu.py tests/test_scale_parity_gpu.py tests/test_stages_gpu.py -q -m gpu -x > gpurun_out/c25/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c25/pytest.log
python bench.py --colour 27 --no-cpu-baseline --steps 30 > gpurun_out/r2/bench_ours_C3_sh27.json 2>/dev/null; echo "bench27 rc=$?"
python bench.py --colour 48 --no-cpu-baseline --steps 30 > gpurun_out/r2/bench_ours_C3_sh48.json 2>/dev/null; echo "bench48 rc=$?"
python - <<'PY'
import json
for d in (27,48):
    b=json.loads(open(f'gpurun_out/r2/bench_ours_C3_sh{d}.json').read().strip().splitlines()[-1])
    print(d, b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], b.get('stage_ms'))
PY
