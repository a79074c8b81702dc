#!/bin/bash
# final single-GPU verification: whole GPU suite, smoke, default bench + reference arm, sanitizer on the tensor-core SH kernels
mkdir -p gpurun_out/c28
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/c28/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/c28/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c28/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c28/smoke.log | cut -c1-120
python bench.py > gpurun_out/c28/bench.json 2> gpurun_out/c28/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/c28/bench_ref.json 2> gpurun_out/c28/bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/c28/bench.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/c28/bench_ref.json').read().strip().splitlines()[-1])
print(b['value'], b['e2e']['value'], 'ref', r['value'], 'ratio', round(b['e2e']['value']/r['value'],2), b['config']['workload']==r['config']['workload'], b.get('gpu_launches'), 'sh', b.get('sh'))
PY
timeout 240 compute-sanitizer --tool memcheck python -m pytest tests/test_frame_gpu.py -q -m gpu -k "sh_vs_oracle and tensor" > gpurun_out/c28/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/c28/memcheck.log | tail -3
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_frame_gpu.py -q -m gpu -k "sh_vs_oracle and tensor and 27-opa1" > gpurun_out/c28/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|hazard" gpurun_out/c28/racecheck.log | tail -5
