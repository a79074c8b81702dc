mkdir -p gpurun_out/c13
( python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -8 ) > gpurun_out/c13/pytest.log 2>&1
python __graft_entry__.py smoke > gpurun_out/c13/smoke.log 2>&1
python bench.py > gpurun_out/c13/bench.json 2> gpurun_out/c13/bench.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/c13/bench_ref.json 2> gpurun_out/c13/bench_ref.err
cat gpurun_out/c13/pytest.log; tail -1 gpurun_out/c13/smoke.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/c13/bench.json")); r=json.load(open("gpurun_out/c13/bench_ref.json"))
print(round(d["value"],1), round(d["e2e"]["value"],1), "ref", round(r["value"],2), "ratio", round(d["e2e"]["value"]/r["e2e"]["value"],2), d["config"]["workload"]==r["config"]["workload"], d["gpu_launches"], d["roofline"]["issue_frac"], d["sh"]["27"].get("vs_reference"))
PY
