mkdir -p gpurun_out/c11
( time python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert" | tail -30 ) > gpurun_out/c11/pytest.log 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c11/bench.json 2> gpurun_out/c11/bench.err
cat gpurun_out/c11/pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/c11/bench.json")); print(round(d["value"],1), d["stage_ms"], {k:(round(v["ms_per_step"],3), v.get("blend_fwd_ms"), v.get("blend_bwd_ms"), v.get("vs_reference")) for k,v in d.get("sh",{}).items()}, "e2e", round(d["e2e"]["value"],1), d["roofline"]["issue_frac"], d["roofline"]["thread_inst_per_pair"])
PY
