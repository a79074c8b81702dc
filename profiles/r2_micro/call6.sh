mkdir -p gpurun_out/c6
python profiles/r2_micro/sweep.py C3 > gpurun_out/c6/sweep.txt 2>&1
( python -m pytest tests/test_frame_gpu.py tests/test_stages_gpu.py tests/test_scale_parity_gpu.py -m gpu -q 2>&1 | tail -4 ) > gpurun_out/c6/pytest.log 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c6/bench_gather.json 2> gpurun_out/c6/bench_gather.err
GS_TUNE_GATHER=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c6/bench_packed.json 2> gpurun_out/c6/bench_packed.err
cat gpurun_out/c6/sweep.txt; cat gpurun_out/c6/pytest.log
python - <<'PY'
import json
for f in ("gather","packed"):
    d=json.load(open(f"gpurun_out/c6/bench_{f}.json")); print(f, round(d["value"],1), d["stage_ms"], {k:(round(v["ms_per_step"],3), v.get("pack_ms"), v.get("blend_fwd_ms"), v.get("blend_bwd_ms")) for k,v in d.get("sh",{}).items()}, "e2e", round(d["e2e"]["value"],1))
PY
