mkdir -p gpurun_out/c2
( time python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -40 ) > gpurun_out/c2/pytest.log 2>&1
for v in "GS_BLEND_V=2 GS_BWD_PX=4" "GS_BLEND_V=2 GS_BWD_PX=8" "GS_BLEND_V=1"; do
  tag=$(echo $v | tr ' =' '__')
  env $v python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c2/bench_$tag.json 2> gpurun_out/c2/bench_$tag.err
done
( GS_BWD_PX=8 timeout 600 python -m pytest tests/test_frame_gpu.py tests/test_stages_gpu.py tests/test_scale_parity_gpu.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/c2/pytest_px8.log 2>&1
( timeout 400 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_frame_gpu.py -k "test_fused_frame_vs_oracle" -x -q 2>&1 | tail -15 ) > gpurun_out/c2/racecheck.log 2>&1
tail -4 gpurun_out/c2/pytest.log; tail -3 gpurun_out/c2/pytest_px8.log; tail -6 gpurun_out/c2/racecheck.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/c2/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), d["stage_ms"])
    except Exception as e: print(f,"ERR",e)
PY
