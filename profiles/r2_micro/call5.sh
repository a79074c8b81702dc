mkdir -p gpurun_out/c5
( time python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error|error|assert" | tail -30 ) > gpurun_out/c5/pytest.log 2>&1
python profiles/r2_micro/sweep.py C3 > gpurun_out/c5/sweep.txt 2>&1
( timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_frame_gpu.py -k "test_fused_frame_vs_oracle" -x -q 2>&1 | tail -6 ) > gpurun_out/c5/racecheck.log 2>&1
( timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_frame_gpu.py -k "test_fused_frame_vs_oracle or test_edge_cases" -x -q 2>&1 | tail -6 ) > gpurun_out/c5/memcheck.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/c5/bench_default.json 2> gpurun_out/c5/bench_default.err
cat gpurun_out/c5/pytest.log; cat gpurun_out/c5/sweep.txt; tail -3 gpurun_out/c5/racecheck.log; tail -3 gpurun_out/c5/memcheck.log; tail -c 1500 gpurun_out/c5/bench_default.err
