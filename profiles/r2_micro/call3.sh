mkdir -p gpurun_out/c3
( time python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | grep -E "P4@C2|passed|failed|FAILED|Error|error|assert" | tail -40 ) > gpurun_out/c3/pytest.log 2>&1
python profiles/r2_micro/sweep.py C3 > gpurun_out/c3/sweep.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 6 -c 2 -o gpurun_out/c3/prof_ws -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/c3/ncu_ws.log 2>&1
GS_TUNE_BWD_WS=0 GS_TUNE_BWD_RQ=8 GS_TUNE_BWD_STAGES=2 GS_TUNE_BWD_MINB=10 GS_TUNE_FWD_KERNEL=0 ncu --set full --clock-control none --import-source on -k regex:"blend_" -s 6 -c 2 -o gpurun_out/c3/prof_t0 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/c3/ncu_t0.log 2>&1
python bench.py --steps 30 --warmup 5 > gpurun_out/c3/bench_default.json 2> gpurun_out/c3/bench_default.err
cat gpurun_out/c3/pytest.log | tail -30; cat gpurun_out/c3/sweep.txt
