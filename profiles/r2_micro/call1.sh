mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1/smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/c1/pytest.log 2>&1
./profiles/r2_micro/ffma2_bench > gpurun_out/c1/ffma2.txt 2>&1
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c1/bench_base.json 2> gpurun_out/c1/bench_base.err
GS_BWD_WARPS=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c1/bench_w1.json 2> gpurun_out/c1/bench_w1.err
tail -3 gpurun_out/c1/pytest.log; cat gpurun_out/c1/ffma2.txt
