mkdir -p gpurun_out/c4
python profiles/r2_micro/sweep.py C3 > gpurun_out/c4/sweep.txt 2>&1
( python -m pytest tests/test_loss_gpu.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/c4/pytest_loss.log 2>&1
cat gpurun_out/c4/sweep.txt; cat gpurun_out/c4/pytest_loss.log
