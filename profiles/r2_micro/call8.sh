mkdir -p gpurun_out/c8
python profiles/r2_micro/sweep.py C3 > gpurun_out/c8/sweep.txt 2>&1
( python -m pytest tests/test_frame_gpu.py tests/test_stages_gpu.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/c8/pytest.log 2>&1
( GS_TUNE_BWD_CH=32 python -m pytest tests/test_frame_gpu.py tests/test_scale_parity_gpu.py -m gpu -q 2>&1 | tail -3 ) > gpurun_out/c8/pytest_ch32.log 2>&1
cat gpurun_out/c8/sweep.txt; cat gpurun_out/c8/pytest.log gpurun_out/c8/pytest_ch32.log
