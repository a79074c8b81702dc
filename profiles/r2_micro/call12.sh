mkdir -p gpurun_out/c12
( python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED" | tail -5 ) > gpurun_out/c12/pytest.log 2>&1
python __graft_entry__.py smoke > gpurun_out/c12/smoke.log 2>&1
bash profiles/r2_profile_commands.sh > gpurun_out/c12/profile.log 2>&1
cat gpurun_out/c12/pytest.log; tail -2 gpurun_out/c12/smoke.log
