# usage: bash profiles/r2_micro/call_multi.sh N
N=$1
mkdir -p gpurun_out/m$N
nvidia-smi -L > gpurun_out/m$N/gpus.txt
( GS_TEST_EXCHANGE_WORLD=$N timeout 600 python -m pytest tests/test_exchange_gpu.py -m gpu -q 2>&1 | tail -15 ) > gpurun_out/m$N/pytest_exchange.log 2>&1
if [ "$N" = "2" ]; then
  ( timeout 900 python -m pytest tests/test_reference_train_gpu.py -m gpu -q -k data_parallel 2>&1 | tail -15 ) > gpurun_out/m$N/pytest_train_dp.log 2>&1
fi
run() { # tag, extra args
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 8 "$@" > gpurun_out/m$N/bench_$tag.json 2> gpurun_out/m$N/bench_$tag.err
}
run push
GS_DP_PUSH_MC=0 run push_nomc
GS_DP_PUSH_MC=1 run push_mc
run nccl --exchange nccl
run c5 --workload C5 --steps 30
if [ "$N" != "8" ]; then python bench.py --gpus 1 --steps 50 --warmup 8 --no-cpu-baseline --no-extra-legs > gpurun_out/m$N/bench_n1.json 2> gpurun_out/m$N/bench_n1.err; fi
tail -3 gpurun_out/m$N/pytest_exchange.log; [ -f gpurun_out/m$N/pytest_train_dp.log ] && tail -3 gpurun_out/m$N/pytest_train_dp.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/m$N/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("exchange_check"), [ (r["view"], r["compute_ms"], r["exchange_ms"]) for r in d.get("per_rank",[])])
    except Exception as e: print(f,"ERR",e)
PY
