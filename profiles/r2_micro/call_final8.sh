N=8
mkdir -p gpurun_out/f8
( GS_TEST_TRAIN_WORLD=8 timeout 900 python -m pytest tests/test_reference_train_gpu.py -m gpu -q -k data_parallel 2>&1 | tail -6 ) > gpurun_out/f8/pytest_train_dp8.log 2>&1
run() { tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 8 "$@" > gpurun_out/f8/bench_$tag.json 2> gpurun_out/f8/bench_$tag.err
}
run push
run nccl --exchange nccl
run c5 --workload C5 --steps 30
tail -3 gpurun_out/f8/pytest_train_dp8.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/f8/bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d.get("exchange_check"))
    except Exception as e: print(f,"ERR",e)
PY
