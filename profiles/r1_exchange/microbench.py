"""Exchange-only timing of the 134 MB gradient bucket (2.4 M Gaussians x 14 floats):
own p2p / multimem kernels (with their two symmetric-memory barriers) vs NCCL all-reduce.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/r1_exchange/microbench.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "3d-gaussian-splatting_b200"))
rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

import gaussian  # noqa: E402

n = 2400000 * 14
buf = symm_mem.empty(n, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, group=dist.group.WORLD.group_name)
plain = torch.empty(n, dtype=torch.float32, device=dev)
mc, ptrs = int(hdl.multicast_ptr), [int(x) for x in hdl.buffer_ptrs]


def timed(fn, it=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def multimem():
    hdl.barrier(channel=0)
    gaussian.allreduce_multimem(mc, n, rank, world, local)
    hdl.barrier(channel=1)


def p2p():
    hdl.barrier(channel=0)
    gaussian.allreduce_p2p(ptrs, n, rank, world, local)
    hdl.barrier(channel=1)


torch.manual_seed(rank)
x = torch.randn(n, device=dev)
want = x.clone()
dist.all_reduce(want)
errs = {}
for name, fn in (("p2p", p2p), ("multimem", multimem)):
    buf.copy_(x)
    fn()
    torch.cuda.synchronize()
    errs[name] = float((buf - want).abs().max())
buf.zero_()
plain.zero_()
r = dict(p2p=timed(p2p), p2p_kernel=timed(lambda: gaussian.allreduce_p2p(ptrs, n, rank, world, local)),
         multimem=timed(multimem), multimem_kernel=timed(lambda: gaussian.allreduce_multimem(mc, n, rank, world, local)),
         nccl=timed(lambda: dist.all_reduce(plain)), barrier=timed(lambda: hdl.barrier(channel=0)))
if rank == 0:
    print("max|err| vs NCCL", errs, "ms", {k: round(v, 4) for k, v in r.items()}, flush=True)
dist.destroy_process_group()
