"""Two-GPU check of the peer-memory (p2p and NVLS multimem) gradient exchange against NCCL's all-reduce of the same
buckets.  Needs >= 2 GPUs of one NVSwitch domain; skipped elsewhere."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", device_id=dev)
        import dp
        import renderer
        import splatter
        import synthetic as S

        n, w, h = 20001, 256, 160                       # N % 4 != 0: exercises the padded segments
        g = S.make_gaussians(n, w, h, 0)
        views = [S.make_view(w, h, k % 8) for k in range(world)]
        vd = [dict(width=v.width, height=v.height, focal_x=v.fx, focal_y=v.fy, rot=v.rot, tran=v.tran) for v in views]
        go = (S.make_grad_output(h, w, 0) * (h * w)).to(dev)

        def grads(bucket):
            sp = splatter.Splatter.from_tensors(g, vd, device=dev)
            params = list(sp.gaussian_3ds.parameters())
            b = bucket(params)
            out = []
            for _ in range(2):                           # second pass re-uses the persistent bucket
                for p in params:
                    p.grad = None
                sp(rank).backward(go)
                b.allreduce()
                out = [p.grad.clone() for p in params]
            return out, b

        ref, _ = grads(lambda ps: dp.GradBucket(ps))
        worst, same = 0.0, True
        for mode in ("p2p", "multimem", "push", "auto"):
            got, b = grads(lambda ps: dp.make_grad_bucket(ps, exchange=mode))
            assert isinstance(b, dp.SymmetricGradBucket) and b.mode == ("push" if mode == "auto" else mode), b.mode
            renderer.set_flat_grad_allocator(None)
            for a, r in zip(got, ref):
                worst = max(worst, float((a - r).abs().max() / (r.abs().max() + 1e-30)))
            # every rank must hold the same bits after the exchange
            flat = torch.cat([a.flatten() for a in got])
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            same = same and all(bool((gathered[0] == t).all()) for t in gathered)
        dist.destroy_process_group()
        q.put((rank, worst, same, None))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, None, None, traceback.format_exc() + str(e)))


@pytest.mark.timeout(300)
def test_peer_memory_exchange_matches_nccl():
    world = int(os.environ.get("GS_TEST_EXCHANGE_WORLD", "2"))      # 2 (default), 4 or 8
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(30)
    for rank, worst, same, err in res:
        assert err is None, f"rank {rank}: {err}"
        assert worst < 1e-5, f"rank {rank}: peer-memory vs NCCL sum differs by {worst}"   # summation order only
        assert same, "ranks disagree after the exchange"
