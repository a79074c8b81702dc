"""N>1 host logic on CPU: world_size=2, gloo backend, 127.0.0.1 rendezvous."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dp
    torch.manual_seed(0)                       # replicas start identical
    params = [torch.nn.Parameter(torch.randn(50, 3)), torch.nn.Parameter(torch.randn(50)),
              torch.nn.Parameter(torch.randn(50, 4))]
    # each rank "renders" its own view: a rank-dependent gradient
    views = [dp.view_for_rank(s, rank, world, 8) for s in range(4)]
    for p in params:
        p.grad = torch.full_like(p, float(rank + 1))
    params[1].grad = None                      # a parameter that got no gradient on this rank
    bucket = dp.GradBucket(params)
    bucket.allreduce()
    stat = torch.tensor([float(rank + 1)])
    dp.allreduce_stats([stat])
    ok = (bool((params[0].grad == 3.0).all()) and bool((params[2].grad == 3.0).all())
          and bool((params[1].grad == 0.0).all()) and float(stat) == 3.0)
    # async variant + averaging
    for p in params:
        p.grad = torch.full_like(p, float(rank))
    work, finish = dp.GradBucket(params, average=True).allreduce(async_op=True)
    work.wait()
    finish()
    ok = ok and bool((params[0].grad == 0.5).all())
    # zero-copy path: gradients that are views of one flat buffer (as the fused backward makes them)
    flat = torch.zeros(150 + 2 + 52 + 200)          # pads of 2 after the first segment
    gviews = [flat[0:150].view(50, 3), flat[152:202].view(50), flat[204:404].view(50, 4)]
    for p, v in zip(params, gviews):
        v.fill_(float(rank + 1))
        p.grad = v
    assert dp._as_one_buffer([p.grad for p in params]) is not None
    dp.GradBucket(params).allreduce()
    ok = ok and bool((flat[0:150] == 3.0).all()) and bool((params[2].grad == 3.0).all()) \
        and params[0].grad.untyped_storage().data_ptr() == flat.untyped_storage().data_ptr()
    # exchange selection: a gloo / CPU group always gets the portable bucket; forcing a peer-memory
    # mode there is an error, not a silent fallback
    ok = ok and type(dp.make_grad_bucket(params)) is dp.GradBucket \
        and type(dp.make_grad_bucket(params, exchange="nccl")) is dp.GradBucket
    try:
        dp.make_grad_bucket(params, exchange="p2p")
        ok = False
    except RuntimeError:
        pass
    out[rank] = (ok, views, bucket.nbytes())
    dist.destroy_process_group()


def test_view_sharding_is_a_partition():
    sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))
    import dp
    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in dp.shard_views(8, r, world))
        assert seen == list(range(8))
        step_views = sorted(dp.view_for_rank(3, r, world, 8) for r in range(world))
        assert len(set(step_views)) == world


def test_gradient_bucket_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world))
    assert out[0][1] == [0, 2, 4, 6] and out[1][1] == [1, 3, 5, 7]
    assert out[0][2] == (150 + 50 + 200) * 4


def test_symmetric_bucket_mode_resolution():
    """Mode selection of SymmetricGradBucket without GPUs: the symmetric-memory allocation and the
    exchange kernels are replaced by stand-ins; only the host logic (auto -> push at 2 / 4 / 8 ranks,
    optional NVSwitch broadcast, slice size, allocator contract) runs."""
    sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))
    import dp

    class Hdl:
        def __init__(self, t, mc):
            self.buffer_ptrs = [t.data_ptr(), t.data_ptr() + 1 << 20]
            self.multicast_ptr = mc

        def barrier(self, channel=0):
            pass

    def make(world, mode, mc=1234):
        b = dp.SymmetricGradBucket.__new__(dp.SymmetricGradBucket)
        b.params = [torch.nn.Parameter(torch.zeros(4))]
        b.average, b.group, b.world, b.rank = False, None, world, 0
        b.mode, b.buf, b.hdl, b.staging, b.staging_hdl, b.per = mode, None, None, None, None, 0
        b._alloc = lambda numel, device: (lambda t: (t, Hdl(t, mc)))(torch.zeros(numel))
        b._reduce = lambda buf, hdl, numel: buf.fill_(world * (world + 1) / 2)
        b._probe()
        return b

    assert make(2, "auto").mode == "push" and not make(2, "auto").push_mc
    assert make(4, "auto").mode == "push" and not make(4, "auto").push_mc      # plain peer stores unless GS_DP_PUSH_MC=1
    os.environ["GS_DP_PUSH_MC"] = "1"
    try:
        assert make(4, "auto").push_mc and not make(4, "auto", mc=0).push_mc
    finally:
        del os.environ["GS_DP_PUSH_MC"]
    assert make(8, "auto", mc=0).mode == "push" and not make(8, "auto", mc=0).push_mc
    assert make(2, "p2p").mode == "p2p" and make(4, "push").mode == "push"
    with pytest.raises(RuntimeError):
        make(4, "multimem", mc=0)
    with pytest.raises(RuntimeError):
        make(3, "p2p")
    b = make(2, "push")
    out = b.allocator(4 * 10 + 2, torch.device("cpu"))       # rounded up to 44 floats
    assert isinstance(out, tuple) and out[0].numel() == 44
    bucket_ptr, staging_ptrs, per, rank = out[1]
    assert per % 4 == 0 and per * 2 >= 44 and b.staging.numel() == 2 * per and len(staging_ptrs) == 2 and rank == 0
    assert bool((b.staging == 0).all())
    assert b.allocator(42, torch.device("cpu"))[0] is out[0]  # persistent across steps
    p = make(2, "p2p")
    assert isinstance(p.allocator(40, torch.device("cpu")), torch.Tensor)
