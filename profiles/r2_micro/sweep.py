"""A/B sweep of the blend-kernel variants on the C3 scene (one process, one scene; knobs via gaussian.tune).
Prints per-variant blend forward / backward stage times (mean of 8 frames, CUDA events recorded by the
library) and checks every variant's gradients against the first one.  Usage: python profiles/r2_micro/sweep.py [C3|C2]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))
import torch  # noqa: E402
import gaussian  # noqa: E402
import splatter  # noqa: E402
import synthetic as S  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
n, w, h = {"C3": (2_400_000, 1920, 1080), "C2": (500_000, 1920, 1080)}[wl]
dev = torch.device("cuda", 0)
g = S.make_gaussians(n, w, h, 0)
v = S.make_view(w, h, 0)
sp = splatter.Splatter.from_tensors(g, [dict(width=w, height=h, focal_x=v.fx, focal_y=v.fy, rot=v.rot, tran=v.tran)], device=dev)
sp._rctx.set_timing(True)
go = S.make_grad_output(h, w, 0).to(dev)
params = list(sp.gaussian_3ds.parameters())

BASE = dict(fwd_kernel=0, fwd_ch=128, fwd_px=4, bwd_kernel=1, bwd_px=8, bwd_ws=0, bwd_unroll=2, bwd_stages=2, bwd_minb=16, bwd_rq=4,
            bwd_ch=64, gather=1, strict=1)
VARIANTS = [
    ("shipped: gather, fwd ch128, bwd2 px8 unroll2 st2 minb16 ch64", dict()),
    ("bwd ch32 st2", dict(bwd_ch=32)),
    ("bwd ch32 st3", dict(bwd_ch=32, bwd_stages=3)),
    ("bwd ch32 st2 unroll4 minb10", dict(bwd_ch=32, bwd_unroll=4, bwd_minb=10)),
    ("bwd ch32 st3 unroll4 minb10", dict(bwd_ch=32, bwd_unroll=4, bwd_minb=10, bwd_stages=3)),
    ("bwd ch64 unroll4 minb10", dict(bwd_unroll=4, bwd_minb=10)),
    ("packed: r1 kernels", dict(gather=0, fwd_ch=256, bwd_kernel=0, bwd_px=4)),
]


def run(cfg, frames=8):
    for k, val in {**BASE, **cfg}.items():
        gaussian.tune(k, val)
    f = b = tot = pk = 0.0
    for it in range(3 + frames):
        for p in params:
            p.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img = sp(0)
        img.backward(go)
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            st = sp._rctx.stage_ms()
            f += st[5] / frames
            b += st[6] / frames
            pk += st[4] / frames
            tot += e0.elapsed_time(e1) / frames
    return f, b, tot, img.detach().clone(), [p.grad.clone() for p in params], pk


ref = None
rows = []
for name, cfg in VARIANTS:
    try:
        f, b, tot, img, grads, pk = run(cfg)
    except Exception as e:
        print(f"{name:48s} FAILED: {str(e)[:120]}", flush=True)
        continue
    if ref is None:
        ref = (img, grads)
        err = 0.0
    else:
        err = max(float((img - ref[0]).abs().max()),
                  max(float((a - r).abs().max() / (r.abs().max() + 1e-30)) for a, r in zip(grads, ref[1])))
    rows.append(dict(name=name, cfg=cfg, blend_fwd_ms=round(f, 4), blend_bwd_ms=round(b, 4), frame_ms=round(tot, 4), max_dev_vs_first=err))
    print(f"{name:56s} pack/ranges {pk:.4f}  fwd {f:.4f}  bwd {b:.4f}  frame {tot:.4f}  dev {err:.1e}", flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"sweep_{wl}.json"), "w"), indent=1)
