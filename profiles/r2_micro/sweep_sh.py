"""A/B of the SH blend kernels on the C3 scene: scalar (blend_sh.cu) vs tensor-core (blend_sh_tc.cu) forward /
backward (knob sh_tc: bit 0 forward, bit 1 backward).  Prints stage times (CUDA events recorded by the library,
mean of 6 frames) and every variant's deviation from the scalar kernels.
Usage: python profiles/r2_micro/sweep_sh.py [27|48] [C3|C2|S]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "3d-gaussian-splatting_b200"))
import torch  # noqa: E402
import gaussian  # noqa: E402
import splatter  # noqa: E402
import synthetic as S  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 27
wl = sys.argv[2] if len(sys.argv) > 2 else "C3"
variants = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 3]
n, w, h = {"C3": (2_400_000, 1920, 1080), "C2": (500_000, 1920, 1080), "S": (20_000, 320, 240)}[wl]
dev = torch.device("cuda", 0)
g = S.make_gaussians(n, w, h, 0, sh_dim=d)
v = S.make_view(w, h, 0)
sp = splatter.Splatter.from_tensors(g, [dict(width=w, height=h, focal_x=v.fx, focal_y=v.fy, rot=v.rot, tran=v.tran)],
                                    device=dev, use_sh_coeff=True)
sp._rctx.set_timing(True)
go = (S.make_grad_output(h, w, 0) * (h * w)).to(dev)
params = list(sp.gaussian_3ds.parameters())


def run(tc, frames=6):
    gaussian.tune("sh_tc", tc)
    f = b = tot = 0.0
    for it in range(2 + frames):
        for p in params:
            p.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        img = sp(0)
        img.backward(go)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            st = sp._rctx.stage_ms()
            f += st[5] / frames
            b += st[6] / frames
            tot += e0.elapsed_time(e1) / frames
    return f, b, tot, img.detach().clone(), [p.grad.clone() for p in params], sp.frame_stats()


ref = None
rows = []
for tc in variants:
    try:
        f, b, tot, img, grads, st = run(tc)
    except Exception as e:
        print(f"sh_tc={tc} FAILED: {str(e)[:200]}", flush=True)
        break
    if ref is None:
        ref = (img, grads)
    ierr = float((img - ref[0]).abs().max())
    gerr = [float((a - r).abs().max() / (r.abs().max() + 1e-30)) for a, r in zip(grads, ref[1])]
    rows.append(dict(sh_tc=tc, d=d, blend_fwd_ms=round(f, 4), blend_bwd_ms=round(b, 4), frame_ms=round(tot, 4), img_abs_dev=ierr,
                     grad_rel_dev=gerr, m_eff=st.get("n_instances_eff"), m_eff_bwd=st.get("n_instances_eff_bwd")))
    print(f"D={d} {wl} sh_tc={tc}  fwd {f:.4f}  bwd {b:.4f}  frame {tot:.4f}  img dev {ierr:.2e}  grad dev {max(gerr):.2e} "
          f"M_eff {st.get('n_instances_eff')} / {st.get('n_instances_eff_bwd')}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"sweep_sh_{d}_{wl}.json"), "w"), indent=1)
