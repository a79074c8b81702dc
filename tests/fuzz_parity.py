"""Randomised parity sweep (not collected by pytest; run on the GPU box:
`python tests/fuzz_parity.py [n_trials]`).  Random sizes / views / opacity regimes / RGB or
SH-27 colour; fused CUDA frame vs the fp64 CPU oracle sorted on the device's depth keys.
Round-1 result (24 trials): 0 mismatches, worst image error 3.2e-6 (budget 1e-4), worst
gradient error 1.4e-5 (budget 1e-3), instance counts identical in every trial."""
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("3d-gaussian-splatting_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import gs_oracle as O  # noqa: E402
import splatter  # noqa: E402
import synthetic as S  # noqa: E402
from helpers import abs_err, device_depth_keys, rel_err, scene  # noqa: E402


def main(trials=24, seed=1):
    dev = torch.device("cuda", 0)
    random.seed(seed)
    worst_img = worst_g = 0.0
    bad = 0
    for t in range(trials):
        n = random.choice([500, 2000, 6000])
        w, h = random.choice([64, 112, 200, 256]), random.choice([48, 80, 120, 192])
        k, sd = random.randrange(8), random.randrange(1000)
        opa = random.choice([(0.005, 0.05), (0.05, 0.9), (0.5, 0.98)])
        sh = random.choice([3, 3, 27])
        g, v, cam = scene(n, w, h, seed=sd, k=k, sh_dim=sh, opa_range=opa)
        go = S.make_grad_output(h, w, sd) * (h * w)
        dk = device_depth_keys(g, cam, dev)
        p = {kk: x.double().clone().requires_grad_(True) for kk, x in g.items()}
        oimg, aux = O.render(p["pos"], p["rgb"], p["opa"], p["quat"], p["scale"], cam, use_sh_coeff=sh != 3,
                             return_aux=True, depth_key=dk)
        oimg.backward(go.double())
        vs = [dict(width=v.width, height=v.height, focal_x=v.fx, focal_y=v.fy, rot=v.rot, tran=v.tran)]
        sp = splatter.Splatter.from_tensors(g, vs, device=dev, use_sh_coeff=sh != 3)
        img = sp(0)
        img.backward(go.to(dev))
        m, mo = sp.frame_stats()["n_instances"], int(aux["accum"][-1])
        ei = abs_err(img, oimg)
        eg = max(rel_err(getattr(sp.gaussian_3ds, kk).grad, p[kk].grad) for kk in g)
        ok = m == mo and ei < 1e-4 and eg < 1e-3
        bad += not ok
        worst_img, worst_g = max(worst_img, ei), max(worst_g, eg)
        print(t, n, w, h, k, sd, opa, sh, "M", m, mo, f"img {ei:.2e} grad {eg:.2e}", "" if ok else "  <-- MISMATCH")
    print("worst img", worst_img, "worst grad", worst_g, "mismatches", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 24) else 0)
