"""Checkpoints.

`save_checkpoint` writes the reference's key schema (train.py:283-291: a dict with the five parameter
tensors `pos, opa, rgb, quat, scale`, so the reference's own `--ckpt` / `Splatter(load_ckpt=...)` path,
splatter.py:417-424, loads our files and we load theirs) and - what the reference lacks - everything a
true resume needs under the extra key `"resume"`: optimizer moments and step count, iteration, the
densification statistics of train.py:82-83, and the RNG states (numpy picks the camera, train.py:93;
torch samples split positions, utils.py:391-402).  Extra keys are ignored by the reference's loader.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

KEYS = ("pos", "opa", "rgb", "quat", "scale")          # train.py:284-290 order


def _optimizer_state(opt):
    if opt is None:
        return None
    if hasattr(opt, "state_dict"):
        return {"kind": "torch", "state": opt.state_dict()}
    # optim.FlatAdam: flat moments + per-group learning rates
    flat = getattr(opt, "_flat", None)
    return {"kind": "flat_adam", "step": opt.step_count, "betas": opt.betas, "eps": opt.eps,
            "lrs": [g["lr"] for g in opt.param_groups],
            "exp_avg": None if flat is None else flat[1].detach().cpu(),
            "exp_avg_sq": None if flat is None else flat[2].detach().cpu()}


def save_checkpoint(splatter, path, optimizer=None, iteration: Optional[int] = None, trainer_state: Optional[dict] = None):
    """`path` is the checkpoint file (train.py writes `<exp>/ckpt.pth`)."""
    g = splatter.gaussian_3ds
    ckpt = {k: getattr(g, k).detach().clone() for k in KEYS}
    ckpt["resume"] = {
        "iteration": iteration,
        "optimizer": _optimizer_state(optimizer),
        "trainer": {k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in (trainer_state or {}).items()},
        "torch_rng": torch.get_rng_state(),
        "cuda_rng": torch.cuda.get_rng_state(splatter.device) if torch.cuda.is_available() else None,
        "numpy_rng": np.random.get_state(),
        "scale_activation": splatter.scale_activation,
        "use_sh_coeff": splatter.use_sh_coeff,
    }
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp"
    torch.save(ckpt, tmp)
    os.replace(tmp, path)                                   # a crash never leaves a truncated checkpoint
    return path


def load_checkpoint(path, splatter=None, optimizer=None, restore_rng=True):
    """Returns the dict; with `splatter` the parameters are replaced in place (new nn.Parameters, like
    adaptive_control - rebuild torch optimizers before passing them here); with `optimizer` its state is restored.
    A reference-written file (five keys only) loads the parameters and returns `resume == None`."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    missing = [k for k in KEYS if k not in ckpt]
    if missing:
        raise KeyError(f"{path}: not a checkpoint of this family, missing {missing}")
    res = ckpt.get("resume")
    if splatter is not None:
        g = splatter.gaussian_3ds
        with torch.no_grad():
            for k in KEYS:
                t = ckpt[k].detach().to(device=splatter.device, dtype=torch.float32).contiguous()
                setattr(g, k, torch.nn.Parameter(t))
        splatter.n_gaussians = g.pos.shape[0]
    if optimizer is not None and res is not None and res.get("optimizer") is not None:
        st = res["optimizer"]
        if st["kind"] == "torch":
            optimizer.load_state_dict(st["state"])
        else:
            optimizer.step_count = st["step"]
            for grp, lr in zip(optimizer.param_groups, st["lrs"]):
                grp["lr"] = lr
            optimizer._resume = (st["exp_avg"], st["exp_avg_sq"])      # consumed by FlatAdam._build
    if restore_rng and res is not None:
        torch.set_rng_state(res["torch_rng"])
        if res.get("cuda_rng") is not None and splatter is not None:
            torch.cuda.set_rng_state(res["cuda_rng"], splatter.device)
        np.random.set_state(res["numpy_rng"])
    return ckpt
