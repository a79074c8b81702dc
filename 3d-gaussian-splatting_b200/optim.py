"""Fused flat-bucket Adam.

Drop-in for the way the reference steps its optimizer (train.py:56-64 / :173-181:
`torch.optim.Adam([{"params": p, "lr": lr} x 5], betas=(0.9, 0.99))`, learning rates rewritten in
`param_groups[i]["lr"]` every iteration, train.py:184-185): same constructor shape, `zero_grad`,
`step`, `param_groups`.  The fused backward leaves the five gradients as views of ONE flat buffer
(renderer._flat_grads; all-reduced in place by dp.GradBucket), so the step is a single kernel
(`gaussian.adam_step` -> `gs_adam_step`) over flat parameter / moment buffers instead of
5 x (foreach) kernel groups.  No CPU / torch fallback: gradients that are not one flat bucket are
an error.
"""
from __future__ import annotations

from typing import List

import torch

import gaussian


class FlatAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups: List[dict] = []
        for g in groups:
            ps = g["params"]
            ps = [ps] if isinstance(ps, torch.Tensor) else list(ps)
            self.param_groups.append({"params": ps, "lr": float(g.get("lr", lr))})
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self._flat = None          # (flat_param, exp_avg, exp_avg_sq, ordered params, segment ends)

    def zero_grad(self, set_to_none: bool = True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def _lr_of(self, p):
        for g in self.param_groups:
            if any(p is q for q in g["params"]):
                return g["lr"]
        raise KeyError("parameter not in any group")

    def _build(self, ordered):
        """Move the parameters into one flat buffer laid out exactly like the gradient bucket."""
        g0 = ordered[0].grad
        base = g0.storage_offset()
        span_end = ordered[-1].grad.storage_offset() + ordered[-1].grad.numel()
        total = (span_end - base + 3) // 4 * 4
        dev = g0.device
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        ends = []
        for i, p in enumerate(ordered):
            o = p.grad.storage_offset() - base
            view = flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view                                   # the Parameter now aliases the flat buffer
            nxt = ordered[i + 1].grad.storage_offset() - base if i + 1 < len(ordered) else total
            ends.append(nxt)
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        resume = getattr(self, "_resume", None)          # moments restored by checkpoint.load_checkpoint
        if resume is not None and resume[0] is not None and resume[0].numel() == flat.numel():
            m.copy_(resume[0])
            v.copy_(resume[1])
            self._keep_step = True
        self._resume = None
        self._flat = (flat, m, v, ordered, ends, base)

    @torch.no_grad()
    def step(self):
        params = [p for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not params:
            return
        ordered = sorted(params, key=lambda p: p.grad.storage_offset())
        g0 = ordered[0].grad
        store = g0.untyped_storage().data_ptr()
        for a, b in zip(ordered[:-1], ordered[1:]):
            gap = b.grad.storage_offset() - (a.grad.storage_offset() + a.grad.numel())
            if b.grad.untyped_storage().data_ptr() != store or not (0 <= gap <= 3):
                raise RuntimeError("FlatAdam needs the gradients to be views of one flat bucket "
                                   "(as produced by renderer.render_frame / render_frame_final)")
        if self._flat is None or len(self._flat[3]) != len(ordered) or any(a is not b for a, b in zip(self._flat[3], ordered)) \
                or any(p.data.untyped_storage().data_ptr() != self._flat[0].untyped_storage().data_ptr() for p in ordered):
            self._build(ordered)
            if not getattr(self, "_keep_step", False):
                self.step_count = 0
            self._keep_step = False
        flat, m, v, _, ends, _ = self._flat
        base = g0.storage_offset()
        gflat = torch.empty(0, dtype=torch.float32, device=g0.device).set_(g0.untyped_storage(), base, (flat.numel(),))
        if (base * 4 + g0.untyped_storage().data_ptr()) % 16:
            raise RuntimeError("FlatAdam: gradient bucket must be 16-byte aligned")
        self.step_count += 1
        gaussian.adam_step(flat, gflat, m, v, ends, [self._lr_of(p) for p in ordered], self.betas[0], self.betas[1],
                           self.eps, self.step_count)
