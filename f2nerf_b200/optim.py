"""FusedAdam — SURVEY §8(f) N1: ``torch::optim::Adam`` (the optimizer the reference builds in
``src/ExpRunner.cpp:54`` from ``Renderer::OptimParamGroups`` and steps in ``:136``) as ONE kernel per parameter.

Same state (``step``, ``exp_avg``, ``exp_avg_sq``), same hyper-parameter handling (per-group ``lr`` / ``betas`` /
``eps`` / ``weight_decay``; ``lr`` is re-read every step so the trainer's schedule, ``ExpRunner.cpp:228-240``, keeps
working by assigning ``group["lr"]``), and bit-identical parameters after every step (the C ABI call reproduces the
ATen elementwise sequence rounding for rounding, see ``csrc/optim.cu``).  For the hash table it additionally
 * touches only the live 17/32 of the pool (the level-overlap quirk leaves the rest with zero gradient and zero
   moments for ever, where the Adam update is the identity), and
 * writes the fp16 shadow the next forward reads, so ``Hash3DAnchored.table_f16()`` stops converting 64 MB per call.
"""
import torch

from ._lib import call, stream


class FusedAdam:
    def __init__(self, param_groups, table_field=None):
        """``param_groups``: list of dicts as returned by ``Renderer.OptimParamGroups()``
        (keys ``params``, ``lr``, ``betas``, ``eps``, optional ``weight_decay``).
        ``table_field``: the ``Hash3DAnchored`` whose ``feat_pool_`` gets the live-prefix / fp16-shadow treatment."""
        self.param_groups = [dict(g) for g in param_groups]
        for g in self.param_groups:
            g.setdefault("betas", (0.9, 0.999)); g.setdefault("eps", 1e-8); g.setdefault("weight_decay", 0.0)
            g["params"] = list(g["params"])
        self.state = {}
        self.table_field = table_field
        if table_field is not None:
            table_field.manage_shadow(True)

    def zero_grad(self):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self):
        for g in self.param_groups:
            b1, b2 = g["betas"]
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.numel() % 4 == 0):
                    raise ValueError("FusedAdam: parameters must be contiguous CUDA fp32 tensors with numel % 4 == 0")
                st = self.state.get(id(p))
                if st is None:
                    st = self.state[id(p)] = dict(step=0, exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p))
                st["step"] += 1
                grad = p.grad.contiguous()
                n = p.numel()
                n_live, shadow = n, None
                f = self.table_field
                if f is not None and p is f.feat_pool_:
                    n_live = min(n, (f.n_levels_ + 1) * f.local_size_)            # halves [0, 17 S) are the only ones ever addressed
                    shadow = f.shadow_for_update()
                call("f2b_adam_step", p, grad, st["exp_avg"], st["exp_avg_sq"], n, n_live, float(g["lr"]), float(b1), float(b2),
                     float(g["eps"]), float(g["weight_decay"]), int(st["step"]), shadow, stream())
                if shadow is not None:
                    f.shadow_updated()
