"""Drop-in for the reference's ``tools/optimizers/adamw.py`` ``AdamW`` (the trainer default,
train.py:38 / trainer.py:344-379): same constructor, same ``step()`` arithmetic — decoupled decay
``weight_decay * p_old`` that is NOT multiplied by the learning rate (adamw.py:91-96), denominator
``sqrt(v) + eps`` with both bias corrections folded into the step size (adamw.py:84-90) — executed by
one ``unpp_adamw`` launch per parameter tensor (the reference issues ~10 ATen kernels per tensor),
or ONE launch for the whole model when the parameters are views of a flat buffer
(``fused.FusedTrainStep`` lays them out that way).  CUDA fp32 parameters only; ``amsgrad`` is not
implemented (the trainer never enables it).
"""
from __future__ import annotations

import torch
from torch.optim import Optimizer

from . import ops


class AdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if amsgrad:
            raise NotImplementedError("amsgrad is not implemented by the sm_100a AdamW")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("the sm_100a AdamW updates CUDA float32 parameters only (there is no CPU path)")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                with torch.cuda.device(p.device):
                    ops.adamw(p.data.view(-1), g.view(-1), state["exp_avg"].view(-1), state["exp_avg_sq"].view(-1), group["lr"], b1, b2, group["eps"],
                              group["weight_decay"], state["step"])
        return loss
