"""Drop-in for the reference's ``tools/optimizers/adamw.py`` ``AdamW`` (the trainer default,
train.py:38 / trainer.py:344-379): same constructor, same ``step()`` arithmetic — decoupled decay
``weight_decay * p_old`` that is NOT multiplied by the learning rate (adamw.py:91-96), denominator
``sqrt(v) + eps`` with both bias corrections folded into the step size (adamw.py:84-90) — executed by
ONE multi-tensor launch per parameter group (``unpp_optim_step_multi``: a pointer table of the group's tensors; the reference issues ~10
ATen kernels per tensor, 74 tensors), one launch per tensor only when the tensors of a group are at different step counts or on
different devices.  ``fused.FusedTrainStep`` goes further and keeps parameters, gradients and moments in flat buffers.  CUDA fp32 parameters only; ``amsgrad`` is not
implemented (the trainer never enables it).

``SGDW`` and ``AdaBound`` are the drop-ins for ``tools/optimizers/sgdw.py`` / ``tools/optimizers/adabound.py``
(``--optimizer sgdw|adabound``, trainer.py:364-376) on the same ``unpp_optim_step`` kernel; ``torch.optim.SGD`` /
``torch.optim.Adam`` (trainer.py:344-355) keep working unchanged on this model's parameters, and their flat-buffer
forms are ``FusedTrainStep(optimizer="sgd"|"adam")``.
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
            entries = []
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
                entries.append((p, g, state["exp_avg"], state["exp_avg_sq"], int(state["step"])))
            _launch_group("adamw", entries, dict(lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"], weight_decay=group["weight_decay"]))
        return loss


def _flat(t):
    return t.view(-1)


def _launch_group(kind, entries, hyper) -> None:
    """``entries``: [(p, grad, state1 | None, state2 | None, step)] of one parameter group.  One multi-tensor launch when they share the
    step count and the device, else one launch per tensor (same kernel arithmetic either way)."""
    if not entries:
        return
    same = len({e[4] for e in entries}) == 1 and len({e[0].device for e in entries}) == 1
    if same and len(entries) > 1:
        with torch.cuda.device(entries[0][0].device):
            ops.optim_step_multi(kind, [(_flat(p.data), _flat(g), None if a is None else _flat(a), None if b is None else _flat(b)) for p, g, a, b, _ in entries],
                                 step=entries[0][4], **hyper)
    else:
        for p, g, a, b, step in entries:
            with torch.cuda.device(p.device):
                ops.optim_step(kind, _flat(p.data), _flat(g), None if a is None else _flat(a), None if b is None else _flat(b), step=step, **hyper)
    torch._C._increment_version([e[0] for e in entries])


def _check(p):
    if p.grad.is_sparse:
        raise RuntimeError("sparse gradients are not supported")
    if not p.is_cuda or p.dtype != torch.float32:
        raise RuntimeError("the sm_100a optimizers update CUDA float32 parameters only (there is no CPU path)")
    return p.grad if p.grad.is_contiguous() else p.grad.contiguous()


class SGDW(Optimizer):
    """Drop-in for the reference's ``SGDW`` (tools/optimizers/sgdw.py:59-110), constructor and behaviour AS SHIPPED: ``step()``
    maintains the momentum buffer (first step ``buf = g``, later ``buf = momentum*buf + (1-dampening)*g``, sgdw.py:95-102) but
    never applies the descent direction — the only parameter update is the decoupled decay ``p -= weight_decay * p``
    (sgdw.py:107-108).  Kept bug-for-bug so that a run switched over reproduces the reference's parameters."""

    def __init__(self, params, lr, momentum=0, dampening=0, weight_decay=0, nesterov=False):
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            entries = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = _check(p)
                state = self.state[p]
                buf = None
                if group["momentum"] != 0:
                    if "momentum_buffer" not in state:
                        state["momentum_buffer"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                        state["step"] = 0
                    buf = state["momentum_buffer"]
                state["step"] = state.get("step", 0) + 1
                entries.append((p, g, buf, None, int(state["step"])))
            _launch_group("sgdw", entries, dict(lr=group["lr"], beta1=group["momentum"], beta2=group["dampening"], weight_decay=group["weight_decay"]))
        return loss


class AdaBound(Optimizer):
    """Drop-in for the reference's ``AdaBound`` (tools/optimizers/adabound.py:10-122): Adam moments with the L2 decay folded
    into the gradient, per-element rate ``clamp(step_size/(sqrt(v)+eps), lower(t), upper(t))`` with ``final_lr`` rescaled by
    ``lr/base_lr`` so that lr schedulers act on the bounds too (adabound.py:117-121).  ``amsbound`` is not implemented."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), final_lr=0.1, gamma=1e-3, eps=1e-8, weight_decay=0, amsbound=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= final_lr:
            raise ValueError("Invalid final learning rate: {}".format(final_lr))
        if not 0.0 <= gamma < 1.0:
            raise ValueError("Invalid gamma parameter: {}".format(gamma))
        if amsbound:
            raise NotImplementedError("amsbound is not implemented by the sm_100a AdaBound")
        super().__init__(params, dict(lr=lr, betas=betas, final_lr=final_lr, gamma=gamma, eps=eps, weight_decay=weight_decay, amsbound=amsbound))
        self.base_lrs = [group["lr"] for group in self.param_groups]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group, base_lr in zip(self.param_groups, self.base_lrs):
            b1, b2 = group["betas"]
            entries = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = _check(p)
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state["step"] += 1
                entries.append((p, g, state["exp_avg"], state["exp_avg_sq"], int(state["step"])))
            _launch_group("adabound", entries, dict(lr=group["lr"], beta1=b1, beta2=b2, eps=group["eps"], weight_decay=group["weight_decay"],
                                                    final_lr=group["final_lr"], gamma=group["gamma"], base_lr=base_lr))
        return loss
