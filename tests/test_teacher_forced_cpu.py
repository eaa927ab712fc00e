"""CPU: the two restatements of the bf16-storage training step agree with each other.

oracle/bf16_emulation.py states the step as torch autograd with bf16 roundings at the CUDA path's storage points;
oracle/teacher_forced.py states every single operation of forward AND backward explicitly (torch.nn.grad) and checks each
stored tensor against its own stored inputs.  Feeding the first into the second pins both (they share no backward code),
and shows that the teacher-forced bound catches a 10 % error injected into ONE gradient edge at exactly that edge."""
import pytest
import torch

from oracle import bf16_emulation as E
from oracle import teacher_forced as T
from oracle import unetpp_oracle as O


def _case(B, H, W, p_drop, seed):
    sd = O.synth_state_dict(seed=seed)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, 3, H, W, generator=g)
    target = torch.rand(B, 4, H, W, generator=g)
    masks = [(torch.rand(B, 16, H, W, generator=g) >= p_drop).to(torch.uint8) for _ in range(3)] if p_drop > 0 else None
    return sd, x, target, masks


@pytest.mark.parametrize("B,H,W,p_drop,loss", [(2, 32, 48, 0.4, "mse"), (1, 32, 32, 0.0, "focal")])
def test_emulated_step_passes_the_teacher_forced_check(B, H, W, p_drop, loss):
    sd, x, target, masks = _case(B, H, W, p_drop, 21)
    cap = {}
    _, heats, grads, _ = E.train_step_grads_bf16(sd, x, target, dropout_masks=masks, p_drop=p_drop, loss=loss, capture=cap)
    rep = T.verify_step(cap, heats, grads, sd, x, target=target, masks=masks, p_drop=p_drop, loss=loss)
    assert len(rep.rows) >= 176  # 94 stored tensors / statistics / heat maps + 82 parameter-gradient rows
    assert rep.check() == [], rep.check()
    # fp32 autograd against the fp64 restatement: far inside the bounds
    assert rep.worst("f32")["max_rel"] < 1e-5 and rep.worst("bf16")["frac_gt_ulp"] == 0.0


@pytest.mark.parametrize("edge", ["dU01", "dZ112", "conv10.dz2", "tmp1", "dP00", "dXh1"])
def test_a_ten_percent_error_in_one_gradient_edge_is_caught_at_that_edge(edge):
    sd, x, target, masks = _case(1, 32, 32, 0.0, 5)
    cap = {}
    _, heats, grads, _ = E.train_step_grads_bf16(sd, x, target, capture=cap, scale_grad={edge: 1.1})
    bad = T.verify_step(cap, heats, grads, sd, x, target=target).check()
    assert [r["name"] for r in bad if r["kind"] == "bf16"] == [edge], bad
    assert 0.05 < bad[0]["max_rel"] < 0.15


def test_emulation_tracks_the_fp32_oracle_within_the_bf16_noise():
    """The emulation is the fp32 oracle plus bf16 roundings, nothing else: same loss to 1e-2, same heat maps to 3e-2, and
    the head / full-resolution decoder gradients (the least chaotic end) within a few per cent."""
    sd, x, target, _ = _case(2, 32, 32, 0.0, 7)
    rl, routs, rg, rstats = O.train_step_grads(sd, x, target)
    el, eouts, eg, estats = E.train_step_grads_bf16(sd, x, target)
    assert abs(float(el) - float(rl)) <= 1e-2 * float(rl)
    for a, b in zip(eouts, routs):
        assert float((a - b).abs().max()) <= 3e-2
    for k in ("final_3.weight", "up_concat03.conv.conv2.0.weight", "up_concat01.conv.conv1.0.weight"):
        assert float((eg[k] - rg[k]).abs().max()) <= 5e-2 * float(rg[k].abs().max()), k
    for k, v in rstats.items():
        if not k.endswith("num_batches_tracked"):
            assert float((estats[k] - v).abs().max()) <= 1e-2 * float(v.abs().max()), k
