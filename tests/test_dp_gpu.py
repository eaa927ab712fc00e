"""GPU, two ranks over NCCL: the REAL data-parallel training step (fused.FusedTrainStep with world = 2) against two
single-GPU steps on the two sub-batches.

Reference semantics (trainer/trainer.py:283-285,336-338, nn.DataParallel): the batch is split evenly over the devices,
every replica normalises with its OWN BatchNorm batch statistics and draws its own dropout masks, the gradients of the
replicas are summed; SURVEY.md section 8(e): "world-N gradients == mean over N single-GPU runs on the N sub-batches".

Every reduction in the step has a fixed order and the all-reduce of two operands is one commutative fp32 add, so the
statement is checked BIT FOR BIT: flat gradient after the all-reduce == g_rank0 + g_rank1 of the single-GPU runs, updated
parameters == the same optimizer kernel applied to that sum with grad_scale = 1/2, BatchNorm running statistics of rank r ==
those of the single-GPU run on sub-batch r, and both ranks end with identical parameters."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, optimizer, loss, p_drop):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import unet_nested4tiny_objects_keypoints_b200 as pkg
    from unet_nested4tiny_objects_keypoints_b200 import fused, ops, parallel
    from oracle import unetpp_oracle as O

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    B, H, W = 4, 64, 64  # global batch; 2 images per rank
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 3, H, W, generator=g)
    target = torch.rand(B, 4, H, W, generator=g)
    xs, ts = parallel.shard_batch(x, world, rank), parallel.shard_batch(target, world, rank)
    sd = O.synth_state_dict(seed=50)
    hyper = dict(lr=1e-3, weight_decay=1e-4, optimizer=optimizer, loss=loss)

    def build():
        m = pkg.UNet_Nested()
        m.load_state_dict(sd)
        m = m.to(dev).train()
        m.drop_out.p = p_drop
        return m

    # ---- single-GPU step on this rank's sub-batch (no process group yet: world = 1); its dropout stream = seed 100 + rank
    m1 = build()
    st1 = fused.FusedTrainStep(m1, B // world, H, W, device=dev, seed=100 + rank, **hyper)
    p0 = st1.flat_p.clone()
    st1.step(xs.pin_memory(), ts.pin_memory())
    torch.cuda.synchronize()
    g_single = st1.flat_g.clone()
    loss_single = st1.loss.clone()
    stats_single = {k: v.clone() for k, v in m1.state_dict().items() if "running_" in k or "num_batches" in k}

    # ---- the data-parallel step
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    m2 = build()
    parallel.broadcast_parameters(m2, src=0)
    st2 = fused.FusedTrainStep(m2, B // world, H, W, device=dev, seed=100, **hyper)
    assert st2.world == world
    st2.step(xs.pin_memory(), ts.pin_memory())
    torch.cuda.synchronize()

    gs = [torch.empty_like(g_single) for _ in range(world)]
    dist.all_gather(gs, g_single)
    expect_g = gs[0] + gs[1]
    assert torch.equal(st2.flat_g, expect_g), f"rank {rank}: all-reduced gradient != sum of the single-GPU gradients (max diff {float((st2.flat_g - expect_g).abs().max()):.3e})"
    assert torch.equal(st2.loss, loss_single)  # the loss is per rank (its own sub-batch), like a replica's
    # the same optimizer kernel on the summed gradient with grad_scale = 1/world
    p_ref, s1, s2 = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    counter, scal = torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(4, device=dev)
    lr_dev = torch.full((1,), float(hyper["lr"]), device=dev)  # (the step count and the learning rate live on the device in the captured step)
    ops.optim_step(optimizer, p_ref, expect_g, s1, s2, grad_scale=1.0 / world, step_counter=counter, lr_dev=lr_dev, scalars=scal, **st2.hyper)
    torch.cuda.synchronize()
    assert torch.equal(st2.flat_p, p_ref), f"rank {rank}: updated parameters differ (max {float((st2.flat_p - p_ref).abs().max()):.3e})"
    assert not torch.equal(st2.flat_p, p0)
    # per-rank BatchNorm statistics (nn.DataParallel keeps them per replica)
    for k, v in m2.state_dict().items():
        if k in stats_single:
            assert torch.equal(v, stats_single[k]), (rank, k)
    # replicas stay identical
    ps = [torch.empty_like(st2.flat_p) for _ in range(world)]
    dist.all_gather(ps, st2.flat_p)
    assert torch.equal(ps[0], ps[1])
    # and a second step keeps working on the captured graphs (fresh dropout masks, step counter on the device)
    st2.step(xs.pin_memory(), ts.pin_memory())
    torch.cuda.synchronize()
    assert torch.isfinite(st2.loss).all() and int(st2.step_counter) == 2
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("optimizer,loss,p_drop", [("adamw", "mse", 0.0), ("adamw", "focal", 0.4), ("sgd", "mse", 0.4)])
def test_two_rank_nccl_step_equals_the_two_single_gpu_steps(optimizer, loss, p_drop):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), optimizer, loss, p_drop), nprocs=2, join=True)
