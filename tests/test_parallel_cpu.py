"""CPU, world_size 2 over gloo: the data-parallel host logic (parallel.py) — sharding, the single
flat-gradient all-reduce, identical optimizer steps on every rank — checked against a single-process
run of the oracle on the concatenated batch (per-rank BatchNorm statistics, like nn.DataParallel)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from unet_nested4tiny_objects_keypoints_b200 import parallel


def test_shard_range_and_divisibility():
    assert parallel.shard_range(256, 8, 3) == (96, 128)
    assert [parallel.shard_range(6, 2, r) for r in range(2)] == [(0, 3), (3, 6)]
    with pytest.raises(ValueError):
        parallel.shard_range(10, 4, 0)  # trainer.py:285: batch must divide by the device count
    with pytest.raises(ValueError):
        parallel.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import unetpp_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    sd = O.synth_state_dict(seed=40 + rank)  # replicas start different on purpose
    import unet_nested4tiny_objects_keypoints_b200 as pkg
    model = pkg.UNet_Nested()
    model.load_state_dict(sd)
    parallel.broadcast_parameters(model, src=0)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 3, 16, 16, generator=g)
    target = torch.rand(4, 4, 16, 16, generator=g)
    xs, ts = parallel.shard_batch(x, world, rank), parallel.shard_batch(target, world, rank)
    # per-rank forward/backward (the oracle stands in for the CUDA engine: this test is about the host logic)
    _, _, grads, _ = O.train_step_grads(sd0, xs, ts, dropout_masks=None)
    names = [k for k, _ in model.named_parameters()]
    flat = torch.cat([grads[k].reshape(-1) for k in names])
    scale = parallel.allreduce_gradients(flat)
    p = torch.cat([sd0[k].reshape(-1) for k in names])
    p_new, _, _ = O.adamw_reference_step(p, flat * scale, torch.zeros_like(p), torch.zeros_like(p), 1, lr=1e-3, weight_decay=1e-4)
    out[rank] = (flat * scale, p_new, sd0["conv00.conv1.0.weight"].clone())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_mean_of_per_shard_gradients():
    from oracle import unetpp_oracle as O
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g0, p0, w0 = out[0]
    g1, p1, w1 = out[1]
    assert torch.equal(w0, w1)          # broadcast made the replicas identical
    assert torch.equal(g0, g1) and torch.equal(p0, p1)  # same reduced gradient, same update everywhere
    # reference: mean over the two shards' gradients, each with its own BatchNorm batch statistics
    sd0 = O.synth_state_dict(seed=40)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 3, 16, 16, generator=g)
    target = torch.rand(4, 4, 16, 16, generator=g)
    import unet_nested4tiny_objects_keypoints_b200 as pkg
    names = [k for k, _ in pkg.UNet_Nested().named_parameters()]
    ref = 0
    for r in range(world):
        _, _, grads, _ = O.train_step_grads(sd0, x[2 * r:2 * r + 2], target[2 * r:2 * r + 2], dropout_masks=None)
        ref = ref + torch.cat([grads[k].reshape(-1) for k in names])
    ref = ref / world
    assert torch.allclose(g0, ref, rtol=1e-5, atol=1e-8)
