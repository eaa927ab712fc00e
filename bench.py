#!/usr/bin/env python
"""bench.py — UNet++ (UNet_Nested) hot-path throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input:
  * ``infer`` (default, BASELINE.json configs[1]): batch 128 of 3x256x256 fp32 images per GPU ->
    UNet_Nested forward (eval) -> heat maps of the deepest head -> arg-max keypoints.  Images are
    independent, so N GPUs run N shards with no data-path collective ("scaling": "weak").
  * ``train`` (configs[2] / configs[3]): fwd + bwd + fused MSE + AdamW at batch 32 on one GPU, or
    global batch 256 split over N GPUs with one NCCL all-reduce of the flat gradient buffer.
One JSON line is printed by rank 0 (see the keys below).  ``--impl reference`` times the reference
algorithm's CPU restatement (oracle/, torch-CPU fp32 — the reference is pure Python on PyTorch and
has no installable package) on the host cores for the same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

S = 256
INFER_B = 128
TRAIN_B1 = 32
TRAIN_GLOBAL = 256
ALGO_BYTES_PER_IMG_INFER = 62.05e6   # fused plan, bf16 activations, 256x256
ALGO_FLOPS_PER_IMG_INFER = 8.789e9
CONFIG_TAG = "configs[1]"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tc=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), which="measured")
    return dict(hbm=6650.0, tc=1590.0, which="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])), mx.append(float(f[2]))
                except ValueError:
                    continue
                for n, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
        except Exception:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    elif n > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    return rank, world, local


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def timed(fn, steps, world):
    """EXACTLY `steps` calls of fn between barrier+synchronize, timed with CUDA events on the launching stream."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def make_model(train: bool, device):
    import unet_nested4tiny_objects_keypoints_b200 as pkg
    torch.manual_seed(0)  # reference constructor init under seed 0, random-init weights
    m = pkg.UNet_Nested().to(device)
    return m.train() if train else m.eval()


def trace_kernels(fn, reps=3):
    """Per-launch durations of conv_tc / wgrad measured live with CUDA events (eager pass, not under a profiler).
    Returns {label: [total_ms over reps, launches over reps, algorithmic bytes per launch, flops per launch]}, reps."""
    from unet_nested4tiny_objects_keypoints_b200 import ops
    fn()
    torch.cuda.synchronize()
    agg = {}
    for _ in range(reps):
        ops.trace = []
        torch.cuda._sleep(int(60e6))  # ~30 ms of GPU spin: the host enqueues the whole pass ahead, so the events bracket pure GPU time
        fn()
        torch.cuda.synchronize()
        for label, e0, e1, nbytes, flops in ops.trace:
            a = agg.setdefault(label, [0.0, 0, nbytes, flops])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
        ops.trace = None
    return agg, reps


def roofline_from_trace(agg, reps, prefix, pk):
    """Roofline of one kernel (all its launches in a step): algorithmic bytes / CUDA-event time vs the measured HBM peak."""
    sel = {k: v for k, v in agg.items() if k.startswith(prefix)}
    if not sel:
        return None
    ms = sum(v[0] for v in sel.values()) / reps
    launches = sum(v[1] for v in sel.values()) // reps
    nbytes = sum(v[2] * v[1] for v in sel.values()) / reps
    flops = sum(v[3] * v[1] for v in sel.values()) / reps
    achieved = nbytes / (ms * 1e-3) / 1e9
    return {"kernel": prefix.strip(), "bound": "hbm", "achieved": round(achieved, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(achieved / pk["hbm"], 4),
            "traffic": None, "peak_source": pk["which"], "launches_per_step": launches, "ms_per_step_in_kernel": round(ms, 4),
            "algorithmic_bytes_per_launch_avg": round(nbytes / launches), "tensor_tflops": round(flops / (ms * 1e-3) / 1e12, 1)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_reference(steps, warmup, sample_b=4, train=False):
    """The reference algorithm on the host cores: oracle/unetpp_oracle.py (torch-CPU fp32, all threads)."""
    from oracle import unetpp_oracle as O
    import numpy as np
    torch.set_num_threads(os.cpu_count() or 1)
    sd = O.synth_state_dict(seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(sample_b, 3, S, S, generator=g)
    target = torch.rand(sample_b, 4, S, S, generator=g)

    def one():
        if train:
            O.train_step_grads(sd, x, target, dropout_masks=None)
        else:
            with torch.no_grad():
                heat = O.forward(sd, x)[2]
            O.argmax_keypoints(heat.numpy())

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return sample_b * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def run_reference_arm(args, rank, world, out):
    if rank != 0:
        return
    train = args.workload == "train"
    sample_b = 2 if train else 4
    ips, ms, cores = cpu_reference(args.steps, args.warmup, sample_b, train)
    sample = f"{sample_b} of the {TRAIN_B1 if train else INFER_B} images of one step per timed step, {args.steps} steps, torch-CPU fp32 oracle port"
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": round(ips, 3), "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args.workload, world),
            "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


def metric_name(workload):
    return "UNet++ images/sec (inference 256x256 + arg-max keypoints)" if workload == "infer" else "UNet++ images/sec (train step 256x256: fwd+bwd+MSE+AdamW)"


def config_of(workload, world):
    if workload == "infer":
        return {"workload": f"{CONFIG_TAG}: UNet_Nested eval forward, batch {INFER_B}/GPU, 3x{S}x{S} fp32 in, fused head + heat-map arg-max keypoints",
                "batch_per_gpu": INFER_B, "image": [3, S, S], "l2": "activation working set (>= 268 MB per level-0 tensor) is larger than the 126 MB L2; no flush needed",
                "parallelism": f"dp{world} (independent shards, no collective)", "cuda_graph": True}
    b = TRAIN_B1 if world == 1 else TRAIN_GLOBAL // world
    return {"workload": ("configs[2]: train step batch 32" if world == 1 else f"configs[3]: data-parallel train step, global batch {TRAIN_GLOBAL}") +
            f", 3x{S}x{S}, BN batch stats, dropout 0.4, MSE on 3 heads, reference AdamW", "batch_per_gpu": b, "image": [3, S, S],
            "l2": "activation working set larger than L2; no flush needed", "parallelism": f"dp{world} (one NCCL all-reduce of the 2.2 MB flat gradient per step)",
            "cuda_graph": True}


# ---------------------------------------------------------------------------------------------- GPU arms
def bench_infer(args, rank, world, local):
    from unet_nested4tiny_objects_keypoints_b200 import fused, ops
    dev = torch.device("cuda", local)
    model = make_model(False, dev)
    sess = fused.InferenceSession(model, INFER_B, S, S, head=2, device=dev)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(INFER_B, 3, S, S, generator=g).pin_memory()
    sess.x.copy_(x_host)
    for _ in range(max(args.warmup, 3)):
        sess.run_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(sess.run_device, args.steps, world)
    clk = clocks.stop() if rank == 0 else None
    value = INFER_B * world * args.steps / (ms * 1e-3)

    # end to end through the public API (InferenceSession.run_many): every step a pinned host batch goes H2D,
    # through forward + arg-max, and its keypoints come back D2H; the copy of batch k+1 overlaps the kernels of batch k
    xy_host = [torch.empty(INFER_B, model.n_classes, 2, dtype=torch.int32).pin_memory() for _ in range(2)]
    val_host = [torch.empty(INFER_B, model.n_classes, dtype=torch.float32).pin_memory() for _ in range(2)]
    x_hosts = [x_host, torch.randn(INFER_B, 3, S, S, generator=g).pin_memory()]

    def e2e_run(nsteps):
        sess.run_many([x_hosts[k & 1] for k in range(nsteps)], [xy_host[k & 1] for k in range(nsteps)], [val_host[k & 1] for k in range(nsteps)])

    e2e_run(2)
    torch.cuda.synchronize()
    ms_e2e = timed(lambda: e2e_run(args.steps), 1, world)
    e2e = {"value": round(INFER_B * world * args.steps / (ms_e2e * 1e-3), 1), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
           "d2h_bytes_per_step": xy_host[0].numel() * 4 + val_host[0].numel() * 4, "ms_per_step": round(ms_e2e / args.steps, 3),
           "pipelining": "H2D of batch k+1 overlaps the kernels of batch k (two input buffers, copy stream)"}

    line = None
    if rank == 0:
        pk = peaks()
        with torch.no_grad():
            agg, reps = trace_kernels(lambda: sess._body(0))
        roof = roofline_from_trace(agg, reps, "conv_tc", pk)
        step_ms = ms / args.steps
        roof["share_of_step"] = round(roof["ms_per_step_in_kernel"] / step_ms, 3)
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and S == 256 and INFER_B == 128:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
            with open(tpath) as f:
                tj = json.load(f)
            roof["traffic"] = round(tj["traffic_bytes_per_launch_avg"])
            roof["traffic_source"] = tj["source"]
        whole = {"algorithmic_bytes_per_image": ALGO_BYTES_PER_IMG_INFER, "achieved_gbs": round(ALGO_BYTES_PER_IMG_INFER * INFER_B / (step_ms * 1e-3) / 1e9, 1),
                 "frac_of_hbm_peak": round(ALGO_BYTES_PER_IMG_INFER * INFER_B / (step_ms * 1e-3) / 1e9 / pk["hbm"], 4),
                 "tflops": round(ALGO_FLOPS_PER_IMG_INFER * INFER_B / (step_ms * 1e-3) / 1e12, 1)}
        line = {"metric": metric_name("infer"), "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_of("infer", world), "clocks": clk, "e2e": e2e, "gpu_launches": sess.launches * args.steps, "roofline": roof,
                "whole_step_roofline": whole,
                "per_kernel": {k: {"ms": round(v[0] / v[1], 4), "GBps": round(v[2] / (v[0] / v[1] * 1e-3) / 1e9, 1), "TFLOPs": round(v[3] / (v[0] / v[1] * 1e-3) / 1e12, 1)}
                               for k, v in sorted(agg.items()) if v[2]}}
    return line


def bench_train(args, rank, world, local):
    from unet_nested4tiny_objects_keypoints_b200 import fused, ops
    dev = torch.device("cuda", local)
    model = make_model(True, dev)
    if world > 1:
        fused.broadcast_parameters(model)
    b = TRAIN_B1 if world == 1 else TRAIN_GLOBAL // world
    loss = getattr(args, "loss", "mse")  # BASELINE.json configs[2]/[3] name the MSE heat-map loss; "focal" = the trainer's shipped criterion
    step = fused.FusedTrainStep(model, b, S, S, device=dev, seed=0, loss=loss)
    g = torch.Generator().manual_seed(99 + rank)
    x_host = torch.randn(b, 3, S, S, generator=g).pin_memory()
    t_host = torch.rand(b, 4, S, S, generator=g).pin_memory()
    step.x.copy_(x_host)
    step.target.copy_(t_host)
    for _ in range(max(args.warmup, 3)):
        step.step_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms = timed(step.step_device, args.steps, world)
    clk = clocks.stop() if rank == 0 else None
    value = b * world * args.steps / (ms * 1e-3)
    # end to end through the public API (FusedTrainStep.step_many): every step its images and targets go H2D from pinned
    # host memory and its loss comes back D2H; the copies of batch k+1 overlap the step of batch k
    x_hosts = [x_host, torch.randn(b, 3, S, S, generator=g).pin_memory()]
    t_hosts = [t_host, torch.rand(b, 4, S, S, generator=g).pin_memory()]
    loss_hosts = [torch.empty(1).pin_memory() for _ in range(2)]

    def e2e_run(nsteps):
        step.step_many([x_hosts[k & 1] for k in range(nsteps)], [t_hosts[k & 1] for k in range(nsteps)], [loss_hosts[k & 1] for k in range(nsteps)])

    e2e_run(2)
    torch.cuda.synchronize()
    ms_e2e = timed(lambda: e2e_run(args.steps), 1, world)
    e2e = {"value": round(b * world * args.steps / (ms_e2e * 1e-3), 1), "unit": "images/s", "h2d_bytes_per_step": (x_host.numel() + t_host.numel()) * 4,
           "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3),
           "pipelining": "H2D of batch k+1 (staging buffers, copy stream) overlaps the step of batch k"}
    line = None
    if rank == 0:
        pk = peaks()
        n0 = ops.launch_count
        step._fwd_bwd()
        step._update()
        launches = ops.launch_count - n0
        from unet_nested4tiny_objects_keypoints_b200 import training as _tr
        side_on, _tr.WGRAD_SIDE_STREAM = _tr.WGRAD_SIDE_STREAM, False  # per-kernel events: one kernel at a time (the timed steps above ran with the side stream)
        try:
            agg, reps = trace_kernels(lambda: (step._fwd_bwd(), step._update()))
        finally:
            _tr.WGRAD_SIDE_STREAM = side_on
        roof = roofline_from_trace(agg, reps, "conv_tc", pk)
        roof_w = roofline_from_trace(agg, reps, "wgrad taps", pk)
        step_ms = ms / args.steps
        roof["share_of_step"] = round(roof["ms_per_step_in_kernel"] / step_ms, 3)
        roof_w["share_of_step"] = round(roof_w["ms_per_step_in_kernel"] / step_ms, 3)
        line = {"metric": metric_name("train"), "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config_of("train", world), "clocks": clk, "e2e": e2e, "gpu_launches": launches * args.steps,
                "roofline": roof, "roofline_wgrad": roof_w, "loss": float(step.loss.item()),
                "per_kernel": {k: {"ms": round(v[0] / v[1], 4), "GBps": round(v[2] / (v[0] / v[1] * 1e-3) / 1e9, 1), "TFLOPs": round(v[3] / (v[0] / v[1] * 1e-3) / 1e12, 1)}
                               for k, v in sorted(agg.items()) if v[2]}}
    return line


def main():
    # Only the JSON line may reach stdout (NCCL and friends print banners there): everything else goes to stderr.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "infer1024"])
    ap.add_argument("--no-extra", action="store_true", help="skip the cpu_baseline leg and the extra training measurement")
    args = ap.parse_args()
    if args.workload == "infer1024":  # BASELINE.json configs[4]: high-resolution inference, batch 16 at 1024x1024
        global S, INFER_B, ALGO_BYTES_PER_IMG_INFER, ALGO_FLOPS_PER_IMG_INFER, CONFIG_TAG
        S, INFER_B, ALGO_BYTES_PER_IMG_INFER, ALGO_FLOPS_PER_IMG_INFER, CONFIG_TAG = 1024, 16, 976.3e6, 140.63e9, "configs[4]"
        args.workload, args.no_extra = "infer", True

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")), out)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = dist_setup(args.gpus)
    line = bench_infer(args, rank, world, local) if args.workload == "infer" else bench_train(args, rank, world, local)
    extra = None
    if args.workload == "infer" and not args.no_extra:
        # the other half of the metric ("infer & train step"): the training step at this world size
        targs = argparse.Namespace(**vars(args))
        targs.steps, targs.warmup = min(args.steps, 10), 3
        torch.cuda.empty_cache()
        t = bench_train(targs, rank, world, local)
        if t is not None:
            extra = {k: t[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e", "gpu_launches", "roofline", "roofline_wgrad", "loss")}
    if rank == 0:
        if world == 1 and not args.no_extra:
            train = args.workload == "train"
            # ~10 s of host work: 150 four-image inference batches (or 25 two-image training steps) of the same workload
            nb, bs = (25, 2) if train else (150, 4)
            ips, ms_cpu, cores = cpu_reference(nb, 2, bs, train)
            line["cpu_baseline"] = {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": f"{bs}-image batches of the same workload, 2 warm-up + {nb} timed, oracle/unetpp_oracle.py on torch-CPU fp32"}
            if not train:  # BASELINE.json configs[0]: the reference's own CPU-runnable case, batch 1 (3 warm-up + 10 timed)
                ips1, ms1, _ = cpu_reference(10, 3, 1, False)
                line["cpu_baseline"]["configs0_batch1_images_per_s"] = round(ips1, 3)
        else:
            line["cpu_baseline"] = None
        if extra is not None:
            line["train_step"] = extra
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
