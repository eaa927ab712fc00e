#!/usr/bin/env python
"""bench.py — UNet++ (UNet_Nested) hot-path throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload infer|train|infer1024] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic input:
  * ``infer`` (default, BASELINE.json configs[1]): batch 128 of 3x256x256 images per GPU -> UNet_Nested forward (eval) ->
    ALL THREE heat maps ``forward`` returns (models/unet.py:300) -> arg-max keypoints of every plane.  Images are independent,
    so N GPUs run N shards with no data-path collective ("scaling": "weak").
  * ``train`` (configs[2] / configs[3]): fwd + bwd + fused MSE + AdamW at batch 32 on one GPU, or global batch 256 split over
    N GPUs with one NCCL all-reduce of the flat gradient buffer.
  * ``infer1024`` (configs[4]): batch 16 at 1024x1024.
One JSON line is printed by rank 0.  The default run also measures the training step, the 1024x1024 case and the unchanged
trainer's eager sequence and reports them under ``train_step`` / ``infer1024`` / ``eager_dropin``.

Roofline accounting: the numerator is the ALGORITHMIC traffic of SURVEY.md section 8(d) — every kernel of
the fused plan reads each logical source once at its native resolution and writes each output once (2-byte activations; concat,
upsample, BatchNorm, ReLU, dropout, sigmoid cost nothing; a pooled copy is an extra quarter-size write) — enumerated by
``plan_rows`` below (62.05 MB / image forward at 256x256), NOT what our launches happen to move (a materialised upsampled
tensor earns no credit).  The denominator is the time of the launches that implement a plan row: per-launch CUDA-event
durations of an eager pass give every launch's SHARE, the absolute scale is the CUDA-graph step the headline is timed on
(a kernel can never be credited with more time than the step holds).  Per row ``bound = max(bytes / peak_HBM, flops / peak_TC)``.

``--impl reference`` times the reference algorithm's CPU restatement (oracle/, torch-CPU fp32 — the reference is pure Python on
PyTorch and has no installable package) on the host cores for the same metric.
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
CONFIG_TAG = "configs[1]"
F = (16, 32, 64, 128)  # filters of the default constructor (models/unet.py:214-216)
OTHER = "other (BatchNorm apply / backward, head backward, pool backward, reductions, packing, optimizer)"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        # a kernel timed inside a long step: the SUSTAINED tensor figure; HBM: the measured copy bandwidth
        return dict(hbm=float(p["hbm_gbs"]), tc=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), which="MEASURED_PEAKS.json (hbm copy, bf16 sustained)")
    return dict(hbm=6650.0, tc=1590.0, which="fallback of B200_PROFILING.md")


# ---------------------------------------------------------------------------------------------- the fused plan of SURVEY.md 8(d)
DEC = {"01": ("10", ("00",), 0), "11": ("20", ("10",), 1), "21": ("30", ("20",), 2), "02": ("11", ("00", "01"), 0), "12": ("21", ("10", "11"), 1),
       "03": ("12", ("00", "01", "02"), 0)}


def plan_rows(side: int, in_ch: int = 3, ncls: int = 4):
    """Forward rows of the fused plan, per image: [(name, bytes read, bytes written, flops)] with 2-byte activations."""
    px = [side * side >> (2 * l) for l in range(4)]
    rows = []
    cin = in_ch
    for l in range(4):
        c = F[l]
        rows.append((f"conv{l}0.c1", cin * px[l] * 2, c * px[l] * 2, 2 * 9 * cin * c * px[l]))
        rows.append((f"conv{l}0.c2", c * px[l] * 2, c * px[l] * 2 + (c * px[l + 1] * 2 if l < 3 else 0), 2 * 9 * c * c * px[l]))
        cin = c
    for tag in ("01", "11", "21", "02", "12", "03"):
        high, lows, l = DEC[tag]
        c = F[l]
        k1 = c * (1 + len(lows))
        rd = F[l + 1] * px[l + 1] * 2 + len(lows) * c * px[l] * 2
        rows.append((f"up{tag}.c1", rd, c * px[l] * 2, 2 * 4 * F[l + 1] * c * px[l + 1] + 2 * 9 * k1 * c * px[l]))
        head = l == 0
        wr = (0 if tag == "03" else c * px[l] * 2) + (ncls * px[l] * 2 if head else 0)
        rows.append((f"up{tag}.c2", c * px[l] * 2, wr, 2 * 9 * c * c * px[l] + (2 * c * ncls * px[l] if head else 0)))
    return rows


def plan_train(side: int, in_ch: int = 3, ncls: int = 4):
    """Per image: (forward bytes, dgrad bytes, wgrad bytes, forward flops).  Backward rule of SURVEY 8(d): per conv the dgrad reads dZ
    and writes dX once, the wgrad reads dZ and the saved input once (dZ is read twice: the two are separate kernels); the first
    conv needs no dgrad; transposed convs and the 1x1 heads follow the same rule."""
    px = [side * side >> (2 * l) for l in range(4)]
    fwd = sum(r[1] + r[2] for r in plan_rows(side, in_ch, ncls))
    flops = sum(r[3] for r in plan_rows(side, in_ch, ncls))
    dgrad = wgrad = 0
    cin = in_ch
    for l in range(4):
        c = F[l]
        for k in (cin, c):
            if not (l == 0 and k == cin):
                dgrad += (c + k) * px[l] * 2
            wgrad += (c + k) * px[l] * 2
        cin = c
    for tag in ("01", "11", "21", "02", "12", "03"):
        high, lows, l = DEC[tag]
        c = F[l]
        k1 = c * (1 + len(lows))
        dgrad += (c + k1) * px[l] * 2 + 2 * c * px[l] * 2            # conv1 (wrt the whole concat) and conv2
        wgrad += (c + k1) * px[l] * 2 + 2 * c * px[l] * 2
        dgrad += c * px[l] * 2 + F[l + 1] * px[l + 1] * 2             # transposed conv: read dU, write the gradient of its low-resolution input
        wgrad += c * px[l] * 2 + F[l + 1] * px[l + 1] * 2             # read dU and the saved low-resolution input
        if l == 0:                                                     # 1x1 head: d(logit) and X
            dgrad += (ncls + c) * px[0] * 2
            wgrad += (ncls + c) * px[0] * 2
    return fwd, dgrad, wgrad, flops


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


def h2d_bandwidth(dev, nbytes=256 << 20):
    """Measured pinned host -> device copy bandwidth of this rank (GB/s): the ceiling of any end-to-end number."""
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return 4 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9


def trace_kernels(fn, reps=3):
    """Per-launch durations measured live with CUDA events (eager pass, not under a profiler).
    Returns {(tag, label): [total_ms over reps, launches over reps]}, reps."""
    from unet_nested4tiny_objects_keypoints_b200 import ops
    fn()
    torch.cuda.synchronize()
    agg = {}
    for _ in range(reps):
        ops.trace = []
        torch.cuda._sleep(int(60e6))  # ~30 ms of GPU spin: the host enqueues the whole pass ahead, so the events bracket pure GPU time
        fn()
        torch.cuda.synchronize()
        for label, e0, e1, nbytes, flops, tag in ops.trace:
            a = agg.setdefault((tag, label), [0.0, 0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1
        ops.trace = None
    return agg, reps


def scaled_times(agg, reps, step_ms):
    """Per-(tag, label) time inside ONE graph step: eager event shares times the graph step (never more than the step holds)."""
    eager = {k: v[0] / reps for k, v in agg.items()}
    total = sum(eager.values())
    scale = min(1.0, step_ms / total) if total > 0 else 1.0
    return {k: v * scale for k, v in eager.items()}, total, scale


def traffic_of(kind):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this round (or null)."""
    for name in (f"r02_traffic_{kind}.json", "r01_traffic.json" if kind == "infer" else ""):
        path = os.path.join(ROOT, "profiles", name)
        if name and os.path.exists(path):
            with open(path) as f:
                tj = json.load(f)
            return round(tj["traffic_bytes_per_launch_avg"]), tj["source"]
    return None, None


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_reference(steps, warmup, sample_b=4, train=False, side=None):
    """The reference algorithm on the host cores: oracle/unetpp_oracle.py (torch-CPU fp32, all threads)."""
    from oracle import unetpp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    side = side or S
    sd = O.synth_state_dict(seed=0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(sample_b, 3, side, side, generator=g)
    target = torch.rand(sample_b, 4, side, side, generator=g)

    def one():
        if train:
            O.train_step_grads(sd, x, target, dropout_masks=None)
        else:
            with torch.no_grad():
                heats = O.forward(sd, x)
            for h in heats:  # the validation loop extracts key points from every output (trainer/trainer.py:212-221)
                O.argmax_keypoints(h.numpy())

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
    sample_b = 2 if train else (1 if S > 256 else 4)
    ips, ms, cores = cpu_reference(args.steps, args.warmup, sample_b, train)
    sample = f"{sample_b} of the {TRAIN_B1 if train else INFER_B} images of one step per timed step, {args.steps} steps, torch-CPU fp32 oracle port"
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": round(ips, 3), "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_of(args.workload, world),
            "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), file=out, flush=True)


def metric_name(workload):
    if workload == "infer":
        return f"UNet++ images/sec (inference {S}x{S}, three heads + arg-max keypoints)"
    return "UNet++ images/sec (train step 256x256: fwd+bwd+MSE+AdamW)"


def config_of(workload, world):
    if workload == "infer":
        return {"workload": f"{CONFIG_TAG}: UNet_Nested eval forward, batch {INFER_B}/GPU, 3x{S}x{S} in, heads computed: 3 of 3 (final_1..3, fused into their "
                            "producing convs), arg-max keypoints of all 12 planes per image",
                "batch_per_gpu": INFER_B, "image": [3, S, S], "l2": "activation working set (>= 268 MB per level-0 tensor) is larger than the 126 MB L2; no flush needed",
                "parallelism": f"dp{world} (independent shards, no collective)", "cuda_graph": True}
    b = TRAIN_B1 if world == 1 else TRAIN_GLOBAL // world
    return {"workload": ("configs[2]: train step batch 32" if world == 1 else f"configs[3]: data-parallel train step, global batch {TRAIN_GLOBAL}") +
            ", 3x256x256, BN batch stats, dropout 0.4, MSE on 3 heads, reference AdamW", "batch_per_gpu": b, "image": [3, 256, 256],
            "l2": "activation working set larger than L2; no flush needed", "parallelism": f"dp{world} (one NCCL all-reduce of the 2.2 MB flat gradient per step)",
            "cuda_graph": True}


# ---------------------------------------------------------------------------------------------- GPU arms
def bench_infer(args, rank, world, local):
    from unet_nested4tiny_objects_keypoints_b200 import fused
    dev = torch.device("cuda", local)
    model = make_model(False, dev)
    heads = (0, 1, 2)
    sess = fused.InferenceSession(model, INFER_B, S, S, head=heads, device=dev)
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
    nh, ncls = len(heads), model.n_classes

    # end to end through the public API (InferenceSession.run_many): every step a pinned host batch goes H2D, through forward + arg-max,
    # and its keypoints come back D2H; the copy of batch k+1 overlaps the kernels of batch k.  Two input forms of the same images:
    #   e2e             8-bit RGB [B,H,W,3] as the reference's dataset decodes them (datasets/datasets_base.py:71-72); ToTensor's 1/255 runs on the device
    #   e2e_fp32_input  the fp32 [B,3,H,W] tensors the reference copies (trainer/trainer.py:109): 4x the bytes
    def e2e_of(s, hosts):
        xy_host = [torch.empty(nh, INFER_B, ncls, 2, dtype=torch.int32).pin_memory() for _ in range(2)]
        val_host = [torch.empty(nh, INFER_B, ncls, dtype=torch.float32).pin_memory() for _ in range(2)]

        def run(nsteps):
            s.run_many([hosts[k & 1] for k in range(nsteps)], [xy_host[k & 1] for k in range(nsteps)], [val_host[k & 1] for k in range(nsteps)])

        run(2)
        torch.cuda.synchronize()
        ms_e = timed(lambda: run(args.steps), 1, world)
        return {"value": round(INFER_B * world * args.steps / (ms_e * 1e-3), 1), "unit": "images/s", "h2d_bytes_per_step": hosts[0].numel() * hosts[0].element_size(),
                "d2h_bytes_per_step": xy_host[0].numel() * 4 + val_host[0].numel() * 4, "ms_per_step": round(ms_e / args.steps, 3),
                "pipelining": "H2D of batch k+1 overlaps the kernels of batch k (two input buffers, copy stream)"}

    e2e_f32 = e2e_of(sess, [x_host, torch.randn(INFER_B, 3, S, S, generator=g).pin_memory()])
    e2e_f32["input"] = "float32 [B,3,H,W] (the tensors the reference copies, trainer.py:109)"
    sess8 = fused.InferenceSession(model, INFER_B, S, S, head=heads, device=dev, input="uint8_nhwc")
    u8 = [torch.randint(0, 256, (INFER_B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
    e2e = e2e_of(sess8, u8)
    e2e["input"] = "uint8 RGB [B,H,W,3] (what the reference's dataset decodes, datasets_base.py:71-72); ToTensor's 1/255 fused into the on-device layout change"
    e2e["h2d_gbs_measured"] = round(h2d_bandwidth(dev), 1)
    del sess8

    line = None
    if rank == 0:
        pk = peaks()
        with torch.no_grad():
            agg, reps = trace_kernels(lambda: sess._body(0))
        step_ms = ms / args.steps
        times, eager_total, scale = scaled_times(agg, reps, step_ms)
        by_tag = {}
        for (tag, label), t in times.items():
            if label.startswith("conv_tc"):
                by_tag[tag] = by_tag.get(tag, 0.0) + t
        # weights (553 260 parameters, 2 bytes each in the packed operands) are read once per launch, i.e. once per STEP: SURVEY 8(d)'s 62.05 MB / image
        # books them per image (60.95 + 1.11); here they are added once, which makes every fraction below ~1.8 % smaller than 62.05 MB x batch would
        rows, t_bound_sum, conv_bytes, conv_flops = [], 0.0, 553260 * 2.0, 0.0
        for name, rd, wr, fl in plan_rows(S):
            nbytes, flops = (rd + wr) * INFER_B, fl * INFER_B
            t_h, t_t = nbytes / (pk["hbm"] * 1e9) * 1e3, flops / (pk["tc"] * 1e12) * 1e3
            t_b, t_m = max(t_h, t_t), by_tag.get(name, 0.0)
            t_bound_sum += t_b
            conv_bytes += nbytes
            conv_flops += flops
            rows.append({"row": name, "ms": round(t_m, 4), "bound": "hbm" if t_h >= t_t else "tensor", "frac": round(t_b / t_m, 3) if t_m else None,
                         "GBps": round(nbytes / (t_m * 1e-3) / 1e9, 1) if t_m else None, "TFLOPs": round(flops / (t_m * 1e-3) / 1e12, 1) if t_m else None})
        t_conv = sum(t for (tag, label), t in times.items() if label.startswith("conv_tc"))
        n_conv = sum(v[1] for (tag, label), v in agg.items() if label.startswith("conv_tc")) // reps
        achieved = conv_bytes / (t_conv * 1e-3) / 1e9
        traffic, tsrc = traffic_of("infer")
        roof = {"kernel": "conv_tc (every launch of one step)", "bound": "hbm", "achieved": round(achieved, 1), "peak": pk["hbm"], "unit": "GB/s",
                "frac": round(achieved / pk["hbm"], 4), "traffic": traffic, "traffic_source": tsrc, "peak_source": pk["which"], "launches_per_step": n_conv,
                "ms_per_step_in_kernel": round(t_conv, 4), "share_of_step": round(t_conv / step_ms, 3),
                "algorithmic_bytes_per_launch_avg": round(conv_bytes / n_conv), "algorithmic_bytes_per_image": round(conv_bytes / INFER_B),
                "tensor_tflops": round(conv_flops / (t_conv * 1e-3) / 1e12, 1),
                "timing": f"CUDA-event share of each launch in an eager pass ({eager_total:.3f} ms serialised) x the graph step ({step_ms:.3f} ms): scale {scale:.3f}"}
        whole = {"algorithmic_bytes_per_image": round(conv_bytes / INFER_B), "achieved_gbs": round(conv_bytes / (step_ms * 1e-3) / 1e9, 1),
                 "frac_of_hbm_peak": round(conv_bytes / (step_ms * 1e-3) / 1e9 / pk["hbm"], 4),
                 "frac_of_per_row_roofline": round(t_bound_sum / step_ms, 4), "roofline_ms": round(t_bound_sum, 4),
                 "tflops": round(conv_flops / (step_ms * 1e-3) / 1e12, 1)}
        other = {label: round(t, 4) for (tag, label), t in sorted(times.items()) if not label.startswith("conv_tc")}
        line = {"metric": metric_name("infer"), "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_of("infer", world), "clocks": clk, "e2e": e2e, "e2e_fp32_input": e2e_f32, "gpu_launches": sess.launches * args.steps,
                "roofline": roof, "whole_step_roofline": whole, "per_kernel": rows, "other_kernels_ms": other}
    return line


def focal_bce_2d(p, t, gamma=3.0):
    """The trainer's heatmap criterion (tools/losses/focal_loss.py:264-301, size_average=False) as plain torch ops: user code of the eager leg."""
    e = 1 - torch.abs(p - t) + 1e-20
    return (-1 * (1 - e) ** gamma * torch.log(e)).sum() / (p.shape[0] * p.shape[1])


def bench_eager_dropin(args, local, steps=5):
    """The UNCHANGED trainer's step on the drop-in module (trainer/trainer.py:109-136): inputs H2D, zero_grad, model(inputs), every output
    moved to the CPU, FocalLoss_BCE_2d there, mean of three, backward (the upstream gradients travel back H2D), AdamW.step().
    Targets are precomputed (the trainer's numpy target synthesis, helper.create_heatmap, is host code outside the path).  Wall clock with
    a device synchronize per run — the loop is host-bound by construction.  Second figure: the same with the loss left on the device."""
    from unet_nested4tiny_objects_keypoints_b200 import optimizers
    dev = torch.device("cuda", local)
    res = {}
    for where in ("trainer_as_shipped_loss_on_cpu", "loss_on_device"):
        model = make_model(True, dev)
        opt = optimizers.AdamW(model.parameters(), lr=3e-6, weight_decay=1e-4)  # train.py:38-45 defaults
        g = torch.Generator().manual_seed(5)
        x_host = torch.randn(TRAIN_B1, 3, 256, 256, generator=g).pin_memory()
        t_host = torch.rand(TRAIN_B1, 4, 256, 256, generator=g)
        t_dev = t_host.to(dev)

        def one():
            x = x_host.to(dev, non_blocking=True)
            opt.zero_grad()
            outs = model(x)
            if where == "loss_on_device":
                loss = sum(focal_bce_2d(o, t_dev) for o in outs) / 3
            else:
                loss = sum(focal_bce_2d(o.cpu(), t_host) for o in outs) / 3
            loss.backward()
            opt.step()
            return loss

        for _ in range(2):
            one()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        res[where] = {"images_per_s": round(TRAIN_B1 / dt, 1), "ms_per_step": round(dt * 1e3, 2)}
        del model, opt
        torch.cuda.empty_cache()
    res["what"] = f"batch {TRAIN_B1}, 256x256, model(x) -> loss -> backward() -> optimizers.AdamW.step(), eager (no CUDA graph), wall clock, {steps} steps"
    return res


def bench_train(args, rank, world, local):
    from unet_nested4tiny_objects_keypoints_b200 import fused, ops
    dev = torch.device("cuda", local)
    model = make_model(True, dev)
    if world > 1:
        fused.broadcast_parameters(model)
    b = TRAIN_B1 if world == 1 else TRAIN_GLOBAL // world
    loss = getattr(args, "loss", "mse")  # BASELINE.json configs[2]/[3] name the MSE heat-map loss; "focal" = the trainer's shipped criterion
    step = fused.FusedTrainStep(model, b, 256, 256, device=dev, seed=0, loss=loss)
    g = torch.Generator().manual_seed(99 + rank)
    x_host = torch.randn(b, 3, 256, 256, generator=g).pin_memory()
    t_host = torch.rand(b, 4, 256, 256, generator=g).pin_memory()
    step.x.copy_(x_host)
    step.target.copy_(t_host)
    for _ in range(max(args.warmup, 3)):
        step.step_device()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    if world > 1:
        step.allreduce_events = []
    ms = timed(step.step_device, args.steps, world)
    clk = clocks.stop() if rank == 0 else None
    ar_us = None
    if world > 1:
        ar_us = statistics.median(e0.elapsed_time(e1) for e0, e1 in step.allreduce_events) * 1e3
        step.allreduce_events = None
    value = b * world * args.steps / (ms * 1e-3)
    # end to end through the public API (FusedTrainStep.step_many): every step its batch goes H2D from pinned host memory and its loss
    # comes back D2H; the copies of batch k+1 overlap the step of batch k.  Two forms of the same batch:
    #   e2e             what the trainer's data loader delivers (trainer.py:106-109): 8-bit RGB images (as the dataset decodes them,
    #                   datasets_base.py:71-72; ToTensor's 1/255 on the device) + the 7 key points per image; the targets of
    #                   helper.create_heatmap (numpy on the CPU every iteration in the reference, trainer.py:122-123) are synthesised on the device
    #   e2e_fp32_input  fp32 images + fp32 target heat maps built beforehand (58.7 MB per step at batch 32)
    loss_hosts = [torch.empty(1).pin_memory() for _ in range(2)]

    def e2e_of(st, xs, ts):
        def run(nsteps):
            st.step_many([xs[k & 1] for k in range(nsteps)], [ts[k & 1] for k in range(nsteps)], [loss_hosts[k & 1] for k in range(nsteps)])

        run(2)
        torch.cuda.synchronize()
        ms_e = timed(lambda: run(args.steps), 1, world)
        return {"value": round(b * world * args.steps / (ms_e * 1e-3), 1), "unit": "images/s",
                "h2d_bytes_per_step": xs[0].numel() * xs[0].element_size() + ts[0].numel() * ts[0].element_size(), "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e / args.steps, 3), "pipelining": "H2D of batch k+1 (staging buffers, copy stream) overlaps the step of batch k"}

    e2e_f32 = e2e_of(step, [x_host, torch.randn(b, 3, 256, 256, generator=g).pin_memory()], [t_host, torch.rand(b, 4, 256, 256, generator=g).pin_memory()])
    e2e_f32["input"] = "float32 images [B,3,H,W] + float32 target heat maps [B,4,H,W]"
    model8 = make_model(True, dev)
    if world > 1:
        fused.broadcast_parameters(model8)
    step8 = fused.FusedTrainStep(model8, b, 256, 256, device=dev, seed=0, loss=loss, input="uint8_nhwc")
    u8 = [torch.randint(0, 256, (b, 256, 256, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
    kps = [(torch.rand(b, 7, 2, generator=g) * 240 + 8).pin_memory() for _ in range(2)]
    e2e = e2e_of(step8, u8, kps)
    e2e["input"] = "uint8 RGB images [B,H,W,3] + 7 key points per image (the trainer's (inputs, labels), trainer.py:106-109); ToTensor and helper.create_heatmap on the device"
    del step8, model8
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
        step_ms = ms / args.steps
        times, eager_total, scale = scaled_times(agg, reps, step_ms)
        fwd_b, dgrad_b, wgrad_b, fwd_fl = plan_train(256)
        fam = {"conv_tc forward": [0.0, 0, fwd_b * b, fwd_fl * b], "conv_tc backward (dgrad)": [0.0, 0, dgrad_b * b, fwd_fl * b], "wgrad_tc": [0.0, 0, wgrad_b * b, fwd_fl * b],
               OTHER: [0.0, 0, 0, 0]}
        for (tag, label), t in times.items():
            k = ("conv_tc forward" if tag == "fwd" else "conv_tc backward (dgrad)") if label.startswith("conv_tc") else "wgrad_tc" if label.startswith("wgrad taps") else OTHER
            fam[k][0] += t
            fam[k][1] += agg[(tag, label)][1] // reps
        per_kernel = []
        for k, (t, n, nbytes, flops) in fam.items():
            t_b = max(nbytes / (pk["hbm"] * 1e9), flops / (pk["tc"] * 1e12)) * 1e3
            per_kernel.append({"family": k, "launches": n, "ms": round(t, 4), "algorithmic_GB": round(nbytes / 1e9, 3), "GBps": round(nbytes / (t * 1e-3) / 1e9, 1) if t and nbytes else None,
                               "frac": round(t_b / t, 3) if t and nbytes else None})
        t_conv, n_conv = fam["conv_tc forward"][0] + fam["conv_tc backward (dgrad)"][0], fam["conv_tc forward"][1] + fam["conv_tc backward (dgrad)"][1]
        conv_bytes = (fwd_b + dgrad_b) * b
        traffic, tsrc = traffic_of("train")
        roof = {"kernel": "conv_tc (forward + dgrad launches of one step)", "bound": "hbm", "achieved": round(conv_bytes / (t_conv * 1e-3) / 1e9, 1), "peak": pk["hbm"], "unit": "GB/s",
                "frac": round(conv_bytes / (t_conv * 1e-3) / 1e9 / pk["hbm"], 4), "traffic": traffic, "traffic_source": tsrc, "peak_source": pk["which"],
                "launches_per_step": n_conv, "ms_per_step_in_kernel": round(t_conv, 4), "share_of_step": round(t_conv / step_ms, 3),
                "algorithmic_bytes_per_launch_avg": round(conv_bytes / n_conv),
                "timing": f"CUDA-event share of each launch in an eager pass ({eager_total:.3f} ms serialised) x the graph step ({step_ms:.3f} ms): scale {scale:.3f}"}
        total_b = (fwd_b + dgrad_b + wgrad_b) * b
        whole = {"algorithmic_bytes_per_image": fwd_b + dgrad_b + wgrad_b, "achieved_gbs": round(total_b / (step_ms * 1e-3) / 1e9, 1),
                 "frac_of_hbm_peak": round(total_b / (step_ms * 1e-3) / 1e9 / pk["hbm"], 4),
                 "note": "enumerated per conv with the rule of SURVEY 8(d) (dgrad: dZ in, dX out; wgrad: dZ + saved input in; transposed convs and heads alike): "
                         "60.9 + 83.8 + 86.2 MB; SURVEY's own estimate is ~3x forward = 190 MB",
                 "frac_of_hbm_peak_at_190MB": round(190e6 * b / (step_ms * 1e-3) / 1e9 / pk["hbm"], 4)}
        line = {"metric": metric_name("train"), "value": round(value, 1), "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": round(step_ms, 4), "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic", "config": config_of("train", world), "clocks": clk, "e2e": e2e, "e2e_fp32_input": e2e_f32,
                "gpu_launches": launches * args.steps,
                "roofline": roof, "whole_step_roofline": whole, "per_kernel": per_kernel,
                "per_launch_label": [{"pass": tag, "label": label, "launches": agg[(tag, label)][1] // reps, "ms": round(t, 4)}
                                     for (tag, label), t in sorted(times.items(), key=lambda kv: -kv[1])],
                "loss": float(step.loss.item()), "loss_kind": loss,
                "batch_per_gpu": b, "images_per_s_per_gpu": round(value / world, 1), "allreduce_us": round(ar_us, 1) if ar_us is not None else None}
    return line


def set_workload(side, batch, tag):
    global S, INFER_B, CONFIG_TAG
    S, INFER_B, CONFIG_TAG = side, batch, tag


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
    ap.add_argument("--loss", default="mse", choices=["mse", "focal"], help="training criterion: BASELINE's MSE or the trainer's FocalLoss_BCE_2d")
    ap.add_argument("--no-extra", action="store_true", help="skip the cpu_baseline leg and the extra measurements (training step, 1024x1024, eager trainer loop)")
    args = ap.parse_args()
    if args.workload == "infer1024":  # BASELINE.json configs[4]: high-resolution inference, batch 16 at 1024x1024
        set_workload(1024, 16, "configs[4]")
        args.workload, args.no_extra = "infer", True

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")), out)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = dist_setup(args.gpus)
    line = bench_infer(args, rank, world, local) if args.workload == "infer" else bench_train(args, rank, world, local)
    extra = extra1024 = eager = None
    if args.workload == "infer" and not args.no_extra:
        # the other half of the metric ("infer & train step"): the training step at this world size
        targs = argparse.Namespace(**vars(args))
        targs.steps, targs.warmup = min(args.steps, 10), 3
        torch.cuda.empty_cache()
        t = bench_train(targs, rank, world, local)
        if t is not None:
            extra = {k: t[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "e2e", "e2e_fp32_input", "gpu_launches", "roofline", "whole_step_roofline", "per_kernel",
                                       "loss", "loss_kind", "batch_per_gpu", "images_per_s_per_gpu", "allreduce_us")}
        if world == 1:
            targs.loss = "focal"  # the criterion the trainer ships (trainer.py:426), next to BASELINE's MSE
            torch.cuda.empty_cache()
            tf = bench_train(targs, rank, world, local)
            extra["focal_criterion"] = {k: tf[k] for k in ("value", "ms_per_step", "loss", "loss_kind")}
            torch.cuda.empty_cache()
            eager = bench_eager_dropin(args, local)
            # BASELINE.json configs[4] in the same run (driver-visible): batch 16 at 1024x1024
            torch.cuda.empty_cache()
            set_workload(1024, 16, "configs[4]")
            t1024 = bench_infer(targs, rank, world, local)
            extra1024 = {k: t1024[k] for k in ("metric", "value", "unit", "ms_per_step", "config", "e2e", "e2e_fp32_input", "gpu_launches", "roofline", "whole_step_roofline",
                                               "other_kernels_ms")}
            set_workload(256, 128, "configs[1]")
    elif args.workload == "train" and not args.no_extra and world == 1:
        eager = bench_eager_dropin(args, local)
    if rank == 0:
        if world == 1 and not args.no_extra:
            train = args.workload == "train"
            # ~10-20 s of host work: four-image inference batches (or two-image training steps) of the same workload
            nb, bs = (25, 2) if train else (100, 4)
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
        if extra1024 is not None:
            line["infer1024"] = extra1024
        if eager is not None:
            line["eager_dropin"] = eager
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
