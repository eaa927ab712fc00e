"""profiles/<out>_summary.md from the files scripts/make_profiles.py left under profiles/ (bench lines, multi-GPU lines) and its stdout
(launch lists and per-launch ncu tables).  usage: python scripts/make_profiles.py <cap> <out> > /tmp/tables.md; python scripts/make_summary.py <out> /tmp/tables.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out, tables_path = sys.argv[1], sys.argv[2]
P = os.path.join(ROOT, "profiles")
d = json.load(open(os.path.join(P, f"{out}_bench.json")))
t, i = d["train_step"], d["infer1024"]
n = {k: json.load(open(os.path.join(P, f"{out}_bench_n{k}.json"))) for k in (2, 4, 8) if os.path.exists(os.path.join(P, f"{out}_bench_n{k}.json"))}
rows = "\n".join("| %s | %.1f | %s | %.3f | %s | %s |" % (r["row"], r["ms"] * 1e3, r["bound"], r["frac"], r["GBps"], r["TFLOPs"]) for r in d["per_kernel"])
fam = "\n".join("| %s | %d | %.3f | %.2f | %s | %s |" % (r["family"][:60], r["launches"], r["ms"], r["algorithmic_GB"], r["GBps"], r["frac"]) for r in t["per_kernel"])
tables = open(tables_path).read()
e = d["eager_dropin"]
md = f"""# Round 2 — state of the hot path (B200)

`scripts/capture_profiles.sh` on one `gpurun` box (every ncu pass after the same program had exited 0 without ncu), turned into the files of this
directory by `scripts/make_profiles.py` and `scripts/make_summary.py`:

* `{out}_bench.json` — the default `bench.py` line (inference B=128 256x256 with all three heads + arg-max; `train_step`, `infer1024`, `eager_dropin`,
  `cpu_baseline` inside); `{out}_bench_train.json` — `--workload train`; `{out}_bench_reference_arm.json` — `--impl reference`;
  `{out}_bench_n{{2,4,8}}.json` — the same default line under `torchrun` on 2 / 4 / 8 GPUs of one box;
* `{out}_{{infer,train}}_launches.csv` — `ncu --metrics gpu__time_duration.sum --clock-control none` launch lists;
* `{out}_conv_tc_infer_ncu.csv`, `{out}_conv_tc_train_ncu.csv`, `{out}_wgrad_tc_train_ncu.csv` — `ncu --set full --clock-control none` per-launch counters
  (selected columns: DRAM bytes, tensor-pipe activity, tensor-core operand fetch from shared memory, issue activity, launch geometry);
  `{out}_traffic_{{infer,train}}.json` = their DRAM totals, what `bench.py` reports as `roofline.traffic`;
* parity: `{out}_teacher_forced_tests.log`, `{out}_gpu_tests_1gpu.log`, `{out}_dp2_nccl_tests.log`, `{out}_robustness_dp_tests_2gpu.log` (two GPUs),
  `{out}_grad_errors_b32_256.md`, `{out}_grad_errors_b2_64.md`, `{out}_keypoint_agreement.md`.

## bench.py (CUDA-graph steps, CUDA events; SM clock {d['clocks']['sm_mhz']:.0f} MHz, throttle reasons seen: {d['clocks']['reasons']})

| workload | device-timed | end to end (every step H2D in, D2H out) | roofline |
|---|---|---|---|
| inference B=128 256x256, **three heads** + arg-max of 12 planes / image (configs[1]) | **{d['value']:.0f} img/s**, {d['ms_per_step']:.3f} ms | **{d['e2e']['value']:.0f} img/s** from 8-bit RGB images (25 MB / step); {d['e2e_fp32_input']['value']:.0f} img/s from fp32 tensors (100.7 MB / step; pinned copy measured at {d['e2e']['h2d_gbs_measured']} GB/s) | conv_tc {d['roofline']['achieved']:.0f} GB/s algorithmic = **{d['roofline']['frac']:.3f}** of 6 522.7; whole step {d['whole_step_roofline']['frac_of_per_row_roofline']:.3f} of the per-row roofline ({d['whole_step_roofline']['roofline_ms']:.3f} ms), {d['whole_step_roofline']['frac_of_hbm_peak']:.3f} of HBM peak; DRAM traffic {d['roofline']['traffic'] / 1e6:.0f} MB / launch vs {d['roofline']['algorithmic_bytes_per_launch_avg'] / 1e6:.0f} MB algorithmic |
| train step B=32, MSE + AdamW (configs[2]) | **{t['value']:.0f} img/s**, {t['ms_per_step']:.3f} ms (focal criterion: {t['focal_criterion']['value']:.0f} img/s) | {t['e2e']['value']:.0f} img/s from 8-bit images + key points (6.3 MB / step, targets synthesised on the device); {t['e2e_fp32_input']['value']:.0f} img/s from fp32 images + fp32 targets (58.7 MB / step) | conv_tc {t['roofline']['frac']:.3f} (DRAM traffic {t['roofline']['traffic'] / 1e6:.0f} MB / launch vs {t['roofline']['algorithmic_bytes_per_launch_avg'] / 1e6:.0f} MB algorithmic); whole step {t['whole_step_roofline']['frac_of_hbm_peak']:.3f} of HBM peak by the enumerated 230.9 MB / image ({t['whole_step_roofline']['frac_of_hbm_peak_at_190MB']:.3f} by SURVEY's 190 MB estimate) |
| inference B=16 1024x1024, three heads (configs[4]) | **{i['value']:.0f} img/s**, {i['ms_per_step']:.3f} ms | {i['e2e']['value']:.0f} img/s | whole step {i['whole_step_roofline']['frac_of_per_row_roofline']:.3f} of the per-row roofline; split arg-max {i['other_kernels_ms']['argmax_peaks_split'] * 1e3:.0f} us for 805 MB |
| unchanged trainer loop on the drop-in (eager: `model(x)` -> loss -> `backward()` -> `optimizers.AdamW.step()`), B=32 | {e['loss_on_device']['images_per_s']:.0f} img/s with the loss on the device | {e['trainer_as_shipped_loss_on_cpu']['images_per_s']:.0f} img/s as shipped (3 x D2H of the outputs, FocalLoss on the CPU, H2D of the gradients: trainer.py:127-135) | |
| reference arm (oracle port, torch-CPU fp32, {d['cpu_baseline']['cores']} host cores) | {d['cpu_baseline']['value']:.1f} img/s (batch 1: {d['cpu_baseline']['configs0_batch1_images_per_s']:.1f}) | | |

Round 1 measured one head (61.7 k img/s, 2.075 ms); the two other heads add two fp32 heat-map writes and 8 more planes to the arg-max (+0.13 ms).
"""
if n:
    md += f"""
### Multi-GPU (`torchrun`, one rank per GPU, same box)

| GPUs | inference img/s (device) | e2e from 8-bit images | e2e from fp32 tensors (pinned-copy GB/s per GPU) | DP train img/s, global batch 256 | per GPU (batch) | all-reduce us | DP e2e (8-bit images + key points / fp32 images + targets) |
|---|---|---|---|---|---|---|---|
| 1 | {d['value']:.0f} | {d['e2e']['value']:.0f} | {d['e2e_fp32_input']['value']:.0f} ({d['e2e']['h2d_gbs_measured']}) | {t['value']:.0f} (batch 32) | {t['value']:.0f} (32) | - | {t['e2e']['value']:.0f} / {t['e2e_fp32_input']['value']:.0f} |
""" + "\n".join(f"| {k} | {n[k]['value']:.0f} | {n[k]['e2e']['value']:.0f} | {n[k]['e2e_fp32_input']['value']:.0f} ({n[k]['e2e']['h2d_gbs_measured']}) | {n[k]['train_step']['value']:.0f} | {n[k]['train_step']['images_per_s_per_gpu']:.0f} ({n[k]['train_step']['batch_per_gpu']}) | {n[k]['train_step']['allreduce_us']} | {n[k]['train_step']['e2e']['value']:.0f} / {n[k]['train_step']['e2e_fp32_input']['value']:.0f} |" for k in sorted(n)) + f"""

End-to-end inference from 8-bit images scales {n[8]['e2e']['value'] / d['e2e']['value']:.2f}x on 8 GPUs ({n[8]['e2e']['value'] / d['e2e']['value'] / 8:.3f} of linear); from fp32 tensors it is bound by the host: the
pinned-copy bandwidth per GPU drops from {d['e2e']['h2d_gbs_measured']} GB/s (1 GPU) to {n[8]['e2e']['h2d_gbs_measured']} GB/s (8 GPUs copying at once, ~{8 * n[8]['e2e']['h2d_gbs_measured']:.0f} GB/s for the box).  Data-parallel training at
8 x 32 runs at {n[8]['train_step']['images_per_s_per_gpu'] / t['value']:.3f} of the single-GPU batch-32 step per GPU; the all-reduce of the 2.2 MB flat gradient is {n[8]['train_step']['allreduce_us']} us of a {n[8]['train_step']['ms_per_step']:.2f} ms step.
"""
md += f"""
### Inference, per row of the fused plan (SURVEY 8(d) bytes / flops; time = launch share x graph step)

| row | us | bound | fraction of its roofline | algorithmic GB/s | TFLOP/s |
|---|---|---|---|---|---|
{rows}

other launches of the step (ms): {d['other_kernels_ms']}

### Training step, per kernel family

| family | launches | ms | algorithmic GB | GB/s | fraction of the HBM roofline |
|---|---|---|---|---|---|
{fam}

## ncu launch lists and per-launch counters
{tables}
"""
open(os.path.join(P, f"{out}_summary.md"), "w").write(md)
print("wrote", os.path.join(P, f"{out}_summary.md"))
