"""Turn the artefacts of scripts/capture_profiles.sh (gpurun_out/<tag>_*) into the tracked files under profiles/:
launch lists (ncu gpu__time_duration pass), per-launch counters of the full captures (incl. tensor-pipe activity), the DRAM-traffic
JSONs bench.py reads for roofline.traffic, the bench lines, and a markdown summary printed to stdout.
usage: python scripts/make_profiles.py <capture tag> <output tag>      e.g.  r02b r02"""
import collections, csv, json, os, shutil, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out_tag = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def launches(path, first):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", "")) / 1e3) for r in rows[h + 1:] if len(r) > vi and r[vi].replace(",", "").replace(".", "").isdigit()]
    seq = seq[max(i for i, (n, _) in enumerate(seq) if first in n):]  # the last (warm) pass, found by the kernel every pass starts with
    agg = collections.OrderedDict()
    for name, us in seq:
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += us
    return seq, agg


for kind, first in (("infer", "nchw_to_nhwc"), ("train", "pack_weights_batched")):
    src = os.path.join(G, f"{tag}_{kind}_launches.csv")
    shutil.copy(src, os.path.join(P, f"{out_tag}_{kind}_launches.csv"))
    seq, agg = launches(src, first)
    tot = sum(a[1] for a in agg.values())
    print(f"\n### {kind}: {len(seq)} launches, {tot:.0f} us serialised under ncu (last pass)\n\n| kernel | launches | time (us) | share |\n|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k[:70]}` | {a[0]} | {a[1]:.0f} | {100 * a[1] / tot:.1f} % |")


def full(name):
    rows = list(csv.reader(open(os.path.join(G, f"{tag}_{name}_ncu.csv"))))
    shutil.copy(os.path.join(G, f"{tag}_{name}_ncu.csv"), os.path.join(P, f"{out_tag}_{name}_ncu.csv"))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:] if len(r) == len(hdr)]


def num(d, key):
    return float(d[key].replace(",", ""))


TP = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"
for name, kind, what in (("conv_tc_infer", "infer", "the 23 conv_tc launches of one eager inference pass, B=128 256x256, three heads"),
                         ("conv_tc_train", "train", "the 57 conv_tc launches (26 forward + 31 dgrad) of one eager training step, B=32 256x256"),
                         ("wgrad_tc_train", None, "the 26 wgrad_tc launches of one eager training step, B=32 256x256")):
    data = full(name)
    rd = sum(num(d, "dram__bytes_read.sum") for d in data) * 1e6
    wr = sum(num(d, "dram__bytes_write.sum") for d in data) * 1e6
    us = sum(num(d, "gpu__time_duration.sum") for d in data)
    if kind:
        tj = {"kernel": "conv_tc", "launches": len(data), "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr, "traffic_bytes_per_launch_avg": (rd + wr) / len(data),
              "source": f"ncu --set full --clock-control none, {what} (profiles/{out_tag}_{name}_ncu.csv)"}
        json.dump(tj, open(os.path.join(P, f"{out_tag}_traffic_{kind}.json"), "w"), indent=1)
    print(f"\n### {name}: {len(data)} launches, {us:.0f} us, DRAM read {rd / 1e9:.2f} GB + write {wr / 1e9:.2f} GB = {(rd + wr) / 1e9:.2f} GB\n")
    print("| # | kernel | grid | regs | smem KB | us | DRAM rd MB | DRAM wr MB | DRAM GB/s | tensor pipe active % | tc operand fetch (smem wavefronts) % | issue active % |\n|---|---|---|---|---|---|---|---|---|---|---|---|")
    for i, d in enumerate(data):
        t = num(d, "gpu__time_duration.sum")
        r, w = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
        kn = d["Kernel Name"]
        kn = "conv_tc" + kn.split("conv_tc_kernel")[1].split("(")[0] if "conv_tc" in kn else "wgrad_tc"
        print("| %d | %s | %s | %s | %.0f | %.1f | %.0f | %.0f | %.0f | %.1f | %.1f | %.1f |" % (
            i, kn, d["Grid Size"], d["launch__registers_per_thread"], num(d, "launch__shared_mem_per_block_dynamic"), t, r, w, (r + w) / t * 1e3, num(d, TP),
            num(d, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), num(d, "smsp__issue_active.avg.pct_of_peak_sustained_active")))
for f in ("bench", "bench_train", "bench_reference_arm"):
    shutil.copy(os.path.join(G, f"{tag}_{f}.json"), os.path.join(P, f"{out_tag}_{f}.json"))
