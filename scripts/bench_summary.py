import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("infer img/s", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "roof frac", d["roofline"]["frac"], "clocks", d["clocks"])
for k, v in d["per_kernel"].items():
    print("  ", k, v)
if "train_step" in d:
    t = d["train_step"]
    print("train img/s", t["value"], "ms", t["ms_per_step"], "conv_tc", t["roofline"]["ms_per_step_in_kernel"], "wgrad", t["roofline_wgrad"]["ms_per_step_in_kernel"], "launches", t["gpu_launches"])
print("cpu", d.get("cpu_baseline"))
