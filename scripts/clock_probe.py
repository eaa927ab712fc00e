"""Runs the inference graph back to back for a few seconds while logging nvidia-smi clocks/power."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import unet_nested4tiny_objects_keypoints_b200 as pkg
from unet_nested4tiny_objects_keypoints_b200 import fused
torch.manual_seed(0)
m = pkg.UNet_Nested().cuda().eval()
sess = fused.InferenceSession(m, 128, 256, 256)
sess.x.normal_()
q = "clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown"
p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader", "-lms", "100"], stdout=open("gpurun_out/clocks_probe.csv", "w"))
time.sleep(0.5)
for secs in (0.2, 3.0):
    torch.cuda.synchronize(); t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(10): sess.run_device()
        n += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    print(f"{secs}s loop: {e0.elapsed_time(e1)/n:.3f} ms/step over {n} steps")
time.sleep(0.3); p.terminate()
print(open("gpurun_out/clocks_probe.csv").read())
