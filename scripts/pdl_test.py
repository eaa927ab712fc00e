"""Does programmatic dependent launch overlap consecutive conv_tc launches?  Back-to-back launches, eager and captured."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_nested4tiny_objects_keypoints_b200 import ops
dev = "cuda"
N, H = 32, 256
src = torch.randn(N, H, H, 16, device=dev).to(torch.bfloat16)
wf = torch.randn(16, 16, 3, 3, device=dev) * 0.05
w, nt = ops.pack_weights_b2(wf, False, 16), ops.NTile(16, b2=True)
outs = [torch.empty(N, H, H, 16, dtype=torch.bfloat16, device=dev) for _ in range(2)]
bias = torch.randn(16, device=dev)
K = 40
def chain():
    a = src
    for i in range(K):
        ops.conv([a], N, H, H, w, 16, nt, 9, bias=bias, relu=True, out=outs[i & 1])
        a = outs[i & 1]
def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / K * 1e3
print("UNPP_PDL =", os.environ.get("UNPP_PDL"))
torch.cuda._sleep(int(1e8)); print("eager  per conv %.2f us" % timed(chain))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    chain()
print("graph  per conv %.2f us" % timed(g.replay))
