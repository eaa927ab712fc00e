import torch, time
x = torch.randn(128, 3, 256, 256).pin_memory()
d = torch.empty_like(x, device="cuda")
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: d.copy_(x, non_blocking=True))
print("single copy  %.3f ms  %.1f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
s = [torch.cuda.Stream() for _ in range(4)]
def multi(k):
    cur = torch.cuda.current_stream()
    ch = x.shape[0] // k
    for i in range(k):
        s[i].wait_stream(cur)
        with torch.cuda.stream(s[i]):
            d[i * ch:(i + 1) * ch].copy_(x[i * ch:(i + 1) * ch], non_blocking=True)
    for i in range(k): cur.wait_stream(s[i])
for k in (2, 4):
    ms = t(lambda: multi(k))
    print("%d streams    %.3f ms  %.1f GB/s" % (k, ms, x.numel() * 4 / ms / 1e6))
h = torch.randn(128, 3, 256, 256).half().pin_memory(); dh = torch.empty_like(h, device="cuda")
ms = t(lambda: dh.copy_(h, non_blocking=True)); print("fp16 copy    %.3f ms  %.1f GB/s" % (ms, h.numel() * 2 / ms / 1e6))
