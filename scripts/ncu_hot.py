"""Top SASS instructions by warp-stall samples from an `ncu --page source --print-source sass --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try:
        data.append((int(r[isamp]), n, r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
top = sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for s, n, r in sorted(top, key=lambda d: d[1]):
    st = sorted(((int(r[i] or 0), h) for i, h in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% #%-5d ex=%-8s %-70s %s" % (s, 100.0 * s / tot, n, r[iex], r[isrc].strip()[:70], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
