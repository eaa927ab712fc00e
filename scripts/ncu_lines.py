"""Executed instructions / stall samples per CUDA source line from `ncu -i X --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and len(r) > iex and r[0] != "":
        try:
            out.append((int(r[iex]), int(r[isamp]), cur, r[0], r[1]))
        except ValueError:
            pass
tot, stot = sum(o[0] for o in out), sum(o[1] for o in out)
print("warp-instructions executed %.2fM, samples %d" % (tot / 1e6, stot))
for ex, sm, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%7.2fM %5.1f%% samp=%4d %s:%s: %s" % (ex / 1e6, 100.0 * ex / tot, sm, f, ln, src.strip()[:105]))
