"""Round-2 additions to the golden fixtures (tests/golden/unetpp_r2.*).  TEST INFRASTRUCTURE ONLY — run in the build
container where the reference is mounted at /root/reference:

    python oracle/make_golden_r2.py

Pins against the REAL reference (its own code, executed here):
  * helper.create_heatmap (tools/misc/helper.py:87-172) with SIX key points: plane 3 then holds one point and is still divided by
    its maximum (helper.py:148-159) — oracle.create_heatmap;
  * Heatmap.extract_points_(plane, num) (tools/misc/heatmap.py:148-208, OpenCV watershed inside) on multi-point target planes
    drawn by the reference's own Heatmap.create_heatmap — oracle.topk_peaks must return the same points in the same order.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("UNPP_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from oracle import unetpp_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    warnings.filterwarnings("ignore")
    import numpy.matlib  # noqa: F401  (helper.py uses np.matlib without importing it explicitly on numpy 2)
    from tools.misc import helper as RH
    from tools.misc.heatmap import Heatmap
    rng = np.random.RandomState(20)
    arrays, meta = {}, {}

    # ---- six key points
    kp6 = (rng.rand(3, 6, 2) * 36 + 2).astype(np.float32)
    r6 = RH.create_heatmap(kp6, 40, 40)
    o6 = O.create_heatmap(kp6, 40, 40)
    assert np.abs(r6 - o6).max() <= 1e-6, np.abs(r6 - o6).max()
    assert np.allclose(r6[:, 3].reshape(3, -1).max(1), 1.0)
    arrays["hm6_keypoints"], arrays["hm6_out"] = kp6, r6

    # ---- multi-point planes: pattern of the trainer (trainer.py:297-298 builds Heatmap from the dataset's pattern)
    H = W = 64
    pattern = [[0], [1, 2, 3], [4], [5, 6]]
    hm = Heatmap(pattern, W, H)
    N = 6
    pts = np.zeros((N, 7, 2), dtype=np.float32)
    for n in range(N):  # well separated integer-free positions (>= 14 px apart): blobs of radius 3 do not touch above the 0.5 threshold
        grid = [(x, y) for x in range(8, W - 7, 16) for y in range(8, H - 7, 16)]
        sel = rng.permutation(len(grid))[:7]
        for k, g in enumerate(sel):
            pts[n, k] = (grid[g][0] + rng.rand() * 3 - 1.5, grid[g][1] + rng.rand() * 3 - 1.5)
    planes = hm.create_heatmap(pts)  # [N, 4, H, W] float32, each plane divided by its maximum
    # a network's output is not a clean target: scale the blobs differently so that the "brightest first" order is exercised
    planes = planes * (0.75 + 0.25 * rng.rand(N, 4, 1, 1)).astype(np.float32)
    ref_pts = []
    for n in range(N):
        ref_pts.append([])
        for c, grp in enumerate(pattern):
            ref_pts[-1].append(hm.extract_points_(planes[n, c], len(grp)))
    merged = 0
    for c, grp in enumerate(pattern):
        oxy, oval, ocnt = O.topk_peaks(planes[:, c:c + 1], len(grp))
        for n in range(N):
            got = [list(map(int, p)) for p in oxy[n, 0][:ocnt[n, 0]]]
            assert len(got) == len(grp), (n, c, got)  # every drawn point is found
            if len(ref_pts[n][c]) == len(grp):
                assert got == ref_pts[n][c], (n, c, got, ref_pts[n][c])
            else:
                # the reference's watershed sometimes merges two separate blobs into one region (e.g. two points 15 px apart on one image
                # column) and then reports fewer points than were drawn; what it does report is a subsequence of ours, in the same order
                merged += 1
                it = iter(got)
                assert all(p in it for p in ref_pts[n][c]), (n, c, got, ref_pts[n][c])
    meta["peaks_planes_where_the_reference_merged_blobs"] = merged
    arrays["peaks_planes"] = planes
    meta["peaks_pattern"] = pattern
    meta["peaks_ref"] = ref_pts
    # the retry path (heatmap.py:187-190): nothing reaches 0.5, a disc of 0.47 with a 0.49 centre survives the retry at 0.45
    yy, xx = np.mgrid[0:H, 0:W]
    low = np.zeros((1, 1, H, W), dtype=np.float32)
    low[0, 0][(yy - 30) ** 2 + (xx - 41) ** 2 <= 36] = 0.47
    low[0, 0, 30, 41] = 0.49
    r_low = hm.extract_points_(low[0, 0], 1)
    oxy, _, ocnt = O.topk_peaks(low, 1)
    assert [list(map(int, oxy[0, 0, 0]))] == r_low and ocnt[0, 0] == 1, (r_low, oxy)
    arrays["peaks_low"] = low
    meta["peaks_low_ref"] = r_low
    np.savez_compressed(os.path.join(GOLD, "unetpp_r2.npz"), **arrays)
    with open(os.path.join(GOLD, "unetpp_r2.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", os.path.join(GOLD, "unetpp_r2.npz"), {k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
