"""CPU: the roofline numerator of bench.py is the fused plan of SURVEY.md section 8(d), enumerated — pinned to the survey's own figures."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


def test_forward_plan_matches_survey_8d():
    rows = bench.plan_rows(256)
    assert len(rows) == 20
    total = sum(r[1] + r[2] for r in rows)
    assert abs((total + 553260 * 2) / 1e6 - 62.05) < 0.02          # 62.05 MB / image incl. the bf16 weights
    assert abs(sum(r[3] for r in rows) / 1e9 - 8.789) < 0.001       # 8.789 GFLOP / image (hook-counted)
    by = {r[0]: r for r in rows}
    mb = lambda v: round(v / 1e6, 2)
    # a few rows of the survey's per-kernel table (MB read / written)
    assert (mb(by["conv00.c1"][1]), mb(by["conv00.c1"][2])) == (0.39, 2.10)
    assert (mb(by["conv00.c2"][1]), mb(by["conv00.c2"][2])) == (2.10, 2.62)
    assert (mb(by["up01.c1"][1]), mb(by["up01.c1"][2])) == (3.15, 2.10)
    assert (mb(by["up12.c1"][1]), mb(by["up12.c1"][2])) == (2.62, 1.05)
    assert (mb(by["up03.c1"][1]), mb(by["up03.c1"][2])) == (7.34, 2.10)
    assert (mb(by["up03.c2"][1]), mb(by["up03.c2"][2])) == (2.10, 0.52)
    rows1024 = bench.plan_rows(1024)
    assert abs((sum(r[1] + r[2] for r in rows1024) + 553260 * 2) / 1e6 - 976.3) < 0.1
    assert abs(sum(r[3] for r in rows1024) / 1e9 - 140.63) < 0.01


def test_training_plan_enumeration():
    fwd, dgrad, wgrad, flops = bench.plan_train(256)
    assert fwd == sum(r[1] + r[2] for r in bench.plan_rows(256))
    # per conv: dgrad = dZ in + dX out, wgrad = dZ + saved input in; the first conv has no dgrad
    assert wgrad - dgrad == (16 + 3) * 65536 * 2
    assert 2.5 * fwd < dgrad + wgrad < 3.0 * fwd
