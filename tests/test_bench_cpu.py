"""CPU checks of bench.py's bookkeeping (no GPU): the nominal-work model mirrors the pipeline's grouping, and the roofline
object is built from a recorded bench line without errors and with the contract's keys."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("sb_bench", os.path.join(ROOT, "bench.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        sys.argv = argv
    return m


def test_nominal_work_model():
    b = _bench()
    w = b.msm_nominal_work(20)
    # 2^20: every group has >= 2^21 entries: four pairwise rounds, then the XYZZ level on E / 16 points
    assert set(w) == {"k_affine_round<Fq>", "k_seg_accum_mixed<Fq>", "k_affine_round<Fq2>", "k_seg_accum_mixed<Fq2>"}
    e_g1 = b.msm_layout(1 << 20, 1 << 20)[1] * (1 << 20)
    assert abs(w["k_affine_round<Fq>"]["fq_products"] - e_g1 * (1 - 1 / 16) * 6) < 1
    assert abs(w["k_seg_accum_mixed<Fq>"]["fq_products"] - e_g1 / 16 * 10) < 1
    # window rule (mirror of msm_layout in csrc/msm.cu): log2(m) - 3 capped at 16 in a large group, log2(m) in a small one
    assert b.msm_layout(1 << 19, 1 << 19)[0] == 16 and b.msm_layout(1 << 18, 1 << 18)[0] == 15
    assert b.msm_layout(1 << 16, 1 << 16)[0] == 16 and b.msm_layout(1 << 10, 1 << 15)[0] == 10 and b.msm_layout(1 << 10, 1 << 18)[0] == 7
    # 2^17 per rank (8 GPUs): no group reaches 2^21 entries except none -> mixed additions only over G2
    w8 = b.msm_nominal_work(17, 8)
    assert "k_affine_round<Fq2>" not in w8 and "k_seg_accum_mixed<Fq2>" in w8


def test_roofline_object_from_a_recorded_line():
    b = _bench()
    rec = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_2e20_n1.json")).read().strip().splitlines()[-1])
    tl = [("k_affine_round<Fq2>", 0.0, 4.5, 18), ("k_affine_round<Fq2>", 5.0, 7.2, 18), ("k_sc_round<sc1,fold>", 8.0, 8.1, -1)]
    r = b.build_roofline(rec["kernels"], tl, b.msm_nominal_work(20), 30.2, 6544.0, "measured (MEASURED_PEAKS.json)", 20, 1)
    for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r
    assert r["kernel"] == "k_affine_round<Fq2>" and r["bound"] == "imad" and 0.3 < r["frac"] < 1.0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["hbm"]["bound"] == "hbm" and 0 < r["hbm"]["frac"] < 1
    ll = r["largest_launch"]
    assert ll["ms"] == 4.5 and ll["traffic"] == r["traffic"] and 1.0 < ll["traffic"] / ll["algorithmic_bytes"] < 1.5
    json.dumps(r)
    # sharded line: no ncu traffic is attached, nothing else changes shape
    r8 = b.build_roofline({"k_seg_accum_mixed<Fq2>": {"launches_per_step": 5, "ms_per_step": 2.1}}, [], b.msm_nominal_work(17, 8), 30.2, 6544.0, "x", 20, 8)
    assert r8["traffic"] is None and "largest_launch" not in r8
    assert b.build_roofline({}, [], b.msm_nominal_work(20), 30.2, 6544.0, "x", 20, 1) is None
