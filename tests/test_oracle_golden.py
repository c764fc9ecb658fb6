"""Pins the CPU oracle (oracle/) against golden vectors produced by the LIVE reference
(tests/golden/gen_golden.py).  Stage functions must agree exactly (or to float64 rounding where the
reference is numba fastmath); end-to-end flows agree to the reference's own rounding floor
(last-bit solver differences flip float32 roundings between levels)."""
import numpy as np
import pytest

from oracle import oracle as O
from conftest import epe_stats, rel_l2


def test_tap_tables(golden):
    g = golden("tables")
    nbad = ntot = 0
    for k, (il, ol, sg) in enumerate(g["cases"]):
        idx, wt = O.resize_tables(int(il), int(ol), float(sg))
        assert np.array_equal(idx, g[f"idx{k}"])
        ref = g[f"wt{k}"]
        # numba fastmath reassociates the Gaussian (x) cubic accumulation (host-SIMD dependent):
        # a handful of weights differ by 1 float32 ulp; never more.
        d = np.abs(wt.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))
        assert d.max() <= 1
        nbad += int((d > 0).sum())
        ntot += d.size
    assert nbad <= 1e-3 * ntot


def test_resize_bit_exact(golden):
    g = golden("resize")
    src = g["src"].astype(np.float64)
    for k, s in enumerate(g["sizes"]):
        out = O.resize(src, tuple(int(v) for v in s)).astype(np.float32)
        assert np.array_equal(out, g[f"out{k}"]), f"size {s}"


def test_warp_exact(golden):
    g = golden("warp")
    f2 = g["f2"].astype(np.float64)
    f1 = g["f1"].astype(np.float64)
    for meth in ("cubic", "linear"):
        out = O.imregister_wrapper(f2, g["u"], g["v"], g["w"], f1, meth)
        assert out.dtype == np.float32 and np.array_equal(out, g[meth])
    big = O.imregister_wrapper(f2, g["u"] * 8, g["v"] * 8, g["w"] * 8, f1, "cubic")
    assert np.array_equal(big, g["cubic_big"])
    raw = O.imregister_wrapper(g["raw_u16"], g["u"].astype(np.float32), g["v"].astype(np.float32),
                               g["w"].astype(np.float32), g["raw_ref"], "cubic")
    assert np.array_equal(raw, g["raw_cubic"])
    with pytest.raises(ValueError):
        O.imregister_wrapper(f2, g["u"], g["v"], g["w"], f1, "nearest")


def test_motion_tensor_exact(golden):
    g = golden("motion_tensor")
    f1 = g["f1"].astype(np.float64)
    # python floats, as the reference driver passes them: with a weak-scalar spacing numpy keeps the
    # float32 warp output in float32 through np.gradient / the second differences (a rounding point
    # the GPU kernel reproduces); an np.float64 spacing would silently promote to float64.
    h = tuple(float(x) for x in g["h"])
    J32 = np.stack(O.get_motion_tensor_gc(f1, g["f2"], *h), 0)
    J64 = np.stack(O.get_motion_tensor_gc(f1, g["f2"].astype(np.float64), *h), 0)
    assert np.array_equal(J32, g["J_f2f32"])
    assert np.array_equal(J64, g["J_f2f64"])


@pytest.mark.parametrize("name", ["c2", "c1s"])
def test_solver(golden, name):
    g = golden("solver")
    it, lag, a_smooth = g[f"{name}_params"]
    h = g[f"{name}_h"]
    out = O.compute_flow_3d(list(g[f"{name}_J"]), g[f"{name}_weight"], g[f"{name}_u"], g[f"{name}_v"],
                            g[f"{name}_w"], g[f"{name}_alpha"], int(it), int(lag), g[f"{name}_a_data"],
                            float(a_smooth), h[2], h[1], h[0])
    ref = g[f"{name}_out"]
    # numba fastmath vs plain C: float64 rounding only
    assert np.abs(out - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("run", ["ml0", "ml2w", "ml1s", "ml2w_uvw"])
def test_get_displacement_small(golden, run):
    g = golden("flow_small")
    kws = {
        "ml0": dict(alpha=(0.25,) * 3, update_lag=5, iterations=30, min_level=0, levels=100, eta=0.8,
                    a_smooth=1.0, a_data=0.45),
        "ml2w": dict(alpha=(0.25, 0.3, 0.2), update_lag=5, iterations=20, min_level=2, levels=100, eta=0.8,
                     a_smooth=1.0, a_data=0.45, weight=np.array([0.3, 0.7])),
        "ml1s": dict(alpha=(0.5,) * 3, update_lag=10, iterations=15, min_level=1, levels=100, eta=0.8,
                     a_smooth=0.5, a_data=0.45),
    }
    kw = dict(kws[run.replace("_uvw", "")])
    if run.endswith("_uvw"):
        kw["uvw"] = g["uvw"].copy()
    flow = O.get_displacement(g["fixed"], g["moving"], **kw)
    mean, mx = epe_stats(flow, g[f"flow_{run}"])
    # north-star tolerance is mean <= 0.01 / max <= 0.05; the oracle sits far inside it
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)


def test_preprocess(golden):
    g = golden("preprocess")
    ref64 = g["ref"].astype(np.float64)
    assert np.array_equal(O.preprocess(ref64, g["sigma"]), g["ref_proc"])
    assert np.array_equal(O.preprocess(g["batch"], g["sigma"], ref64), g["batch_proc"])
    assert np.array_equal(O.preprocess(g["batch"], g["sigma"], ref64, "separate"), g["batch_proc_sep"])
    assert np.array_equal(O.preprocess(g["batch_u16"], g["sigma"], ref64), g["batch_u16_proc"])


def test_sequence(golden):
    g = golden("sequence")
    ml, it, lag, buf = (int(v) for v in g["params"])
    reg, w = O.compensate_arr(g["video"], g["ref"], min_level=ml, iterations=it, update_lag=lag,
                              buffer_size=buf, weight=[0.5, 0.5])
    mean, mx = epe_stats(w, g["w"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)
    assert rel_l2(reg, g["registered"]) <= 1e-5


def test_config1_default_options(golden):
    """BASELINE config 1: 64x128x128 pair, OFOptions-default parameters, vs the live reference."""
    from tests_inputs import synth_volume
    g = golden("config1")
    V = synth_volume((64, 128, 128), 1)
    assert np.allclose([V.astype(np.float64).sum(), float(V[7, 11, 13])], g["fixed_checksum"], rtol=0, atol=1e-6)
    V64 = V.astype(np.float64)[..., None]
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]])
    fp = O.preprocess(V64, sigma)
    mp = O.preprocess(g["low_moving"].astype(np.float64)[..., None], sigma, V64)
    flow = O.get_displacement(fp, mp, alpha=(0.25,) * 3, levels=100, min_level=5, eta=0.8, update_lag=5,
                              iterations=100, a_smooth=1.0, a_data=0.45, weight=np.ones((64, 128, 128, 1)))
    f32 = flow.astype(np.float32)
    mean, mx = epe_stats(f32[::2, ::2, ::2], g["low_flow_s2"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)
    reg = O.imregister_wrapper(g["low_moving"], f32[..., 0], f32[..., 1], f32[..., 2], V64[..., 0], "cubic")
    assert rel_l2(reg[::2, ::2, ::2], g["low_reg_s2"]) <= 1e-5


def test_preprocess_temporal(golden):
    g = golden("preprocess_t")
    assert np.array_equal(O.preprocess(g["batch"], g["sigma"], g["ref"].astype(np.float64)), g["batch_proc"])


def test_sequence_update_reference(golden):
    g, gs = golden("sequence_update_ref"), golden("sequence")
    ml, it, lag, buf = (int(v) for v in g["params"])
    video, ref = gs["video"][:6, :12, :24, :28], gs["ref"][:12, :24, :28]
    reg, w = O.compensate_arr(video, ref, min_level=ml, iterations=it, update_lag=lag, buffer_size=buf,
                              weight=[0.5, 0.5], update_reference=True)
    mean, mx = epe_stats(w, g["w"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)
    assert rel_l2(reg, g["registered"]) <= 1e-5


def test_flow_nonlinear_smoothness(golden):
    g = golden("flow_small")
    kw = dict(alpha=(0.5,) * 3, update_lag=10, iterations=15, min_level=1, levels=100, eta=0.8, a_smooth=0.5, a_data=0.45)
    mean, mx = epe_stats(O.get_displacement(g["fixed"], g["moving"], **kw), g["flow_ml1s"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)


def test_oracle_motion_tensor_gray_and_cs(golden):
    """The oracle's restatements of the two unused constancy variants equal the live reference bit for bit."""
    g = golden("motion_tensor_alt")
    assert np.array_equal(np.stack(O.get_motion_tensor_gray(g["f1"], g["f2"], *g["h"]), 0), g["J_gray"])
    assert np.array_equal(np.stack(O.get_motion_tensor_cs(g["f1"], g["f2"], *g["h"]), 0), g["J_cs"])
