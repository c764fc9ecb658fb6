"""Stage-wise parity of the CUDA path (through the C ABI) against golden vectors of the live
reference and against the oracle.  Tolerances are written next to each assertion:
bit-exact for the resize, warp, median and pre-filter (float32 outputs / order statistics),
float64 rounding for the motion tensor and the solver."""
import numpy as np
import pytest
from scipy.ndimage import median_filter

from conftest import ulp_diff
from oracle import oracle as O

INNER = (slice(1, -1),) * 3


def test_tap_tables_match_oracle_and_reference(backend, golden):
    from flowreg3d_b200 import plan
    g = golden("tables")
    for k, (il, ol, sg) in enumerate(g["cases"]):
        R, gt = plan.gauss_taps(float(sg))
        idx = np.empty((int(ol), 2 * R + 4), np.int32)
        wt = np.empty((int(ol), 2 * R + 4), np.float32)
        from flowreg3d_b200 import _lib
        assert _lib.load().fr3d_fill_resize_table(int(il), int(ol), gt.ctypes.data, R, idx.ctypes.data,
                                                  wt.ctypes.data) == 0
        oi, ow = O.resize_tables(int(il), int(ol), float(sg))
        assert np.array_equal(idx, oi) and np.array_equal(wt, ow)          # product == oracle, bit for bit
        assert np.array_equal(idx, g[f"idx{k}"])                            # indices == reference
        assert ulp_diff(wt, g[f"wt{k}"]).max() <= 1                         # weights within 1 float32 ulp


def test_resize_bit_exact(backend, golden):
    from flowreg3d_b200 import core
    g = golden("resize")
    src = g["src"].astype(np.float64)
    for k, s in enumerate(g["sizes"]):
        out = core.resize(src, tuple(int(v) for v in s)).astype(np.float32)
        assert np.array_equal(out, g[f"out{k}"]), f"size {s}"


def test_resize_edge_shapes(backend):
    """ragged / tiny / single-plane inputs and identity resize vs the oracle."""
    from flowreg3d_b200 import core
    rng = np.random.default_rng(3)
    for shp, size in [((1, 7, 9), (1, 5, 4)), ((5, 6, 7), (5, 6, 7)), ((3, 4, 5), (9, 11, 13)),
                      ((17, 3, 31), (6, 2, 12)), ((2, 2, 2), (4, 1, 3)),
                      ((4, 5, 130), (6, 7, 100)), ((3, 4, 40), (5, 9, 150)), ((3, 6, 200), (2, 5, 129))]:  # wide rows: 4-output runs
        a = rng.random(shp).astype(np.float32)
        assert np.array_equal(core.resize(a, size), O.resize(a, size)), (shp, size)
    a = rng.random((6, 7, 8)).astype(np.float32)
    assert np.array_equal(core.resize(a, a.shape), a)  # scale 1 is an exact copy


@pytest.mark.parametrize("meth", ["cubic", "linear"])
def test_warp_bit_exact(backend, golden, meth):
    import flowreg3d_b200 as F
    g = golden("warp")
    out = F.imregister_wrapper(g["f2"].astype(np.float64), g["u"], g["v"], g["w"], g["f1"].astype(np.float64), meth)
    assert out.dtype == np.float32
    d = ulp_diff(out, g[meth])
    # float32 result of float64 spline math; the CUDA pow() in the prefilter boundary term may differ
    # from libm in the last bit -> allow 1 ulp on a vanishing fraction of voxels
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4


def test_warp_out_of_volume_and_integer_input(backend, golden):
    import flowreg3d_b200 as F
    g = golden("warp")
    f2, f1 = g["f2"].astype(np.float64), g["f1"].astype(np.float64)
    big = F.imregister_wrapper(f2, g["u"] * 8, g["v"] * 8, g["w"] * 8, f1, "cubic")
    d = ulp_diff(big, g["cubic_big"])
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4
    raw = F.imregister_wrapper(g["raw_u16"], g["u"].astype(np.float32), g["v"].astype(np.float32),
                               g["w"].astype(np.float32), g["raw_ref"], "cubic")
    assert np.array_equal(raw, g["raw_cubic"])  # integer source: scipy rounds into uint16, exact
    with pytest.raises(ValueError):
        F.imregister_wrapper(f2, g["u"], g["v"], g["w"], f1, "nearest")
    # single-channel input: channel axis squeezed like the reference (:72-73)
    one = F.imregister_wrapper(f2[..., 0], g["u"], g["v"], g["w"], f1[..., 0], "linear")
    assert one.shape == f2.shape[:3]
    assert np.array_equal(one, g["linear"][..., 0])


def test_warp_long_lines_segmented_prefilter(backend):
    """Lines long enough for the tiled prefilter to cut them into warm-started segments (Y: 3, X: 8
    segments) -- same bit-level agreement with scipy as the short golden case."""
    import flowreg3d_b200 as F
    from tests_inputs import smooth_flow
    rng = np.random.default_rng(11)
    shp = (5, 150, 420)
    f2 = rng.random(shp + (2,))
    f2[:, :, :3] *= 50.0                                       # strong edge at a line start
    f1 = rng.random(shp + (2,))
    g = smooth_flow(shp, 5, 3.0, 4.0).astype(np.float64)
    out = F.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic")
    ref = O.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic")
    d = ulp_diff(out, ref)
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4, (d.max(), (d > 0).mean())


@pytest.mark.parametrize("shape,C", [((6, 9, 13), 1), ((5, 10, 16), 2), ((7, 8, 11), 3)])
def test_warp_rough_flow(backend, shape, C):
    """Rough (per-voxel random) displacements: neighbouring outputs gather from unrelated tap boxes, odd X, large
    displacements (out-of-volume replacement, clipped taps), 1-3 channels and an integer source; every case must
    equal scipy (through the oracle)."""
    import flowreg3d_b200 as F
    rng = np.random.default_rng(shape[2] + C)
    f2 = rng.random(shape + (C,))
    f1 = rng.random(shape + (C,))
    for scale in (0.4, 1.5, 6.0):
        u, v, w = (scale * rng.standard_normal(shape) for _ in range(3))
        out = F.imregister_wrapper(f2, u, v, w, f1, "cubic")
        ref = O.imregister_wrapper(f2, u, v, w, f1, "cubic")
        assert out.shape == ref.shape
        d = ulp_diff(out, ref)
        assert d.max() <= 1 and (d > 0).mean() <= 2e-3, (scale, d.max(), (d > 0).mean())
    raw = rng.integers(0, 60000, shape + (C,)).astype(np.uint16)
    rref = rng.integers(0, 60000, shape + (C,)).astype(np.uint16)
    u, v, w = (1.5 * rng.standard_normal(shape).astype(np.float32) for _ in range(3))
    assert np.array_equal(F.imregister_wrapper(raw, u, v, w, rref, "cubic"),
                          O.imregister_wrapper(raw, u, v, w, rref, "cubic"))


def test_motion_tensor(backend, golden):
    from flowreg3d_b200 import core
    g = golden("motion_tensor")
    h = [float(x) for x in g["h"]]
    f1 = g["f1"].astype(np.float64)
    for key, f2 in (("J_f2f32", g["f2"]), ("J_f2f64", g["f2"].astype(np.float64))):
        J = core.motion_tensor(f1, f2, *h)
        ref = g[key][(slice(None),) + INNER]
        assert np.array_equal(g[key][:, 0], np.zeros_like(g[key][:, 0]))  # the reference ring is zero
        # same float32/float64 rounding points as numpy -> agreement to float64 rounding
        assert np.abs(J - ref).max() <= 1e-12 * np.abs(ref).max()


def _solver_case(g, name):
    J = np.moveaxis(g[f"{name}_J"][(slice(None),) + INNER], -1, 0)
    wgt = np.moveaxis(g[f"{name}_weight"][INNER], -1, 0)
    uvw = np.stack([g[f"{name}_{c}"][INNER] for c in "uvw"], 0)
    ref = np.moveaxis(g[f"{name}_out"][INNER], -1, 0)
    it, lag, a_smooth = g[f"{name}_params"]
    return J, wgt, uvw, ref, int(it), int(lag), float(a_smooth)


def test_solver_wavefront_reproduces_lexicographic_order(backend, golden):
    from flowreg3d_b200 import core
    g = golden("solver")
    J, wgt, uvw, ref, it, lag, _ = _solver_case(g, "c2")
    d = core.sor_level(J, wgt, uvw, g["c2_alpha"], g["c2_h"], it, lag, g["c2_a_data"])
    # reference = numba fastmath float64; ours = float64 without contraction, different association
    assert np.abs(d - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    # iteration counts that are not multiples of the lag, lag 1, single iteration
    for it2, lag2 in ((1, 5), (7, 1), (4, 3)):
        d = core.sor_level(J, wgt, uvw, g["c2_alpha"], g["c2_h"], it2, lag2, g["c2_a_data"])
        Jr = [np.pad(np.moveaxis(J[:, q], 0, -1), ((1, 1), (1, 1), (1, 1), (0, 0))) for q in range(10)]
        o = O.compute_flow_3d(Jr, g["c2_weight"], g["c2_u"], g["c2_v"], g["c2_w"], g["c2_alpha"], it2, lag2,
                              g["c2_a_data"], 1.0, g["c2_h"][2], g["c2_h"][1], g["c2_h"][0])
        assert np.abs(d - np.moveaxis(o[INNER], -1, 0)).max() <= 1e-10


def test_solver_item_scheduling_options_are_bit_identical(backend, golden):
    """FR3D_OPT_SOR_SCHED only changes which warp runs which work item of a wave (psi-refresh items dealt first and
    evenly = bit 7; the last percent of a wave through a ticket counter = bits 0-6): same items, same arithmetic."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import _lib, core, device as dev
    g = golden("solver")
    J, wgt, uvw, ref, it, lag, _ = _solver_case(g, "c2")
    ctx = core.bare_context()

    def with_sched(value, fn):
        core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_SOR_SCHED, value))
        try:
            return fn()
        finally:
            core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_SOR_SCHED, -1))

    for it2, lag2 in ((it, lag), (1, 5), (7, 1), (4, 3), (23, 2), (11, 50)):
        solve = lambda: core.sor_level(J, wgt, uvw, g["c2_alpha"], g["c2_h"], it2, lag2, g["c2_a_data"])  # noqa: E731
        base = with_sched(0, solve)
        for sched in (128, 128 + 20, 20, 100, 128 + 100):
            assert np.array_equal(with_sched(sched, solve), base), (it2, lag2, sched)
    # several frames per launch (frame pairs per work item, an odd frame left over), whole pyramid
    gs = golden("flow_small")
    fixed, moving = gs["fixed"].astype(np.float32), gs["moving"].astype(np.float32)
    mv = np.stack([moving, np.roll(moving, 1, 2), np.roll(moving, -2, 1)], 0)
    fp = F.FlowParams(alpha=(0.25, 0.3, 0.2), update_lag=5, iterations=12, min_level=1, levels=100, eta=0.8,
                      a_smooth=1.0, a_data=0.45)
    flows = []
    for sched in (0, 128 + 20):
        reg = F.Registration(fixed.shape[:3], fixed.shape[3], fp, max_batch=3)
        core._check(reg.ctx.h, reg.ctx.lib.fr3d_set_option(reg.ctx.h, _lib.OPT_SOR_SCHED, sched))
        reg.set_reference(fixed)
        flows.append(dev.to_host(reg.get_displacement(mv)).copy())
        reg.sync()
    assert np.array_equal(flows[0], flows[1])
    if backend == "emu":            # the emulator reports which code paths ran: the balanced order must have been one
        ctx.profile(True)
        with_sched(128, lambda: core.sor_level(J, wgt, uvw, g["c2_alpha"], g["c2_h"], 3, 2, g["c2_a_data"]))
        assert "fr3d_sor_wavefront_balanced" in ctx.profile_report()
        ctx.profile(False)


def test_solver_redblack_mode_matches_redblack_restatement(backend, golden):
    """The opt-in checkerboard sweep equals a CPU restatement with the same (non-reference) order, and
    differs from the lexicographic result (which is why it is not the default)."""
    from flowreg3d_b200 import core
    from flowreg3d_b200.plan import SWEEP_REDBLACK
    g = golden("solver")
    J, wgt, uvw, ref, it, lag, _ = _solver_case(g, "c2")
    d = core.sor_level(J, wgt, uvw, g["c2_alpha"], g["c2_h"], it, lag, g["c2_a_data"], sweep=SWEEP_REDBLACK)
    Jr = [np.pad(np.moveaxis(J[:, q], 0, -1), ((1, 1), (1, 1), (1, 1), (0, 0))) for q in range(10)]
    O.set_sweep_order(1)
    try:
        o = O.compute_flow_3d(Jr, g["c2_weight"], g["c2_u"], g["c2_v"], g["c2_w"], g["c2_alpha"], it, lag,
                              g["c2_a_data"], 1.0, g["c2_h"][2], g["c2_h"][1], g["c2_h"][0])
    finally:
        O.set_sweep_order(0)
    assert np.abs(d - np.moveaxis(o[INNER], -1, 0)).max() <= 1e-10
    assert np.abs(d - ref).max() > 1e-6      # a different iteration, not the reference's


def test_solver_nonlinear_smoothness(backend, golden):
    """a_smooth != 1: psi_s recomputed every sweep (with the reference's stale ring), golden from the live
    reference; plus lag / iteration edge cases against the oracle."""
    from flowreg3d_b200 import core
    g = golden("solver")
    J, wgt, uvw, ref, it, lag, a_smooth = _solver_case(g, "c1s")
    assert a_smooth != 1.0
    d = core.sor_level(J, wgt, uvw, g["c1s_alpha"], g["c1s_h"], it, lag, g["c1s_a_data"], a_smooth=a_smooth)
    assert np.abs(d - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())
    Jr = [np.pad(np.moveaxis(J[:, q], 0, -1), ((1, 1), (1, 1), (1, 1), (0, 0))) for q in range(10)]
    for it2, lag2, a2 in ((1, 5, 0.5), (2, 1, 0.7), (9, 4, 0.5)):
        d = core.sor_level(J, wgt, uvw, g["c1s_alpha"], g["c1s_h"], it2, lag2, g["c1s_a_data"], a_smooth=a2)
        o = O.compute_flow_3d(Jr, g["c1s_weight"], g["c1s_u"], g["c1s_v"], g["c1s_w"], g["c1s_alpha"], it2, lag2,
                              g["c1s_a_data"], a2, g["c1s_h"][2], g["c1s_h"][1], g["c1s_h"][0])
        assert np.abs(d - np.moveaxis(o[INNER], -1, 0)).max() <= 1e-10, (it2, lag2, a2)


def test_median_exact(backend):
    from flowreg3d_b200 import core
    rng = np.random.default_rng(0)
    v = rng.standard_normal((3, 9, 13, 17))
    v[1] = np.round(v[1], 1)                                  # heavy ties
    v[2] = 1.0 + rng.integers(0, 50, v[2].shape) * 1e-12      # distinct in float64, tied in float32
    out = core.median5(v)
    ref = np.stack([median_filter(x, size=(5, 5, 5), mode="mirror") for x in v])
    assert np.array_equal(out, ref)
    small = rng.standard_normal((1, 6, 6, 7))                 # smallest grid the driver filters (min > 5)
    assert np.array_equal(core.median5(small)[0], median_filter(small[0], size=(5, 5, 5), mode="mirror"))


def test_preprocess_bit_exact(backend, golden):
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    from flowreg3d_b200.compensate import normalization_range
    g = golden("preprocess")
    ref, batch, sigma = g["ref"], g["batch"], g["sigma"]
    Z, Y, X, C = ref.shape
    reg = F.Registration((Z, Y, X), C, F.FlowParams(min_level=1, a_smooth=1.0), max_batch=3, sigma=sigma)
    r64 = ref.astype(np.float64)
    lo, den = normalization_range(r64, "joint")
    f32 = lambda a: np.asarray(a).astype(np.float32)
    assert np.array_equal(dev.to_host(reg.preprocess(ref[None], lo, den))[0], f32(g["ref_proc"]))
    assert np.array_equal(dev.to_host(reg.preprocess(batch, lo, den)), f32(g["batch_proc"]))
    assert np.array_equal(dev.to_host(reg.preprocess(g["batch_u16"], lo, den)), f32(g["batch_u16_proc"]))
    lo, den = normalization_range(r64, "separate")
    assert np.array_equal(dev.to_host(reg.preprocess(batch, lo, den)), f32(g["batch_proc_sep"]))
    reg.ctx.close()


def test_preprocess_temporal_filter(backend, golden):
    """sigma_t >= 0.125: the frames of a batch are filtered across time first (4-D filter of the reference),
    reflect at the batch ends; the fixed volume (4-D input) is not."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    from flowreg3d_b200.compensate import normalization_range
    g = golden("preprocess_t")
    ref, batch, sigma = g["ref"], g["batch"], g["sigma"]
    Z, Y, X, C = ref.shape
    reg = F.Registration((Z, Y, X), C, F.FlowParams(min_level=1, a_smooth=1.0), max_batch=5, sigma=sigma)
    assert reg.plan.temporal
    lo, den = normalization_range(ref.astype(np.float64), "joint")
    got = dev.to_host(reg.preprocess(batch, lo, den))
    assert np.array_equal(got, g["batch_proc"].astype(np.float32))
    one = dev.to_host(reg.preprocess(batch[:1], lo, den, temporal=False))
    spatial_only = O.preprocess(batch[:1], np.concatenate([sigma[:, :3], np.full((2, 1), 0.1)], 1), ref.astype(np.float64))
    assert np.array_equal(one, spatial_only.astype(np.float32))
    reg.ctx.close()


def test_warp_factored_option_within_one_ulp(backend):
    """FR3D_OPT_WARP_FACTORED (1 = default since round 2: 7.61 -> 6.18 ms per 16 config-2 frames on a B200; the only
    knob that may change results): the factored separable sum agrees with scipy's ((c*wz)*wy)*wx association
    (option 0) to <= 1 float32 ulp on a vanishing fraction of the voxels; integer sources stay exact."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import _lib, core
    from tests_inputs import smooth_flow
    rng = np.random.default_rng(2)
    shp = (9, 40, 64)
    f2, f1 = rng.random(shp + (2,)), rng.random(shp + (2,))
    g = smooth_flow(shp, 5, 3.0, 4.0).astype(np.float64)
    raw = rng.integers(0, 60000, shp + (1,)).astype(np.uint16)
    ctx = core.bare_context()
    core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_WARP_FACTORED, 0))
    base = F.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic")
    assert np.array_equal(base, O.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic")) or \
        ulp_diff(base, O.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic")).max() <= 1
    core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_WARP_FACTORED, 1))
    ctx.profile(True)
    try:
        fact = F.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic")
        ran = list(ctx.profile_report())
        assert any("WarpGatherLeanK" in k and ("Li1EEE" in k or "0, 1>" in k) for k in ran), ran   # the factored functor
        fraw = F.imregister_wrapper(raw, g[..., 0].astype(np.float32), g[..., 1].astype(np.float32),
                                    g[..., 2].astype(np.float32), raw, "cubic")
    finally:
        ctx.profile(False)
    d = ulp_diff(fact, base)
    assert d.max() <= 1 and (d > 0).mean() <= 1e-4, (d.max(), (d > 0).mean())
    assert np.array_equal(fraw, O.imregister_wrapper(raw, g[..., 0].astype(np.float32), g[..., 1].astype(np.float32),
                                                     g[..., 2].astype(np.float32), raw, "cubic"))
    core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_WARP_FACTORED, 0))
    try:
        assert np.array_equal(F.imregister_wrapper(f2, g[..., 0], g[..., 1], g[..., 2], f1, "cubic"), base)  # off again
    finally:
        core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_WARP_FACTORED, 1))


def test_slab_restricted_sweeps(backend, golden):
    """fr3d_level_sweeps_slab (the z-slab multi-GPU seam): a call updates the voxels of its planes only, and running
    every wave slab by slab reproduces the full solve bit for bit -- within a wave no voxel reads a value written in
    that wave, so the slabs of one wave commute.  (The 2-/3-rank gloo tests cannot see a mask that lets everything
    through: redundant work gives the same answer.  Kernel-logic emulator only; the plane range reaches the CUDA
    kernel through its flag word, which no GPU run has exercised yet.)"""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    from flowreg3d_b200.core import _check
    from flowreg3d_b200.multigpu import _level_info
    g = golden("flow_small")
    fixed, moving = g["fixed"][:9, :15, :17].astype(np.float32), g["moving"][:9, :15, :17].astype(np.float32)
    T = 7
    fp = F.FlowParams(alpha=(0.25, 0.3, 0.2), update_lag=3, iterations=T, min_level=0, levels=100, eta=0.8,
                      a_smooth=1.0, a_data=0.45)
    reg = F.Registration(fixed.shape[:3], fixed.shape[3], fp, max_batch=2)
    reg.set_reference(fixed)
    mv = np.stack([moving, np.roll(moving, 1, 2)], 0)
    ref = dev.to_host(reg.get_displacement(mv))
    lib, h = reg.ctx.lib, reg.ctx.h
    mvd = reg._as_dev(mv, np.float32, None)
    B = 2
    out = dev.empty((B,) + reg.shape + (3,), np.float32, reg.device)
    checked_mask = False
    for li in range(lib.fr3d_level_count(h)):
        _check(h, lib.fr3d_level_begin(h, li, dev.ptr(mvd), None, B))
        (pz, py, px), S, _, _ = _level_info(reg, li)
        cuts = [0, max(1, pz // 3), max(2, (2 * pz) // 3), pz] if pz >= 3 else [0, pz]
        for q in range(S + 2 * (T - 1)):
            for a, b in zip(cuts, cuts[1:]):
                if b > a:
                    _check(h, lib.fr3d_level_sweeps_slab(h, li, q, q + 1, a, b))
                if q == 0 and a == 0 and pz >= 3 and not checked_mask:
                    # after the first slab's wave 0 only plane 0 .. cuts[1]-1 may be non-zero
                    buf = dev.empty((B, pz * py * px, 3), np.float64, reg.device)   # float64 state: {du, dv, dw}
                    _check(h, lib.fr3d_level_planes(h, li, 0, dev.ptr(buf), 0, pz))
                    reg.sync()
                    st = dev.to_host(buf).reshape(B, pz, py, px, 3)
                    assert np.abs(st[:, 0, 0, 0, :3]).sum() > 0 and not np.abs(st[:, cuts[1]:]).any()
                    checked_mask = True
        _check(h, lib.fr3d_level_end(h, li))
    _check(h, lib.fr3d_flow_finish(h, dev.ptr(out), reg._code(out)))
    reg.sync()
    assert checked_mask and np.array_equal(dev.to_host(out), ref)
    # a later slab alone must not touch the first plane
    _check(h, lib.fr3d_level_begin(h, 0, dev.ptr(mvd), None, B))
    (pz, py, px), S, _, _ = _level_info(reg, 0)
    _check(h, lib.fr3d_level_sweeps_slab(h, 0, 0, 1, pz // 2, pz))
    buf = dev.empty((B, pz * py * px, 3), np.float64, reg.device)
    _check(h, lib.fr3d_level_planes(h, 0, 0, dev.ptr(buf), 0, pz))
    reg.sync()
    assert not np.abs(dev.to_host(buf)).any()        # wave 0 is voxel (0,0,0): not in [pz/2, pz)
    _check(h, lib.fr3d_level_sweeps(h, 0, -1, -1, -1, -1))
    _check(h, lib.fr3d_level_end(h, 0))
    for li in range(1, lib.fr3d_level_count(h)):
        _check(h, lib.fr3d_level_begin(h, li, dev.ptr(mvd), None, B))
        _check(h, lib.fr3d_level_sweeps(h, li, -1, -1, -1, -1))
        _check(h, lib.fr3d_level_end(h, li))


def test_imregister_interleaved_flow_views_equal_separate_planes(backend):
    """imregister_wrapper called with the three component views of ONE float32 (Z,Y,X,3) flow array (what the
    reference's executors pass, sequential_3d.py:148-160) takes a single-upload path; it must equal the call with
    three separate planes bit for bit, for both interpolations, and equal the oracle like the plain path does."""
    import flowreg3d_b200 as F
    from tests_inputs import smooth_flow
    rng = np.random.default_rng(4)
    shp = (7, 30, 44)
    f2, f1 = rng.random(shp + (2,)).astype(np.float32), rng.random(shp + (2,)).astype(np.float32)
    flow = (smooth_flow(shp, 6, 2.5, 4.0) * np.array([1.0, -1.5, 0.7])).astype(np.float32)
    for meth in ("cubic", "linear"):
        a = F.imregister_wrapper(f2, flow[..., 0], flow[..., 1], flow[..., 2], f1, meth)
        b = F.imregister_wrapper(f2, flow[..., 0].copy(), flow[..., 1].copy(), flow[..., 2].copy(), f1, meth)
        assert np.array_equal(a, b), meth
        o = O.imregister_wrapper(f2, flow[..., 0], flow[..., 1], flow[..., 2], f1, meth)
        assert ulp_diff(a, o).max() <= 1


@pytest.mark.parametrize("kind", ["gray", "cs"])
def test_motion_tensor_gray_and_cs_variants(backend, golden, kind):
    """get_motion_tensor_gray / get_motion_tensor_cs (core/optical_flow_3d.py:155-259; never called by the reference's
    driver, stage functions here too) against the live reference's golden: bit-equal (same operations, same order; the
    library is built without FMA contraction)."""
    from flowreg3d_b200 import core
    g = golden("motion_tensor_alt")
    J = core.motion_tensor(g["f1"], g["f2"], *g["h"], kind=kind)
    ref = g["J_" + kind]
    assert not ref[:, 0].any() and not ref[:, :, :, -1].any()              # the reference's zero ring
    inner = ref[:, 1:-1, 1:-1, 1:-1]
    assert np.abs(J - inner).max() <= 1e-12 * np.abs(inner).max()
    with pytest.raises(ValueError):
        core.motion_tensor(g["f1"], g["f2"], *g["h"], kind="nope")
