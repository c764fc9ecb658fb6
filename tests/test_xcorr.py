"""Rigid cross-correlation pre-alignment (OFOptions.cc_initialization) through the C ABI,
against the oracle (oracle/xcorr.py) and the live-reference golden (tests/golden/xcorr.npz: the reference's own
estimate_rigid_xcorr_3d and SequentialExecutor3D.process_batch with the scikit-image call stubbed by the oracle's
restatement).  Tolerances next to each assertion."""
import numpy as np
import pytest
from scipy import ndimage as ndi

from oracle import oracle as O
from oracle import xcorr as OX


def _golden_inputs(golden):
    from tests_inputs import synth_volume
    g = golden("xcorr")
    ref = np.stack([synth_volume((24, 72, 96), 50 + c) for c in range(2)], -1)
    assert float(ref[5, 7, 11, 1]) == g["ref_checksum"][1]
    return g, ref


def test_wrap_shift_matches_scipy(backend):
    """The periodic cubic-spline shift used by the wrap disambiguation == scipy.ndimage.shift(mode="grid-wrap"),
    float32 result: bit-exact for the prefilter + interpolation in float64 (same recursion as scipy)."""
    from flowreg3d_b200 import core, device as dev
    rng = np.random.default_rng(0)
    img = rng.random((3, 17, 20)).astype(np.float32)
    ctx = core.bare_context()
    c = np.zeros((3, 17, 20, 2))
    c[..., 0] = img
    t = dev.to_device(c, ctx.device)
    work = dev.empty((3, 17, 20), np.float64, ctx.device)
    out = dev.empty((3, 17, 20), np.float64, ctx.device)
    sh = np.array([[2.3, -4.6], [3.0, -2.0], [-0.5, 7.25]])
    core._check(ctx.h, ctx.lib.fr3d_cc_wrap_shift(ctx.h, dev.ptr(t), 3, 17, 20, sh.ctypes.data, dev.ptr(work),
                                                  dev.ptr(out)))
    ctx.sync()
    o = dev.to_host(out)
    for b in range(3):
        order = 3 if np.any(sh[b] % 1 != 0) else 0
        r = ndi.shift(img[b], sh[b], mode="grid-wrap", order=order)
        assert np.abs(o[b] - r).max() <= 1e-6, b          # <= 1 float32 ulp of O(1) data (CUDA vs libm pow-free path)


@pytest.mark.parametrize("case", [0, 1, 2])
def test_estimate_equals_oracle(backend, case):
    """estimate_rigid_xcorr_3d on the device == the oracle, bit for bit (the estimate is quantised to 1/up of a
    projection pixel; the float64 DFT products can only matter at ties), over down-sampled / full / z-scaled
    projections and integer (up = 1) estimates."""
    from flowreg3d_b200 import xcorr as PX
    rng = np.random.default_rng(10 + case)
    shp = (24, 72, 96)
    z, y, x = np.ogrid[:shp[0], :shp[1], :shp[2]]
    ref = (rng.random(shp) * 0.5 + 5 * np.exp(-((z - 12) ** 2 + (y - 36) ** 2 + (x - 50) ** 2) / 150)).astype(np.float32)
    d = [(3.2, -1.5, 2.0), (-4.6, 2.4, -1.0), (0.0, 6.3, 0.5)][case]
    mov = (ndi.shift(ref, shift=(d[2], d[1], d[0]), order=1, mode="nearest")
           + 0.05 * rng.standard_normal(shp)).astype(np.float32)
    for kw in (dict(target_hw=(48, 64), up=10), dict(target_hw=None, up=20),
               dict(target_hw=(72, 48), target_z=12, up=5), dict(target_hw=None, up=1)):
        e_o = OX.estimate_rigid_xcorr_3d(ref, mov, **kw)
        e_p = PX.estimate_rigid_xcorr_3d(ref, mov, **kw)
        assert e_p.dtype == np.float32 and np.array_equal(e_o, e_p), (kw, e_o, e_p)
    assert np.abs(OX.estimate_rigid_xcorr_3d(ref, mov, target_hw=None, up=20) - np.array(d)).max() < 0.5


def test_estimate_matches_live_reference_golden(backend, golden):
    from flowreg3d_b200 import xcorr as PX
    g, ref = _golden_inputs(golden)
    for k in range(2):
        mov = g["batch"][k]
        e = PX.estimate_rigid_xcorr_3d(ref, mov, target_hw=(48, 64), up=10, weight=np.array([0.3, 0.7], np.float32))
        assert np.array_equal(e, g[f"est{k}_w"]), (e, g[f"est{k}_w"])
        e = PX.estimate_rigid_xcorr_3d(ref, mov, target_hw=None, up=20)
        assert np.array_equal(e, g[f"est{k}_full"])
        e = PX.estimate_rigid_xcorr_3d(ref[..., 0], mov[..., 0], target_hw=(72, 48), target_z=12, up=5)
        assert np.array_equal(e, g[f"est{k}_z"])


def test_executor_with_cc_initialization_matches_live_reference(backend, golden):
    """B200Executor3D.process_batch with cc_initialization=True against the reference's sequential executor
    (golden): flows within 1e-4 mean / 5e-3 max voxel, registered frames within 1e-4 relative L2 (the north-star
    tolerances are 0.01 / 0.05 / 1e-4); and the reference's behaviour for C > 1 (ValueError)."""
    import flowreg3d_b200 as F
    g, ref = _golden_inputs(golden)
    ref1 = ref[..., :1]
    batch = g["batch"][..., :1]
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]])
    rp = O.preprocess(ref1.astype(np.float64), sigma)
    bp = O.preprocess(batch.astype(np.float64), sigma, ref1.astype(np.float64))
    Z, Y, X = ref1.shape[:3]
    fp = dict(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, eta=0.8, update_lag=4, iterations=8, a_smooth=1.0,
              a_data=0.45, weight=np.ones((Z, Y, X, 1)), cc_initialization=True, cc_hw=(48, 64), cc_up=10)
    ex = F.B200Executor3D(max_batch=2)
    try:
        reg, flows = ex.process_batch(batch, bp, ref1, rp, g["w_init"], None, None, "cubic", None, flow_params=fp)
        assert reg.dtype == batch.dtype and flows.dtype == np.float32
        for t in range(2):
            d = np.sqrt(((flows[t][:, ::2, ::2] - g["flows_s2"][t]) ** 2).sum(-1))
            assert d.mean() <= 1e-4 and d.max() <= 5e-3, (t, d.mean(), d.max())
            r = g["registered_s2"][t]
            assert np.linalg.norm(reg[t][:, ::2, ::2] - r) <= 1e-4 * np.linalg.norm(r)
        with pytest.raises(ValueError):
            b2 = g["batch"]
            ex.process_batch(b2, b2.astype(np.float64), ref, ref.astype(np.float64), g["w_init"], None, None,
                             "cubic", None, flow_params=dict(fp, weight=np.full((Z, Y, X, 2), 0.5)))
    finally:
        ex.cleanup()


def test_sequence_with_cc_initialization(backend):
    """compensate_arr_3D with OFOptions(cc_initialization=True) and cc_prealign=True (the pre-alignment the
    reference's executors implement, run for every frame): a rigid offset beyond the reach of a shallow pyramid is
    recovered; the result equals the oracle's per-frame restatement of the executor steps chained with the
    reference's w_init bookkeeping (zero field for the first batch -- no bootstrap solve --, then the mean of the
    previous batch's flows; the chaining itself is pinned by test_sequence_with_cc_initialization_live_golden)."""
    import flowreg3d_b200 as F
    from tests_inputs import synth_volume
    shape = (12, 40, 48)
    ref = synth_volume(shape, 3)
    d = np.array([4.0, -3.0, 1.0])
    r64 = ref.astype(np.float64)
    mov = O.imregister_wrapper(r64, np.full(shape, -d[0]), np.full(shape, -d[1]), np.full(shape, -d[2]), r64, "linear")
    rng = np.random.default_rng(7)
    video = np.stack([mov + 0.002 * rng.standard_normal(shape).astype(np.float32) for _ in range(3)], 0)[..., None]
    opts = F.OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=10, update_lag=5, buffer_size=3,
                       cc_initialization=True, cc_hw=64, cc_up=10)
    reg, w = F.compensate_arr_3D(video, ref, opts, cc_prealign=True)
    assert reg.shape == video.shape and w.shape == (3,) + shape + (3,)
    core = (slice(None), slice(3, -3), slice(8, -8), slice(8, -8))
    assert np.abs(w[core].reshape(-1, 3).mean(0) - d).max() < 0.5
    # the same through the oracle
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]])
    ref4 = r64[..., None]
    rp = O.preprocess(ref4, sigma)
    bp = O.preprocess(video.astype(np.float64), sigma, ref4)
    params = dict(alpha=(0.25,) * 3, update_lag=5, iterations=10, min_level=2, levels=100, eta=0.8, a_smooth=1.0,
                  a_data=0.45, weight=np.ones(shape + (1,)))
    # compensate_recording_3D.py:346-356: with cc_initialization the first batch starts from w_init = 0 (no bootstrap)
    w_init = np.zeros(shape + (3,), np.float32)
    for t in range(3):
        fo, _ = OX.flow_with_cc_initialization(rp, bp[t], w_init.astype(np.float32), params, cc_hw=64, cc_up=10)
        e = np.sqrt(((w[t] - fo) ** 2).sum(-1))
        assert e.mean() <= 1e-4 and e.max() <= 5e-3, (t, e.mean(), e.max())
    with pytest.raises(ValueError):
        F.compensate_arr_3D(np.repeat(video, 2, -1), np.stack([ref, ref], -1), opts, cc_prealign=True)


def test_sequence_with_cc_initialization_live_golden(backend, golden):
    """compensate_arr_3D(cc_initialization=True) over two batches against the LIVE reference's BatchMotionCorrector
    (tests/golden/xcorr_sequence.npz): first batch from w_init = 0 with NO bootstrap solve
    (compensate_recording_3D.py:346-356), second batch from the mean of the first batch's flows -- and NO rigid
    pre-alignment, because the reference pipeline never hands the cc_* keys to its executors (:301-315)."""
    import flowreg3d_b200 as F
    from tests_inputs import synth_volume
    g = golden("xcorr_sequence")
    video = g["video"]
    ref = synth_volume(video.shape[1:4], 3)
    opts = F.OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=10, update_lag=5, buffer_size=2,
                       cc_initialization=True, cc_hw=64, cc_up=10)
    reg, w = F.compensate_arr_3D(video, ref[..., None], opts)
    e = np.sqrt(((w.astype(np.float64) - g["w"]) ** 2).sum(-1))
    assert e.mean() <= 1e-4 and e.max() <= 5e-3, (e.mean(), e.max())     # tolerance: 0.01 / 0.05
    assert np.linalg.norm(reg - g["registered"]) <= 1e-5 * np.linalg.norm(g["registered"])


def test_block_scan_option_gives_the_same_estimates(backend):
    """FR3D_OPT_CC_BLOCK_SCANS (1 = default since round 2: block-cooperative arg-max / tile sums / plane mean, 30.1 ->
    5.5 ms of pre-alignment kernels per 10 frames on a B200) changes no estimate against the one-thread-per-plane
    scans (option 0)."""
    from flowreg3d_b200 import _lib, core, xcorr as PX
    rng = np.random.default_rng(3)
    shp = (20, 60, 70)
    z, y, x = np.ogrid[:shp[0], :shp[1], :shp[2]]
    ref = (rng.random(shp) * 0.5 + 5 * np.exp(-((z - 10) ** 2 + (y - 30) ** 2 + (x - 33) ** 2) / 120)).astype(np.float32)
    mov = (ndi.shift(ref, shift=(1.5, -2.5, 3.75), order=1, mode="nearest")
           + 0.05 * rng.standard_normal(shp)).astype(np.float32)
    ctx = core.bare_context()
    core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_CC_BLOCK_SCANS, 0))
    base = [PX.estimate_rigid_xcorr_3d(ref, mov, **kw) for kw in (dict(target_hw=(40, 48), up=10), dict(target_hw=None, up=1))]
    core._check(ctx.h, ctx.lib.fr3d_set_option(ctx.h, _lib.OPT_CC_BLOCK_SCANS, 1))
    ctx.profile(True)
    try:
        blk = [PX.estimate_rigid_xcorr_3d(ref, mov, **kw) for kw in (dict(target_hw=(40, 48), up=10), dict(target_hw=None, up=1))]
        ran = list(ctx.profile_report())
        for name in ("CcAbsArgmaxTileK", "CcTileSumsTileK", "CcPlaneMeanTileK"):
            assert any(name in k for k in ran), (name, ran)
    finally:
        ctx.profile(False)
    for a, b, kw in zip(base, blk, ("down-sampled", "integer")):
        assert np.array_equal(a, b), (kw, a, b)
        assert np.array_equal(a, OX.estimate_rigid_xcorr_3d(ref, mov, **(dict(target_hw=(40, 48), up=10)
                                                                       if kw == "down-sampled" else dict(target_hw=None, up=1))))


def test_estimate_edge_cases_equal_oracle(emu_backend):
    """Odd / small / wide volumes, zero shift (empty disambiguation tiles), shifts beyond half the projection (wrap
    candidates), targets larger than the volume (no resize), coarse targets, z down-sampling, integer estimates."""
    from flowreg3d_b200 import xcorr as PX
    rng = np.random.default_rng(1)
    for shp in [(7, 33, 41), (16, 64, 64), (5, 20, 130)]:
        z, y, x = np.ogrid[:shp[0], :shp[1], :shp[2]]
        ref = (rng.random(shp) * 0.5 + 3 * np.exp(-((z - shp[0] / 2) ** 2 + (y - shp[1] / 2) ** 2
                                                    + (x - shp[2] / 2) ** 2) / 60)).astype(np.float32)
        for d in [(0, 0, 0), (1, 0, 0), (-7.5, 3.25, 1.0), (12.0, -9.0, 2.0)]:
            mov = ndi.shift(ref, shift=(d[2], d[1], d[0]), order=1, mode="nearest").astype(np.float32)
            for kw in (dict(target_hw=(256, 256), up=10), dict(target_hw=(16, 24), up=4),
                       dict(target_hw=None, up=1, target_z=4)):
                eo = OX.estimate_rigid_xcorr_3d(ref, mov, **kw)
                ep = PX.estimate_rigid_xcorr_3d(ref, mov, **kw)
                assert np.array_equal(eo, ep), (shp, d, kw, eo, ep)
