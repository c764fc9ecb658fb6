"""Oracle for the rigid cross-correlation pre-alignment (oracle/xcorr.py) against the
reference's own known-answer tests for this path (/root/reference/tests/util/test_xcorr_prealignment.py:13-330),
mirrored here with fixed seeds and the reference's tolerances.  scikit-image is absent, so these KATs are the
only pin of the restated phase_cross_correlation (the oracle's header says so).  CPU only."""
import numpy as np
import pytest
from scipy.ndimage import shift as ndi_shift

from oracle import oracle as O
from oracle import xcorr as X


def _blob(shape, centre, width, amp):
    z, y, x = np.ogrid[:shape[0], :shape[1], :shape[2]]
    return amp * np.exp(-((z - centre[0]) ** 2 + (y - centre[1]) ** 2 + (x - centre[2]) ** 2) / width)


def _moved(ref, d, noise, rng):
    mov = ndi_shift(ref, shift=(d[2], d[1], d[0]), order=1, mode="nearest")
    return mov + noise * rng.standard_normal(ref.shape)


def test_pure_translation():  # reference test :13-60
    rng = np.random.default_rng(0)
    ref = rng.random((40, 128, 128)).astype(np.float32) + _blob((40, 128, 128), (20, 64, 64), 200, 10)
    true = np.array([3.2, -1.5, 2.0], np.float32)
    mov = _moved(ref, true, 0.1, rng)
    est = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(128, 128), up=20)
    assert est.dtype == np.float32 and est.shape == (3,)
    assert np.allclose(est, true, atol=0.3), est
    aligned = O.imregister_wrapper(mov, est[0], est[1], est[2], ref, "linear")
    assert np.mean(np.abs(aligned - ref)) < 0.5


def test_multichannel_alignment():  # :63-111
    rng = np.random.default_rng(1)
    ref = rng.random((30, 64, 64, 2)).astype(np.float32)
    ref[..., 0] += _blob((30, 64, 64), (15, 32, 32), 100, 5)
    ref[..., 1] += _blob((30, 64, 64), (15, 20, 40), 100, 5)
    true = np.array([2.5, -1.0, 1.5], np.float32)
    mov = np.empty_like(ref)
    for c in range(2):
        mov[..., c] = ndi_shift(ref[..., c], shift=(true[2], true[1], true[0]), order=1, mode="nearest")
    mov += 0.05 * rng.standard_normal(ref.shape)
    est = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(64, 64), up=10, weight=np.array([0.5, 0.5], np.float32))
    assert np.allclose(est, true, atol=0.5), est
    aligned = O.imregister_wrapper(mov, est[0], est[1], est[2], ref, "linear")
    assert np.mean(np.abs(aligned - ref)) < 0.5
    # no weights: plain channel mean (:33-36)
    est2 = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(64, 64), up=10)
    assert np.allclose(est2, true, atol=0.5), est2


def test_downsampling_accuracy():  # :114-154
    rng = np.random.default_rng(2)
    shape = (50, 256, 256)
    ref = rng.random(shape).astype(np.float32) * 0.5 + _blob(shape, (25, 128, 128), 500, 10)
    true = np.array([5.0, -3.0, 4.0], np.float32)
    mov = _moved(ref, true, 0.1, rng)
    for target in (256, 128, 64):
        est = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(target, target), up=10)
        assert np.allclose(est, true, atol=0.5 if target >= 128 else 1.0), (target, est)


def test_sign_convention():  # :157-209
    ref = np.zeros((20, 50, 50), np.float32)
    ref[10, 25, 25] = 100
    mov = np.zeros_like(ref)
    mov[10, 28, 22] = 100
    est = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=None, up=1)
    assert np.allclose(est, np.array([-3.0, 3.0, 0.0], np.float32), atol=1.0), est
    aligned = O.imregister_wrapper(mov, est[0], est[1], est[2], ref, "linear")
    pos = np.unravel_index(np.argmax(aligned), aligned.shape)
    assert all(abs(int(a) - b) <= 1 for a, b in zip(pos, (10, 25, 25))), pos


def test_z_axis_scaling():  # :212-266
    rng = np.random.default_rng(3)
    shape = (80, 128, 128)
    ref = rng.random(shape).astype(np.float32) * 0.5 + _blob(shape, (40, 64, 64), 300, 10)
    true = np.array([2.5, -1.5, 3.5], np.float32)
    mov = _moved(ref, true, 0.1, rng)
    e0 = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(128, 128), target_z=None, up=10)
    e1 = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(128, 128), target_z=40, up=10)
    assert np.abs(e0 - true).max() < 0.5, e0
    assert np.abs(e1 - true).max() < 0.8, e1


def test_pipeline_integration():  # :269-330
    rng = np.random.default_rng(4)
    shape = (40, 100, 100)
    ref = rng.random(shape).astype(np.float32) * 0.2 + _blob(shape, (20, 50, 50), 150, 1)
    true = np.array([4.5, -2.3, 1.8], np.float32)
    mov = _moved(ref, true, 0.05, rng)
    rigid = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(100, 100), up=20)
    assert np.allclose(rigid, true, atol=0.3), rigid
    pre = O.imregister_wrapper(mov, rigid[0], rigid[1], rigid[2], ref, "linear")
    assert np.mean(np.abs(pre - ref)) < 0.3


def test_phase_cross_correlation_known_shifts():
    """Circular integer and sub-pixel shifts of a band-limited image are recovered exactly / to 1/up,
    and `disambiguate` resolves the wrap-around ambiguity for a non-periodic scene."""
    rng = np.random.default_rng(5)
    a = rng.random((64, 80)).astype(np.float32)
    b = np.roll(a, (5, -7), (0, 1))
    s = X.phase_cross_correlation(a, b, upsample_factor=1)
    assert np.array_equal(s, [-5.0, 7.0])
    # sub-pixel: Fourier shift by (2.3, -4.6)
    fy, fx = np.fft.fftfreq(64)[:, None], np.fft.fftfreq(80)[None, :]
    sm = np.fft.ifft2(np.fft.fft2(np.exp(-((np.arange(64)[:, None] - 30) ** 2 + (np.arange(80)[None, :] - 40) ** 2)
                                         / 50.0)) * np.exp(-2j * np.pi * (2.3 * fy - 4.6 * fx))).real
    base = np.exp(-((np.arange(64)[:, None] - 30) ** 2 + (np.arange(80)[None, :] - 40) ** 2) / 50.0)
    s = X.phase_cross_correlation(base.astype(np.float32), sm.astype(np.float32), upsample_factor=10,
                                  normalization=None)
    assert np.allclose(s, [-2.3, 4.6], atol=0.1001), s
    # a scene moved by more than half the image: without disambiguation the wrapped shift comes back
    big = np.zeros((40, 40), np.float32)
    big[5:12, 6:14] = rng.random((7, 8)) + 1
    mv = np.zeros_like(big)
    mv[5 + 24:12 + 24, 6:14] = big[5:12, 6:14]
    s0 = X.phase_cross_correlation(big, mv, upsample_factor=1, disambiguate=False)
    s1 = X.phase_cross_correlation(big, mv, upsample_factor=1, disambiguate=True)
    assert s0[0] == 16.0 and s1[0] == -24.0 and s1[1] == 0.0, (s0, s1)


def test_flow_with_cc_initialization_small():
    """The six executor steps (sequential_3d.py:89-145): a rigid offset larger than the coarse-to-fine range of a
    shallow pyramid is recovered by the pre-alignment, and total = w_init + rigid + residual."""
    from tests_inputs import synth_volume
    shape = (12, 40, 48)
    ref = synth_volume(shape, 3).astype(np.float64)[..., None]
    d = np.array([4.0, -3.0, 1.0])
    mov = O.imregister_wrapper(ref, np.full(shape, -d[0]), np.full(shape, -d[1]), np.full(shape, -d[2]), ref,
                               "linear").astype(np.float64)[..., None]
    params = dict(alpha=(0.25,) * 3, update_lag=5, iterations=10, min_level=2, levels=100, eta=0.8, a_smooth=1.0,
                  a_data=0.45, weight=np.ones(shape + (1,)))
    w0 = np.zeros(shape + (3,))
    flow, w_cross = X.flow_with_cc_initialization(ref, mov, w0, params, cc_hw=64, cc_up=10)
    assert flow.dtype == np.float32 and flow.shape == shape + (3,)
    assert np.allclose(w_cross, d, atol=0.5), w_cross
    core = (slice(3, -3), slice(8, -8), slice(8, -8))
    assert np.abs(flow[core].reshape(-1, 3).mean(0) - d).max() < 0.5


# ----------------------------------------------------------------------------------------------
# golden of the LIVE reference (tests/golden/gen_golden.py xcorr): the reference's own
# estimate_rigid_xcorr_3d and SequentialExecutor3D.process_batch(cc_initialization=True), run with
# `skimage.registration.phase_cross_correlation` stubbed by the oracle's restatement.  Pins every
# line of reference code around that call.
# ----------------------------------------------------------------------------------------------
def _xcorr_inputs(golden):
    from tests_inputs import synth_volume
    g = golden("xcorr")
    ref = np.stack([synth_volume((24, 72, 96), 50 + c) for c in range(2)], -1)
    assert np.isclose(ref.astype(np.float64).sum(), g["ref_checksum"][0], rtol=1e-12)
    assert float(ref[5, 7, 11, 1]) == g["ref_checksum"][1]
    return g, ref


def test_estimate_matches_live_reference(golden):
    g, ref = _xcorr_inputs(golden)
    for k in range(2):
        mov = g["batch"][k]
        e = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=(48, 64), up=10, weight=np.array([0.3, 0.7], np.float32))
        assert np.array_equal(e, g[f"est{k}_w"]), (e, g[f"est{k}_w"])
        e = X.estimate_rigid_xcorr_3d(ref, mov, target_hw=None, up=20)
        assert np.array_equal(e, g[f"est{k}_full"])
        e = X.estimate_rigid_xcorr_3d(ref[..., 0], mov[..., 0], target_hw=(72, 48), target_z=12, up=5)
        assert np.array_equal(e, g[f"est{k}_z"])
        # and they are what the generator applied (reference tolerance for down-sampled projections: 0.5)
        assert np.abs(g[f"est{k}_full"] - g["shifts"][k]).max() < 0.5


def test_executor_steps_match_live_reference(golden):
    g, ref = _xcorr_inputs(golden)
    assert int(g["two_channel_raises"][0]) == 1      # reference behaviour: C > 1 + cc_initialization raises
    ref1 = ref[..., :1].astype(np.float64)
    batch = g["batch"][..., :1].astype(np.float64)
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]])
    rp = O.preprocess(ref1, sigma)
    bp = O.preprocess(batch, sigma, ref1)
    params = dict(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, eta=0.8, update_lag=4, iterations=8,
                  a_smooth=1.0, a_data=0.45, weight=np.ones(ref1.shape))
    for t in range(2):
        flow, _ = X.flow_with_cc_initialization(rp, bp[t], g["w_init"].astype(np.float32), params,
                                                cc_hw=(48, 64), cc_up=10)
        d = np.sqrt(((flow[:, ::2, ::2] - g["flows_s2"][t]) ** 2).sum(-1))
        assert d.mean() <= 1e-4 and d.max() <= 5e-3, (t, d.mean(), d.max())
        reg = O.imregister_wrapper(g["batch"][t][..., :1].astype(np.float64), flow[..., 0], flow[..., 1],
                                   flow[..., 2], ref1, "cubic")
        reg = reg[..., None] if reg.ndim == 3 else reg
        r = g["registered_s2"][t]
        assert np.linalg.norm(reg[:, ::2, ::2] - r) <= 1e-4 * np.linalg.norm(r)
