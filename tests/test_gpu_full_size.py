"""GPU-only tests at BASELINE.json's full sizes (32x512x512x2) through size-independent properties,
and CUDA-vs-oracle parity at sizes the oracle finishes in seconds."""
import numpy as np
import pytest

from conftest import epe_stats, rel_l2
from oracle import oracle as O
from tests_inputs import smooth_flow, synth_volume

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_ref():
    return np.stack([synth_volume((32, 512, 512), 10 + c) for c in range(2)], -1)


def test_full_size_identity_and_translation(cuda_backend, full_ref):
    """Warp properties at 32x512x512x2: zero flow is the identity (cubic B-spline interpolates its
    samples), an integer shift moves voxels exactly, out-of-volume voxels take the reference."""
    import flowreg3d_b200 as F
    Z, Y, X, C = full_ref.shape
    reg = F.Registration((Z, Y, X), C, F.FlowParams(min_level=5, a_smooth=1.0, iterations=2), max_batch=1)
    from flowreg3d_b200 import device as dev
    zero = np.zeros((1, Z, Y, X, 3), np.float32)
    out = dev.to_host(reg.compensate(full_ref[None], zero, ref_raw=full_ref))[0]
    assert np.abs(out - full_ref).max() <= 2e-6          # prefilter + interpolation round trip, float64 math
    shift = zero.copy()
    shift[..., 0] = 3.0
    shift[..., 1] = -2.0
    other = (1.0 - full_ref).astype(np.float32)
    out = dev.to_host(reg.compensate(full_ref[None], shift, ref_raw=other))[0]
    assert np.abs(out[:, 2:, :-3] - full_ref[:, :-2, 3:]).max() <= 2e-6
    assert np.array_equal(out[:, :, -3:], other[:, :, -3:])   # x + 3 >= X  -> reference value
    assert np.array_equal(out[:, :2], other[:, :2])           # y - 2 < 0   -> reference value
    reg.ctx.close()


def test_full_size_default_options_recovers_flow(cuda_backend, full_ref):
    """Config 2 frame: the estimated flow explains the synthetic motion, identical frames give zero
    flow, and a frame of the batch does not depend on its batch mates."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    Z, Y, X, C = full_ref.shape
    g = smooth_flow((Z, Y, X), 1000, 2.0, 12.0)
    r64 = full_ref.astype(np.float64)
    mov = F.imregister_wrapper(r64, -g[..., 0], -g[..., 1], -g[..., 2], r64, "linear")
    opts = F.OFOptions(buffer_size=3)
    seq = F.SequenceCorrector(full_ref, opts, max_batch=3)
    batch = np.stack([mov, full_ref, mov], 0)
    proc = seq.reg.preprocess(batch, seq.lo, seq.den)
    flows = dev.to_host(seq.reg.get_displacement(proc))
    assert np.array_equal(flows[0], flows[2])                  # batch members are independent
    assert np.abs(flows[1]).max() <= 1e-6                      # moving == fixed -> zero flow
    mean, mx = epe_stats(flows[0], g)
    assert mean <= 0.15, mean                                  # the reference reaches 0.08 here (oracle run)
    single = dev.to_host(seq.reg.get_displacement(proc[:1]))[0]
    assert np.array_equal(single, flows[0])                    # B = 1 == B = 3
    regd = dev.to_host(seq.reg.compensate(batch[:1], flows[:1]))[0]
    assert rel_l2(regd, full_ref) < rel_l2(mov, full_ref) * 0.35
    seq.close()


def test_mid_size_cuda_vs_oracle(cuda_backend):
    """CUDA vs oracle on a 2-channel 20x96x112 pair at OFOptions-default parameters."""
    import flowreg3d_b200 as F
    Z, Y, X = 20, 96, 112
    fixed = np.stack([synth_volume((Z, Y, X), 50 + c) for c in range(2)], -1)
    g = smooth_flow((Z, Y, X), 3, 1.5, 8.0)
    f64 = fixed.astype(np.float64)
    moving = O.imregister_wrapper(f64, -g[..., 0], -g[..., 1], -g[..., 2], f64, "linear")
    kw = dict(alpha=(0.25,) * 3, update_lag=5, iterations=100, min_level=2, levels=100, eta=0.8, a_smooth=1.0,
              a_data=0.45, weight=np.array([0.5, 0.5]))
    flow = F.get_displacement(fixed, moving, **kw)
    ref = O.get_displacement(fixed, moving, **kw)
    mean, mx = epe_stats(flow, ref)
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)             # tolerance: 0.01 / 0.05


def test_sor_many_ctas_and_batches(cuda_backend, golden):
    """The cooperative wavefront kernel with more work than CTAs and B > 1 equals B independent solves."""
    from flowreg3d_b200 import core
    rng = np.random.default_rng(9)
    p, m, n, C = 12, 70, 90, 2
    J = rng.random((C, 10, p, m, n)) * 0.1
    J[:, :4] += 0.5                                            # positive diagonal
    wgt = np.full((C, p, m, n), 0.5)
    uvw = rng.standard_normal((3, p, m, n))
    d = core.sor_level(J, wgt, uvw, (0.3, 0.4, 0.5), (1.5, 1.2, 1.1), 12, 5, 0.45)
    Jr = [np.pad(np.moveaxis(J[:, q], 0, -1), ((1, 1), (1, 1), (1, 1), (0, 0))) for q in range(10)]
    pad = lambda a: np.pad(a, 1, mode="edge")
    o = O.compute_flow_3d(Jr, np.pad(np.moveaxis(wgt, 0, -1), ((1, 1), (1, 1), (1, 1), (0, 0))), pad(uvw[0]),
                          pad(uvw[1]), pad(uvw[2]), (0.3, 0.4, 0.5), 12, 5, np.full(C, 0.45), 1.0, 1.1, 1.2, 1.5)
    inner = (slice(1, -1),) * 3
    assert np.abs(d - np.moveaxis(o[inner], -1, 0)).max() <= 1e-9


def _recoil_jitter_frames(ref, T):
    """BASELINE config 3 style motion: injection every few frames with exponential
    recoil (expansion about the volume centre) plus a scanning jitter, deterministic in t."""
    import flowreg3d_b200 as F
    Z, Y, X, C = ref.shape
    zz, yy, xx = np.meshgrid(np.arange(Z) - Z / 2, np.arange(Y) - Y / 2, np.arange(X) - X / 2, indexing="ij")
    frames = []
    r64 = ref.astype(np.float64)
    for t in range(T):
        mag = 0.04 * np.exp(-(t % 5) / 2.0)
        jit = 1.0 * np.sin(2 * np.pi * t / 17.0)
        u = xx * mag + jit * np.sin(2 * np.pi * yy / Y)
        v = yy * mag
        w = zz * mag * 0.5
        frames.append(O.imregister_wrapper(r64, -u, -v, -w, r64, "linear"))
    return np.stack(frames, 0).astype(np.float32)


def test_config3_sequence_cuda_vs_oracle(cuda_backend):
    """Config 3 (injection / recoil + jitter) at a size the oracle finishes in seconds: the whole
    sequence path (GPU pre-filter, w_init bootstrap + chaining over three batches, ragged last batch,
    pipelined host streaming) against the oracle's restatement of BatchMotionCorrector."""
    import flowreg3d_b200 as F
    Z, Y, X = 24, 64, 72
    ref = np.stack([synth_volume((Z, Y, X), 70 + c) for c in range(2)], -1)
    video = _recoil_jitter_frames(ref, 7)
    opts = F.OFOptions(min_level=2, iterations=40, update_lag=5, buffer_size=3, weight=[0.5, 0.5])
    reg, w = F.compensate_arr_3D(video, ref, opts)
    oreg, ow = O.compensate_arr(video, ref, min_level=2, iterations=40, update_lag=5, buffer_size=3,
                                weight=[0.5, 0.5])
    mean, mx = epe_stats(w, ow)
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)              # tolerance: 0.01 / 0.05
    assert rel_l2(reg, oreg) <= 1e-5                            # tolerance: 1e-4


@pytest.mark.parametrize("min_level,alpha,iters", [(3, 0.1, 20), (3, 2.0, 50), (0, 0.25, 50), (0, 1.0, 20)])
def test_config5_solver_sweep_cuda_vs_oracle(cuda_backend, min_level, alpha, iters):
    """Config 5 (pyramid depth / alpha / iteration sweep) cells on a 1-channel 20x72x80 pair."""
    import flowreg3d_b200 as F
    Z, Y, X = 20, 72, 80
    fixed = synth_volume((Z, Y, X), 90)
    g = smooth_flow((Z, Y, X), 4, 1.5, 7.0)
    f64 = fixed.astype(np.float64)
    moving = O.imregister_wrapper(f64, -g[..., 0], -g[..., 1], -g[..., 2], f64, "linear")
    kw = dict(alpha=(alpha,) * 3, update_lag=5, iterations=iters, min_level=min_level, levels=100, eta=0.8,
              a_smooth=1.0, a_data=0.45)
    flow = F.get_displacement(fixed, moving, **kw)
    ref = O.get_displacement(fixed, moving, **kw)
    mean, mx = epe_stats(flow, ref)
    assert mean <= 1e-4 and mx <= 5e-3, (min_level, alpha, iters, mean, mx)   # tolerance: 0.01 / 0.05
