"""End-to-end parity of the drop-in entry points against the live reference's golden outputs and
the oracle.  North-star tolerances (BASELINE.json): flow mean EPE <= 0.01 voxel, max EPE <= 0.05
voxel, corrected volume relative L2 <= 1e-4.  The assertions below are much tighter: the path
reproduces the reference's rounding points, so what remains is the reference's own float64
last-bit noise (numba fastmath) amplified by float32 re-rounding between levels."""
import numpy as np
import pytest

from conftest import epe_stats, rel_l2
from oracle import oracle as O
from tests_inputs import synth_volume

RUNS = {
    "ml1s": dict(alpha=(0.5,) * 3, update_lag=10, iterations=15, min_level=1, levels=100, eta=0.8,
                 a_smooth=0.5, a_data=0.45),                     # nonlinear smoothness term
    "ml0": dict(alpha=(0.25,) * 3, update_lag=5, iterations=30, min_level=0, levels=100, eta=0.8,
                a_smooth=1.0, a_data=0.45),
    "ml2w": dict(alpha=(0.25, 0.3, 0.2), update_lag=5, iterations=20, min_level=2, levels=100, eta=0.8,
                 a_smooth=1.0, a_data=0.45, weight=np.array([0.3, 0.7])),
}


@pytest.mark.parametrize("state", ["f32", "f64"])
@pytest.mark.parametrize("run", ["ml2w", "ml0", "ml2w_uvw", "ml1s"])
def test_get_displacement_vs_reference_and_oracle(backend, golden, run, state, monkeypatch):
    """state f64 = the shipped default; f32 = reduced-traffic mode (solver increments stored in float32)."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import core
    monkeypatch.setattr(core, "STATE_DTYPE", np.float32 if state == "f32" else np.float64)
    g = golden("flow_small")
    kw = dict(RUNS[run.replace("_uvw", "")])
    if run.endswith("_uvw"):
        kw["uvw"] = g["uvw"].copy()
    flow = F.get_displacement(g["fixed"], g["moving"], **kw)
    assert flow.dtype == np.float64 and flow.shape == g["fixed"].shape[:3] + (3,)
    tol_mean, tol_max = (1e-4, 5e-3) if state == "f64" else (5e-4, 2.5e-2)   # north-star tolerance: 0.01 / 0.05
    mean, mx = epe_stats(flow, g[f"flow_{run}"])
    assert mean <= tol_mean and mx <= tol_max, ("vs reference", mean, mx)
    mean, mx = epe_stats(flow, O.get_displacement(g["fixed"], g["moving"], **kw))
    if state == "f64":
        assert mean <= 1e-6 and mx <= 1e-4, ("vs oracle", mean, mx)
    else:
        assert mean <= tol_mean and mx <= tol_max, ("vs oracle", mean, mx)


def test_get_displacement_argument_handling(backend, golden):
    import flowreg3d_b200 as F
    g = golden("flow_small")
    fx, mv = g["fixed"][:12, :20, :24, 0], g["moving"][:12, :20, :24, 0]
    kw = dict(alpha=0.3, update_lag=5, iterations=5, min_level=0, levels=100, eta=0.8, a_smooth=1.0)
    f3 = F.get_displacement(fx, mv, **kw)                      # 3-D input, scalar alpha
    f4 = F.get_displacement(fx[..., None], mv[..., None], **{**kw, "alpha": (0.3,) * 3}, const_assumption="gray")
    assert np.array_equal(f3, f4)                              # const_assumption is ignored, like the reference
    o = O.get_displacement(fx, mv, **{**kw, "alpha": (0.3,) * 3})
    assert epe_stats(f3, o)[1] <= 1e-6
    # the reference's own defaults (alpha 2, 20 iterations, lag 10, a_smooth 0.5 = nonlinear smoothness)
    fd = F.get_displacement(fx, mv)
    mean, mx = epe_stats(fd, O.get_displacement(fx, mv))
    assert mean <= 1e-6 and mx <= 1e-4, (mean, mx)
    with pytest.raises(ValueError):
        F.get_displacement(fx, mv[:-1], **kw)


def test_executor_matches_sequential_semantics(backend, golden):
    """Boundary B1: process_batch == the reference's per-frame loop (oracle.process_batch)."""
    import flowreg3d_b200 as F
    g = golden("sequence")
    ml, it, lag, _ = (int(v) for v in g["params"])
    ref_raw = g["ref"].astype(np.float64)
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]] * 2)
    ref_proc = O.preprocess(ref_raw, sigma)
    batch = g["video"][:3]
    bproc = O.preprocess(batch, sigma, ref_raw)
    fp = dict(alpha=(0.25,) * 3, weight=np.full(ref_raw.shape, 0.5), levels=100, min_level=ml, eta=0.8,
              update_lag=lag, iterations=it, a_smooth=1.0, a_data=0.45)
    w0 = (0.3 * np.ones(ref_raw.shape[:3] + (3,))).astype(np.float32)
    seen = []
    ex = F.B200Executor3D(max_batch=2)  # 3 frames in chunks of 2: ragged last chunk
    with ex:
        reg, flows = ex.process_batch(batch, bproc, ref_raw, ref_proc, w0, None, None, "cubic", seen.append,
                                      flow_params=fp)
        reg2, flows2 = ex.process_batch(batch[:1], bproc[:1], ref_raw, ref_proc, w0, None, None, "linear",
                                        None, flow_params=fp)
    oreg, oflows = O.process_batch(batch, bproc, ref_raw, ref_proc, w0, "cubic", fp)
    assert reg.dtype == batch.dtype and flows.dtype == np.float32 and sum(seen) == 3
    mean, mx = epe_stats(flows, oflows)
    assert mean <= 1e-6 and mx <= 1e-4
    assert rel_l2(reg, oreg) <= 1e-6
    oreg2, _ = O.process_batch(batch[:1], bproc[:1], ref_raw, ref_proc, w0, "linear", fp)
    assert rel_l2(reg2, oreg2) <= 1e-6
    # cc_initialization on a multi-channel batch: ValueError, as the reference pipeline raises (the full weight array
    # meets estimate_rigid_xcorr_3d's tensordot, xcorr_prealignment.py:26-30); single channel: tests/test_xcorr.py
    with pytest.raises(ValueError):
        ex.process_batch(batch, bproc, ref_raw, ref_proc, w0, flow_params={**fp, "cc_initialization": True})


def test_compensate_arr_3D_vs_reference(backend, golden):
    """Sequence entry point incl. GPU pre-processing, w_init bootstrap and chaining."""
    import flowreg3d_b200 as F
    g = golden("sequence")
    ml, it, lag, buf = (int(v) for v in g["params"])
    opts = F.OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=ml, iterations=it, update_lag=lag,
                       buffer_size=buf, weight=[0.5, 0.5], output_typename=None)
    ticks = []
    reg, w = F.compensate_arr_3D(g["video"], g["ref"], opts, progress_callback=lambda c, t: ticks.append((c, t)))
    assert reg.shape == g["video"].shape and reg.dtype == g["video"].dtype and w.dtype == np.float32
    assert ticks[-1] == (g["video"].shape[0], g["video"].shape[0])
    mean, mx = epe_stats(w, g["w"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)              # tolerance: 0.01 / 0.05
    assert rel_l2(reg, g["registered"]) <= 1e-5                  # tolerance: 1e-4
    assert opts.buffer_size == buf                               # caller's options are not mutated


def test_compensate_arr_3D_shapes(backend, golden):
    import flowreg3d_b200 as F
    g = golden("sequence")
    v = g["video"][:2, :10, :20, :24, 0]
    r = g["ref"][:10, :20, :24, 0]
    opts = F.OFOptions(min_level=1, iterations=4, buffer_size=2)
    reg4, w4 = F.compensate_arr_3D(v, r, opts)                   # (T,Z,Y,X) + (Z,Y,X)
    assert reg4.shape == v.shape and reg4.dtype == np.float64 and w4.shape == v.shape + (3,)
    reg5, w5 = F.compensate_arr_3D(v[..., None], r[..., None], opts)
    assert reg5.shape == v.shape + (1,)
    np.testing.assert_allclose(reg4, reg5[..., 0], rtol=1e-4)    # reference test_compensate_arr_3D.py:402-427
    reg3, w3 = F.compensate_arr_3D(v[0], r, opts)                # single volume
    assert reg3.shape == r.shape and w3.shape == r.shape + (3,)
    with pytest.raises(ValueError):
        F.compensate_arr_3D(np.empty((0, 4, 4, 4)), r, opts)


def test_config1_default_options(backend, golden):
    """BASELINE config 1: 64x128x128 pair, OFOptions-default parameters, vs the live reference."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    from flowreg3d_b200.compensate import normalization_range
    g = golden("config1")
    V = synth_volume((64, 128, 128), 1)
    V64 = V.astype(np.float64)[..., None]
    mov = g["low_moving"][..., None]
    opts = F.OFOptions(sigma=[[1.0, 1.0, 1.0, 0.1]], weight=[1.0])
    seq = F.SequenceCorrector(V64, opts, max_batch=1)
    lo, den = normalization_range(V64, "joint")
    mp = seq.reg.preprocess(mov[None], lo, den)
    flow = seq.reg.get_displacement(mp)
    reg = seq.reg.compensate(mov[None], flow)
    seq.reg.sync()
    flow, reg = dev.to_host(flow)[0], dev.to_host(reg)[0, ..., 0]
    seq.close()
    mean, mx = epe_stats(flow[::2, ::2, ::2], g["low_flow_s2"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)              # tolerance: 0.01 / 0.05
    assert rel_l2(reg[::2, ::2, ::2], g["low_reg_s2"]) <= 1e-5   # tolerance: 1e-4


def test_split_streams_identical(backend, golden):
    """The batch split over two contexts / CUDA streams gives bit-identical results (frames are independent)."""
    import flowreg3d_b200 as F
    from flowreg3d_b200 import device as dev
    g = golden("sequence")
    ml, it, lag, _ = (int(v) for v in g["params"])
    opts = F.OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=ml, iterations=it, update_lag=lag,
                       buffer_size=3, weight=[0.5, 0.5])
    outs = []
    for streams in (1, 2):
        seq = F.SequenceCorrector(g["ref"], opts, max_batch=3, streams=streams)
        reg, fl = seq.process_batch(g["video"][:3])
        seq.reg.sync()
        outs.append((dev.to_host(reg).copy(), dev.to_host(fl).copy()))
        seq.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("C", [1, 3, 4])
def test_channel_counts_and_spatial_weight(backend, golden, C):
    """1 / 3 / 4 channels (the templated solver paths), a spatially varying (Z,Y,X,C) weight and per-channel
    a_data against the oracle."""
    import flowreg3d_b200 as F
    g = golden("flow_small")
    rng = np.random.default_rng(C)
    fx = np.stack([g["fixed"][:14, :26, :30, c % 2] * (1.0 + 0.1 * c) for c in range(C)], -1)
    mv = np.stack([g["moving"][:14, :26, :30, c % 2] * (1.0 + 0.1 * c) for c in range(C)], -1)
    weight = 0.2 + rng.random(fx.shape)
    kw = dict(alpha=(0.3, 0.25, 0.2), update_lag=4, iterations=9, min_level=1, levels=100, eta=0.8, a_smooth=1.0,
              a_data=np.linspace(0.45, 0.8, C), weight=weight)
    flow = F.get_displacement(fx, mv, **kw)
    mean, mx = epe_stats(flow, O.get_displacement(fx, mv, **kw))
    assert mean <= 1e-6 and mx <= 1e-4, (C, mean, mx)


def test_uint16_recording(backend, golden):
    """Integer raw frames: pre-processing from uint16, scipy's integer output rounding in the compensation
    warp, result cast back to the input dtype (sequential_3d.py:163-169)."""
    import flowreg3d_b200 as F
    g = golden("sequence")
    video = np.clip(g["video"][:4, :12, :24, :28] * 4000.0 + 100.0, 0, 65535).astype(np.uint16)
    ref = np.clip(g["ref"][:12, :24, :28] * 4000.0 + 100.0, 0, 65535).astype(np.uint16)
    opts = F.OFOptions(min_level=2, iterations=8, update_lag=4, buffer_size=3, weight=[0.5, 0.5],
                       output_typename=None)
    reg, w = F.compensate_arr_3D(video, ref, opts)
    oreg, ow = O.compensate_arr(video, ref, min_level=2, iterations=8, update_lag=4, buffer_size=3, weight=[0.5, 0.5])
    assert reg.dtype == np.uint16 and reg.shape == video.shape
    mean, mx = epe_stats(w, ow)
    assert mean <= 1e-5 and mx <= 1e-3, (mean, mx)
    assert (reg != oreg).mean() <= 1e-4          # identical up to a vanishing number of rounding ties


def test_sequence_with_temporal_prefilter(backend, golden):
    """OFOptions.sigma with a temporal component: frames of each batch are coupled in the pre-filter."""
    import flowreg3d_b200 as F
    g = golden("sequence")
    video, ref = g["video"][:5, :12, :24, :28], g["ref"][:12, :24, :28]
    sigma = [[1.0, 1.0, 1.0, 0.7], [1.0, 1.0, 1.0, 0.7]]
    opts = F.OFOptions(min_level=2, iterations=6, update_lag=3, buffer_size=3, weight=[0.5, 0.5], sigma=sigma,
                       output_typename=None)
    reg, w = F.compensate_arr_3D(video, ref, opts)
    oreg, ow = O.compensate_arr(video, ref, min_level=2, iterations=6, update_lag=3, buffer_size=3, weight=[0.5, 0.5],
                                sigma=sigma)
    mean, mx = epe_stats(w, ow)
    assert mean <= 1e-5 and mx <= 1e-3, (mean, mx)
    assert rel_l2(reg, oreg) <= 1e-6


def test_flow_statistics(backend, golden):
    """mean / max displacement, mean divergence and mean translation per frame, as BatchMotionCorrector computes
    them from the returned flows (compensate_recording_3D.py:488-508)."""
    import flowreg3d_b200 as F
    g = golden("sequence")
    video, ref = g["video"][:4, :12, :24, :28], g["ref"][:12, :24, :28]
    opts = F.OFOptions(min_level=2, iterations=6, update_lag=3, buffer_size=3, weight=[0.5, 0.5])
    seq = F.SequenceCorrector(ref, opts, statistics=True)
    _, flows = seq.run_pipelined([video[:3], video[3:]])
    st = seq.statistics()
    seq.close()
    w = np.concatenate([f.numpy() for f in flows], 0)
    mag = np.sqrt(w[..., 0] ** 2 + w[..., 1] ** 2 + w[..., 2] ** 2)
    np.testing.assert_allclose(st["mean_disp"], mag.mean(axis=(1, 2, 3)), rtol=1e-5)
    np.testing.assert_allclose(st["max_disp"], mag.max(axis=(1, 2, 3)), rtol=1e-6)
    div = [float(np.mean(np.gradient(w[t, ..., 0], axis=2) + np.gradient(w[t, ..., 1], axis=1) +
                         np.gradient(w[t, ..., 2], axis=0))) for t in range(4)]
    np.testing.assert_allclose(st["mean_div"], div, rtol=1e-4, atol=1e-7)
    tr = [float(np.sqrt(sum(float(np.mean(w[t, ..., q])) ** 2 for q in range(3)))) for t in range(4)]
    np.testing.assert_allclose(st["mean_translation"], tr, rtol=1e-5)


def test_update_reference(backend, golden):
    """update_reference=True: the fixed volume is re-averaged from the compensated pre-processed frames after every
    batch (compensate_recording_3D.py:395-429); golden from the live reference, and the oracle restatement."""
    import flowreg3d_b200 as F
    g, gs = golden("sequence_update_ref"), golden("sequence")
    video, ref = gs["video"][:6, :12, :24, :28], gs["ref"][:12, :24, :28]
    opts = F.OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=8, update_lag=4, buffer_size=3,
                       weight=[0.5, 0.5], update_reference=True, output_typename=None)
    reg, w = F.compensate_arr_3D(video, ref, opts)
    mean, mx = epe_stats(w, g["w"])
    assert mean <= 1e-4 and mx <= 5e-3, (mean, mx)              # tolerance: 0.01 / 0.05
    assert rel_l2(reg, g["registered"]) <= 1e-5                  # tolerance: 1e-4
    oreg, ow = O.compensate_arr(video, ref, min_level=2, iterations=8, update_lag=4, buffer_size=3, weight=[0.5, 0.5],
                                update_reference=True)
    mean, mx = epe_stats(w, ow)
    assert mean <= 1e-6 and mx <= 1e-4, (mean, mx)
    # and it matters: the second batch differs from a run with a fixed reference
    _, w_fixed = F.compensate_arr_3D(video, ref, F.OFOptions(**{**opts.model_dump(), "update_reference": False}))
    assert np.abs(w[3:] - w_fixed[3:]).max() > 1e-3


def test_per_pair_functions_are_thread_safe(backend, golden):
    """The reference's get_displacement / imregister_wrapper are pure functions that ThreadingExecutor3D calls from
    worker threads (threading_3d.py:211-225).  Here every thread gets its own context; concurrent calls (different
    moving volumes, two different fixed volumes, so the cached-pyramid path is exercised too) equal the serial ones."""
    import threading
    import flowreg3d_b200 as F
    g = golden("flow_small")
    fixed, moving = g["fixed"][:12, :32, :36].astype(np.float32), g["moving"][:12, :32, :36].astype(np.float32)
    kw = dict(alpha=(0.25, 0.3, 0.2), update_lag=5, iterations=10, min_level=1, levels=100, eta=0.8, a_smooth=1.0,
              a_data=0.45)
    jobs = [(fixed, moving), (fixed, np.roll(moving, 1, 2)), (np.roll(fixed, 1, 1), moving), (fixed, np.roll(moving, 2, 1))]

    def run(job):
        f, m = job
        flow = F.get_displacement(f, m, **kw)
        f32 = flow.astype(np.float32)
        return flow, F.imregister_wrapper(m, f32[..., 0], f32[..., 1], f32[..., 2], f, "cubic")

    serial = [run(j) for j in jobs]
    out = [None] * len(jobs)
    errs = []

    def worker(idx):
        try:
            for rep in range(2):                 # second round: cached contexts and pyramids
                for k in range(idx, len(jobs), 2):
                    out[k] = run(jobs[k])
        except Exception as e:                   # pragma: no cover
            errs.append(e)

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for (fa, ra), (fb, rb) in zip(serial, out):
        assert np.array_equal(fa, fb) and np.array_equal(ra, rb)


def test_run_stream_from_a_reader_like_iterator(backend, golden):
    """SequenceCorrector.run_stream: batches pulled from an iterator (the reference's reader.read_batch() pattern,
    util/io/_base_3d.py:230-251) and handed to a sink in order (writer.write_frames) give exactly what the in-memory
    entry point gives."""
    import flowreg3d_b200 as F
    g = golden("sequence")
    v, r = g["video"][:, :12, :24, :28], g["ref"][:12, :24, :28]
    opts = F.OFOptions(min_level=3, iterations=6, update_lag=3, buffer_size=3, weight=[0.5, 0.5], output_typename=None)
    reg1, w1 = F.compensate_arr_3D(v, r, opts)

    class Reader:                      # the reference's reader protocol: read_batch() -> array or None
        def __init__(self, arr, bs):
            self.arr, self.bs, self.at = arr, bs, 0

        def read_batch(self):
            if self.at >= self.arr.shape[0]:
                return None
            b = self.arr[self.at:self.at + self.bs]
            self.at += self.bs
            return b

    got_reg, got_w, order = [], [], []
    seq = F.SequenceCorrector(r, opts)
    try:
        rd = Reader(v, 3)

        def batches():                 # `while reader.has_batch(): yield reader.read_batch()`; None ends the stream
            while True:
                yield rd.read_batch()
        seq.run_stream(batches(),
                       sink=lambda k, reg, fl: (order.append(k), got_reg.append(reg.numpy().copy()), got_w.append(fl.numpy().copy())))
    finally:
        seq.close()
    assert order == list(range(len(order))) and len(order) == 3
    assert np.array_equal(np.concatenate(got_w, 0), w1) and np.array_equal(np.concatenate(got_reg, 0), reg1)
