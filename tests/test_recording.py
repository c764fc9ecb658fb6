"""Reader -> GPU -> writer entry (flowreg3d_b200.recording): the reference's compensate_recording / BatchMotionCorrector
(motion_correction/compensate_recording_3D.py:32-613) over the streaming path, against compensate_arr_3D on the same
input, the statistics' numpy definitions, the reference's own reader / writer objects and its own pipeline."""
import os
import sys
import types
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/src")
KW = dict(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=8, update_lag=4, buffer_size=3,
          weight=[0.5, 0.5])


def _case(golden):
    g = golden("sequence")
    return g["video"][:7, :12, :24, :28], g["ref"][:12, :24, :28]


def _stats(w):
    # compensate_recording_3D.py:488-508
    mag = np.sqrt(w[..., 0] ** 2 + w[..., 1] ** 2 + w[..., 2] ** 2)
    div = [float(np.mean(np.gradient(w[t, ..., 0], axis=2) + np.gradient(w[t, ..., 1], axis=1) +
                         np.gradient(w[t, ..., 2], axis=0))) for t in range(w.shape[0])]
    tr = [float(np.sqrt(sum(float(np.mean(w[t, ..., c])) ** 2 for c in range(3)))) for t in range(w.shape[0])]
    return mag.mean(axis=(1, 2, 3)), mag.max(axis=(1, 2, 3)), np.array(div), np.array(tr)


def test_compensate_recording_matches_array_entry(backend, golden, tmp_path):
    import flowreg3d_b200 as F
    video, ref = _case(golden)
    reg_a, w_a = F.compensate_arr_3D(video, ref, F.OFOptions(output_typename=None, **KW))
    seen = []
    opts = F.OFOptions(input_file=video, output_format="ARRAY", output_path=str(tmp_path / "out"), reference_frames=ref,
                       save_w=True, save_meta_info=True, **KW)
    pipe = F.BatchMotionCorrector(opts, F.RegistrationConfig(n_jobs=1))
    pipe.register_progress_callback(lambda c, t: seen.append((c, t)))
    pipe.register_progress_callback(lambda c, t: 1 / 0)            # a failing callback only warns (:158-162)
    with pytest.warns(UserWarning, match="Progress callback error"):
        ref_used = pipe.run()
    assert ref_used.dtype == np.float64 and np.array_equal(ref_used, ref.astype(np.float64))
    reg_r, w_r = pipe.video_writer.get_array(), pipe.w_writer.get_array()
    assert reg_r.dtype == video.dtype and reg_r.shape == video.shape and w_r.dtype == np.float32
    assert np.array_equal(w_r, w_a) and np.array_equal(reg_r, reg_a)
    assert seen == [(3, 7), (6, 7), (7, 7)]
    md, mx, dv, tr = _stats(w_r)
    assert np.allclose(pipe.mean_disp, md, rtol=1e-5, atol=1e-7) and np.allclose(pipe.max_disp, mx, rtol=1e-6)
    assert np.allclose(pipe.mean_div, dv, rtol=1e-4, atol=1e-7) and np.allclose(pipe.mean_translation, tr, rtol=1e-5, atol=1e-7)
    st = np.load(tmp_path / "out" / "statistics.npz")
    assert np.array_equal(st["mean_disp"], np.array(pipe.mean_disp)) and len(st["mean_div"]) == 7
    assert np.array_equal(np.load(tmp_path / "out" / "reference_frame.npy"), ref_used)
    assert pipe.w_init.shape == ref.shape[:3] + (3,) and pipe.w_init.dtype == np.float64
    assert np.allclose(pipe.w_init, w_r[6:].mean(axis=0), atol=1e-6)          # chained from the last batch (:481-485)


def test_compensate_recording_npy_files_and_reference_indices(emu_backend, golden, tmp_path):
    """Out-of-core form: .npy in (memory-mapped, integer dtype), .npy out, reference = mean of reader[indices]."""
    import flowreg3d_b200 as F
    video, _ = _case(golden)
    raw = np.round(video * 1000).astype(np.uint16)
    np.save(tmp_path / "in.npy", raw)
    opts = F.OFOptions(input_file=str(tmp_path / "in.npy"), output_format="NPY", output_path=str(tmp_path / "res"),
                       reference_frames=[0, 2, 5], save_w=True, save_meta_info=False, **KW)
    ref_used = F.compensate_recording(opts)
    assert np.array_equal(ref_used, raw[[0, 2, 5]].mean(axis=0))
    out, w = np.load(tmp_path / "res" / "compensated.npy"), np.load(tmp_path / "res" / "w.npy")
    assert out.dtype == np.uint16 and out.shape == raw.shape and w.shape == raw.shape[:4] + (3,)
    assert not (tmp_path / "res" / "statistics.npz").exists()
    reg_a, w_a = F.compensate_arr_3D(raw, ref_used, F.OFOptions(output_typename=None, **KW))
    assert np.array_equal(out, reg_a) and np.array_equal(w, w_a)
    with pytest.raises(IndexError):
        F.compensate_recording(F.OFOptions(input_file=raw, output_format="ARRAY", output_path=str(tmp_path / "x"),
                                           reference_frames=[0, 99], **KW))
    with pytest.raises(NotImplementedError):
        F.compensate_recording(F.OFOptions(input_file=raw, output_format="HDF5", output_path=str(tmp_path / "x"),
                                           reference_frames=[0], **KW))


def test_registration_config_and_progress_protocol():
    """The reference's own unit tests for this surface (tests/motion_correction/test_compensate_recording_3D.py:24-47,
    435-481, 532-560), on this package's objects."""
    from unittest.mock import patch
    import flowreg3d_b200 as F
    c = F.RegistrationConfig()
    assert (c.n_jobs, c.batch_size, c.verbose, c.parallelization) == (-1, 10, False, None)
    c = F.RegistrationConfig(n_jobs=2, batch_size=5, verbose=True, parallelization="threading3d")
    assert (c.n_jobs, c.batch_size, c.verbose, c.parallelization) == (2, 5, True, "threading3d")
    opts = F.OFOptions(input_file="dummy.h5", output_path="unused", quality_setting="fast", levels=2, iterations=5)
    pipe = F.BatchMotionCorrector(opts)
    calls = []
    cb = lambda cur, tot: calls.append((cur, tot))  # noqa: E731
    pipe.register_progress_callback(cb)
    pipe.register_progress_callback(cb)
    assert len(pipe.progress_callbacks) == 1
    pipe._total_frames = 100
    for n in (10, 15, 20, 25, 30):
        pipe._notify_progress(n)
    pipe._notify_progress(7, task_id="other")          # other tasks are tracked, not reported
    assert calls == [(10, 100), (25, 100), (45, 100), (70, 100), (100, 100)]
    ref = np.random.rand(4, 16, 16, 2).astype(np.float32)
    with patch.object(F.BatchMotionCorrector, "run") as run:
        run.return_value = ref
        cfg = F.RegistrationConfig(n_jobs=1, batch_size=5, verbose=True, parallelization="sequential3d")
        assert np.array_equal(F.compensate_recording(opts, ref, cfg), ref)
        run.assert_called_once_with(ref)
    with patch.object(F.BatchMotionCorrector, "run") as run:
        run.return_value = ref
        assert np.array_equal(F.compensate_recording(opts), ref)
        run.assert_called_once_with(None)
    with pytest.raises(NotImplementedError):            # an .h5 path needs the reference's reader object
        F.BatchMotionCorrector(F.OFOptions(input_file="dummy.h5", output_path="/tmp/fr3d_unused_out"))._setup_io()


def test_array_reader_protocol():
    from flowreg3d_b200.recording import ArrayReader3D, ArrayWriter3D, NpyFileWriter3D
    a = np.arange(7 * 2 * 3 * 4 * 1, dtype=np.float32).reshape(7, 2, 3, 4, 1)
    r = ArrayReader3D(a, buffer_size=3)
    assert len(r) == 7 and r.shape == (7, 2, 3, 4, 1) and r.has_batch()
    got = []
    while r.has_batch():
        got.append(r.read_batch())
    assert [b.shape[0] for b in got] == [3, 3, 1] and r.read_batch() is None
    assert np.array_equal(np.concatenate(got), a) and np.array_equal(r[[0, -1]], a[[0, 6]])
    r.reset()
    assert r.has_batch() and ArrayReader3D(a[0]).shape == (1, 2, 3, 4, 1) and ArrayReader3D(a[0, ..., 0]).shape == (1, 2, 3, 4, 1)
    with pytest.raises(IndexError):
        r[7]
    with pytest.raises(ValueError):
        ArrayReader3D(np.zeros((2, 2)))
    w = ArrayWriter3D()
    assert w.get_array() is None
    w.write_frames(a[:2])
    w.write_frames(a[2])
    assert np.array_equal(w.get_array(), a[:3])
    with pytest.raises(ValueError):
        NpyFileWriter3D("/tmp/never_written.npy", 1).write_frames(a[:2])


@pytest.mark.skipif(not (REF / "flowreg3d").is_dir(), reason="reference source tree not present")
def test_against_the_reference_pipeline_and_with_its_reader_writer(emu_backend, golden, tmp_path):
    """(i) The reference's own ArrayReader3D / ArrayWriter3D objects drive this pipeline unchanged; (ii) the reference's
    compensate_recording (sequential executor) on the same options agrees within the path's tolerance, statistics
    included."""
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.path.insert(0, str(REF))
    try:
        for m in ("tifffile", "h5py", "hdf5storage"):
            if m not in sys.modules:
                try:
                    __import__(m)
                except Exception:
                    sys.modules[m] = types.ModuleType(m)
        from flowreg3d.util.io._arr_3d import ArrayReader3D as RefReader, ArrayWriter3D as RefWriter
        from flowreg3d.motion_correction.OF_options_3D import OFOptions as RefOptions, OutputFormat
        from flowreg3d.motion_correction.compensate_recording_3D import (BatchMotionCorrector as RefPipe,
                                                                          RegistrationConfig as RefConfig)
        import flowreg3d.motion_correction.parallelization  # noqa: F401
        import flowreg3d_b200 as F
        video, ref = _case(golden)
        ours = F.BatchMotionCorrector(
            F.OFOptions(input_file=RefReader(video, buffer_size=3), output_path=str(tmp_path / "a"), reference_frames=ref,
                        save_w=True, save_meta_info=False, **KW),
            video_writer=RefWriter(), w_writer=RefWriter())
        ours.run()
        reg_o, w_o = ours.video_writer.get_array(), ours.w_writer.get_array()
        theirs = RefPipe(RefOptions(input_file=video, output_format=OutputFormat.ARRAY, output_path=str(tmp_path / "b"),
                                    reference_frames=ref, save_w=True, save_meta_info=False, **KW),
                         RefConfig(n_jobs=1, parallelization="sequential", verbose=True))
        theirs.run()
        reg_t, w_t = theirs.video_writer.get_array(), theirs.w_writer.get_array()
        assert reg_o.shape == reg_t.shape and reg_o.dtype == reg_t.dtype
        e = np.sqrt(((w_o.astype(np.float64) - w_t) ** 2).sum(-1))
        assert e.mean() <= 1e-4 and e.max() <= 5e-3, (e.mean(), e.max())            # tolerance: 0.01 / 0.05
        assert np.linalg.norm(reg_o.astype(np.float64) - reg_t) <= 1e-5 * np.linalg.norm(reg_t)     # tolerance: 1e-4
        for name in ("mean_disp", "max_disp", "mean_div", "mean_translation"):
            a, b = np.array(getattr(ours, name)), np.array(getattr(theirs, name))
            assert a.shape == b.shape and np.allclose(a, b, rtol=5e-3, atol=1e-4), (name, a, b)
    finally:
        sys.path.remove(str(REF))
