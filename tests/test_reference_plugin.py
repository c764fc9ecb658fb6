"""Boundary B1 against the UNMODIFIED reference: flowreg3D's own BatchMotionCorrector driving the B200
executor plugin (kernel-logic emulator here; the CUDA library on a GPU box) and its own sequential
executor, on the same input.  Runs only where the reference source tree is present (the build
container); it cannot travel to the GPU box and is skipped there."""
import os
import sys
import types
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference/src")
pytestmark = pytest.mark.skipif(not (REF / "flowreg3d").is_dir(), reason="reference source tree not present")


@pytest.fixture(scope="module")
def reference():
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    sys.path.insert(0, str(REF))
    for m in ("tifffile", "h5py", "hdf5storage"):      # optional I/O packages the array path never touches
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    from flowreg3d.motion_correction.compensate_arr_3D import compensate_arr_3D
    from flowreg3d.motion_correction.OF_options_3D import OFOptions
    from flowreg3d._runtime import RuntimeContext
    import flowreg3d.motion_correction.parallelization  # noqa: F401  (registers the reference executors)
    yield types.SimpleNamespace(compensate_arr_3D=compensate_arr_3D, OFOptions=OFOptions, RuntimeContext=RuntimeContext)
    sys.path.remove(str(REF))


def test_reference_pipeline_with_b200_executor(emu_backend, golden, reference, monkeypatch):
    import importlib
    import flowreg3d_b200.executor as ex
    importlib.reload(ex)                               # pick up the reference's BaseExecutor3D as the base class
    from flowreg3d.motion_correction.parallelization.base_3d import BaseExecutor3D
    assert issubclass(ex.B200Executor3D, BaseExecutor3D)
    assert ex.B200Executor3D.register()
    assert "b2003d" in reference.RuntimeContext.get_available_parallelization()
    g = golden("sequence")
    video, ref = g["video"][:5, :12, :24, :28], g["ref"][:12, :24, :28]
    kw = dict(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=8, update_lag=4, buffer_size=3,
              weight=[0.5, 0.5])
    import flowreg3d.motion_correction.compensate_recording_3D as cr
    seen = []
    orig = cr.BatchMotionCorrector._setup_executor

    def pick(name):
        def setup(self):
            self.config.parallelization = name
            orig(self)
            seen.append(type(self.executor).__name__)
        return setup

    monkeypatch.setattr(cr.BatchMotionCorrector, "_setup_executor", pick("b200"))
    reg_b, w_b = reference.compensate_arr_3D(video, ref, reference.OFOptions(**kw))
    monkeypatch.setattr(cr.BatchMotionCorrector, "_setup_executor", pick("sequential"))
    reg_s, w_s = reference.compensate_arr_3D(video, ref, reference.OFOptions(**kw))
    assert seen == ["B200Executor3D", "SequentialExecutor3D"]
    e = np.sqrt(((w_b.astype(np.float64) - w_s) ** 2).sum(-1))
    assert e.mean() <= 1e-4 and e.max() <= 5e-3, (e.mean(), e.max())      # tolerance: 0.01 / 0.05
    assert np.linalg.norm(reg_b.astype(np.float64) - reg_s) <= 1e-5 * np.linalg.norm(reg_s)   # tolerance: 1e-4


def test_reference_executor_consistency_bar(emu_backend, reference, monkeypatch):
    """The reference's own cross-executor test (tests/motion_correction/test_parallelization.py:152-198: noise
    video, quality 'fast', levels 2, 5 iterations, rtol 1e-5 / atol 1e-6) with the B200 executor against
    sequential3d -- a registered '*3d' executor is swept into that test automatically."""
    import importlib
    import flowreg3d_b200.executor as ex
    importlib.reload(ex)
    assert ex.B200Executor3D.register()
    from flowreg3d.motion_correction.compensate_recording_3D import BatchMotionCorrector, RegistrationConfig
    from flowreg3d.motion_correction.OF_options_3D import OutputFormat
    T, Z, Y, X, C = 8, 6, 12, 12, 2
    np.random.seed(42)
    video = np.random.rand(T, Z, Y, X, C).astype(np.float32)
    ref = np.mean(video[:2], axis=0)
    results = {}
    for name in ("sequential3d", "b2003d"):
        options = reference.OFOptions(quality_setting="fast", levels=2, iterations=5)
        options.input_file = video.copy()
        options.reference_frames = ref.copy()
        options.output_format = OutputFormat.ARRAY
        options.save_w = True
        options.save_meta_info = False
        comp = BatchMotionCorrector(options, RegistrationConfig(parallelization=name, n_jobs=2))
        comp.run()
        assert type(comp.executor).__name__ == ("B200Executor3D" if name == "b2003d" else "SequentialExecutor3D")
        results[name] = comp.video_writer.get_array()
    np.testing.assert_allclose(results["b2003d"], results["sequential3d"], rtol=1e-5, atol=1e-6)


def test_reference_pipeline_with_cc_initialization(emu_backend, golden, reference, monkeypatch):
    """OFOptions(cc_initialization=True) through the UNMODIFIED reference pipeline: the B200 executor against the
    reference's sequential executor on a single-channel recording with a rigid offset.  The reference's executor
    needs skimage.registration.phase_cross_correlation (absent here): it is given the oracle's restatement, the B200
    executor uses its own device path (no scikit-image)."""
    import importlib
    import flowreg3d_b200.executor as ex
    importlib.reload(ex)
    assert ex.B200Executor3D.register()
    from oracle import oracle as O
    from oracle import xcorr as OX
    sk, skr = types.ModuleType("skimage"), types.ModuleType("skimage.registration")
    skr.phase_cross_correlation = lambda a, b, **kw: (OX.phase_cross_correlation(a, b, **kw), None, None)
    sk.registration = skr
    monkeypatch.setitem(sys.modules, "skimage", sk)
    monkeypatch.setitem(sys.modules, "skimage.registration", skr)
    from tests_inputs import synth_volume
    shape = (12, 40, 48)
    ref = synth_volume(shape, 3)
    r64 = ref.astype(np.float64)
    rng = np.random.default_rng(5)
    video = []
    for t in range(4):
        d = np.array([4.0, -3.0, 1.0]) + 0.3 * t
        mv = O.imregister_wrapper(r64, np.full(shape, -d[0]), np.full(shape, -d[1]), np.full(shape, -d[2]), r64, "linear")
        video.append(mv + 0.002 * rng.standard_normal(shape).astype(np.float32))
    video = np.stack(video, 0)[..., None].astype(np.float32)
    kw = dict(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=8, update_lag=4, buffer_size=2,
              cc_initialization=True, cc_hw=(32, 40), cc_up=10)
    import flowreg3d.motion_correction.compensate_recording_3D as cr
    orig = cr.BatchMotionCorrector._setup_executor
    seen = []

    def pick(name):
        def setup(self):
            self.config.parallelization = name
            orig(self)
            seen.append(type(self.executor).__name__)
        return setup

    monkeypatch.setattr(cr.BatchMotionCorrector, "_setup_executor", pick("b200"))
    reg_b, w_b = reference.compensate_arr_3D(video, ref[..., None], reference.OFOptions(**kw))
    monkeypatch.setattr(cr.BatchMotionCorrector, "_setup_executor", pick("sequential"))
    reg_s, w_s = reference.compensate_arr_3D(video, ref[..., None], reference.OFOptions(**kw))
    assert seen == ["B200Executor3D", "SequentialExecutor3D"]
    e = np.sqrt(((w_b.astype(np.float64) - w_s) ** 2).sum(-1))
    assert e.mean() <= 1e-4 and e.max() <= 5e-3, (e.mean(), e.max())      # tolerance: 0.01 / 0.05
    assert np.linalg.norm(reg_b.astype(np.float64) - reg_s) <= 1e-5 * np.linalg.norm(reg_s)   # tolerance: 1e-4
    core = (slice(None), slice(3, -3), slice(8, -8), slice(8, -8))
    # sanity only (how close 8 iterations on a 3-level pyramid get to the ground truth is the method's property)
    assert np.abs(w_b[core].reshape(4, -1, 3).mean(1)[0] - np.array([4.0, -3.0, 1.0])).max() < 1.0
