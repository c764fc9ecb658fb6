"""Host-side logic and the C-ABI surface (CPU only; no compute call on the CUDA library)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "fr3d.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fr3d_[a-z0-9_]+)\s*\(", txt)))


def test_cuda_library_exports_every_declared_symbol():
    """libfr3d.so (sm_100a build) loads without a GPU and exports the whole header."""
    from flowreg3d_b200 import build, _lib
    so = build.build()
    lib = ctypes.CDLL(str(so))
    names = _declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fr3d.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    lib.fr3d_abi_version.restype = ctypes.c_int
    assert lib.fr3d_abi_version() == _lib.ABI_VERSION


def test_cuda_library_contains_sm100a_code():
    import subprocess
    from flowreg3d_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", str(build.build())], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_library(monkeypatch, tmp_path):
    """A missing libfr3d.so raises; nothing in the environment can redirect the product loader (the kernel-logic
    emulator is installed by test code only, through _lib._select_for_tests in its own process)."""
    from flowreg3d_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "_test_library", None)
    monkeypatch.setattr(_lib, "HERE", tmp_path)          # a checkout without the built library
    with pytest.raises(ImportError):
        _lib.load()
    monkeypatch.setattr(_lib, "HERE", ROOT / "flowreg3d_b200")
    for var in ("FR3D_LIBRARY_OVERRIDE", "FR3D_LIBRARY_VARIANT"):
        monkeypatch.setenv(var, str(tmp_path / "libfr3d_emu.so"))
    assert _lib.library_path() == ROOT / "flowreg3d_b200" / "libfr3d.so" and not _lib.is_emulator()
    src = (ROOT / "flowreg3d_b200" / "_lib.py").read_text() + (ROOT / "flowreg3d_b200" / "device.py").read_text()
    assert "os.environ" not in src and "getenv" not in src
    # ... and the CUDA library reads no environment either: every tuning aid is an fr3d_set_option value
    for f in (ROOT / "flowreg3d_b200" / "csrc").iterdir():
        assert "getenv" not in f.read_text(), f


def test_product_never_imports_oracle():
    for py in (ROOT / "flowreg3d_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), py


def test_level_schedule_matches_reference_sizes():
    from flowreg3d_b200 import plan
    # SURVEY.md section 8 (measured on the reference)
    s, ml = plan.level_schedule((32, 512, 512), 0.8, 100, 5)
    assert [x[1] for x in s] == [(8, 134, 134), (10, 168, 168)] and ml == 5
    s, ml = plan.level_schedule((64, 128, 128), 0.8, 100, 5)
    assert s[0][1] == (9, 17, 17) and s[-1][1] == (21, 42, 42) and len(s) == 5
    s, ml = plan.level_schedule((32, 512, 512), 0.8, 100, 6)     # min_level >= top is clamped to top-1
    assert ml == 5 and len(s) == 2
    # level sizes and per-level alpha scaling recorded from the LIVE reference driver (tests/golden/gen_golden.py:
    # gen_schedule patches the per-level work out of flowreg3d.core.optical_flow_3d.get_displacement)
    import json
    for case in json.loads((ROOT / "tests" / "golden" / "schedule.json").read_text()):
        shape, eta = tuple(case["shape"]), case["eta"]
        for impl in (plan.level_schedule, O.level_schedule):
            s, ml = impl(shape, eta, case["levels"], case["min_level"])
            assert [list(x[1]) for x in s] == case["sizes"], (impl.__module__, case)
            alpha = [[(1.0 if i == ml else eta ** (-0.5 * i)) * a for a in (1.0, 2.0, 3.0)] for i, _ in s]
            assert np.allclose(alpha, case["alpha"], rtol=1e-15, atol=0), (impl.__module__, case)


def test_gaussian_kernel_matches_scipy():
    from scipy.ndimage import gaussian_filter1d
    from flowreg3d_b200 import plan
    for s in (0.5, 1.0, 1.5, 2.3):
        w = plan.gaussian_half_kernel(s)
        x = np.zeros(2 * len(w) + 5)
        x[len(x) // 2] = 1.0
        k = gaussian_filter1d(x, s, mode="reflect", truncate=4.0)
        c = len(x) // 2
        assert np.array_equal(k[c:c + len(w)], w)
    assert np.array_equal(plan.gaussian_half_kernel(0.0), [1.0])
    assert len(plan.gaussian_half_kernel(0.1)) == 1
    assert np.array_equal(plan.sigma_t([[1, 1, 1, 0.5], [1, 1, 1, 0.1]], 3), [0.5, 0.1, 0.1])
    assert np.array_equal(plan.sigma_zyx([[1, 2, 3, 0.1]], 2), [[3, 2, 1], [3, 2, 1]])


def test_ofoptions_defaults_and_validators():
    """Mirrors reference tests/motion_correction/test_OF_options_3D.py:28-42 and the validators."""
    from flowreg3d_b200 import OFOptions
    o = OFOptions()
    assert o.alpha == (0.25, 0.25, 0.25) and o.buffer_size == 10 and o.min_level == 5
    assert o.sigma == [[1.0, 1.0, 1.0, 0.1]] * 2 and o.levels == 100 and o.iterations == 100
    assert o.update_lag == 5 and o.a_smooth == 1.0 and o.a_data == 0.45 and o.eta == 0.8
    assert o.interpolation_method.value == "cubic" and o.quality_setting.value == "custom"
    assert OFOptions(alpha=2).alpha == (2.0, 2.0, 2.0)
    assert OFOptions(alpha=(1, 3)).alpha == (1.0, 1.0, 3.0)
    with pytest.raises(Exception):
        OFOptions(alpha=-1)
    with pytest.raises(Exception):
        OFOptions(backend="cuda")                       # extra="forbid", like the reference
    assert OFOptions(weight=[2, 6]).weight == [0.25, 0.75]
    assert OFOptions(sigma=[1, 2, 0.1]).sigma == [[1.0, 2.0, 1.0, 0.1]]
    assert OFOptions(min_level=-1, quality_setting="fast").effective_min_level == 6
    assert OFOptions(min_level=-1, quality_setting="balanced").effective_min_level == 4
    assert OFOptions(min_level=-1, quality_setting="quality").effective_min_level == 0
    d = o.to_dict()
    assert set(d) == {"alpha", "weight", "levels", "min_level", "eta", "iterations", "update_lag", "a_data",
                      "a_smooth", "const_assumption"} and d["min_level"] == 5
    c = o.copy()
    c.buffer_size = 3
    assert o.buffer_size == 10
    assert OFOptions(weight=[0.5, 0.5]).get_weight_at(0, 1) == 1.0   # truncation + renormalisation
    assert OFOptions(weight=[0.2, 0.8]).get_weight_at(1, 2) == pytest.approx(0.8)
    assert OFOptions(weight=[0.2, 0.8]).get_weight_at(2, 3) == pytest.approx(1 / 3)


def test_shard_bounds_cover_every_frame_once():
    from flowreg3d_b200.compensate import shard_bounds
    for n in (0, 1, 7, 10, 16, 200):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_executor_registers_with_reference_runtime_when_present():
    import flowreg3d_b200 as F
    try:
        import flowreg3d  # noqa: F401
    except Exception:
        assert F.B200Executor3D.register() is False
        return
    from flowreg3d._runtime import RuntimeContext
    assert F.B200Executor3D.register() is True
    assert RuntimeContext.get_parallelization_executor("b2003d") is F.B200Executor3D


def test_options_accept_every_reference_field_and_reference_from_indices(emu_backend):
    """OFOptions takes every keyword the reference's OFOptions takes (OF_options_3D.py:141-231; file I/O fields are
    carried and ignored on the array path), still forbids unknown ones, and get_reference_frame reproduces the 3-D
    branch of the reference (:496-503): the mean over the listed frames, numpy's arithmetic."""
    import flowreg3d_b200 as F
    o = F.OFOptions(input_file=None, input_dim_order="TZYXC", output_path="results", output_format="HDF5",
                    output_file_name=None, channel_idx=None, bin_size=1, n_references=1, min_frames_per_reference=20,
                    save_meta_info=False, save_w=True, save_valid_mask=False, save_valid_idx=False,
                    naming_convention="default", preproc_funct=None, reference_frames=[1, 3])
    with pytest.raises(Exception):
        F.OFOptions(no_such_option=1)
    v = np.random.default_rng(0).random((5, 6, 10, 12, 2)).astype(np.float32)
    r = o.get_reference_frame(v)
    assert np.array_equal(r, v[[1, 3]].mean(axis=0)) and r.dtype == np.float32
    with pytest.raises(IndexError):
        F.OFOptions(reference_frames=[7]).get_reference_frame(v)
    arr = np.ones((6, 10, 12, 2))
    assert F.OFOptions(reference_frames=arr).get_reference_frame(v) is not None
    with pytest.warns(UserWarning):
        many = F.OFOptions(reference_frames=[0, 1], n_references=3).get_reference_frame(v)
    assert len(many) == 3 and np.array_equal(many[0], v[[0, 1]].mean(axis=0))
    # compensate_arr_3D(c_ref=None): the fixed volume the options describe
    opts = F.OFOptions(reference_frames=[0, 2], min_level=2, iterations=4, update_lag=2, buffer_size=3, weight=[0.5, 0.5])
    small = v[:3, :, :, :, :]
    reg_a, w_a = F.compensate_arr_3D(small, None, opts)
    reg_b, w_b = F.compensate_arr_3D(small, small[[0, 2]].mean(axis=0), opts)
    assert np.array_equal(reg_a, reg_b) and np.array_equal(w_a, w_b)
