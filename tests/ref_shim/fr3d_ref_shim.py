"""pytest plugin (-p fr3d_ref_shim): installs `flowreg3d.*` alias modules that resolve to flowreg3d_b200, so that the
reference's OWN test files run unmodified against this package (tests/test_reference_suite.py drives it in a
subprocess).  Only names the reference's tests import are aliased; optional I/O packages that are absent become empty
stub modules (tests that really need them fail and are listed as such by the driver)."""
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import flowreg3d_b200 as F  # noqa: E402
from flowreg3d_b200 import _lib, io_factory, options as O, recording as R  # noqa: E402

_lib._select_for_tests(str(ROOT / "tests" / "emu" / "_build" / "libfr3d_emu.so"), emulator=True)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


def _alias(name, module):
    sys.modules[name] = module
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, module)


class RuntimeContext:
    """Stand-in for flowreg3d._runtime.RuntimeContext (the reference's registry of CPU executors) so that test modules
    importing it can be collected; this package has no worker pools to select from."""
    _store = {"available_parallelization": set()}

    @classmethod
    def init(cls, force=False):
        return None

    @classmethod
    def get(cls, key, default=None):
        return cls._store.get(key, default)

    @classmethod
    def get_parallelization_executor(cls, name):
        return None

    @classmethod
    def get_available_parallelization(cls):
        return set()


from flowreg3d_b200 import compensate as C  # noqa: E402

_mod("flowreg3d", get_displacement=F.get_displacement, imregister_wrapper=F.imregister_wrapper)
_mod("flowreg3d._runtime", RuntimeContext=RuntimeContext)
_mod("flowreg3d.motion_correction")
_alias("flowreg3d.motion_correction.OF_options_3D", O)          # the real modules: mock.patch on the alias reaches them
_alias("flowreg3d.motion_correction.compensate_arr_3D", C)
_alias("flowreg3d.motion_correction.compensate_recording_3D", R)
_mod("flowreg3d.util")
from flowreg3d_b200 import core as K, xcorr as X  # noqa: E402

_alias("flowreg3d.util.xcorr_prealignment", X)
_mod("flowreg3d.core")
_alias("flowreg3d.core.optical_flow_3d", K)
_mod("flowreg3d.util.io")
_alias("flowreg3d.util.io.factory", io_factory)
_alias("flowreg3d.util.io._arr_3d", R)
for name in ("tifffile", "h5py", "hdf5storage"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
