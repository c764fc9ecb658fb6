"""flowreg3d_b200 -- B200-native (sm_100a) drop-in for flowreg3D's dense 3-D optical-flow
registration path.  Python host code + hand-written CUDA kernels behind a C ABI (libfr3d.so).

Reference-facing surface (same names and argument meaning as FlowRegSuite/flowreg3D):
    get_displacement, imregister_wrapper      core/optical_flow_3d.py
    OFOptions                                 motion_correction/OF_options_3D.py
    compensate_arr_3D                         motion_correction/compensate_arr_3D.py
    compensate_recording, BatchMotionCorrector, RegistrationConfig
                                              motion_correction/compensate_recording_3D.py
    ArrayReader3D, ArrayWriter3D              util/io/_arr_3d.py (+ .npy memory-map reader / writer)
    B200Executor3D (BaseExecutor3D plugin)    motion_correction/parallelization/base_3d.py
"""
from .core import Context, Registration, get_displacement, imregister_wrapper  # noqa: F401
from .plan import FlowParams  # noqa: F401
from .options import OFOptions  # noqa: F401
from .executor import B200Executor3D  # noqa: F401
from .compensate import SequenceCorrector, compensate_arr_3D, compensate_arr_3D_sharded  # noqa: F401
from .recording import (ArrayReader3D, ArrayWriter3D, BatchMotionCorrector, NpyFileReader3D,  # noqa: F401
                        NpyFileWriter3D, RegistrationConfig, compensate_recording)

__all__ = ["get_displacement", "imregister_wrapper", "OFOptions", "compensate_arr_3D",
           "compensate_arr_3D_sharded", "B200Executor3D", "SequenceCorrector", "Registration",
           "FlowParams", "Context", "compensate_recording", "BatchMotionCorrector", "RegistrationConfig",
           "ArrayReader3D", "ArrayWriter3D", "NpyFileReader3D", "NpyFileWriter3D"]
