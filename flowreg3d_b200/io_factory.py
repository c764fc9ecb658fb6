"""Reader / writer factory with the reference's call signatures (util/io/factory.py: `get_video_file_reader(input_source,
buffer_size, bin_size, **kwargs)`, `get_video_file_writer(file_path, output_format, **kwargs)`), limited to the formats
that need nothing but numpy: arrays and `.npy` files (memory maps).  Objects that already follow the reader protocol
are returned as they are, so the reference's HDF5 / TIFF / MAT readers pass through."""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np


def _is_reader(x) -> bool:
    return hasattr(x, "has_batch") and hasattr(x, "read_batch")


def get_video_file_reader(input_source, buffer_size: int = 10, bin_size: int = 1, **kwargs):
    from .recording import ArrayReader3D, NpyFileReader3D
    if isinstance(input_source, np.ndarray):
        return ArrayReader3D(input_source, buffer_size, bin_size)
    if _is_reader(input_source):
        return input_source
    if isinstance(input_source, (str, os.PathLike)) and str(input_source).lower().endswith(".npy"):
        return NpyFileReader3D(input_source, buffer_size, bin_size)
    if input_source is None:
        raise ValueError("no input: options.input_file is not set")
    raise NotImplementedError(
        f"no reader for {input_source!r} in this package (arrays and .npy files only): pass a reader object with the "
        "reference's protocol instead, e.g. flowreg3d.util.io.factory.get_video_file_reader(...)")


def get_video_file_writer(file_path, output_format, frame_count=None, **kwargs):
    from .recording import ArrayWriter3D, NpyFileWriter3D
    fmt = str(getattr(output_format, "value", output_format)).upper()
    if fmt == "ARRAY":
        return ArrayWriter3D()
    if fmt == "NPY":
        if frame_count is None:
            raise ValueError("an .npy writer needs frame_count (the header fixes the shape up front)")
        return NpyFileWriter3D(Path(file_path), frame_count)
    raise NotImplementedError(
        f"no writer for output_format {fmt} in this package (ARRAY and NPY only): pass a writer object with "
        "write_frames / close instead, e.g. flowreg3d.util.io.factory.get_video_file_writer(...)")
