"""Reader -> GPU -> writer pipeline: the reference's top-level entry `compensate_recording(options, reference_frame,
config)` and its `BatchMotionCorrector` (motion_correction/compensate_recording_3D.py:32-613) for recordings that are
streamed batch by batch, on top of `SequenceCorrector.run_stream` (pinned double-buffered copies, the batch being
computed overlapping the host->device copy of the next and the device->host copy of the previous one).

Readers and writers are duck-typed on the reference's protocol (util/io/_base_3d.py): a reader has `has_batch()`,
`read_batch() -> (T,Z,Y,X,C)`, `reader[list_of_indices]`, `len(reader)`, `close()`; a writer has `write_frames(frames)`
and `close()`.  The reference's own reader / writer objects (HDF5, TIFF, MAT: they need h5py / tifffile / hdf5storage)
can be passed as they are -- `options.input_file = reader`, `video_writer=writer` -- this package ships the two
formats that need nothing but numpy: in-memory arrays (`ArrayReader3D` / `ArrayWriter3D`, util/io/_arr_3d.py) and `.npy`
files through memory maps (`NpyFileReader3D` / `NpyFileWriter3D`), which is what an out-of-core run uses here."""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from pathlib import Path
from typing import Any, Callable, List, Optional

import numpy as np
import torch

from .compensate import SequenceCorrector


@dataclass
class RegistrationConfig:
    """compensate_recording_3D.py:20-29.  Accepted for signature parity: n_jobs / parallelization select the reference's
    CPU worker pool, which does not exist here (one process drives one GPU); batch_size is not read by the reference's
    pipeline either (the reader's buffer_size decides)."""
    n_jobs: int = -1
    batch_size: int = 10
    verbose: bool = False
    parallelization: Optional[str] = None


def _as_5d(array: np.ndarray) -> np.ndarray:
    # util/io/_arr_3d.py:30-38
    if array.ndim == 4:
        return array[np.newaxis, ...]
    if array.ndim == 3:
        return array[np.newaxis, ..., np.newaxis]
    if array.ndim == 5:
        return array
    raise ValueError(f"Array must be 3D, 4D or 5D, got shape {array.shape}")


class ArrayReader3D:
    """In-memory recording as a batch reader (util/io/_arr_3d.py:11-75, batch logic util/io/_base_3d.py:118-262)."""

    def __init__(self, array: np.ndarray, buffer_size: int = 10, bin_size: int = 1):
        self.array = _as_5d(np.asarray(array))
        self.buffer_size = int(buffer_size)
        self.bin_size = int(bin_size)          # "not used for arrays" (_arr_3d.py:25)
        self.frame_count, self.depth, self.height, self.width, self.n_channels = self.array.shape
        self.dtype = self.array.dtype
        self.current_frame = 0

    def _read(self, idx):
        return np.array(self.array[idx])       # a copy, as the reference's reader returns

    def __getitem__(self, key):
        if isinstance(key, (int, np.integer)):
            k = int(key)
            if k < -self.frame_count or k >= self.frame_count:
                raise IndexError(f"Index {k} out of range for {self.frame_count} binned frames")
            return self._read(k % self.frame_count)
        if isinstance(key, slice):
            return self._read(key)
        if isinstance(key, (list, tuple, np.ndarray)):
            idx = [int(i) for i in key]
            for i in idx:
                if i < -self.frame_count or i >= self.frame_count:
                    raise IndexError(f"Index {i} out of range for {self.frame_count} binned frames")
            return self._read([i % self.frame_count for i in idx])
        raise TypeError(f"Invalid index type: {type(key)}")

    def has_batch(self) -> bool:
        return self.current_frame < self.frame_count

    def read_batch(self) -> Optional[np.ndarray]:
        if not self.has_batch():
            return None
        end = min(self.current_frame + self.buffer_size, self.frame_count)
        out = self._read(slice(self.current_frame, end))
        self.current_frame = end
        return out

    def reset(self):
        self.current_frame = 0

    def __len__(self) -> int:
        return self.frame_count

    @property
    def shape(self):
        return (self.frame_count, self.depth, self.height, self.width, self.n_channels)

    def close(self):
        pass


class NpyFileReader3D(ArrayReader3D):
    """A `.npy` recording (T,Z,Y,X[,C]) read through a memory map: only the frames of a batch are paged in."""

    def __init__(self, path, buffer_size: int = 10, bin_size: int = 1):
        self.path = Path(path)
        super().__init__(np.load(str(self.path), mmap_mode="r"), buffer_size, bin_size)

    def close(self):
        self.array = None


class ArrayWriter3D:
    """Accumulates the written volumes in memory (util/io/_arr_3d.py:78-127)."""

    def __init__(self):
        self._vid: List[np.ndarray] = []

    def write_frames(self, frames: np.ndarray):
        frames = np.asarray(frames)
        if frames.ndim == 4:
            frames = frames[np.newaxis, ...]
        if frames.ndim != 5:
            raise ValueError(f"Expected 4D or 5D array, got {frames.ndim}D")
        self._vid.append(np.array(frames))     # the caller's buffer is reused for the next batch

    def get_array(self) -> Optional[np.ndarray]:
        if not self._vid:
            return None
        return np.concatenate(self._vid, axis=0)

    def close(self):
        pass


class NpyFileWriter3D:
    """Writes (T,Z,Y,X,C) into a `.npy` file through a memory map created at the first write; `frame_count` is the
    number of volumes the recording will hold (a `.npy` header fixes the shape up front)."""

    def __init__(self, path, frame_count: int):
        self.path = Path(path)
        self.frame_count = int(frame_count)
        self._mm = None
        self._at = 0

    def write_frames(self, frames: np.ndarray):
        frames = np.asarray(frames)
        if frames.ndim == 4:
            frames = frames[np.newaxis, ...]
        if self._mm is None:
            self.path.parent.mkdir(parents=True, exist_ok=True)
            self._mm = np.lib.format.open_memmap(str(self.path), mode="w+", dtype=frames.dtype,
                                                 shape=(self.frame_count,) + frames.shape[1:])
        n = frames.shape[0]
        if self._at + n > self.frame_count:
            raise ValueError(f"writer was created for {self.frame_count} volumes, got {self._at + n}")
        self._mm[self._at:self._at + n] = frames
        self._at += n

    def close(self):
        if self._mm is not None:
            self._mm.flush()
            self._mm = None


def _fmt(x) -> str:
    return str(getattr(x, "value", x)).upper()


class BatchMotionCorrector:
    """The reference's pipeline object (compensate_recording_3D.py:32-588): same attributes (`mean_disp`, `max_disp`,
    `mean_div`, `mean_translation`, `reference_raw`, `w_init`, `video_reader`, `video_writer`, `w_writer`), same
    `register_progress_callback` / `run(reference_frame) -> reference_raw`.  Keyword-only extras: `video_writer` /
    `w_writer` take writer objects (for formats whose packages this environment lacks), `device`, `cc_prealign`
    (see SequenceCorrector), `lookahead` (batches read ahead of the one being computed)."""

    def __init__(self, options: Any, config: Optional[RegistrationConfig] = None, *, video_writer=None, w_writer=None,
                 device: Optional[torch.device] = None, cc_prealign: bool = False, lookahead: int = 1):
        self.options = options
        self.config = config or RegistrationConfig()
        self.mean_disp: List[float] = []
        self.max_disp: List[float] = []
        self.mean_div: List[float] = []
        self.mean_translation: List[float] = []
        self.reference_raw: Optional[np.ndarray] = None
        self.weight: Optional[np.ndarray] = None
        self.w_init: Optional[np.ndarray] = None
        self.video_reader = None
        self.video_writer = video_writer
        self.w_writer = w_writer
        self.progress_callbacks: List[Callable[[int, int], None]] = []
        self._progress_trackers: dict = {}      # task id -> (frames done, total), compensate_recording_3D.py:60-62
        self._total_frames: Optional[int] = None
        self.device = device
        self.cc_prealign = bool(cc_prealign)
        self.lookahead = int(lookahead)
        # (:64-74) the reference sizes its CPU worker pool here; kept as attributes, the GPU path has one worker
        if self.config.n_jobs == -1:
            import os
            self.n_workers = os.cpu_count() or 4
        else:
            self.n_workers = self.config.n_jobs
        # (:76-121) the executor object of the pipeline: this package's BaseExecutor3D implementation.  run() drives
        # the same kernels through SequenceCorrector (batches stay on the device between the stages); the executor is
        # what the reference's own BatchMotionCorrector would call (tests/test_reference_plugin.py).
        from .executor import B200Executor3D
        self.executor = B200Executor3D(n_workers=1, device=device)

    def register_progress_callback(self, callback: Callable[[int, int], None]) -> None:
        # compensate_recording_3D.py:124-135: callback(current_frame, total_frames), a callback registers once
        if callback not in self.progress_callbacks:
            self.progress_callbacks.append(callback)

    def _notify_progress(self, frames_completed: int, task_id: str = "main") -> None:
        # compensate_recording_3D.py:137-162: per-task cumulative counters; only the main task (whose total is the
        # reader's length) reaches the callbacks, a failing callback warns
        if task_id not in self._progress_trackers:
            self._progress_trackers[task_id] = (0, self._total_frames if task_id == "main" else None)
        current, total = self._progress_trackers[task_id]
        current += frames_completed
        self._progress_trackers[task_id] = (current, total)
        if task_id == "main" and total and self.progress_callbacks:
            for cb in self.progress_callbacks:
                try:
                    cb(current, total)
                except Exception as e:
                    warnings.warn(f"Progress callback error: {e}")

    # -- I/O (compensate_recording_3D.py:164-196, OF_options_3D.py:405-463) ---------------------------------------
    def _setup_io(self):
        o = self.options
        out_dir = Path(getattr(o, "output_path", "results"))
        out_dir.mkdir(parents=True, exist_ok=True)
        if hasattr(o, "get_video_reader"):
            self.video_reader = o.get_video_reader()
        else:       # an options object without the getters: the same factory
            from . import io_factory
            self.video_reader = io_factory.get_video_file_reader(getattr(o, "input_file", None),
                                                                 buffer_size=int(o.buffer_size),
                                                                 bin_size=int(getattr(o, "bin_size", 1)))
        fmt = _fmt(getattr(o, "output_format", "ARRAY"))
        if self.video_writer is None:
            if hasattr(o, "get_video_writer"):
                self.video_writer = o.get_video_writer()
            else:
                from . import io_factory
                self.video_writer = io_factory.get_video_file_writer(str(out_dir / f"compensated.{fmt}"), fmt,
                                                                     frame_count=len(self.video_reader))
        if getattr(o, "save_w", False) and self.w_writer is None:
            from . import io_factory
            try:
                if fmt == "ARRAY":       # (:175-182) flows stay in memory when the frames do
                    self.w_writer = io_factory.get_video_file_writer(None, "ARRAY")
                else:                    # (:183-190: w.h5 with datasets u, v, w in the reference; w.npy (T,Z,Y,X,3) here)
                    n = len(self.video_reader) if hasattr(self.video_reader, "__len__") else 0
                    self.w_writer = io_factory.get_video_file_writer(str(out_dir / "w.npy"), "NPY", frame_count=n,
                                                                     dataset_names=["u", "v", "w"])
            except Exception as e:       # (:191-196)
                warnings.warn(f"Failed to create displacement writer: {e}. Displacements will not be saved.")
                self.w_writer = None
                o.save_w = False

    def _setup_reference(self, reference_frame):
        o = self.options
        if reference_frame is None:
            rf = getattr(o, "reference_frames", None)
            if isinstance(rf, (list, tuple)) and not isinstance(rf, np.ndarray):
                # OF_options_3D.py:496-503: frames = reader[indices]; 5-D -> mean over time
                frames = np.asarray(self.video_reader[[int(i) for i in rf]])
                reference_frame = frames.mean(axis=0) if frames.ndim == 5 else frames
            else:
                reference_frame = o.get_reference_frame(None)
        self.reference_raw = np.asarray(reference_frame).astype(np.float64)
        # (:207-224) the per-channel data weights as a full (Z,Y,X,C) array
        Z, Y, X = self.reference_raw.shape[:3]
        n_channels = self.reference_raw.shape[3] if self.reference_raw.ndim == 4 else 1
        self.weight = np.ones((Z, Y, X, n_channels), dtype=np.float64)
        for c in range(n_channels):
            self.weight[..., c] = o.get_weight_at(c, n_channels)

    # -- the pipeline (compensate_recording_3D.py:431-557) -----------------------------------------------------------
    def run(self, reference_frame: Optional[np.ndarray] = None) -> np.ndarray:
        self._setup_io()
        self._setup_reference(reference_frame)
        reader, writer, o = self.video_reader, self.video_writer, self.options
        self._total_frames = len(reader) if hasattr(reader, "__len__") else None     # :438-439
        seq = SequenceCorrector(self.reference_raw, o, device=self.device, statistics=True,
                                cc_prealign=self.cc_prealign)
        dtypes: List[np.dtype] = []

        def batches():
            while reader.has_batch():
                b = reader.read_batch()
                if b is None:
                    return
                b = np.asarray(b)
                dtypes.append(b.dtype)
                yield b

        def sink(k, registered, flows):
            # the executors return `registered` in the batch's dtype (sequential_3d.py:71-73: np.empty_like(batch)) and
            # float32 flows; the host tensors handed in here are reused for later batches, the writers copy
            reg = registered.numpy() if isinstance(registered, torch.Tensor) else np.asarray(registered)
            fl = flows.numpy() if isinstance(flows, torch.Tensor) else np.asarray(flows)
            writer.write_frames(reg.astype(dtypes[k], copy=True))
            if getattr(o, "save_w", False) and self.w_writer is not None:
                self.w_writer.write_frames(np.array(fl))
            self._notify_progress(reg.shape[0])

        try:
            seq.run_stream(batches(), sink, lookahead=self.lookahead)
            st = seq.statistics()
            self.mean_disp, self.max_disp = st["mean_disp"], st["max_disp"]
            self.mean_div, self.mean_translation = st["mean_div"], st["mean_translation"]
            if seq.w_init is not None:
                from . import device as dev
                seq.reg.sync()
                self.w_init = dev.to_host(seq.w_init).astype(np.float64)
        finally:
            seq.close()
        self._save_metadata()
        self._cleanup()
        return self.reference_raw

    def _save_metadata(self):
        # compensate_recording_3D.py:559-581
        if not getattr(self.options, "save_meta_info", False):
            return
        out_dir = Path(self.options.output_path)
        np.savez(str(out_dir / "statistics.npz"), mean_disp=np.array(self.mean_disp), max_disp=np.array(self.max_disp),
                 mean_div=np.array(self.mean_div), mean_translation=np.array(self.mean_translation))
        if self.reference_raw is not None:
            np.save(str(out_dir / "reference_frame.npy"), self.reference_raw)

    def _cleanup(self):
        if self.video_writer is not None:
            self.video_writer.close()
        if self.w_writer is not None:
            self.w_writer.close()


def compensate_recording(options: Any, reference_frame: Optional[np.ndarray] = None,
                         config: Optional[RegistrationConfig] = None, **kwargs) -> np.ndarray:
    """Drop-in for flowreg3d.motion_correction.compensate_recording (compensate_recording_3D.py:591-613): runs the
    pipeline the options describe and returns the reference volume used.  The registered volumes go to the writer
    (`options.output_format = "ARRAY"`: read them back with `options_writer.get_array()`, i.e. keep the
    BatchMotionCorrector, or pass `video_writer=`)."""
    return BatchMotionCorrector(options, config, **kwargs).run(reference_frame)
