"""OFOptions -- the hot-path subset of the reference's configuration object (drop-in boundary B3).

Mirrors flowreg3d.motion_correction.OF_options_3D.OFOptions (OF_options_3D.py:130-686): same field
names, defaults, validators (alpha -> 3-tuple, weight normalisation, sigma -> (n,4)), quality
presets, ``effective_min_level``, ``get_weight_at``, ``to_dict``, ``copy``, ``save_options`` / ``load_options``,
``get_video_reader`` / ``get_video_writer`` (through ``io_factory``: arrays and .npy files; objects with the
reference's reader / writer protocol pass through) and ``get_mcp_schema``.  Unknown fields are rejected exactly
like the reference (extra="forbid").  When the
real flowreg3d package is importable its own OFOptions can be passed to every entry point here
instead -- only the attributes below are read.
"""
from __future__ import annotations

import json
from datetime import date
from enum import Enum
from pathlib import Path
from typing import Any, List, Optional, Tuple, Union

import numpy as np
from pydantic import BaseModel, ConfigDict, Field, PrivateAttr, StrictInt, field_validator, model_validator


class QualitySetting(str, Enum):
    QUALITY = "quality"
    BALANCED = "balanced"
    FAST = "fast"
    CUSTOM = "custom"


class ChannelNormalization(str, Enum):
    JOINT = "joint"
    SEPARATE = "separate"


class InterpolationMethod(str, Enum):
    NEAREST = "nearest"
    LINEAR = "linear"
    CUBIC = "cubic"


class OutputFormat(str, Enum):
    # OF_options_3D.py:86-103 (file formats need the reference's writers; ARRAY and NPY are served here)
    TIFF = "TIFF"
    HDF5 = "HDF5"
    MAT = "MAT"
    MULTIFILE_TIFF = "MULTIFILE_TIFF"
    MULTIFILE_MAT = "MULTIFILE_MAT"
    MULTIFILE_HDF5 = "MULTIFILE_HDF5"
    CAIMAN_HDF5 = "CAIMAN_HDF5"
    ARRAY = "ARRAY"
    NPY = "NPY"


class NamingConvention(str, Enum):
    DEFAULT = "default"
    BATCH = "batch"


class ConstancyAssumption(str, Enum):
    GRAY = "gray"
    GRADIENT = "gc"


class OFOptions(BaseModel):
    model_config = ConfigDict(arbitrary_types_allowed=True, validate_assignment=False, extra="forbid",
                              populate_by_name=True)

    # I/O and bookkeeping fields of the reference (OF_options_3D.py:141-153, 178, 190-199, 222): accepted so that every
    # keyword the reference's OFOptions takes constructs here too; file readers / writers are out of scope,
    # the array path ignores them exactly as the reference's ArrayReader3D / ArrayWriter do (bin_size: "not used for
    # arrays", util/io/_arr_3d.py:25; n_references > 1: the reference repeats ONE reference, :469-476)
    input_file: Any = Field(None)
    input_dim_order: str = Field("TZYX")
    output_path: Any = Field(Path("results"))
    output_format: OutputFormat = Field(OutputFormat.MAT)
    output_file_name: Optional[str] = Field(None)
    channel_idx: Optional[List[int]] = Field(None)
    bin_size: StrictInt = Field(1, ge=1)
    n_references: StrictInt = Field(1, ge=1)
    min_frames_per_reference: StrictInt = Field(20, ge=1)
    save_meta_info: bool = Field(True)
    save_w: bool = Field(False)
    save_valid_mask: bool = Field(False)
    save_valid_idx: bool = Field(False)
    naming_convention: NamingConvention = Field(NamingConvention.DEFAULT)
    preproc_funct: Optional[Any] = Field(None, exclude=True)
    # flow parameters (OF_options_3D.py:155-174)
    alpha: Union[float, Tuple[float, float], Tuple[float, float, float]] = Field((0.25, 0.25, 0.25))
    weight: Union[List[float], np.ndarray] = Field([0.5, 0.5])
    levels: StrictInt = Field(100, ge=1)
    min_level: StrictInt = Field(5, ge=-1)
    quality_setting: QualitySetting = Field(QualitySetting.QUALITY)
    eta: float = Field(0.8, gt=0, le=1)
    update_lag: StrictInt = Field(5, ge=1)
    iterations: StrictInt = Field(100, ge=1)
    a_smooth: float = Field(1.0, ge=0)
    a_data: float = Field(0.45, gt=0, le=1)
    # preprocessing (:177-182)
    sigma: Any = Field([[1.0, 1.0, 1.0, 0.1], [1.0, 1.0, 1.0, 0.1]])
    buffer_size: StrictInt = Field(10, ge=1)
    # reference (:185-192)
    reference_frames: Union[List[int], str, Any, np.ndarray] = Field(list(range(50, 500)))
    update_reference: bool = Field(False)
    # processing options (:197-224)
    verbose: bool = Field(False)
    output_typename: Optional[str] = Field("double")
    channel_normalization: ChannelNormalization = Field(ChannelNormalization.JOINT)
    interpolation_method: InterpolationMethod = Field(InterpolationMethod.CUBIC)
    cc_initialization: bool = Field(False)
    cc_hw: Union[int, Tuple[int, int]] = Field(256)
    cc_up: int = Field(10, ge=1)
    update_initialization_w: bool = Field(True)
    constancy_assumption: ConstancyAssumption = Field(ConstancyAssumption.GRADIENT, alias="constancy")

    _quality_setting_old: QualitySetting = PrivateAttr(default=QualitySetting.QUALITY)
    _video_reader: Any = PrivateAttr(default=None)
    _video_writer: Any = PrivateAttr(default=None)

    @field_validator("alpha", mode="before")
    @classmethod
    def _alpha(cls, v):
        # OF_options_3D.py:236-263: scalar -> (a,a,a); (a,b) -> (a,a,b); all positive
        if isinstance(v, (int, float)):
            vals, one = (v, v, v), True
        elif isinstance(v, (list, tuple)) and len(v) in (1, 2, 3):
            vals = (v[0],) * 3 if len(v) == 1 else ((v[0], v[0], v[1]) if len(v) == 2 else tuple(v))
            one = len(v) == 1
        else:
            raise ValueError("Alpha must be scalar, 2-element, or 3-element tuple")
        if any(a <= 0 for a in vals):
            # the reference's two messages (:244-258)
            raise ValueError("Alpha must be positive" if one else "All alpha values must be positive")
        return tuple(float(a) for a in vals)

    @field_validator("weight", mode="before")
    @classmethod
    def _weight(cls, v):
        # :265-283: 1-D weights are normalised to sum 1
        if isinstance(v, np.ndarray):
            if v.ndim == 1 and v.sum() > 0:
                return (v / v.sum()).tolist()
            return v.tolist()
        if isinstance(v, (list, tuple)):
            arr = np.asarray(v, dtype=float)
            if arr.ndim == 1 and arr.sum() > 0:
                return (arr / arr.sum()).tolist()
        return v

    @field_validator("sigma", mode="before")
    @classmethod
    def _sigma(cls, v):
        # :285-310: rows [sx, sy, sz, st]; 3-vectors get sz = 1 inserted
        sig = np.asarray(v, dtype=float)
        if sig.ndim == 1:
            if sig.size == 3:
                sig = np.insert(sig, 2, 1.0)
            elif sig.size != 4:
                raise ValueError("1D sigma must be [sx, sy, sz, st] or [sx, sy, st] for 3D")
            return sig.reshape(1, 4).tolist()
        if sig.ndim == 2:
            if sig.shape[1] == 3:
                sig = np.insert(sig, 2, 1.0, axis=1)
            elif sig.shape[1] != 4:
                raise ValueError("2D sigma must be (n_channels, 4) for 3D")
            return sig.tolist()
        raise ValueError("Sigma must be [sx,sy,sz,st] or (n_channels, 4) for 3D")

    @model_validator(mode="after")
    def _quality(self):
        # :312-327
        if not isinstance(self.output_path, Path):
            self.output_path = Path(self.output_path)
        if self.quality_setting != QualitySetting.CUSTOM:
            self._quality_setting_old = self.quality_setting
        if self.min_level >= 0:
            self.quality_setting = QualitySetting.CUSTOM
        elif self.min_level == -1 and self.quality_setting == QualitySetting.CUSTOM:
            self.quality_setting = self._quality_setting_old
        return self

    @property
    def effective_min_level(self) -> int:
        # :329-341
        if self.min_level >= 0:
            return self.min_level
        return {QualitySetting.QUALITY: 0, QualitySetting.BALANCED: 4, QualitySetting.FAST: 6,
                QualitySetting.CUSTOM: max(self.min_level, 0)}.get(self.quality_setting, 0)

    @property
    def constancy(self) -> str:
        return self.constancy_assumption.value

    def get_sigma_at(self, i: int) -> np.ndarray:
        sig = np.asarray(self.sigma, dtype=float)
        if sig.ndim == 1:
            return sig
        return sig[0] if i >= sig.shape[0] else sig[i]

    def get_weight_at(self, i: int, n_channels: int):
        # :371-399
        w = np.asarray(self.weight, dtype=float)
        if w.ndim <= 1:
            if w.size == 1:
                return float(w.reshape(-1)[0])
            if w.size > n_channels:
                w = w[:n_channels]
                w = w / w.sum()
                self.weight = w.tolist()
            if i >= w.size:
                return 1.0 / n_channels
            return float(w[i])
        if i >= w.shape[0]:
            return np.ones(w.shape[1:]) / n_channels
        return w[i]

    def get_reference_frame(self, video=None):
        """The fixed volume the options describe, for array-backed recordings (OF_options_3D.py:466-503): an ndarray is
        returned as it is; a list of frame indices selects frames of `video` (T,Z,Y,X[,C]) and returns their mean over
        time -- `video[indices].mean(axis=0)`, numpy's arithmetic, as the reference's 3-D branch (:496-503); indices
        outside the recording raise IndexError like the reference's reader (util/io/_base_3d.py:141-144).  File paths
        belong to the reference's readers (out of scope here).  n_references > 1 repeats the single reference, as the
        reference does (:469-476)."""
        if self.n_references > 1:
            import warnings
            warnings.warn("Multi-reference mode not fully implemented; repeating a single computed reference")
            one = self.model_copy(update={"n_references": 1})
            return [one.get_reference_frame(video)] * self.n_references
        rf = self.reference_frames
        if isinstance(rf, np.ndarray):
            return rf
        if isinstance(rf, (list, tuple)) and video is not None:
            idx = [int(i) for i in rf]
            if isinstance(video, np.ndarray):
                T = video.shape[0]
                for i in idx:
                    if i < -T or i >= T:
                        raise IndexError(f"Index {i} out of range for {T} binned frames")
            frames = np.asarray(video[idx])          # an array or any reader: `video_reader[indices]` (:496)
            if frames.ndim == 5:
                return frames.mean(axis=0)           # (T,Z,Y,X,C) -> mean over time (:499-503)
            if frames.ndim == 4 and isinstance(video, np.ndarray) and video.ndim == 4:
                return frames.mean(axis=0)           # a channel-less array recording (T,Z,Y,X)
            return frames                            # a single volume (:504-511)
        if isinstance(rf, (str, bytes)) or hasattr(rf, "__fspath__"):
            p = Path(rf if not isinstance(rf, bytes) else rf.decode())
            if p.suffix.lower() == ".npy":
                return np.load(str(p))
            if p.suffix.lower() in (".tif", ".tiff"):
                try:
                    import tifffile
                except ImportError as e:     # (:484-493: the reference hard-requires tifffile for this branch)
                    raise RuntimeError(f"Unable to read reference image: {p} (tifffile is not installed)") from e
                return tifffile.imread(str(p))
            raise RuntimeError(f"Unable to read reference image: {p}")
        return np.asarray(rf)

    # -- readers / writers (:405-463) ---------------------------------------------------------------------------
    def get_video_reader(self):
        """Cached reader of `input_file`; the reader is stored back in `input_file` as the reference does (:426-427)."""
        if self._video_reader is not None:
            return self._video_reader
        from . import io_factory
        if io_factory._is_reader(self.input_file):
            self._video_reader = self.input_file
            return self._video_reader
        self._video_reader = io_factory.get_video_file_reader(self.input_file, buffer_size=self.buffer_size,
                                                              bin_size=self.bin_size, dim_order=self.input_dim_order)
        self.input_file = self._video_reader
        return self._video_reader

    def get_video_writer(self):
        """Cached writer for `output_format`; file name as the reference builds it (:437-457)."""
        if self._video_writer is not None:
            return self._video_writer
        from . import io_factory
        fmt = self.output_format
        if self.output_file_name:
            filename = self.output_file_name
        else:
            ext = "HDF5" if fmt == OutputFormat.HDF5 else getattr(fmt, "value", str(fmt))
            if self.naming_convention == NamingConvention.DEFAULT:
                filename = str(Path(self.output_path) / f"compensated.{ext}")
            else:
                reader = self.get_video_reader()
                stem = Path(getattr(reader, "input_file_name", "output")).stem
                filename = str(Path(self.output_path) / f"{stem}_compensated.{ext}")
        kw = {}
        if getattr(fmt, "value", fmt) == "NPY":
            filename = filename[:-4] + ".npy" if filename.endswith(".NPY") else filename
            kw["frame_count"] = len(self.get_video_reader())
        self._video_writer = io_factory.get_video_file_writer(filename, getattr(fmt, "value", fmt), **kw)
        return self._video_writer

    # -- persistence (:603-668) ---------------------------------------------------------------------------------
    def save_options(self, filepath=None) -> None:
        """JSON with the one-line header the MATLAB toolbox writes ("Compensation options <date>").  An ndarray
        reference is stored next to it (reference_frames.tif with tifffile, reference_frames.npy without)."""
        path = Path(filepath) if filepath else Path(self.output_path) / "options.json"
        path.parent.mkdir(parents=True, exist_ok=True)
        data = self.model_dump(by_alias=True, exclude={"preproc_funct"})
        for k, v in list(data.items()):
            if isinstance(v, Path):
                data[k] = str(v)
            elif isinstance(v, np.ndarray):
                data[k] = v.tolist()
            elif isinstance(v, Enum):
                data[k] = v.value
            elif isinstance(v, tuple):
                data[k] = list(v)
            elif k == "input_file" and v is not None and not isinstance(v, (str, int, float, list, dict)):
                data[k] = None                       # a reader object cannot be serialised
        if isinstance(self.reference_frames, np.ndarray):
            try:
                import tifffile
                ref_path = path.parent / "reference_frames.tif"
                tifffile.imwrite(str(ref_path), self.reference_frames)
            except (ImportError, AttributeError):
                ref_path = path.parent / "reference_frames.npy"
                np.save(str(ref_path), self.reference_frames)
            data["reference_frames"] = str(ref_path)
        with path.open("w", encoding="utf-8") as f:
            f.write(f"Compensation options {date.today().isoformat()}\n\n")
            json.dump(data, f, indent=2)
        if self.verbose:
            print(f"Options saved to {path}")

    @classmethod
    def load_options(cls, filepath) -> "OFOptions":
        """Reads what save_options (or the MATLAB toolbox) wrote: header lines before the first "{" are skipped."""
        with Path(filepath).open("r", encoding="utf-8") as f:
            lines = f.readlines()
        start = next((i for i, line in enumerate(lines) if line.strip().startswith("{")), 0)
        data = json.loads("".join(lines[start:]))
        ref = data.get("reference_frames")
        if isinstance(ref, str):
            rp = Path(ref)
            if rp.exists() and rp.suffix.lower() == ".npy":
                data["reference_frames"] = np.load(str(rp))
            elif rp.exists() and rp.suffix.lower() in (".tif", ".tiff"):
                import tifffile
                data["reference_frames"] = tifffile.imread(str(rp))
        return cls(**data)

    def copy(self) -> "OFOptions":
        return self.model_copy(deep=True)

    def to_dict(self) -> dict:
        # :667-680
        return {"alpha": self.alpha, "weight": self.weight, "levels": self.levels,
                "min_level": self.effective_min_level, "eta": self.eta, "iterations": self.iterations,
                "update_lag": self.update_lag, "a_data": self.a_data, "a_smooth": self.a_smooth,
                "const_assumption": self.constancy_assumption.value}

    def __repr__(self) -> str:
        return (f"OFOptions(quality={self.quality_setting.value}, alpha={self.alpha}, "
                f"levels={self.levels}, min_level={self.effective_min_level})")


def get_mcp_schema() -> dict:
    """JSON schema of the options model (OF_options_3D.py:741-745)."""
    return OFOptions.model_json_schema(mode="serialization")
