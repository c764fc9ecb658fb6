"""CPU ORACLE for the rigid cross-correlation pre-alignment.

TEST INFRASTRUCTURE -- not product code (same import rule as ``oracle.py``: tests, smoke(), the
CPU legs of bench.py only).  No product code for this row exists yet: this file is the "oracle
first" step for it.

What it restates (``file:line`` relative to ``/root/reference/src/flowreg3d``):
    util/xcorr_prealignment.py:8-13     _proj_xy, _proj_xz (mean projections)
    util/xcorr_prealignment.py:15-99    estimate_rigid_xcorr_3d
    util/resize_util_3D.py:159-166      imresize2d_gauss_cubic (= the 3-D fused resize of a (1,H,W)
                                        volume with per_axis=True: one sigma per axis, :120-123)
    motion_correction/parallelization/sequential_3d.py:89-145   the six executor steps around it

THIRD-PARTY ARITHMETIC THAT IS ABSENT HERE: the reference calls
``skimage.registration.phase_cross_correlation(ref, mov, upsample_factor=up, normalization="phase",
disambiguate=True)`` (xcorr_prealignment.py:60-66, 91-97).  scikit-image (pyproject.toml:31 pins
``scikit-image>=0.24.0``) is NOT installed in the build container or on the GPU box and cannot be
installed (no network), so ``phase_cross_correlation`` below is a restatement of its published
algorithm -- Guizar-Sicairos, Thurman, Fienup, "Efficient subpixel image registration algorithms",
Opt. Lett. 33, 156 (2008): FFT cross-power spectrum, optional phase normalisation
(divide by max(|.|, 100 eps)), integer peak, refinement by a matrix-multiply DFT of a
ceil(1.5 up)^2 neighbourhood up-sampled `up` times -- plus scikit-image 0.24's documented
``disambiguate`` step (the peak of a circular correlation is ambiguous by the image size per axis:
the 2^ndim candidate shifts are ranked by the Pearson correlation of the overlapping tiles of the
reference and the circularly shifted moving image; shift applied with ``scipy.ndimage.shift``,
``mode="grid-wrap"``, spline order 3 for sub-pixel shifts, 0 otherwise).

PARITY STATUS: **parity unpinned against scikit-image itself**; pinned only by the reference's
own known-answer tests for this path (``/root/reference/tests/util/test_xcorr_prealignment.py``:
pure translation, multichannel + weights, down-sampling 256/128/64, sign convention, z scaling,
pipeline composition), which ``tests/test_oracle_xcorr.py`` mirrors with their tolerances.
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy import fft as sfft
from scipy import ndimage as ndi

from . import oracle as O


# --------------------------------------------------------------------------------------------
# imresize2d_gauss_cubic  (util/resize_util_3D.py:159-166 -> :114-156 with per_axis=True)
# --------------------------------------------------------------------------------------------
def imresize2d_gauss_cubic(img2d, out_hw, sigma_coeff=0.6):
    x = np.ascontiguousarray(np.asarray(img2d)[None, ...].astype(np.float32, copy=False))
    oh, ow = int(out_hw[0]), int(out_hw[1])
    sy, sx = oh / x.shape[1], ow / x.shape[2]
    sigx = sigma_coeff / sx if sx < 1.0 else 0.0       # :120-123, one sigma per axis
    sigy = sigma_coeff / sy if sy < 1.0 else 0.0
    tabs = (O.resize_tables(x.shape[2], ow, sigx), O.resize_tables(x.shape[1], oh, sigy),
            O.resize_tables(1, 1, 0.0))
    y = O._resize3(x, (1, oh, ow), tabs)
    return y[0].astype(np.asarray(img2d).dtype, copy=False)


# --------------------------------------------------------------------------------------------
# phase_cross_correlation (scikit-image >= 0.24, unmasked path) -- restated, see the header
# --------------------------------------------------------------------------------------------
def _upsampled_dft(data, upsampled_region_size, upsample_factor, axis_offsets):
    """Matrix-multiply DFT of a small up-sampled neighbourhood (Guizar-Sicairos 2008, sec. 3):
    out[a, b] = sum_{j,k} exp(-2 pi i ((a - off_0) f_j + (b - off_1) f_k)) data[j, k],
    f = fftfreq(n, upsample_factor); kernels in the precision of `data`."""
    im2pi = 1j * 2 * np.pi
    sizes = [upsampled_region_size] * data.ndim if np.isscalar(upsampled_region_size) \
        else list(upsampled_region_size)
    props = list(zip(data.shape, sizes, axis_offsets))
    for n_items, ups_size, ax_offset in props[::-1]:
        kernel = (np.arange(ups_size) - ax_offset)[:, None] * sfft.fftfreq(n_items, upsample_factor)
        kernel = np.exp(-im2pi * kernel).astype(data.dtype, copy=False)
        data = np.tensordot(kernel, data, axes=(1, -1))
    return data


def _disambiguate_shift(reference_image, moving_image, shift):
    shape = reference_image.shape
    positive_shift = [s_i % n for s_i, n in zip(shift, shape)]
    negative_shift = [s_i - n for s_i, n in zip(positive_shift, shape)]
    subpixel = np.any(np.array(shift) % 1 != 0)
    interp_order = 3 if subpixel else 0
    shifted = ndi.shift(moving_image, shift, mode="grid-wrap", order=interp_order)
    indices = np.round(positive_shift).astype(int)
    splits_per_dim = [(slice(0, i), slice(i, None)) for i in indices]
    max_corr = -1.0
    max_slice = None
    for test_slice in itertools.product(*splits_per_dim):
        reference_tile = np.reshape(reference_image[test_slice], -1)
        moving_tile = np.reshape(shifted[test_slice], -1)
        corr = -1.0
        if reference_tile.size > 2:
            with np.errstate(invalid="ignore", divide="ignore"):
                corr = np.corrcoef(reference_tile, moving_tile)[0, 1]
        if corr > max_corr:
            max_corr = corr
            max_slice = test_slice
    if max_slice is None:
        return np.asarray(shift)
    real_shift = []
    for sl, pos, neg in zip(max_slice, positive_shift, negative_shift):
        real_shift.append(pos if sl.stop is None else neg)
    return np.array(real_shift)


def phase_cross_correlation(reference_image, moving_image, upsample_factor=1, normalization="phase",
                            disambiguate=False):
    """Shift (in pixels, one entry per axis) that registers `moving_image` with `reference_image`."""
    reference_image = np.asarray(reference_image)
    moving_image = np.asarray(moving_image)
    if reference_image.shape != moving_image.shape:
        raise ValueError("images must be same shape")
    src_freq = sfft.fftn(reference_image)            # float32 input -> complex64, as scipy.fft does
    target_freq = sfft.fftn(moving_image)
    shape = src_freq.shape
    image_product = src_freq * target_freq.conj()
    if normalization == "phase":
        eps = np.finfo(image_product.real.dtype).eps
        image_product /= np.maximum(np.abs(image_product), 100 * eps)
    elif normalization is not None:
        raise ValueError("normalization must be either phase or None")
    cross_correlation = sfft.ifftn(image_product)
    maxima = np.unravel_index(np.argmax(np.abs(cross_correlation)), cross_correlation.shape)
    midpoint = np.array([np.fix(n / 2) for n in shape])
    float_dtype = image_product.real.dtype
    shift = np.stack(maxima).astype(float_dtype, copy=False)
    shift[shift > midpoint] -= np.array(shape)[shift > midpoint]
    if upsample_factor > 1:
        up = np.array(upsample_factor, dtype=float_dtype)
        shift = np.round(shift * up) / up
        region = np.ceil(up * 1.5)
        dftshift = np.fix(region / 2.0)
        offset = dftshift - shift * up
        cc = _upsampled_dft(image_product.conj(), region, up, offset).conj()
        maxima = np.unravel_index(np.argmax(np.abs(cc)), cc.shape)
        maxima = np.stack(maxima).astype(float_dtype, copy=False)
        maxima -= dftshift
        shift = shift + maxima / up
    for dim in range(src_freq.ndim):
        if shape[dim] == 1:
            shift[dim] = 0
    if disambiguate:
        shift = _disambiguate_shift(reference_image, moving_image, shift)
    return shift


# --------------------------------------------------------------------------------------------
# estimate_rigid_xcorr_3d  (util/xcorr_prealignment.py:15-99)
# --------------------------------------------------------------------------------------------
def _windowed(p):
    """:50-58 / :81-89 -- float32, zero mean, separable Hann window."""
    p = p.astype(np.float32, copy=False)
    p = p - p.mean()
    h0 = np.hanning(p.shape[0]).astype(np.float32)
    h1 = np.hanning(p.shape[1]).astype(np.float32)
    return p * (h0[:, None] * h1[None, :])


def estimate_rigid_xcorr_3d(ref_vol, mov_vol, target_hw=(256, 256), target_z=None, up=10,
                            normalization="phase", disambiguate=True, weight=None):
    """Returns -[dx, dy, dz] (float32): the displacement to ADD to the backward-warp field so that
    `mov_vol` lines up with `ref_vol`."""
    ref_vol = np.asarray(ref_vol)
    mov_vol = np.asarray(mov_vol)
    if ref_vol.ndim == 4 and ref_vol.shape[3] > 1:                     # :25-36
        if weight is not None:
            w = np.asarray(weight).reshape(-1).astype(np.float32)
            w = w / w.sum()
            ref_vol = np.tensordot(ref_vol, w, axes=([3], [0]))
            mov_vol = np.tensordot(mov_vol, w, axes=([3], [0]))
        else:
            ref_vol = ref_vol.mean(axis=3)
            mov_vol = mov_vol.mean(axis=3)
    elif ref_vol.ndim == 4:
        ref_vol = ref_vol[..., 0]
        mov_vol = mov_vol[..., 0]
    Z, H, W = ref_vol.shape
    Th = H if target_hw is None else min(H, int(target_hw[0]))
    Tw = W if target_hw is None else min(W, int(target_hw[1]))
    sy, sx = H / Th, W / Tw
    pxy_r, pxy_m = ref_vol.mean(axis=0), mov_vol.mean(axis=0)          # :44-48
    if (Th, Tw) != (H, W):
        pxy_r = imresize2d_gauss_cubic(pxy_r, (Th, Tw))
        pxy_m = imresize2d_gauss_cubic(pxy_m, (Th, Tw))
    s_xy = phase_cross_correlation(_windowed(pxy_r), _windowed(pxy_m), upsample_factor=up,
                                   normalization=normalization, disambiguate=disambiguate)
    dy = float(s_xy[0]) * sy
    dx = float(s_xy[1]) * sx
    Tz = Z if target_z is None else min(Z, int(target_z))              # :71-78
    sz = Z / Tz
    pxz_r, pxz_m = ref_vol.mean(axis=1), mov_vol.mean(axis=1)
    if Tz != Z or Tw != W:
        pxz_r = imresize2d_gauss_cubic(pxz_r, (Tz, Tw))
        pxz_m = imresize2d_gauss_cubic(pxz_m, (Tz, Tw))
    s_xz = phase_cross_correlation(_windowed(pxz_r), _windowed(pxz_m), upsample_factor=up,
                                   normalization=normalization, disambiguate=disambiguate)
    dz = float(s_xz[0]) * sz
    return -np.array([dx, dy, dz], dtype=np.float32)


# --------------------------------------------------------------------------------------------
# the executor steps around it  (parallelization/sequential_3d.py:89-145)
# --------------------------------------------------------------------------------------------
def flow_with_cc_initialization(reference_proc, moving_proc, w_init, flow_params, cc_hw=256, cc_up=1):
    """One frame of `process_batch` with `cc_initialization=True`:
    (1) warp the moving frame by w_init (linear), (2) rigid residual by phase correlation,
    (3) w_init + rigid, (4) warp the ORIGINAL moving frame by the combined field (linear),
    (5) residual non-rigid flow from zero, (6) total = combined + residual, float32."""
    params = {k: v for k, v in flow_params.items() if k not in ("cc_initialization", "cc_hw", "cc_up")}
    hw = (cc_hw, cc_hw) if isinstance(cc_hw, int) else cc_hw
    w_init = np.asarray(w_init)
    part = O.imregister_wrapper(moving_proc, w_init[..., 0], w_init[..., 1], w_init[..., 2],
                                reference_proc, "linear")
    ref_cc = reference_proc[..., 0] if reference_proc.ndim == 4 and reference_proc.shape[3] == 1 \
        else reference_proc
    mov_cc = part[..., 0] if part.ndim == 4 and part.shape[3] == 1 else part
    w_cross = estimate_rigid_xcorr_3d(ref_cc, mov_cc, target_hw=hw, up=cc_up, weight=params.get("weight"))
    comb = w_init.copy()
    for q in range(3):
        comb[..., q] += w_cross[q]
    aligned = O.imregister_wrapper(moving_proc, comb[..., 0], comb[..., 1], comb[..., 2], reference_proc,
                                   "linear")
    if aligned.ndim == 3:
        aligned = aligned[..., None]
    resid = O.get_displacement(reference_proc, aligned, uvw=np.zeros_like(w_init), **params)
    return (comb + resid).astype(np.float32, copy=False), w_cross
