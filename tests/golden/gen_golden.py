#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference (flowreg3D at /root/reference).

Runs only in the build container (the reference cannot travel to the GPU box):

    NUMBA_CACHE_DIR=/tmp/numba_cache python tests/golden/gen_golden.py

Every file stores the exact inputs fed to the reference function and the outputs it returned,
so the oracle (oracle/oracle.py) and the CUDA path can be checked against them anywhere.
The reference is imported read-only from /root/reference/src; the three optional I/O packages it
hard-imports but that are absent here (tifffile, h5py, hdf5storage) are stubbed with empty
modules -- they are not on the array path.
"""
import os
import sys
import types
from pathlib import Path

import numpy as np
from scipy.ndimage import gaussian_filter

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference/src")
for _m in ("tifffile", "h5py", "hdf5storage"):
    if _m not in sys.modules:
        try:
            __import__(_m)
        except Exception:
            sys.modules[_m] = types.ModuleType(_m)

from flowreg3d.core import optical_flow_3d as R  # noqa: E402
from flowreg3d.core.level_solver_3d import compute_flow_3d as ref_solver  # noqa: E402
from flowreg3d.motion_generation import (  # noqa: E402
    get_high_disp_3d_generator,
    get_low_disp_3d_generator,
    get_test_3d_generator,
)
from flowreg3d.util import image_processing_3D as im3d  # noqa: E402
from flowreg3d.util.resize_util_3D import (  # noqa: E402
    _precompute_fused_gauss_cubic,
    imresize_fused_gauss_cubic3D,
)

OUT = Path(__file__).resolve().parent


def synth_volume(shape, seed):
    """SURVEY 8(d) recipe: min-max of a sigma-weighted sum of Gaussian-smoothed uniform noise."""
    rng = np.random.default_rng(seed)
    n = rng.random(shape)
    v = sum(s * gaussian_filter(n, s) for s in (1.5, 4.0, 8.0))
    return ((v - v.min()) / (v.max() - v.min())).astype(np.float32)


def save(name, **arrs):
    path = OUT / f"{name}.npz"
    np.savez_compressed(path, **arrs)
    print(f"{name}: {path.stat().st_size / 1e6:.2f} MB")


def gen_tables():
    cases = [(56, 22, 0.6 / (22 / 56)), (48, 48, 0.0), (22, 56, 0.0), (512, 168, 0.6 / (10 / 32)),
             (42, 22, 0.6 / (22 / 42)), (128, 17, 0.6 / (9 / 64)), (9, 21, 0.0)]
    d = {"cases": np.array(cases, np.float64)}
    for k, (il, ol, sg) in enumerate(cases):
        idx, wt = _precompute_fused_gauss_cubic(int(il), int(ol), float(sg))
        d[f"idx{k}"] = idx
        d[f"wt{k}"] = wt
    save("tables", **d)


def gen_resize():
    v = np.stack([synth_volume((24, 48, 56), 10 + c) for c in range(2)], -1)
    d = {"src": v}
    sizes = [(10, 20, 22), (24, 48, 56), (30, 50, 60), (8, 16, 18), (19, 38, 45)]
    d["sizes"] = np.array(sizes)
    for k, s in enumerate(sizes):
        d[f"out{k}"] = imresize_fused_gauss_cubic3D(v.astype(np.float64), s).astype(np.float32)
    save("resize", **d)


def small_pair():
    Z, Y, X, C = 24, 48, 56, 2
    fixed = np.stack([synth_volume((Z, Y, X), 10 + c) for c in range(C)], -1)
    np.random.seed(3)
    g = get_low_disp_3d_generator()(Z, Y, X)[0] * 0.5
    f64 = fixed.astype(np.float64)
    moving = R.imregister_wrapper(f64, -g[..., 0], -g[..., 1], -g[..., 2], f64, "linear")
    return fixed, moving, g


def gen_warp():
    fixed, moving, g = small_pair()
    u = (g[..., 0] * 1.3).astype(np.float64)
    v = g[..., 1].astype(np.float64)
    w = (g[..., 2] * 2.0 - 0.7).astype(np.float64)
    d = dict(f2=moving, f1=fixed, u=u, v=v, w=w)
    for meth in ("cubic", "linear"):
        d[meth] = R.imregister_wrapper(moving.astype(np.float64), u, v, w, fixed.astype(np.float64), meth)
    # large displacement: exercises the out-of-volume -> reference-value rule
    d["cubic_big"] = R.imregister_wrapper(moving.astype(np.float64), u * 8, v * 8, w * 8,
                                          fixed.astype(np.float64), "cubic")
    # uint16 raw frame through the final-warp path
    raw = np.round(moving * 4000).astype(np.uint16)
    ref_raw = np.round(fixed * 4000).astype(np.float64)
    uf, vf, wf = (a.astype(np.float32) for a in (u, v, w))
    d["raw_u16"] = raw
    d["raw_ref"] = ref_raw
    d["raw_cubic"] = R.imregister_wrapper(raw, uf, vf, wf, ref_raw, "cubic")
    save("warp", **d)


def gen_motion_tensor():
    fixed, moving, _ = small_pair()
    f1 = fixed[..., 0].astype(np.float64)[:12, :20, :22]
    f2_32 = moving[..., 0].astype(np.float32)[:12, :20, :22]
    h = (1.25, 1.1, 1.3)
    d = dict(f1=f1.astype(np.float32), f2=f2_32, h=np.array(h))
    J32 = R.get_motion_tensor_gc(f1, f2_32, *h)  # f2 float32: the level-warp dtype (levels below top)
    J64 = R.get_motion_tensor_gc(f1, f2_32.astype(np.float64), *h)  # f2 float64: the top level
    d["J_f2f32"] = np.stack(J32, 0)
    d["J_f2f64"] = np.stack(J64, 0)
    save("motion_tensor", **d)


def gen_motion_tensor_alt():
    """The two constancy variants the reference's driver never calls: get_motion_tensor_gray and get_motion_tensor_cs
    (core/optical_flow_3d.py:155-259) on a small pair, intensities scaled to 0..255 (cs uses eps = 80)."""
    fixed, moving, _ = small_pair()
    f1 = (fixed[..., 0].astype(np.float64)[:11, :18, :20]) * 255.0
    f2 = (moving[..., 1].astype(np.float64)[:11, :18, :20]) * 255.0
    h = (1.25, 1.1, 1.3)
    save("motion_tensor_alt", f1=f1, f2=f2, h=np.array(h), J_gray=np.stack(R.get_motion_tensor_gray(f1, f2, *h), 0),
         J_cs=np.stack(R.get_motion_tensor_cs(f1, f2, *h), 0))


def gen_solver():
    rng = np.random.default_rng(5)
    fixed, moving, _ = small_pair()
    d = {}
    for name, C, a_smooth, iters, lag in (("c2", 2, 1.0, 23, 5), ("c1s", 1, 0.5, 12, 5)):
        size = (10, 18, 20)
        f1 = imresize_fused_gauss_cubic3D(fixed.astype(np.float64), size)[..., :C]
        f2 = imresize_fused_gauss_cubic3D(moving.astype(np.float64), size)[..., :C].astype(np.float32)
        h = (24 / size[0], 48 / size[1], 56 / size[2])
        J = [np.zeros((size[0] + 2, size[1] + 2, size[2] + 2, C)) for _ in range(10)]
        for c in range(C):
            Jc = R.get_motion_tensor_gc(f1[..., c], f2[..., c], *h)
            for q in range(10):
                J[q][..., c] = Jc[q]
        wl = np.pad(np.full(size + (C,), 1.0 / C), ((1, 1), (1, 1), (1, 1), (0, 0)))
        u, v, w = (np.pad(gaussian_filter(rng.standard_normal(size), 2.0) * 3, 1, mode="edge") for _ in range(3))
        alpha = (0.4, 0.5, 0.6)
        a_data = np.full(C, 0.45)
        out = ref_solver(*[np.ascontiguousarray(a) for a in J], np.ascontiguousarray(wl), u, v, w,
                         alpha[0], alpha[1], alpha[2], iters, lag, a_data, a_smooth, h[2], h[1], h[0])
        d.update({f"{name}_J": np.stack(J, 0), f"{name}_weight": wl, f"{name}_u": u, f"{name}_v": v,
                  f"{name}_w": w, f"{name}_alpha": np.array(alpha), f"{name}_h": np.array(h),
                  f"{name}_params": np.array([iters, lag, a_smooth]), f"{name}_a_data": a_data,
                  f"{name}_out": out})
    save("solver", **d)


def gen_flow_small():
    fixed, moving, g = small_pair()
    d = dict(fixed=fixed, moving=moving, g=g.astype(np.float32))
    runs = {
        "ml0": dict(alpha=(0.25,) * 3, update_lag=5, iterations=30, min_level=0, levels=100, eta=0.8,
                    a_smooth=1.0, a_data=0.45),
        "ml2w": dict(alpha=(0.25, 0.3, 0.2), update_lag=5, iterations=20, min_level=2, levels=100, eta=0.8,
                     a_smooth=1.0, a_data=0.45, weight=np.array([0.3, 0.7])),
        "ml1s": dict(alpha=(0.5,) * 3, update_lag=10, iterations=15, min_level=1, levels=100, eta=0.8,
                     a_smooth=0.5, a_data=0.45),
    }
    uvw = (g * 0.8).astype(np.float32)
    for k, kw in runs.items():
        d[f"flow_{k}"] = R.get_displacement(fixed, moving, **kw).astype(np.float32)
    d["uvw"] = uvw
    d["flow_ml2w_uvw"] = R.get_displacement(fixed, moving, uvw=uvw.copy(), **runs["ml2w"]).astype(np.float32)
    save("flow_small", **d)


def gen_preprocess():
    rng = np.random.default_rng(11)
    ref = (np.stack([synth_volume((12, 24, 28), 20 + c) for c in range(2)], -1) * 3000 + 100).astype(np.float32)
    batch = (ref[None] * (1 + 0.05 * rng.standard_normal((3,) + ref.shape))).astype(np.float32)
    sigma = np.array([[1.0, 1.0, 1.0, 0.1], [1.5, 1.0, 0.5, 0.1]])
    d = dict(ref=ref, batch=batch, sigma=sigma)
    r64 = ref.astype(np.float64)
    d["ref_proc"] = im3d.apply_gaussian_filter(im3d.normalize(r64, ref=None, channel_normalization="joint"),
                                               sigma, mode="reflect", truncate=4.0)
    d["batch_proc"] = im3d.apply_gaussian_filter(im3d.normalize(batch, ref=r64, channel_normalization="joint"),
                                                 sigma, mode="reflect", truncate=4.0)
    d["batch_proc_sep"] = im3d.apply_gaussian_filter(
        im3d.normalize(batch, ref=r64, channel_normalization="separate"), sigma, mode="reflect", truncate=4.0)
    bu16 = batch.astype(np.uint16)
    d["batch_u16"] = bu16
    d["batch_u16_proc"] = im3d.apply_gaussian_filter(
        im3d.normalize(bu16, ref=r64, channel_normalization="joint"), sigma, mode="reflect", truncate=4.0)
    save("preprocess", **d)


def gen_preprocess_t():
    """Temporal pre-filter: 5-D batches are filtered in 4-D (t, z, y, x) per channel (image_processing_3D.py:140-156)."""
    rng = np.random.default_rng(12)
    ref = (np.stack([synth_volume((8, 14, 18), 30 + c) for c in range(2)], -1) * 3000 + 100).astype(np.float32)
    batch = (ref[None] * (1 + 0.05 * rng.standard_normal((5,) + ref.shape))).astype(np.float32)
    sigma = np.array([[1.0, 1.0, 1.0, 0.8], [1.5, 1.0, 0.7, 1.2]])
    r64 = ref.astype(np.float64)
    save("preprocess_t", ref=ref, batch=batch, sigma=sigma,
         batch_proc=im3d.apply_gaussian_filter(im3d.normalize(batch, ref=r64, channel_normalization="joint"),
                                               sigma, mode="reflect", truncate=4.0))


def gen_sequence_update_ref():
    """compensate_arr_3D with update_reference=True (the fixed volume re-averaged from compensated frames after
    every batch, compensate_recording_3D.py:395-429)."""
    from flowreg3d.motion_correction.compensate_arr_3D import compensate_arr_3D
    from flowreg3d.motion_correction.OF_options_3D import OFOptions
    g = np.load(OUT / "sequence.npz")
    video, ref = g["video"][:6, :12, :24, :28], g["ref"][:12, :24, :28]
    opts = OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=8, update_lag=4, buffer_size=3,
                     weight=[0.5, 0.5], update_reference=True)
    reg, w = compensate_arr_3D(video, ref, opts)
    save("sequence_update_ref", registered=reg.astype(np.float32), w=w.astype(np.float32),
         params=np.array([2, 8, 4, 3]))


def gen_sequence():
    """compensate_arr_3D through the reference's own BatchMotionCorrector (sequential executor)."""
    from flowreg3d.motion_correction.OF_options_3D import OFOptions
    from flowreg3d.motion_correction.compensate_arr_3D import compensate_arr_3D
    from flowreg3d.motion_correction.compensate_recording_3D import (
        BatchMotionCorrector, RegistrationConfig)
    import flowreg3d.motion_correction.compensate_arr_3D as mod

    Z, Y, X, C, T = 16, 40, 44, 2, 7
    ref = np.stack([synth_volume((Z, Y, X), 30 + c) for c in range(C)], -1)
    rng = np.random.default_rng(12)
    frames = []
    for t in range(T):
        np.random.seed(1000 + t)
        g = get_low_disp_3d_generator()(Z, Y, X)[0] * 0.4
        r64 = ref.astype(np.float64)
        fr = R.imregister_wrapper(r64, -g[..., 0], -g[..., 1], -g[..., 2], r64, "linear")
        frames.append(fr + 0.01 * rng.standard_normal(fr.shape).astype(np.float32))
    video = np.stack(frames, 0).astype(np.float32)
    opts = OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=20, update_lag=5,
                     buffer_size=3, weight=[0.5, 0.5], save_meta_info=False, output_typename=None)
    # force the sequential executor (the reference's own reference implementation)
    orig = mod.BatchMotionCorrector

    class _Seq(BatchMotionCorrector):
        def __init__(self, options, config=None):
            super().__init__(options, RegistrationConfig(parallelization="sequential", verbose=True))

    mod.BatchMotionCorrector = _Seq
    try:
        reg, w = compensate_arr_3D(video, ref, opts)
    finally:
        mod.BatchMotionCorrector = orig
    save("sequence", video=video, ref=ref, registered=np.asarray(reg, np.float32), w=np.asarray(w, np.float32),
         params=np.array([2, 20, 5, 3]))


def gen_config1():
    """BASELINE config 1: 64x128x128x1 pair, default OFOptions parameters (subsampled outputs)."""
    Z, Y, X = 64, 128, 128
    V = synth_volume((Z, Y, X), 1)
    d = {}
    for name, gen, seed, scale in (("low", get_low_disp_3d_generator, 1, 1.0),):
        np.random.seed(seed)
        g = gen()(Z, Y, X)[0] * scale
        V64 = V.astype(np.float64)
        mov = R.imregister_wrapper(V64, -g[..., 0], -g[..., 1], -g[..., 2], V64, "linear")
        sigma = np.array([[1.0, 1.0, 1.0, 0.1]])
        fp = im3d.apply_gaussian_filter(im3d.normalize(V64[..., None], ref=None), sigma)
        mp = im3d.apply_gaussian_filter(im3d.normalize(mov.astype(np.float64)[..., None], ref=V64[..., None]), sigma)
        flow = R.get_displacement(fp, mp, alpha=(0.25,) * 3, levels=100, min_level=5, eta=0.8, update_lag=5,
                                  iterations=100, a_smooth=1.0, a_data=0.45, weight=np.ones((Z, Y, X, 1)))
        f32 = flow.astype(np.float32)
        reg = R.imregister_wrapper(mov, f32[..., 0], f32[..., 1], f32[..., 2], V64, "cubic")
        d[f"{name}_moving"] = mov.astype(np.float32)
        d[f"{name}_flow_s2"] = f32[::2, ::2, ::2]
        d[f"{name}_reg_s2"] = reg[::2, ::2, ::2].astype(np.float32)
        d[f"{name}_flow_stats"] = np.array([np.abs(flow).max(), np.sqrt((flow ** 2).sum(-1)).mean(),
                                            np.sqrt(((flow - g) ** 2).sum(-1)).mean()])
        print(name, "EPE of reference vs ground truth:", d[f"{name}_flow_stats"])
    d["fixed_checksum"] = np.array([V.astype(np.float64).sum(), float(V[7, 11, 13])])
    save("config1", **d)


def gen_schedule():
    """Level schedules of the LIVE reference driver (core/optical_flow_3d.py:389-408, 485-490): get_displacement is
    run with its per-level work patched out (resize -> zeros of the requested size, warp / motion tensor / median ->
    trivial, level_solver -> records the level size and the scaled alpha it was handed)."""
    import json
    rec = []

    def fake_solver(J11, *a):
        alpha_tmp = a[13]
        rec.append((tuple(int(x) - 2 for x in J11.shape[:3]), [float(x) for x in alpha_tmp]))
        z = np.zeros(J11.shape[:3])
        return z, z.copy(), z.copy()

    def fake_resize(img, size):
        return np.zeros(tuple(size) + tuple(np.shape(img)[3:]), np.float64)

    saved = {k: getattr(R, k) for k in ("level_solver", "resize", "imregister_wrapper", "get_motion_tensor_gc",
                                        "median_filter")}
    R.level_solver = fake_solver
    R.resize = fake_resize
    R.imregister_wrapper = lambda f2, u, v, w, f1, *a, **k: f2
    R.get_motion_tensor_gc = lambda f1, f2, hz, hy, hx: [np.zeros(tuple(x + 2 for x in f1.shape))] * 10
    R.median_filter = lambda a, **k: a
    out = []
    try:
        cases = [((32, 512, 512), 0.8, 100, ml) for ml in (0, 2, 5, 6, 9)]
        cases += [((64, 128, 128), 0.8, 100, ml) for ml in (0, 2, 5)]
        cases += [((64, 256, 256), 0.8, 100, ml) for ml in (0, 5)]
        cases += [((128, 1024, 1024), 0.8, 100, ml) for ml in (0, 2, 5)]
        cases += [((24, 48, 56), 0.8, 100, 0), ((24, 48, 56), 0.5, 3, 0), ((16, 40, 44), 0.8, 100, 2),
                  ((9, 9, 200), 0.8, 100, 0), ((33, 31, 17), 0.7, 100, 1), ((12, 300, 20), 0.9, 7, 3)]
        for shape, eta, levels, ml in cases:
            rec.clear()
            z = np.zeros(shape + (1,))
            R.get_displacement(z, z, alpha=(1.0, 2.0, 3.0), update_lag=1, iterations=1, min_level=ml, levels=levels,
                               eta=eta, a_smooth=1.0, a_data=0.45)
            out.append({"shape": list(shape), "eta": eta, "levels": levels, "min_level": ml,
                        "sizes": [list(r[0]) for r in rec], "alpha": [r[1] for r in rec]})
            print(shape, eta, levels, ml, "->", len(rec), "levels, finest", rec[-1][0])
    finally:
        for k, v in saved.items():
            setattr(R, k, v)
    (OUT / "schedule.json").write_text(json.dumps(out, indent=0))


def gen_config2():
    """BASELINE config 2: ONE frame of the 2-channel 32x512x512 recording through the live reference at OFOptions
    defaults (normalise + Gaussian pre-filter, get_displacement with min_level 5 -> levels 8x134x134 and 10x168x168,
    100 iterations, then the cubic compensation warp of the raw frame).  Inputs are rebuilt by the test from
    tests_inputs (synth_volume seeds 10, 11; smooth_flow seed 1000; noise default_rng(2000)); outputs are stored on
    every second z plane and every fourth y / x sample."""
    sys.path.insert(0, str(OUT.parent))
    from tests_inputs import smooth_flow
    Z, Y, X, C = 32, 512, 512, 2
    ref = np.stack([synth_volume((Z, Y, X), 10 + c) for c in range(C)], -1)
    g = smooth_flow((Z, Y, X), 1000, 2.0, 12.0)
    r64 = ref.astype(np.float64)
    mov = R.imregister_wrapper(r64, -g[..., 0], -g[..., 1], -g[..., 2], r64, "linear")
    mov = (mov + 0.01 * np.random.default_rng(2000).standard_normal(mov.shape)).astype(np.float32)
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]] * C)
    fp = im3d.apply_gaussian_filter(im3d.normalize(r64, ref=None), sigma)
    mp = im3d.apply_gaussian_filter(im3d.normalize(mov.astype(np.float64), ref=r64), sigma)
    flow = R.get_displacement(fp, mp, alpha=(0.25,) * 3, levels=100, min_level=5, eta=0.8, update_lag=5,
                              iterations=100, a_smooth=1.0, a_data=0.45, weight=np.full((Z, Y, X, C), 0.5))
    f32 = flow.astype(np.float32)
    reg = R.imregister_wrapper(mov, f32[..., 0], f32[..., 1], f32[..., 2], r64, "cubic")
    epe = np.sqrt(((flow - g) ** 2).sum(-1))
    print("config2: EPE of the reference vs ground truth mean %.4f max %.4f" % (epe.mean(), epe.max()))
    save("config2", flow_s=f32[::2, ::4, ::4], reg_s=np.asarray(reg, np.float32)[::2, ::4, ::4],
         moving_checksum=np.array([mov.astype(np.float64).sum(), float(mov[7, 11, 13, 1])]),
         proc_s=mp[::4, ::8, ::8].astype(np.float32),
         stats=np.array([epe.mean(), epe.max(), np.abs(flow).max()]))


def _stub_skimage():
    """scikit-image is absent here: stub skimage.registration.phase_cross_correlation with the oracle's restatement."""
    sys.path.insert(0, str(OUT.parent.parent))
    from oracle import xcorr as OX
    sk = types.ModuleType("skimage")
    skr = types.ModuleType("skimage.registration")
    skr.phase_cross_correlation = lambda a, b, **kw: (OX.phase_cross_correlation(a, b, **kw), None, None)
    sk.registration = skr
    sys.modules["skimage"] = sk
    sys.modules["skimage.registration"] = skr


def gen_xcorr_sequence():
    """compensate_arr_3D of the LIVE reference with cc_initialization=True over two batches (sequential executor,
    skimage stubbed as in gen_xcorr): pins the w_init bookkeeping of that mode -- a ZERO field for the first batch, no
    bootstrap solve (compensate_recording_3D.py:346-356), then the mean of the previous batch's flows (:481-485)."""
    _stub_skimage()
    from flowreg3d.motion_correction.OF_options_3D import OFOptions
    from flowreg3d.motion_correction.compensate_recording_3D import BatchMotionCorrector, RegistrationConfig
    import flowreg3d.motion_correction.compensate_arr_3D as mod
    sys.path.insert(0, str(OUT.parent))
    from scipy.ndimage import shift as ndi_shift
    Z, Y, X, T = 12, 40, 48, 4
    ref = synth_volume((Z, Y, X), 3)
    rng = np.random.default_rng(7)
    shifts = [(4.0, -3.0, 1.0), (3.5, -2.5, 0.5), (-2.0, 1.5, -1.0), (1.0, 2.0, 0.0)]
    video = np.stack([ndi_shift(ref, shift=(s[2], s[1], s[0]), order=1, mode="nearest")
                      + 0.002 * rng.standard_normal(ref.shape).astype(np.float32) for s in shifts], 0)
    video = video.astype(np.float32)[..., None]
    opts = OFOptions(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, iterations=10, update_lag=5, buffer_size=2,
                     cc_initialization=True, cc_hw=64, cc_up=10, save_meta_info=False, output_typename=None)
    orig = mod.BatchMotionCorrector

    class _Seq(BatchMotionCorrector):
        def __init__(self, options, config=None):
            super().__init__(options, RegistrationConfig(parallelization="sequential", verbose=True))

    mod.BatchMotionCorrector = _Seq
    try:
        reg, w = mod.compensate_arr_3D(video, ref[..., None], opts)
    finally:
        mod.BatchMotionCorrector = orig
    print("xcorr_sequence: mean flow per frame", np.asarray(w).reshape(T, -1, 3).mean(1))
    save("xcorr_sequence", video=video, shifts=np.array(shifts), registered=np.asarray(reg, np.float32),
         w=np.asarray(w, np.float32))


def gen_xcorr():
    """Rigid cross-correlation pre-alignment (util/xcorr_prealignment.py, sequential_3d.py:89-145).
    scikit-image is absent here, so the live reference is run with `skimage.registration` STUBBED by the oracle's
    restatement of phase_cross_correlation (oracle/xcorr.py): the golden pins everything the reference itself
    does around that call (projections, 2-D resize, whitening, Hann window, scaling, sign, the six executor
    steps), not scikit-image's arithmetic."""
    _stub_skimage()
    from flowreg3d.util.xcorr_prealignment import estimate_rigid_xcorr_3d
    from flowreg3d.motion_correction.parallelization.sequential_3d import SequentialExecutor3D
    from scipy.ndimage import shift as ndi_shift
    rng = np.random.default_rng(21)
    Z, Y, X, C = 24, 72, 96, 2
    ref = np.stack([synth_volume((Z, Y, X), 50 + c) for c in range(C)], -1)
    d = {}
    shifts = [(3.2, -1.5, 2.0), (-4.6, 2.4, -1.0)]
    movs = []
    for k, sh in enumerate(shifts):
        mov = np.stack([ndi_shift(ref[..., c], shift=(sh[2], sh[1], sh[0]), order=1, mode="nearest")
                        for c in range(C)], -1)
        mov = (mov + 0.01 * rng.standard_normal(mov.shape)).astype(np.float32)
        movs.append(mov)
        d[f"est{k}_w"] = estimate_rigid_xcorr_3d(ref, mov, target_hw=(48, 64), up=10,
                                                 weight=np.array([0.3, 0.7], np.float32))
        d[f"est{k}_full"] = estimate_rigid_xcorr_3d(ref, mov, target_hw=None, up=20)
        d[f"est{k}_z"] = estimate_rigid_xcorr_3d(ref[..., 0], mov[..., 0], target_hw=(72, 48), target_z=12, up=5)
        print("xcorr", sh, d[f"est{k}_w"], d[f"est{k}_full"], d[f"est{k}_z"])
    # The executor steps.  NOTE (reference behaviour): BatchMotionCorrector passes the FULL (Z,Y,X,C) weight array in
    # flow_params (compensate_recording_3D.py:212-222, 303), which estimate_rigid_xcorr_3d reshapes to a vector and
    # contracts with the channel axis (xcorr_prealignment.py:26-30) -> ValueError("shape-mismatch for sum") for
    # C > 1.  cc_initialization therefore only runs for single-channel recordings in the reference; the golden
    # uses C = 1 and additionally records that the 2-channel call raises.
    batch = np.stack(movs, 0)[..., :1]
    ref1 = ref[..., :1]
    sigma = np.array([[1.0, 1.0, 1.0, 0.1]])
    rp = im3d.apply_gaussian_filter(im3d.normalize(ref1.astype(np.float64), ref=None), sigma)
    bp = im3d.apply_gaussian_filter(im3d.normalize(batch.astype(np.float64), ref=ref1.astype(np.float64)), sigma)
    w_init = np.zeros((Z, Y, X, 3), np.float32)
    w_init[..., 0] = 0.5
    fp = dict(alpha=(0.25, 0.25, 0.25), levels=100, min_level=2, eta=0.8, update_lag=4, iterations=8,
              a_smooth=1.0, a_data=0.45, weight=np.ones((Z, Y, X, 1)), cc_initialization=True, cc_hw=(48, 64),
              cc_up=10)
    ex = SequentialExecutor3D()
    reg, flows = ex.process_batch(batch, bp, ref1, rp, w_init, R.get_displacement, R.imregister_wrapper, "cubic",
                                  None, flow_params=fp)
    raised = 0
    try:
        fp2 = dict(fp, weight=np.full((Z, Y, X, 2), 0.5))
        b2 = np.stack(movs, 0)
        ex.process_batch(b2, b2.astype(np.float64), ref, ref.astype(np.float64), w_init, R.get_displacement,
                         R.imregister_wrapper, "cubic", None, flow_params=fp2)
    except ValueError as e:
        raised = 1
        print("2-channel cc_initialization raises:", e)
    d["two_channel_raises"] = np.array([raised])
    # ref = stack(synth_volume((24,72,96), 50 + c)) is rebuilt by the test (tests_inputs.synth_volume); flows and the
    # registered frames are stored on every other y/x sample
    save("xcorr", ref_checksum=np.array([ref.astype(np.float64).sum(), float(ref[5, 7, 11, 1])]),
         batch=np.stack(movs, 0), shifts=np.array(shifts), w_init=w_init,
         registered_s2=np.asarray(reg, np.float32)[:, :, ::2, ::2], flows_s2=np.asarray(flows, np.float32)[:, :, ::2, ::2],
         **d)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tables", "resize", "warp", "motion_tensor", "motion_tensor_alt", "solver", "flow_small",
                             "preprocess", "preprocess_t", "sequence", "sequence_update_ref", "config1", "config2", "schedule", "xcorr", "xcorr_sequence"]
    for w in which:
        globals()[f"gen_{w}"]()
