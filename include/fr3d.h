/*
 * fr3d.h -- C ABI of libfr3d: B200-native (sm_100a) dense 3-D variational optical-flow
 * registration, the hot path of flowreg3D.
 *
 * The reference (FlowRegSuite/flowreg3D) is pure Python and has no FFI; the seams this library
 * sits under are the reference's Python-level interfaces (paths relative to src/flowreg3d/):
 *   B1  BaseExecutor3D.process_batch        motion_correction/parallelization/base_3d.py:38-71
 *   B2  get_displacement / imregister_wrapper   core/optical_flow_3d.py:319-333 / :22
 *   B3  OFOptions                           motion_correction/OF_options_3D.py:130-686
 * flowreg3d_b200/ mirrors those interfaces in Python and calls the entry points below through
 * ctypes.  INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - return 0 on success, a negative fr3d_status otherwise; never throws; message via
 *     fr3d_last_error().
 *   - unless a parameter says "host", pointers are CUDA DEVICE pointers owned by the caller.
 *   - volumes are row-major (Z, Y, X[, C]) channels-last, exactly as the reference lays them out;
 *     flow fields are (Z, Y, X, 3) with [...,0]=u=dx (X axis), [...,1]=v=dy, [...,2]=w=dz, in
 *     full-resolution voxel units.
 *   - all work is enqueued on the context's stream; calls return without synchronising unless
 *     stated.  A context is not thread-safe; use one per thread/stream.
 *   - the host computes the (tiny) level schedule and tap tables (numpy expressions identical to
 *     the reference's, see flowreg3d_b200/plan.py) and passes them in fr3d_plan; the library does
 *     every per-voxel computation on the GPU.  There is no CPU fallback.
 */
#ifndef FR3D_H
#define FR3D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FR3D_MAX_CHANNELS 4
#define FR3D_MAX_LEVELS 64
#define FR3D_ABI_VERSION 5

typedef enum {
    FR3D_OK = 0,
    FR3D_ERR_ARG = -1,     /* bad argument / unsupported option */
    FR3D_ERR_CUDA = -2,    /* CUDA runtime error */
    FR3D_ERR_NOMEM = -3,   /* device or host allocation failed */
    FR3D_ERR_STATE = -4    /* call order (e.g. no reference set) */
} fr3d_status;

typedef enum {
    FR3D_F32 = 0,
    FR3D_F64 = 1,
    FR3D_U8 = 2,
    FR3D_U16 = 3,
    FR3D_I16 = 4,
    FR3D_I32 = 5
} fr3d_dtype;

typedef enum {
    FR3D_SWEEP_LEXICOGRAPHIC = 0, /* hyperplane wavefront: reproduces the reference's sweep order */
    FR3D_SWEEP_REDBLACK = 1       /* checkerboard half-sweeps (even k+j+i first), du->dv->dw sequential inside a
                                   * voxel; opt-in throughput mode: it does NOT reproduce the reference's
                                   * lexicographic result (mean 0.015, max 0.4 voxel apart) */
} fr3d_sweep;

/* Per-axis resampling table of the fused Gauss (x) Keys-cubic resize
 * (util/resize_util_3D.py:76-111): out[i] = sum_p wt[i*P+p] * src[idx[i*P+p]].  HOST pointers. */
typedef struct {
    int32_t in_len, out_len, P;
    const int32_t* idx; /* host, out_len*P */
    const float* wt;    /* host, out_len*P */
} fr3d_axis_table;

/* One pyramid level (core/optical_flow_3d.py:403-529), listed coarse -> fine. */
typedef struct {
    int32_t size[3];              /* level grid (pz, py, px) */
    double h[3];                  /* (hz, hy, hx) = full_dim / level_dim */
    double alpha[3];              /* (x, y, z) regularisation, already scaled for the level */
    int32_t median;               /* 1: 5x5x5 median of the increments (min(size) > 5) */
    fr3d_axis_table from_full[3]; /* (x, y, z) tables full resolution -> this level */
    fr3d_axis_table from_prev[3]; /* (x, y, z) tables previous (coarser) level -> this level; unused at [0] */
} fr3d_level;

typedef struct {
    int32_t abi_version;          /* FR3D_ABI_VERSION */
    int32_t Z, Y, X, C;           /* full-resolution volume, channels */
    int32_t max_batch;            /* frames processed concurrently per call */
    int32_t n_levels;
    const fr3d_level* levels;     /* host, n_levels entries, coarse -> fine */
    fr3d_axis_table to_full[3];   /* (x, y, z) finest solved level -> full resolution; P = 0 if min_level == 0 */
    int32_t iterations, update_lag;
    double a_data[FR3D_MAX_CHANNELS];
    double a_smooth;              /* 1.0: linear smoothness; otherwise the nonlinear term psi_s = a (|grad|^2 + 1e-5)^(a-1),
                                   * recomputed every sweep (level_solver_3d.py:262-311); lexicographic sweep only */
    int32_t sweep;                /* fr3d_sweep */
    int32_t interp;               /* compensation warp: 3 = cubic B-spline, 1 = trilinear */
    int32_t state_dtype;          /* solver state storage (du,dv,dw and the constant Laplacian term): FR3D_F64 (what the
                                   * Python host passes by default: the reference to float64 rounding) or FR3D_F32
                                   * (24 % fewer solver bytes; inside the 0.01 / 0.05 voxel tolerance on configs 2 and
                                   * 4, outside on one measured workload -- DESIGN.md 4.1).  The system matrix and all
                                   * arithmetic are float64 either way. */
    /* pre-filter (util/image_processing_3D.py:95-162): normalised Gaussian half-kernels
     * w[0..r] (w[0] = centre) per channel and axis (z, y, x); r = 0 means identity.  HOST. */
    int32_t gauss_radius[FR3D_MAX_CHANNELS][3];
    const double* gauss_w[FR3D_MAX_CHANNELS][3];
    /* temporal axis of the 4-D (t,z,y,x) filter the reference applies to 5-D batches
     * (image_processing_3D.py:140-156): filtered first, reflect at the ends of the B frames of one
     * fr3d_preprocess call.  r = 0 (sigma_t < 0.125, the reference default 0.1) means identity. */
    int32_t gauss_radius_t[FR3D_MAX_CHANNELS];
    const double* gauss_w_t[FR3D_MAX_CHANNELS];
} fr3d_plan;

typedef struct fr3d_ctx fr3d_ctx;

/* ---- context ---------------------------------------------------------------------------- */
/* plan may be NULL: a "bare" context that only serves the stage entry points. stream: a
 * cudaStream_t cast to void* (NULL = the legacy default stream). */
int fr3d_create(fr3d_ctx** out, int device, const fr3d_plan* plan, void* stream);
void fr3d_destroy(fr3d_ctx* ctx);
const char* fr3d_last_error(const fr3d_ctx* ctx); /* ctx may be NULL: last create() error */
int fr3d_abi_version(void);
int fr3d_synchronize(fr3d_ctx* ctx);
/* number of kernel launches issued by this context so far (for bench.py's gpu_launches) */
int64_t fr3d_launch_count(const fr3d_ctx* ctx);
/* bytes of device memory currently held by the context */
int64_t fr3d_device_bytes(const fr3d_ctx* ctx);

/* Tuning knobs (do not change results, except FR3D_OPT_WARP_FACTORED as documented). */
typedef enum {
    FR3D_OPT_SOR_CTAS_PER_SM = 1, /* resident CTAs per SM the persistent solver kernel may claim (0 = all it can
                                   * get; 1 leaves room for a second stream's kernels on every SM) */
    FR3D_OPT_WARP_FACTORED = 3,   /* cubic warps of 1-2 channels: 1 = factored separable sum (3x fewer float64
                                   * operations; float32 results differ from scipy's association in the last bit on
                                   * ~1e-7 of the voxels), 0 (default) = scipy's ((c*wz)*wy)*wx accumulate, bit-equal.
                                   * THE ONE KNOB THAT CAN CHANGE RESULTS (by <= 1 float32 ulp). */
    FR3D_OPT_CC_BLOCK_SCANS = 2,  /* rigid pre-alignment: 1 = block-cooperative plane scans (arg-max, tile sums,
                                   * plane mean: one CTA per plane) instead of one thread per plane (0) */
    FR3D_OPT_SOR_KERNEL = 4,      /* level solver kernel: 0 (default) = direct-load wavefront; 1 = staged wavefront
                                   * (cp.async.bulk + mbarrier ring per warp); 2 = time-blocked skewed tiles, increments
                                   * resident in shared memory for Tb sweeps.  Same arithmetic, same update order:
                                   * results are bit-identical (1 and 2 measured slower on B200) */
    FR3D_OPT_SOR_STAGES = 5,      /* shared-memory stages per warp of the staged solver kernel (0 = built-in default) */
    FR3D_OPT_SPLINE_TMA = 7,      /* B-spline prefilter X pass: 1 (default) = the block's lines are staged with ONE bulk copy
                                   * global -> shared (cp.async.bulk + mbarrier) and written back with one bulk copy
                                   * shared -> global; 0 = per-thread staging loops.  Same arithmetic: bit-identical */
    FR3D_OPT_RESIZE_X_ROWS = 8,   /* pyramid X pass: 1 (default) = a thread resamples 4 rows at one output position and
                                   * looks the taps up once; 0 = one output per thread.  Bit-identical */
    FR3D_OPT_SOR_SCHED = 9,       /* direct-load wavefront kernel, how a wave's work items reach the warps: bits 0-6 =
                                   * percent (0..100) of the items handed out through a per-wave ticket counter instead
                                   * of round-robin; bit 7 = psi-refresh items enumerated first, equally many per warp
                                   * (measured slower on B200).  -1 (default) = 20 for the float64 state, 0 for float32.
                                   * Scheduling only: results are bit-identical */
    FR3D_OPT_SOR_FRAMES_PER_ITEM = 10, /* frames one warp work item of the wavefront kernels covers (0 = default 2; tuning aid,
                                   * bit-identical) */
    FR3D_OPT_WARP_TILE = 11,      /* gather block shape in outputs, tx | ty << 8 | tz << 16 with tx*ty*tz == 256 (0 = default
                                   * 32 x 8 x 1; tuning aid, bit-identical) */
    FR3D_OPT_SOR_TILE = 6         /* tile kernel geometry: Tb | K << 8 | J << 16 | I << 24 (sweeps per time block and
                                   * tile extents; a 0 field keeps its default: 5 sweeps, 8 x 8 x 8) */
} fr3d_option;
int fr3d_set_option(fr3d_ctx* ctx, int option, int64_t value);

/* Per-kernel timing with CUDA events on the context's stream (used by bench.py for the roofline
 * figures).  fr3d_profile_report synchronises, writes "name<TAB>launches<TAB>total_ms" lines into buf
 * (host) and clears the record; returns the byte length of the full report or a negative status. */
int fr3d_profile_enable(fr3d_ctx* ctx, int on);
int64_t fr3d_profile_report(fr3d_ctx* ctx, char* buf, int64_t cap);

/* ---- pipeline (needs a plan) ------------------------------------------------------------ */
/* Normalise + Gaussian pre-filter (compensate_recording_3D.py:229-254): out = G * ((raw-lo)/den),
 * float64 math, one rounding to float32 (the reference's first resize rounds it the same way).
 * raw: (B,Z,Y,X,C) of `dtype`; lo, den: host, C doubles; out: (B,Z,Y,X,C) float32.  With a temporal radius
 * in the plan and temporal != 0 the B frames of the call are one batch of the reference (filtered across
 * frames first); temporal = 0 is the reference's 4-D (Z,Y,X,C) case (fixed volume: spatial filter only).
 * out64 (optional, may be NULL): the same result before its rounding to float32 -- what the reference keeps in
 * float64 and warps when it re-averages the reference volume (update_reference). */
int fr3d_preprocess(fr3d_ctx* ctx, const void* raw, int dtype, int B, const double* lo,
                    const double* den, int temporal, float* out, double* out64);

/* Cache the fixed volume's pyramid and the weight pyramid (frame-invariant).
 * ref_proc: (Z,Y,X,C) float32 pre-processed reference.  weight: (Z,Y,X,C) float32 or NULL, in which
 * case weight_const (host, C doubles, already normalised) is broadcast. */
int fr3d_set_reference(fr3d_ctx* ctx, const float* ref_proc, const float* weight,
                       const double* weight_const);

/* get_displacement for B frames against the cached reference (core/optical_flow_3d.py:319-542).
 * moving_proc: (B,Z,Y,X,C) float32.  uvw_init: (Z,Y,X,3) float32 shared by all B frames, or NULL.
 * flow_out: (B,Z,Y,X,3) of out_dtype (FR3D_F32 or FR3D_F64). */
int fr3d_get_displacement(fr3d_ctx* ctx, const float* moving_proc, const float* uvw_init, int B,
                          void* flow_out, int out_dtype);

/* The same computation one pyramid level at a time (levels 0 .. fr3d_level_count()-1, coarse to fine):
 *   fr3d_level_begin   everything before the solve (level image, flow from the coarser level, warp, assembly)
 *   fr3d_level_sweeps  sweeps t_begin <= t < t_end restricted to waves q_begin <= q < q_end of the global
 *                      wavefront schedule q = (k+j+i) + 2t (negative bounds = everything).  t_begin must be a
 *                      multiple of update_lag.  Lexicographic sweep, a_smooth == 1 only for partial ranges.
 *   fr3d_level_state   copy the increments of solver-storage slots [slot_begin, slot_end) of all B frames
 *                      out of (direction 0) or into (direction 1) `ext` (device, B x (slot_end-slot_begin)
 *                      state vectors: {du,dv,dw,-} float32 (16 bytes) or {du,dv,dw} float64 (24 bytes)).  Hyperplane s = k+j+i occupies the slots
 *                      hyperplane_start[s] .. hyperplane_start[s+1] reported by fr3d_level_info, so a range of
 *                      hyperplanes is one contiguous block per frame.
 *   fr3d_level_end     increments -> 5^3 median -> accumulate into the flow
 *   fr3d_flow_finish   flow of the finest level -> (B,Z,Y,X,3) full resolution
 * This is the seam of the sweep-pipelined multi-GPU solve of a single large volume
 * (flowreg3d_b200/multigpu.py): rank r runs the sweeps [t_r, t_{r+1}) of every level; the increments
 * stream rank r -> r+1 hyperplane block by hyperplane block over NVLink (one direction only, because sweep
 * t+1 of a hyperplane needs nothing but sweep t of itself and its two neighbours). */
int fr3d_level_count(const fr3d_ctx* ctx);
/* size: (pz,py,px); hyperplane_start: host, n_hyperplanes+1 entries (may be NULL). */
int fr3d_level_info(const fr3d_ctx* ctx, int level, int32_t size[3], int32_t* n_hyperplanes,
                    int64_t* n_slots, int32_t* hyperplane_start);
int fr3d_level_begin(fr3d_ctx* ctx, int level, const float* moving_proc, const float* uvw_init, int B);
int fr3d_level_sweeps(fr3d_ctx* ctx, int level, int t_begin, int t_end, int q_begin, int q_end);
int fr3d_level_state(fr3d_ctx* ctx, int level, int direction, void* ext, int64_t slot_begin,
                     int64_t slot_end);
/* z-slab variant of the seam (flowreg3d_b200/multigpu.py, mode "zslab"): rank r owns the planes [k_begin, k_end)
 * of the level.  fr3d_level_sweeps_slab runs the waves [q_begin, q_end) of ALL sweeps restricted to those planes
 * (a voxel of wave q reads nothing newer than wave q-1, so the ranks exchange, after every wave, the boundary planes
 * they own); fr3d_level_planes copies whole planes of the increments out of (direction 0) or into (1) `ext`
 * (device, B x (k_end-k_begin) x py x px state vectors: 4 float32 or 3 float64 components). */
int fr3d_level_sweeps_slab(fr3d_ctx* ctx, int level, int q_begin, int q_end, int k_begin, int k_end);
int fr3d_level_planes(fr3d_ctx* ctx, int level, int direction, void* ext, int k_begin, int k_end);
/* The cells of plane k that wave q updated (one anti-diagonal per sweep in flight) out of (0) / into (1) `ext`
 * (device, B x iterations x py state vectors; entries without a cell are not touched): the per-wave halo message. */
int fr3d_level_wave_cells(fr3d_ctx* ctx, int level, int direction, void* ext, int k, int q);
/* The same z-slab solve with the halo exchange INSIDE the persistent kernel (one launch per level instead of one
 * launch + two pack kernels + a host synchronisation + a message pair per wave).  One process per GPU; the increment
 * array of the open level (which = 0) and a pair of flag words (which = 1) are shared with the z-neighbours through
 * CUDA IPC: fr3d_ipc_export fills a 64-byte cudaIpcMemHandle_t, fr3d_ipc_open maps a neighbour's handle (peer access
 * over NVLink), fr3d_ipc_close unmaps it.  fr3d_level_sweeps_slab_p2p runs ALL waves of the level on the planes
 * [k_begin, k_end): a voxel on plane k_begin / k_end - 1 is also stored into lo_d / hi_d (the neighbours' arrays),
 * and after every wave the ranks publish flag_base + (waves done) in each other's flag words and wait for their own.
 * flag_base must be the same on all ranks and grow by the level's wave count from launch to launch (the words are
 * never reset).  lo_* / hi_* are NULL for the first / last slab.  No reference counterpart. */
int fr3d_ipc_export(fr3d_ctx* ctx, int which, void* handle_out);
int fr3d_ipc_open(fr3d_ctx* ctx, const void* handle, void** ptr_out);
int fr3d_ipc_close(fr3d_ctx* ctx, void* ptr);
int fr3d_level_sweeps_slab_p2p(fr3d_ctx* ctx, int level, int k_begin, int k_end, void* lo_d, void* hi_d, void* lo_flags,
                               void* hi_flags, int64_t flag_base);
int fr3d_level_end(fr3d_ctx* ctx, int level);
/* fr3d_level_end restricted to the planes z_begin <= z < z_end of the level's flow (the median of a plane needs
 * the increments of two planes on either side, which every rank holds); the other planes of the flow are then
 * fetched from the ranks that computed them:  fr3d_flow_slab copies the planes [z_begin, z_end) of the level's
 * flow out of (direction 0) or into (direction 1) `ext` (device, B x 3 x (z_end - z_begin) x py x px float64). */
int fr3d_level_end_range(fr3d_ctx* ctx, int level, int z_begin, int z_end);
int fr3d_flow_slab(fr3d_ctx* ctx, int level, int direction, double* ext, int z_begin, int z_end);
int fr3d_flow_finish(fr3d_ctx* ctx, void* flow_out, int out_dtype);

/* Compensation warp of B raw frames (parallelization/sequential_3d.py:153-160):
 * out(x) = vol(x + flow(x)), plan.interp interpolation, out-of-volume voxels take ref's value.
 * vol: (B,Z,Y,X,C) of vol_dtype; flow: (B,Z,Y,X,3) float32; ref: (Z,Y,X,C) of ref_dtype;
 * out: (B,Z,Y,X,C) float32. */
int fr3d_compensate(fr3d_ctx* ctx, const void* vol, int vol_dtype, const float* flow,
                    const void* ref, int ref_dtype, int B, float* out);

/* ---- stage entry points (any context; explicit sizes; used by the stage-wise parity tests
 *      and by the per-pair imregister_wrapper shim) ---------------------------------------- */
/* src: nvol planar volumes (D,H,W) float32 -> dst: nvol x (od,oh,ow); tables host, order (x,y,z). */
int fr3d_resize3d(fr3d_ctx* ctx, const float* src, int nvol, int D, int H, int W,
                  const fr3d_axis_table tables[3], float* dst);

/* imregister_wrapper (core/optical_flow_3d.py:22-74) on an arbitrary volume:
 * vol (Z,Y,X,C) vol_dtype; u,v,w (Z,Y,X) float64 displacements in voxels; ref (Z,Y,X,C) ref_dtype;
 * out (Z,Y,X,C) float32; interp 3|1. */
int fr3d_warp(fr3d_ctx* ctx, const void* vol, int vol_dtype, const double* u, const double* v,
              const double* w, const void* ref, int ref_dtype, int Z, int Y, int X, int C,
              int interp, float* out);

/* get_motion_tensor_gc (core/optical_flow_3d.py:92-152) for one channel:
 * f1, f2 (p,m,n) float32; f2_f32_math != 0 reproduces numpy's float32 arithmetic on a float32 f2
 * (every level below the top); J: (10,p,m,n) float64, order J11,J22,J33,J44,J12,J13,J23,J14,J24,J34
 * (the reference's zero ring is not stored). */
int fr3d_motion_tensor(fr3d_ctx* ctx, const float* f1, const float* f2, int p, int m, int n,
                       double hz, double hy, double hx, int f2_f32_math, double* J);

/* get_motion_tensor_gray (core/optical_flow_3d.py:218-259; kind = 1) and get_motion_tensor_cs (:155-215; kind = 2)
 * for one channel of float64 images f1, f2 (p,m,n); J as above.  Stage functions only: like in the reference, the
 * driver never calls them (get_displacement hard-wires the gradient-constancy tensor, :457). */
int fr3d_motion_tensor_alt(fr3d_ctx* ctx, int kind, const double* f1, const double* f2, int p, int m, int n,
                           double hz, double hy, double hx, double* J);

/* compute_flow_3d (core/level_solver_3d.py:314-546) on interior arrays (no ring):
 * J (C,10,p,m,n) float64; weight (C,p,m,n) float64; uvw (3,p,m,n) float64; alpha (x,y,z) host;
 * a_data host C doubles; state_dtype FR3D_F32|FR3D_F64 (see fr3d_plan); out d (3,p,m,n) float64 =
 * du,dv,dw. */
int fr3d_sor_level(fr3d_ctx* ctx, const double* J, const double* weight, const double* uvw, int p,
                   int m, int n, int C, const double* alpha, double hz, double hy, double hx,
                   int iterations, int update_lag, const double* a_data, double a_smooth,
                   int sweep, int state_dtype, double* d);

/* scipy.ndimage.median_filter(size=5^3, mode="mirror") on nvol planar float64 volumes. */
int fr3d_median5(fr3d_ctx* ctx, const double* src, int nvol, int p, int m, int n, double* dst);

/* numpy.mean(frames, axis=0) of T float32 arrays of n elements: the w_init bootstrap / chaining of
 * BatchMotionCorrector (compensate_recording_3D.py:388, 481-485), float32 accumulation in frame order. */
int fr3d_mean_frames(fr3d_ctx* ctx, const float* frames, int T, int64_t n, float* out);

/* numpy.mean(axis=0) of T float32 arrays accumulated in float64 (the reference re-averaging its fixed volume from
 * compensated frames, compensate_recording_3D.py:395-429); out: n float64. */
int fr3d_mean_frames_f64(fr3d_ctx* ctx, const float* frames, int T, int64_t n, double* out);

/* Per-frame statistics BatchMotionCorrector keeps (compensate_recording_3D.py:488-508), computed where the flow
 * lives: flow (B,Z,Y,X,3) float32 -> out (B,4) float64 (device) = mean |w|, max |w|, mean divergence
 * (numpy.gradient semantics, unit spacing), |mean translation|.  float64 accumulation (the reference
 * accumulates in float32: agreement to ~1e-6 relative, not bit-exact). */
int fr3d_flow_stats(fr3d_ctx* ctx, const float* flow, int B, int Z, int Y, int X, double* out);

/* ---- rigid cross-correlation pre-alignment: cc_initialization=True -----------------------------
 * (util/xcorr_prealignment.py:15-99 estimate_rigid_xcorr_3d; executor steps
 * motion_correction/parallelization/sequential_3d.py:89-145).  Stage entry points; the host logic
 * between them (peak -> refinement window -> wrap disambiguation, i.e. what the reference delegates
 * to skimage.registration.phase_cross_correlation) is flowreg3d_b200/xcorr.py.  "complex" arrays are
 * interleaved (re, im) float64, row-major. */
/* imregister_wrapper for B frames with per-frame float32 flows and an explicit interpolation
 * (sequential_3d.py:92-99, 124-131 use "linear" whatever OFOptions.interpolation_method says). */
int fr3d_warp_flow(fr3d_ctx* ctx, const void* vol, int vol_dtype, const float* flow, const void* ref,
                   int ref_dtype, int B, int Z, int Y, int X, int C, int interp, float* out);
/* _proj_xy / _proj_xz (xcorr_prealignment.py:8-13): vol (B,Z,Y,X) float32 -> pxy (B,Y,X), pxz (B,Z,X)
 * float32; acc64 = 0: numpy's float32 slice-by-slice mean of the float32 warp output, 1: float64
 * accumulation (the float64 reference volume). */
int fr3d_cc_project(fr3d_ctx* ctx, const float* vol, int B, int Z, int Y, int X, int acc64, float* pxy,
                    float* pxz);
/* p - p.mean(), times the separable Hann window hy[:,None]*hx[None,:], all float32
 * (xcorr_prealignment.py:50-58, 81-89); p (B,H,W) float32, hy/hx device float32; out (B,H,W) complex. */
int fr3d_cc_window(fr3d_ctx* ctx, const float* p, int B, int H, int W, const float* hy, const float* hx,
                   double* out);
/* C[b] = A[b] (M,K) x Bm[b] (K,N), complex; batch strides in complex elements, 0 = shared matrix.  The 2-D DFTs
 * of phase_cross_correlation (forward, inverse and the up-sampled refinement window) are products with host-built
 * DFT matrices. */
int fr3d_cc_cgemm(fr3d_ctx* ctx, const double* A, int64_t a_stride, const double* Bm, int64_t b_stride,
                  double* Cm, int M, int N, int K, int nbatch);
/* P[b] = Fr * conj(Fm[b]) (n complex each); normalize != 0: P /= max(|P|, 100 eps_float32) ("phase"). */
int fr3d_cc_cross_power(fr3d_ctx* ctx, const double* Fr, const double* Fm, double* P, int64_t n, int nbatch,
                        int normalize);
/* idx[b] = numpy.argmax(numpy.abs(cc[b])) (first maximum), cc (nbatch, n) complex; idx device int64. */
int fr3d_cc_abs_argmax(fr3d_ctx* ctx, const double* cc, int64_t n, int nbatch, int64_t* idx);
/* scipy.ndimage.shift(img[b], shift[b], mode="grid-wrap", order = 3 if the shift is fractional else 0) ->
 * float32 values (skimage _disambiguate_shift); img_c (B,H,W) complex (real part used), shift_host (B,2) host
 * (sy, sx); work, out (B,H,W) float64 device. */
int fr3d_cc_wrap_shift(fr3d_ctx* ctx, const double* img_c, int B, int H, int W, const double* shift_host,
                       double* work, double* out);
/* sums for the Pearson correlation of the four tiles split at (sy, sx) of the reference plane ref_c (H,W) complex
 * and shifted[b] (H,W): out (B,4,6) device = n, Sa, Sb, Saa, Sbb, Sab; tile = 2*(y >= sy) + (x >= sx). */
int fr3d_cc_tile_sums(fr3d_ctx* ctx, const double* ref_c, const double* shifted, int B, int H, int W,
                      const int* split_host, double* out);
/* w_combined[b] = w_init + rigid[b] (sequential_3d.py:117-121), float32; rigid_host (B,3) host. */
int fr3d_rigid_flow(fr3d_ctx* ctx, const float* w_init, const float* rigid_host, int B, int64_t nvox,
                    float* out);
/* (w_combined + w_residual).astype(float32) (sequential_3d.py:140-141): n elements. */
int fr3d_add_flow(fr3d_ctx* ctx, const float* comb, const double* resid, int64_t n, float* out);

/* ---- host helpers ----------------------------------------------------------------------- */
/* util/resize_util_3D.py:76-95: fill idx/wt (host, out_len*(2R+4)) from the float32 Gaussian g
 * (host, 2R+1 taps; numpy-computed by the caller so that it is bit-identical to the reference). */
int fr3d_fill_resize_table(int in_len, int out_len, const float* g, int R, int32_t* idx, float* wt);

#ifdef __cplusplus
}
#endif
#endif /* FR3D_H */
