// fr3d_api.cu -- C ABI (include/fr3d.h) and the coarse-to-fine level driver
// (core/optical_flow_3d.py:389-541 of the reference) on top of the kernels in fr3d_kernels.h and
// fr3d_sor.h.  Built for sm_100a by flowreg3d_b200/build.py; the FR3D_EMU build of this same file
// is the test-only kernel-logic emulator (tests/emu).
#include <memory>
#ifndef FR3D_EMU
#include <cxxabi.h>
#endif

#include "fr3d_sor.h"
#include "fr3d_xcorr.h"

using namespace fr3d;

namespace {

struct DevTable {
    int in_len = 0, out_len = 0, P = 0;
    Buf<int32_t> idx;
    Buf<float> wt;
    void upload(Device& d, const fr3d_axis_table& t)
    {
        in_len = t.in_len;
        out_len = t.out_len;
        P = t.P;
        if (P > 0) {
            FR3D_REQUIRE(t.idx && t.wt && t.in_len > 0 && t.out_len > 0, "axis table with null data");
            idx.upload(d, t.idx, (size_t)t.out_len * t.P);
            wt.upload(d, t.wt, (size_t)t.out_len * t.P);
        }
    }
};

// Hyperplane-major solver storage of one (p,m,n) grid: host-built row table, device-built
// neighbour / permutation tables (struct HPView in fr3d_kernels.h).
struct HPGeom {
    int p = 0, m = 0, n = 0, S = 0;
    int32_t npad = 0;
    std::vector<int32_t> pe_host;
    Buf<int32_t> rowbase, start, pe, pe4, nbr, perm;
    mutable Buf<int32_t> peR; // chunk prefix over hyperplanes s, s - 2 lag, ... (kind-balanced dealing), built on first use
    mutable int peR_lag = 0;
    HPView view() const
    {
        return HPView{p, m, n, S, npad, rowbase.p, start.p, pe.p, nbr.p, perm.p};
    }
    const int32_t* refresh_prefix(Device& dev, int lag) const
    {
        if (lag < 1 || S == 0 || pe_host.empty())
            return nullptr;
        if (peR_lag != lag) {
            std::vector<int32_t> t(S);
            for (int s = 0; s < S; ++s)
                t[s] = pe_host[s] - (s >= 2 ? pe_host[s - 2] : 0) + (s >= 2 * lag ? t[s - 2 * lag] : 0);
            peR.upload(dev, t.data(), t.size());
            peR_lag = lag;
        }
        return peR.p;
    }
    void build(Device& dev, int p_, int m_, int n_)
    {
        peR_lag = 0;
        p = p_;
        m = m_;
        n = n_;
        S = p + m + n - 2;
        FR3D_REQUIRE((int64_t)p * m * n + 32LL * S < 2147483647LL, "level %dx%dx%d too large for 32-bit slots", p, m, n);
        std::vector<int32_t> rb((size_t)S * p), st((size_t)S + 1);
        pe_host.assign(S, 0);
        int64_t at = 0;
        for (int s = 0; s < S; ++s) {
            st[s] = (int32_t)at;
            int64_t run = 0;
            for (int k = 0; k < p; ++k) {
                int jlo = s - k - (n - 1);
                jlo = jlo < 0 ? 0 : jlo;
                int jhi = s - k;
                jhi = jhi > m - 1 ? m - 1 : jhi;
                rb[(size_t)s * p + k] = (int32_t)(at + run - jlo);
                if (jhi >= jlo)
                    run += jhi - jlo + 1;
            }
            const int64_t padded = (run + 31) / 32 * 32;
            pe_host[s] = (int32_t)(padded / 32) + (s >= 2 ? pe_host[s - 2] : 0);
            at += padded;
        }
        st[S] = (int32_t)at;
        npad = (int32_t)at;
        rowbase.upload(dev, rb.data(), rb.size());
        start.upload(dev, st.data(), st.size());
        pe.upload(dev, pe_host.data(), pe_host.size());
        std::vector<int32_t> p4(S);
        for (int s = 0; s < S; ++s)
            p4[s] = (st[s + 1] - st[s]) / 32 + (s >= 4 ? p4[s - 4] : 0);
        pe4.upload(dev, p4.data(), p4.size());
        nbr.ensure(dev, (size_t)6 * npad);
        perm.ensure(dev, (size_t)npad);
        launch(dev, HPFillK{nbr.p, perm.p, npad}, npad);
        launch(dev, HPBuildK{view(), nbr.p, perm.p}, (int64_t)p * m * n);
    }
};

struct LevelDev {
    int pz = 0, py = 0, px = 0;
    int64_t N = 0;
    double hz = 1, hy = 1, hx = 1;
    double alpha[3] = {0, 0, 0};
    bool median = false;
    DevTable full[3], prev[3];
    HPGeom hp;
    Buf<float> f1;   // (C, N) natural planar: the reference pyramid level
    Buf<double> whp; // (C, npad) solver storage: resized channel weights
};

// element strides of a 5-D view (o0, o1, z, y, x)
struct View {
    int64_t s0, s1, sz, sy, sx;
};
inline View planar(int64_t n1, int64_t D, int64_t H, int64_t W)
{
    return View{n1 * D * H * W, D * H * W, H * W, W, 1};
}

} // namespace

struct fr3d_ctx {
    Device dev;
    std::string err;
    bool has_plan = false, ref_set = false;
    int Z = 0, Y = 0, X = 0, C = 0, max_batch = 0;
    int iterations = 0, update_lag = 1, sweep = 0, interp = 3, state_dtype = FR3D_F32;
    double a_data[FR3D_MAX_CHANNELS] = {0, 0, 0, 0};
    double a_smooth = 1.0;
    std::vector<std::unique_ptr<LevelDev>> levels;
    DevTable to_full[3];
    PreGauss pre;
    Buf<double> pre_w[FR3D_MAX_CHANNELS][3], pre_wt[FR3D_MAX_CHANNELS], gt;
    // workspaces (grow-only)
    Buf<float> t1, t2, f2, tmp, fscr;
    Buf<double> uvw_a, uvw_b, coef, J, AB, dnat, g1, g2, wnat;
    Buf<double> cc_shift;      // rigid pre-alignment: small per-frame parameter blocks
    Buf<int> cc_int;
    Buf<float> cc_f32;
    Buf<char> L, d, U, dold; // (B, npad) Vec4 of the state dtype
    Buf<double> psi_c, psi_r;
    Buf<unsigned> bar;
    Buf<unsigned> p2p_flags; // z-slab halo flags written by the z-neighbours (fr3d_ipc_export which = 1)
    Buf<int32_t> slab_pe, slab_start; // per-launch item tables of the z-slab solve (chunks of the own planes only)
    DevTable stage_tab[3];
    std::unique_ptr<HPGeom> stage_hp; // geometry cache of fr3d_sor_level
    // level-by-level execution state (fr3d_level_begin / _sweeps / _end / fr3d_flow_finish)
    int run_B = 0, run_level = -1;    // frames of the run, level whose solve is open (-1: none)
    int run_done = -1;                // last level completed
    bool run_flip = false;            // which of uvw_a / uvw_b holds the current flow
    SorParams<float> spf;
    SorParams<double> spd;
    const HPGeom* sp_hp = nullptr;
};

static thread_local std::string g_create_err;

// ---- building blocks -------------------------------------------------------------------------
// One resampling pass.  n[], ss[], ds[]: extents and element strides of the 5-D item space, axis r
// (1..3) is the resampled one.  inner4: the thread walks axis 4 itself (short interleaved axis whose tap
// table look-ups and index arithmetic are then shared); otherwise axis 4 must have been folded away
// (n[4] == 1).
template <class SrcT, class DstT>
static void resize_pass(fr3d_ctx* c, const SrcT* src, const int64_t ss[5], DstT* dst, const int64_t ds[5],
                        const int64_t n[5], int r, const DevTable& t)
{
    FR3D_REQUIRE(r >= 1 && r <= 3, "resize_pass: resampled axis must be 1..3");
    const int64_t inner = n[1] * n[2] * n[3];
    FR3D_REQUIRE(inner > 0 && inner < (1LL << 31), "resize_pass: one batch entry has %lld outputs", (long long)inner);
    // items of a launch are decoded with 32-bit fast division: split the batch axis if needed
    const int64_t per = ((1LL << 31) / inner) < 1 ? 1 : ((1LL << 31) / inner);
    for (int64_t b0 = 0; b0 < n[0]; b0 += per) {
        const int64_t nb = n[0] - b0 < per ? n[0] - b0 : per;
        if (r != 3 && n[3] >= 96) {
            // 4 outputs per thread along axis 3 (lane + 32u inside blocks of 128) share the tap look-ups
            ResizePassK<SrcT, DstT, 4> k;
            const int64_t runs = (n[3] + 127) / 128;
            k.src = src + b0 * ss[0];
            k.dst = dst + b0 * ds[0];
            for (int q = 0; q < 5; ++q) {
                k.ss[q] = ss[q];
                k.ds[q] = ds[q];
                k.fd[q] = FastDiv((uint32_t)(q == 3 ? runs : n[q]));
            }
            k.r = r;
            k.P = t.P;
            k.n4 = (int)n[4];
            k.n3 = (int)n[3];
            k.idx = t.idx.p;
            k.wt = t.wt.p;
            launch(c->dev, k, nb * n[1] * n[2] * runs * 32);
            continue;
        }
        if (r == 3 && n[2] >= 8 && c->dev.resize_x_rows) {
            // X pass: 4 rows per thread share the tap look-ups of their output position
            constexpr int RY = 4;
            ResizeXRowsK<SrcT, DstT, RY> k;
            const int64_t groups = (n[2] + RY - 1) / RY;
            k.src = src + b0 * ss[0];
            k.dst = dst + b0 * ds[0];
            for (int q = 0; q < 5; ++q) {
                k.ss[q] = ss[q];
                k.ds[q] = ds[q];
                k.fd[q] = FastDiv((uint32_t)(q == 2 ? groups : n[q]));
            }
            k.P = t.P;
            k.n4 = (int)n[4];
            k.n2 = (int)n[2];
            k.idx = t.idx.p;
            k.wt = t.wt.p;
            launch(c->dev, k, nb * n[1] * groups * n[3]);
            continue;
        }
        ResizePassK<SrcT, DstT> k;
        k.src = src + b0 * ss[0];
        k.dst = dst + b0 * ds[0];
        for (int q = 0; q < 5; ++q) {
            k.ss[q] = ss[q];
            k.ds[q] = ds[q];
            k.fd[q] = FastDiv((uint32_t)n[q]);
        }
        k.r = r;
        k.P = t.P;
        k.n4 = (int)n[4];
        k.n3 = (int)n[3];
        k.idx = t.idx.p;
        k.wt = t.wt.p;
        launch(c->dev, k, nb * inner);
    }
}

// Separable resize of n0*n1 volumes (D,H,W) -> (od,oh,ow): X, then Y, then Z pass
// (util/resize_util_3D.py:132-145), float32 intermediates.
template <class SrcT, class DstT>
static void resize3(fr3d_ctx* c, const SrcT* src, View sv, int64_t n0, int64_t n1, int D, int H, int W,
                    DstT* dst, View dv, const DevTable tabs[3])
{
    const int ow = tabs[0].out_len, oh = tabs[1].out_len, od = tabs[2].out_len;
    if (n1 == 1) { // a single-entry axis carries no stride of its own
        sv.s1 = sv.s0;
        dv.s1 = dv.s0;
    }
    FR3D_REQUIRE(tabs[0].in_len == W && tabs[1].in_len == H && tabs[2].in_len == D,
                 "resize tables (%d,%d,%d) do not match volume (%d,%d,%d)", tabs[2].in_len, tabs[1].in_len,
                 tabs[0].in_len, D, H, W);
    float* a = c->t1.ensure(c->dev, (size_t)(n0 * n1) * D * H * ow);
    float* b = c->t2.ensure(c->dev, (size_t)(n0 * n1) * D * oh * ow);
    const View av = planar(n1, D, H, ow), bv = planar(n1, D, oh, ow);
    // The item order of a pass is free (ResizePassK takes explicit extents and strides): when the
    // source / destination is channel-interleaved (stride of the n1 axis == 1) the n1 axis is made the
    // fastest one so that the interleaved side is accessed contiguously.
    // Item space (n0, A, B, C | inner): when the source / destination is channel-interleaved (stride of the
    // n1 axis == 1) and short, the thread walks the n1 axis itself (axis 4) so that the interleaved side
    // is accessed contiguously and the tap look-ups are shared; otherwise n0 and n1 are folded into
    // axis 0 (planar views: s0 == n1 * s1).
    const bool src_il = sv.s1 == 1 && n1 > 1 && n1 <= 4;
    const bool dst_il = dv.s1 == 1 && n1 > 1 && n1 <= 4;
    const bool src_fold = sv.s0 == n1 * sv.s1 || n0 == 1, dst_fold = dv.s0 == n1 * dv.s1 || n0 == 1;
    FR3D_REQUIRE((src_il || src_fold) && (dst_il || dst_fold), "resize3: unsupported view");
    if (src_il) {
        const int64_t n[5] = {n0, D, H, ow, n1};
        const int64_t ss[5] = {sv.s0, sv.sz, sv.sy, sv.sx, sv.s1};
        const int64_t ds[5] = {av.s0, av.sz, av.sy, av.sx, av.s1};
        resize_pass<SrcT, float>(c, src, ss, a, ds, n, 3, tabs[0]);
    } else {
        const int64_t n[5] = {n0 * n1, D, H, ow, 1};
        const int64_t ss[5] = {sv.s1, sv.sz, sv.sy, sv.sx, 0};
        const int64_t ds[5] = {av.s1, av.sz, av.sy, av.sx, 0};
        resize_pass<SrcT, float>(c, src, ss, a, ds, n, 3, tabs[0]);
    }
    {
        const int64_t n[5] = {n0 * n1, D, oh, ow, 1};
        const int64_t ss[5] = {av.s1, av.sz, av.sy, av.sx, 0};
        const int64_t ds[5] = {bv.s1, bv.sz, bv.sy, bv.sx, 0};
        resize_pass<float, float>(c, a, ss, b, ds, n, 2, tabs[1]);
    }
    if (dst_il) {
        const int64_t n[5] = {n0, od, oh, ow, n1};
        const int64_t ss[5] = {bv.s0, bv.sz, bv.sy, bv.sx, bv.s1};
        const int64_t ds[5] = {dv.s0, dv.sz, dv.sy, dv.sx, dv.s1};
        resize_pass<float, DstT>(c, b, ss, dst, ds, n, 1, tabs[2]);
    } else {
        const int64_t n[5] = {n0 * n1, od, oh, ow, 1};
        const int64_t ss[5] = {bv.s1, bv.sz, bv.sy, bv.sx, 0};
        const int64_t ds[5] = {dv.s1, dv.sz, dv.sy, dv.sx, 0};
        resize_pass<float, DstT>(c, b, ss, dst, ds, n, 1, tabs[2]);
    }
}

// One in-place prefilter pass over `nlines` lines of N samples through the shared-memory tiled kernel;
// lines too long for a tile fall back to the one-thread-per-line kernel.
template <class Fallback>
static void spline_tile_pass(fr3d_ctx* c, double* coef, int N, int64_t nlines, int64_t per_group, int64_t group_stride,
                             int64_t line_stride, int64_t elem_stride, int64_t first, int line_fast,
                             const Fallback& fallback)
{
    const int L3 = N + 3;
    const int Lv = N + 2 * FR3D_SPLINE_PAD;
    // segments of >= 32 samples (+ 40 warm-up), at most 16 per line; tile = power-of-two number of lines with
    // lines x segments <= 256 threads and <= 72 KB of shared memory (three resident blocks per SM, so the
    // staging copies of one overlap the filter phases of the others).  Measured on B200 (515-sample lines):
    // 8 segments 5.6 ms, 16 segments 4.2 ms, 24 / 32 segments 5.6 ms per step.
    int nseg = Lv / 32;
    nseg = nseg < 1 ? 1 : (nseg > 16 ? 16 : nseg);
    int TL = 32;
    while (TL > 1 && TL * nseg > 256)
        TL /= 2;
    const size_t budget = 72 * 1024;
    while (TL > 1 && (size_t)TL * (L3 + nseg + FR3D_SPLINE_PAD) * sizeof(double) > budget)
        TL /= 2;
    if ((size_t)TL * (L3 + nseg + FR3D_SPLINE_PAD) * sizeof(double) > budget || TL < 4) {
        fallback();
        return;
    }
    SplineTileK k;
    k.coef = coef;
    k.N = N;
    k.TL = TL;
    k.NSEG = nseg;
    k.W = 40;
    k.nlines = nlines;
    k.per_group = per_group;
    k.group_stride = group_stride;
    k.line_stride = line_stride;
    k.elem_stride = elem_stride;
    k.first = first;
    k.line_fast = line_fast;
    // bulk-copy (TMA) staging when a block's lines are one contiguous 16-byte aligned run: unit element stride, lines
    // back to back within and across groups, an even number of 8-byte slots per block start
    k.tma = (!line_fast && elem_stride == 1 && line_stride == L3 &&
             (nlines <= per_group || group_stride == per_group * line_stride) &&
             ((int64_t)TL * L3) % 2 == 0 && first % 2 == 0 && c->dev.spline_tma)
                ? 1
                : 0;
    int threads = (TL * nseg + 31) / 32 * 32;
    launch_tiles(c->dev, k, (nlines + TL - 1) / TL, threads,
                 (size_t)TL * (L3 + nseg + FR3D_SPLINE_PAD) * sizeof(double) + 16);
}

// Cubic B-spline coefficients of B*C volumes (any dtype / strides) -> c->coef.
static void spline_prefilter(fr3d_ctx* c, const void* src, int dt, int64_t sb, int64_t sc, int64_t sz,
                             int64_t sy, int64_t sx, int B, int C, int Z, int Y, int X)
{
    const size_t n = (size_t)B * C * (Z + 3) * (Y + 3) * (X + 3);
    double* coef = c->coef.ensure(c->dev, n);
    // (measured slower than this in-place pass, 2.1 ms per 16-frame step: shared-memory staged 3.6 ms, per-thread
    // local line 2.7 ms)
    SplineZK kz{src, dt, sb, sc, sz, sy, sx, coef, B, C, Z, Y, X};
    launch(c->dev, kz, (int64_t)B * Y * X * C);
    spline_tile_pass(c, coef, Y, (int64_t)B * C * (Z + 3) * X, X, (int64_t)(Y + 3) * (X + 3), 1, X + 3, 1, 1,
                     [&] { launch(c->dev, SplineYK{coef, Z, Y, X}, (int64_t)B * C * (Z + 3) * X); });
    spline_tile_pass(c, coef, X, (int64_t)B * C * (Z + 3) * (Y + 3), (int64_t)1 << 62, 0, X + 3, 1, 0, 0,
                     [&] { launch(c->dev, SplineXK{coef, X}, (int64_t)B * C * (Z + 3) * (Y + 3)); });
}

// Gather launch.  The kernel is latency-bound (ncu: 53 % long-scoreboard stalls at 16 warps per SM), so what pays is
// resident warps and independent work per thread, not fewer bytes (profiles/r01_sor_variants.txt, gather
// experiments): cubic warps of 1 or 2 channels run the channel-interleaved kernel with a rolled z loop at 64
// registers (4 CTAs per SM): 7.6 ms against 10.5 ms for 16 frames of 32x512x512x2; other channel counts run the
// generic kernel capped to 64 registers (8.7 ms).  Blocks cover 32x8x1 outputs (FR3D_OPT_WARP_TILE: A/B aid).
static void launch_gather(fr3d_ctx* c, WarpGatherK g)
{
    // block shape of the gather in outputs (FR3D_OPT_WARP_TILE, an A/B aid; 0 fields keep 32 x 8 x 1)
    int tx = c->dev.warp_tile & 0xff, ty = (c->dev.warp_tile >> 8) & 0xff, tz = (c->dev.warp_tile >> 16) & 0xff;
    if (tx <= 0 || ty <= 0 || tz <= 0 || tx * ty * tz != 256) {
        tx = 32;
        ty = 8;
        tz = 1;
    }
    g.set_tile(tx, ty, tz);
    const int64_t n = g.items();
    if (g.order == 3 && g.C == 2 && c->dev.warp_factored)
        launch_occ<4>(c->dev, WarpGatherLeanK<2, 0, 1>{g}, n);
    else if (g.order == 3 && g.C == 1 && c->dev.warp_factored)
        launch_occ<4>(c->dev, WarpGatherLeanK<1, 0, 1>{g}, n);
    else if (g.order == 3 && g.C == 2)
        launch_occ<4>(c->dev, WarpGatherLeanK<2, 0>{g}, n);
    else if (g.order == 3 && g.C == 1)
        launch_occ<4>(c->dev, WarpGatherLeanK<1, 0>{g}, n);
    else if (g.order == 3)
        launch_occ<4>(c->dev, g, n);
    else
        launch(c->dev, g, n);
}

static void check_dtype(int dt)
{
    FR3D_REQUIRE(dtype_size(dt) != 0, "unsupported dtype code %d", dt);
}

// Frames per warp work item: enough items to balance the busiest wave over the grid.
static int sor_frame_group(const Device& dev, int B)
{
    if (dev.sor_frames_per_item > 0) // FR3D_OPT_SOR_FRAMES_PER_ITEM (tuning aid)
        return dev.sor_frames_per_item < B ? dev.sor_frames_per_item : B;
    return B >= 2 ? 2 : 1;
}

// Level solve in solver storage: assembles J (when f1/f2 are given; otherwise Jpre is used as is), the
// Laplacian term L and zero increments, then runs the wavefront solver.  Result in c->d.
template <class ST>
static SorParams<ST>& sor_params(fr3d_ctx* c);
template <>
SorParams<float>& sor_params<float>(fr3d_ctx* c) { return c->spf; }
template <>
SorParams<double>& sor_params<double>(fr3d_ctx* c) { return c->spd; }

// Prepare a level solve in solver storage: assembles J (when f1/f2 are given; otherwise Jpre is used as
// is), the Laplacian term L and zero increments; the parameters are kept in the context.
template <class ST>
static void sor_prepare_t(fr3d_ctx* c, const HPGeom& hp, int B, int C, const float* f1, const float* f2, int f2f32,
                          const double* Jpre, const double* uvw, const double* whp, double hz, double hy, double hx,
                          const double* alpha, int T, int lag, const double* a_data, int sweep, double a_smooth)
{
    Device& dev = c->dev;
    const int64_t np = hp.npad;
    SorParams<ST>& P = sor_params<ST>(c);
    P.g = hp.view();
    P.C = C;
    P.B = B;
    P.T = T;
    P.lag = lag;
    P.fg = sor_frame_group(c->dev, B);
    P.redblack = sweep == FR3D_SWEEP_REDBLACK;
    // default: a fifth of each wave by ticket for the float64 state (-3.4 % on a B200), round-robin for float32 (tickets
    // cost that kernel registers: +3 %)
    P.sched = c->dev.sor_sched >= 0 ? c->dev.sor_sched : (sizeof(ST) == 8 ? 20 : 0);
    P.t_begin = 0;
    P.t_end = T;
    P.q_begin = 0;
    P.q_end = sor_num_waves(P);
    P.ax = alpha[0] / (hx * hx);
    P.ay = alpha[1] / (hy * hy);
    P.az = alpha[2] / (hz * hz);
    for (int q = 0; q < FR3D_MAX_CHANNELS; ++q)
        P.a_data[q] = q < C ? a_data[q] : 1.0;
    double* J = const_cast<double*>(Jpre);
    if (!J)
        J = c->J.ensure(dev, (size_t)B * C * 10 * np);
    P.J = J;
    P.wgt = whp;
    P.L = (Vec4<ST>*)c->L.ensure(dev, (size_t)B * np * sizeof(Vec4<ST>));
    P.d = (Vec4<ST>*)c->d.ensure(dev, (size_t)B * np * sizeof(Vec4<ST>));
    P.AB = c->AB.ensure(dev, (size_t)B * 9 * np);
    P.a_smooth = a_smooth;
    P.hx = hx;
    P.hy = hy;
    P.hz = hz;
    P.pe4 = hp.pe4.p;
    P.peR = (P.sched & 128) ? hp.refresh_prefix(dev, lag) : nullptr;
    P.U = nullptr;
    P.dold = nullptr;
    P.psi_c = nullptr;
    P.psi_r = nullptr;
    if (a_smooth != 1.0) {
        FR3D_REQUIRE(sweep == FR3D_SWEEP_LEXICOGRAPHIC, "the red-black sweep is only implemented for a_smooth == 1");
        P.U = (Vec4<ST>*)c->U.ensure(dev, (size_t)B * np * sizeof(Vec4<ST>));
        P.dold = (Vec4<ST>*)c->dold.ensure(dev, (size_t)B * np * sizeof(Vec4<ST>));
        P.psi_c = c->psi_c.ensure(dev, (size_t)B * np);
        P.psi_r = c->psi_r.ensure(dev, (size_t)B * 3 * np);
    }
    AssembleK<ST> as;
    as.f1 = f1;
    as.f2 = f2;
    as.uvw = uvw;
    as.J = Jpre ? nullptr : J;
    as.L = const_cast<Vec4<ST>*>(P.L);
    as.U = const_cast<Vec4<ST>*>(P.U);
    as.hp = P.g;
    as.g = MTGeom{hp.p, hp.m, hp.n, hz, hy, hx, f2f32};
    as.B = B;
    as.C = C;
    as.ax = P.ax;
    as.ay = P.ay;
    as.az = P.az;
    // pad slots of the vectors the solver or the state copies may touch: cleared once, the kernel fills the rest
    dev.zero(P.d, (size_t)B * np * sizeof(Vec4<ST>));
    dev.zero(const_cast<Vec4<ST>*>(P.L), (size_t)B * np * sizeof(Vec4<ST>));
    if (P.U) {
        dev.zero(const_cast<Vec4<ST>*>(P.U), (size_t)B * np * sizeof(Vec4<ST>));
        dev.zero(P.dold, (size_t)B * np * sizeof(Vec4<ST>));
    }
    launch_occ2(dev, as, (int64_t)B * hp.p * hp.m * hp.n);
    c->sp_hp = &hp;
}

static void sor_prepare(fr3d_ctx* c, int state_dtype, const HPGeom& hp, int B, int C, const float* f1, const float* f2,
                        int f2f32, const double* Jpre, const double* uvw, const double* whp, double hz, double hy,
                        double hx, const double* alpha, int T, int lag, const double* a_data, int sweep, double a_smooth)
{
    FR3D_REQUIRE(sweep == FR3D_SWEEP_LEXICOGRAPHIC || sweep == FR3D_SWEEP_REDBLACK, "unknown sweep order %d", sweep);
    if (state_dtype == FR3D_F64)
        sor_prepare_t<double>(c, hp, B, C, f1, f2, f2f32, Jpre, uvw, whp, hz, hy, hx, alpha, T, lag, a_data, sweep, a_smooth);
    else
        sor_prepare_t<float>(c, hp, B, C, f1, f2, f2f32, Jpre, uvw, whp, hz, hy, hx, alpha, T, lag, a_data, sweep, a_smooth);
}

// Run sweeps [t0,t1) x waves [q0,q1) of the prepared solve (negative bounds: the whole range).
template <class ST>
static void sor_launch_t(fr3d_ctx* c, int t0, int t1, int q0, int q1, int k0 = 0, int k1 = 0)
{
    SorParams<ST> P = sor_params<ST>(c);
    if (k1 > 0) {
        FR3D_REQUIRE(!P.redblack && P.a_smooth == 1.0, "z-slab sweeps need the lexicographic sweep with a_smooth == 1");
        FR3D_REQUIRE(k0 >= 0 && k0 < k1 && k1 <= P.g.p, "bad plane range [%d, %d) of %d", k0, k1, P.g.p);
    }
    const int nw = sor_num_waves(P);
    const bool partial = !(t0 <= 0 && (t1 < 0 || t1 >= P.T) && q0 <= 0 && (q1 < 0 || q1 >= nw));
    if (partial) {
        FR3D_REQUIRE(!P.redblack && P.a_smooth == 1.0,
                     "partial sweep ranges need the lexicographic sweep with a_smooth == 1");
        FR3D_REQUIRE(t0 >= 0 && t1 <= P.T && t0 < t1 && q0 >= 0 && q1 <= nw && q0 <= q1, "bad sweep / wave range");
        FR3D_REQUIRE(t0 % P.lag == 0, "a partial solve must start at a multiple of update_lag (psi refresh)");
        P.t_begin = t0;
        P.t_end = t1;
        P.q_begin = q0;
        P.q_end = q1;
        if (q0 == q1)
            return;
    }
    c->dev.sor_k0 = k0;
    c->dev.sor_k1 = k1 > 0 ? k1 : 0;
    try {
        sor_run(c->dev, P, c->bar.ensure(c->dev, FR3D_SOR_BAR_WORDS), c->sp_hp->pe_host.data());
    } catch (...) {
        c->dev.sor_k0 = c->dev.sor_k1 = 0;
        throw;
    }
    c->dev.sor_k0 = c->dev.sor_k1 = 0;
}

static void sor_launch(fr3d_ctx* c, int state_dtype, int t0 = -1, int t1 = -1, int q0 = -1, int q1 = -1, int k0 = 0,
                       int k1 = 0)
{
    FR3D_REQUIRE(c->sp_hp != nullptr, "no level solve has been prepared");
    if (state_dtype == FR3D_F64)
        sor_launch_t<double>(c, t0, t1, q0, q1, k0, k1);
    else
        sor_launch_t<float>(c, t0, t1, q0, q1, k0, k1);
}

#ifndef FR3D_EMU
template <class ST>
static void sor_launch_p2p_t(fr3d_ctx* c, int k0, int k1, void* lo_d, void* hi_d, void* lo_flags, void* hi_flags,
                             int64_t flag_base)
{
    SorParams<ST> P = sor_params<ST>(c);
    FR3D_REQUIRE(!P.redblack && P.a_smooth == 1.0, "z-slab sweeps need the lexicographic sweep with a_smooth == 1");
    FR3D_REQUIRE(k0 >= 0 && k0 < k1 && k1 <= P.g.p, "bad plane range [%d, %d) of %d", k0, k1, P.g.p);
    FR3D_REQUIRE((lo_d == nullptr) == (lo_flags == nullptr) && (hi_d == nullptr) == (hi_flags == nullptr),
                 "a neighbour needs both its increment array and its flag words");
    FR3D_REQUIRE(c->p2p_flags.p != nullptr, "export the flag words first (fr3d_ipc_export which = 1)");
    SorPeers<ST> pr;
    pr.lo_d = (Vec4<ST>*)lo_d;
    pr.hi_d = (Vec4<ST>*)hi_d;
    pr.lo_flag = lo_flags ? (unsigned*)lo_flags + 1 : nullptr; // the lower neighbour's "from my upper neighbour" word
    pr.hi_flag = hi_flags ? (unsigned*)hi_flags + 0 : nullptr; // the upper neighbour's "from my lower neighbour" word
    pr.my_flags = c->p2p_flags.p;
    pr.base = (unsigned)flag_base;
    // Item tables restricted to the slab: in hyperplane-major storage the voxels of hyperplane s are ordered by plane,
    // so the planes [k0, k1) are ONE contiguous slot range of every hyperplane; only the 32-slot chunks that touch it
    // become work items (the kernel still masks the few foreign voxels of the two end chunks).  Without this every
    // rank would walk all chunks of every wave and skip half of them -- measured: no speed-up at all.
    const HPGeom& hp = *c->sp_hp;
    const int S = hp.S, p = hp.p, m = hp.m, n = hp.n;
    std::vector<int32_t> spe(S), sst((size_t)S + 1);
    int64_t at = 0;
    for (int s = 0; s < S; ++s) {
        int64_t run = 0, lo = 0, hi = 0;
        for (int k = 0; k < p; ++k) {
            if (k == k0)
                lo = run;
            int jlo = s - k - (n - 1);
            jlo = jlo < 0 ? 0 : jlo;
            int jhi = s - k;
            jhi = jhi > m - 1 ? m - 1 : jhi;
            if (jhi >= jlo)
                run += jhi - jlo + 1;
            if (k == k1 - 1)
                hi = run;
        }
        const int64_t c0 = lo / 32, c1 = hi > lo ? (hi + 31) / 32 : c0;
        sst[s] = (int32_t)(at + 32 * c0);
        spe[s] = (int32_t)(c1 - c0) + (s >= 2 ? spe[s - 2] : 0);
        at += (run + 31) / 32 * 32;
    }
    sst[S] = (int32_t)at;
    c->slab_pe.upload(c->dev, spe.data(), spe.size());
    c->slab_start.upload(c->dev, sst.data(), sst.size());
    FR3D_CUDA(cudaStreamSynchronize(c->dev.stream)); // the host vectors go out of scope
    P.g.pe = c->slab_pe.p;
    P.g.start = c->slab_start.p;
    sor_run_p2p_any(c->dev, P, c->bar.ensure(c->dev, FR3D_SOR_BAR_WORDS), spe.data(), k0, k1, pr);
}
#endif

static void run_sor(fr3d_ctx* c, int state_dtype, const HPGeom& hp, int B, int C, const float* f1, const float* f2,
                    int f2f32, const double* Jpre, const double* uvw, const double* whp, double hz, double hy,
                    double hx, const double* alpha, int T, int lag, const double* a_data, int sweep, double a_smooth)
{
    sor_prepare(c, state_dtype, hp, B, C, f1, f2, f2f32, Jpre, uvw, whp, hz, hy, hx, alpha, T, lag, a_data, sweep, a_smooth);
    sor_launch(c, state_dtype);
}

// increments in solver storage -> natural planar float64 (B, 3, N)
static void sor_result(fr3d_ctx* c, int state_dtype, const HPGeom& hp, int B, double* dnat)
{
    const int64_t n = (int64_t)B * 3 * hp.p * hp.m * hp.n;
    if (state_dtype == FR3D_F64)
        launch(c->dev, FromHPK<double>{(const Vec4<double>*)c->d.p, dnat, hp.view()}, n);
    else
        launch(c->dev, FromHPK<float>{(const Vec4<float>*)c->d.p, dnat, hp.view()}, n);
}

// ---- C ABI -------------------------------------------------------------------------------------
#define FR3D_API_BEGIN(ctx_)                  \
    fr3d_ctx* _c = (ctx_);                    \
    if (!_c)                                  \
        return FR3D_ERR_ARG;                  \
    try {
#define FR3D_API_END()                        \
    }                                         \
    catch (const Error& e)                    \
    {                                         \
        _c->err = e.msg;                      \
        return e.code;                        \
    }                                         \
    catch (const std::exception& e)           \
    {                                         \
        _c->err = e.what();                   \
        return FR3D_ERR_NOMEM;                \
    }                                         \
    return FR3D_OK;

extern "C" {

int fr3d_abi_version(void) { return FR3D_ABI_VERSION; }

const char* fr3d_last_error(const fr3d_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int64_t fr3d_launch_count(const fr3d_ctx* ctx) { return ctx ? ctx->dev.launches : 0; }
int64_t fr3d_device_bytes(const fr3d_ctx* ctx) { return ctx ? ctx->dev.bytes : 0; }

int fr3d_create(fr3d_ctx** out, int device, const fr3d_plan* plan, void* stream)
{
    if (!out)
        return FR3D_ERR_ARG;
    *out = nullptr;
    fr3d_ctx* c = nullptr;
    try {
        c = new fr3d_ctx();
#ifndef FR3D_EMU
        FR3D_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        FR3D_CUDA(cudaGetDeviceProperties(&prop, device));
        FR3D_REQUIRE(prop.major >= 10, "libfr3d is built for sm_100a (B200); device %d is sm_%d%d", device,
                     prop.major, prop.minor);
        FR3D_REQUIRE(prop.cooperativeLaunch, "device %d lacks cooperative launch", device);
        c->dev.sm_count = prop.multiProcessorCount;
        c->dev.stream = (cudaStream_t)stream;
#else
        (void)device;
        (void)stream;
#endif
        if (plan) {
            FR3D_REQUIRE(plan->abi_version == FR3D_ABI_VERSION, "plan ABI %d != library ABI %d",
                         plan->abi_version, FR3D_ABI_VERSION);
            FR3D_REQUIRE(plan->Z > 0 && plan->Y > 0 && plan->X > 0, "bad volume size");
            FR3D_REQUIRE(plan->C >= 1 && plan->C <= FR3D_MAX_CHANNELS, "C must be 1..%d", FR3D_MAX_CHANNELS);
            FR3D_REQUIRE(plan->max_batch >= 1, "max_batch must be >= 1");
            FR3D_REQUIRE(plan->n_levels >= 1 && plan->n_levels <= FR3D_MAX_LEVELS && plan->levels, "bad level list");
            FR3D_REQUIRE(plan->iterations >= 1 && plan->update_lag >= 1, "iterations/update_lag must be >= 1");
            FR3D_REQUIRE(plan->a_smooth > 0.0, "a_smooth must be positive");
            FR3D_REQUIRE(plan->a_smooth == 1.0 || plan->sweep == FR3D_SWEEP_LEXICOGRAPHIC,
                         "the red-black sweep is only implemented for a_smooth == 1");
            FR3D_REQUIRE(plan->interp == 3 || plan->interp == 1, "interp must be 3 (cubic) or 1 (linear)");
            FR3D_REQUIRE(plan->sweep == FR3D_SWEEP_LEXICOGRAPHIC || plan->sweep == FR3D_SWEEP_REDBLACK,
                         "unknown sweep order %d", plan->sweep);
            c->Z = plan->Z;
            c->Y = plan->Y;
            c->X = plan->X;
            c->C = plan->C;
            c->max_batch = plan->max_batch;
            c->iterations = plan->iterations;
            c->update_lag = plan->update_lag;
            c->sweep = plan->sweep;
            c->interp = plan->interp;
            c->a_smooth = plan->a_smooth;
            FR3D_REQUIRE(plan->state_dtype == FR3D_F32 || plan->state_dtype == FR3D_F64,
                         "state_dtype must be FR3D_F32 or FR3D_F64");
            c->state_dtype = plan->state_dtype;
            for (int q = 0; q < FR3D_MAX_CHANNELS; ++q)
                c->a_data[q] = plan->a_data[q];
            for (int li = 0; li < plan->n_levels; ++li) {
                const fr3d_level& s = plan->levels[li];
                c->levels.emplace_back(new LevelDev());
                LevelDev& L = *c->levels[li];
                L.pz = s.size[0];
                L.py = s.size[1];
                L.px = s.size[2];
                FR3D_REQUIRE(L.pz > 0 && L.py > 0 && L.px > 0, "level %d has an empty grid", li);
                L.N = (int64_t)L.pz * L.py * L.px;
                L.hp.build(c->dev, L.pz, L.py, L.px);
                L.hz = s.h[0];
                L.hy = s.h[1];
                L.hx = s.h[2];
                for (int q = 0; q < 3; ++q)
                    L.alpha[q] = s.alpha[q];
                L.median = s.median != 0;
                for (int q = 0; q < 3; ++q) {
                    L.full[q].upload(c->dev, s.from_full[q]);
                    FR3D_REQUIRE(L.full[q].P > 0, "level %d: from_full table %d missing", li, q);
                    if (li > 0) {
                        L.prev[q].upload(c->dev, s.from_prev[q]);
                        FR3D_REQUIRE(L.prev[q].P > 0, "level %d: from_prev table %d missing", li, q);
                    }
                }
                FR3D_REQUIRE(L.full[0].out_len == L.px && L.full[1].out_len == L.py && L.full[2].out_len == L.pz,
                             "level %d: from_full tables do not produce the level size", li);
            }
            for (int q = 0; q < 3; ++q)
                c->to_full[q].upload(c->dev, plan->to_full[q]);
            for (int ch = 0; ch < c->C; ++ch)
                for (int a = 0; a < 3; ++a) {
                    const int r = plan->gauss_radius[ch][a];
                    FR3D_REQUIRE(r >= 0 && (r == 0 || plan->gauss_w[ch][a]), "bad Gaussian kernel (c=%d, axis=%d)", ch, a);
                    c->pre.r[ch][a] = r;
                    std::vector<double> one(1, 1.0);
                    const double* hw = plan->gauss_w[ch][a] ? plan->gauss_w[ch][a] : one.data();
                    c->pre.w[ch][a] = c->pre_w[ch][a].upload(c->dev, hw, (size_t)r + 1);
                }
            for (int ch = 0; ch < c->C; ++ch) {
                const int r = plan->gauss_radius_t[ch];
                FR3D_REQUIRE(r >= 0 && (r == 0 || plan->gauss_w_t[ch]), "bad temporal Gaussian kernel (c=%d)", ch);
                c->pre.rt[ch] = r;
                std::vector<double> one(1, 1.0);
                const double* hw = plan->gauss_w_t[ch] ? plan->gauss_w_t[ch] : one.data();
                c->pre.wt[ch] = c->pre_wt[ch].upload(c->dev, hw, (size_t)r + 1);
            }
            c->has_plan = true;
        }
        c->dev.sync();
    } catch (const Error& e) {
        g_create_err = e.msg;
        delete c;
        return e.code;
    } catch (const std::exception& e) {
        g_create_err = e.what();
        delete c;
        return FR3D_ERR_NOMEM;
    }
    *out = c;
    return FR3D_OK;
}

void fr3d_destroy(fr3d_ctx* ctx)
{
    if (!ctx)
        return;
#ifndef FR3D_EMU
    cudaStreamSynchronize(ctx->dev.stream);
#endif
    delete ctx;
}

int fr3d_synchronize(fr3d_ctx* ctx)
{
    FR3D_API_BEGIN(ctx)
    _c->dev.sync();
    FR3D_API_END()
}

int fr3d_set_option(fr3d_ctx* ctx, int option, int64_t value)
{
    FR3D_API_BEGIN(ctx)
    switch (option) {
    case FR3D_OPT_SOR_CTAS_PER_SM:
        FR3D_REQUIRE(value >= 0 && value <= 32, "FR3D_OPT_SOR_CTAS_PER_SM: %lld", (long long)value);
        _c->dev.sor_ctas_per_sm = (int)value;
        break;
    case FR3D_OPT_WARP_FACTORED:
        FR3D_REQUIRE(value == 0 || value == 1, "FR3D_OPT_WARP_FACTORED: %lld", (long long)value);
        _c->dev.warp_factored = (int)value;
        break;
    case FR3D_OPT_CC_BLOCK_SCANS:
        FR3D_REQUIRE(value == 0 || value == 1, "FR3D_OPT_CC_BLOCK_SCANS: %lld", (long long)value);
        _c->dev.cc_block_scans = (int)value;
        break;
    case FR3D_OPT_SOR_KERNEL:
        FR3D_REQUIRE(value >= 0 && value <= 2, "FR3D_OPT_SOR_KERNEL: %lld", (long long)value);
        _c->dev.sor_kernel = (int)value;
        break;
    case FR3D_OPT_RESIZE_X_ROWS:
        FR3D_REQUIRE(value == 0 || value == 1, "FR3D_OPT_RESIZE_X_ROWS: %lld", (long long)value);
        _c->dev.resize_x_rows = (int)value;
        break;
    case FR3D_OPT_SPLINE_TMA:
        FR3D_REQUIRE(value == 0 || value == 1, "FR3D_OPT_SPLINE_TMA: %lld", (long long)value);
        _c->dev.spline_tma = (int)value;
        break;
    case FR3D_OPT_SOR_TILE: {
        const int tb = (int)(value & 0xff), tk = (int)((value >> 8) & 0xff), tj = (int)((value >> 16) & 0xff),
                  ti = (int)((value >> 24) & 0xff);
        FR3D_REQUIRE(value >= 0 && (value >> 32) == 0 && tb <= 32 && (tk == 0 || tk >= 2) && (tj == 0 || tj >= 2) &&
                         (ti == 0 || ti >= 2),
                     "FR3D_OPT_SOR_TILE: %lld", (long long)value);
        _c->dev.sor_tile_sweeps = tb;
        _c->dev.sor_tile_k = tk;
        _c->dev.sor_tile_j = tj;
        _c->dev.sor_tile_i = ti;
        break;
    }
    case FR3D_OPT_SOR_FRAMES_PER_ITEM:
        FR3D_REQUIRE(value >= 0 && value <= 64, "FR3D_OPT_SOR_FRAMES_PER_ITEM: %lld", (long long)value);
        _c->dev.sor_frames_per_item = (int)value;
        break;
    case FR3D_OPT_WARP_TILE: {
        const int tx = (int)(value & 0xff), ty = (int)((value >> 8) & 0xff), tz = (int)((value >> 16) & 0xff);
        FR3D_REQUIRE(value == 0 || ((value >> 24) == 0 && tx * ty * tz == 256), "FR3D_OPT_WARP_TILE: %lld (tx*ty*tz must be 256)",
                     (long long)value);
        _c->dev.warp_tile = (int)value;
        break;
    }
    case FR3D_OPT_SOR_SCHED:
        FR3D_REQUIRE(value >= -1 && (value < 0 || ((value & 127) <= 100 && (value >> 11) == 0)), "FR3D_OPT_SOR_SCHED: %lld",
                     (long long)value);
        _c->dev.sor_sched = (int)value;
        break;
    case FR3D_OPT_SOR_STAGES:
        FR3D_REQUIRE(value >= 0 && value <= 16 && value != 1, "FR3D_OPT_SOR_STAGES: %lld", (long long)value);
        _c->dev.sor_stages = (int)value;
        break;
    default: FR3D_THROW(FR3D_ERR_ARG, "unknown option %d", option);
    }
    FR3D_API_END()
}

int fr3d_preprocess(fr3d_ctx* ctx, const void* raw, int dtype, int B, const double* lo, const double* den,
                    int temporal_on, float* out, double* out64)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->has_plan, "fr3d_preprocess needs a context created with a plan");
    FR3D_REQUIRE(raw && out && lo && den && B >= 1, "null argument");
    check_dtype(dtype);
    const int Z = _c->Z, Y = _c->Y, X = _c->X, C = _c->C;
    PreGauss g = _c->pre;
    for (int ch = 0; ch < C; ++ch) {
        g.lo[ch] = lo[ch];
        g.den[ch] = den[ch];
        FR3D_REQUIRE(den[ch] != 0.0, "normalisation denominator is zero");
    }
    const size_t n = (size_t)B * C * Z * Y * X;
    bool temporal = false;
    for (int ch = 0; ch < C; ++ch)
        temporal = temporal || (temporal_on && g.rt[ch] > 0);
    if (temporal) {
        // 4-D filter of the reference: the T axis first, on the normalised values; the spatial passes
        // then read that float64 batch (lo = 0, den = 1 leaves it unchanged)
        double* tb = _c->gt.ensure(_c->dev, n);
        launch(_c->dev, PreTK{raw, dtype, tb, B, (int64_t)Z * Y * X * C, C, g}, (int64_t)n);
        raw = tb;
        dtype = FR3D_F64;
        for (int ch = 0; ch < C; ++ch) {
            g.lo[ch] = 0.0;
            g.den[ch] = 1.0;
        }
    }
    double* a = _c->g1.ensure(_c->dev, n);
    int rz = g.r[0][0];
    for (int ch = 1; ch < C; ++ch)
        if (g.r[ch][0] != rz)
            rz = -1;
    const int64_t ncol = (int64_t)B * Y * X * C;
    switch (rz) {
    case 1: launch(_c->dev, PreZWinK<1>{raw, dtype, a, B, Z, Y, X, C, g}, ncol); break;
    case 2: launch(_c->dev, PreZWinK<2>{raw, dtype, a, B, Z, Y, X, C, g}, ncol); break;
    case 3: launch(_c->dev, PreZWinK<3>{raw, dtype, a, B, Z, Y, X, C, g}, ncol); break;
    case 4: launch(_c->dev, PreZWinK<4>{raw, dtype, a, B, Z, Y, X, C, g}, ncol); break;
    case 5: launch(_c->dev, PreZWinK<5>{raw, dtype, a, B, Z, Y, X, C, g}, ncol); break;
    case 6: launch(_c->dev, PreZWinK<6>{raw, dtype, a, B, Z, Y, X, C, g}, ncol); break;
    default: launch(_c->dev, PreZK{raw, dtype, a, B, Z, Y, X, C, g}, (int64_t)n); break; // mixed / large / zero radii
    }
    int RY = 0, RX = 0;
    for (int ch = 0; ch < C; ++ch) {
        RY = g.r[ch][1] > RY ? g.r[ch][1] : RY;
        RX = g.r[ch][2] > RX ? g.r[ch][2] : RX;
    }
    bool uniform = RY == RX && RY >= 1 && RY <= 6;
    for (int ch = 0; ch < C; ++ch)
        uniform = uniform && g.r[ch][1] == RY && g.r[ch][2] == RY;
    if (uniform) {
        const int ty = (Y + 15) / 16, tx = (X + 63) / 64;
        const int64_t nblk = (int64_t)B * Z * ty * tx;
#define FR3D_PRE_WIN(R_)                                                                                   \
    case R_:                                                                                               \
        launch_tiles(_c->dev, PreYXWinK<R_>{a, out, out64, B, Z, Y, X, C, g, ty, tx}, nblk, 256,            \
                     PreYXWinK<R_>::smem_bytes(C));                                                        \
        break;
        switch (RY) {
            FR3D_PRE_WIN(1)
            FR3D_PRE_WIN(2)
            FR3D_PRE_WIN(3)
            FR3D_PRE_WIN(4)
            FR3D_PRE_WIN(5)
            FR3D_PRE_WIN(6)
        }
#undef FR3D_PRE_WIN
    } else if (RY <= 16 && RX <= 16) {
        PreYXTileK k;
        k.in = a;
        k.out = out;
        k.out64 = out64;
        k.B = B;
        k.Z = Z;
        k.Y = Y;
        k.X = X;
        k.C = C;
        k.RY = RY;
        k.RX = RX;
        k.g = g;
        k.tiles_y = (Y + PreYXTileK::TY - 1) / PreYXTileK::TY;
        k.tiles_x = (X + PreYXTileK::TX - 1) / PreYXTileK::TX;
        const int IW = PreYXTileK::TX + 2 * RX, IH = PreYXTileK::TY + 2 * RY;
        const size_t smem = (size_t)(IH * IW + PreYXTileK::TY * IW + PreYXTileK::TY * PreYXTileK::TX * C) * sizeof(double);
        launch_tiles(_c->dev, k, (int64_t)B * Z * k.tiles_y * k.tiles_x, 256, smem);
    } else {
        double* b = _c->g2.ensure(_c->dev, n);
        launch(_c->dev, PreYK{a, b, C, Z, Y, X, g}, (int64_t)n);
        launch(_c->dev, PreXK{b, out, out64, B, Z, Y, X, C, g}, (int64_t)n);
    }
    FR3D_API_END()
}

int fr3d_set_reference(fr3d_ctx* ctx, const float* ref_proc, const float* weight, const double* weight_const)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->has_plan, "fr3d_set_reference needs a context created with a plan");
    FR3D_REQUIRE(ref_proc && (weight || weight_const), "null argument");
    const int Z = _c->Z, Y = _c->Y, X = _c->X, C = _c->C;
    const View cl{0, 1, (int64_t)Y * X * C, (int64_t)X * C, C}; // channels-last (Z,Y,X,C), o1 = channel
    const float* wsrc = weight;
    if (!wsrc) {
        FillK<float> f;
        f.dst = _c->fscr.ensure(_c->dev, (size_t)Z * Y * X * C);
        f.C = C;
        f.inner = 0;
        for (int q = 0; q < FR3D_MAX_CHANNELS; ++q)
            f.v[q] = q < C ? weight_const[q] : 0.0;
        launch(_c->dev, f, (int64_t)Z * Y * X * C);
        wsrc = f.dst;
    }
    for (auto& Lp : _c->levels) {
        LevelDev& L = *Lp;
        float* f1 = L.f1.ensure(_c->dev, (size_t)C * L.N);
        resize3<float, float>(_c, ref_proc, cl, 1, C, Z, Y, X, f1, planar(C, L.pz, L.py, L.px), L.full);
        // weights: resized like an image (core/optical_flow_3d.py:475), widened to float64, in solver storage
        double* wn = _c->wnat.ensure(_c->dev, (size_t)C * L.N);
        resize3<float, double>(_c, wsrc, cl, 1, C, Z, Y, X, wn, planar(C, L.pz, L.py, L.px), L.full);
        double* ws = L.whp.ensure(_c->dev, (size_t)C * L.hp.npad);
        launch(_c->dev, ToHPK<double>{wn, ws, L.hp.view()}, (int64_t)C * L.hp.npad);
    }
    _c->ref_set = true;
    FR3D_API_END()
}

// ---- coarse-to-fine driver, one level at a time (core/optical_flow_3d.py:403-541) -----------------
static void run_buffers(fr3d_ctx* c, int B, double*& ucur, double*& uprev)
{
    int64_t Nmax = 0;
    for (const auto& L : c->levels)
        Nmax = L->N > Nmax ? L->N : Nmax;
    double* a = c->uvw_a.ensure(c->dev, (size_t)B * 3 * Nmax);
    double* b = c->uvw_b.ensure(c->dev, (size_t)B * 3 * Nmax);
    ucur = c->run_flip ? b : a;
    uprev = c->run_flip ? a : b;
}

// Everything of level li that precedes its solve: moving image of the level, flow from the coarser
// level, cubic warp, system assembly.
static void level_begin(fr3d_ctx* c, int li, const float* moving, const float* uvw_init, int B)
{
    FR3D_REQUIRE(c->has_plan, "needs a context created with a plan");
    if (!c->ref_set)
        FR3D_THROW(FR3D_ERR_STATE, "fr3d_set_reference has not been called");
    FR3D_REQUIRE(moving, "null argument");
    FR3D_REQUIRE(B >= 1 && B <= c->max_batch, "B=%d outside 1..max_batch=%d", B, c->max_batch);
    FR3D_REQUIRE(li >= 0 && li < (int)c->levels.size(), "level %d outside 0..%d", li, (int)c->levels.size() - 1);
    if (li == 0) {
        c->run_B = B;
        c->run_done = -1;
        c->run_flip = false;
    } else if (c->run_done != li - 1 || c->run_B != B || c->run_level != -1) {
        FR3D_THROW(FR3D_ERR_STATE, "levels must be processed in order, coarse to fine, one at a time");
    }
    Device& dev = c->dev;
    const int Z = c->Z, Y = c->Y, X = c->X, C = c->C;
    const int64_t NF = (int64_t)Z * Y * X;
    int64_t Nmax = 0;
    for (const auto& L : c->levels)
        Nmax = L->N > Nmax ? L->N : Nmax;
    float* f2 = c->f2.ensure(dev, (size_t)B * C * Nmax);
    float* tmp = c->tmp.ensure(dev, (size_t)B * C * Nmax);
    if (li > 0)
        c->run_flip = !c->run_flip;
    double *ucur, *uprev;
    run_buffers(c, B, ucur, uprev);
    const View mov_cl{NF * C, 1, (int64_t)Y * X * C, (int64_t)X * C, C};
    LevelDev& L = *c->levels[li];
    const int p = L.pz, m = L.py, n = L.px;
    const int64_t N = L.N;
    // moving image at this level: always resampled from full resolution (:409-410)
    resize3<float, float>(c, moving, mov_cl, B, C, Z, Y, X, f2, planar(C, p, m, n), L.full);
    const float* warped = f2;
    int f2f32 = 0;
    if (li == 0) {
        if (uvw_init) {
            // (:417-420) the initial field is shared by the batch: resample once, broadcast
            float* u0 = c->fscr.ensure(dev, (size_t)3 * N);
            const View uv{0, 1, (int64_t)Y * X * 3, (int64_t)X * 3, 3};
            resize3<float, float>(c, uvw_init, uv, 1, 3, Z, Y, X, u0, planar(3, p, m, n), L.full);
            launch(dev, BroadcastF32toF64K{u0, ucur, 3 * N}, (int64_t)B * 3 * N);
        } else {
            dev.zero(ucur, (size_t)B * 3 * N * sizeof(double));
        }
    } else {
        // (:424-434) upsample the flow from the coarser level, then warp the moving image
        const LevelDev& Lp = *c->levels[li - 1];
        resize3<double, double>(c, uprev, planar(3, Lp.pz, Lp.py, Lp.px), B, 3, Lp.pz, Lp.py, Lp.px, ucur,
                                planar(3, p, m, n), L.prev);
        spline_prefilter(c, f2, FR3D_F32, C * N, N, (int64_t)m * n, n, 1, B, C, p, m, n);
        WarpGatherK g;
        g.order = 3;
        g.coef = c->coef.p;
        g.src = nullptr;
        g.sdt = FR3D_F32;
        g.sb = g.sc = g.sz = g.sy = g.sx = 0;
        g.disp64 = ucur;
        g.disp32 = nullptr;
        g.hx = L.hx;
        g.hy = L.hy;
        g.hz = L.hz;
        g.ref = L.f1.p;
        g.rdt = FR3D_F32;
        g.rc = N;
        g.rz = (int64_t)m * n;
        g.ry = n;
        g.rx = 1;
        g.out = tmp;
        g.ob = C * N;
        g.oc = N;
        g.oz = (int64_t)m * n;
        g.oy = n;
        g.ox = 1;
        g.B = B;
        g.C = C;
        g.Z = p;
        g.Y = m;
        g.X = n;
        launch_gather(c, g);
        warped = tmp;
        f2f32 = 1; // numpy keeps the float32 warp output in float32 through its derivatives
    }
    sor_prepare(c, c->state_dtype, L.hp, B, C, L.f1.p, warped, f2f32, nullptr, ucur, L.whp.p, L.hz, L.hy, L.hx,
                L.alpha, c->iterations, c->update_lag, c->a_data, c->sweep, c->a_smooth);
    c->run_level = li;
}

// Increments of the open level -> natural layout, 5^3 median, accumulation into the flow (:517-529).
// [k0, k0+kn): z range of the flow this call updates (the other planes are left to other ranks; kn < 0 = all).
static void level_end(fr3d_ctx* c, int li, int k0 = 0, int kn = -1)
{
    FR3D_REQUIRE(c->run_level == li, "level %d is not open", li);
    if (kn < 0) {
        k0 = 0;
        kn = c->levels[li]->pz;
    }
    FR3D_REQUIRE(k0 >= 0 && kn >= 0 && k0 + kn <= c->levels[li]->pz, "bad z range");
    Device& dev = c->dev;
    const int B = c->run_B;
    LevelDev& L = *c->levels[li];
    const int p = L.pz, m = L.py, n = L.px;
    int64_t Nmax = 0;
    for (const auto& Lq : c->levels)
        Nmax = Lq->N > Nmax ? Lq->N : Nmax;
    double* dnat = c->dnat.ensure(dev, (size_t)B * 3 * Nmax);
    double *ucur, *uprev;
    run_buffers(c, B, ucur, uprev);
    sor_result(c, c->state_dtype, L.hp, B, dnat);
    if (L.median)
        launch_occ2(dev, Median5PairK{dnat, ucur, ucur, p, m, n, (n + 1) / 2, k0, kn}, (int64_t)B * 3 * kn * m * ((n + 1) / 2));
    else if (kn == p)
        launch(dev, AddK{ucur, dnat, ucur}, (int64_t)B * 3 * L.N);
    else
        for (int v = 0; v < B * 3; ++v) { // unfiltered level: add the slab of every field
            const int64_t o = (int64_t)v * L.N + (int64_t)k0 * m * n;
            launch(dev, AddK{ucur + o, dnat + o, ucur + o}, (int64_t)kn * m * n);
        }
    c->run_level = -1;
    c->run_done = li;
}

// (:530-541) flow of the finest solved level -> full resolution, interleaved (B,Z,Y,X,3).
static void flow_finish(fr3d_ctx* c, void* flow_out, int out_dtype)
{
    FR3D_REQUIRE(flow_out, "null argument");
    FR3D_REQUIRE(out_dtype == FR3D_F32 || out_dtype == FR3D_F64, "flow_out dtype must be F32 or F64");
    FR3D_REQUIRE(c->run_done == (int)c->levels.size() - 1 && c->run_level == -1, "the finest level has not been solved");
    Device& dev = c->dev;
    const int B = c->run_B;
    const int Z = c->Z, Y = c->Y, X = c->X;
    const int64_t NF = (int64_t)Z * Y * X;
    double *ucur, *uprev;
    run_buffers(c, B, ucur, uprev);
    const LevelDev& Lf = *c->levels.back();
    if (c->to_full[0].P > 0) {
        // resample the flow to full resolution; values are not rescaled
        const View fo{NF * 3, 1, (int64_t)Y * X * 3, (int64_t)X * 3, 3};
        if (out_dtype == FR3D_F32)
            resize3<double, float>(c, ucur, planar(3, Lf.pz, Lf.py, Lf.px), B, 3, Lf.pz, Lf.py, Lf.px,
                                   (float*)flow_out, fo, c->to_full);
        else
            resize3<double, double>(c, ucur, planar(3, Lf.pz, Lf.py, Lf.px), B, 3, Lf.pz, Lf.py, Lf.px,
                                    (double*)flow_out, fo, c->to_full);
    } else {
        FR3D_REQUIRE(Lf.pz == Z && Lf.py == Y && Lf.px == X, "finest level is not full resolution but to_full is empty");
        if (out_dtype == FR3D_F32)
            launch(dev, InterleaveFlowK<float>{ucur, (float*)flow_out, NF}, (int64_t)B * NF * 3);
        else
            launch(dev, InterleaveFlowK<double>{ucur, (double*)flow_out, NF}, (int64_t)B * NF * 3);
    }
}

int fr3d_get_displacement(fr3d_ctx* ctx, const float* moving, const float* uvw_init, int B, void* flow_out,
                          int out_dtype)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(moving && flow_out, "null argument");
    FR3D_REQUIRE(out_dtype == FR3D_F32 || out_dtype == FR3D_F64, "flow_out dtype must be F32 or F64");
    _c->run_level = -1;
    for (int li = 0; li < (int)_c->levels.size(); ++li) {
        level_begin(_c, li, moving, uvw_init, B);
        sor_launch(_c, _c->state_dtype);
        level_end(_c, li);
    }
    flow_finish(_c, flow_out, out_dtype);
    FR3D_API_END()
}

int fr3d_level_count(const fr3d_ctx* ctx) { return ctx ? (int)ctx->levels.size() : 0; }

int fr3d_level_info(const fr3d_ctx* ctx, int level, int32_t size[3], int32_t* n_hyperplanes, int64_t* n_slots,
                    int32_t* hyperplane_start)
{
    if (!ctx || level < 0 || level >= (int)ctx->levels.size())
        return FR3D_ERR_ARG;
    const LevelDev& L = *ctx->levels[level];
    if (size) {
        size[0] = L.pz;
        size[1] = L.py;
        size[2] = L.px;
    }
    if (n_hyperplanes)
        *n_hyperplanes = L.hp.S;
    if (n_slots)
        *n_slots = L.hp.npad;
    if (hyperplane_start) {
        // start[s] = 32 * (chunks of hyperplanes < s); rebuilt from the parity prefix kept on the host
        int64_t at = 0;
        for (int s = 0; s < L.hp.S; ++s) {
            hyperplane_start[s] = (int32_t)at;
            at += 32LL * (L.hp.pe_host[s] - (s >= 2 ? L.hp.pe_host[s - 2] : 0));
        }
        hyperplane_start[L.hp.S] = (int32_t)at;
    }
    return FR3D_OK;
}

int fr3d_level_begin(fr3d_ctx* ctx, int level, const float* moving_proc, const float* uvw_init, int B)
{
    FR3D_API_BEGIN(ctx)
    if (level == 0)
        _c->run_level = -1;
    level_begin(_c, level, moving_proc, uvw_init, B);
    FR3D_API_END()
}

int fr3d_level_sweeps(fr3d_ctx* ctx, int level, int t_begin, int t_end, int q_begin, int q_end)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_level == level, "level %d is not open", level);
    sor_launch(_c, _c->state_dtype, t_begin, t_end, q_begin, q_end);
    FR3D_API_END()
}

int fr3d_level_sweeps_slab(fr3d_ctx* ctx, int level, int q_begin, int q_end, int k_begin, int k_end)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_level == level, "level %d is not open", level);
    FR3D_REQUIRE(k_end > 0, "empty plane range");
    sor_launch(_c, _c->state_dtype, 0, _c->iterations, q_begin, q_end, k_begin, k_end);
    FR3D_API_END()
}

// ---- z-slab solve with the halo exchange inside the kernel (peer memory over NVLink, CUDA IPC) ----------------------
int fr3d_ipc_export(fr3d_ctx* ctx, int which, void* handle_out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(handle_out && (which == 0 || which == 1), "bad argument");
#ifdef FR3D_EMU
    FR3D_THROW(FR3D_ERR_ARG, "CUDA IPC is not available in the kernel-logic emulator");
#else
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    void* p = nullptr;
    if (which == 0) {
        // reserve the increment array for the largest level once, so that the handle (and the neighbours' mappings
        // of it) stay valid for the life of the context: the buffer is grow-only and would otherwise move
        size_t need = 0;
        for (auto& L : _c->levels) {
            const size_t n = (size_t)_c->max_batch * L->hp.npad *
                             (_c->state_dtype == FR3D_F64 ? sizeof(Vec4<double>) : sizeof(Vec4<float>));
            need = n > need ? n : need;
        }
        FR3D_REQUIRE(need > 0, "the context has no plan");
        if (_c->d.cap < need) {
            FR3D_REQUIRE(_c->run_level < 0, "export the increment array before a level is opened");
            _c->d.ensure(_c->dev, need);
        }
        p = _c->d.p;
    } else {
        if (!_c->p2p_flags.p) {
            _c->p2p_flags.ensure(_c->dev, 16);
            _c->dev.zero(_c->p2p_flags.p, 16 * sizeof(unsigned));
            FR3D_CUDA(cudaStreamSynchronize(_c->dev.stream));
        }
        p = _c->p2p_flags.p;
    }
    cudaIpcMemHandle_t h;
    FR3D_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out, &h, sizeof(h));
#endif
    FR3D_API_END()
}

int fr3d_ipc_open(fr3d_ctx* ctx, const void* handle, void** ptr_out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(handle && ptr_out, "bad argument");
#ifdef FR3D_EMU
    FR3D_THROW(FR3D_ERR_ARG, "CUDA IPC is not available in the kernel-logic emulator");
#else
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    FR3D_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
#endif
    FR3D_API_END()
}

int fr3d_ipc_close(fr3d_ctx* ctx, void* ptr)
{
    FR3D_API_BEGIN(ctx)
#ifdef FR3D_EMU
    (void)ptr;
    FR3D_THROW(FR3D_ERR_ARG, "CUDA IPC is not available in the kernel-logic emulator");
#else
    if (ptr) {
        FR3D_CUDA(cudaStreamSynchronize(_c->dev.stream));
        FR3D_CUDA(cudaIpcCloseMemHandle(ptr));
    }
#endif
    FR3D_API_END()
}

int fr3d_level_sweeps_slab_p2p(fr3d_ctx* ctx, int level, int k_begin, int k_end, void* lo_d, void* hi_d, void* lo_flags,
                               void* hi_flags, int64_t flag_base)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_level == level, "level %d is not open", level);
#ifdef FR3D_EMU
    (void)k_begin, (void)k_end, (void)lo_d, (void)hi_d, (void)lo_flags, (void)hi_flags, (void)flag_base;
    FR3D_THROW(FR3D_ERR_ARG, "the peer-memory halo exchange needs CUDA devices");
#else
    FR3D_REQUIRE(_c->sp_hp != nullptr, "no level solve has been prepared");
    if (_c->state_dtype == FR3D_F64)
        sor_launch_p2p_t<double>(_c, k_begin, k_end, lo_d, hi_d, lo_flags, hi_flags, flag_base);
    else
        sor_launch_p2p_t<float>(_c, k_begin, k_end, lo_d, hi_d, lo_flags, hi_flags, flag_base);
#endif
    FR3D_API_END()
}

int fr3d_level_planes(fr3d_ctx* ctx, int level, int direction, void* ext, int k_begin, int k_end)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_level == level, "level %d is not open", level);
    FR3D_REQUIRE(ext && (direction == 0 || direction == 1), "bad argument");
    const HPGeom& hp = _c->levels[level]->hp;
    FR3D_REQUIRE(k_begin >= 0 && k_begin < k_end && k_end <= hp.p, "bad plane range");
    const int64_t n = (int64_t)_c->run_B * (k_end - k_begin) * hp.m * hp.n;
    if (_c->state_dtype == FR3D_F64)
        launch(_c->dev, SlabPlanesK<double>{(Vec4<double>*)_c->d.p, (Vec4<double>*)ext, hp.view(), k_begin, k_end, direction}, n);
    else
        launch(_c->dev, SlabPlanesK<float>{(Vec4<float>*)_c->d.p, (Vec4<float>*)ext, hp.view(), k_begin, k_end, direction}, n);
    FR3D_API_END()
}

int fr3d_level_wave_cells(fr3d_ctx* ctx, int level, int direction, void* ext, int k, int q)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_level == level, "level %d is not open", level);
    FR3D_REQUIRE(ext && (direction == 0 || direction == 1), "bad argument");
    const HPGeom& hp = _c->levels[level]->hp;
    FR3D_REQUIRE(k >= 0 && k < hp.p && q >= 0, "bad plane / wave");
    const int T = _c->iterations;
    const int64_t n = (int64_t)_c->run_B * T * hp.m;
    if (_c->state_dtype == FR3D_F64)
        launch(_c->dev, SlabWaveCellsK<double>{(Vec4<double>*)_c->d.p, (Vec4<double>*)ext, hp.view(), k, q, T, direction}, n);
    else
        launch(_c->dev, SlabWaveCellsK<float>{(Vec4<float>*)_c->d.p, (Vec4<float>*)ext, hp.view(), k, q, T, direction}, n);
    FR3D_API_END()
}

int fr3d_level_state(fr3d_ctx* ctx, int level, int direction, void* ext, int64_t slot_begin, int64_t slot_end)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_level == level, "level %d is not open", level);
    FR3D_REQUIRE(ext && (direction == 0 || direction == 1), "bad argument");
    const HPGeom& hp = _c->levels[level]->hp;
    FR3D_REQUIRE(slot_begin >= 0 && slot_begin <= slot_end && slot_end <= hp.npad, "bad slot range");
    const size_t esz = _c->state_dtype == FR3D_F64 ? sizeof(Vec4<double>) : sizeof(Vec4<float>);
    const size_t n = (size_t)(slot_end - slot_begin) * esz;
    for (int b = 0; b < _c->run_B; ++b) {
        char* in = _c->d.p + ((size_t)b * hp.npad + slot_begin) * esz;
        char* ex = (char*)ext + (size_t)b * n;
        if (direction == 0)
            _c->dev.d2d(ex, in, n);
        else
            _c->dev.d2d(in, ex, n);
    }
    FR3D_API_END()
}

int fr3d_level_end(fr3d_ctx* ctx, int level)
{
    FR3D_API_BEGIN(ctx)
    level_end(_c, level);
    FR3D_API_END()
}

int fr3d_level_end_range(fr3d_ctx* ctx, int level, int z_begin, int z_end)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(z_begin <= z_end, "bad z range");
    level_end(_c, level, z_begin, z_end - z_begin);
    FR3D_API_END()
}

int fr3d_flow_slab(fr3d_ctx* ctx, int level, int direction, double* ext, int z_begin, int z_end)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->run_done == level && _c->run_level == -1, "level %d has not been completed", level);
    FR3D_REQUIRE(ext && (direction == 0 || direction == 1), "bad argument");
    const LevelDev& L = *_c->levels[level];
    FR3D_REQUIRE(z_begin >= 0 && z_begin <= z_end && z_end <= L.pz, "bad z range");
    double *ucur, *uprev;
    run_buffers(_c, _c->run_B, ucur, uprev);
    const int64_t plane = (int64_t)L.py * L.px;
    launch(_c->dev, SlabCopyK{ucur, ext, plane, L.pz, z_begin, z_end - z_begin, direction == 0},
           (int64_t)_c->run_B * 3 * (z_end - z_begin) * plane);
    FR3D_API_END()
}

int fr3d_flow_finish(fr3d_ctx* ctx, void* flow_out, int out_dtype)
{
    FR3D_API_BEGIN(ctx)
    flow_finish(_c, flow_out, out_dtype);
    FR3D_API_END()
}

static void warp_common(fr3d_ctx* c, const void* vol, int vdt, const double* d64, const float* d32,
                        const void* ref, int rdt, int B, int Z, int Y, int X, int C, int interp, float* out)
{
    check_dtype(vdt);
    check_dtype(rdt);
    FR3D_REQUIRE(interp == 3 || interp == 1, "Unsupported interpolation method. Use 'linear' or 'cubic'.");
    const int64_t NF = (int64_t)Z * Y * X;
    if (interp == 3)
        spline_prefilter(c, vol, vdt, NF * C, 1, (int64_t)Y * X * C, (int64_t)X * C, C, B, C, Z, Y, X);
    WarpGatherK g;
    g.order = interp;
    g.coef = c->coef.p;
    g.src = vol;
    g.sdt = vdt;
    g.sb = NF * C;
    g.sc = 1;
    g.sz = (int64_t)Y * X * C;
    g.sy = (int64_t)X * C;
    g.sx = C;
    g.disp64 = d64;
    g.disp32 = d32;
    g.hx = g.hy = g.hz = 1.0;
    g.ref = ref;
    g.rdt = rdt;
    g.rc = 1;
    g.rz = (int64_t)Y * X * C;
    g.ry = (int64_t)X * C;
    g.rx = C;
    g.out = out;
    g.ob = NF * C;
    g.oc = 1;
    g.oz = (int64_t)Y * X * C;
    g.oy = (int64_t)X * C;
    g.ox = C;
    g.B = B;
    g.C = C;
    g.Z = Z;
    g.Y = Y;
    g.X = X;
    launch_gather(c, g);
}

int fr3d_compensate(fr3d_ctx* ctx, const void* vol, int vol_dtype, const float* flow, const void* ref,
                    int ref_dtype, int B, float* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(_c->has_plan, "fr3d_compensate needs a context created with a plan");
    FR3D_REQUIRE(vol && flow && ref && out && B >= 1, "null argument");
    warp_common(_c, vol, vol_dtype, nullptr, flow, ref, ref_dtype, B, _c->Z, _c->Y, _c->X, _c->C, _c->interp, out);
    FR3D_API_END()
}

int fr3d_warp(fr3d_ctx* ctx, const void* vol, int vol_dtype, const double* u, const double* v, const double* w,
              const void* ref, int ref_dtype, int Z, int Y, int X, int C, int interp, float* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(vol && u && v && w && ref && out, "null argument");
    FR3D_REQUIRE(Z > 0 && Y > 0 && X > 0 && C >= 1 && C <= FR3D_MAX_CHANNELS, "bad shape");
    const int64_t NF = (int64_t)Z * Y * X;
    double* d = _c->dnat.ensure(_c->dev, (size_t)3 * NF);
    _c->dev.d2d(d, u, NF * sizeof(double));
    _c->dev.d2d(d + NF, v, NF * sizeof(double));
    _c->dev.d2d(d + 2 * NF, w, NF * sizeof(double));
    warp_common(_c, vol, vol_dtype, d, nullptr, ref, ref_dtype, 1, Z, Y, X, C, interp, out);
    FR3D_API_END()
}

int fr3d_resize3d(fr3d_ctx* ctx, const float* src, int nvol, int D, int H, int W, const fr3d_axis_table tables[3],
                  float* dst)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(src && dst && tables && nvol >= 1, "null argument");
    for (int q = 0; q < 3; ++q) {
        _c->stage_tab[q].upload(_c->dev, tables[q]);
        FR3D_REQUIRE(_c->stage_tab[q].P > 0, "empty table");
    }
    const int od = tables[2].out_len, oh = tables[1].out_len, ow = tables[0].out_len;
    resize3<float, float>(_c, src, planar(nvol, D, H, W), 1, nvol, D, H, W, dst, planar(nvol, od, oh, ow),
                          _c->stage_tab);
    FR3D_API_END()
}

int fr3d_motion_tensor(fr3d_ctx* ctx, const float* f1, const float* f2, int p, int m, int n, double hz, double hy,
                       double hx, int f2_f32_math, double* J)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(f1 && f2 && J && p > 0 && m > 0 && n > 0, "bad argument");
    MotionTensorK k{MTImages{f1, f2, MTGeom{p, m, n, hz, hy, hx, f2_f32_math}}, J};
    launch(_c->dev, k, (int64_t)p * m * n);
    FR3D_API_END()
}

int fr3d_motion_tensor_alt(fr3d_ctx* ctx, int kind, const double* f1, const double* f2, int p, int m, int n, double hz,
                           double hy, double hx, double* J)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(f1 && f2 && J && p > 0 && m > 0 && n > 0, "bad argument");
    FR3D_REQUIRE(kind == 1 || kind == 2, "kind must be 1 (gray) or 2 (cs), got %d", kind);
    MotionTensorAltK k{f1, f2, J, p, m, n, kind, hz, hy, hx};
    launch(_c->dev, k, (int64_t)p * m * n);
    FR3D_API_END()
}

int fr3d_sor_level(fr3d_ctx* ctx, const double* J, const double* weight, const double* uvw, int p, int m, int n,
                   int C, const double* alpha, double hz, double hy, double hx, int iterations, int update_lag,
                   const double* a_data, double a_smooth, int sweep, int state_dtype, double* d)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(J && weight && uvw && alpha && a_data && d, "null argument");
    FR3D_REQUIRE(p > 0 && m > 0 && n > 0 && C >= 1 && C <= FR3D_MAX_CHANNELS, "bad shape");
    FR3D_REQUIRE(iterations >= 1 && update_lag >= 1, "iterations/update_lag must be >= 1");
    FR3D_REQUIRE(a_smooth > 0.0, "a_smooth must be positive");
    FR3D_REQUIRE(state_dtype == FR3D_F32 || state_dtype == FR3D_F64, "state_dtype must be FR3D_F32 or FR3D_F64");
    if (!_c->stage_hp || _c->stage_hp->p != p || _c->stage_hp->m != m || _c->stage_hp->n != n) {
        _c->stage_hp.reset(new HPGeom());
        _c->stage_hp->build(_c->dev, p, m, n);
    }
    const HPGeom& hp = *_c->stage_hp;
    const int64_t np = hp.npad;
    double* Js = _c->J.ensure(_c->dev, (size_t)C * 10 * np);
    double* ws = _c->wnat.ensure(_c->dev, (size_t)C * np);
    launch(_c->dev, ToHPK<double>{J, Js, hp.view()}, (int64_t)C * 10 * np);
    launch(_c->dev, ToHPK<double>{weight, ws, hp.view()}, (int64_t)C * np);
    run_sor(_c, state_dtype, hp, 1, C, nullptr, nullptr, 0, Js, uvw, ws, hz, hy, hx, alpha, iterations, update_lag,
            a_data, sweep, a_smooth);
    sor_result(_c, state_dtype, hp, 1, d);
    FR3D_API_END()
}

int fr3d_median5(fr3d_ctx* ctx, const double* src, int nvol, int p, int m, int n, double* dst)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(src && dst && nvol >= 1 && p > 0 && m > 0 && n > 0, "bad argument");
    FR3D_REQUIRE(src != dst, "fr3d_median5 cannot run in place");
    launch_occ2(_c->dev, Median5PairK{src, dst, nullptr, p, m, n, (n + 1) / 2, 0, p}, (int64_t)nvol * p * m * ((n + 1) / 2));
    FR3D_API_END()
}

int fr3d_profile_enable(fr3d_ctx* ctx, int on)
{
    FR3D_API_BEGIN(ctx)
    _c->dev.profile_collect();
    _c->dev.profiling = on != 0;
    FR3D_API_END()
}

// Writes "name\tcount\ttotal_ms\n" lines (kernel names demangled); returns the number of bytes the
// full report needs (excluding the terminator), or a negative status.
int64_t fr3d_profile_report(fr3d_ctx* ctx, char* buf, int64_t cap)
{
    if (!ctx)
        return FR3D_ERR_ARG;
    std::string out;
    for (auto& kv : ctx->dev.profile_collect()) {
        std::string name = kv.first;
#ifndef FR3D_EMU
        int st = 0;
        char* dm = abi::__cxa_demangle(name.c_str(), nullptr, nullptr, &st);
        if (st == 0 && dm) {
            name = dm;
            free(dm);
        }
#endif
        char line[256];
        snprintf(line, sizeof(line), "\t%lld\t%.6f\n", (long long)kv.second.first, kv.second.second);
        out += name + line;
    }
    if (buf && cap > 0) {
        const size_t n = out.size() < (size_t)cap - 1 ? out.size() : (size_t)cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int64_t)out.size();
}

int fr3d_flow_stats(fr3d_ctx* ctx, const float* flow, int B, int Z, int Y, int X, double* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(flow && out && B >= 1 && Z > 0 && Y > 0 && X > 0, "bad argument");
    const int64_t N = (int64_t)Z * Y * X;
    double* acc = _c->wnat.ensure(_c->dev, (size_t)B * 6);
    _c->dev.zero(acc, (size_t)B * 6 * sizeof(double));
    FlowStatsK k;
    k.flow = flow;
    k.acc = acc;
    k.Z = Z;
    k.Y = Y;
    k.X = X;
    k.chunk = 256 * 64;
    k.chunks = (int)((N + k.chunk - 1) / k.chunk);
    launch_tiles(_c->dev, k, (int64_t)B * k.chunks, 256, (size_t)256 * 6 * sizeof(double));
    launch(_c->dev, FlowStatsFinishK{acc, out, 1.0 / (double)N}, B);
    FR3D_API_END()
}

int fr3d_mean_frames(fr3d_ctx* ctx, const float* frames, int T, int64_t n, float* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(frames && out && T >= 1 && n >= 1, "bad argument");
    launch(_c->dev, MeanFramesK{frames, out, T, n}, n);
    FR3D_API_END()
}

int fr3d_mean_frames_f64(fr3d_ctx* ctx, const float* frames, int T, int64_t n, double* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(frames && out && T >= 1 && n >= 1, "bad argument");
    launch(_c->dev, MeanFramesF64K{frames, out, T, n}, n);
    FR3D_API_END()
}

// util/resize_util_3D.py:76-95 (numba fastmath hoists 1/scale: follow the compiled arithmetic so
// that floor() lands on the same tap window as the reference at exact-integer sample positions).
int fr3d_fill_resize_table(int in_len, int out_len, const float* g, int R, int32_t* idx, float* wt)
{
    if (in_len < 1 || out_len < 1 || R < 0 || !g || !idx || !wt)
        return FR3D_ERR_ARG;
    const int P = 2 * R + 4;
    const double A = -0.75;
    const double rscale = 1.0 / ((double)out_len / (double)in_len);
    for (int i = 0; i < out_len; ++i) {
        const double x = ((double)i + 0.5) * rscale - 0.5;
        const int64_t left = (int64_t)floor(x - 2.0) - R;
        double ssum = 0.0;
        for (int p = 0; p < P; ++p) {
            const int64_t j = left + p;
            int64_t jj = j;
            if (in_len <= 1)
                jj = 0;
            else
                while (jj < 0 || jj >= in_len)
                    jj = jj < 0 ? -jj - 1 : 2 * (int64_t)in_len - 1 - jj;
            idx[(int64_t)i * P + p] = (int32_t)jj;
            const double dd = x - (double)j;
            double acc = 0.0;
            for (int u = -R; u <= R; ++u) {
                const double ax = fabs(dd - (double)u);
                double kx = 0.0;
                if (ax < 1.0)
                    kx = (A + 2.0) * ax * ax * ax - (A + 3.0) * ax * ax + 1.0;
                else if (ax < 2.0)
                    kx = A * ax * ax * ax - 5.0 * A * ax * ax + 8.0 * A * ax - 4.0 * A;
                acc += (double)g[u + R] * kx;
            }
            wt[(int64_t)i * P + p] = (float)acc;
            ssum += acc;
        }
        const double inv = 1.0 / ssum;
        for (int p = 0; p < P; ++p)
            wt[(int64_t)i * P + p] = (float)((double)wt[(int64_t)i * P + p] * inv);
    }
    return FR3D_OK;
}

// ---- rigid cross-correlation pre-alignment (fr3d_xcorr.h) -------------------------------------------------------
int fr3d_warp_flow(fr3d_ctx* ctx, const void* vol, int vol_dtype, const float* flow, const void* ref, int ref_dtype,
                   int B, int Z, int Y, int X, int C, int interp, float* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(vol && flow && ref && out && B >= 1, "null argument");
    FR3D_REQUIRE(Z > 0 && Y > 0 && X > 0 && C >= 1 && C <= FR3D_MAX_CHANNELS, "bad shape");
    warp_common(_c, vol, vol_dtype, nullptr, flow, ref, ref_dtype, B, Z, Y, X, C, interp, out);
    FR3D_API_END()
}

int fr3d_cc_project(fr3d_ctx* ctx, const float* vol, int B, int Z, int Y, int X, int acc64, float* pxy, float* pxz)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(vol && pxy && pxz && B >= 1 && Z > 0 && Y > 0 && X > 0, "bad argument");
    launch(_c->dev, CcProjectK{vol, pxy, pxz, B, Z, Y, X, acc64}, (int64_t)B * X * ((int64_t)Y + Z));
    FR3D_API_END()
}

int fr3d_cc_window(fr3d_ctx* ctx, const float* p, int B, int H, int W, const float* hy, const float* hx, double* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(p && hy && hx && out && B >= 1 && H > 0 && W > 0, "bad argument");
    float* mean = _c->cc_f32.ensure(_c->dev, (size_t)B);
    if (_c->dev.cc_block_scans)
        launch_tiles(_c->dev, CcPlaneMeanTileK{p, mean, (int64_t)H * W}, B, 256, (size_t)256 * sizeof(double));
    else
        launch(_c->dev, CcPlaneMeanK{p, mean, (int64_t)H * W}, B);
    launch(_c->dev, CcWindowK{p, mean, hy, hx, out, H, W}, (int64_t)B * H * W);
    FR3D_API_END()
}

int fr3d_cc_cgemm(fr3d_ctx* ctx, const double* A, int64_t a_stride, const double* Bm, int64_t b_stride, double* Cm,
                  int M, int N, int K, int nbatch)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(A && Bm && Cm && M > 0 && N > 0 && K > 0 && nbatch >= 1 && a_stride >= 0 && b_stride >= 0,
                 "bad argument");
    launch(_c->dev, CcGemmK{A, Bm, Cm, a_stride, b_stride, M, N, K}, (int64_t)nbatch * M * N);
    FR3D_API_END()
}

int fr3d_cc_cross_power(fr3d_ctx* ctx, const double* Fr, const double* Fm, double* P, int64_t n, int nbatch,
                        int normalize)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(Fr && Fm && P && n > 0 && nbatch >= 1, "bad argument");
    launch(_c->dev, CcCrossPowerK{Fr, Fm, P, n, normalize}, (int64_t)nbatch * n);
    FR3D_API_END()
}

int fr3d_cc_abs_argmax(fr3d_ctx* ctx, const double* cc, int64_t n, int nbatch, int64_t* idx)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(cc && idx && n > 0 && nbatch >= 1, "bad argument");
    if (_c->dev.cc_block_scans)
        launch_tiles(_c->dev, CcAbsArgmaxTileK{cc, idx, n}, nbatch, 256, (size_t)256 * 2 * sizeof(double));
    else
        launch(_c->dev, CcAbsArgmaxK{cc, idx, n}, nbatch);
    FR3D_API_END()
}

int fr3d_cc_wrap_shift(fr3d_ctx* ctx, const double* img_c, int B, int H, int W, const double* shift_host,
                       double* work, double* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(img_c && shift_host && work && out && B >= 1 && H > 0 && W > 0, "bad argument");
    std::vector<int> order((size_t)B);
    bool any3 = false;
    for (int b = 0; b < B; ++b) {
        // scikit-image: spline order 3 when any component of the shift is fractional, 0 otherwise
        const double sy = shift_host[2 * b], sx = shift_host[2 * b + 1];
        order[b] = (sy != floor(sy) || sx != floor(sx)) ? 3 : 0;
        any3 = any3 || order[b] == 3;
    }
    double* dshift = _c->cc_shift.upload(_c->dev, shift_host, (size_t)2 * B);
    int* dorder = _c->cc_int.upload(_c->dev, order.data(), (size_t)B);
    const int64_t n = (int64_t)H * W;
    launch(_c->dev, StridedCopyK{img_c, work, 2}, (int64_t)B * n); // real parts of the complex planes
    if (any3) {
        // axis 0 (lines along y: element stride W, one line per x), then axis 1
        launch(_c->dev, CcSplineWrapK{work, H, (int64_t)W, 1, n, W}, (int64_t)B * W);
        launch(_c->dev, CcSplineWrapK{work, W, 1, (int64_t)W, n, H}, (int64_t)B * H);
    }
    launch(_c->dev, CcWrapShiftK{work, img_c, dshift, dorder, out, H, W}, (int64_t)B * n);
    FR3D_API_END()
}

int fr3d_cc_tile_sums(fr3d_ctx* ctx, const double* ref_c, const double* shifted, int B, int H, int W,
                      const int* split_host, double* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(ref_c && shifted && split_host && out && B >= 1 && H > 0 && W > 0, "bad argument");
    int* dsplit = _c->cc_int.upload(_c->dev, split_host, (size_t)2 * B);
    if (_c->dev.cc_block_scans)
        launch_tiles(_c->dev, CcTileSumsTileK{ref_c, shifted, dsplit, out, H, W}, (int64_t)B * 4, 256,
                     (size_t)256 * 6 * sizeof(double));
    else
        launch(_c->dev, CcTileSumsK{ref_c, shifted, dsplit, out, H, W}, (int64_t)B * 4);
    FR3D_API_END()
}

int fr3d_rigid_flow(fr3d_ctx* ctx, const float* w_init, const float* rigid_host, int B, int64_t nvox, float* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(w_init && rigid_host && out && B >= 1 && nvox > 0, "bad argument");
    float* dr = _c->cc_f32.upload(_c->dev, rigid_host, (size_t)3 * B);
    launch(_c->dev, CcRigidFlowK{w_init, dr, out, 3 * nvox}, (int64_t)B * 3 * nvox);
    FR3D_API_END()
}

int fr3d_add_flow(fr3d_ctx* ctx, const float* comb, const double* resid, int64_t n, float* out)
{
    FR3D_API_BEGIN(ctx)
    FR3D_REQUIRE(comb && resid && out && n > 0, "bad argument");
    launch(_c->dev, CcAddFlowK{comb, resid, out}, n);
    FR3D_API_END()
}

} // extern "C"

