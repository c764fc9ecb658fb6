// fr3d_sor.h -- nonlinear SOR level solver (core/level_solver_3d.py:314-546), GPU formulation.
//
// The reference sweeps the grid lexicographically (z outer, y, x inner) in place, omega = 1.95,
// for a fixed number of sweeps.  With a 7-point stencil, voxel (k,j,i) of sweep t needs
//   (k-1,j,i), (k,j-1,i), (k,j,i-1)   from sweep t      (already updated)
//   (k+1,j,i), (k,j+1,i), (k,j,i+1)   from sweep t-1    (not yet updated)
// so every (voxel, sweep) pair with equal  q = (k+j+i) + 2t  is independent of every other such
// pair, and all its inputs carry wave index q-1 (or q-2 for its own previous value).  Executing
// waves q = 0, 1, ... therefore reproduces the lexicographic Gauss-Seidel order EXACTLY while
// keeping up to min(T, S/2) sweeps in flight (S = p+m+n-2 hyperplanes).
//
// Boundary handling (reference: set_boundary_3d copies the first interior layer into the ring at
// the start of each sweep; u's ring is an edge replicate): an out-of-domain neighbour contributes
// the voxel's own increment from sweep t-1 and no u difference.
//
// The data-term nonlinearity psi (refreshed every `update_lag` sweeps from the voxel's own
// du,dv,dw) is voxel-local, so the refresh is fused into the sweep: at sweeps t % lag == 0 the
// voxel recomputes psi from the 10C motion-tensor entries, pre-combines the 3x3 system
//     A = sum_c w_c psi_c J_c[:3,:3],  b = sum_c w_c psi_c J_c[:3,3]
// stores it (9 doubles; the diagonal as 1/(2(ax+ay+az) + A_qq)) and uses it for the next sweeps.
//
// Storage is HYPERPLANE-MAJOR (struct HPView): the voxels of hyperplane s = k+j+i are stored
// contiguously (ordered by k, then j), each hyperplane padded to a multiple of 32 slots, so a warp
// work item is "32 consecutive slots of one hyperplane" with no idle lanes except in the pad, and
// the six stencil neighbours come from a per-voxel neighbour table that is shared by every frame
// of the batch.  The increments are kept as {du,dv,dw,-} vectors (one 16-byte access per voxel in
// the float32 state mode).
//
// Arithmetic: float64 with explicit fma() (the reference solver is numba fastmath=True, i.e.
// contraction/reassociation/reciprocal are already licensed there; bitwise equality with it is not
// defined).  State storage (du,dv,dw and the constant Laplacian term) is float64 by default (agrees
// with the oracle to rounding) or float32 (29 % less traffic; 1e-5..1e-4 / 4e-4..1e-2 voxel mean / max
// from the reference, tolerance 0.01 / 0.05); the system matrix is always float64.
#pragma once
#include "fr3d_kernels.h"

namespace fr3d {

#define FR3D_SOR_OMEGA 1.95
#ifndef FR3D_SOR_FRAME_FAST
#define FR3D_SOR_FRAME_FAST 0 /* item order inside a wave: 0 = chunk index fastest, 1 = frame group fastest */
#endif

// Per-state-dtype tuning (measured on B200, config 2, B = 16; profiles/r01_sor_variants.txt):
//   float64 state: two frames in flight per lane, 2 CTAs/SM (128 registers), neighbour increments via L1
//   float32 state: one frame in flight per lane, 3 CTAs/SM (80 registers)
// Neighbour loads through L1 are safe because (i) a wave only writes hyperplanes of its own parity
// while neighbours live on hyperplanes of the other parity, (ii) the voxel's own value bypasses L1,
// and (iii) every CTA executes a gpu-scope fence (which invalidates its SM's L1) at the wave barrier.
#ifndef FR3D_SOR_F32_MINB
#define FR3D_SOR_F32_MINB 3
#endif
#ifndef FR3D_SOR_F64_MINB
#define FR3D_SOR_F64_MINB 2
#endif
#ifndef FR3D_SOR_F64_PAIR
#define FR3D_SOR_F64_PAIR 1
#endif
template <class ST>
struct SorTune;
//   ticket-dealt kernel, float64 state: ONE CTA of 384 threads per SM (12 warps at 168 registers, no spills) -- with
//   round-robin dealing fewer warps lose (52.2 vs 50.1 ms), with tickets they win (47.3 vs 48.4 ms; results/r02_sor_sched.md)
#ifndef FR3D_SOR_F64_DYN_THREADS
#define FR3D_SOR_F64_DYN_THREADS 384
#endif
#ifndef FR3D_SOR_F64_DYN_MINB
#define FR3D_SOR_F64_DYN_MINB 1
#endif
template <>
struct SorTune<double> {
    static constexpr int kPair = FR3D_SOR_F64_PAIR, kMinBlocks = FR3D_SOR_F64_MINB, kNbrCa = 1;
    static constexpr int kDynThreads = FR3D_SOR_F64_DYN_THREADS, kDynMinBlocks = FR3D_SOR_F64_DYN_MINB;
};
template <>
struct SorTune<float> {
    static constexpr int kPair = 0, kMinBlocks = FR3D_SOR_F32_MINB, kNbrCa = 0;
    static constexpr int kDynThreads = 256, kDynMinBlocks = FR3D_SOR_F32_MINB;
};

template <class ST>
struct SorParams {
    HPView g;
    int C, B, T, lag, fg; // fg: frames handled by one warp work item
    int redblack;         // 0: waves q = s + 2t (lexicographic order); 1: waves q = 2t + colour (checkerboard)
    int sched;            // wavefront kernel (FR3D_OPT_SOR_SCHED): bits 0-6 percent of a wave's items handed out through a
                          // ticket counter instead of round-robin; bit 7: psi-refresh items dealt before the plain ones
    // partial execution (sweep-pipelined multi-GPU solve): only sweeps t_begin <= t < t_end and waves
    // q_begin <= q < q_end of the global schedule q = s + 2t; the full solve is [0,T) x [0,num_waves)
    int t_begin, t_end, q_begin, q_end;
    double ax, ay, az;    // alpha_{x,y,z} / h_{x,y,z}^2
    double a_data[FR3D_MAX_CHANNELS];
    const double* J;      // (B, C, 10, npad): J11,J22,J33,J44,J12,J13,J23,J14,J24,J34
    const double* wgt;    // (C, npad), shared by all frames
    const Vec4<ST>* L;    // (B, npad): alpha-weighted Laplacian of u, v, w
    Vec4<ST>* d;          // (B, npad): du, dv, dw (zero-initialised)
    double* AB;           // (B, npad/32, 9, 32) CHUNK-MAJOR: 1/den_u, 1/den_v, 1/den_w, A12, A13, A23, b1-Lu, b2-Lv, b3-Lw
                          // of the 32 slots of a chunk are 2304 contiguous bytes (one bulk copy of the staged kernel)
                          // (nonlinear smoothness: A11, A22, A33 themselves -- the denominator changes per sweep)
    // nonlinear smoothness term (a_smooth != 1), see "Nonlinear smoothness" below
    double a_smooth, hx, hy, hz;
    const int32_t* pe4;   // (S): chunks in hyperplanes s, s-4, s-8, ...
    const int32_t* peR;   // (S): chunks in hyperplanes s, s-2*lag, s-4*lag, ... (kind-balanced dealing; nullptr: not built)
    const Vec4<ST>* U;    // (B, npad): u, v, w
    Vec4<ST>* dold;       // (B, npad): increments before the voxel's latest update
    double* psi_c;        // (B, npad): psi_s at the voxel
    double* psi_r;        // (B, 3, npad): psi_s at the ring voxel next to a boundary voxel along x / y / z
};

// element e (0..8) of the pre-combined system of slot a, frame b
template <class ST>
FR3D_HD int64_t sor_ab_at(const SorParams<ST>& P, int b, int64_t a, int e)
{
    return (((int64_t)b * (P.g.npad >> 5) + (a >> 5)) * 9 + e) * 32 + (a & 31);
}

// psi refresh + pre-combination for one voxel (sweeps with t % lag == 0)
template <int C>
FR3D_HD void sor_refresh(const double* a_data, const double* Jb, const double* wgt, int64_t np, int64_t a,
                         double du, double dv, double dw, double den0, double* A, bool inv_diag = true)
{
    double S[9];
#pragma unroll
    for (int q = 0; q < 9; ++q)
        S[q] = 0.0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double* Jc = Jb + (int64_t)c * 10 * np + a;
        const double J11 = Jc[0], J22 = Jc[np], J33 = Jc[2 * np], J44 = Jc[3 * np], J12 = Jc[4 * np],
                     J13 = Jc[5 * np], J23 = Jc[6 * np], J14 = Jc[7 * np], J24 = Jc[8 * np], J34 = Jc[9 * np];
        double ww = wgt[(int64_t)c * np + a];
        const double adc = a_data[c];
        if (adc != 1.0) {
            // E = (du,dv,dw,1) J (du,dv,dw,1)^T  (level_solver_3d.py:363-375)
            const double ru = fma(J11, du, fma(J12, dv, fma(J13, dw, J14)));
            const double rv = fma(J12, du, fma(J22, dv, fma(J23, dw, J24)));
            const double rw = fma(J13, du, fma(J23, dv, fma(J33, dw, J34)));
            const double r1 = fma(J14, du, fma(J24, dv, fma(J34, dw, J44)));
            double val = fma(ru, du, fma(rv, dv, fma(rw, dw, r1)));
            if (val < 0.0)
                val = 0.0;
            // adc * (val + 1e-6)^(adc - 1)
            ww *= adc * exp((adc - 1.0) * log(val + 1e-6));
        }
        S[0] = fma(ww, J11, S[0]);
        S[1] = fma(ww, J22, S[1]);
        S[2] = fma(ww, J33, S[2]);
        S[3] = fma(ww, J12, S[3]);
        S[4] = fma(ww, J13, S[4]);
        S[5] = fma(ww, J23, S[5]);
        S[6] = fma(ww, J14, S[6]);
        S[7] = fma(ww, J24, S[7]);
        S[8] = fma(ww, J34, S[8]);
    }
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const double den = den0 + S[q];
        A[q] = inv_diag ? (den != 0.0 ? 1.0 / den : 0.0) : S[q]; // reference: denom == 0 -> update is 0
    }
#pragma unroll
    for (int q = 3; q < 9; ++q)
        A[q] = S[q];
}

// Loaded inputs of one voxel update.
template <class ST>
struct SorIn {
    Vec4<ST> own, L, xm, ym, zm, xp, yp, zp;
    double A[9];
};

template <class ST>
FR3D_HD Vec4<ST> sor_update(const SorParams<ST>& P, const SorIn<ST>& r)
{
    const double du = (double)r.own.x, dv = (double)r.own.y, dw = (double)r.own.z;
    const double num_u = fma(P.ax, (double)r.xp.x + (double)r.xm.x,
                             fma(P.ay, (double)r.yp.x + (double)r.ym.x,
                                 fma(P.az, (double)r.zp.x + (double)r.zm.x, (double)r.L.x)));
    const double num_v = fma(P.ax, (double)r.xp.y + (double)r.xm.y,
                             fma(P.ay, (double)r.yp.y + (double)r.ym.y,
                                 fma(P.az, (double)r.zp.y + (double)r.zm.y, (double)r.L.y)));
    const double num_w = fma(P.ax, (double)r.xp.z + (double)r.xm.z,
                             fma(P.ay, (double)r.yp.z + (double)r.ym.z,
                                 fma(P.az, (double)r.zp.z + (double)r.zm.z, (double)r.L.z)));
    const double* A = r.A;
    const double om = FR3D_SOR_OMEGA, om1 = 1.0 - FR3D_SOR_OMEGA;
    const double u1 = (num_u - fma(A[3], dv, fma(A[4], dw, A[6]))) * A[0];
    const double du_n = fma(om, u1, om1 * du);
    const double v1 = (num_v - fma(A[3], du_n, fma(A[5], dw, A[7]))) * A[1];
    const double dv_n = fma(om, v1, om1 * dv);
    const double w1 = (num_w - fma(A[4], du_n, fma(A[5], dv_n, A[8]))) * A[2];
    const double dw_n = fma(om, w1, om1 * dw);
    Vec4<ST> o;
    o.x = (ST)du_n;
    o.y = (ST)dv_n;
    o.z = (ST)dw_n;
    set_pad(o);
    return o;
}

// Wave bookkeeping shared by the CUDA kernel and the emulation loop.  A wave's work is a list of
// warp items (frame group, 32-slot chunk of one of the hyperplanes in flight).
struct SorWave {
    int s_lo, nT;   // hyperplanes in flight: s_lo, s_lo+2, ..., s_lo+2(nT-1)
    int base;       // pe[s_lo-2] (chunks before the first hyperplane in flight, same parity)
    int chunks;     // total chunks of this wave (one frame)
    int items;      // chunks * frame groups
};
// The two small per-hyperplane tables every item consults (shared-memory copies in the CUDA kernel).
struct SorTabs {
    const int32_t* pe;
    const int32_t* start;
};
template <class ST>
FR3D_HD int sor_num_waves(const SorParams<ST>& P)
{
    return P.redblack ? 2 * P.T : P.g.S + 2 * (P.T - 1);
}
template <class ST>
FR3D_HD SorWave sor_wave(const SorParams<ST>& P, const SorTabs& tb, int q)
{
    const int S = P.g.S;
    SorWave w;
    if (P.redblack) {
        // half-sweep q: every hyperplane of parity q & 1 (a voxel's six neighbours have the other parity)
        const int c = q & 1;
        w.s_lo = c;
        w.nT = c < S ? (S - 1 - c) / 2 + 1 : 0;
        w.base = 0;
        w.chunks = w.nT > 0 ? tb.pe[c + 2 * (w.nT - 1)] : 0;
        w.items = w.chunks * ((P.B + P.fg - 1) / P.fg);
        return w;
    }
    int tlo = q - (S - 1);
    tlo = tlo > 0 ? (tlo + 1) / 2 : 0;
    tlo = tlo < P.t_begin ? P.t_begin : tlo;
    int thi = q / 2;
    if (thi > P.t_end - 1)
        thi = P.t_end - 1;
    w.nT = thi >= tlo ? thi - tlo + 1 : 0;
    w.s_lo = q - 2 * thi;
    w.base = 0;
    w.chunks = 0;
    if (w.nT > 0) {
        w.base = w.s_lo >= 2 ? tb.pe[w.s_lo - 2] : 0;
        w.chunks = tb.pe[q - 2 * tlo] - w.base;
    }
    w.items = w.chunks * ((P.B + P.fg - 1) / P.fg);
    return w;
}

// Where a lane works for one warp item: its slot, the six neighbour slots, the frames.
struct SorLoc {
    int64_t a;
    int n0, n1, n2, n3, n4, n5; // n0 < 0: pad slot (nothing to do)
    int b0, b1;
    bool refresh;
};

// Locate lane `lane` of warp item `item` of wave q (table look-ups only; issued one item ahead of
// the arithmetic so that their latency hides behind the previous item's loads).
template <class ST>
FR3D_HD SorLoc sor_locate(const SorParams<ST>& P, const SorTabs& tb, int q, const SorWave& w, int item, int lane)
{
    const HPView& g = P.g;
#if FR3D_SOR_FRAME_FAST
    const int nfg = (P.B + P.fg - 1) / P.fg;
    const int f = item / nfg;
    const int fgi = item - f * nfg;
#else
    const int fgi = item / w.chunks;
    const int f = item - fgi * w.chunks;
#endif
    // hyperplane holding chunk f: smallest r with pe[s_lo + 2r] - base > f   (uniform per warp)
    int lo = 0, hi = w.nT - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tb.pe[w.s_lo + 2 * mid] - w.base > f)
            hi = mid;
        else
            lo = mid + 1;
    }
    const int s = w.s_lo + 2 * lo;
    const int before = (s >= 2 ? tb.pe[s - 2] : 0) - w.base;
    const int t = P.redblack ? (q >> 1) : ((q - s) >> 1);
    SorLoc L;
    L.a = (int64_t)tb.start[s] + 32 * (f - before) + lane;
    const int32_t* nb = g.nbr + HPView::nbr_at(0, L.a);
    L.n0 = nb[0];
    L.n1 = nb[32];
    L.n2 = nb[64];
    L.n3 = nb[96];
    L.n4 = nb[128];
    L.n5 = nb[160];
    L.refresh = (t % P.lag) == 0;
    L.b0 = fgi * P.fg;
    L.b1 = L.b0 + P.fg < P.B ? L.b0 + P.fg : P.B;
    return L;
}

// KIND-BALANCED item order of a wave (FR3D_OPT_SOR_SCHED bit 7).  A psi-refresh item (sweep t % lag == 0: reads 10C
// planes of J, evaluates the robust weights, writes the system) costs ~2.6 plain items, and one wave mixes both kinds --
// every lag-th of its hyperplanes is a refresh hyperplane.  Dealt round-robin in hyperplane order, a warp draws 0..5
// refresh items among its ~9 and the wave lasts as long as its unluckiest warp (measured on a B200, config 2: the wave
// barriers cost 7 ms at lag 1 or with no refresh at all, 11.7 ms at lag 5).  Here the wave's items are enumerated
// refresh chunks first, then plain chunks, as ONE index range dealt round-robin: every warp gets the same number of
// refresh items (+-1) and the heavy items run first.  Refresh hyperplanes of wave q are s = q (mod 2 lag); peR is
// the chunk prefix over that residue class.
// MEASURED SLOWER than the mixed order (52.4 vs 50.1 ms, results/r02_sor_sched.md): a wave of two homogeneous phases --
// all warps streaming J, then all warps sweeping -- overlaps worse than the mix.  Kept as an option (bit 7), not a default.
struct SorKinds {
    int nR;     // refresh hyperplanes in the wave
    int s_min;  // the lowest of them
    int baseR;  // refresh chunks before s_min (same residue class)
    int Rc, Pc; // refresh / plain chunks of the wave
    int r_items; // refresh items = Rc * frame groups
};
template <class ST>
FR3D_HD SorKinds sor_wave_kinds(const SorParams<ST>& P, int q, const SorWave& w)
{
    SorKinds k{0, 0, 0, 0, w.chunks, 0};
    if (w.nT <= 0)
        return k;
    const int thi = (q - w.s_lo) >> 1, tlo = thi - w.nT + 1;
    const int t_first = (tlo + P.lag - 1) / P.lag * P.lag, t_last = thi / P.lag * P.lag;
    if (t_first > t_last)
        return k;
    const int L2 = 2 * P.lag;
    k.nR = (t_last - t_first) / P.lag + 1;
    k.s_min = q - 2 * t_last;
    k.baseR = k.s_min >= L2 ? P.peR[k.s_min - L2] : 0;
    k.Rc = P.peR[q - 2 * t_first] - k.baseR;
    k.Pc = w.chunks - k.Rc;
    k.r_items = k.Rc * ((P.B + P.fg - 1) / P.fg);
    return k;
}

// refresh chunks of the wave in hyperplanes <= s
template <class ST>
FR3D_HD int sor_refresh_through(const SorParams<ST>& P, const SorKinds& k, int s)
{
    if (k.nR == 0 || s < k.s_min)
        return 0;
    const int L2 = 2 * P.lag;
    return P.peR[k.s_min + (s - k.s_min) / L2 * L2] - k.baseR;
}

// Item u of the kind-balanced order: u < r_items is a refresh item, the rest are plain items.
template <class ST>
FR3D_HD SorLoc sor_locate_bal(const SorParams<ST>& P, const SorTabs& tb, int q, const SorWave& w, const SorKinds& k,
                              int u, int lane)
{
    const HPView& g = P.g;
    int s, fgi, in_plane;
    bool refresh;
    if (u < k.r_items) {
        const int L2 = 2 * P.lag;
        fgi = u / k.Rc;
        const int f = u - fgi * k.Rc;
        int lo = 0, hi = k.nR - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (P.peR[k.s_min + L2 * mid] - k.baseR > f)
                hi = mid;
            else
                lo = mid + 1;
        }
        s = k.s_min + L2 * lo;
        in_plane = f - ((s >= L2 ? P.peR[s - L2] : 0) - k.baseR);
        refresh = true;
    } else {
        const int v = u - k.r_items;
        fgi = v / k.Pc;
        const int f = v - fgi * k.Pc;
        // hyperplane holding plain chunk f: smallest r with (chunks - refresh chunks) through s_lo + 2r  >  f
        int lo = 0, hi = w.nT - 1;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            const int sm = w.s_lo + 2 * mid;
            if (tb.pe[sm] - w.base - sor_refresh_through(P, k, sm) > f)
                hi = mid;
            else
                lo = mid + 1;
        }
        s = w.s_lo + 2 * lo;
        const int before = lo > 0 ? tb.pe[s - 2] - w.base - sor_refresh_through(P, k, s - 2) : 0;
        in_plane = f - before;
        refresh = false;
    }
    SorLoc L;
    L.a = (int64_t)tb.start[s] + 32 * in_plane + lane;
    const int32_t* nb = g.nbr + HPView::nbr_at(0, L.a);
    L.n0 = nb[0];
    L.n1 = nb[32];
    L.n2 = nb[64];
    L.n3 = nb[96];
    L.n4 = nb[128];
    L.n5 = nb[160];
    L.refresh = refresh;
    L.b0 = fgi * P.fg;
    L.b1 = L.b0 + P.fg < P.B ? L.b0 + P.fg : P.B;
    return L;
}

template <class ST>
FR3D_HD void sor_load(const SorParams<ST>& P, const SorLoc& L, int b, bool with_ab, SorIn<ST>& r)
{
    const int64_t np = P.g.npad;
    const Vec4<ST>* d = P.d + (int64_t)b * np;
    r.own = ld4_cg(d + L.a);
    if (SorTune<ST>::kNbrCa) {
        r.xm = ld4_ca(d + L.n0);
        r.ym = ld4_ca(d + L.n1);
        r.zm = ld4_ca(d + L.n2);
        r.xp = ld4_ca(d + L.n3);
        r.yp = ld4_ca(d + L.n4);
        r.zp = ld4_ca(d + L.n5);
    } else {
        r.xm = ld4_cg(d + L.n0);
        r.ym = ld4_cg(d + L.n1);
        r.zm = ld4_cg(d + L.n2);
        r.xp = ld4_cg(d + L.n3);
        r.yp = ld4_cg(d + L.n4);
        r.zp = ld4_cg(d + L.n5);
    }
    if (with_ab) {
        // plain sweep: the constant Laplacian term was folded into b1..b3 at the last refresh
        r.L.x = r.L.y = r.L.z = (ST)0;
        const double* AB = P.AB + sor_ab_at(P, b, L.a, 0);
#pragma unroll
        for (int k = 0; k < 9; ++k)
            r.A[k] = FR3D_LDCG(AB + k * 32);
    } else {
        r.L = ld4_cg(P.L + (int64_t)b * np + L.a);
    }
}

// Device-side halo exchange of the z-slab solve (one process per GPU, peer memory over NVLink): the increment arrays
// and the flag words of the two z-neighbours, opened through CUDA IPC.  A voxel on the slab's lowest / highest plane
// is stored to the neighbour's copy of the level as well; flags carry "waves completed" (monotonic over launches).
template <class ST>
struct SorPeers {
    Vec4<ST>* lo_d = nullptr;     // increments of the rank that owns the planes below k_begin (nullptr: none)
    Vec4<ST>* hi_d = nullptr;     // ... above k_end - 1
    unsigned* lo_flag = nullptr;  // lower neighbour's "my upper neighbour has finished wave" word
    unsigned* hi_flag = nullptr;  // upper neighbour's "my lower neighbour has finished wave" word
    unsigned* my_flags = nullptr; // [0]: written by my lower neighbour, [1]: by my upper neighbour
    unsigned base = 0;            // flag value before this launch's first wave
};

template <class ST, bool SLAB>
FR3D_HD void sor_store(const SorParams<ST>& P, int b, int64_t a, const Vec4<ST>& v, int k, int kb, int ke,
                       const SorPeers<ST>* pr)
{
    const int64_t at = (int64_t)b * P.g.npad + a;
    st4_cg(P.d + at, v);
    if (SLAB && pr) {
        if (k == kb && pr->lo_d)
            st4_cg(pr->lo_d + at, v);
        if (k == ke - 1 && pr->hi_d)
            st4_cg(pr->hi_d + at, v);
    }
}

// Update the lane's voxel in every frame of the item.
// SLAB (z-slab multi-GPU solve): only voxels of the planes kb <= k < ke are updated.  It is a separate
// instantiation (and the plane range travels outside SorParams) so that the full solve's kernel is untouched.
template <class ST, int C, bool SLAB = false>
FR3D_HD void sor_process(const SorParams<ST>& P, const SorLoc& L, int kb = 0, int ke = 0,
                         const SorPeers<ST>* pr = nullptr)
{
    if (L.n0 < 0)
        return; // pad slot
    int k = 0;
    if (SLAB) {
        k = P.g.perm[L.a] / (P.g.m * P.g.n); // perm = natural index (k*m + j)*n + i
        if (k < kb || k >= ke)
            return; // another rank's plane
    }
    const int64_t np = P.g.npad;
    const int64_t a = L.a;
    if (L.refresh) {
        const double den0 = 2.0 * P.ax + 2.0 * P.ay + 2.0 * P.az;
        for (int b = L.b0; b < L.b1; ++b) {
            double* AB = P.AB + sor_ab_at(P, b, a, 0);
            SorIn<ST> r;
            sor_load(P, L, b, false, r);
            sor_refresh<C>(P.a_data, P.J + (int64_t)b * C * 10 * np, P.wgt, np, a, (double)r.own.x, (double)r.own.y,
                           (double)r.own.z, den0, r.A);
            // fold the constant Laplacian term of u, v, w into the right-hand side: the plain sweeps then
            // read 9 system entries and no L  (num - b == (num - L) - (b - L))
            r.A[6] -= (double)r.L.x;
            r.A[7] -= (double)r.L.y;
            r.A[8] -= (double)r.L.z;
            r.L.x = r.L.y = r.L.z = (ST)0;
#pragma unroll
            for (int e = 0; e < 9; ++e)
                FR3D_STCG(AB + e * 32, r.A[e]);
            sor_store<ST, SLAB>(P, b, a, sor_update(P, r), k, kb, ke, pr);
        }
        return;
    }
    // plain sweep: two frames in flight per lane (all loads of both issued before the first use)
    int b = L.b0;
    for (; SorTune<ST>::kPair && b + 1 < L.b1; b += 2) {
        SorIn<ST> r[2];
#pragma unroll
        for (int e = 0; e < 2; ++e)
            sor_load(P, L, b + e, true, r[e]);
#pragma unroll
        for (int e = 0; e < 2; ++e)
            sor_store<ST, SLAB>(P, b + e, a, sor_update(P, r[e]), k, kb, ke, pr);
    }
    for (; b < L.b1; ++b) {
        SorIn<ST> r;
        sor_load(P, L, b, true, r);
        sor_store<ST, SLAB>(P, b, a, sor_update(P, r), k, kb, ke, pr);
    }
}

// ------------------------------------------------------------------------------------------------
// Nonlinear smoothness (a_smooth != 1; level_solver_3d.py:262-311, 352-355, 400-493).
// The reference recomputes psi_s = a (|grad(u+du)|^2 + 1e-5)^(a-1) over the whole ring-padded field at
// the start of EVERY sweep t from the increments after sweep t-1 -- and, because set_boundary_3d runs
// after it, from ring values that still hold the state after sweep t-2 -- and weights each face by
// 0.5 (psi_s[c] + psi_s[nb]) alpha / h^2.  psi_s of a voxel therefore depends on its six neighbours'
// previous-sweep values, which stretches the wave schedule to
//      psi task (voxel on hyperplane s, sweep t)   -> wave s + 4t
//      sweep task (same voxel, sweep t)            -> wave s + 4t + 2
// (every input of either task is final at least one wave earlier, and nothing it reads is overwritten
// before it ran).  `dold` keeps each voxel's increments from before its latest update = the stale ring
// values the reference sees.
struct SorSet {       // hyperplanes s_lo, s_lo+4, ... of one task kind in one wave
    int s_lo, nT, base, chunks;
};
template <class ST>
FR3D_HD int sor_nl_num_waves(const SorParams<ST>& P) { return P.g.S + 4 * (P.T - 1) + 2; }
template <class ST>
FR3D_HD SorSet sor_nl_set(const SorParams<ST>& P, int qq) // tasks with s + 4t == qq
{
    SorSet w{0, 0, 0, 0};
    if (qq < 0)
        return w;
    const int S = P.g.S;
    int tlo = qq - (S - 1);
    tlo = tlo > 0 ? (tlo + 3) / 4 : 0;
    int thi = qq / 4;
    if (thi > P.T - 1)
        thi = P.T - 1;
    if (thi < tlo)
        return w;
    w.nT = thi - tlo + 1;
    w.s_lo = qq - 4 * thi;
    w.base = w.s_lo >= 4 ? P.pe4[w.s_lo - 4] : 0;
    w.chunks = P.pe4[qq - 4 * tlo] - w.base;
    return w;
}
// slot of lane `lane` of chunk f of the set; returns the hyperplane through s
template <class ST>
FR3D_HD int64_t sor_nl_slot(const SorParams<ST>& P, const SorSet& w, int f, int lane, int& s)
{
    int lo = 0, hi = w.nT - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (P.pe4[w.s_lo + 4 * mid] - w.base > f)
            hi = mid;
        else
            lo = mid + 1;
    }
    s = w.s_lo + 4 * lo;
    const int before = (s >= 4 ? P.pe4[s - 4] : 0) - w.base;
    return (int64_t)P.g.start[s] + 32 * (f - before) + lane;
}

template <class ST>
struct V3 {
    double x, y, z;
};
template <class ST>
FR3D_HD V3<ST> v3add(const Vec4<ST>& a, const Vec4<ST>& b)
{
    return V3<ST>{(double)a.x + (double)b.x, (double)a.y + (double)b.y, (double)a.z + (double)b.z};
}

// psi_s at voxel a and at the ring voxels next to it, frame b.
template <class ST>
FR3D_HD void sor_nl_psi(const SorParams<ST>& P, int64_t a, int b)
{
    const int64_t np = P.g.npad;
    const int32_t* nbt = P.g.nbr + HPView::nbr_at(0, a);
    const int nb[6] = {nbt[0], nbt[32], nbt[64], nbt[96], nbt[128], nbt[160]};
    if (nb[0] < 0)
        return; // pad slot
    const Vec4<ST>* d = P.d + (int64_t)b * np;
    const Vec4<ST>* dold = P.dold + (int64_t)b * np;
    const Vec4<ST>* U = P.U + (int64_t)b * np;
    const Vec4<ST> Uc = ld4_cg(U + a);
    const V3<ST> cur = v3add(Uc, ld4_cg(d + a));       // u + du after sweep t-1
    const V3<ST> ring = v3add(Uc, ld4_cg(dold + a));   // what a ring copy of this voxel holds: state after sweep t-2
    V3<ST> uu[6], ro[6]; // neighbour values as the interior sees them / as the ring holds them (state t-2)
    bool in[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        in[q] = nb[q] != (int)a;
        if (in[q]) {
            const Vec4<ST> Un = ld4_cg(U + nb[q]);
            uu[q] = v3add(Un, ld4_cg(d + nb[q]));
            ro[q] = v3add(Un, ld4_cg(dold + nb[q]));
        } else {
            uu[q] = ring;
            ro[q] = ring;
        }
    }
    const double h2[3] = {2.0 * P.hx, 2.0 * P.hy, 2.0 * P.hz};
    double g = 0.0;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const double dx = (uu[ax + 3].x - uu[ax].x) / h2[ax], dy = (uu[ax + 3].y - uu[ax].y) / h2[ax],
                     dz = (uu[ax + 3].z - uu[ax].z) / h2[ax];
        g = fma(dx, dx, fma(dy, dy, fma(dz, dz, g)));
    }
    if (g < 0.0)
        g = 0.0;
    const double am1 = P.a_smooth - 1.0;
    P.psi_c[(int64_t)b * np + a] = P.a_smooth * exp(am1 * log(g + 1e-5));
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (in[ax] && in[ax + 3])
            continue;
        // ring voxel beyond this voxel along ax: one-sided difference along ax (this voxel's current value
        // against the ring's stale copy), central differences of stale ring copies along the other two axes
        double gr = 0.0;
        {
            const double dx = (cur.x - ring.x) / h2[ax], dy = (cur.y - ring.y) / h2[ax], dz = (cur.z - ring.z) / h2[ax];
            gr = fma(dx, dx, fma(dy, dy, fma(dz, dz, gr)));
        }
#pragma unroll
        for (int bx = 0; bx < 3; ++bx) {
            if (bx == ax)
                continue;
            const double dx = (ro[bx + 3].x - ro[bx].x) / h2[bx], dy = (ro[bx + 3].y - ro[bx].y) / h2[bx],
                         dz = (ro[bx + 3].z - ro[bx].z) / h2[bx];
            gr = fma(dx, dx, fma(dy, dy, fma(dz, dz, gr)));
        }
        if (gr < 0.0)
            gr = 0.0;
        P.psi_r[((int64_t)b * 3 + ax) * np + a] = P.a_smooth * exp(am1 * log(gr + 1e-5));
    }
}

// One sweep update of voxel a, frame b, sweep t (nonlinear smoothness).
template <class ST, int C>
FR3D_HD void sor_nl_update(const SorParams<ST>& P, int64_t a, int b, int t)
{
    const int64_t np = P.g.npad;
    const int32_t* nbt = P.g.nbr + HPView::nbr_at(0, a);
    const int nb[6] = {nbt[0], nbt[32], nbt[64], nbt[96], nbt[128], nbt[160]};
    if (nb[0] < 0)
        return;
    Vec4<ST>* d = P.d + (int64_t)b * np;
    const Vec4<ST>* U = P.U + (int64_t)b * np;
    const double* pc = P.psi_c + (int64_t)b * np;
    const double* pr = P.psi_r + (int64_t)b * 3 * np;
    const Vec4<ST> own = ld4_cg(d + a);
    const Vec4<ST> Uc = ld4_cg(U + a);
    const double psc = FR3D_LDCG(pc + a);
    const double aw[3] = {P.ax, P.ay, P.az};
    double num[3] = {0.0, 0.0, 0.0}, den = 0.0;
    const int order[6] = {2, 5, 1, 4, 0, 3}; // z-, z+, y-, y+, x-, x+ (level_solver_3d.py:400-493)
#pragma unroll
    for (int o = 0; o < 6; ++o) {
        const int q = order[o], ax = q % 3;
        const bool in = nb[q] != (int)a;
        const double psn = in ? FR3D_LDCG(pc + nb[q]) : FR3D_LDCG(pr + (int64_t)ax * np + a);
        const double tmp = 0.5 * (psc + psn) * aw[ax];
        const Vec4<ST> dn = in ? ld4_cg(d + nb[q]) : own;
        const Vec4<ST> Un = in ? ld4_cg(U + nb[q]) : Uc;
        num[0] = fma(tmp, ((double)Un.x + (double)dn.x) - (double)Uc.x, num[0]);
        num[1] = fma(tmp, ((double)Un.y + (double)dn.y) - (double)Uc.y, num[1]);
        num[2] = fma(tmp, ((double)Un.z + (double)dn.z) - (double)Uc.z, num[2]);
        den += tmp;
    }
    double A[9];
    double* AB = P.AB + sor_ab_at(P, b, a, 0);
    const double du = (double)own.x, dv = (double)own.y, dw = (double)own.z;
    if ((t % P.lag) == 0) {
        sor_refresh<C>(P.a_data, P.J + (int64_t)b * C * 10 * np, P.wgt, np, a, du, dv, dw, 0.0, A, false);
#pragma unroll
        for (int e = 0; e < 9; ++e)
            FR3D_STCG(AB + e * 32, A[e]);
    } else {
#pragma unroll
        for (int e = 0; e < 9; ++e)
            A[e] = FR3D_LDCG(AB + e * 32);
    }
    const double om = FR3D_SOR_OMEGA, om1 = 1.0 - FR3D_SOR_OMEGA;
    const double den_u = den + A[0], den_v = den + A[1], den_w = den + A[2];
    const double u1 = den_u != 0.0 ? (num[0] - fma(A[3], dv, fma(A[4], dw, A[6]))) / den_u : 0.0;
    const double du_n = fma(om, u1, om1 * du);
    const double v1 = den_v != 0.0 ? (num[1] - fma(A[3], du_n, fma(A[5], dw, A[7]))) / den_v : 0.0;
    const double dv_n = fma(om, v1, om1 * dv);
    const double w1 = den_w != 0.0 ? (num[2] - fma(A[4], du_n, fma(A[5], dv_n, A[8]))) / den_w : 0.0;
    const double dw_n = fma(om, w1, om1 * dw);
    Vec4<ST> o;
    o.x = (ST)du_n;
    o.y = (ST)dv_n;
    o.z = (ST)dw_n;
    set_pad(o);
    st4_cg(P.dold + (int64_t)b * np + a, own);
    st4_cg(d + a, o);
}

// Warp item `item` of wave q: psi items first, then sweep items; each for all frames.
template <class ST, int C>
FR3D_HD void sor_nl_item(const SorParams<ST>& P, int q, const SorSet& wp, const SorSet& ws, int item, int lane)
{
    int s;
    if (item < wp.chunks) {
        const int64_t a = sor_nl_slot(P, wp, item, lane, s);
        for (int b = 0; b < P.B; ++b)
            sor_nl_psi(P, a, b);
    } else {
        const int64_t a = sor_nl_slot(P, ws, item - wp.chunks, lane, s);
        const int t = (q - 2 - s) >> 2;
        for (int b = 0; b < P.B; ++b)
            sor_nl_update<ST, C>(P, a, b, t);
    }
}

// peak number of warp items over all waves (host; pe_host = host copy of the chunk prefix)
inline int sor_peak_items(int S, int T, int B, int fg, const int32_t* pe_host, int redblack)
{
    int peak = 0;
    if (redblack) {
        for (int c = 0; c < 2 && c < S; ++c) {
            const int last = c + 2 * ((S - 1 - c) / 2);
            peak = pe_host[last] > peak ? pe_host[last] : peak;
        }
        return peak * ((B + fg - 1) / fg);
    }
    const int nw = S + 2 * (T - 1);
    for (int q = 0; q < nw; ++q) {
        int tlo = q - (S - 1);
        tlo = tlo > 0 ? (tlo + 1) / 2 : 0;
        int thi = q / 2;
        if (thi > T - 1)
            thi = T - 1;
        if (thi < tlo)
            continue;
        const int s_lo = q - 2 * thi;
        const int c = pe_host[q - 2 * tlo] - (s_lo >= 2 ? pe_host[s_lo - 2] : 0);
        if (c > peak)
            peak = c;
    }
    return peak * ((B + fg - 1) / fg);
}

} // namespace fr3d
#include "fr3d_sor_tile.h"
namespace fr3d {

// Which levels the time-blocked tile kernel takes (FR3D_OPT_SOR_KERNEL = 2, the default): the plain lexicographic
// full solve.  Partial sweep ranges (sweep-pipelined multi-GPU solve), z-slab launches, the red-black order and the
// nonlinear smoothness term keep the wavefront kernels.
template <class ST>
inline bool sor_tile_applicable(const Device& dev, const SorParams<ST>& P)
{
    return dev.sor_kernel == 2 && dev.sor_k1 <= 0 && !P.redblack && P.a_smooth == 1.0 && P.t_begin == 0 &&
           P.t_end == P.T && P.q_begin == 0 && P.q_end == sor_num_waves(P);
}
inline SorTileGeom sor_tile_geom_for(const Device& dev, int p, int m, int n, int T, int lag)
{
    // sweeps per time block: the update lag when a tile can hold it (psi refreshes then fall on local sweep 0 and run
    // as one parallel pass), else its largest divisor <= 8, else 5 with the refresh inside the waves
    int Tb = dev.sor_tile_sweeps;
    if (Tb <= 0) {
        Tb = 5;
        for (int d = 8; d >= 2; --d)
            if (lag % d == 0) {
                Tb = d;
                break;
            }
    }
    const int tk = dev.sor_tile_k > 0 ? dev.sor_tile_k : 8, tj = dev.sor_tile_j > 0 ? dev.sor_tile_j : 8,
              ti = dev.sor_tile_i > 0 ? dev.sor_tile_i : 8;
    return sor_tile_geom(p, m, n, T, lag, Tb, tk, tj, ti);
}

// words of the barrier buffer: [0] arrival counter, [32 + wave] item tickets of the wavefront kernel (one per wave)
#define FR3D_SOR_BAR_WORDS (32 + 16384)

#ifdef FR3D_EMU
template <class ST, int C>
inline void sor_run_tiles(Device& dev, const SorParams<ST>& P)
{
    const SorTileGeom G = sor_tile_geom_for(dev, P.g.p, P.g.m, P.g.n, P.T, P.lag);
    std::vector<unsigned char> smem(sor_tile_smem<ST>(G) + 16);
    sor_tile_run_block<ST, C>(P, G, smem.data(), 0, 1, SorSerialPar(), [](int) {});
    dev.emu_count("fr3d_sor_tiles");
    dev.launches++;
}

template <class ST, int C>
inline void sor_run_c(Device& dev, const SorParams<ST>& P, unsigned*)
{
    if (sor_tile_applicable(dev, P)) {
        sor_run_tiles<ST, C>(dev, P);
        return;
    }
    if (P.a_smooth != 1.0) {
        const int nw = sor_nl_num_waves(P);
        for (int q = 0; q < nw; ++q) {
            const SorSet wp = sor_nl_set(P, q), ws = sor_nl_set(P, q - 2);
            for (int item = 0; item < wp.chunks + ws.chunks; ++item)
                for (int lane = 0; lane < 32; ++lane)
                    sor_nl_item<ST, C>(P, q, wp, ws, item, lane);
        }
        dev.launches++;
        return;
    }
    const SorTabs tb{P.g.pe, P.g.start};
    const bool balanced = (P.sched & 128) && P.peR && dev.sor_k1 <= 0 && !P.redblack;
    if (balanced)
        dev.emu_count("fr3d_sor_wavefront_balanced");
    for (int q = P.q_begin; q < P.q_end; ++q) {
        const SorWave w = sor_wave(P, tb, q);
        if (balanced) { // the kind-balanced enumeration of the wave's items (same items, other order)
            const SorKinds kd = sor_wave_kinds(P, q, w);
            for (int u = w.items - 1; u >= 0; --u)
                for (int lane = 0; lane < 32; ++lane)
                    sor_process<ST, C>(P, sor_locate_bal(P, tb, q, w, kd, u, lane));
            continue;
        }
        for (int item = 0; item < w.items; ++item)
            for (int lane = 0; lane < 32; ++lane) {
                if (dev.sor_k1 > 0)
                    sor_process<ST, C, true>(P, sor_locate(P, tb, q, w, item, lane), dev.sor_k0, dev.sor_k1);
                else
                    sor_process<ST, C>(P, sor_locate(P, tb, q, w, item, lane));
            }
    }
    dev.launches++;
}
#else
#ifndef FR3D_SOR_THREADS
#define FR3D_SOR_THREADS 256
#endif

#ifndef FR3D_BARRIER_VARIANT
#define FR3D_BARRIER_VARIANT 0
#endif
// Grid barrier between waves: every CTA's thread 0 arrives on a global counter and spins until all CTAs of
// the generation have arrived.  The gpu-scope fence / acquire after the spin also invalidates the SM's L1
// (SASS: CCTL.IVALL), which the L1-cached neighbour loads of the next wave rely on.
__device__ __forceinline__ void fr3d_grid_barrier(unsigned* ctr, unsigned target)
{
#ifdef FR3D_SOR_TIMING_NO_BARRIER /* DIAGNOSTIC BUILD ONLY (wrong results): what the waves cost without their barrier */
    __syncthreads();
    return;
#endif
    __syncthreads();
    if (threadIdx.x == 0) {
#if FR3D_BARRIER_VARIANT == 0
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*((volatile unsigned*)ctr) < target) {
        }
        __threadfence();
#elif FR3D_BARRIER_VARIANT == 3 /* DIAGNOSTIC (wrong results): the fences of the barrier without its synchronisation */
        __threadfence();
        __threadfence();
        (void)ctr;
        (void)target;
#elif FR3D_BARRIER_VARIANT == 4 /* DIAGNOSTIC for .ca neighbour loads (no L1 invalidation): release-arrive, relaxed poll */
        unsigned v;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
#else
        unsigned v;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        do {
#if FR3D_BARRIER_VARIANT == 2
            __nanosleep(32);
#endif
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
        __threadfence();
#endif
    }
    __syncthreads();
}

// L2 prefetch of the streamed inputs of a located item (the warp's NEXT item, while the current one is processed):
// the pre-combined system (chunk-major: 2304 contiguous bytes per frame) and the voxels' own increments, or, for a
// psi-refresh item, the 10C motion-tensor planes and the Laplacian term.  prefetch.global.L2 holds no register and no
// scoreboard entry, so the bytes in flight per SM are no longer bounded by the register file; the demand loads of
// the next item then find their lines in L2.  Neighbour increments were written one wave ago and are not prefetched.
// MEASURED AND REJECTED (B200, config 2, B = 25, results/r02_sor_experiments.md): 55.3 vs 53.2 ms (float64 state),
// 50.1 vs 44.6 ms (float32 state) -- more bytes in flight make the kernel slower, i.e. it is not waiting for
// latency that a deeper queue could hide.  Kept off by default as a build option.
#ifndef FR3D_SOR_PREFETCH
#define FR3D_SOR_PREFETCH 0
#endif
template <class ST, int C>
__device__ __forceinline__ void sor_prefetch_item(const SorParams<ST>& P, const SorLoc& L, int lane)
{
#if FR3D_SOR_PREFETCH
    const int64_t np = P.g.npad;
    const int64_t a0 = L.a - lane; // first slot of the chunk
    for (int b = L.b0; b < L.b1; ++b) {
        const char* own = reinterpret_cast<const char*>(P.d + (int64_t)b * np + a0);
        constexpr int kOwnLines = 32 * (int)sizeof(Vec4<ST>) / 128;
        if (!L.refresh) {
            const char* ab = reinterpret_cast<const char*>(P.AB + sor_ab_at(P, b, a0, 0));
            if (lane < 18)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ab + lane * 128));
            else if (lane < 18 + kOwnLines)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(own + (lane - 18) * 128));
        } else {
            const double* Jb = P.J + (int64_t)b * C * 10 * np + a0;
            for (int li = lane; li < C * 10 * 2; li += 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(Jb + (int64_t)(li >> 1) * np) + (li & 1) * 128));
            const char* Lp = reinterpret_cast<const char*>(P.L + (int64_t)b * np + a0);
            if (lane < kOwnLines)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(own + lane * 128));
            else if (lane < 2 * kOwnLines)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(Lp + (lane - kOwnLines) * 128));
        }
    }
#endif
}

// Persistent cooperative kernel: all waves of one level solve, one grid barrier per wave.
// Dynamic shared memory: copies of the pe / start tables (tabs_in_smem) so that locating an item
// costs shared-memory latency only.
template <class ST, int C, bool SLAB = false>
__global__ void __launch_bounds__(FR3D_SOR_THREADS, SorTune<ST>::kMinBlocks)
fr3d_sor_wavefront(const SorParams<ST> P, unsigned* bar, int tabs_in_smem)
{
    // SLAB: the plane range rides in the upper bits of the flag word (tabs | k_begin << 1 | k_end << 16)
    const int kb = SLAB ? ((tabs_in_smem >> 1) & 0x7fff) : 0, ke = SLAB ? (tabs_in_smem >> 16) : 0;
    if (SLAB)
        tabs_in_smem &= 1;
    extern __shared__ int32_t fr3d_sor_smem[];
    SorTabs tb{P.g.pe, P.g.start};
    if (tabs_in_smem) {
        const int S = P.g.S;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            fr3d_sor_smem[i] = P.g.pe[i];
        for (int i = threadIdx.x; i <= S; i += blockDim.x)
            fr3d_sor_smem[S + i] = P.g.start[i];
        __syncthreads();
        tb.pe = fr3d_sor_smem;
        tb.start = fr3d_sor_smem + S;
    }
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * wpb;
    unsigned gen = 0;
    for (int q = P.q_begin; q < P.q_end; ++q) {
        const SorWave w = sor_wave(P, tb, q);
        int item = blockIdx.x * wpb + warp;
        if (item < w.items) {
            SorLoc cur = sor_locate(P, tb, q, w, item, lane);
            for (;;) {
                const int next = item + stride;
                const bool more = next < w.items;
                SorLoc nxt;
                if (more) {
                    nxt = sor_locate(P, tb, q, w, next, lane); // neighbour-table loads fly during the update below
                    sor_prefetch_item<ST, C>(P, nxt, lane);
                }
                sor_process<ST, C, SLAB>(P, cur, kb, ke);
                if (!more)
                    break;
                cur = nxt;
                item = next;
            }
        }
        ++gen;
        fr3d_grid_barrier(bar, gen * gridDim.x);
    }
}

// The same kernel with two changes to HOW a wave's items reach the warps (FR3D_OPT_SOR_SCHED, full solves only; every
// item runs the same code on the same inputs, so results are bit-identical):
//  * BAL: the kind-balanced item order (sor_locate_bal above) -- refresh items first, equally many per warp;
//  * tickets: the first rounds are dealt round-robin (item = warp + r * stride, no traffic); the last `pct` percent of
//    the wave's items are handed out through a per-wave ticket counter (bar[32 + wave], zeroed before the launch), so
//    that warps which drew cheap items (L2 hits) take more of them and the wave ends within one item's time on all
//    warps.  A ticket is drawn one item ahead of its use (the atomic's latency hides behind the update in between).
template <class ST, int C, bool BAL>
__global__ void __launch_bounds__(SorTune<ST>::kDynThreads, SorTune<ST>::kDynMinBlocks)
fr3d_sor_wavefront_dyn(const SorParams<ST> P, unsigned* bar, int tabs_in_smem)
{
    extern __shared__ int32_t fr3d_sor_smem[];
    SorTabs tb{P.g.pe, P.g.start};
    if (tabs_in_smem) {
        const int S = P.g.S;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            fr3d_sor_smem[i] = P.g.pe[i];
        for (int i = threadIdx.x; i <= S; i += blockDim.x)
            fr3d_sor_smem[S + i] = P.g.start[i];
        __syncthreads();
        tb.pe = fr3d_sor_smem;
        tb.start = fr3d_sor_smem + S;
    }
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * wpb;
    const int pct = P.sched & 127;
    unsigned gen = 0;
    for (int q = P.q_begin; q < P.q_end; ++q) {
        const SorWave w = sor_wave(P, tb, q);
        int item = blockIdx.x * wpb + warp;
        if (item < w.items) {
            SorKinds kd;
            if (BAL)
                kd = sor_wave_kinds(P, q, w);
            int dyn_base = 0x3fffffff; // items below are dealt round-robin
            if (pct) {
                const int rs = (int)(((int64_t)w.items * (100 - pct)) / (100 * (int64_t)stride));
                dyn_base = (rs < 1 ? 1 : rs) * stride;
            }
            unsigned* ticket = bar + 32 + (q - P.q_begin);
            SorLoc cur = BAL ? sor_locate_bal(P, tb, q, w, kd, item, lane) : sor_locate(P, tb, q, w, item, lane);
            bool t_dyn = item + stride >= dyn_base;
            int t_val = t_dyn ? (lane == 0 ? (int)atomicAdd(ticket, 1u) : 0) : item + stride;
            for (;;) {
                const int next = t_dyn ? dyn_base + __shfl_sync(0xffffffffu, t_val, 0) : t_val;
                const bool more = next < w.items;
                SorLoc nxt;
                if (more) {
                    nxt = BAL ? sor_locate_bal(P, tb, q, w, kd, next, lane) : sor_locate(P, tb, q, w, next, lane);
                    t_dyn = next + stride >= dyn_base;
                    t_val = t_dyn ? (lane == 0 ? (int)atomicAdd(ticket, 1u) : 0) : next + stride;
                }
                sor_process<ST, C, false>(P, cur, 0, 0);
                if (!more)
                    break;
                cur = nxt;
            }
        }
        ++gen;
        fr3d_grid_barrier(bar, gen * gridDim.x);
    }
}

#ifdef FR3D_SOR_SPLIT_EXPERIMENT
// SPLIT-PHASE waves (EXPERIMENT, compiled with -DFR3D_SOR_SPLIT_EXPERIMENT only; then FR3D_OPT_SOR_SCHED bit 10 selects
// it and bits 8-9 its arrival mode).  MEASURED AND REJECTED on a B200 (config 2, 25 frames, results/r02_sor_sched.md):
// 51.8 - 53.6 ms in its four modes against 50.1 ms for the barrier kernel.  Measured on a B200 (config 2, 25 frames): with the
// grid barrier compiled out the same waves take 38.4 instead of 50.5 ms -- a quarter of the solve is the barrier's
// latency plus the drain of each wave's last items and the refill after it, ~12 us on each of ~1000 waves.  The frames
// of a batch are independent systems, so the kernel splits them into two groups and alternates  A_q, B_q, A_q+1, ...:
// a warp ARRIVES on group A's counter when its A_q items are stored, goes straight on to its B_q items, and only then
// WAITS for "all warps have arrived for A_q" -- which by then happened long ago.  No CTA-wide or grid-wide stall is
// left: arrival and wait are per warp (stores -> fence -> red on the group's counter;  poll -> fence, whose acquire
// also drops the SM's L1 lines that the .ca neighbour loads of the next wave must not see).  With items enumerated
// chunk-fastest, a group's items are one contiguous index range of the wave.  Same items, same arithmetic, same order
// of dependent updates: bit-identical.
template <class ST, int C>
__global__ void __launch_bounds__(FR3D_SOR_THREADS, SorTune<ST>::kMinBlocks)
fr3d_sor_wavefront_split(const SorParams<ST> P, unsigned* bar, int tabs_in_smem)
{
    extern __shared__ int32_t fr3d_sor_smem[];
    SorTabs tb{P.g.pe, P.g.start};
    if (tabs_in_smem) {
        const int S = P.g.S;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            fr3d_sor_smem[i] = P.g.pe[i];
        for (int i = threadIdx.x; i <= S; i += blockDim.x)
            fr3d_sor_smem[S + i] = P.g.start[i];
        __syncthreads();
        tb.pe = fr3d_sor_smem;
        tb.start = fr3d_sor_smem + S;
    }
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * wpb;
    const int gw = blockIdx.x * wpb + warp;
    const int nfg = (P.B + P.fg - 1) / P.fg;
    const int nfg0 = (nfg + 1) >> 1; // frame groups of the first half
    // arrival / wait granularity and primitives (experiment switches, FR3D_OPT_SOR_SCHED bits 8, 9):
    //   mode 0: per warp, __threadfence both sides;  1: per warp, red.release / ld.acquire;  2: per CTA (thread 0
    //   arrives after a __syncthreads, thread 32 polls the other group meanwhile);  3: per CTA with release / acquire
    const int mode = (P.sched >> 8) & 3;
    const bool per_cta = mode >= 2, relacq = mode & 1;
    const unsigned members = per_cta ? gridDim.x : (unsigned)stride;
    unsigned gen = 0;
    for (int q = P.q_begin; q < P.q_end; ++q, ++gen) {
        const SorWave w = sor_wave(P, tb, q);
        const int split = w.chunks * nfg0;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            const int hi = g ? w.items : (split < w.items ? split : w.items);
            int item = (g ? split : 0) + gw;
            unsigned* ctr = bar + 32 + 32 * g;
            // This group's previous wave must be complete everywhere.  EVERY member waits, also one without items in
            // this phase: a member then arrives for wave q only after it saw all arrivals of wave q - 1, which is what
            // makes "counter >= gen * members" mean "all members have arrived gen times".
            if (gen) {
                const unsigned target = gen * members;
                if (per_cta ? threadIdx.x == 32 : lane == 0) {
                    if (relacq) {
                        unsigned v;
                        do {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
                        } while (v < target);
                    } else {
                        while (*((volatile unsigned*)ctr) < target) {
                        }
                        __threadfence();
                    }
                }
            }
            if (per_cta)
                __syncthreads(); // (thread 0 has arrived for the other group, thread 32 has seen this group's arrivals)
            else
                __syncwarp();
            if (item < hi) {
                SorLoc cur = sor_locate(P, tb, q, w, item, lane);
                for (;;) {
                    const int next = item + stride;
                    const bool more = next < hi;
                    SorLoc nxt;
                    if (more)
                        nxt = sor_locate(P, tb, q, w, next, lane);
                    sor_process<ST, C, false>(P, cur, 0, 0);
                    if (!more)
                        break;
                    cur = nxt;
                    item = next;
                }
            }
            if (per_cta) {
                __syncthreads();
                if (threadIdx.x == 0) {
                    if (relacq) {
                        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
                    } else {
                        __threadfence();
                        atomicAdd(ctr, 1u);
                    }
                }
            } else {
                if (!relacq)
                    __threadfence();
                __syncwarp();
                if (lane == 0) {
                    if (relacq)
                        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
                    else
                        atomicAdd(ctr, 1u);
                }
            }
        }
    }
}
#endif

// z-slab solve with the halo exchange INSIDE the persistent kernel: every rank sweeps its own planes wave by wave;
// a boundary-plane voxel is written to the z-neighbour's memory as it is produced (peer store over NVLink), and after
// every wave the ranks hand each other a "wave q done" flag.  One launch per level instead of one launch, two pack
// kernels, a host synchronisation and an NCCL message pair per wave.
//   wave end:  all stores of the CTA -> __syncthreads -> thread 0: system-scope fence, arrive on the local counter,
//              wait for all CTAs;  block 0 then publishes base + q + 1 in both neighbours' flag words;
//   wave start: thread 0 of every CTA waits until both of its own flag words have reached base + q (the neighbours'
//              boundary values of wave q - 1 are then in this GPU's memory), fences, releases the CTA.
__device__ __forceinline__ void sor_p2p_wait(volatile unsigned* flag, unsigned target)
{
    long long t0 = 0;
    while ((int)(*flag - target) < 0) {
        if (t0 == 0)
            t0 = clock64();
        else if (clock64() - t0 > 40000000000LL)
            __trap(); // ~20 s: a neighbour died or the ranks disagree on the schedule -- fail instead of hanging
    }
}
template <class ST, int C>
__global__ void __launch_bounds__(FR3D_SOR_THREADS, SorTune<ST>::kMinBlocks)
fr3d_sor_wavefront_p2p(const SorParams<ST> P, unsigned* bar, int tabs_in_smem, int kb, int ke, const SorPeers<ST> pr)
{
    extern __shared__ int32_t fr3d_sor_smem[];
    SorTabs tb{P.g.pe, P.g.start};
    if (tabs_in_smem) {
        const int S = P.g.S;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            fr3d_sor_smem[i] = P.g.pe[i];
        for (int i = threadIdx.x; i <= S; i += blockDim.x)
            fr3d_sor_smem[S + i] = P.g.start[i];
        __syncthreads();
        tb.pe = fr3d_sor_smem;
        tb.start = fr3d_sor_smem + S;
    }
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x * wpb;
    unsigned gen = 0;
    for (int q = P.q_begin; q < P.q_end; ++q) {
        if (q > P.q_begin) {
            // the neighbours' boundary values of wave q - 1.  (Measured and rejected on 2 B200: letting every warp
            // wait only before its first boundary-plane item, so that the flag latency hides behind interior items,
            // costs a perm load and a vote per item and is SLOWER: 420 vs 379 ms at config 4, min_level 2.)
            if (threadIdx.x == 0) {
                const unsigned target = pr.base + (unsigned)(q - P.q_begin);
                if (pr.lo_d)
                    sor_p2p_wait(pr.my_flags + 0, target);
                if (pr.hi_d)
                    sor_p2p_wait(pr.my_flags + 1, target);
                __threadfence_system();
            }
            __syncthreads();
        }
        const SorWave w = sor_wave(P, tb, q);
        int item = blockIdx.x * wpb + warp;
        if (item < w.items) {
            SorLoc cur = sor_locate(P, tb, q, w, item, lane);
            for (;;) {
                const int next = item + stride;
                const bool more = next < w.items;
                SorLoc nxt;
                if (more)
                    nxt = sor_locate(P, tb, q, w, next, lane);
                sor_process<ST, C, true>(P, cur, kb, ke, &pr);
                if (!more)
                    break;
                cur = nxt;
                item = next;
            }
        }
        ++gen;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();                  // this CTA's local and peer stores, before it counts as arrived
            atomicAdd(bar, 1u);
            while (*((volatile unsigned*)bar) < gen * gridDim.x) {
            }
            __threadfence();
            if (blockIdx.x == 0) {
                const unsigned done = pr.base + (unsigned)(q - P.q_begin) + 1u;
                __threadfence_system();
                if (pr.lo_flag)
                    *((volatile unsigned*)pr.lo_flag) = done;
                if (pr.hi_flag)
                    *((volatile unsigned*)pr.hi_flag) = done;
            }
        }
        __syncthreads();
    }
    // do not leave while a neighbour may still be storing its last wave into this GPU's arrays
    if (threadIdx.x == 0 && P.q_end > P.q_begin) {
        const unsigned target = pr.base + (unsigned)(P.q_end - P.q_begin);
        if (pr.lo_d)
            sor_p2p_wait(pr.my_flags + 0, target);
        if (pr.hi_d)
            sor_p2p_wait(pr.my_flags + 1, target);
        __threadfence_system();
    }
}

template <class ST, int C>
inline void sor_run_p2p(Device& dev, const SorParams<ST>& P, unsigned* bar, int peak_items, int kb, int ke,
                        const SorPeers<ST>& pr)
{
    const size_t smem = (size_t)(2 * P.g.S + 1) * sizeof(int32_t);
    const int tabs_in_smem = smem <= 40 * 1024;
    const size_t dyn = tabs_in_smem ? smem : 0;
    int per_sm = 0;
    FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront_p2p<ST, C>, FR3D_SOR_THREADS, dyn));
    FR3D_REQUIRE(per_sm >= 1, "the z-slab solver kernel does not fit on an SM");
    const int wpb = FR3D_SOR_THREADS / 32;
    int64_t want = ((int64_t)peak_items + wpb - 1) / wpb;
    int grid = dev.sm_count * per_sm;
    if (want < grid)
        grid = (int)(want < 1 ? 1 : want);
    FR3D_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), dev.stream));
    SorParams<ST> Pc = P;
    SorPeers<ST> prc = pr;
    int tis = tabs_in_smem;
    void* args[] = {(void*)&Pc, (void*)&bar, (void*)&tis, (void*)&kb, (void*)&ke, (void*)&prc};
    dev.span_begin("fr3d_sor_wavefront_p2p");
    FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront_p2p<ST, C>, dim3(grid), dim3(FR3D_SOR_THREADS), args,
                                          dyn, dev.stream));
    dev.span_end();
    dev.launches++;
}

template <class ST>
inline void sor_run_p2p_any(Device& dev, const SorParams<ST>& P, unsigned* bar, const int32_t* pe_host, int kb, int ke,
                            const SorPeers<ST>& pr)
{
    const int peak = sor_peak_items(P.g.S, P.T, P.B, P.fg, pe_host, 0);
    switch (P.C) {
    case 1: sor_run_p2p<ST, 1>(dev, P, bar, peak, kb, ke, pr); break;
    case 2: sor_run_p2p<ST, 2>(dev, P, bar, peak, kb, ke, pr); break;
    case 3: sor_run_p2p<ST, 3>(dev, P, bar, peak, kb, ke, pr); break;
    case 4: sor_run_p2p<ST, 4>(dev, P, bar, peak, kb, ke, pr); break;
    default: FR3D_THROW(FR3D_ERR_ARG, "unsupported channel count %d", P.C);
    }
}

// ------------------------------------------------------------------------------------------------
// STAGED wavefront kernel (default for a_smooth == 1, full-volume solves).
//
// The plain wavefront kernel above is latency-bound: a warp issues the ~20 loads of one item, waits for
// DRAM, computes, stores, and only then starts the next item (ncu, round 1: 25 % warps active, 48-57 %
// long-scoreboard stalls, 3.9 TB/s).  Here every warp owns a ring of NS shared-memory stages that the TMA
// engine fills with cp.async.bulk (SASS: UBLKCP) while the warp computes: the streamed inputs of an item
// -- the chunk's neighbour table (768 B), the pre-combined system (2304 B per frame, chunk-major so that
// it is ONE contiguous run) and the voxels' own increments (512 B per frame) -- are in flight NS items
// ahead, signalled per stage by an mbarrier (complete_tx).  The sequence of items a warp will process is a
// pure function of (block, warp, wave), so the prefetch runs ACROSS the grid barrier into the next wave:
// what an item of wave q+1 streams (its system, its own value of sweep t-1) was last written in wave q-1
// and is final once wave q has started.  Only the six neighbour increments (written in wave q) must wait
// for the barrier; they are gathered with ordinary loads one item ahead of the arithmetic (LOOKAHEAD).
// psi-refresh items (one sweep in `lag`) stream the neighbour table and the own value and read J directly.
#ifndef FR3D_STG_FG
#define FR3D_STG_FG 1      /* frames per warp item */
#endif
#ifndef FR3D_STG_LOOK
#define FR3D_STG_LOOK 1    /* gather the neighbour increments of item i+1 before computing item i */
#endif
#ifndef FR3D_STG_MINB
#define FR3D_STG_MINB 2    /* resident CTAs per SM the register allocation must allow */
#endif
#ifndef FR3D_STG_NS
#define FR3D_STG_NS 3      /* default stages per warp (runtime-adjustable: FR3D_OPT_SOR_STAGES) */
#endif
#ifndef FR3D_STG_NBRCA
#define FR3D_STG_NBRCA 0   /* neighbour gathers through L1 (1) or L2 only (0) */
#endif

namespace stg {
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    long long t0 = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (!ok) {
            // a stage that never completes is a bug (byte count / alignment): fail instead of hanging the GPU
            if (t0 == 0)
                t0 = clock64();
            else if (clock64() - t0 > 8000000000LL)
                __trap();
        }
    } while (!ok);
}
// global -> shared bulk copy (TMA engine, no tensor map); bytes, both addresses: multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order this thread's generic-proxy accesses with async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
} // namespace stg

template <class ST, int FG>
struct SorStage {
    static constexpr int kNbr = 6 * 32 * 4;
    static constexpr int kAB = 9 * 32 * 8;
    static constexpr int kOwn = 32 * (int)sizeof(Vec4<ST>);
    static constexpr int kFrame = kAB + kOwn;
    static constexpr int kBytes = kNbr + FG * kFrame;
};

// Warp-uniform description of one item: the chunk (first slot), its frames, whether the sweep refreshes psi.
struct SorItem {
    int32_t a0;
    int b0, nb;
    bool refresh;
};
template <class ST, int FG>
__device__ __forceinline__ SorItem sor_item(const SorParams<ST>& P, const SorTabs& tb, int q, const SorWave& w, int item)
{
    const int fgi = item / w.chunks;
    const int f = item - fgi * w.chunks;
    int lo = 0, hi = w.nT - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tb.pe[w.s_lo + 2 * mid] - w.base > f)
            hi = mid;
        else
            lo = mid + 1;
    }
    const int s = w.s_lo + 2 * lo;
    const int before = (s >= 2 ? tb.pe[s - 2] : 0) - w.base;
    const int t = P.redblack ? (q >> 1) : ((q - s) >> 1);
    SorItem I;
    I.a0 = tb.start[s] + 32 * (f - before);
    I.refresh = (t % P.lag) == 0;
    I.b0 = fgi * FG;
    I.nb = P.B - I.b0 < FG ? P.B - I.b0 : FG;
    return I;
}

template <class ST, int FG>
struct SorNbrs {
    Vec4<ST> v[FG][6];
    int n0; // < 0: pad slot
};

template <class ST, int C, int FG>
__global__ void __launch_bounds__(FR3D_SOR_THREADS, FR3D_STG_MINB)
fr3d_sor_staged(const SorParams<ST> P, unsigned* bar, int tabs_in_smem, int NS)
{
    typedef SorStage<ST, FG> SG;
    extern __shared__ __align__(128) unsigned char fr3d_stg_smem[];
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* const stages = fr3d_stg_smem + (size_t)warp * NS * SG::kBytes;
    uint64_t* const mb = reinterpret_cast<uint64_t*>(fr3d_stg_smem + (size_t)wpb * NS * SG::kBytes) + warp * NS;
    int32_t* const tabs = reinterpret_cast<int32_t*>(fr3d_stg_smem + (size_t)wpb * NS * (SG::kBytes + 8));
    SorTabs tb{P.g.pe, P.g.start};
    if (tabs_in_smem) {
        const int S = P.g.S;
        for (int i = threadIdx.x; i < S; i += blockDim.x)
            tabs[i] = P.g.pe[i];
        for (int i = threadIdx.x; i <= S; i += blockDim.x)
            tabs[S + i] = P.g.start[i];
        tb.pe = tabs;
        tb.start = tabs + S;
    }
    if (lane == 0) {
        for (int i = 0; i < NS; ++i)
            stg::mbar_init(stg::s32(mb + i), 1);
        stg::fence_barrier_init();
    }
    __syncthreads();

    const int64_t np = P.g.npad;
    const int64_t nch = np >> 5;
    const int first = blockIdx.x * wpb + warp;
    const int stride = gridDim.x * wpb;

    // ---- producer cursor: the next item of this warp's sequence that has not been requested yet
    int pq = P.q_begin, pitem = first, pst = 0;
    bool pvalid = pq < P.q_end;
    SorWave pw;
    if (pvalid)
        pw = sor_wave(P, tb, pq);
    int inflight = 0; // stages requested and not yet released
    auto p_norm = [&]() {
        while (pvalid && pitem >= pw.items) {
            ++pq;
            if (pq >= P.q_end) {
                pvalid = false;
                break;
            }
            pw = sor_wave(P, tb, pq);
            pitem = first;
        }
    };
    p_norm();
    auto produce = [&](int q_now) {
        while (pvalid && inflight < NS && pq <= q_now + 1) {
            const SorItem I = sor_item<ST, FG>(P, tb, pq, pw, pitem);
            if (lane == 0) {
                const uint32_t ba = stg::s32(mb + pst);
                const uint32_t dst = stg::s32(stages + (size_t)pst * SG::kBytes);
                const uint32_t bytes = SG::kNbr + I.nb * SG::kOwn + (I.refresh ? 0 : I.nb * SG::kAB);
                stg::fence_async_shared(); // the stage's previous contents were read through the generic proxy
                stg::mbar_expect_tx(ba, bytes);
                stg::bulk_g2s(dst, P.g.nbr + (int64_t)(I.a0 >> 5) * 192, SG::kNbr, ba);
                for (int e = 0; e < I.nb; ++e) {
                    const int b = I.b0 + e;
                    const uint32_t fd = dst + SG::kNbr + e * SG::kFrame;
                    if (!I.refresh)
                        stg::bulk_g2s(fd, P.AB + ((int64_t)b * nch + (I.a0 >> 5)) * 288, SG::kAB, ba);
                    stg::bulk_g2s(fd + SG::kAB, P.d + (int64_t)b * np + I.a0, SG::kOwn, ba);
                }
            }
            ++inflight;
            pst = pst + 1 == NS ? 0 : pst + 1;
            pitem += stride;
            p_norm();
        }
    };

    // ---- consumer
    int cst = 0;
    uint32_t cph = 0;
    auto gather = [&](const unsigned char* sg, const SorItem& I, SorNbrs<ST, FG>& N) {
        const int32_t* nb = reinterpret_cast<const int32_t*>(sg) + lane;
        const int n0 = nb[0];
        N.n0 = n0;
        const int i0 = n0 < 0 ? I.a0 + lane : n0;
        const int i1 = nb[32], i2 = nb[64], i3 = nb[96], i4 = nb[128], i5 = nb[160];
#pragma unroll
        for (int e = 0; e < FG; ++e) {
            if (e < I.nb) {
                const Vec4<ST>* d = P.d + (int64_t)(I.b0 + e) * np;
                if (FR3D_STG_NBRCA) {
                    N.v[e][0] = ld4_ca(d + i0);
                    N.v[e][1] = ld4_ca(d + i1);
                    N.v[e][2] = ld4_ca(d + i2);
                    N.v[e][3] = ld4_ca(d + i3);
                    N.v[e][4] = ld4_ca(d + i4);
                    N.v[e][5] = ld4_ca(d + i5);
                } else {
                    N.v[e][0] = ld4_cg(d + i0);
                    N.v[e][1] = ld4_cg(d + i1);
                    N.v[e][2] = ld4_cg(d + i2);
                    N.v[e][3] = ld4_cg(d + i3);
                    N.v[e][4] = ld4_cg(d + i4);
                    N.v[e][5] = ld4_cg(d + i5);
                }
            }
        }
    };
    auto compute = [&](const unsigned char* sg, const SorItem& I, const SorNbrs<ST, FG>& N) {
        if (N.n0 < 0)
            return; // pad slot
        const int64_t a = I.a0 + lane;
#pragma unroll
        for (int e = 0; e < FG; ++e) {
            if (e >= I.nb)
                break;
            const int b = I.b0 + e;
            const unsigned char* fr = sg + SG::kNbr + e * SG::kFrame;
            SorIn<ST> r;
            r.own = reinterpret_cast<const Vec4<ST>*>(fr + SG::kAB)[lane];
            r.xm = N.v[e][0];
            r.ym = N.v[e][1];
            r.zm = N.v[e][2];
            r.xp = N.v[e][3];
            r.yp = N.v[e][4];
            r.zp = N.v[e][5];
            if (I.refresh) {
                const double den0 = 2.0 * P.ax + 2.0 * P.ay + 2.0 * P.az;
                r.L = ld4_cg(P.L + (int64_t)b * np + a);
                sor_refresh<C>(P.a_data, P.J + (int64_t)b * C * 10 * np, P.wgt, np, a, (double)r.own.x, (double)r.own.y,
                               (double)r.own.z, den0, r.A);
                r.A[6] -= (double)r.L.x;
                r.A[7] -= (double)r.L.y;
                r.A[8] -= (double)r.L.z;
                r.L.x = r.L.y = r.L.z = (ST)0;
                double* AB = P.AB + sor_ab_at(P, b, a, 0);
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    FR3D_STCG(AB + k * 32, r.A[k]);
            } else {
                r.L.x = r.L.y = r.L.z = (ST)0;
                const double* A = reinterpret_cast<const double*>(fr) + lane;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    r.A[k] = A[k * 32];
            }
            st4_cg(P.d + (int64_t)b * np + a, sor_update(P, r));
        }
    };

    unsigned gen = 0;
    for (int q = P.q_begin; q < P.q_end; ++q) {
        const SorWave w = sor_wave(P, tb, q);
        produce(q);
        int item = first;
        if (item < w.items) {
            SorItem I = sor_item<ST, FG>(P, tb, q, w, item);
            SorNbrs<ST, FG> cur;
            stg::mbar_wait(stg::s32(mb + cst), cph);
            gather(stages + (size_t)cst * SG::kBytes, I, cur);
            for (;;) {
                const int next = item + stride;
                const bool more = next < w.items;
                const int nst = cst + 1 == NS ? 0 : cst + 1;
                const uint32_t nph = nst == 0 ? cph ^ 1u : cph;
                SorItem In = I;
                SorNbrs<ST, FG> nx;
                if (more) {
                    In = sor_item<ST, FG>(P, tb, q, w, next);
                    if (FR3D_STG_LOOK) {
                        stg::mbar_wait(stg::s32(mb + nst), nph);
                        gather(stages + (size_t)nst * SG::kBytes, In, nx);
                    }
                }
                compute(stages + (size_t)cst * SG::kBytes, I, cur);
                __syncwarp();
                --inflight;
                cst = nst;
                cph = nph;
                produce(q);
                if (!more)
                    break;
                if (!FR3D_STG_LOOK) {
                    stg::mbar_wait(stg::s32(mb + cst), cph);
                    gather(stages + (size_t)cst * SG::kBytes, In, nx);
                }
                I = In;
                cur = nx;
                item = next;
            }
        }
        ++gen;
        stg::fence_async_global(); // this wave's stores (generic proxy) before other SMs' bulk copies of them
        fr3d_grid_barrier(bar, gen * gridDim.x);
        stg::fence_async_global();
    }
}

template <class ST, int C>
inline bool sor_run_staged(Device& dev, const SorParams<ST>& Pin, unsigned* bar, const int32_t* pe_host)
{
    constexpr int FG = FR3D_STG_FG;
    typedef SorStage<ST, FG> SG;
    SorParams<ST> P = Pin;
    P.fg = FG;
    const int peak = sor_peak_items(P.g.S, P.T, P.B, P.fg, pe_host, P.redblack);
    const int wpb = FR3D_SOR_THREADS / 32;
    int NS = dev.sor_stages > 0 ? dev.sor_stages : FR3D_STG_NS;
    const size_t tab_bytes = (size_t)(2 * P.g.S + 1) * sizeof(int32_t);
    const int tabs_in_smem = tab_bytes <= 40 * 1024;
    static bool configured = false; // per instantiation
    if (!configured) {
        FR3D_CUDA(cudaFuncSetAttribute(fr3d_sor_staged<ST, C, FG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
        configured = true;
    }
    int per_sm = 0;
    size_t dyn = 0;
    const int want_per_sm = dev.sor_ctas_per_sm > 0 ? dev.sor_ctas_per_sm : FR3D_STG_MINB;
    for (; NS >= 2; --NS) {
        dyn = (size_t)wpb * NS * (SG::kBytes + 8) + (tabs_in_smem ? tab_bytes : 0);
        if (dyn > 227 * 1024)
            continue;
        FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_staged<ST, C, FG>, FR3D_SOR_THREADS, dyn));
        if (per_sm >= want_per_sm || NS == 2)
            break;
    }
    if (per_sm < 1)
        return false; // does not fit: the caller falls back to the direct-load kernel
    if (per_sm > want_per_sm)
        per_sm = want_per_sm;
    int64_t want = ((int64_t)peak + wpb - 1) / wpb;
    int grid = dev.sm_count * per_sm;
    if (want < grid)
        grid = (int)(want < 1 ? 1 : want);
    FR3D_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), dev.stream));
    int tis = tabs_in_smem;
    void* args[] = {(void*)&P, (void*)&bar, (void*)&tis, (void*)&NS};
    dev.span_begin("fr3d_sor_staged");
    FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_staged<ST, C, FG>, dim3(grid), dim3(FR3D_SOR_THREADS), args,
                                          dyn, dev.stream));
    dev.span_end();
    dev.launches++;
    return true;
}

// ------------------------------------------------------------------------------------------------
// TIME-BLOCKED TILE kernel (fr3d_sor_tile.h): persistent cooperative grid, one task per thread block at a time,
// one grid barrier per TILE wave.
#ifndef FR3D_TILE_THREADS
#define FR3D_TILE_THREADS 128
#endif
#ifndef FR3D_TILE_MINB
#define FR3D_TILE_MINB 3
#endif
template <class ST, int C>
__global__ void __launch_bounds__(FR3D_TILE_THREADS, FR3D_TILE_MINB)
fr3d_sor_tiles(const SorParams<ST> P, const SorTileGeom G, unsigned* bar)
{
    extern __shared__ __align__(16) unsigned char fr3d_tile_smem[];
    unsigned gen = 0;
    sor_tile_run_block<ST, C>(P, G, fr3d_tile_smem, (int)blockIdx.x, (int)gridDim.x, SorBlockPar(), [&](int) {
        ++gen;
        fr3d_grid_barrier(bar, gen * gridDim.x);
    });
}

template <class ST, int C>
inline bool sor_run_tiles(Device& dev, const SorParams<ST>& P, unsigned* bar)
{
    const SorTileGeom G = sor_tile_geom_for(dev, P.g.p, P.g.m, P.g.n, P.T, P.lag);
    const size_t dyn = sor_tile_smem<ST>(G);
    if (dyn > 227 * 1024)
        return false;
    static bool configured = false; // per instantiation
    if (!configured) {
        FR3D_CUDA(cudaFuncSetAttribute(fr3d_sor_tiles<ST, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int per_sm = 0;
    FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_tiles<ST, C>, FR3D_TILE_THREADS, dyn));
    if (per_sm < 1)
        return false;
    if (dev.sor_ctas_per_sm > 0 && per_sm > dev.sor_ctas_per_sm)
        per_sm = dev.sor_ctas_per_sm;
    // the widest tile wave bounds the useful grid
    int64_t widest = 0;
    for (int w = 0; w < G.nwaves; ++w) {
        int64_t cnt = 0;
        for (int tau = 0; tau < G.ntau; ++tau) {
            const int e = w - G.D * tau;
            if (e >= 0 && e < G.ne)
                cnt += sor_tile_count3(G, e);
        }
        widest = cnt > widest ? cnt : widest;
    }
    widest *= P.B;
    int grid = dev.sm_count * per_sm;
    if (widest < grid)
        grid = (int)(widest < 1 ? 1 : widest);
    FR3D_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), dev.stream));
    SorParams<ST> Pc = P;
    SorTileGeom Gc = G;
    void* args[] = {(void*)&Pc, (void*)&Gc, (void*)&bar};
    dev.span_begin("fr3d_sor_tiles");
    FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_tiles<ST, C>, dim3(grid), dim3(FR3D_TILE_THREADS), args, dyn,
                                          dev.stream));
    dev.span_end();
    dev.launches++;
#ifdef FR3D_TILE_TIMING
    {
        unsigned long long clk[8];
        FR3D_CUDA(cudaStreamSynchronize(dev.stream));
        FR3D_CUDA(cudaMemcpyFromSymbol(clk, fr3d_tile_clk, sizeof(clk)));
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        FR3D_CUDA(cudaMemcpyToSymbol(fr3d_tile_clk, z, sizeof(z)));
        double tot = 0;
        for (int q = 0; q < 7; ++q)
            tot += (double)clk[q];
        fprintf(stderr,
                "[tile timing] level %dx%dx%d B=%d grid=%d x %d/SM tile %dx%dx%d Tb=%d waves=%d: rowbase %.1f%% box %.1f%% "
                "prepass %.1f%% waves %.1f%% writeback %.1f%% decode %.1f%% barrier %.1f%%  (%.0f kcycles per block)\n",
                P.g.p, P.g.m, P.g.n, P.B, grid, per_sm, G.K, G.J, G.I, G.Tb, G.nwaves, 100 * clk[0] / tot, 100 * clk[1] / tot,
                100 * clk[2] / tot, 100 * clk[3] / tot, 100 * clk[4] / tot, 100 * clk[5] / tot, 100 * clk[6] / tot,
                tot / grid / 1e3);
    }
#endif
    return true;
}

// Nonlinear-smoothness variant: psi and sweep tasks of a wave, one grid barrier per wave.
template <class ST, int C>
__global__ void __launch_bounds__(FR3D_SOR_THREADS, 2) fr3d_sor_wavefront_nl(const SorParams<ST> P, unsigned* bar)
{
    const int nw = sor_nl_num_waves(P);
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned gen = 0;
    for (int q = 0; q < nw; ++q) {
        const SorSet wp = sor_nl_set(P, q), ws = sor_nl_set(P, q - 2);
        const int items = wp.chunks + ws.chunks;
        for (int item = blockIdx.x * wpb + warp; item < items; item += gridDim.x * wpb)
            sor_nl_item<ST, C>(P, q, wp, ws, item, lane);
        ++gen;
        fr3d_grid_barrier(bar, gen * gridDim.x);
    }
}

template <class ST, int C>
inline void sor_run_nl(Device& dev, const SorParams<ST>& P, unsigned* bar)
{
    int per_sm = 0;
    FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront_nl<ST, C>,
                                                            FR3D_SOR_THREADS, 0));
    FR3D_REQUIRE(per_sm >= 1, "SOR kernel does not fit on an SM");
    const int wpb = FR3D_SOR_THREADS / 32;
    // an upper bound of the busiest wave: every chunk of one residue class mod 4, twice
    int64_t want = ((int64_t)P.g.npad / 32 / 2 + wpb) / wpb;
    int grid = dev.sm_count * per_sm;
    if (want < grid)
        grid = (int)(want < 1 ? 1 : want);
    FR3D_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), dev.stream));
    SorParams<ST> Pc = P;
    void* args[] = {(void*)&Pc, (void*)&bar};
    dev.span_begin("fr3d_sor_wavefront_nl");
    FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront_nl<ST, C>, dim3(grid), dim3(FR3D_SOR_THREADS),
                                          args, 0, dev.stream));
    dev.span_end();
    dev.launches++;
}

template <class ST, int C>
inline void sor_run_c(Device& dev, const SorParams<ST>& P, unsigned* bar, int peak_items, const int32_t* pe_host)
{
    if (P.a_smooth != 1.0) {
        sor_run_nl<ST, C>(dev, P, bar);
        return;
    }
    if (sor_tile_applicable(dev, P) && sor_run_tiles<ST, C>(dev, P, bar))
        return;
    if (dev.sor_kernel == 1 && dev.sor_k1 <= 0 && sor_run_staged<ST, C>(dev, P, bar, pe_host))
        return;
    const size_t smem = (size_t)(2 * P.g.S + 1) * sizeof(int32_t);
    const int tabs_in_smem = smem <= 40 * 1024;
    const size_t dyn = tabs_in_smem ? smem : 0;
    const bool slab = dev.sor_k1 > 0;
    FR3D_REQUIRE(!slab || dev.sor_k1 < 32768, "z-slab solve: more than 32767 planes");
    int per_sm = 0;
    int threads = FR3D_SOR_THREADS;
    SorParams<ST> Pc = P;
    const int nwaves = P.q_end - P.q_begin;
    if (slab || nwaves < 0 || 32 + nwaves > FR3D_SOR_BAR_WORDS)
        Pc.sched = 0; // slab solves and solves with more waves than tickets: round-robin only
#ifdef FR3D_SOR_SPLIT_EXPERIMENT
    const bool split = (Pc.sched & 1024) != 0;
#else
    const bool split = false;
#endif
    if (!Pc.peR || Pc.redblack)
        Pc.sched &= ~128; // no refresh-class prefix table / checkerboard waves are homogeneous
    const bool bal = !split && (Pc.sched & 128) != 0;
    const bool tickets = !split && (Pc.sched & 127) != 0;
    if (slab)
        FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront<ST, C, true>,
                                                                FR3D_SOR_THREADS, dyn));
#ifdef FR3D_SOR_SPLIT_EXPERIMENT
    else if (split)
        FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront_split<ST, C>,
                                                                FR3D_SOR_THREADS, dyn));
#endif
    else if (bal || tickets) {
        threads = SorTune<ST>::kDynThreads;
        if (bal)
            FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront_dyn<ST, C, true>, threads, dyn));
        else
            FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront_dyn<ST, C, false>, threads, dyn));
    }
    else
        FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront<ST, C>,
                                                                FR3D_SOR_THREADS, dyn));
    FR3D_REQUIRE(per_sm >= 1, "SOR kernel does not fit on an SM");
    if (dev.sor_ctas_per_sm > 0 && per_sm > dev.sor_ctas_per_sm)
        per_sm = dev.sor_ctas_per_sm;
    // no more CTAs than the busiest wave can use
    const int wpb = threads / 32;
    int64_t want = ((int64_t)peak_items + wpb - 1) / wpb;
    int grid = dev.sm_count * per_sm;
    if (want < grid)
        grid = (int)(want < 1 ? 1 : want);
    FR3D_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned) * (tickets ? 32 + nwaves : 96), dev.stream));
    int tis = tabs_in_smem | (slab ? ((dev.sor_k0 << 1) | (dev.sor_k1 << 16)) : 0);
    void* args[] = {(void*)&Pc, (void*)&bar, (void*)&tis};
    dev.span_begin("fr3d_sor_wavefront");
    if (slab)
        FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront<ST, C, true>, dim3(grid),
                                              dim3(FR3D_SOR_THREADS), args, dyn, dev.stream));
#ifdef FR3D_SOR_SPLIT_EXPERIMENT
    else if (split)
        FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront_split<ST, C>, dim3(grid), dim3(FR3D_SOR_THREADS),
                                              args, dyn, dev.stream));
#endif
    else if (bal)
        FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront_dyn<ST, C, true>, dim3(grid), dim3(threads), args,
                                              dyn, dev.stream));
    else if (tickets)
        FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront_dyn<ST, C, false>, dim3(grid), dim3(threads), args,
                                              dyn, dev.stream));
    else
        FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront<ST, C>, dim3(grid), dim3(FR3D_SOR_THREADS),
                                              args, dyn, dev.stream));
    dev.span_end();
    dev.launches++;
}
#endif

template <class ST>
inline void sor_run(Device& dev, const SorParams<ST>& P, unsigned* bar, const int32_t* pe_host)
{
#ifdef FR3D_EMU
    (void)pe_host;
#define FR3D_SOR_GO(C_) sor_run_c<ST, C_>(dev, P, bar)
#else
    const int peak = sor_peak_items(P.g.S, P.T, P.B, P.fg, pe_host, P.redblack);
#define FR3D_SOR_GO(C_) sor_run_c<ST, C_>(dev, P, bar, peak, pe_host)
#endif
    switch (P.C) {
    case 1: FR3D_SOR_GO(1); break;
    case 2: FR3D_SOR_GO(2); break;
    case 3: FR3D_SOR_GO(3); break;
    case 4: FR3D_SOR_GO(4); break;
    default: FR3D_THROW(FR3D_ERR_ARG, "C=%d outside 1..%d", P.C, FR3D_MAX_CHANNELS);
    }
#undef FR3D_SOR_GO
}

} // namespace fr3d
