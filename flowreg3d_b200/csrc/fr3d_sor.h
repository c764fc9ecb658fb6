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
// stores it (9 doubles) and uses it for the next `lag` sweeps.
//
// Storage is the rotated-skew layout of fr3d_kernels.h (struct Skew): a hyperplane is a contiguous
// slab segment and the six neighbours sit at fixed offsets in the two adjacent slabs.
#pragma once
#include "fr3d_kernels.h"

namespace fr3d {

struct SorParams {
    int p, m, n, C, B, T, lag;
    double ax, ay, az; // alpha_{x,y,z} / h_{x,y,z}^2
    double a_data[FR3D_MAX_CHANNELS];
    const double* J;   // (B, C, 10, N) skewed: J11,J22,J33,J44,J12,J13,J23,J14,J24,J34
    const double* wgt; // (C, N) skewed, shared by all frames
    const double* L;   // (B, 3, N) skewed: alpha-weighted Laplacian of u, v, w
    double* AB;        // (B, 9, N) skewed: A11,A22,A33,A12,A13,A23,b1,b2,b3
    double* d;         // (B, 3, N) skewed: du, dv, dw (zero-initialised)
};

#define FR3D_SOR_OMEGA 1.95

// psi refresh + pre-combination for one voxel (sweeps with t % lag == 0): reads the 10C tensor
// entries and the C weights, returns the 9 system entries A11,A22,A33,A12,A13,A23,b1,b2,b3.
FR3D_HD void sor_refresh(const SorParams& P, const double* Jb, int64_t N, int64_t a0, double du, double dv,
                         double dw, double* A)
{
#pragma unroll
    for (int q = 0; q < 9; ++q)
        A[q] = 0.0;
#pragma unroll
    for (int c = 0; c < FR3D_MAX_CHANNELS; ++c) {
        if (c < P.C) {
            const double* Jc = Jb + (int64_t)c * 10 * N + a0;
            const double J11 = Jc[0], J22 = Jc[N], J33 = Jc[2 * N], J44 = Jc[3 * N], J12 = Jc[4 * N],
                         J13 = Jc[5 * N], J23 = Jc[6 * N], J14 = Jc[7 * N], J24 = Jc[8 * N], J34 = Jc[9 * N];
            double ww = P.wgt[(int64_t)c * N + a0];
            const double adc = P.a_data[c];
            if (adc != 1.0) {
                double val = J11 * du * du + J22 * dv * dv + J33 * dw * dw + 2.0 * J12 * du * dv +
                             2.0 * J13 * du * dw + 2.0 * J23 * dv * dw + 2.0 * J14 * du + 2.0 * J24 * dv +
                             2.0 * J34 * dw + J44;
                if (val < 0.0)
                    val = 0.0;
                ww *= adc * pow(val + 1e-6, adc - 1.0);
            }
            A[0] += ww * J11;
            A[1] += ww * J22;
            A[2] += ww * J33;
            A[3] += ww * J12;
            A[4] += ww * J13;
            A[5] += ww * J23;
            A[6] += ww * J14;
            A[7] += ww * J24;
            A[8] += ww * J34;
        }
    }
}

// Everything of a (frame b, sweep t, plane k) row that does not depend on j.
struct SorRow {
    int64_t N;
    int64_t base0, basem, basep; // slab*pm + k*m for the voxel's slab and the two adjacent slabs
    double* d;                   // frame's (3, N) increments
    double* AB;                  // frame's (9, N) system
    const double* L;             // frame's (3, N)
    const double* J;             // frame's (C, 10, N)
    int s, k, refresh;
    bool hz0, hz1;
};

FR3D_HD SorRow sor_row(const SorParams& P, int b, int t, int s, int k)
{
    SorRow r;
    const int64_t pm = (int64_t)P.p * P.m;
    r.N = pm * P.n;
    const int cs = s % P.n;
    const int cm = cs == 0 ? P.n - 1 : cs - 1;
    const int cp = cs == P.n - 1 ? 0 : cs + 1;
    const int64_t row = (int64_t)k * P.m;
    r.base0 = cs * pm + row;
    r.basem = cm * pm + row;
    r.basep = cp * pm + row;
    r.d = P.d + (int64_t)b * 3 * r.N;
    r.AB = P.AB + (int64_t)b * 9 * r.N;
    r.L = P.L + (int64_t)b * 3 * r.N;
    r.J = P.J + (int64_t)b * P.C * 10 * r.N;
    r.s = s;
    r.k = k;
    r.refresh = (t % P.lag) == 0;
    r.hz0 = k > 0;
    r.hz1 = k < P.p - 1;
    return r;
}

// One voxel (row r, column j) of sweep t on hyperplane s = k+j+i.  All loads are unconditional (an
// out-of-domain neighbour reads the voxel's own, not yet updated, increment) and are issued before
// the first use, so the memory latency is paid once per voxel.
FR3D_HD void sor_voxel(const SorParams& P, const SorRow& r, int j)
{
    const int i = r.s - r.k - j;
    const int64_t N = r.N;
    const int64_t a0 = r.base0 + j;
    const int64_t am = r.basem + j, ap = r.basep + j;
    // neighbour addresses; minus side = this sweep, plus side = previous sweep
    const int64_t axm = i > 0 ? am : a0, axp = i < P.n - 1 ? ap : a0;
    const int64_t aym = j > 0 ? am - 1 : a0, ayp = j < P.m - 1 ? ap + 1 : a0;
    const int64_t azm = r.hz0 ? am - P.m : a0, azp = r.hz1 ? ap + P.m : a0;
    double* d0 = r.d;
    double* d1 = d0 + N;
    double* d2 = d1 + N;
    const double du = FR3D_LDCG(d0 + a0), dv = FR3D_LDCG(d1 + a0), dw = FR3D_LDCG(d2 + a0);
    double A[9];
    double* AB = r.AB + a0;
    if (!r.refresh) {
#pragma unroll
        for (int q = 0; q < 9; ++q)
            A[q] = FR3D_LDCG(AB + q * N);
    }
    const double uxp = FR3D_LDCG(d0 + axp), uxm = FR3D_LDCG(d0 + axm), uyp = FR3D_LDCG(d0 + ayp),
                 uym = FR3D_LDCG(d0 + aym), uzp = FR3D_LDCG(d0 + azp), uzm = FR3D_LDCG(d0 + azm);
    const double vxp = FR3D_LDCG(d1 + axp), vxm = FR3D_LDCG(d1 + axm), vyp = FR3D_LDCG(d1 + ayp),
                 vym = FR3D_LDCG(d1 + aym), vzp = FR3D_LDCG(d1 + azp), vzm = FR3D_LDCG(d1 + azm);
    const double wxp = FR3D_LDCG(d2 + axp), wxm = FR3D_LDCG(d2 + axm), wyp = FR3D_LDCG(d2 + ayp),
                 wym = FR3D_LDCG(d2 + aym), wzp = FR3D_LDCG(d2 + azp), wzm = FR3D_LDCG(d2 + azm);
    const double* Lb = r.L + a0;
    const double Lu = Lb[0], Lv = Lb[N], Lw = Lb[2 * N];
    if (r.refresh) {
        sor_refresh(P, r.J, N, a0, du, dv, dw, A);
#pragma unroll
        for (int q = 0; q < 9; ++q)
            FR3D_STCG(AB + q * N, A[q]);
    }
    const double den0 = 2.0 * P.ax + 2.0 * P.ay + 2.0 * P.az;
    const double num_u = Lu + P.ax * (uxp + uxm) + P.ay * (uyp + uym) + P.az * (uzp + uzm);
    const double num_v = Lv + P.ax * (vxp + vxm) + P.ay * (vyp + vym) + P.az * (vzp + vzm);
    const double num_w = Lw + P.ax * (wxp + wxm) + P.ay * (wyp + wym) + P.az * (wzp + wzm);
    const double den_u = den0 + A[0], den_v = den0 + A[1], den_w = den0 + A[2];

    const double u1 = den_u != 0.0 ? (num_u - (A[6] + A[3] * dv + A[4] * dw)) / den_u : 0.0;
    const double du_n = (1.0 - FR3D_SOR_OMEGA) * du + FR3D_SOR_OMEGA * u1;
    const double v1 = den_v != 0.0 ? (num_v - (A[7] + A[3] * du_n + A[5] * dw)) / den_v : 0.0;
    const double dv_n = (1.0 - FR3D_SOR_OMEGA) * dv + FR3D_SOR_OMEGA * v1;
    const double w1 = den_w != 0.0 ? (num_w - (A[8] + A[4] * du_n + A[5] * dv_n)) / den_w : 0.0;
    const double dw_n = (1.0 - FR3D_SOR_OMEGA) * dw + FR3D_SOR_OMEGA * w1;
    FR3D_STCG(d0 + a0, du_n);
    FR3D_STCG(d1 + a0, dv_n);
    FR3D_STCG(d2 + a0, dw_n);
}

// Wave bookkeeping shared by the CUDA kernel and the emulation loop.  A wave's work is a list of
// rows (b, t, k); a warp takes a row and walks its valid j range 32 columns at a time.
struct SorWave {
    int tlo, nT;
    int rows;
};
FR3D_HD int sor_num_waves(const SorParams& P) { return (P.p + P.m + P.n - 2) + 2 * (P.T - 1); }
FR3D_HD SorWave sor_wave(const SorParams& P, int q)
{
    const int S = P.p + P.m + P.n - 2;
    int tlo = q - (S - 1);
    tlo = tlo > 0 ? (tlo + 1) / 2 : 0;
    int thi = q / 2;
    if (thi > P.T - 1)
        thi = P.T - 1;
    SorWave w;
    w.tlo = tlo;
    w.nT = thi >= tlo ? thi - tlo + 1 : 0;
    w.rows = P.B * w.nT * P.p;
    return w;
}
// Execute lane `lane` of row `row` of wave q.
FR3D_HD void sor_do_row(const SorParams& P, int q, const SorWave& w, int row, int lane)
{
    const int k = row % P.p;
    int r = row / P.p;
    const int t = w.tlo + r % w.nT;
    const int b = r / w.nT;
    const int s = q - 2 * t;
    int jlo = s - k - (P.n - 1);
    jlo = jlo < 0 ? 0 : jlo;
    int jhi = s - k;
    jhi = jhi > P.m - 1 ? P.m - 1 : jhi;
    if (jlo > jhi)
        return;
    const SorRow rc = sor_row(P, b, t, s, k);
    for (int j = (jlo & ~31) + lane; j <= jhi; j += 32)
        if (j >= jlo)
            sor_voxel(P, rc, j);
}

#ifdef FR3D_EMU
inline void sor_run(Device& dev, const SorParams& P, unsigned*)
{
    const int nw = sor_num_waves(P);
    for (int q = 0; q < nw; ++q) {
        const SorWave w = sor_wave(P, q);
        for (int row = 0; row < w.rows; ++row)
            for (int lane = 0; lane < 32; ++lane)
                sor_do_row(P, q, w, row, lane);
    }
    dev.launches++;
}
#else
#define FR3D_SOR_THREADS 256

__device__ __forceinline__ void fr3d_grid_barrier(unsigned* ctr, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*((volatile unsigned*)ctr) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// Persistent cooperative kernel: all waves of one level solve, one grid barrier per wave.
__global__ void __launch_bounds__(FR3D_SOR_THREADS, 2) fr3d_sor_wavefront(const SorParams P, unsigned* bar)
{
    const int nw = sor_num_waves(P);
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned gen = 0;
    for (int q = 0; q < nw; ++q) {
        const SorWave w = sor_wave(P, q);
        for (int row = blockIdx.x * wpb + warp; row < w.rows; row += gridDim.x * wpb)
            sor_do_row(P, q, w, row, lane);
        ++gen;
        fr3d_grid_barrier(bar, gen * gridDim.x);
    }
}

inline void sor_run(Device& dev, const SorParams& P, unsigned* bar)
{
    int per_sm = 0;
    FR3D_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fr3d_sor_wavefront,
                                                            FR3D_SOR_THREADS, 0));
    FR3D_REQUIRE(per_sm >= 1, "SOR kernel does not fit on an SM");
    FR3D_REQUIRE((int64_t)P.B * P.T * P.p < 2147483647LL, "level too large for 32-bit row ids");
    // no more CTAs than the busiest wave can use
    int64_t peak = 0;
    {
        const int nw = sor_num_waves(P);
        for (int q = 0; q < nw; q += 1) {
            const SorWave w = sor_wave(P, q);
            if (w.rows > peak)
                peak = w.rows;
        }
    }
    const int wpb = FR3D_SOR_THREADS / 32;
    int64_t want = (peak + wpb - 1) / wpb;
    int grid = dev.sm_count * per_sm;
    if (want < grid)
        grid = (int)(want < 1 ? 1 : want);
    FR3D_CUDA(cudaMemsetAsync(bar, 0, sizeof(unsigned), dev.stream));
    SorParams Pc = P;
    void* args[] = {(void*)&Pc, (void*)&bar};
    dev.span_begin("fr3d_sor_wavefront");
    FR3D_CUDA(cudaLaunchCooperativeKernel((void*)fr3d_sor_wavefront, dim3(grid), dim3(FR3D_SOR_THREADS),
                                          args, 0, dev.stream));
    dev.span_end();
    dev.launches++;
}
#endif

} // namespace fr3d
