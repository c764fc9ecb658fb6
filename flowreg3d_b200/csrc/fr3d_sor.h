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

// One voxel of sweep t on hyperplane s = k+j+i.
FR3D_HD void sor_voxel(const SorParams& P, int b, int t, int s, int k, int j)
{
    const int i = s - k - j;
    const int64_t pm = (int64_t)P.p * P.m;
    const int64_t N = pm * P.n;
    const int cs = s % P.n;
    const int cm = (cs + P.n - 1) % P.n;
    const int cp = (cs + 1) % P.n;
    const int64_t o = (int64_t)k * P.m + j;
    const int64_t a0 = cs * pm + o;
    double* d = P.d + (int64_t)b * 3 * N;
    const double du = FR3D_LDCG(d + a0), dv = FR3D_LDCG(d + N + a0), dw = FR3D_LDCG(d + 2 * N + a0);

    double A11, A22, A33, A12, A13, A23, b1, b2, b3;
    double* AB = P.AB + (int64_t)b * 9 * N + a0;
    if (t % P.lag == 0) {
        A11 = A22 = A33 = A12 = A13 = A23 = b1 = b2 = b3 = 0.0;
        for (int c = 0; c < P.C; ++c) {
            const double* Jc = P.J + (((int64_t)b * P.C + c) * 10) * N + a0;
            const double J11 = Jc[0], J22 = Jc[N], J33 = Jc[2 * N], J44 = Jc[3 * N], J12 = Jc[4 * N],
                         J13 = Jc[5 * N], J23 = Jc[6 * N], J14 = Jc[7 * N], J24 = Jc[8 * N],
                         J34 = Jc[9 * N];
            double ww = P.wgt[(int64_t)c * N + a0];
            const double adc = P.a_data[c];
            if (adc != 1.0) {
                double val = J11 * du * du + J22 * dv * dv + J33 * dw * dw + 2.0 * J12 * du * dv +
                             2.0 * J13 * du * dw + 2.0 * J23 * dv * dw + 2.0 * J14 * du +
                             2.0 * J24 * dv + 2.0 * J34 * dw + J44;
                if (val < 0.0)
                    val = 0.0;
                ww *= adc * pow(val + 1e-6, adc - 1.0);
            }
            A11 += ww * J11;
            A22 += ww * J22;
            A33 += ww * J33;
            A12 += ww * J12;
            A13 += ww * J13;
            A23 += ww * J23;
            b1 += ww * J14;
            b2 += ww * J24;
            b3 += ww * J34;
        }
        FR3D_STCG(AB, A11);
        FR3D_STCG(AB + N, A22);
        FR3D_STCG(AB + 2 * N, A33);
        FR3D_STCG(AB + 3 * N, A12);
        FR3D_STCG(AB + 4 * N, A13);
        FR3D_STCG(AB + 5 * N, A23);
        FR3D_STCG(AB + 6 * N, b1);
        FR3D_STCG(AB + 7 * N, b2);
        FR3D_STCG(AB + 8 * N, b3);
    } else {
        A11 = FR3D_LDCG(AB);
        A22 = FR3D_LDCG(AB + N);
        A33 = FR3D_LDCG(AB + 2 * N);
        A12 = FR3D_LDCG(AB + 3 * N);
        A13 = FR3D_LDCG(AB + 4 * N);
        A23 = FR3D_LDCG(AB + 5 * N);
        b1 = FR3D_LDCG(AB + 6 * N);
        b2 = FR3D_LDCG(AB + 7 * N);
        b3 = FR3D_LDCG(AB + 8 * N);
    }

    // neighbour increments: minus side from this sweep, plus side from the previous one
    const int64_t am = cm * pm + o, ap = cp * pm + o;
    const bool hx0 = i > 0, hx1 = i < P.n - 1, hy0 = j > 0, hy1 = j < P.m - 1, hz0 = k > 0,
               hz1 = k < P.p - 1;
    double sx[3], sy[3], sz[3];
    const double own[3] = {du, dv, dw};
    for (int q = 0; q < 3; ++q) {
        const double* dq = d + q * N;
        const double xm = hx0 ? FR3D_LDCG(dq + am) : own[q];
        const double xp = hx1 ? FR3D_LDCG(dq + ap) : own[q];
        const double ym = hy0 ? FR3D_LDCG(dq + am - 1) : own[q];
        const double yp = hy1 ? FR3D_LDCG(dq + ap + 1) : own[q];
        const double zm = hz0 ? FR3D_LDCG(dq + am - P.m) : own[q];
        const double zp = hz1 ? FR3D_LDCG(dq + ap + P.m) : own[q];
        sx[q] = xp + xm;
        sy[q] = yp + ym;
        sz[q] = zp + zm;
    }
    const double* Lb = P.L + (int64_t)b * 3 * N + a0;
    const double den0 = 2.0 * P.ax + 2.0 * P.ay + 2.0 * P.az;
    const double num_u = Lb[0] + P.ax * sx[0] + P.ay * sy[0] + P.az * sz[0];
    const double num_v = Lb[N] + P.ax * sx[1] + P.ay * sy[1] + P.az * sz[1];
    const double num_w = Lb[2 * N] + P.ax * sx[2] + P.ay * sy[2] + P.az * sz[2];
    const double den_u = den0 + A11, den_v = den0 + A22, den_w = den0 + A33;

    const double u1 = den_u != 0.0 ? (num_u - (b1 + A12 * dv + A13 * dw)) / den_u : 0.0;
    const double du_n = (1.0 - FR3D_SOR_OMEGA) * du + FR3D_SOR_OMEGA * u1;
    const double v1 = den_v != 0.0 ? (num_v - (b2 + A12 * du_n + A23 * dw)) / den_v : 0.0;
    const double dv_n = (1.0 - FR3D_SOR_OMEGA) * dv + FR3D_SOR_OMEGA * v1;
    const double w1 = den_w != 0.0 ? (num_w - (b3 + A13 * du_n + A23 * dv_n)) / den_w : 0.0;
    const double dw_n = (1.0 - FR3D_SOR_OMEGA) * dw + FR3D_SOR_OMEGA * w1;
    FR3D_STCG(d + a0, du_n);
    FR3D_STCG(d + N + a0, dv_n);
    FR3D_STCG(d + 2 * N + a0, dw_n);
}

// Wave bookkeeping shared by the CUDA kernel and the emulation loop.
struct SorWave {
    int tlo, nT;
    int64_t items; // warp-items: (b, t, k, j-chunk of 32)
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
    w.items = (int64_t)P.B * w.nT * P.p * ((P.m + 31) / 32);
    return w;
}
// Execute lane `lane` of warp-item `it` of wave q.
FR3D_HD void sor_item(const SorParams& P, int q, const SorWave& w, int64_t it, int lane)
{
    const int chunks = (P.m + 31) / 32;
    const int jc = (int)(it % chunks);
    int64_t r = it / chunks;
    const int k = (int)(r % P.p);
    r /= P.p;
    const int t = w.tlo + (int)(r % w.nT);
    const int b = (int)(r / w.nT);
    const int s = q - 2 * t;
    const int j = jc * 32 + lane;
    const int i = s - k - j;
    if (j < P.m && i >= 0 && i < P.n)
        sor_voxel(P, b, t, s, k, j);
}

#ifdef FR3D_EMU
inline void sor_run(Device& dev, const SorParams& P, unsigned*)
{
    const int nw = sor_num_waves(P);
    for (int q = 0; q < nw; ++q) {
        const SorWave w = sor_wave(P, q);
        for (int64_t it = 0; it < w.items; ++it)
            for (int lane = 0; lane < 32; ++lane)
                sor_item(P, q, w, it, lane);
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
__global__ void __launch_bounds__(FR3D_SOR_THREADS) fr3d_sor_wavefront(const SorParams P, unsigned* bar)
{
    const int nw = sor_num_waves(P);
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned gen = 0;
    for (int q = 0; q < nw; ++q) {
        const SorWave w = sor_wave(P, q);
        for (int64_t it = (int64_t)blockIdx.x * wpb + warp; it < w.items; it += (int64_t)gridDim.x * wpb)
            sor_item(P, q, w, it, lane);
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
    if (per_sm > 2)
        per_sm = 2;
    // no more CTAs than the busiest wave can use
    int64_t peak = 0;
    {
        const int nw = sor_num_waves(P);
        for (int q = 0; q < nw; q += 1) {
            const SorWave w = sor_wave(P, q);
            if (w.items > peak)
                peak = w.items;
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
