// fr3d_xcorr.h -- device side of the rigid cross-correlation pre-alignment
// (util/xcorr_prealignment.py:15-99; executor steps parallelization/sequential_3d.py:89-145).
//
// Everything here works on the two small mean projections of a volume (XY: (Th,Tw), XZ: (Tz,Tw),
// at most a few hundred samples per side), B frames at a time.  The phase correlation
// (skimage.registration.phase_cross_correlation in the reference) needs 2-D DFTs of those planes;
// at these sizes a DFT is two dense complex products with precomputed DFT matrices (built on the
// host), done by the plain float64 kernel below -- ~70 M complex multiply-adds per frame, no
// library dependency.  The scalar bookkeeping between the stages (peak -> refinement window ->
// wrap disambiguation) is host logic in flowreg3d_b200/xcorr.py, as in the reference.
#pragma once
#include "fr3d_kernels.h"

namespace fr3d {

// Mean projections of a single-channel volume (xcorr_prealignment.py:8-13, 44-45, 73-74).
// numpy's mean of a float32 array along a non-contiguous axis adds the slices in order in float32
// and divides in float32 (acc64 == 0); the float64 reference volume is accumulated in float64.
// item = output sample of pxy (B*Y*X of them) followed by the samples of pxz (B*Z*X).
struct CcProjectK {
    const float* vol; // (B,Z,Y,X)
    float* pxy;       // (B,Y,X)
    float* pxz;       // (B,Z,X)
    int B, Z, Y, X, acc64;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t nxy = (int64_t)B * Y * X;
        if (item < nxy) {
            const int x = (int)(item % X);
            const int y = (int)((item / X) % Y);
            const int b = (int)(item / ((int64_t)X * Y));
            const float* p = vol + (int64_t)b * Z * Y * X + (int64_t)y * X + x;
            if (acc64) {
                double s = 0.0;
                for (int z = 0; z < Z; ++z)
                    s += (double)p[(int64_t)z * Y * X];
                pxy[item] = (float)(s / (double)Z);
            } else {
                float s = 0.0f;
                for (int z = 0; z < Z; ++z)
                    s += p[(int64_t)z * Y * X];
                pxy[item] = s / (float)Z;
            }
        } else {
            const int64_t it = item - nxy;
            const int x = (int)(it % X);
            const int z = (int)((it / X) % Z);
            const int b = (int)(it / ((int64_t)X * Z));
            const float* p = vol + ((int64_t)b * Z + z) * Y * X + x;
            if (acc64) {
                double s = 0.0;
                for (int y = 0; y < Y; ++y)
                    s += (double)p[(int64_t)y * X];
                pxz[it] = (float)(s / (double)Y);
            } else {
                float s = 0.0f;
                for (int y = 0; y < Y; ++y)
                    s += p[(int64_t)y * X];
                pxz[it] = s / (float)Y;
            }
        }
    }
};

// Mean of each (H,W) plane (float64 accumulation, rounded to float32); item = plane.
struct CcPlaneMeanK {
    const float* p;
    float* mean;
    int64_t n;
    FR3D_HD void operator()(int64_t b) const
    {
        const float* q = p + b * n;
        double s = 0.0;
        for (int64_t i = 0; i < n; ++i)
            s += (double)q[i];
        mean[b] = (float)(s / (double)n);
    }
};

// Whitening + Hann window in float32 (xcorr_prealignment.py:50-58, 81-89), written as a complex
// float64 plane (im = 0) ready for the DFT products.  item = (b, y, x).
struct CcWindowK {
    const float* p;
    const float* mean;
    const float* hy;
    const float* hx;
    double* out; // (B,H,W,2)
    int H, W;
    FR3D_HD void operator()(int64_t item) const
    {
        const int x = (int)(item % W);
        const int y = (int)((item / W) % H);
        const int64_t b = item / ((int64_t)W * H);
        const float v = p[item] - mean[b];
        const float w = hy[y] * hx[x];
        out[2 * item] = (double)(v * w);
        out[2 * item + 1] = 0.0;
    }
};

// C[b] = A[b] * Bm[b], complex float64, row-major (M,K) x (K,N); a batch stride of 0 shares the
// matrix between the frames.  item = (b, m, n).
struct CcGemmK {
    const double* A;
    const double* Bm;
    double* Cm;
    int64_t a_stride, b_stride; // in complex elements
    int M, N, K;
    FR3D_HD void operator()(int64_t item) const
    {
        const int n = (int)(item % N);
        const int m = (int)((item / N) % M);
        const int64_t b = item / ((int64_t)N * M);
        const double* a = A + 2 * (b * a_stride + (int64_t)m * K);
        const double* q = Bm + 2 * (b * b_stride + n);
        double re = 0.0, im = 0.0;
        for (int k = 0; k < K; ++k) {
            const double ar = a[2 * k], ai = a[2 * k + 1];
            const double br = q[2 * (int64_t)k * N], bi = q[2 * (int64_t)k * N + 1];
            re += ar * br - ai * bi;
            im += ar * bi + ai * br;
        }
        Cm[2 * item] = re;
        Cm[2 * item + 1] = im;
    }
};

// Cross-power spectrum P = Fr * conj(Fm), optionally phase-normalised: P /= max(|P|, 100 eps),
// eps of float32 -- the precision scipy.fft gives the reference's float32 planes.  item = (b, i).
struct CcCrossPowerK {
    const double* Fr; // (n) shared by the frames
    const double* Fm; // (B,n)
    double* P;
    int64_t n;
    int normalize;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t i = item % n;
        const double ar = Fr[2 * i], ai = Fr[2 * i + 1];
        const double br = Fm[2 * item], bi = -Fm[2 * item + 1];
        double re = ar * br - ai * bi, im = ar * bi + ai * br;
        if (normalize) {
            double m = sqrt(re * re + im * im);
            const double floor_ = 100.0 * 1.1920928955078125e-07;
            m = m > floor_ ? m : floor_;
            re /= m;
            im /= m;
        }
        P[2 * item] = re;
        P[2 * item + 1] = im;
    }
};

// Index of the first maximum of |cc| in each plane (numpy.argmax of numpy.abs); item = plane.
struct CcAbsArgmaxK {
    const double* cc; // (B,n) complex
    int64_t* idx;
    int64_t n;
    FR3D_HD void operator()(int64_t b) const
    {
        const double* q = cc + 2 * b * n;
        double best = -1.0;
        int64_t at = 0;
        for (int64_t i = 0; i < n; ++i) {
            const double re = q[2 * i], im = q[2 * i + 1];
            const double m = sqrt(re * re + im * im);
            if (m > best) {
                best = m;
                at = i;
            }
        }
        idx[b] = at;
    }
};

// Periodic cubic B-spline prefilter of one line (scipy.ndimage spline_filter1d, order 3,
// mode="grid-wrap": gain 6, pole sqrt(3)-2, exact periodic initialisation of both recursions).
// item = line; lines of length n with element stride `es`, consecutive lines `ls` apart inside a
// plane of `per_plane` lines, planes `ps` apart.
struct CcSplineWrapK {
    double* c;
    int n;
    int64_t es, ls, ps;
    int per_plane;
    FR3D_HD void operator()(int64_t item) const
    {
        double* p = c + (item / per_plane) * ps + (item % per_plane) * ls;
        const double z = FR3D_SPLINE_POLE;
        if (n < 2)
            return;
        for (int i = 0; i < n; ++i)
            p[i * es] *= 6.0;
        {
            double zi = z, c0 = p[0];
            for (int i = 1; i < n; ++i) {
                c0 += zi * p[(n - i) * es];
                zi *= z;
            }
            p[0] = c0 / (1.0 - zi);
        }
        for (int i = 1; i < n; ++i)
            p[i * es] += z * p[(i - 1) * es];
        {
            double zi = z, cl = p[(n - 1) * es];
            for (int i = 0; i < n - 1; ++i) {
                cl += zi * p[i * es];
                zi *= z;
            }
            p[(n - 1) * es] = cl * (z / (zi - 1.0));
        }
        for (int i = n - 2; i >= 0; --i)
            p[i * es] = z * (p[(i + 1) * es] - p[i * es]);
    }
};

// dst[i] = src[i * stride] (real parts of interleaved complex planes); item = element
struct StridedCopyK {
    const double* src;
    double* dst;
    int stride;
    FR3D_HD void operator()(int64_t i) const { dst[i] = src[i * stride]; }
};

FR3D_HD int cc_wrap(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

// scipy.ndimage.shift(img, shift, mode="grid-wrap", order = 3 | 0) of B planes, each with its own
// shift; float32 result (the reference's planes are float32).  item = (b, y, x).
struct CcWrapShiftK {
    const double* coef;  // prefiltered planes (B,H,W), read by the order-3 frames
    const double* img_c; // the planes themselves, complex (B,H,W,2), read by the order-0 frames
    const double* shift; // (B,2) device: (sy, sx)
    const int* order;    // (B) device: 3 or 0
    double* out;         // (B,H,W), float32-rounded values
    int H, W;
    FR3D_HD void operator()(int64_t item) const
    {
        const int x = (int)(item % W);
        const int y = (int)((item / W) % H);
        const int64_t b = item / ((int64_t)W * H);
        const double* q = coef + b * (int64_t)H * W;
        const double cy = (double)y - shift[2 * b], cx = (double)x - shift[2 * b + 1];
        if (order[b] == 0) {
            const int iy = cc_wrap((int)floor(cy + 0.5), H), ix = cc_wrap((int)floor(cx + 0.5), W);
            out[item] = (double)(float)img_c[2 * (b * (int64_t)H * W + (int64_t)iy * W + ix)];
            return;
        }
        const double fy = floor(cy), fx = floor(cx);
        double wy[4], wx[4];
        bspline3_weights(cy - fy, wy);
        bspline3_weights(cx - fx, wx);
        const int iy = (int)fy - 1, ix = (int)fx - 1;
        double t = 0.0;
        for (int a = 0; a < 4; ++a) {
            const double* row = q + (int64_t)cc_wrap(iy + a, H) * W;
            for (int e = 0; e < 4; ++e)
                t += row[cc_wrap(ix + e, W)] * wy[a] * wx[e];
        }
        out[item] = (double)(float)t;
    }
};

// Sums for the Pearson correlation of the four tiles [0,sy)|[sy,H) x [0,sx)|[sx,W) of the reference
// plane and a shifted plane: n, Sa, Sb, Saa, Sbb, Sab.  item = (b, tile); tile = 2*(y part) + (x part).
struct CcTileSumsK {
    const double* ref;     // (H,W,2) complex windowed reference plane (real part used)
    const double* shifted; // (B,H,W)
    const int* split;      // (B,2) device: (sy, sx)
    double* out;           // (B,4,6)
    int H, W;
    FR3D_HD void operator()(int64_t item) const
    {
        const int tile = (int)(item & 3);
        const int64_t b = item >> 2;
        const int sy = split[2 * b], sx = split[2 * b + 1];
        const int y0 = (tile >> 1) ? sy : 0, y1 = (tile >> 1) ? H : sy;
        const int x0 = (tile & 1) ? sx : 0, x1 = (tile & 1) ? W : sx;
        double n = 0.0, sa = 0.0, sb = 0.0, saa = 0.0, sbb = 0.0, sab = 0.0;
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                const double a = ref[2 * ((int64_t)y * W + x)];
                const double v = shifted[(b * H + y) * (int64_t)W + x];
                n += 1.0;
                sa += a;
                sb += v;
                saa += a * a;
                sbb += v * v;
                sab += a * v;
            }
        double* o = out + item * 6;
        o[0] = n;
        o[1] = sa;
        o[2] = sb;
        o[3] = saa;
        o[4] = sbb;
        o[5] = sab;
    }
};

// ---- block-cooperative versions of the three plane scans (FR3D_OPT_CC_BLOCK_SCANS) ---------------------------
// Two phases (tile-kernel contract, fr3d_common.h): every thread reduces a strided subset into shared memory,
// thread 0 combines the partials in thread order -- deterministic, and for the arg-max identical to the serial scan
// (largest |.|, smallest index among equals).
struct CcPlaneMeanTileK {
    static constexpr int PHASES = 2;
    const float* p;
    float* mean;
    int64_t n;
    FR3D_HD void phase(int ph, int64_t b, int tid, int nthreads, double* sm) const
    {
        if (ph == 0) {
            const float* q = p + b * n;
            double s = 0.0;
            for (int64_t i = tid; i < n; i += nthreads)
                s += (double)q[i];
            sm[tid] = s;
            return;
        }
        if (tid != 0)
            return;
        double s = 0.0;
        for (int k = 0; k < nthreads; ++k)
            s += sm[k];
        mean[b] = (float)(s / (double)n);
    }
};

struct CcAbsArgmaxTileK {
    static constexpr int PHASES = 2;
    const double* cc;
    int64_t* idx;
    int64_t n;
    FR3D_HD void phase(int ph, int64_t b, int tid, int nthreads, double* sm) const
    {
        if (ph == 0) {
            const double* q = cc + 2 * b * n;
            double best = -1.0;
            int64_t at = 0;
            for (int64_t i = tid; i < n; i += nthreads) {
                const double re = q[2 * i], im = q[2 * i + 1];
                const double m = sqrt(re * re + im * im);
                if (m > best) {
                    best = m;
                    at = i;
                }
            }
            sm[2 * tid] = best;
            sm[2 * tid + 1] = (double)at; // exact: plane sizes are far below 2^53
            return;
        }
        if (tid != 0)
            return;
        double best = -1.0;
        int64_t at = 0;
        for (int k = 0; k < nthreads; ++k) {
            const double v = sm[2 * k];
            const int64_t a = (int64_t)sm[2 * k + 1];
            if (v > best || (v == best && v >= 0.0 && a < at)) {
                best = v;
                at = a;
            }
        }
        idx[b] = at;
    }
};

struct CcTileSumsTileK {
    static constexpr int PHASES = 2;
    const double* ref;
    const double* shifted;
    const int* split;
    double* out;
    int H, W;
    FR3D_HD void phase(int ph, int64_t item, int tid, int nthreads, double* sm) const
    {
        const int tile = (int)(item & 3);
        const int64_t b = item >> 2;
        if (ph == 0) {
            const int sy = split[2 * b], sx = split[2 * b + 1];
            const int y0 = (tile >> 1) ? sy : 0, y1 = (tile >> 1) ? H : sy;
            const int x0 = (tile & 1) ? sx : 0, x1 = (tile & 1) ? W : sx;
            const int w = x1 - x0;
            const int64_t cnt = w > 0 && y1 > y0 ? (int64_t)(y1 - y0) * w : 0;
            double n = 0.0, sa = 0.0, sb = 0.0, saa = 0.0, sbb = 0.0, sab = 0.0;
            for (int64_t q = tid; q < cnt; q += nthreads) {
                const int y = y0 + (int)(q / w), x = x0 + (int)(q % w);
                const double a = ref[2 * ((int64_t)y * W + x)];
                const double v = shifted[(b * H + y) * (int64_t)W + x];
                n += 1.0;
                sa += a;
                sb += v;
                saa += a * a;
                sbb += v * v;
                sab += a * v;
            }
            double* o = sm + (size_t)tid * 6;
            o[0] = n;
            o[1] = sa;
            o[2] = sb;
            o[3] = saa;
            o[4] = sbb;
            o[5] = sab;
            return;
        }
        if (tid != 0)
            return;
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < nthreads; ++k)
            for (int e = 0; e < 6; ++e)
                t[e] += sm[(size_t)k * 6 + e];
        for (int e = 0; e < 6; ++e)
            out[item * 6 + e] = t[e];
    }
};

// w_combined = w_init + w_cross (sequential_3d.py:117-121): float32 add of the frame's rigid offset
// to every voxel of the shared initial field.  item = (b, voxel, component).
struct CcRigidFlowK {
    const float* w_init; // (n3)
    const float* rigid;  // (B,3) device
    float* out;          // (B,n3)
    int64_t n3;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t i = item % n3;
        const int64_t b = item / n3;
        out[item] = w_init[i] + rigid[3 * b + (int)(i % 3)];
    }
};

// flow = (w_combined + w_residual).astype(float32) (sequential_3d.py:140-141): float32 field plus the
// float64 residual flow, one rounding.  item = element.
struct CcAddFlowK {
    const float* comb;
    const double* resid;
    float* out;
    FR3D_HD void operator()(int64_t item) const { out[item] = (float)((double)comb[item] + resid[item]); }
};

} // namespace fr3d
