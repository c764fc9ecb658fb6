// fr3d_kernels.h -- per-item kernel bodies (functors) of the flowreg3D hot path.
// One CUDA thread executes operator()(item).  Rounding points follow the reference exactly
// (SURVEY.md 7.3-D); the translation unit is compiled with -fmad=false so that no product/sum is
// silently contracted into an FMA.
//
// Reference citations are relative to /root/reference/src/flowreg3d/.
#pragma once
#include "fr3d_common.h"

namespace fr3d {

// Half-sample symmetric reflection  (d c b a | a b c d | d c b a):
// util/resize_util_3D.py:64-73 and scipy.ndimage mode="reflect".
FR3D_HD int reflect_idx(int j, int n)
{
    if (n <= 1)
        return 0;
    while (j < 0 || j >= n)
        j = (j < 0) ? (-j - 1) : (2 * n - 1 - j);
    return j;
}

// Whole-sample symmetric reflection (d c b | a b c d | c b a): scipy.ndimage mode="mirror".
FR3D_HD int mirror_idx(int j, int n)
{
    if (n <= 1)
        return 0;
    while (j < 0 || j >= n)
        j = (j < 0) ? (-j) : (2 * n - 2 - j);
    return j;
}

// ------------------------------------------------------------------------------------------
// Fused Gauss (x) cubic resize, one separable pass (util/resize_util_3D.py:8-50).
// Output iteration space n[0..4] (n[4] fastest); dimension r is the resampled one.
// Each product is rounded to float32, accumulated in float64 in tap order, stored with one
// rounding to float32 (then widened if DstT is double).
// RUN > 1: the thread produces RUN outputs along axis 3, 32 apart (lane + 32u inside a block of 32*RUN, so that
// every access stays coalesced across the warp); only when axis 3 is not the resampled one.  The tap
// look-ups and the index arithmetic are shared by the run.
template <class SrcT, class DstT, int RUN = 1>
struct ResizePassK {
    const SrcT* src;
    DstT* dst;
    int64_t ss[5], ds[5];
    FastDiv fd[5]; // divisors of the item space at fd[1..3] (axis 3 counted in runs; item < 2^32); axis 4 is walked inside the thread
    int r, P, n4;  // n4 = extent of axis 4 (1 unless the innermost axis is a short interleaved one)
    int n3;        // extent of axis 3 (outputs)
    const int32_t* idx;
    const float* wt;
    FR3D_HD void operator()(int64_t item) const
    {
        uint32_t i[4];
        uint32_t e = (uint32_t)item;
        uint32_t lane = 0;
        if (RUN > 1) {
            lane = e & 31u;
            e >>= 5;
        }
        fd[3].divmod(e, e, i[3]);
        fd[2].divmod(e, e, i[2]);
        fd[1].divmod(e, i[0], i[1]);
        if (RUN > 1)
            i[3] = i[3] * (32 * RUN) + lane;
        int64_t so = 0, dof = 0;
        uint32_t ir = 0;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            dof += (int64_t)i[d] * ds[d];
            if (d != r)
                so += (int64_t)i[d] * ss[d];
            else
                ir = i[d];
        }
        const int32_t* ix = idx + (int64_t)ir * P;
        const float* w = wt + (int64_t)ir * P;
        const int64_t sr = ss[r];
        if (RUN == 1) {
            for (int q = 0; q < n4; ++q) {
                const SrcT* sp = src + so + q * ss[4];
                double acc = 0.0;
                for (int p = 0; p < P; ++p) {
                    const float a = (float)sp[(int64_t)ix[p] * sr];
                    const float prod = a * w[p];
                    acc += (double)prod;
                }
                dst[dof + q * ds[4]] = (DstT)(float)acc;
            }
        } else {
            int nrun = 0;
#pragma unroll
            for (int u = 0; u < RUN; ++u)
                nrun += (int)i[3] + 32 * u < n3;
            for (int q = 0; q < n4; ++q) {
                const SrcT* sp = src + so + q * ss[4];
                double acc[RUN];
#pragma unroll
                for (int u = 0; u < RUN; ++u)
                    acc[u] = 0.0;
                for (int p = 0; p < P; ++p) {
                    const SrcT* tp = sp + (int64_t)ix[p] * sr;
                    const float wp = w[p];
#pragma unroll
                    for (int u = 0; u < RUN; ++u) {
                        if (u < nrun) {
                            const float a = (float)tp[(int64_t)(32 * u) * ss[3]];
                            const float prod = a * wp;
                            acc[u] += (double)prod;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < RUN; ++u)
                    if (u < nrun)
                        dst[dof + (int64_t)(32 * u) * ds[3] + q * ds[4]] = (DstT)(float)acc[u];
            }
        }
    }
};

// Resampling along the FASTEST axis (axis 3, the X pass): a thread produces the outputs of RY consecutive rows
// (axis 2) at one output position, so the tap indices and weights of that position are looked up once for RY rows
// (and all interleaved channels) instead of once per output -- the X pass is bound by load instructions (ncu, round 2:
// 24 % of its stall samples are LSU-queue throttling, 42 table + sample loads per output).  Lanes stay on consecutive
// output positions, as in ResizePassK; per output the products are formed and accumulated in the same order.
template <class SrcT, class DstT, int RY>
struct ResizeXRowsK {
    const SrcT* src;
    DstT* dst;
    int64_t ss[5], ds[5];
    FastDiv fd[5]; // fd[2]: row groups (ceil(n2 / RY)); fd[1], fd[3] as in ResizePassK
    int P, n4, n2;
    const int32_t* idx;
    const float* wt;
    FR3D_HD void operator()(int64_t item) const
    {
        uint32_t i[4];
        uint32_t e = (uint32_t)item;
        fd[3].divmod(e, e, i[3]);
        fd[2].divmod(e, e, i[2]);
        fd[1].divmod(e, i[0], i[1]);
        const int y0 = (int)i[2] * RY;
        const int64_t so = (int64_t)i[0] * ss[0] + (int64_t)i[1] * ss[1] + (int64_t)y0 * ss[2];
        const int64_t dof = (int64_t)i[0] * ds[0] + (int64_t)i[1] * ds[1] + (int64_t)y0 * ds[2] + (int64_t)i[3] * ds[3];
        const int32_t* ix = idx + (int64_t)i[3] * P;
        const float* w = wt + (int64_t)i[3] * P;
        const int64_t sr = ss[3];
        int nrow = n2 - y0;
        nrow = nrow > RY ? RY : nrow;
        for (int q = 0; q < n4; ++q) {
            const SrcT* sp = src + so + q * ss[4];
            double acc[RY];
#pragma unroll
            for (int u = 0; u < RY; ++u)
                acc[u] = 0.0;
            for (int p = 0; p < P; ++p) {
                const SrcT* tp = sp + (int64_t)ix[p] * sr;
                const float wp = w[p];
#pragma unroll
                for (int u = 0; u < RY; ++u) {
                    if (u < nrow) {
                        const float a = (float)tp[(int64_t)u * ss[2]];
                        const float prod = a * wp;
                        acc[u] += (double)prod;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RY; ++u)
                if (u < nrow)
                    dst[dof + (int64_t)u * ds[2] + q * ds[4]] = (DstT)(float)acc[u];
        }
    }
};

// ------------------------------------------------------------------------------------------
// Pre-processing (util/image_processing_3D.py:12-162): (x - lo)/den then a separable Gaussian,
// scipy correlate1d symmetric form  t = x[l]*w0 + sum_{j=r..1} (x[l-j] + x[l+j])*w[j]  in float64,
// mode="reflect".  Axis order Z, Y, X as scipy.ndimage.gaussian_filter applies them.
struct PreGauss {
    int r[FR3D_MAX_CHANNELS][3];
    const double* w[FR3D_MAX_CHANNELS][3]; // device
    double lo[FR3D_MAX_CHANNELS], den[FR3D_MAX_CHANNELS];
    int rt[FR3D_MAX_CHANNELS];             // temporal axis (filtered first, across the frames of the call)
    const double* wt[FR3D_MAX_CHANNELS];
};

// T pass of the 4-D filter: raw (B,Z,Y,X,C) of dtype -> normalised, temporally filtered float64 of the same
// layout; item = (b, voxel, c).  Same symmetric summation as the other axes, reflect over the B frames.
struct PreTK {
    const void* raw;
    int dt;
    double* out;
    int B;
    int64_t nvc; // Z*Y*X*C
    int C;
    PreGauss g;
    FR3D_HD double nrm(int64_t o, int b, int c) const
    {
        return (load_as_double(raw, dt, (int64_t)b * nvc + o) - g.lo[c]) / g.den[c];
    }
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t o = item % nvc;
        const int b = (int)(item / nvc);
        const int c = (int)(o % C);
        const int r = g.rt[c];
        const double* w = g.wt[c];
        double t = nrm(o, b, c) * w[0];
        for (int j = r; j >= 1; --j)
            t += (nrm(o, reflect_idx(b - j, B), c) + nrm(o, reflect_idx(b + j, B), c)) * w[j];
        out[item] = t;
    }
};

// Z pass: raw (B,Z,Y,X,C) of dtype -> planar float64 (B,C,Z,Y,X); item = (b,z,y,x,c), c fastest.
struct PreZK {
    const void* raw;
    int dt;
    double* out;
    int B, Z, Y, X, C;
    PreGauss g;
    FR3D_HD double nrm(int64_t base, int z, int c) const
    {
        // base = offset of (b, 0, y, x, c)
        const double v = load_as_double(raw, dt, base + (int64_t)z * Y * X * C);
        return (v - g.lo[c]) / g.den[c];
    }
    FR3D_HD void operator()(int64_t item) const
    {
        const int c = (int)(item % C);
        item /= C;
        const int x = (int)(item % X);
        item /= X;
        const int y = (int)(item % Y);
        item /= Y;
        const int z = (int)(item % Z);
        const int b = (int)(item / Z);
        const int64_t base = (((int64_t)b * Z * Y + y) * X + x) * C + c;
        const int r = g.r[c][0];
        const double* w = g.w[c][0];
        double t = nrm(base, z, c) * w[0];
        for (int j = r; j >= 1; --j)
            t += (nrm(base, reflect_idx(z - j, Z), c) + nrm(base, reflect_idx(z + j, Z), c)) * w[j];
        out[((((int64_t)b * C + c) * Z + z) * Y + y) * X + x] = t;
    }
};

// Z pass, sliding-window form (all channels share the compile-time radius R): one thread walks a
// whole (b, y, x, c) column keeping the 2R+1 normalised samples in registers, so every raw sample is
// read and divided once instead of 2R+1 times.  Same summation order as PreZK.
template <int R>
struct PreZWinK {
    const void* raw;
    int dt;
    double* out;
    int B, Z, Y, X, C;
    PreGauss g;
    FR3D_HD void operator()(int64_t item) const
    {
        const int c = (int)(item % C);
        item /= C;
        const int x = (int)(item % X);
        item /= X;
        const int y = (int)(item % Y);
        const int b = (int)(item / Y);
        const int64_t base = (((int64_t)b * Z * Y + y) * X + x) * C + c;
        const int64_t zs = (int64_t)Y * X * C;
        const double lo = g.lo[c], den = g.den[c];
        double w[R + 1];
#pragma unroll
        for (int j = 0; j <= R; ++j)
            w[j] = g.w[c][0][j];
        double win[2 * R + 1];
#pragma unroll
        for (int j = 0; j <= 2 * R; ++j)
            win[j] = (load_as_double(raw, dt, base + (int64_t)reflect_idx(j - R, Z) * zs) - lo) / den;
        double* o = out + (((int64_t)b * C + c) * Z * Y + y) * X + x;
        const int64_t os = (int64_t)Y * X;
        // (Measured and rejected, round 2: requesting the column's samples six planes ahead as raw bits -- the ncu
        // source view puts half of this kernel's stall samples on the conversion behind its one load per plane --
        // made the kernel slower, 3.4 vs 2.8 ms per 25-frame step: more registers, same latency chain.)
        for (int z = 0; z < Z; ++z) {
            double t = win[R] * w[0];
#pragma unroll
            for (int j = R; j >= 1; --j)
                t += (win[R - j] + win[R + j]) * w[j];
            o[(int64_t)z * os] = t;
#pragma unroll
            for (int j = 0; j < 2 * R; ++j)
                win[j] = win[j + 1];
            if (z + 1 < Z)
                win[2 * R] = (load_as_double(raw, dt, base + (int64_t)reflect_idx(z + 1 + R, Z) * zs) - lo) / den;
        }
    }
};

// Y pass: planar float64 -> planar float64; item = (vol, z, y, x).
struct PreYK {
    const double* in;
    double* out;
    int C, Z, Y, X;
    PreGauss g;
    FR3D_HD void operator()(int64_t item) const
    {
        const int x = (int)(item % X);
        int64_t q = item / X;
        const int y = (int)(q % Y);
        q /= Y; // q = vol*Z + z
        const int c = (int)((q / Z) % C);
        const double* line = in + q * Y * X + x;
        const int r = g.r[c][1];
        const double* w = g.w[c][1];
        double t = line[(int64_t)y * X] * w[0];
        for (int j = r; j >= 1; --j)
            t += (line[(int64_t)reflect_idx(y - j, Y) * X] + line[(int64_t)reflect_idx(y + j, Y) * X]) * w[j];
        out[item] = t;
    }
};

// X pass: planar float64 -> channels-last float32 (B,Z,Y,X,C); item = (b,z,y,x,c), c fastest.
struct PreXK {
    const double* in;
    float* out;
    double* out64; // optional: the unrounded float64 result (what the reference keeps), same layout
    int B, Z, Y, X, C;
    PreGauss g;
    FR3D_HD void operator()(int64_t item) const
    {
        const int c = (int)(item % C);
        int64_t q = item / C;
        const int x = (int)(q % X);
        q /= X; // q = (b*Z + z)*Y + y
        const int64_t zy = q % ((int64_t)Z * Y);
        const int b = (int)(q / ((int64_t)Z * Y));
        const double* line = in + (((int64_t)b * C + c) * Z * Y + zy) * X;
        const int r = g.r[c][2];
        const double* w = g.w[c][2];
        double t = line[x] * w[0];
        for (int j = r; j >= 1; --j)
            t += (line[reflect_idx(x - j, X)] + line[reflect_idx(x + j, X)]) * w[j];
        out[item] = (float)t;
        if (out64)
            out64[item] = t;
    }
};

// Y and X passes fused through shared memory: a block owns a TY x TX output tile of one (b, z)
// plane; per channel it stages the (TY+2Ry) x (TX+2Rx) float64 neighbourhood of the Z-pass output
// (reflect applied on the global coordinates), filters along Y into a second shared buffer (the
// float64 intermediate scipy would store), filters along X, and finally writes the float32 tile of
// all channels interleaved.  Summation order per output identical to PreYK / PreXK.
struct PreYXTileK {
    static constexpr int PHASES = 3 * FR3D_MAX_CHANNELS + 1;
    static constexpr int TY = 16, TX = 64;
    const double* in; // (B, C, Z, Y, X) planar
    float* out;       // (B, Z, Y, X, C)
    double* out64;    // optional: unrounded float64 result, same layout
    int B, Z, Y, X, C;
    int RY, RX;       // halo = max radius over channels
    PreGauss g;
    int tiles_y, tiles_x;
    FR3D_HD void phase(int ph, int64_t blk, int tid, int nthreads, double* sm) const
    {
        const int tx = (int)(blk % tiles_x);
        int64_t q = blk / tiles_x;
        const int ty = (int)(q % tiles_y);
        q /= tiles_y; // b*Z + z
        const int z = (int)(q % Z);
        const int b = (int)(q / Z);
        const int y0 = ty * TY, x0 = tx * TX;
        const int IW = TX + 2 * RX, IH = TY + 2 * RY;
        double* tin = sm;                       // [IH][IW]
        double* mid = tin + IH * IW;            // [TY][IW]
        double* tout = mid + TY * IW;           // [TY][TX][C]
        if (ph == 3 * FR3D_MAX_CHANNELS) {
            const int ny = Y - y0 < TY ? Y - y0 : TY, nx = X - x0 < TX ? X - x0 : TX;
            const int rowlen = nx * C;
            for (int e = tid; e < ny * rowlen; e += nthreads) {
                const int yy = e / rowlen, r = e - yy * rowlen;
                const int64_t o = ((((int64_t)b * Z + z) * Y + (y0 + yy)) * X + x0) * C + r;
                out[o] = (float)tout[yy * TX * C + r];
                if (out64)
                    out64[o] = tout[yy * TX * C + r];
            }
            return;
        }
        const int c = ph / 3, sub = ph % 3;
        if (c >= C)
            return;
        if (sub == 0) {
            const double* plane = in + (((int64_t)b * C + c) * Z + z) * Y * X;
            for (int e = tid; e < IH * IW; e += nthreads) {
                const int yy = e / IW, xx = e - yy * IW;
                tin[e] = plane[(int64_t)reflect_idx(y0 - RY + yy, Y) * X + reflect_idx(x0 - RX + xx, X)];
            }
        } else if (sub == 1) {
            const int r = g.r[c][1];
            const double* w = g.w[c][1];
            for (int e = tid; e < TY * IW; e += nthreads) {
                const int yy = e / IW, xx = e - yy * IW;
                const double* col = tin + (yy + RY) * IW + xx;
                double t = col[0] * w[0];
                for (int j = r; j >= 1; --j)
                    t += (col[-j * IW] + col[j * IW]) * w[j];
                mid[e] = t;
            }
        } else {
            const int r = g.r[c][2];
            const double* w = g.w[c][2];
            for (int e = tid; e < TY * TX; e += nthreads) {
                const int yy = e / TX, xx = e - yy * TX;
                const double* row = mid + yy * IW + xx + RX;
                double t = row[0] * w[0];
                for (int j = r; j >= 1; --j)
                    t += (row[-j] + row[j]) * w[j];
                tout[(yy * TX + xx) * C + c] = t;
            }
        }
    }
};

// Register-window variant of PreYXTileK for the common case that every channel uses the same radius R
// along Y and X: a thread produces a run of 8 outputs from 8 + 2R staged samples held in registers
// (2 shared-memory reads per output instead of 2R + 1).  Same summation order per output.
template <int R>
struct PreYXWinK {
    static constexpr int PHASES = 3 * FR3D_MAX_CHANNELS + 1;
    static constexpr int TY = 16, TX = 64, RUN = 8;
    static constexpr int IW = TX + 2 * R, IH = TY + 2 * R, MW = IW + 1; // MW: padded row of the Y-pass buffer
    const double* in; // (B, C, Z, Y, X) planar
    float* out;       // (B, Z, Y, X, C)
    double* out64;    // optional: unrounded float64 result, same layout
    int B, Z, Y, X, C;
    PreGauss g;
    int tiles_y, tiles_x;
    static size_t smem_bytes(int C_) { return (size_t)(IH * IW + TY * MW + TY * TX * C_) * sizeof(double); }
    FR3D_HD void phase(int ph, int64_t blk, int tid, int nthreads, double* sm) const
    {
        const int tx = (int)(blk % tiles_x);
        int64_t q = blk / tiles_x;
        const int ty = (int)(q % tiles_y);
        q /= tiles_y; // b*Z + z
        const int z = (int)(q % Z);
        const int b = (int)(q / Z);
        const int y0 = ty * TY, x0 = tx * TX;
        double* tin = sm;                       // [IH][IW]
        double* mid = tin + IH * IW;            // [TY][MW]
        double* tout = mid + TY * MW;           // [TY][TX][C]
        if (ph == 3 * FR3D_MAX_CHANNELS) {
            const int ny = Y - y0 < TY ? Y - y0 : TY, nx = X - x0 < TX ? X - x0 : TX;
            const int rowlen = nx * C;
            const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
            for (int yy = warp; yy < ny; yy += nwarps) {
                const int64_t ob = ((((int64_t)b * Z + z) * Y + (y0 + yy)) * X + x0) * C;
                for (int r = lane; r < rowlen; r += 32) {
                    out[ob + r] = (float)tout[yy * TX * C + r];
                    if (out64)
                        out64[ob + r] = tout[yy * TX * C + r];
                }
            }
            return;
        }
        const int c = ph / 3, sub = ph % 3;
        if (c >= C)
            return;
        if (sub == 0) {
            const double* plane = in + (((int64_t)b * C + c) * Z + z) * Y * X;
            const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
            // all loads of a thread are issued before its first store to shared memory: a store waits for its load,
            // and the warp issues in order, so the plain load/store loop serialised up to 9 DRAM latencies per
            // thread (ncu, round 2: 44 % of the kernel's stall samples on that STS)
            constexpr int XR = (IW + 31) / 32;
            for (int yy = warp; yy < IH; yy += 2 * nwarps) {
                double v[2][XR];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int y2 = yy + u * nwarps;
                    const double* row = plane + (int64_t)reflect_idx(y0 - R + (y2 < IH ? y2 : yy), Y) * X;
#pragma unroll
                    for (int k = 0; k < XR; ++k) {
                        const int xx = lane + 32 * k;
                        v[u][k] = row[reflect_idx(x0 - R + (xx < IW ? xx : lane), X)];
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int y2 = yy + u * nwarps;
                    if (y2 < IH) {
#pragma unroll
                        for (int k = 0; k < XR; ++k) {
                            const int xx = lane + 32 * k;
                            if (xx < IW)
                                tin[y2 * IW + xx] = v[u][k];
                        }
                    }
                }
            }
        } else if (sub == 1) {
            // task = (column xx, run of 8 rows)
            for (int task = tid; task < IW * (TY / RUN); task += nthreads) {
                const int xx = task % IW, run = task / IW;
                double w[R + 1], win[RUN + 2 * R];
#pragma unroll
                for (int j = 0; j <= R; ++j)
                    w[j] = g.w[c][1][j];
#pragma unroll
                for (int k = 0; k < RUN + 2 * R; ++k)
                    win[k] = tin[(run * RUN + k) * IW + xx];
#pragma unroll
                for (int o = 0; o < RUN; ++o) {
                    double t = win[o + R] * w[0];
#pragma unroll
                    for (int j = R; j >= 1; --j)
                        t += (win[o + R - j] + win[o + R + j]) * w[j];
                    mid[(run * RUN + o) * MW + xx] = t;
                }
            }
        } else {
            // task = (row yy, run of 8 columns); consecutive threads take consecutive rows (bank spread)
            for (int task = tid; task < TY * (TX / RUN); task += nthreads) {
                const int yy = task % TY, run = task / TY;
                double w[R + 1], win[RUN + 2 * R];
#pragma unroll
                for (int j = 0; j <= R; ++j)
                    w[j] = g.w[c][2][j];
#pragma unroll
                for (int k = 0; k < RUN + 2 * R; ++k)
                    win[k] = mid[yy * MW + run * RUN + k];
#pragma unroll
                for (int o = 0; o < RUN; ++o) {
                    double t = win[o + R] * w[0];
#pragma unroll
                    for (int j = R; j >= 1; --j)
                        t += (win[o + R - j] + win[o + R + j]) * w[j];
                    tout[(yy * TX + run * RUN + o) * C + c] = t;
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Cubic B-spline prefilter = scipy.ndimage.spline_filter(order=3, mode="nearest") applied to the
// volume edge-padded by 12 (scipy/ndimage/_interpolation.py:212-225, ni_splines.c).  One thread per
// line.  The 12-voxel pad is virtual (clamped reads); only coefficients -1..N+1 along each axis are
// kept because the clipped sample positions never reach further (a sample clipped to N-1 touches
// N+1 with a weight that is zero up to rounding).  Coefficient volume layout: planar
// (vol, Z+3, Y+3, X+3) float64, coefficient b stored at slot b+1.
#define FR3D_SPLINE_PAD 12
#define FR3D_SPLINE_EXT 3 /* extra coefficient slots per axis: b = -1, N, N+1 */
#define FR3D_SPLINE_POLE (-0.26794919243112270647) /* sqrt(3) - 2 */

template <class Get, class Put, class GetC>
FR3D_HD void spline_line(int N, const Get& get, const Put& put, const GetC& getc)
{
    // get(b): input sample at clamped position, already scaled by the gain 6.
    // put(b, v) / getc(b): store / reload coefficient b in [-1, N+1].
    const double z = FR3D_SPLINE_POLE;
    const int len = N + 2 * FR3D_SPLINE_PAD;
    const double zn = pow(z, (double)len);
    const double c0 = get(-FR3D_SPLINE_PAD);
    double acc = c0 + zn * get(N + FR3D_SPLINE_PAD - 1);
    double zi = z;
    const int K = len - 1 < 48 ? len - 1 : 48; // |z|^48 ~ 4e-28: the remaining terms are below 1 ulp
    for (int i = 1; i <= K; ++i) {
        acc += zi * (get(i - FR3D_SPLINE_PAD) + zn * get(N + FR3D_SPLINE_PAD - 1 - i));
        zi *= z;
    }
    acc *= z / (1.0 - zn * zn);
    acc += c0;
    double prev = acc; // causal coefficient at b = -PAD
    double tail[FR3D_SPLINE_PAD];
    // the Y/X passes run in place: the last input sample is overwritten by coefficient N-1 before
    // the clamped reads b >= N need it, so keep it in a register
    const double xhi = get(N - 1);
    for (int b = -FR3D_SPLINE_PAD + 1; b <= N + FR3D_SPLINE_PAD - 1; ++b) {
        const double cur = (b >= N ? xhi : get(b)) + z * prev;
        if (b >= -1) {
            if (b <= N + 1)
                put(b, cur);
            else
                tail[b - N - 2] = cur;
        }
        prev = cur;
    }
    double cm = prev * (z / (z - 1.0)); // anticausal start at b = N + PAD - 1
    for (int b = N + FR3D_SPLINE_PAD - 2; b >= -1; --b) {
        const double cp = (b > N + 1) ? tail[b - N - 2] : getc(b);
        cm = z * (cm - cp);
        if (b <= N + 1)
            put(b, cm);
    }
}

// Z-axis pass: source volume (any dtype, arbitrary strides) -> coef[vol][a+1][y+1][x+1].
// item = (b, y, x, c) with c fastest (channels-last sources read coalesced).
struct SplineZK {
    const void* src;
    int dt;
    int64_t sb, sc, sz, sy, sx; // source strides (elements)
    double* coef;
    int B, C, Z, Y, X;
    FR3D_HD void operator()(int64_t item) const
    {
        const int c = (int)(item % C);
        item /= C;
        const int x = (int)(item % X);
        item /= X;
        const int y = (int)(item % Y);
        const int b = (int)(item / Y);
        const int64_t so = b * sb + c * sc + y * sy + x * sx;
        const int64_t pl = (int64_t)(Y + 3) * (X + 3);
        double* col = coef + (((int64_t)b * C + c) * (Z + 3)) * pl + (int64_t)(y + 1) * (X + 3) + (x + 1);
        const void* s = src;
        const int d = dt;
        const int64_t zs = sz;
        const int ZZ = Z;
        spline_line(
            Z, [=](int bb) { return 6.0 * load_as_double(s, d, so + (int64_t)clampi(bb, 0, ZZ - 1) * zs); },
            [=](int bb, double v) { col[(int64_t)(bb + 1) * pl] = v; },
            [=](int bb) { return col[(int64_t)(bb + 1) * pl]; });
    }
};

// Y-axis pass, in place: item = (vol, a in [0,Z+3), x in [0,X)).
struct SplineYK {
    double* coef;
    int Z, Y, X;
    FR3D_HD void operator()(int64_t item) const
    {
        const int x = (int)(item % X);
        const int64_t va = item / X; // vol*(Z+3) + a
        const int64_t rowlen = X + 3;
        double* base = coef + va * (int64_t)(Y + 3) * rowlen + (x + 1);
        const int YY = Y;
        spline_line(
            Y, [=](int bb) { return 6.0 * base[(int64_t)(clampi(bb, 0, YY - 1) + 1) * rowlen]; },
            [=](int bb, double v) { base[(int64_t)(bb + 1) * rowlen] = v; },
            [=](int bb) { return base[(int64_t)(bb + 1) * rowlen]; });
    }
};

// X-axis pass, in place: item = (vol, a, b in [0,Y+3)).
struct SplineXK {
    double* coef;
    int X;
    FR3D_HD void operator()(int64_t item) const
    {
        double* base = coef + item * (int64_t)(X + 3);
        const int XX = X;
        spline_line(
            X, [=](int bb) { return 6.0 * base[clampi(bb, 0, XX - 1) + 1]; },
            [=](int bb, double v) { base[bb + 1] = v; }, [=](int bb) { return base[bb + 1]; });
    }
};

// Y / X passes, shared-memory tiled.  A block stages TL whole lines in shared memory (coalesced
// loads), filters them in place and writes them back: one global read and one write per
// coefficient instead of two each, and coalesced for the X pass too.  Each line is cut into NSEG
// segments handled by different threads: the recursions  c+[b] = x[b] + z c+[b-1]  and
// c[b] = z (c[b+1] - c+[b])  forget their start value as |z|^k (|z| = 0.268), so a segment that
// starts from 0 a warm-up length W = 40 samples early reproduces the sequential result to
// |z|^40 ~ 1e-23 relative, far below float64 rounding; segments near the line ends use scipy's
// exact boundary initialisation.  Same recurrences, constants and operation order as spline_line.
// TMA staging of a tile whose lines are contiguous in memory (the X pass): ONE bulk copy global -> shared
// (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP) replaces the per-thread load / store staging loops, and one bulk
// copy shared -> global writes the filtered lines back.  No registers, no LSU instructions, full-line bursts.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void fr3d_tma_load_tile(double* smem_dst, const double* gsrc, uint32_t bytes, uint64_t* mbar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok)
                     : "r"(bar), "r"(0u)
                     : "memory");
    } while (!ok);
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fr3d_tma_store_tile(double* gdst, const double* smem_src, uint32_t bytes)
{
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(smem_src);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the tile was written through the generic proxy
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); // shared memory may be reused / released
}
#endif

struct SplineTileK {
    static constexpr int PHASES = 6;
    double* coef;
    int N;          // samples per line; N+3 coefficient slots (b = -1 .. N+1 at slot b+1)
    int TL, NSEG, W;
    int64_t nlines;
    // global address of slot p of line l: first + (l / per_group)*group_stride + (l % per_group)*line_stride + p*elem_stride
    int64_t per_group, group_stride, line_stride, elem_stride, first;
    int line_fast;  // 1: consecutive threads load consecutive lines (Y pass), 0: consecutive slots (X pass)
    int tma;        // 1: the TL lines of a block are contiguous and 16-byte aligned -> bulk-copy (TMA) staging

    FR3D_HD int64_t gaddr(int64_t l, int p) const
    {
        return first + (l / per_group) * group_stride + (l % per_group) * line_stride + (int64_t)p * elem_stride;
    }
    FR3D_HD int sidx(int line, int p) const { return line_fast ? p * TL + line : line * (N + 3) + p; }
    FR3D_HD int seglen() const { return (N + 2 * FR3D_SPLINE_PAD + NSEG - 1) / NSEG; }

    FR3D_HD void phase(int ph, int64_t blk, int tid, int nthreads, double* sm) const
    {
        const int L3 = N + 3;
        const int64_t l0 = blk * TL;
        int nl = (int)(nlines - l0 < TL ? nlines - l0 : TL);
        double* prevs = sm + (size_t)TL * L3;              // [TL][NSEG]
        double* tails = prevs + (size_t)TL * NSEG;         // [TL][PAD]
        if (ph == 0 || ph == 5) {
#if defined(__CUDA_ARCH__)
            if (tma) {
                // lines l0 .. l0+nl-1 are one contiguous, 16-byte aligned run of nl * (N+3) doubles
                const uint32_t bytes = (uint32_t)((size_t)nl * L3 * sizeof(double));
                if ((bytes & 15u) == 0) {
                    if (tid == 0) {
                        if (ph == 0)
                            fr3d_tma_load_tile(sm, coef + gaddr(l0, 0), bytes, reinterpret_cast<uint64_t*>(tails + (size_t)TL * FR3D_SPLINE_PAD));
                        else
                            fr3d_tma_store_tile(coef + gaddr(l0, 0), sm, bytes);
                    }
                    return;
                }
            }
#endif
            // staging copy global <-> shared: no per-element division, eight independent accesses in
            // flight per thread
            if (line_fast) {
                // TL is a power of two: element e -> (slot p = e / TL, line = e % TL); consecutive
                // threads touch consecutive lines (contiguous in memory)
                int sh = 0;
                while ((1 << sh) < TL)
                    ++sh;
                const int line = tid & (TL - 1);
                if (line >= nl)
                    return;
                double* gp = coef + gaddr(l0 + line, 0);
                const int pstep = nthreads >> sh;
                for (int p0 = tid >> sh; p0 < L3; p0 += 8 * pstep) {
                    double v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int p = p0 + u * pstep;
                        if (p < L3)
                            v[u] = ph == 0 ? gp[(int64_t)p * elem_stride] : sm[p * TL + line];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int p = p0 + u * pstep;
                        if (p < L3) {
                            if (ph == 0)
                                sm[p * TL + line] = v[u];
                            else
                                gp[(int64_t)p * elem_stride] = v[u];
                        }
                    }
                }
            } else {
                // consecutive lanes = consecutive slots of one line, warps stride over lines
                const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
                for (int line = warp; line < nl; line += nwarps) {
                    double* gp = coef + gaddr(l0 + line, 0);
                    double* sp = sm + line * L3;
                    for (int p0 = lane; p0 < L3; p0 += 8 * 32) {
                        double v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int p = p0 + u * 32;
                            if (p < L3)
                                v[u] = ph == 0 ? gp[(int64_t)p * elem_stride] : sp[p];
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int p = p0 + u * 32;
                            if (p < L3) {
                                if (ph == 0)
                                    sp[p] = v[u];
                                else
                                    gp[(int64_t)p * elem_stride] = v[u];
                            }
                        }
                    }
                }
            }
            return;
        }
        if (tid >= TL * NSEG)
            return;
        const int line = tid % TL, seg = tid / TL;
        if (line >= nl)
            return;
        const double z = FR3D_SPLINE_POLE;
        const int blo = -FR3D_SPLINE_PAD, bhi = N + FR3D_SPLINE_PAD - 1;
        const int sl = seglen();
        const int bs = blo + seg * sl;
        int be = bs + sl - 1;
        be = be > bhi ? bhi : be;
        if (bs > bhi)
            return;
        const int NN = N;
        auto S = [&](int b) -> double& { return sm[sidx(line, b + 1)]; };
        auto get = [&](int b) { return 6.0 * sm[sidx(line, clampi(b, 0, NN - 1) + 1)]; };
        double& slot = prevs[line * NSEG + seg];
        double* tail = tails + line * FR3D_SPLINE_PAD;
        if (ph == 1) {
            // value of c+ just before the segment's first sample
            double prev;
            int b;
            if (seg == 0 || bs - W <= blo) {
                const int len = N + 2 * FR3D_SPLINE_PAD;
                const double zn = pow(z, (double)len);
                const double c0 = get(blo);
                double acc = c0 + zn * get(bhi);
                double zi = z;
                const int K = len - 1 < 48 ? len - 1 : 48;
                for (int i = 1; i <= K; ++i) {
                    acc += zi * (get(i - FR3D_SPLINE_PAD) + zn * get(bhi - i));
                    zi *= z;
                }
                acc *= z / (1.0 - zn * zn);
                acc += c0;
                prev = acc; // c+[blo]
                b = blo + 1;
            } else {
                prev = 0.0;
                b = bs - W;
            }
            for (; b < bs; ++b)
                prev = get(b) + z * prev;
            slot = prev;
            if (be == bhi)
                tail[FR3D_SPLINE_PAD - 1] = get(N - 1); // x[N-1], needed after slot N is overwritten
            return;
        }
        if (ph == 2) {
            double prev = slot;
            const double xhi = tail[FR3D_SPLINE_PAD - 1];
            for (int b = (bs > blo + 1 ? bs : blo + 1); b <= be; ++b) {
                const double cur = (b >= N ? xhi : get(b)) + z * prev;
                if (b >= -1) {
                    if (b <= N + 1)
                        S(b) = cur;
                    else
                        tail[b - N - 2] = cur;
                }
                prev = cur;
            }
            if (be == bhi)
                tail[FR3D_SPLINE_PAD - 2] = prev; // c+[bhi] (bhi - N - 2 = PAD - 3 is the last tail slot used above)
            return;
        }
        // c+ at position b after phase 2 (b in [-1, bhi - 1])
        auto cplus = [&](int b) { return b > N + 1 ? tail[b - N - 2] : S(b); };
        if (ph == 3) {
            // value of c just above the segment's last sample
            if (be < -1)
                return;
            double cm;
            int b;
            if (be == bhi || be + W >= N + 1) {
                cm = tails[line * FR3D_SPLINE_PAD + FR3D_SPLINE_PAD - 2] * (z / (z - 1.0)); // c[bhi]
                b = bhi - 1;
            } else {
                cm = 0.0;
                b = be + W;
            }
            for (; b > be; --b)
                cm = z * (cm - cplus(b));
            slot = cm;
            return;
        }
        if (ph == 4) {
            if (be < -1)
                return;
            double cm = slot;
            int b = be < bhi - 1 ? be : bhi - 1;
            const int stop = bs > -1 ? bs : -1;
            for (; b >= stop; --b) {
                cm = z * (cm - cplus(b));
                if (b <= N + 1)
                    S(b) = cm;
            }
            return;
        }
    }
};

// scipy ni_splines.c get_spline_interpolation_weights(order=3); verified bit-exact against
// scipy 1.18.1 (w3 is 1 - w0 - w1 - w2, not t^3/6).
FR3D_HD void bspline3_weights(double t, double* w)
{
    const double y = t, z = 1.0 - t;
    w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
    w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
    w[0] = z * z * z / 6.0;
    w[3] = 1.0 - w[0] - w[1] - w[2];
}

// Backward warp gather (core/optical_flow_3d.py:22-74 + scipy NI_GeometricTransform):
// coordinate = float32(grid + displacement); out-of-volume test on the unclipped float32
// coordinate; clip to [0, N-1]; order-3: 4x4x4 taps of the prefiltered coefficients, order-1: 8 taps
// of the source; accumulate t += ((c*wz)*wy)*wx in (z,y,x) tap order; round once to float32.
// item = (b, z, y, x); loops over channels.
//
// Integer sources: the reference calls map_coordinates without `output`, so scipy allocates the
// result in the INPUT dtype and rounds the interpolated value into it (ni_interpolation.c
// CASE_INTERP_OUT_UINT / _INT: add 0.5 toward +-inf, clip to the dtype range, truncate) before the
// reference widens it to float32.  integer_round() restates that.
FR3D_HD double integer_round(double t, int dt)
{
    double lo, hi;
    switch (dt) {
    case FR3D_U8: lo = 0.0; hi = 255.0; break;
    case FR3D_U16: lo = 0.0; hi = 65535.0; break;
    case FR3D_I16: lo = -32768.0; hi = 32767.0; break;
    case FR3D_I32: lo = -2147483648.0; hi = 2147483647.0; break;
    default: return t;
    }
    if (lo == 0.0)
        t = t > 0.0 ? t + 0.5 : 0.0;
    else
        t = t > 0.0 ? t + 0.5 : t - 0.5;
    t = t > hi ? hi : t;
    t = t < lo ? lo : t;
    return trunc(t);
}

struct WarpGatherK {
    int order;            // 3 or 1
    const double* coef;   // order 3: (B*C, Z+3, Y+3, X+3)
    const void* src;      // order 1: source volume
    int sdt;
    int64_t sb, sc, sz, sy, sx;
    // displacement: either three planar float64 fields per frame (level warp; divided by h) or an
    // interleaved float32 flow (compensation warp; h = 1)
    const double* disp64; // (B,3,Z,Y,X) or null
    const float* disp32;  // (B,Z,Y,X,3) or null
    double hx, hy, hz;
    const void* ref;      // out-of-volume replacement (shared by all frames)
    int rdt;
    int64_t rc, rz, ry, rx;
    float* out;
    int64_t ob, oc, oz, oy, ox;
    int B, C, Z, Y, X;
    // Output tiling: a 256-thread block covers TX x TY x TZ outputs (TX*TY*TZ == 256; TX == 0: one x-run of 256).
    // The 4x4x4 tap boxes of a compact tile overlap in y and z as well as in x, so the block pulls
    // (TY+3)(TZ+3)(TX+3) coefficients through L2 instead of 16 (256+3).
    int TX, TY, TZ, nbx, nby, nbz;

    void set_tile(int tx, int ty, int tz)
    {
        TX = tx;
        TY = ty;
        TZ = tz;
        if (tx > 0) {
            nbx = (X + tx - 1) / tx;
            nby = (Y + ty - 1) / ty;
            nbz = (Z + tz - 1) / tz;
        } else {
            nbx = nby = nbz = 0;
        }
    }
    int64_t items() const
    {
        return TX > 0 ? (int64_t)B * nbz * nby * nbx * 256 : (int64_t)B * Z * Y * X;
    }

    FR3D_HD void operator()(int64_t item) const
    {
        int x, y, z, b;
        if (TX > 0) {
            const int t = (int)(item & 255);
            int64_t blk = item >> 8;
            const int tx = t % TX, tyz = t / TX;
            x = (int)(blk % nbx) * TX + tx;
            blk /= nbx;
            y = (int)(blk % nby) * TY + tyz % TY;
            blk /= nby;
            z = (int)(blk % nbz) * TZ + tyz / TY;
            b = (int)(blk / nbz);
            if (x >= X || y >= Y || z >= Z)
                return;
        } else {
            x = (int)(item % X);
            int64_t q = item / X;
            y = (int)(q % Y);
            q /= Y;
            z = (int)(q % Z);
            b = (int)(q / Z);
        }
        double dx, dy, dz;
        if (disp64) {
            const int64_t nvox = (int64_t)Z * Y * X;
            const int64_t o = ((int64_t)z * Y + y) * X + x;
            const double* d = disp64 + (int64_t)b * 3 * nvox;
            dx = d[o] / hx;
            dy = d[nvox + o] / hy;
            dz = d[2 * nvox + o] / hz;
        } else {
            const float* d = disp32 + ((((int64_t)b * Z + z) * Y + y) * X + x) * 3;
            dx = (double)d[0];
            dy = (double)d[1];
            dz = (double)d[2];
        }
        float mx = (float)((double)x + dx);
        float my = (float)((double)y + dy);
        float mz = (float)((double)z + dz);
        const bool oob = (mx < 0.0f) | (mx >= (float)X) | (my < 0.0f) | (my >= (float)Y) |
                         (mz < 0.0f) | (mz >= (float)Z);
        const int64_t obase = b * ob + z * oz + y * oy + x * ox;
        if (oob) {
            for (int c = 0; c < C; ++c)
                out[obase + c * oc] =
                    (float)load_as_double(ref, rdt, c * rc + z * rz + y * ry + x * rx);
            return;
        }
        // np.clip on the float32 coordinates (NaN cannot occur: displacements are finite)
        mx = fminf(fmaxf(mx, 0.0f), (float)(X - 1));
        my = fminf(fmaxf(my, 0.0f), (float)(Y - 1));
        mz = fminf(fmaxf(mz, 0.0f), (float)(Z - 1));
        const double cx = (double)mx, cy = (double)my, cz = (double)mz;
        const double fx = floor(cx), fy = floor(cy), fz = floor(cz);
        const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
        if (order == 3) {
            double wx[4], wy[4], wz[4];
            bspline3_weights(cx - fx, wx);
            bspline3_weights(cy - fy, wy);
            bspline3_weights(cz - fz, wz);
            // taps i0-1..i0+2 live at slots i0..i0+3 of the (N+3)-long coefficient axes
            const int64_t rowlen = X + 3, pl = (int64_t)(Y + 3) * rowlen;
            for (int c = 0; c < C; ++c) {
                const double* cf = coef + ((int64_t)b * C + c) * (Z + 3) * pl + (int64_t)iz * pl +
                                   (int64_t)iy * rowlen + ix;
                double t = 0.0;
                for (int a = 0; a < 4; ++a)
                    for (int bb = 0; bb < 4; ++bb) {
                        const double* row = cf + a * pl + bb * rowlen;
                        for (int cc = 0; cc < 4; ++cc) {
                            double v = row[cc];
                            v *= wz[a];
                            v *= wy[bb];
                            v *= wx[cc];
                            t += v;
                        }
                    }
                out[obase + c * oc] = (float)integer_round(t, sdt);
            }
        } else {
            const double wx[2] = {1.0 - (cx - fx), cx - fx};
            const double wy[2] = {1.0 - (cy - fy), cy - fy};
            const double wz[2] = {1.0 - (cz - fz), cz - fz};
            for (int c = 0; c < C; ++c) {
                const int64_t so = b * sb + c * sc;
                double t = 0.0;
                for (int a = 0; a < 2; ++a)
                    for (int bb = 0; bb < 2; ++bb)
                        for (int cc = 0; cc < 2; ++cc) {
                            const int zz = clampi(iz + a, 0, Z - 1), yy = clampi(iy + bb, 0, Y - 1),
                                      xx = clampi(ix + cc, 0, X - 1);
                            double v = load_as_double(src, sdt, so + zz * sz + yy * sy + xx * sx);
                            v *= wz[a];
                            v *= wy[bb];
                            v *= wx[cc];
                            t += v;
                        }
                out[obase + c * oc] = (float)integer_round(t, sdt);
            }
        }
    }
};

// Order-3 gather with the channel loop innermost (CH independent accumulation chains per thread: the
// 64-term float64 sum of one channel is a serial DADD chain, so a second channel doubles the work in
// flight at the same occupancy) and the z-plane loop rolled (UNROLL_A == 0: 4*CH..16*CH loads live
// instead of 64: fewer registers, more resident warps -- the kernel is latency-bound).  Same
// reference lines, same arithmetic and tap order per channel as WarpGatherK, bit for bit.
// FACT == 1 (FR3D_OPT_WARP_FACTORED, off by default): the separable sum is factored,
// sum_a wz (sum_b wy (sum_c wx c)), 21 fused multiply-adds per plane and channel instead of 64
// float64 operations.  That changes the association of the 64-term sum: the float32 result differs
// from scipy's in the last bit on a ~1e-7 fraction of the voxels -- inside every tolerance of the
// path but not the bit-equality the default keeps.
template <int CH, int UNROLL_A, int FACT = 0>
struct WarpGatherLeanK {
    WarpGatherK g; // order == 3, g.C == CH

    FR3D_HD void operator()(int64_t item) const
    {
        int x, y, z, b;
        {
            const int t = (int)(item & 255);
            int64_t blk = item >> 8;
            const int tx = t % g.TX, tyz = t / g.TX;
            x = (int)(blk % g.nbx) * g.TX + tx;
            blk /= g.nbx;
            y = (int)(blk % g.nby) * g.TY + tyz % g.TY;
            blk /= g.nby;
            z = (int)(blk % g.nbz) * g.TZ + tyz / g.TY;
            b = (int)(blk / g.nbz);
            if (x >= g.X || y >= g.Y || z >= g.Z)
                return;
        }
        double dx, dy, dz;
        if (g.disp64) {
            const int64_t nvox = (int64_t)g.Z * g.Y * g.X;
            const int64_t o = ((int64_t)z * g.Y + y) * g.X + x;
            const double* d = g.disp64 + (int64_t)b * 3 * nvox;
            dx = d[o] / g.hx;
            dy = d[nvox + o] / g.hy;
            dz = d[2 * nvox + o] / g.hz;
        } else {
            const float* d = g.disp32 + ((((int64_t)b * g.Z + z) * g.Y + y) * g.X + x) * 3;
            dx = (double)d[0];
            dy = (double)d[1];
            dz = (double)d[2];
        }
        float mx = (float)((double)x + dx);
        float my = (float)((double)y + dy);
        float mz = (float)((double)z + dz);
        const bool oob = (mx < 0.0f) | (mx >= (float)g.X) | (my < 0.0f) | (my >= (float)g.Y) | (mz < 0.0f) |
                         (mz >= (float)g.Z);
        const int64_t obase = b * g.ob + z * g.oz + y * g.oy + x * g.ox;
        if (oob) {
            for (int c = 0; c < CH; ++c)
                g.out[obase + c * g.oc] =
                    (float)load_as_double(g.ref, g.rdt, c * g.rc + z * g.rz + y * g.ry + x * g.rx);
            return;
        }
        mx = fminf(fmaxf(mx, 0.0f), (float)(g.X - 1));
        my = fminf(fmaxf(my, 0.0f), (float)(g.Y - 1));
        mz = fminf(fmaxf(mz, 0.0f), (float)(g.Z - 1));
        const double cx = (double)mx, cy = (double)my, cz = (double)mz;
        const double fx = floor(cx), fy = floor(cy), fz = floor(cz);
        double wx[4], wy[4], wz[4];
        bspline3_weights(cx - fx, wx);
        bspline3_weights(cy - fy, wy);
        bspline3_weights(cz - fz, wz);
        const int rowlen = g.X + 3;
        const int64_t pl = (int64_t)(g.Y + 3) * rowlen, cvol = (int64_t)(g.Z + 3) * pl;
        const double* cf = g.coef + (int64_t)b * CH * cvol + (int64_t)(int)fz * pl + (int64_t)(int)fy * rowlen + (int)fx;
        double t[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c)
            t[c] = 0.0;
        if (UNROLL_A) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
                plane(cf + a * pl, cvol, rowlen, wz[a], wy, wx, t);
        } else {
#pragma unroll 1
            for (int a = 0; a < 4; ++a) {
                const double wza = a == 0 ? wz[0] : (a == 1 ? wz[1] : (a == 2 ? wz[2] : wz[3]));
                plane(cf, cvol, rowlen, wza, wy, wx, t);
                cf += pl;
            }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c)
            g.out[obase + c * g.oc] = (float)integer_round(t[c], g.sdt);
    }

    // one z-plane of taps: 4 rows x 4 taps x CH channels, all loads first
    FR3D_HD static void plane(const double* p, int64_t cvol, int rowlen, double wza, const double* wy, const double* wx,
                              double* t)
    {
        double v[CH][4][4];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb)
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    v[c][bb][cc] = p[c * cvol + bb * rowlen + cc];
        if (FACT) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                double s = 0.0;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    double r = v[c][bb][0] * wx[0];
#pragma unroll
                    for (int cc = 1; cc < 4; ++cc)
                        r = fma(v[c][bb][cc], wx[cc], r);
                    s = fma(r, wy[bb], s);
                }
                t[c] = fma(s, wza, t[c]);
            }
        } else {
#pragma unroll
            for (int bb = 0; bb < 4; ++bb)
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        double q = v[c][bb][cc];
                        q *= wza;
                        q *= wy[bb];
                        q *= wx[cc];
                        t[c] += q;
                    }
        }
    }
};

// ------------------------------------------------------------------------------------------
// Motion tensor, gradient constancy (core/optical_flow_3d.py:92-152), one channel at one voxel.
// f1, f2: level images (float32-exact values).  All derivatives are central differences on
// replicate-clamped data.  When f2f32 is set, every quantity numpy derives from the float32 warp
// output alone (its gradient and second differences) is evaluated in float32, as numpy does for a
// float32 array with python-float spacings; everything else is float64.
struct MTGeom {
    int p, m, n;
    double hz, hy, hx;
    int f2f32;
};

struct MTImages {
    const float* f1;
    const float* f2;
    MTGeom g;
    FR3D_HD float at1(int k, int j, int i) const
    {
        return f1[((int64_t)clampi(k, 0, g.p - 1) * g.m + clampi(j, 0, g.m - 1)) * g.n + clampi(i, 0, g.n - 1)];
    }
    FR3D_HD float at2(int k, int j, int i) const
    {
        return f2[((int64_t)clampi(k, 0, g.p - 1) * g.m + clampi(j, 0, g.m - 1)) * g.n + clampi(i, 0, g.n - 1)];
    }
    // first derivative of F = (f1, f2 pair) along axis ax (0=z,1=y,2=x) at a (clamped) voxel:
    // 0.5 * (g1 + g2), g1 in float64, g2 in float32 or float64
    FR3D_HD double d1(int k, int j, int i, int ax) const
    {
        k = clampi(k, 0, g.p - 1);
        j = clampi(j, 0, g.m - 1);
        i = clampi(i, 0, g.n - 1);
        const int dk = ax == 0, dj = ax == 1, di = ax == 2;
        const double h = ax == 0 ? g.hz : (ax == 1 ? g.hy : g.hx);
        const double g1 = ((double)at1(k + dk, j + dj, i + di) - (double)at1(k - dk, j - dj, i - di)) / (2.0 * h);
        double g2;
        if (g.f2f32) {
            const float df = at2(k + dk, j + dj, i + di) - at2(k - dk, j - dj, i - di);
            g2 = (double)(df / (float)(2.0 * h));
        } else {
            g2 = ((double)at2(k + dk, j + dj, i + di) - (double)at2(k - dk, j - dj, i - di)) / (2.0 * h);
        }
        return 0.5 * (g1 + g2);
    }
    FR3D_HD double ft(int k, int j, int i) const { return (double)at2(k, j, i) - (double)at1(k, j, i); }
    // 3-point second difference, averaged over the two images
    FR3D_HD double d2(int k, int j, int i, int ax) const
    {
        const int dk = ax == 0, dj = ax == 1, di = ax == 2;
        const double h = ax == 0 ? g.hz : (ax == 1 ? g.hy : g.hx);
        const double a1 = ((double)at1(k - dk, j - dj, i - di) - 2.0 * (double)at1(k, j, i) +
                           (double)at1(k + dk, j + dj, i + di)) / (h * h);
        double a2;
        if (g.f2f32) {
            const float s = at2(k - dk, j - dj, i - di) - 2.0f * at2(k, j, i) + at2(k + dk, j + dj, i + di);
            a2 = (double)(s / (float)(h * h));
        } else {
            a2 = ((double)at2(k - dk, j - dj, i - di) - 2.0 * (double)at2(k, j, i) +
                  (double)at2(k + dk, j + dj, i + di)) / (h * h);
        }
        return 0.5 * (a1 + a2);
    }
    // J[0..9] = J11,J22,J33,J44,J12,J13,J23,J14,J24,J34
    FR3D_HD void tensor(int k, int j, int i, double* J) const
    {
        const double fxx = d2(k, j, i, 2), fyy = d2(k, j, i, 1), fzz = d2(k, j, i, 0);
        const double fxy = (d1(k, j + 1, i, 2) - d1(k, j - 1, i, 2)) / (2.0 * g.hy);
        const double fxz = (d1(k + 1, j, i, 2) - d1(k - 1, j, i, 2)) / (2.0 * g.hz);
        const double fyz = (d1(k + 1, j, i, 1) - d1(k - 1, j, i, 1)) / (2.0 * g.hz);
        const double fxt = (ft(k, j, clampi(i + 1, 0, g.n - 1)) - ft(k, j, clampi(i - 1, 0, g.n - 1))) / (2.0 * g.hx);
        const double fyt = (ft(k, clampi(j + 1, 0, g.m - 1), i) - ft(k, clampi(j - 1, 0, g.m - 1), i)) / (2.0 * g.hy);
        const double fzt = (ft(clampi(k + 1, 0, g.p - 1), j, i) - ft(clampi(k - 1, 0, g.p - 1), j, i)) / (2.0 * g.hz);
        double s;
        s = sqrt(fxx * fxx + fxy * fxy + fxz * fxz);
        const double rx = 1.0 / (s * s + 1e-6);
        s = sqrt(fxy * fxy + fyy * fyy + fyz * fyz);
        const double ry = 1.0 / (s * s + 1e-6);
        s = sqrt(fxz * fxz + fyz * fyz + fzz * fzz);
        const double rz = 1.0 / (s * s + 1e-6);
        J[0] = rx * (fxx * fxx) + ry * (fxy * fxy) + rz * (fxz * fxz);
        J[1] = rx * (fxy * fxy) + ry * (fyy * fyy) + rz * (fyz * fyz);
        J[2] = rx * (fxz * fxz) + ry * (fyz * fyz) + rz * (fzz * fzz);
        J[3] = rx * (fxt * fxt) + ry * (fyt * fyt) + rz * (fzt * fzt);
        J[4] = rx * fxx * fxy + ry * fxy * fyy + rz * fxz * fyz;
        J[5] = rx * fxx * fxz + ry * fxy * fyz + rz * fxz * fzz;
        J[6] = rx * fxy * fxz + ry * fyy * fyz + rz * fyz * fzz;
        J[7] = rx * fxx * fxt + ry * fxy * fyt + rz * fxz * fzt;
        J[8] = rx * fxy * fxt + ry * fyy * fyt + rz * fyz * fzt;
        J[9] = rx * fxz * fxt + ry * fyz * fyt + rz * fzz * fzt;
    }
};

// ------------------------------------------------------------------------------------------
// Solver storage: HYPERPLANE-MAJOR.  The voxels of hyperplane s = k+j+i of a (p,m,n) level are
// stored contiguously, ordered by (k, j); every hyperplane is padded to a multiple of 32 slots.
//     addr(k,j,i) = rowbase[(k+j+i)*p + k] + j
// (rowbase folds in the hyperplane start, the sizes of the rows k' < k and the first valid j of
// row k; it is built on the host, O(S*p)).  All voxels a wavefront sweep may update together are
// therefore one dense range, and their six neighbours sit in the two adjacent hyperplanes at the
// addresses listed in nbr (see fr3d_sor.h).
struct HPView {
    int p, m, n, S;         // S = p+m+n-2 hyperplanes
    int32_t npad;           // storage slots (multiple of 32)
    const int32_t* rowbase; // (S, p)
    const int32_t* start;   // (S+1): first slot of hyperplane s
    const int32_t* pe;      // (S): 32-slot chunks in hyperplanes s, s-2, s-4, ... (same-parity prefix sum)
    const int32_t* nbr;     // (npad/32, 6, 32) CHUNK-MAJOR: slots of x-, y-, z-, x+, y+, z+ of the 32 slots of a chunk
                            // (own slot if outside; [0] = -1 on pad slots).  One chunk's table is 768 contiguous
                            // bytes = one bulk copy of the staged solver kernel (fr3d_sor.h)
    const int32_t* perm;    // (npad): natural linear index (k*m + j)*n + i, or -1 on pad slots
    FR3D_HD int64_t nvox() const { return (int64_t)p * m * n; }
    FR3D_HD int32_t addr(int k, int j, int i) const { return rowbase[(k + j + i) * p + k] + j; }
    // index of neighbour q (0..5) of slot a in nbr
    static FR3D_HD int64_t nbr_at(int q, int64_t a) { return ((a >> 5) * 6 + q) * 32 + (a & 31); }
};

// pad-slot defaults of the tables; item = slot
struct HPFillK {
    int32_t* nbr;
    int32_t* perm;
    int32_t npad;
    FR3D_HD void operator()(int64_t a) const
    {
        perm[a] = -1;
        nbr[HPView::nbr_at(0, a)] = -1;
        for (int q = 1; q < 6; ++q)
            nbr[HPView::nbr_at(q, a)] = (int32_t)a;
    }
};

// item = natural voxel index
struct HPBuildK {
    HPView g;
    int32_t* nbr;
    int32_t* perm;
    FR3D_HD void operator()(int64_t item) const
    {
        const int i = (int)(item % g.n);
        const int j = (int)((item / g.n) % g.m);
        const int k = (int)(item / ((int64_t)g.n * g.m));
        const int32_t a = g.addr(k, j, i);
        perm[a] = (int32_t)item;
        nbr[HPView::nbr_at(0, a)] = i > 0 ? g.addr(k, j, i - 1) : a;
        nbr[HPView::nbr_at(1, a)] = j > 0 ? g.addr(k, j - 1, i) : a;
        nbr[HPView::nbr_at(2, a)] = k > 0 ? g.addr(k - 1, j, i) : a;
        nbr[HPView::nbr_at(3, a)] = i < g.n - 1 ? g.addr(k, j, i + 1) : a;
        nbr[HPView::nbr_at(4, a)] = j < g.m - 1 ? g.addr(k, j + 1, i) : a;
        nbr[HPView::nbr_at(5, a)] = k < g.p - 1 ? g.addr(k + 1, j, i) : a;
    }
};

// Assemble the level system directly in solver storage: motion tensor J (B,C,10,npad), the constant
// part of the smoothness term L = ax*(u_ip+u_im-2u) + ay*(...) + az*(...) for u,v,w, with replicate
// boundary (the reference's edge-padded ring, core/optical_flow_3d.py:88-89,418-426), and the
// (the increments are zero-initialised by the caller).
template <class ST>
struct AssembleK {
    const float* f1;   // (C, N) planar natural, shared by all frames
    const float* f2;   // (B, C, N) planar natural (the warped moving image)
    const double* uvw; // (B, 3, N) natural
    double* J;         // (B, C, 10, npad) or null
    Vec4<ST>* L;       // (B, npad)
    Vec4<ST>* U;       // (B, npad) u, v, w in solver storage, or null (nonlinear smoothness only)
    HPView hp;
    MTGeom g;
    int B, C;
    double ax, ay, az; // alpha/h^2
    // item = (b, natural voxel index): the 33-point image stencils and the Laplacian read coalesced rows; the
    // results are scattered to the voxel's solver slot (pad slots are cleared by the caller's memset)
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t N = hp.nvox(), np = hp.npad;
        const int64_t nat = item % N;
        const int b = (int)(item / N);
        const int i = (int)(nat % g.n);
        const int j = (int)((nat / g.n) % g.m);
        const int k = (int)(nat / ((int64_t)g.n * g.m));
        const int64_t a = hp.addr(k, j, i);
        if (J) {
            for (int c = 0; c < C; ++c) {
                MTImages im{f1 + (int64_t)c * N, f2 + ((int64_t)b * C + c) * N, g};
                double Jv[10];
                im.tensor(k, j, i, Jv);
                double* Jo = J + (((int64_t)b * C + c) * 10) * np + a;
                for (int q = 0; q < 10; ++q)
                    Jo[q * np] = Jv[q];
            }
        }
        const int km = clampi(k - 1, 0, g.p - 1), kp = clampi(k + 1, 0, g.p - 1);
        const int jm = clampi(j - 1, 0, g.m - 1), jp = clampi(j + 1, 0, g.m - 1);
        const int im_ = clampi(i - 1, 0, g.n - 1), ip = clampi(i + 1, 0, g.n - 1);
        double Lq[3], Uq[3];
        for (int q = 0; q < 3; ++q) {
            const double* f = uvw + ((int64_t)b * 3 + q) * N;
            const double c0 = f[((int64_t)k * g.m + j) * g.n + i];
            Uq[q] = c0;
            const double lx = f[((int64_t)k * g.m + j) * g.n + ip] + f[((int64_t)k * g.m + j) * g.n + im_] - 2.0 * c0;
            const double ly = f[((int64_t)k * g.m + jp) * g.n + i] + f[((int64_t)k * g.m + jm) * g.n + i] - 2.0 * c0;
            const double lz = f[((int64_t)kp * g.m + j) * g.n + i] + f[((int64_t)km * g.m + j) * g.n + i] - 2.0 * c0;
            Lq[q] = ax * lx + ay * ly + az * lz;
        }
        Vec4<ST> Lv;
        Lv.x = (ST)Lq[0];
        Lv.y = (ST)Lq[1];
        Lv.z = (ST)Lq[2];
        set_pad(Lv);
        L[(int64_t)b * np + a] = Lv;
        if (U) {
            Vec4<ST> Uv;
            Uv.x = (ST)Uq[0];
            Uv.y = (ST)Uq[1];
            Uv.z = (ST)Uq[2];
            set_pad(Uv);
            U[(int64_t)b * np + a] = Uv;
        }
    }
};

// Stage-API variant: motion tensor of one channel in natural layout (10, N).
struct MotionTensorK {
    MTImages im;
    double* J;
    FR3D_HD void operator()(int64_t item) const
    {
        const int i = (int)(item % im.g.n);
        const int64_t q = item / im.g.n;
        const int j = (int)(q % im.g.m);
        const int k = (int)(q / im.g.m);
        double Jv[10];
        im.tensor(k, j, i, Jv);
        const int64_t N = (int64_t)im.g.p * im.g.m * im.g.n;
        for (int t = 0; t < 10; ++t)
            J[t * N + item] = Jv[t];
    }
};

// The two other constancy assumptions of the reference, get_motion_tensor_gray (core/optical_flow_3d.py:218-259) and
// get_motion_tensor_cs (:155-215).  Neither is reachable through the reference's driver (get_displacement hard-wires
// the gradient-constancy tensor, :457); they are stage functions here as there, for float64 images, one channel.
//   gray: fx = (d/dx f1 + d/dx f2) / 2 with central differences over the replicate-padded images (numpy.gradient with
//         spacings h at interior points of the padded array), ft = f2 - f1;  J = (fx,fy,fz,ft)(fx,fy,fz,ft)^T.
//   cs:   mean over the 26 neighbours d of  wgt_d * (del_d g)(del_d g)^T  with g = (gx,gy,gz,It), the UNIT-spacing
//         central differences of the replicate-padded f2 and It = f2 - f1, both replicate-extended again,
//         del_d a = a(v + d) - a(v),  wgt_d = eps^4 / (4 (eps^2 + (del_d f2)^2)^3),  eps = 80; neighbours are summed in
//         the reference's order (dz, dy, dx ascending, centre skipped) with its operation order.
struct MotionTensorAltK {
    const double* f1;
    const double* f2;
    double* J;          // (10, N)
    int p, m, n, kind;  // kind 1: gray, 2: cs
    double hz, hy, hx;
    FR3D_HD double at(const double* f, int k, int j, int i) const
    {
        return f[((int64_t)clampi(k, 0, p - 1) * m + clampi(j, 0, m - 1)) * n + clampi(i, 0, n - 1)];
    }
    // central difference of f along ax at a clamped voxel, divided by 2*h
    FR3D_HD double grad(const double* f, int k, int j, int i, int ax, double h) const
    {
        k = clampi(k, 0, p - 1);
        j = clampi(j, 0, m - 1);
        i = clampi(i, 0, n - 1);
        const int dk = ax == 0, dj = ax == 1, di = ax == 2;
        return (at(f, k + dk, j + dj, i + di) - at(f, k - dk, j - dj, i - di)) / (2.0 * h);
    }
    FR3D_HD void operator()(int64_t item) const
    {
        const int i = (int)(item % n);
        const int64_t q = item / n;
        const int j = (int)(q % m);
        const int k = (int)(q / m);
        const int64_t N = (int64_t)p * m * n;
        double Jv[10];
        if (kind == 1) {
            const double fx = 0.5 * (grad(f1, k, j, i, 2, hx) + grad(f2, k, j, i, 2, hx));
            const double fy = 0.5 * (grad(f1, k, j, i, 1, hy) + grad(f2, k, j, i, 1, hy));
            const double fz = 0.5 * (grad(f1, k, j, i, 0, hz) + grad(f2, k, j, i, 0, hz));
            const double ft = at(f2, k, j, i) - at(f1, k, j, i);
            Jv[0] = fx * fx;
            Jv[1] = fy * fy;
            Jv[2] = fz * fz;
            Jv[3] = ft * ft;
            Jv[4] = fx * fy;
            Jv[5] = fx * fz;
            Jv[6] = fy * fz;
            Jv[7] = fx * ft;
            Jv[8] = fy * ft;
            Jv[9] = fz * ft;
        } else {
            const double eps2 = 80.0 * 80.0, eps4 = eps2 * eps2;
            for (int t = 0; t < 10; ++t)
                Jv[t] = 0.0;
            const double c0 = at(f2, k, j, i);
            const double gx0 = grad(f2, k, j, i, 2, 1.0), gy0 = grad(f2, k, j, i, 1, 1.0), gz0 = grad(f2, k, j, i, 0, 1.0);
            const double it0 = c0 - at(f1, k, j, i);
            for (int dz = -1; dz <= 1; ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dz == 0 && dy == 0 && dx == 0)
                            continue;
                        const int kk = k + dz, jj = j + dy, ii = i + dx;
                        const double dIm = at(f2, kk, jj, ii) - c0;
                        const double den = eps2 + dIm * dIm;
                        const double wgt = eps4 / (4.0 * den * den * den);
                        const double dIx = grad(f2, kk, jj, ii, 2, 1.0) - gx0;
                        const double dIy = grad(f2, kk, jj, ii, 1, 1.0) - gy0;
                        const double dIz = grad(f2, kk, jj, ii, 0, 1.0) - gz0;
                        const double dIt = (at(f2, kk, jj, ii) - at(f1, kk, jj, ii)) - it0;
                        Jv[0] += wgt * dIx * dIx;
                        Jv[1] += wgt * dIy * dIy;
                        Jv[2] += wgt * dIz * dIz;
                        Jv[3] += wgt * dIt * dIt;
                        Jv[4] += wgt * dIx * dIy;
                        Jv[5] += wgt * dIx * dIz;
                        Jv[6] += wgt * dIy * dIz;
                        Jv[7] += wgt * dIx * dIt;
                        Jv[8] += wgt * dIy * dIt;
                        Jv[9] += wgt * dIz * dIt;
                    }
            const double invN = 1.0 / 26.0;
            for (int t = 0; t < 10; ++t)
                Jv[t] *= invN;
        }
        for (int t = 0; t < 10; ++t)
            J[t * N + item] = Jv[t];
    }
};

// natural (nvol, N) -> solver storage (nvol, npad), pad slots zero; item = (vol, slot)
template <class SrcT>
struct ToHPK {
    const SrcT* nat;
    double* hpv;
    HPView hp;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t a = item % hp.npad;
        const int64_t vol = item / hp.npad;
        const int32_t o = hp.perm[a];
        hpv[item] = o < 0 ? 0.0 : (double)nat[vol * hp.nvox() + o];
    }
};

// increments in solver storage (B, npad){du,dv,dw} -> natural planar (B, 3, N) float64;
// item = (b, q, natural index)
template <class ST>
struct FromHPK {
    const Vec4<ST>* d;
    double* nat;
    HPView hp;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t N = hp.nvox();
        const int64_t o = item % N;
        const int64_t r = item / N;
        const int q = (int)(r % 3);
        const int64_t b = r / 3;
        const int i = (int)(o % hp.n);
        const int j = (int)((o / hp.n) % hp.m);
        const int k = (int)(o / ((int64_t)hp.n * hp.m));
        const Vec4<ST>& v = d[b * hp.npad + hp.addr(k, j, i)];
        nat[item] = (double)(q == 0 ? v.x : (q == 1 ? v.y : v.z));
    }
};

// Planes [k0, k1) of the increments in solver storage <-> a dense buffer (B, k1-k0, m, n) of 4-vectors
// (dir 0: out of the state, 1: into it): the halo / slab exchange of the z-slab multi-GPU solve.
// item = (b, k - k0, j, i)
template <class ST>
struct SlabPlanesK {
    Vec4<ST>* d;
    Vec4<ST>* ext;
    HPView hp;
    int k0, k1, dir;
    FR3D_HD void operator()(int64_t item) const
    {
        const int i = (int)(item % hp.n);
        const int j = (int)((item / hp.n) % hp.m);
        const int64_t r = item / ((int64_t)hp.n * hp.m);
        const int k = k0 + (int)(r % (k1 - k0));
        const int64_t b = r / (k1 - k0);
        Vec4<ST>* slot = d + b * hp.npad + hp.addr(k, j, i);
        if (dir == 0)
            ext[item] = *slot;
        else
            *slot = ext[item];
    }
};

// The cells of plane k that wave q of the lexicographic schedule updates -- one anti-diagonal j + i = s - k per
// sweep t in flight (s = q - 2t) -- <-> a dense buffer (B, T, m) of 4-vectors (entry (t, j) holds cell
// (k, j, s-k-j); entries without a cell are left alone).  This is all a z-neighbour needs after wave q:
// <= T*m cells instead of the m*n of the whole plane.  item = (b, t, j)
template <class ST>
struct SlabWaveCellsK {
    Vec4<ST>* d;
    Vec4<ST>* ext;
    HPView hp;
    int k, q, T, dir;
    FR3D_HD void operator()(int64_t item) const
    {
        const int j = (int)(item % hp.m);
        const int t = (int)((item / hp.m) % T);
        const int64_t b = item / ((int64_t)hp.m * T);
        const int s = q - 2 * t;
        if (s < 0 || s >= hp.S)
            return;
        const int i = s - k - j;
        if (i < 0 || i >= hp.n)
            return;
        Vec4<ST>* slot = d + b * hp.npad + hp.addr(k, j, i);
        if (dir == 0)
            ext[item] = *slot;
        else
            *slot = ext[item];
    }
};

// ------------------------------------------------------------------------------------------
// 5x5x5 median, scipy.ndimage.median_filter(mode="mirror") (core/optical_flow_3d.py:517-526).
// Exact order statistic of float64 data through float32 keys: rounding to float32 is monotone,
// so rank 62 of the keys is the rounding of rank 62 of the data; the float64 value is then
// recovered among the (almost always single) candidates that round to that key.
// Selection: "forgetful" min/max elimination on a 64-entry working set held in registers.
template <int N>
struct ForgetStep {
    // move the minimum of a[0..N) to a[0] and the maximum to a[N-1]
    static FR3D_HD void minmax(float* a)
    {
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            const float lo = fminf(a[i], a[N - 1 - i]);
            const float hi = fmaxf(a[i], a[N - 1 - i]);
            a[i] = lo;
            a[N - 1 - i] = hi;
        }
        // lows live in a[0 .. (N-1)/2], highs in a[N/2 .. N-1]
#pragma unroll
        for (int i = 1; i <= (N - 1) / 2; ++i) {
            const float lo = fminf(a[0], a[i]);
            const float hi = fmaxf(a[0], a[i]);
            a[0] = lo;
            a[i] = hi;
        }
#pragma unroll
        for (int i = N / 2; i < N - 1; ++i) {
            const float lo = fminf(a[i], a[N - 1]);
            const float hi = fmaxf(a[i], a[N - 1]);
            a[i] = lo;
            a[N - 1] = hi;
        }
    }
    // rest: the 125 - 64 keys not yet in the working set, consumed one per step
    static FR3D_HD float run(float* a, const float* rest)
    {
        minmax(a);
        a[0] = rest[64 - N]; // replace the discarded minimum; the maximum a[N-1] falls off the end
        return ForgetStep<N - 1>::run(a, rest);
    }
};
template <>
struct ForgetStep<3> {
    static FR3D_HD float run(float* a, const float*)
    {
        const float lo = fminf(a[0], a[1]), hi = fmaxf(a[0], a[1]);
        return fmaxf(lo, fminf(hi, a[2]));
    }
};

// Pair version: one thread produces the medians of two x-adjacent voxels (x = 2*ip, 2*ip + 1).
// Their windows share 4 of 5 x-columns = 100 samples.  A shared sample with fewer than 37 or more
// than 62 shared samples below it cannot have exactly 62 of the 125 window samples below it in
// either window, so the 37 smallest and 37 largest shared keys are discarded ONCE (forgetful
// min/max elimination, working set 64 -> 28); each output is then the median (rank 25 of 51) of the
// 26 surviving shared keys and its own 25 keys.  ~3.7k min/max operations per output instead of 6.2k.
// shared-core keys 64..99 and the own-column keys are fetched on demand (they are L1 hits) so that only
// the 64-entry working set lives in registers
struct MedianSrc {
    const double* f;
    const int* zi;
    const int* yi;
    const int* xs;
    int m, n;
    template <int T>
    FR3D_HD float core() const // shared key T in [0, 100): (zy = T / 4, column xs[1 + T % 4])
    {
        return (float)f[((int64_t)zi[(T >> 2) / 5] * m + yi[(T >> 2) % 5]) * n + xs[1 + (T & 3)]];
    }
    template <int T>
    FR3D_HD float own(int xo) const // own key T in [0, 25) of the column with mirrored index xo
    {
        return (float)f[((int64_t)zi[T / 5] * m + yi[T % 5]) * n + xo];
    }
};
template <int N>
struct CorePrune {
    static FR3D_HD void run(float* a, const MedianSrc& s)
    {
        ForgetStep<N>::minmax(a);
        a[0] = s.template core<64 + (64 - N)>(); // the discarded minimum's slot takes the next unseen key
        CorePrune<N - 1>::run(a, s);
    }
};
template <>
struct CorePrune<28> {
    static FR3D_HD void run(float* a, const MedianSrc&) { ForgetStep<28>::minmax(a); } // survivors: a[1..26]
};
template <int N>
struct Forget27 { // median of 51 = 27 in the working set + 24 fetched
    static FR3D_HD float run(float* a, const MedianSrc& s, int xo)
    {
        ForgetStep<N>::minmax(a);
        a[0] = s.template own<1 + (27 - N)>(xo);
        return Forget27<N - 1>::run(a, s, xo);
    }
};
template <>
struct Forget27<3> {
    static FR3D_HD float run(float* a, const MedianSrc&, int)
    {
        const float lo = fminf(a[0], a[1]), hi = fmaxf(a[0], a[1]);
        return fmaxf(lo, fminf(hi, a[2]));
    }
};

struct Median5PairK {
    const double* src; // (nvol, p, m, n)
    double* dst;       // (nvol, p, m, n): dst = (add ? add : 0) + median
    const double* add; // optional (nvol, p, m, n)
    int p, m, n, npair; // npair = (n + 1) / 2
    int k0, kn;         // z range [k0, k0 + kn) of the outputs (the whole volume: 0, p)
    // float64 value of rank 62 among the window samples whose float32 key equals med
    FR3D_HD double recover(const double* f, const int* zi, const int* yi, int i, float med) const
    {
        int xi[5];
#pragma unroll
        for (int d = 0; d < 5; ++d)
            xi[d] = mirror_idx(i + d - 2, n);
        int less = 0, eq = 0;
        double tmin = 0.0, tmax = 0.0;
        for (int t = 0; t < 125; ++t) {
            const double v = f[((int64_t)zi[t / 25] * m + yi[(t / 5) % 5]) * n + xi[t % 5]];
            const float key = (float)v;
            less += key < med;
            if (key == med) {
                tmin = (eq == 0 || v < tmin) ? v : tmin;
                tmax = (eq == 0 || v > tmax) ? v : tmax;
                ++eq;
            }
        }
        if (tmin == tmax)
            return tmin;
        int want = 62 - less;
        double cur = tmin;
        for (;;) {
            int mult = 0;
            double next = tmax;
            for (int t = 0; t < 125; ++t) {
                const double v = f[((int64_t)zi[t / 25] * m + yi[(t / 5) % 5]) * n + xi[t % 5]];
                if ((float)v != med)
                    continue;
                mult += v == cur;
                if (v > cur && v < next)
                    next = v;
            }
            if (want < mult || cur == tmax)
                break;
            want -= mult;
            cur = next;
        }
        return cur;
    }
    // Fast path of recover(): almost always exactly ONE window sample rounds to the median key, and then it is the
    // answer whatever the other samples are.  One pass over the window with hoisted row pointers; the test
    // "(float)v == med" is made in the float64 domain against the closed interval of doubles that can round to med
    // (two DSETP on the otherwise idle fp64 pipe instead of a conversion and two compares on the ALU pipe that the
    // selection saturates) and confirmed exactly for the rare candidates.  Ties fall back to recover().
    FR3D_HD double recover_fast(const double* f, const int* zi, const int* yi, int i, float med) const
    {
        int xi[5];
#pragma unroll
        for (int d = 0; d < 5; ++d)
            xi[d] = mirror_idx(i + d - 2, n);
        // neighbours of med in float32; the doubles that round to med lie between the midpoints (inclusive bound:
        // a superset, confirmed below)
        const float up = nextafterf(med, INFINITY), dn = nextafterf(med, -INFINITY);
        const double hi = 0.5 * ((double)med + (double)up), lo = 0.5 * ((double)med + (double)dn);
        int eq = 0;
        double val = 0.0;
        bool differ = false;
#pragma unroll 1
        for (int a = 0; a < 5; ++a) {
#pragma unroll
            for (int b = 0; b < 5; ++b) {
                const double* row = f + ((int64_t)zi[a] * m + yi[b]) * n;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    const double v = row[xi[c]];
                    if (v >= lo && v <= hi) {
                        if ((float)v == med) {
                            differ = differ || (eq > 0 && v != val);
                            val = v;
                            ++eq;
                        }
                    }
                }
            }
        }
        if (eq >= 1 && !differ)
            return val;
        return recover(f, zi, yi, i, med);
    }
    FR3D_HD void operator()(int64_t item) const
    {
        const int ip = (int)(item % npair);
        int64_t q = item / npair;
        const int j = (int)(q % m);
        q /= m;
        const int k = k0 + (int)(q % kn);
        const int64_t vol = q / kn;
        const int64_t N = (int64_t)p * m * n;
        const double* f = src + vol * N;
        const int i0 = 2 * ip;
        const bool two = i0 + 1 < n;
        int zi[5], yi[5], xs[6]; // xs: mirrored x of columns i0-2 .. i0+3
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            zi[d] = mirror_idx(k + d - 2, p);
            yi[d] = mirror_idx(j + d - 2, m);
        }
#pragma unroll
        for (int d = 0; d < 6; ++d)
            xs[d] = mirror_idx(i0 + d - 2 < n + 2 ? i0 + d - 2 : n + 1, n); // (clamped only for the unused odd tail)
        // shared core: columns xs[1..4] x 25 (z,y) positions; key t -> (zy = t / 4, col = 1 + t % 4)
        const MedianSrc ms{f, zi, yi, xs, m, n};
        float a[64];
        load_core<0>(a, ms);
        CorePrune<64>::run(a, ms);
        const int64_t oA = vol * N + ((int64_t)k * m + j) * n + i0;
#pragma unroll 1
        for (int o = 0; o < (two ? 2 : 1); ++o) {
            // output o (voxel i0 + o): survivors a[1..26] + own column xs[0] / xs[5]
            const int xo = o ? xs[5] : xs[0];
            float wa[27];
#pragma unroll
            for (int t = 0; t < 26; ++t)
                wa[t] = a[1 + t];
            wa[26] = ms.own<0>(xo);
            const float med = Forget27<27>::run(wa, ms, xo);
            const double r = recover_fast(f, zi, yi, i0 + o, med);
            dst[oA + o] = add ? add[oA + o] + r : r;
        }
    }
    template <int T>
    static FR3D_HD void load_core(float* a, const MedianSrc& ms)
    {
        if constexpr (T < 64) {
            a[T] = ms.core<T>();
            load_core<T + 1>(a, ms);
        }
    }
};

// dst = a + b (element-wise, float64) -- levels too small for the median
struct AddK {
    const double* a;
    const double* b;
    double* dst;
    FR3D_HD void operator()(int64_t i) const { dst[i] = a[i] + b[i]; }
};

// z-slab [k0, k0+kn) of nvol planar volumes (p, plane) <-> packed (nvol, kn, plane); item = packed index
struct SlabCopyK {
    double* vol;
    double* packed;
    int64_t plane; // m * n
    int p, k0, kn, to_packed;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t per = (int64_t)kn * plane;
        const int64_t v = item / per, o = item % per;
        double* a = vol + v * (int64_t)p * plane + (int64_t)k0 * plane + o;
        if (to_packed)
            packed[item] = *a;
        else
            *a = packed[item];
    }
};

// broadcast a float32 field set (3, N) to (B, 3, N) float64
struct BroadcastF32toF64K {
    const float* src;
    double* dst;
    int64_t n; // 3*N
    FR3D_HD void operator()(int64_t i) const { dst[i] = (double)src[i % n]; }
};

// planar (B,3,N) float64 -> interleaved (B,N,3) float32/float64 (min_level == 0 output path)
template <class DstT>
struct InterleaveFlowK {
    const double* src;
    DstT* dst;
    int64_t N;
    FR3D_HD void operator()(int64_t item) const
    {
        const int q = (int)(item % 3);
        const int64_t r = item / 3;
        const int64_t o = r % N;
        const int64_t b = r / N;
        dst[item] = (DstT)src[(b * 3 + q) * N + o];
    }
};

// numpy.mean(w, axis=0) of float32 frames (compensate_recording_3D.py:388, 481-485): sequential
// float32 accumulation over the frame axis, then one float32 division by the count.
struct MeanFramesK {
    const float* src; // (T, n)
    float* dst;       // (n)
    int T;
    int64_t n;
    FR3D_HD void operator()(int64_t i) const
    {
        float acc = src[i];
        for (int t = 1; t < T; ++t)
            acc += src[(int64_t)t * n + i];
        dst[i] = acc / (float)T;
    }
};

// Per-frame flow statistics of BatchMotionCorrector (compensate_recording_3D.py:488-508): mean and max of the
// displacement magnitude, mean divergence (numpy.gradient: central differences, one-sided at the ends, unit
// spacing) and the mean of each component.  Tile kernel: block = (frame, chunk of voxels); partial sums are
// reduced in shared memory and added to the frame's accumulators (float64; max through its bit pattern, which
// orders like the value for non-negative numbers).  acc: (B, 6) = sum|w|, max|w| bits, sum div, sum u, sum v, sum w.
struct FlowStatsK {
    static constexpr int PHASES = 2;
    const float* flow; // (B, Z, Y, X, 3)
    double* acc;       // (B, 6), zero-initialised
    int Z, Y, X;
    int64_t chunk;     // voxels per block
    int chunks;        // blocks per frame
    FR3D_HD float comp(const float* f, int z, int y, int x, int q) const
    {
        return f[(((int64_t)z * Y + y) * X + x) * 3 + q];
    }
    FR3D_HD float grad(const float* f, int z, int y, int x, int q, int ax) const
    {
        const int n = ax == 0 ? Z : (ax == 1 ? Y : X);
        const int c = ax == 0 ? z : (ax == 1 ? y : x);
        if (n < 2)
            return 0.0f;
        const int lo = c > 0 ? c - 1 : c, hi = c < n - 1 ? c + 1 : c;
        const float a = comp(f, ax == 0 ? lo : z, ax == 1 ? lo : y, ax == 2 ? lo : x, q);
        const float b = comp(f, ax == 0 ? hi : z, ax == 1 ? hi : y, ax == 2 ? hi : x, q);
        return (b - a) / (float)(hi - lo);
    }
    FR3D_HD void phase(int ph, int64_t blk, int tid, int nthreads, double* sm) const
    {
        const int b = (int)(blk / chunks);
        const int64_t v0 = (blk % chunks) * chunk;
        const int64_t N = (int64_t)Z * Y * X;
        const int64_t v1 = v0 + chunk < N ? v0 + chunk : N;
        const float* f = flow + (int64_t)b * N * 3;
        double* part = sm + (size_t)tid * 6;
        if (ph == 0) {
            double s_mag = 0.0, s_div = 0.0, s_u = 0.0, s_v = 0.0, s_w = 0.0;
            float mx = 0.0f;
            for (int64_t v = v0 + tid; v < v1; v += nthreads) {
                const int x = (int)(v % X);
                const int y = (int)((v / X) % Y);
                const int z = (int)(v / ((int64_t)X * Y));
                const float u = f[v * 3], vv = f[v * 3 + 1], w = f[v * 3 + 2];
                const float mag = sqrtf(u * u + vv * vv + w * w);
                s_mag += (double)mag;
                mx = fmaxf(mx, mag);
                s_div += (double)(grad(f, z, y, x, 0, 2) + grad(f, z, y, x, 1, 1) + grad(f, z, y, x, 2, 0));
                s_u += (double)u;
                s_v += (double)vv;
                s_w += (double)w;
            }
            part[0] = s_mag;
            part[1] = (double)mx;
            part[2] = s_div;
            part[3] = s_u;
            part[4] = s_v;
            part[5] = s_w;
            return;
        }
        if (tid != 0)
            return;
        double t[6] = {0, 0, 0, 0, 0, 0};
        for (int k = 0; k < nthreads; ++k) {
            const double* p = sm + (size_t)k * 6;
            t[0] += p[0];
            t[1] = p[1] > t[1] ? p[1] : t[1];
            t[2] += p[2];
            t[3] += p[3];
            t[4] += p[4];
            t[5] += p[5];
        }
        double* a = acc + (int64_t)b * 6;
#ifdef __CUDA_ARCH__
        atomicAdd(a + 0, t[0]);
        atomicMax(reinterpret_cast<unsigned long long*>(a + 1), (unsigned long long)__double_as_longlong(t[1]));
        atomicAdd(a + 2, t[2]);
        atomicAdd(a + 3, t[3]);
        atomicAdd(a + 4, t[4]);
        atomicAdd(a + 5, t[5]);
#else
        a[0] += t[0];
        a[1] = t[1] > a[1] ? t[1] : a[1];
        a[2] += t[2];
        a[3] += t[3];
        a[4] += t[4];
        a[5] += t[5];
#endif
    }
};

// (B,6) accumulators -> (B,4): mean |w|, max |w|, mean divergence, |mean translation|
struct FlowStatsFinishK {
    const double* acc;
    double* out;
    double inv_n;
    FR3D_HD void operator()(int64_t b) const
    {
        const double* a = acc + b * 6;
        const double mu = a[3] * inv_n, mv = a[4] * inv_n, mw = a[5] * inv_n;
        out[b * 4 + 0] = a[0] * inv_n;
        out[b * 4 + 1] = a[1];
        out[b * 4 + 2] = a[2] * inv_n;
        out[b * 4 + 3] = sqrt(mu * mu + mv * mv + mw * mw);
    }
};

// numpy.mean(x, axis=0) of T float32 arrays held in a float64 array (compensate_recording_3D.py:425:
// sequential float64 accumulation over the frame axis, one division by the count)
struct MeanFramesF64K {
    const float* src; // (T, n)
    double* dst;      // (n)
    int T;
    int64_t n;
    FR3D_HD void operator()(int64_t i) const
    {
        double acc = (double)src[i];
        for (int t = 1; t < T; ++t)
            acc += (double)src[(int64_t)t * n + i];
        dst[i] = acc / (double)T;
    }
};

template <class T>
struct FillK {
    T* dst;
    int64_t inner, C; // dst[(i*C + c)] = v[c]
    double v[FR3D_MAX_CHANNELS];
    FR3D_HD void operator()(int64_t i) const { dst[i] = (T)v[i % C]; }
};

} // namespace fr3d
