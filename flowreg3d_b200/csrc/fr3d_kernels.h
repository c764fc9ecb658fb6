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
template <class SrcT, class DstT>
struct ResizePassK {
    const SrcT* src;
    DstT* dst;
    int64_t n[5], ss[5], ds[5];
    int r, P;
    const int32_t* idx;
    const float* wt;
    FR3D_HD void operator()(int64_t item) const
    {
        int64_t i[5];
        i[4] = item % n[4];
        item /= n[4];
        i[3] = item % n[3];
        item /= n[3];
        i[2] = item % n[2];
        item /= n[2];
        i[1] = item % n[1];
        i[0] = item / n[1];
        int64_t so = 0, dof = 0;
        for (int d = 0; d < 5; ++d) {
            dof += i[d] * ds[d];
            if (d != r)
                so += i[d] * ss[d];
        }
        const int32_t* ix = idx + i[r] * P;
        const float* w = wt + i[r] * P;
        const int64_t sr = ss[r];
        double acc = 0.0;
        for (int p = 0; p < P; ++p) {
            const float a = (float)src[so + (int64_t)ix[p] * sr];
            const float prod = a * w[p];
            acc += (double)prod;
        }
        dst[dof] = (DstT)(float)acc;
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

    FR3D_HD void operator()(int64_t item) const
    {
        const int x = (int)(item % X);
        int64_t q = item / X;
        const int y = (int)(q % Y);
        q /= Y;
        const int z = (int)(q % Z);
        const int b = (int)(q / Z);
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
    const int32_t* nbr;     // (6, npad): slots of x-, y-, z-, x+, y+, z+ (own slot if outside); [0] = -1 on pad slots
    const int32_t* perm;    // (npad): natural linear index (k*m + j)*n + i, or -1 on pad slots
    FR3D_HD int64_t nvox() const { return (int64_t)p * m * n; }
    FR3D_HD int32_t addr(int k, int j, int i) const { return rowbase[(k + j + i) * p + k] + j; }
};

// pad-slot defaults of the tables; item = slot
struct HPFillK {
    int32_t* nbr;
    int32_t* perm;
    int32_t npad;
    FR3D_HD void operator()(int64_t a) const
    {
        perm[a] = -1;
        nbr[a] = -1;
        for (int q = 1; q < 6; ++q)
            nbr[(int64_t)q * npad + a] = (int32_t)a;
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
        const int64_t np = g.npad;
        perm[a] = (int32_t)item;
        nbr[a] = i > 0 ? g.addr(k, j, i - 1) : a;
        nbr[np + a] = j > 0 ? g.addr(k, j - 1, i) : a;
        nbr[2 * np + a] = k > 0 ? g.addr(k - 1, j, i) : a;
        nbr[3 * np + a] = i < g.n - 1 ? g.addr(k, j, i + 1) : a;
        nbr[4 * np + a] = j < g.m - 1 ? g.addr(k, j + 1, i) : a;
        nbr[5 * np + a] = k < g.p - 1 ? g.addr(k + 1, j, i) : a;
    }
};

// Assemble the level system directly in solver storage: motion tensor J (B,C,10,npad), the constant
// part of the smoothness term L = ax*(u_ip+u_im-2u) + ay*(...) + az*(...) for u,v,w, with replicate
// boundary (the reference's edge-padded ring, core/optical_flow_3d.py:88-89,418-426), and the
// zero-initialised increments.  item = (b, slot).
template <class ST>
struct AssembleK {
    const float* f1;   // (C, N) planar natural, shared by all frames
    const float* f2;   // (B, C, N) planar natural (the warped moving image)
    const double* uvw; // (B, 3, N) natural
    double* J;         // (B, C, 10, npad) or null
    Vec4<ST>* L;       // (B, npad)
    Vec4<ST>* d;       // (B, npad)
    HPView hp;
    MTGeom g;
    int B, C;
    double ax, ay, az; // alpha/h^2
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t N = hp.nvox(), np = hp.npad;
        const int64_t a = item % np;
        const int b = (int)(item / np);
        Vec4<ST> zero;
        zero.x = zero.y = zero.z = zero.w = (ST)0;
        d[(int64_t)b * np + a] = zero;
        const int32_t nat = hp.perm[a];
        if (nat < 0) {
            L[(int64_t)b * np + a] = zero;
            return;
        }
        const int i = nat % g.n;
        const int j = (nat / g.n) % g.m;
        const int k = nat / (g.n * g.m);
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
        double Lq[3];
        for (int q = 0; q < 3; ++q) {
            const double* f = uvw + ((int64_t)b * 3 + q) * N;
            const double c0 = f[((int64_t)k * g.m + j) * g.n + i];
            const double lx = f[((int64_t)k * g.m + j) * g.n + ip] + f[((int64_t)k * g.m + j) * g.n + im_] - 2.0 * c0;
            const double ly = f[((int64_t)k * g.m + jp) * g.n + i] + f[((int64_t)k * g.m + jm) * g.n + i] - 2.0 * c0;
            const double lz = f[((int64_t)kp * g.m + j) * g.n + i] + f[((int64_t)km * g.m + j) * g.n + i] - 2.0 * c0;
            Lq[q] = ax * lx + ay * ly + az * lz;
        }
        Vec4<ST> Lv;
        Lv.x = (ST)Lq[0];
        Lv.y = (ST)Lq[1];
        Lv.z = (ST)Lq[2];
        Lv.w = (ST)0;
        L[(int64_t)b * np + a] = Lv;
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

struct Median5K {
    const double* src; // (nvol, p, m, n)
    double* dst;       // (nvol, p, m, n): dst = (add ? add : 0) + median
    const double* add; // optional (nvol, p, m, n)
    int p, m, n;
    FR3D_HD void operator()(int64_t item) const
    {
        const int64_t N = (int64_t)p * m * n;
        const int64_t vol = item / N;
        const int64_t o = item % N;
        const int i = (int)(o % n);
        const int j = (int)((o / n) % m);
        const int k = (int)(o / ((int64_t)n * m));
        const double* f = src + vol * N;
        int zi[5], yi[5], xi[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            zi[d] = mirror_idx(k + d - 2, p);
            yi[d] = mirror_idx(j + d - 2, m);
            xi[d] = mirror_idx(i + d - 2, n);
        }
        float a[64], rest[61];
#pragma unroll
        for (int t = 0; t < 125; ++t) {
            const float key = (float)f[((int64_t)zi[t / 25] * m + yi[(t / 5) % 5]) * n + xi[t % 5]];
            if (t < 64)
                a[t] = key;
            else
                rest[t - 64] = key;
        }
        const float med = ForgetStep<64>::run(a, rest);
        // recover the float64 value of rank 62 (0-based)
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
        double res = tmin;
        if (tmin != tmax) {
            // distinct float64 values share the median key (rare; mirrored duplicates are equal and
            // never get here): walk the tied values in increasing order up to rank (62 - less)
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
            res = cur;
        }
        dst[item] = add ? add[item] + res : res;
    }
};

// dst = a + b (element-wise, float64) -- levels too small for the median
struct AddK {
    const double* a;
    const double* b;
    double* dst;
    FR3D_HD void operator()(int64_t i) const { dst[i] = a[i] + b[i]; }
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

template <class T>
struct FillK {
    T* dst;
    int64_t inner, C; // dst[(i*C + c)] = v[c]
    double v[FR3D_MAX_CHANNELS];
    FR3D_HD void operator()(int64_t i) const { dst[i] = (T)v[i % C]; }
};

} // namespace fr3d
