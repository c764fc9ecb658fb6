// fr3d_common.h -- device abstraction shared by every kernel of libfr3d.
//
// Product build: nvcc, sm_100a.  Every kernel body is a functor whose operator()(int64_t item)
// is executed by one CUDA thread (fr3d_launch).
//
// FR3D_EMU build (tests/emu only, never shipped or loaded by flowreg3d_b200): the same functors
// are compiled by g++ and run as a serial loop, so the not-gpu tests can check the kernel LOGIC
// against the oracle in a container without a GPU.  It is a test harness, not a fallback: the
// Python package only ever loads the CUDA library and raises if it is missing.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <typeinfo>
#include <vector>

#include "../../include/fr3d.h"

#ifdef FR3D_EMU
#define FR3D_HD inline
#define FR3D_D inline
#define FR3D_LDCG(ptr) (*(ptr))
#define FR3D_STCG(ptr, v) (*(ptr) = (v))
typedef void* fr3d_stream_t;
#else
#include <cuda_runtime.h>
#define FR3D_HD __host__ __device__ __forceinline__
#define FR3D_D __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define FR3D_LDCG(ptr) __ldcg(ptr)
#define FR3D_STCG(ptr, v) __stcg(ptr, v)
#else
#define FR3D_LDCG(ptr) (*(ptr))
#define FR3D_STCG(ptr, v) (*(ptr) = (v))
#endif
typedef cudaStream_t fr3d_stream_t;
#endif

namespace fr3d {

struct Error {
    int code;
    std::string msg;
};

#define FR3D_THROW(code_, ...)                         \
    do {                                               \
        char _b[512];                                  \
        snprintf(_b, sizeof(_b), __VA_ARGS__);         \
        throw ::fr3d::Error{(code_), std::string(_b)}; \
    } while (0)

#define FR3D_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond))                            \
            FR3D_THROW(FR3D_ERR_ARG, __VA_ARGS__); \
    } while (0)

#ifndef FR3D_EMU
#define FR3D_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            FR3D_THROW(_e == cudaErrorMemoryAllocation ? FR3D_ERR_NOMEM : FR3D_ERR_CUDA,  \
                       "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                       __LINE__);                                                         \
    } while (0)
#endif

// ---- device memory ------------------------------------------------------------------------
struct Device {
    fr3d_stream_t stream = nullptr;
    int64_t launches = 0;
    int64_t bytes = 0;
    int sm_count = 1;
    int sor_ctas_per_sm = 0; // 0 = as many as fit
    int sor_kernel = 0;      // FR3D_OPT_SOR_KERNEL: 2 time-blocked tiles (fr3d_sor_tile.h), 1 staged wavefront (TMA bulk
                             // copies + mbarrier ring), 0 direct-load wavefront
    int sor_sched = -1;      // wavefront kernel item dealing (FR3D_OPT_SOR_SCHED): -1 = default for the state dtype, else
                             // percent by ticket | 128 (refresh items first)
    int sor_frames_per_item = 0; // frames one warp work item of the wavefront kernels covers, 0 = default 2 (FR3D_OPT_SOR_FRAMES_PER_ITEM)
    int warp_tile = 0;       // gather block shape tx | ty << 8 | tz << 16, 0 = 32 x 8 x 1 (FR3D_OPT_WARP_TILE)
    int sor_stages = 0;      // stages per warp of the staged kernel, 0 = default (FR3D_OPT_SOR_STAGES)
    int resize_x_rows = 1;   // FR3D_OPT_RESIZE_X_ROWS: X resampling pass with 4 rows per thread sharing the tap look-ups
    int spline_tma = 1;      // FR3D_OPT_SPLINE_TMA: bulk-copy (TMA) staging of the spline X pass
    int sor_tile_sweeps = 0, sor_tile_k = 0, sor_tile_j = 0, sor_tile_i = 0; // FR3D_OPT_SOR_TILE (0 = defaults)
    int sor_k0 = 0, sor_k1 = 0; // plane range of a z-slab solver launch (0, 0: all planes); set around the launch
    int cc_block_scans = 1;  // FR3D_OPT_CC_BLOCK_SCANS (block-cooperative scans: 30.1 -> 5.5 ms per 10 frames on a B200)
    int warp_factored = 1;   // FR3D_OPT_WARP_FACTORED (factored separable gather: 7.61 -> 6.18 ms per 16 frames on a B200)

    // Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline
    // figures).  Off by default; when on, every launch is bracketed by two event records.
    bool profiling = false;
#ifndef FR3D_EMU
    struct Span {
        const char* name;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get_event()
    {
        cudaEvent_t e;
        if (!pool.empty()) {
            e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEventCreate(&e);
        return e;
    }
#endif
#ifdef FR3D_EMU
    // kernel-logic emulator: no timing, but the profile report still lists WHICH functors ran (tests assert on it)
    std::map<std::string, int64_t> emu_counts;
    void emu_count(const char* name)
    {
        if (profiling)
            emu_counts[name] += 1;
    }
#endif
    void span_begin(const char* name)
    {
#ifndef FR3D_EMU
        if (!profiling)
            return;
        Span s{name, get_event(), get_event()};
        cudaEventRecord(s.a, stream);
        spans.push_back(s);
#else
        (void)name;
#endif
    }
    void span_end()
    {
#ifndef FR3D_EMU
        if (profiling)
            cudaEventRecord(spans.back().b, stream);
#endif
    }
    // name -> (launch count, total milliseconds); synchronises the stream; clears the spans
    std::map<std::string, std::pair<int64_t, double>> profile_collect()
    {
        std::map<std::string, std::pair<int64_t, double>> out;
#ifndef FR3D_EMU
        cudaStreamSynchronize(stream);
        for (Span& s : spans) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, s.a, s.b);
            auto& e = out[s.name];
            e.first += 1;
            e.second += ms;
            pool.push_back(s.a);
            pool.push_back(s.b);
        }
        spans.clear();
#else
        for (auto& kv : emu_counts)
            out[kv.first] = std::make_pair(kv.second, 0.0);
        emu_counts.clear();
#endif
        return out;
    }

    void* alloc(size_t n)
    {
        if (n == 0)
            n = 8;
        void* p = nullptr;
#ifdef FR3D_EMU
        p = malloc(n);
        if (!p)
            FR3D_THROW(FR3D_ERR_NOMEM, "host alloc of %zu bytes failed", n);
#else
        FR3D_CUDA(cudaMalloc(&p, n));
#endif
        bytes += (int64_t)n;
        return p;
    }
    void release(void* p, size_t n)
    {
        if (!p)
            return;
#ifdef FR3D_EMU
        free(p);
#else
        cudaFree(p);
#endif
        bytes -= (int64_t)(n ? n : 8);
    }
    void h2d(void* dst, const void* src, size_t n)
    {
        if (!n)
            return;
#ifdef FR3D_EMU
        memcpy(dst, src, n);
#else
        // tables and small parameter blocks only: pageable source, synchronous semantics wanted
        FR3D_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, stream));
        FR3D_CUDA(cudaStreamSynchronize(stream));
#endif
    }
    void d2d(void* dst, const void* src, size_t n)
    {
        if (!n)
            return;
#ifdef FR3D_EMU
        memmove(dst, src, n);
#else
        FR3D_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, stream));
#endif
    }
    void zero(void* dst, size_t n)
    {
        if (!n)
            return;
#ifdef FR3D_EMU
        memset(dst, 0, n);
#else
        FR3D_CUDA(cudaMemsetAsync(dst, 0, n, stream));
#endif
    }
    void sync()
    {
#ifndef FR3D_EMU
        FR3D_CUDA(cudaStreamSynchronize(stream));
#endif
    }
};

// Grow-only typed device buffer.
template <class T>
struct Buf {
    T* p = nullptr;
    size_t cap = 0; // elements
    Device* dev = nullptr;
    Buf() {}
    Buf(const Buf&) = delete;
    Buf& operator=(const Buf&) = delete;
    ~Buf() { reset(); }
    void reset()
    {
        if (p && dev)
            dev->release(p, cap * sizeof(T));
        p = nullptr;
        cap = 0;
    }
    T* ensure(Device& d, size_t n)
    {
        if (n > cap) {
            reset();
            dev = &d;
            p = (T*)d.alloc(n * sizeof(T));
            cap = n;
        }
        return p;
    }
    T* upload(Device& d, const T* host, size_t n)
    {
        ensure(d, n);
        d.h2d(p, host, n * sizeof(T));
        return p;
    }
};

// ---- launch -------------------------------------------------------------------------------
#ifndef FR3D_EMU
template <class K>
__global__ void __launch_bounds__(256) fr3d_kernel(const K k, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        k(i);
}
#endif

#ifndef FR3D_EMU
// Variant that asks ptxas for two resident CTAs per SM (<= 128 registers per thread).
template <class K>
__global__ void __launch_bounds__(256, 2) fr3d_kernel_occ2(const K k, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        k(i);
}
#endif

#ifndef FR3D_EMU
// Variant with a chosen number of resident CTAs per SM (register cap 65536 / (256 MB) per thread).
template <class K, int MB>
__global__ void __launch_bounds__(256, MB) fr3d_kernel_occ(const K k, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        k(i);
}
#endif

template <int MB, class K>
void launch_occ(Device& dev, const K& k, int64_t n)
{
    if (n <= 0)
        return;
#ifdef FR3D_EMU
    dev.emu_count(typeid(K).name());
    for (int64_t i = 0; i < n; ++i)
        k(i);
#else
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    FR3D_REQUIRE(blocks < (int64_t)2147483647, "launch too large: %lld items", (long long)n);
    dev.span_begin(typeid(K).name());
    fr3d_kernel_occ<K, MB><<<(unsigned)blocks, threads, 0, dev.stream>>>(k, n);
    dev.span_end();
    FR3D_CUDA(cudaGetLastError());
#endif
    dev.launches++;
}

template <class K>
void launch_occ2(Device& dev, const K& k, int64_t n)
{
    if (n <= 0)
        return;
#ifdef FR3D_EMU
    dev.emu_count(typeid(K).name());
    for (int64_t i = 0; i < n; ++i)
        k(i);
#else
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    FR3D_REQUIRE(blocks < (int64_t)2147483647, "launch too large: %lld items", (long long)n);
    dev.span_begin(typeid(K).name());
    fr3d_kernel_occ2<K><<<(unsigned)blocks, threads, 0, dev.stream>>>(k, n);
    dev.span_end();
    FR3D_CUDA(cudaGetLastError());
#endif
    dev.launches++;
}

// Division of a 32-bit unsigned value by a runtime constant through multiply-high + shift
// (Granlund-Montgomery); exact for every 32-bit n.  Built on the host.
struct FastDiv {
    uint32_t d, mul, sh;
    FastDiv() : d(1), mul(0), sh(0) {}
    explicit FastDiv(uint32_t d_) : d(d_), mul(0), sh(0)
    {
        if (d <= 1)
            return;
        uint32_t l = 0;
        while (l < 32 && (1ull << l) < d)
            ++l;
        sh = l - 1;
        mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    }
    FR3D_HD uint32_t div(uint32_t n) const
    {
        if (d <= 1)
            return n;
#ifdef __CUDA_ARCH__
        const uint32_t t = __umulhi(mul, n);
#else
        const uint32_t t = (uint32_t)(((uint64_t)mul * n) >> 32);
#endif
        return (t + ((n - t) >> 1)) >> sh;
    }
    FR3D_HD void divmod(uint32_t n, uint32_t& q, uint32_t& r) const
    {
        q = div(n);
        r = n - q * d;
    }
};

// One logical thread per item.
template <class K>
void launch(Device& dev, const K& k, int64_t n)
{
    if (n <= 0)
        return;
#ifdef FR3D_EMU
    dev.emu_count(typeid(K).name());
    for (int64_t i = 0; i < n; ++i)
        k(i);
#else
    const int threads = 256;
    const int64_t blocks = (n + threads - 1) / threads;
    FR3D_REQUIRE(blocks < (int64_t)2147483647, "launch too large: %lld items", (long long)n);
    dev.span_begin(typeid(K).name());
    fr3d_kernel<K><<<(unsigned)blocks, threads, 0, dev.stream>>>(k, n);
    dev.span_end();
    FR3D_CUDA(cudaGetLastError());
#endif
    dev.launches++;
}

// Solver state vector {du, dv, dw}.  float: padded to four components (one 16-byte access per voxel).  double: THREE
// components, 24 bytes per voxel (round 2: the fourth word was 8 of the 32 + 32 bytes the float64 state moves per
// voxel and sweep for its own value, and of every 32-byte neighbour gather; the accesses become three 8-byte ones,
// still contiguous across a warp).  The name is kept from the padded layout; `set_pad` clears the fourth component
// where there is one.
template <class T>
struct alignas(16) Vec4 {
    T x, y, z, w;
};
template <>
struct alignas(8) Vec4<double> {
    double x, y, z;
};
FR3D_HD void set_pad(Vec4<float>& v) { v.w = 0.0f; }
FR3D_HD void set_pad(Vec4<double>&) {}

// L2-only (cache-global) vector load / store: data written by other SMs in the previous wave
FR3D_HD Vec4<float> ld4_cg(const Vec4<float>* p)
{
#ifdef __CUDA_ARCH__
    const float4 v = __ldcg(reinterpret_cast<const float4*>(p));
    Vec4<float> r;
    r.x = v.x;
    r.y = v.y;
    r.z = v.z;
    r.w = v.w;
    return r;
#else
    return *p;
#endif
}
FR3D_HD Vec4<double> ld4_cg(const Vec4<double>* p)
{
#ifdef __CUDA_ARCH__
    const double* q = reinterpret_cast<const double*>(p);
    Vec4<double> r;
    r.x = __ldcg(q);
    r.y = __ldcg(q + 1);
    r.z = __ldcg(q + 2);
    return r;
#else
    return *p;
#endif
}
// L1-cached vector loads (neighbour increments: hyperplanes that are not written during the current wave;
// every CTA invalidates its SM's L1 at the wave barrier -- see fr3d_sor.h)
FR3D_HD Vec4<float> ld4_ca(const Vec4<float>* p)
{
#ifdef __CUDA_ARCH__
    const float4 v = *reinterpret_cast<const float4*>(p);
    Vec4<float> r;
    r.x = v.x;
    r.y = v.y;
    r.z = v.z;
    r.w = v.w;
    return r;
#else
    return *p;
#endif
}
FR3D_HD Vec4<double> ld4_ca(const Vec4<double>* p) { return *p; }
FR3D_HD void st4_cg(Vec4<float>* p, const Vec4<float>& v)
{
#ifdef __CUDA_ARCH__
    __stcg(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, v.w));
#else
    *p = v;
#endif
}
FR3D_HD void st4_cg(Vec4<double>* p, const Vec4<double>& v)
{
#ifdef __CUDA_ARCH__
    double* q = reinterpret_cast<double*>(p);
    __stcg(q, v.x);
    __stcg(q + 1, v.y);
    __stcg(q + 2, v.z);
#else
    *p = v;
#endif
}

// ---- block-cooperative launch ---------------------------------------------------------------
// A "tile kernel" is a functor with  phase(ph, block, tid, nthreads, smem)  executed for
// ph = 0 .. K::PHASES-1 with a block-wide barrier after every phase; all cross-thread communication
// goes through `smem` (registers do not survive a phase).  CUDA: one CTA per block, dynamic shared
// memory.  FR3D_EMU: the phases run as serial loops over the threads of a block.
#ifndef FR3D_EMU
template <class K>
__global__ void __launch_bounds__(256) fr3d_tile_kernel(const K k)
{
    extern __shared__ double fr3d_smem[];
    for (int ph = 0; ph < K::PHASES; ++ph) {
        k.phase(ph, (int64_t)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, fr3d_smem);
        __syncthreads();
    }
}
#endif

template <class K>
void launch_tiles(Device& dev, const K& k, int64_t nblocks, int threads, size_t smem_bytes)
{
    if (nblocks <= 0)
        return;
#ifdef FR3D_EMU
    dev.emu_count(typeid(K).name());
    std::vector<double> smem(smem_bytes / sizeof(double) + 1);
    for (int64_t b = 0; b < nblocks; ++b)
        for (int ph = 0; ph < K::PHASES; ++ph)
            for (int t = 0; t < threads; ++t)
                k.phase(ph, b, t, threads, smem.data());
#else
    FR3D_REQUIRE(threads >= 1 && threads <= 256, "tile kernel: %d threads", threads);
    FR3D_REQUIRE(nblocks < (int64_t)2147483647, "launch too large: %lld blocks", (long long)nblocks);
    static bool configured = false; // per kernel type
    if (!configured) {
        FR3D_CUDA(cudaFuncSetAttribute(fr3d_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    dev.span_begin(typeid(K).name());
    fr3d_tile_kernel<K><<<(unsigned)nblocks, threads, smem_bytes, dev.stream>>>(k);
    dev.span_end();
    FR3D_CUDA(cudaGetLastError());
#endif
    dev.launches++;
}

FR3D_HD int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

FR3D_HD size_t dtype_size(int dt)
{
    switch (dt) {
    case FR3D_F32: return 4;
    case FR3D_F64: return 8;
    case FR3D_U8: return 1;
    case FR3D_U16: return 2;
    case FR3D_I16: return 2;
    case FR3D_I32: return 4;
    }
    return 0;
}

// The same read split in two: the raw bits now (a load whose result nobody waits for yet), the conversion later.
FR3D_HD uint64_t load_raw_bits(const void* base, int dt, int64_t i)
{
    switch (dt) {
    case FR3D_F32: return (uint64_t)((const uint32_t*)base)[i];
    case FR3D_F64: return ((const uint64_t*)base)[i];
    case FR3D_U8: return (uint64_t)((const uint8_t*)base)[i];
    case FR3D_U16: return (uint64_t)((const uint16_t*)base)[i];
    case FR3D_I16: return (uint64_t)(uint16_t)((const int16_t*)base)[i];
    case FR3D_I32: return (uint64_t)(uint32_t)((const int32_t*)base)[i];
    }
    return 0;
}
FR3D_HD double raw_bits_as_double(uint64_t bits, int dt)
{
    switch (dt) {
    case FR3D_F32: {
        const uint32_t b32 = (uint32_t)bits;
        float f;
        memcpy(&f, &b32, 4);
        return (double)f;
    }
    case FR3D_F64: {
        double d;
        memcpy(&d, &bits, 8);
        return d;
    }
    case FR3D_U8: return (double)(uint8_t)bits;
    case FR3D_U16: return (double)(uint16_t)bits;
    case FR3D_I16: return (double)(int16_t)(uint16_t)bits;
    case FR3D_I32: return (double)(int32_t)(uint32_t)bits;
    }
    return 0.0;
}

// Read one element of a dtype-tagged array as double (exact for every supported dtype).
FR3D_HD double load_as_double(const void* base, int dt, int64_t i)
{
    switch (dt) {
    case FR3D_F32: return (double)((const float*)base)[i];
    case FR3D_F64: return ((const double*)base)[i];
    case FR3D_U8: return (double)((const uint8_t*)base)[i];
    case FR3D_U16: return (double)((const uint16_t*)base)[i];
    case FR3D_I16: return (double)((const int16_t*)base)[i];
    case FR3D_I32: return (double)((const int32_t*)base)[i];
    }
    return 0.0;
}

} // namespace fr3d
