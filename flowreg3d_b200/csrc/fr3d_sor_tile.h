// fr3d_sor_tile.h -- TIME-BLOCKED level solver: skewed space-time tiles, increments resident in shared memory.
//
// Same arithmetic and the same update order as the wavefront kernel (fr3d_sor.h): voxel (k,j,i) of sweep t still
// sees (k-1,j,i), (k,j-1,i), (k,j,i-1) of sweep t and (k+1,..), (..,j+1,..), (..,i+1) of sweep t-1 -- any schedule
// that respects those dependencies produces bit-identical increments.  What changes is WHO does the work WHEN:
//
//   * The level is cut into tiles of K x J x I voxels and the T sweeps into time blocks of Tb sweeps.  One
//     thread block executes one TASK = (tile (a,b,c), time block tau, frame): all Tb sweeps of that tile, with the
//     tile's increments held in shared memory for the whole task.
//   * Gauss-Seidel dependencies forbid redundant halo work, so the tile is SKEWED in time: at local sweep r the
//     task covers the region [aK - r, aK + K - r) x [bJ - r, ..) x [cI - r, ..) (axes that are not tiled do not
//     shift).  Then every "+1 neighbour of the previous sweep" lies inside the task's own previous region, every
//     "-1 neighbour of the same sweep" lies in the region or in a tile with a smaller index, and a voxel that the
//     shift moves out of the region is picked up, with its latest value, by the next tile.
//   * Task (a,b,c,tau) depends only on tasks with a smaller TILE WAVE  w = a + b + c + D*tau,  D = 1 + (number of
//     tiled axes): within a time block on (a-1,b,c), (a,b-1,c), (a,b,c-1); across time blocks on at most
//     (a+1,b+1,c+1,tau-1), because Tb <= K, J, I.  A level solve is therefore ~ na+nb+nc + D*T/Tb grid-wide steps
//     (about 100 at config 2) instead of S + 2T (about 540), each with thousands of independent tasks.
//   * Inside a task the (voxel, sweep) pairs are ordered by the internal wave  omega = (k'+j'+i') + (2 - nsk)*r
//     (k',j',i' local to the shifted region, nsk = number of tiled axes): all inputs of a pair carry omega-1 or
//     omega-2, so one __syncthreads per internal wave suffices; with two tiled axes all Tb sweeps advance in
//     lockstep over the K+J+I-2 local hyperplanes.
//   * Global memory keeps ONE copy of the increments, as in the sequential algorithm.  A task loads the box
//     "union of its regions + one halo cell per side" at its start (every value it will ever need from outside is
//     final by then: it was produced by tasks of smaller tile waves), updates it in place in shared memory and
//     writes back exactly the cells it updated.  The pre-combined system AB stays in global memory (written at the
//     psi refresh, re-read on the following sweeps from L2): a voxel that migrates to the next tile finds it there.
//
// Traffic: the increments cost one box load + one write-back per Tb sweeps instead of 8 vector accesses per sweep,
// the neighbour table is not read at all, and AB is re-read within ~10 us by the same SM, i.e. from L2.  DRAM sees
// J, L once per refresh, AB once or twice per refresh and the increments about twice per time block.
#pragma once

namespace fr3d {

struct SorTileGeom {
    int K, J, I;       // tile extents (an axis that is not tiled has extent = level extent)
    int sk, sj, si;    // 1: axis tiled, regions shift by -1 per local sweep along it
    int nsk, D;        // number of tiled axes; tile-wave stride of a time block = 1 + nsk
    int na, nb, nc;    // tiles per axis (the shift needs ceil((extent + Tb - 1) / tile))
    int Tb, ntau;      // sweeps per time block, time blocks
    int nwaves;        // tile waves
    int FK, FJ, FI;    // shared-memory box (FI padded to an even length, see sor_tile_geom)
    int nbox;          // FK * FJ * FI
    int nhp;           // local hyperplanes K + J + I - 2
    int off;           // omega = (k'+j'+i') + (2 - nsk) * r + off  >= 0
    int nomega;        // internal waves per task
    int ncell;         // K * J * I
    int ne;            // values of a + b + c: na + nb + nc - 2
    int RS;            // hyperplanes k+j+i crossed by the box: FK + FJ + FI - 2 (rows of the rowbase slice)
    int prepass;       // 1: psi refreshes fall on local sweep 0 only (lag % Tb == 0) and are done for the whole
                       //    region in one parallel pass before the internal waves
};

inline int sor_tile_axis(int extent, int want, int whole_upto, int Tb, int& tiled, int& ntiles)
{
    if (extent <= whole_upto || extent <= want) {
        tiled = 0;
        ntiles = 1;
        return extent;
    }
    tiled = 1;
    ntiles = (extent + Tb - 1 + want - 1) / want;
    return want;
}

// Tile geometry of a p x m x n level solved for T sweeps.  tile = requested extent of a tiled axis, Tb = requested
// sweeps per time block (clamped to the tile extent: the cross-block dependency reaches one tile up at most).
inline SorTileGeom sor_tile_geom(int p, int m, int n, int T, int lag, int Tb, int tk, int tj, int ti)
{
    SorTileGeom G;
    if (Tb > T)
        Tb = T;
    int lim = tk < tj ? tk : tj;
    lim = ti < lim ? ti : lim;
    if (Tb > lim)
        Tb = lim;
    if (Tb < 1)
        Tb = 1;
    G.Tb = Tb;
    G.ntau = (T + Tb - 1) / Tb;
    G.K = sor_tile_axis(p, tk, 12, Tb, G.sk, G.na);
    G.J = sor_tile_axis(m, tj, 0, Tb, G.sj, G.nb);
    G.I = sor_tile_axis(n, ti, 0, Tb, G.si, G.nc);
    G.nsk = G.sk + G.sj + G.si;
    G.D = 1 + G.nsk;
    G.nwaves = (G.na - 1) + (G.nb - 1) + (G.nc - 1) + G.D * (G.ntau - 1) + 1;
    G.FK = G.K + G.sk * (Tb - 1) + 2;
    G.FJ = G.J + G.sj * (Tb - 1) + 2;
    G.FI = G.I + G.si * (Tb - 1) + 2;
    // consecutive lanes walk a local hyperplane along (j+1, i-1): their box cells are FI - 1 apart, which must be
    // odd for conflict-free shared-memory access (4- and 8-byte elements alike)
    if ((G.FI & 1) == 1)
        G.FI += 1;
    G.nbox = G.FK * G.FJ * G.FI;
    G.RS = G.FK + G.FJ + G.FI - 2;
    G.prepass = (lag % Tb) == 0 ? 1 : 0;
    G.nhp = G.K + G.J + G.I - 2;
    G.off = G.nsk > 2 ? (G.nsk - 2) * (Tb - 1) : 0;
    const int drift = G.nsk >= 2 ? G.nsk - 2 : 2 - G.nsk;
    G.nomega = G.nhp + drift * (Tb - 1);
    G.ncell = G.K * G.J * G.I;
    G.ne = G.na + G.nb + G.nc - 2;
    return G;
}

// pairs (b, c) in [0,nb) x [0,nc) with b + c = f
FR3D_HD int sor_tile_count2(const SorTileGeom& G, int f)
{
    const int lo = f - (G.nc - 1) > 0 ? f - (G.nc - 1) : 0;
    const int hi = f < G.nb - 1 ? f : G.nb - 1;
    return hi >= lo ? hi - lo + 1 : 0;
}
// triples (a, b, c) with a + b + c = e
FR3D_HD int sor_tile_count3(const SorTileGeom& G, int e)
{
    int n = 0;
    for (int a = 0; a < G.na && a <= e; ++a)
        n += sor_tile_count2(G, e - a);
    return n;
}

// dynamic shared memory of the tile kernel, in bytes
template <class ST>
inline size_t sor_tile_smem(const SorTileGeom& G)
{
    size_t b = (size_t)3 * G.nbox * sizeof(ST);
    b = (b + 15) & ~(size_t)15;
    b += (size_t)G.ncell * 4;           // cells
    b += (size_t)(G.nhp + 2) * 4;       // cell_start (+1)
    b += (size_t)(G.nhp + 1) * 4;       // per-hyperplane counts (table build)
    b += (size_t)G.RS * G.FK * 4;       // rowbase slice of the current task
    b += (size_t)G.ne * 4;              // cnt3
    b += (size_t)(G.ntau + 1) * 4;      // tau prefix of the current wave
    b += 8 * 4;                         // decoded task
    return b;
}

// Shared-memory tables of a thread block.
struct SorTileTabs {
    uint32_t* cells;   // (ncell): local cells k' | j' << 8 | i' << 16, grouped by local hyperplane
    int* cell_start;   // (nhp + 1)
    int* cursor;       // (nhp + 1) scratch of the table build
    int* rb;           // (RS, FK): rowbase[(s) * p + k] for the hyperplanes / planes of the current task's box
    int* cnt3;         // (ne)
    int* tau_pref;     // (ntau + 1)
    int* task;         // (8): frame, a, b, c, tau
};

template <class ST>
FR3D_HD SorTileTabs sor_tile_tabs(const SorTileGeom& G, unsigned char* smem)
{
    size_t b = (size_t)3 * G.nbox * sizeof(ST);
    b = (b + 15) & ~(size_t)15;
    SorTileTabs t;
    t.cells = reinterpret_cast<uint32_t*>(smem + b);
    t.cell_start = reinterpret_cast<int*>(t.cells + G.ncell);
    t.cursor = t.cell_start + (G.nhp + 2);
    t.rb = t.cursor + (G.nhp + 1);
    t.cnt3 = t.rb + G.RS * G.FK;
    t.tau_pref = t.cnt3 + G.ne;
    t.task = t.tau_pref + (G.ntau + 1);
    return t;
}

// L2 prefetch of one 32-byte sector (no register, no scoreboard): the psi-refresh inputs of a task are requested while
// its box is being loaded
FR3D_HD void sor_tile_prefetch(const void* ptr)
{
#ifdef __CUDA_ARCH__
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

// Execution policies: the same task code runs as a CUDA thread block or as one serial "thread" in the emulator.
struct SorSerialPar {
    FR3D_HD int tid() const { return 0; }
    FR3D_HD int nth() const { return 1; }
    FR3D_HD void sync() const {}
    FR3D_HD int add(int* p, int v) const
    {
        const int o = *p;
        *p = o + v;
        return o;
    }
    FR3D_HD void mark(int) const {}
};
#ifdef __CUDACC__
#ifdef FR3D_TILE_TIMING
// tuning builds only: cycles per task phase, summed over all blocks (thread 0 of each block)
__device__ unsigned long long fr3d_tile_clk[8];
#endif
struct SorBlockPar {
#ifdef FR3D_TILE_TIMING
    mutable long long last = 0;
    __device__ __forceinline__ void mark(int phase) const
    {
        if (threadIdx.x == 0) {
            const long long t = clock64();
            if (phase >= 0)
                atomicAdd(&fr3d_tile_clk[phase], (unsigned long long)(t - last));
            last = t;
        }
    }
#else
    __device__ __forceinline__ void mark(int) const {}
#endif
    __device__ __forceinline__ int tid() const { return threadIdx.x; }
    __device__ __forceinline__ int nth() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }
    __device__ __forceinline__ int add(int* p, int v) const { return atomicAdd(p, v); }
};
#endif

// cells (j', i') of row k' of local hyperplane h: j' in [lo, hi]
FR3D_HD void sor_tile_row(const SorTileGeom& G, int h, int kk, int& lo, int& hi)
{
    const int f = h - kk;                       // j' + i'
    lo = f - (G.I - 1) > 0 ? f - (G.I - 1) : 0;
    hi = f < G.J - 1 ? f : G.J - 1;
}

// Build the per-block tables (once per launch).  Cells of a local hyperplane are ordered by (k', j'): consecutive
// entries of a row are consecutive slots of the hyperplane-major solver storage (coalesced AB / J accesses).
template <class Par>
FR3D_HD void sor_tile_build_tabs(const SorTileGeom& G, const SorTileTabs& tb, const Par& par)
{
    for (int e = par.tid(); e < G.ne; e += par.nth())
        tb.cnt3[e] = sor_tile_count3(G, e);
    for (int h = par.tid(); h < G.nhp; h += par.nth()) {
        int cnt = 0;
        for (int kk = 0; kk < G.K && kk <= h; ++kk) {
            int lo, hi;
            sor_tile_row(G, h, kk, lo, hi);
            cnt += hi >= lo ? hi - lo + 1 : 0;
        }
        tb.cursor[h] = cnt;
    }
    par.sync();
    if (par.tid() == 0) {
        int sum = 0;
        for (int h = 0; h < G.nhp; ++h) {
            tb.cell_start[h] = sum;
            sum += tb.cursor[h];
        }
        tb.cell_start[G.nhp] = sum;
    }
    par.sync();
    for (int hk = par.tid(); hk < G.nhp * G.K; hk += par.nth()) {
        const int h = hk / G.K, kk = hk - h * G.K;
        if (kk > h)
            continue;
        int pos = tb.cell_start[h];
        for (int k2 = 0; k2 < kk; ++k2) {
            int lo, hi;
            sor_tile_row(G, h, k2, lo, hi);
            pos += hi >= lo ? hi - lo + 1 : 0;
        }
        int lo, hi;
        sor_tile_row(G, h, kk, lo, hi);
        for (int jj = lo; jj <= hi; ++jj)
            tb.cells[pos++] = (uint32_t)kk | ((uint32_t)jj << 8) | ((uint32_t)(h - kk - jj) << 16);
    }
    par.sync();
}

// Tile-wave bookkeeping: tau_pref[tau] = tasks (per frame) of wave w in time blocks < tau.  Returns via tau_pref[ntau]
// the number of tasks per frame.
template <class Par>
FR3D_HD void sor_tile_wave_prefix(const SorTileGeom& G, const SorTileTabs& tb, int w, const Par& par)
{
    if (par.tid() == 0) {
        int s = 0;
        for (int tau = 0; tau < G.ntau; ++tau) {
            tb.tau_pref[tau] = s;
            const int e = w - G.D * tau;
            if (e >= 0 && e < G.ne)
                s += tb.cnt3[e];
        }
        tb.tau_pref[G.ntau] = s;
    }
    par.sync();
}

// Decode task `idx` of wave w (0 <= idx < B * tau_pref[ntau]) into tb.task = {frame, a, b, c, tau}.
FR3D_HD void sor_tile_decode(const SorTileGeom& G, const SorTileTabs& tb, int w, int idx)
{
    const int per_frame = tb.tau_pref[G.ntau];
    const int frame = idx / per_frame;
    int rem = idx - frame * per_frame;
    int tau = 0;
    while (tau + 1 < G.ntau && tb.tau_pref[tau + 1] <= rem)
        ++tau;
    rem -= tb.tau_pref[tau];
    const int e = w - G.D * tau;
    int a = e - (G.nb - 1) - (G.nc - 1);
    a = a > 0 ? a : 0;
    for (;; ++a) {
        const int c2 = sor_tile_count2(G, e - a);
        if (rem < c2)
            break;
        rem -= c2;
    }
    const int f = e - a;
    const int b = (f - (G.nc - 1) > 0 ? f - (G.nc - 1) : 0) + rem;
    tb.task[0] = frame;
    tb.task[1] = a;
    tb.task[2] = b;
    tb.task[3] = f - b;
    tb.task[4] = tau;
}

// One task: Tb sweeps of tile (ta, tb_, tc) in time block tau of frame `frame`.  sm: 3 * nbox values (du | dv | dw).
template <class ST, int C, class Par>
FR3D_HD void sor_tile_task(const SorParams<ST>& P, const SorTileGeom& G, const SorTileTabs& tabs, ST* sm, int frame, int ta,
                           int tb_, int tc, int tau, const Par& par)
{
    const HPView& g = P.g;
    const int p = g.p, m = g.m, n = g.n;
    const int64_t np = g.npad;
    const int Tb = G.Tb;
    const int nr = P.T - tau * Tb < Tb ? P.T - tau * Tb : Tb;
    const int k0 = ta * G.K, j0 = tb_ * G.J, i0 = tc * G.I;             // region origin at local sweep 0
    const int ok = k0 - G.sk * (Tb - 1) - 1, oj = j0 - G.sj * (Tb - 1) - 1, oi = i0 - G.si * (Tb - 1) - 1;
    const int os = ok + oj + oi;                                        // first hyperplane of the rowbase slice
    const int FJI = G.FJ * G.FI;
    const int nbox = G.nbox;
    ST* const su = sm;
    ST* const sv = sm + nbox;
    ST* const sw = sm + 2 * nbox;
    Vec4<ST>* const dg = P.d + (int64_t)frame * np;
    const int* const rb = tabs.rb;
    // slot of level voxel (k, j, i) of the box: rowbase[(k+j+i) * p + k] + j, the row bases from shared memory
#define FR3D_TILE_SLOT(k_, j_, i_) ((int64_t)rb[((k_) + (j_) + (i_) - os) * G.FK + ((k_) - ok)] + (j_))

    par.mark(-1);
    // ---- 0. row bases of the hyperplanes / planes the box crosses
    for (int c = par.tid(); c < G.RS * G.FK; c += par.nth()) {
        const int rs = c / G.FK, bk = c - rs * G.FK;
        const int sg = os + rs, k = ok + bk;
        tabs.rb[c] = (sg >= 0 && sg < g.S && k >= 0 && k < p) ? g.rowbase[sg * p + k] : 0;
    }
    par.sync();
    par.mark(0);

    // ---- 1. box load: union of the regions + one halo cell per side (cells outside the level stay unset: an
    //         out-of-level neighbour is replaced by the voxel's own value below)
    for (int c = par.tid(); c < nbox; c += par.nth()) {
        const int bk = c / FJI, rem = c - bk * FJI, bj = rem / G.FI, bi = rem - bj * G.FI;
        const int k = ok + bk, j = oj + bj, i = oi + bi;
        if (k >= 0 && k < p && j >= 0 && j < m && i >= 0 && i < n) {
            const Vec4<ST> v = ld4_cg(dg + FR3D_TILE_SLOT(k, j, i));
            su[c] = v.x;
            sv[c] = v.y;
            sw[c] = v.z;
        }
    }
    const bool refresh_block = ((tau * Tb) % P.lag) == 0;
    if (G.prepass && refresh_block) {
        // request the motion tensor and the Laplacian term of the region while the box loads are in flight
        for (int it = par.tid(); it < G.ncell; it += par.nth()) {
            const uint32_t cell = tabs.cells[it];
            const int k = k0 + (int)(cell & 0xff), j = j0 + (int)((cell >> 8) & 0xff), i = i0 + (int)(cell >> 16);
            if (k >= p || j >= m || i >= n)
                continue;
            const int64_t a = FR3D_TILE_SLOT(k, j, i);
            const double* Jb = P.J + (int64_t)frame * C * 10 * np + a;
#pragma unroll
            for (int e = 0; e < 10 * C; ++e)
                sor_tile_prefetch(Jb + (int64_t)e * np);
            sor_tile_prefetch(P.L + (int64_t)frame * np + a);
        }
    }
    par.sync();
    par.mark(1);

    const double den0 = 2.0 * P.ax + 2.0 * P.ay + 2.0 * P.az;
    // ---- 1b. psi refresh of the whole region in one parallel pass (prepass geometry: refresh sweeps are local sweep
    //          0 of a time block; a voxel's psi depends on its own increments before that sweep = the box values)
    if (G.prepass && refresh_block) {
        for (int it = par.tid(); it < G.ncell; it += par.nth()) {
            const uint32_t cell = tabs.cells[it];
            const int k = k0 + (int)(cell & 0xff), j = j0 + (int)((cell >> 8) & 0xff), i = i0 + (int)(cell >> 16);
            if (k >= p || j >= m || i >= n)
                continue;
            const int c0 = (k - ok) * FJI + (j - oj) * G.FI + (i - oi);
            const int64_t a = FR3D_TILE_SLOT(k, j, i);
            double A[9];
            const Vec4<ST> L = ld4_cg(P.L + (int64_t)frame * np + a);
            sor_refresh<C>(P.a_data, P.J + (int64_t)frame * C * 10 * np, P.wgt, np, a, (double)su[c0], (double)sv[c0],
                           (double)sw[c0], den0, A);
            A[6] -= (double)L.x;
            A[7] -= (double)L.y;
            A[8] -= (double)L.z;
            double* AB = P.AB + sor_ab_at(P, frame, a, 0);
#pragma unroll
            for (int e = 0; e < 9; ++e)
                FR3D_STCG(AB + e * 32, A[e]);
        }
        par.sync();     // the waves re-read AB through L2
    }
    par.mark(2);

    // ---- 2. internal waves
    const int drift = 2 - G.nsk;
    for (int om = 0; om < G.nomega; ++om) {
        // items of this wave: for every local sweep r the cells of local hyperplane h_r = om - drift*r - off
        int total = 0;
        for (int r = 0; r < nr; ++r) {
            const int h = om - drift * r - G.off;
            if (h >= 0 && h < G.nhp)
                total += tabs.cell_start[h + 1] - tabs.cell_start[h];
        }
        for (int it = par.tid(); it < total; it += par.nth()) {
            int r = 0, rem = it, h = 0;
            for (;; ++r) {
                h = om - drift * r - G.off;
                const int cnt = (h >= 0 && h < G.nhp) ? tabs.cell_start[h + 1] - tabs.cell_start[h] : 0;
                if (rem < cnt)
                    break;
                rem -= cnt;
            }
            const uint32_t cell = tabs.cells[tabs.cell_start[h] + rem];
            const int k = k0 - G.sk * r + (int)(cell & 0xff);
            const int j = j0 - G.sj * r + (int)((cell >> 8) & 0xff);
            const int i = i0 - G.si * r + (int)(cell >> 16);
            if (k < 0 || k >= p || j < 0 || j >= m || i < 0 || i >= n)
                continue;
            const int c0 = (k - ok) * FJI + (j - oj) * G.FI + (i - oi);
            const int64_t a = FR3D_TILE_SLOT(k, j, i);
            double* AB = P.AB + sor_ab_at(P, frame, a, 0);
            SorIn<ST> in;
            const int t = tau * Tb + r;
            const bool refresh = !G.prepass && (t % P.lag) == 0;
            if (!refresh) {
#pragma unroll
                for (int e = 0; e < 9; ++e)
                    in.A[e] = FR3D_LDCG(AB + e * 32);
            }
            const int cxm = i > 0 ? c0 - 1 : c0, cxp = i < n - 1 ? c0 + 1 : c0;
            const int cym = j > 0 ? c0 - G.FI : c0, cyp = j < m - 1 ? c0 + G.FI : c0;
            const int czm = k > 0 ? c0 - FJI : c0, czp = k < p - 1 ? c0 + FJI : c0;
            in.own.x = su[c0], in.own.y = sv[c0], in.own.z = sw[c0];
            in.xm.x = su[cxm], in.xm.y = sv[cxm], in.xm.z = sw[cxm];
            in.xp.x = su[cxp], in.xp.y = sv[cxp], in.xp.z = sw[cxp];
            in.ym.x = su[cym], in.ym.y = sv[cym], in.ym.z = sw[cym];
            in.yp.x = su[cyp], in.yp.y = sv[cyp], in.yp.z = sw[cyp];
            in.zm.x = su[czm], in.zm.y = sv[czm], in.zm.z = sw[czm];
            in.zp.x = su[czp], in.zp.y = sv[czp], in.zp.z = sw[czp];
            in.L.x = in.L.y = in.L.z = (ST)0;
            if (refresh) {
                // general geometry (lag % Tb != 0): the refresh rides inside the wave, as in the wavefront kernel
                const Vec4<ST> L = ld4_cg(P.L + (int64_t)frame * np + a);
                sor_refresh<C>(P.a_data, P.J + (int64_t)frame * C * 10 * np, P.wgt, np, a, (double)in.own.x,
                               (double)in.own.y, (double)in.own.z, den0, in.A);
                in.A[6] -= (double)L.x;
                in.A[7] -= (double)L.y;
                in.A[8] -= (double)L.z;
#pragma unroll
                for (int e = 0; e < 9; ++e)
                    FR3D_STCG(AB + e * 32, in.A[e]);
            }
            const Vec4<ST> o = sor_update(P, in);
            su[c0] = o.x;
            sv[c0] = o.y;
            sw[c0] = o.z;
        }
        par.sync();
    }

    par.mark(3);
    // ---- 3. write back the cells this task updated (those inside one of its regions)
    for (int c = par.tid(); c < nbox; c += par.nth()) {
        const int bk = c / FJI, rem = c - bk * FJI, bj = rem / G.FI, bi = rem - bj * G.FI;
        const int k = ok + bk, j = oj + bj, i = oi + bi;
        if (k < 0 || k >= p || j < 0 || j >= m || i < 0 || i >= n)
            continue;
        bool upd = false;
        for (int r = 0; r < nr && !upd; ++r)
            upd = k >= k0 - G.sk * r && k < k0 + G.K - G.sk * r && j >= j0 - G.sj * r && j < j0 + G.J - G.sj * r &&
                  i >= i0 - G.si * r && i < i0 + G.I - G.si * r;
        if (upd) {
            Vec4<ST> v;
            v.x = su[c];
            v.y = sv[c];
            v.z = sw[c];
            set_pad(v);
            st4_cg(dg + FR3D_TILE_SLOT(k, j, i), v);
        }
    }
    par.sync();
    par.mark(4);
#undef FR3D_TILE_SLOT
}

// All tile waves, executed by `nblocks` cooperating blocks (block = this caller).  wave_sync(w) is the grid-wide
// step between tile waves.
template <class ST, int C, class Par, class WaveSync>
FR3D_HD void sor_tile_run_block(const SorParams<ST>& P, const SorTileGeom& G, unsigned char* smem, int block, int nblocks,
                                const Par& par, WaveSync wave_sync)
{
    const SorTileTabs tabs = sor_tile_tabs<ST>(G, smem);
    ST* const sm = reinterpret_cast<ST*>(smem);
    sor_tile_build_tabs(G, tabs, par);
    for (int w = 0; w < G.nwaves; ++w) {
        sor_tile_wave_prefix(G, tabs, w, par);
        const int ntasks = tabs.tau_pref[G.ntau] * P.B;
        for (int idx = block; idx < ntasks; idx += nblocks) {
            par.mark(-1);
            if (par.tid() == 0)
                sor_tile_decode(G, tabs, w, idx);
            par.sync();
            par.mark(5);
            const int frame = tabs.task[0], ta = tabs.task[1], tb_ = tabs.task[2], tc = tabs.task[3], tau = tabs.task[4];
            sor_tile_task<ST, C>(P, G, tabs, sm, frame, ta, tb_, tc, tau, par);
        }
        par.mark(-1);
        wave_sync(w);
        par.mark(6);
    }
}

} // namespace fr3d
