// lbvh_build.cu -- LBVH builder for sm_100a: scene bounds, 63-bit Morton keys,
// a hand-written LSD radix sort (no CUB), Karras 2012 hierarchy, atomic
// bottom-up refit, and the leaf-collapsing emit of the 64-byte traversal nodes.
//
// Replaces Embree's rtcCommitScene, which Open3D's RaycastingScene runs lazily
// on the first query after add_triangles (reference call sites:
// pyQSM/viz/ray_casting.py:66,156,219,242,276,317 then :168,223,279,319).
// The key/ordering/topology arithmetic mirrors oracle/qsmrt_oracle.c
// (orc_commit) so the builder can be checked bit-for-bit on the CPU.
#include <algorithm>
#include "common.cuh"
#include "build.h"

namespace {

// ---------------------------------------------------------------- bounds
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__device__ __forceinline__ void tri_bounds(const float *__restrict__ verts, const uint32_t *__restrict__ idx,
                                           uint64_t t, float lo[3], float hi[3])
{
    uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float p0 = verts[3ull * i0 + a], p1 = verts[3ull * i1 + a], p2 = verts[3ull * i2 + a];
        lo[a] = fminf(p0, fminf(p1, p2));
        hi[a] = fmaxf(p0, fmaxf(p1, p2));
    }
}

__global__ void k_init_bounds(uint32_t *ob)
{
    if (threadIdx.x < 3) ob[threadIdx.x] = 0xFFFFFFFFu;        // running min
    else if (threadIdx.x < 6) ob[threadIdx.x] = 0u;            // running max
}

__global__ void __launch_bounds__(256)
k_scene_bounds(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t n, uint32_t *ob)
{
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        float l[3], h[3];
        tri_bounds(verts, idx, t, l, h);
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], l[a]); hi[a] = fmaxf(hi[a], h[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&ob[a], f2ord(lo[a]));
            atomicMax(&ob[3 + a], f2ord(hi[a]));
        }
    }
}

__global__ void k_finalize_bounds(const uint32_t *ob, BuildParams *bp)
{
    if (threadIdx.x != 0) return;
    float m = 0.0f;
    for (int a = 0; a < 3; ++a) {
        float lo = ord2f(ob[a]), hi = ord2f(ob[3 + a]);
        float ext = __fsub_rn(hi, lo);
        bp->slo[a] = lo; bp->shi[a] = hi;
        bp->scale[a] = ext > 0.0f ? __fdiv_rn(2097152.0f, ext) : 0.0f;
        m = fmaxf(m, fabsf(lo)); m = fmaxf(m, fabsf(hi)); m = fmaxf(m, ext);
    }
    float pad = __fmul_rn(m, 7.62939453125e-06f);   // 2^-17, see DESIGN.md "box padding"
    bp->pad = pad > 0.0f ? pad : 1e-30f;
    // quantisation grid: padded bounds plus a 32-cell margin so the +-3-cell widening never clamps
    for (int a = 0; a < 3; ++a) {
        float ext = bp->shi[a] - bp->slo[a];
        float e = bp->pad + ext * 4.8828125e-04f;       // 2^-11
        bp->glo[a] = bp->slo[a] - e;
        bp->cell[a] = (ext + 2.0f * e) / 65535.0f;
        bp->inv_cell[a] = 1.0f / bp->cell[a];
    }
    bp->leaf_diag_sum = 0.0f;
}

// ---------------------------------------------------------------- Morton
__device__ __forceinline__ uint64_t spread21(uint32_t x)
{
    uint64_t v = x & 0x1FFFFFu;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8))  & 0x100F00F00F00F00Full;
    v = (v | (v << 4))  & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2))  & 0x1249249249249249ull;
    return v;
}

__global__ void __launch_bounds__(256)
k_morton(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t n,
         const BuildParams *__restrict__ bp, uint64_t *__restrict__ keys, uint32_t *__restrict__ order)
{
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    float lo[3], hi[3];
    tri_bounds(verts, idx, t, lo, hi);
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float c = __fmul_rn(__fadd_rn(lo[a], hi[a]), 0.5f);
        float f = __fmul_rn(__fsub_rn(c, bp->slo[a]), bp->scale[a]);
        f = fminf(fmaxf(f, 0.0f), 2097151.0f);
        q[a] = (uint32_t)f;
    }
    keys[t] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
    order[t] = (uint32_t)t;
}

// ------------------------------------------------------------ radix sort
// LSD, 8-bit digits, stable.  Per pass: per-tile digit histogram -> per-digit
// exclusive scan over tiles -> ranked scatter (warp match_any multi-split).
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;     // keys per tile

__global__ void __launch_bounds__(RS_THREADS)
k_rs_tile_hist(const uint64_t *__restrict__ keys, uint64_t n, int shift, uint32_t *__restrict__ tile_hist, uint32_t ntiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        uint64_t i = base + (uint64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    tile_hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// one block per digit: exclusive scan of tile_hist[d][0..ntiles) in place; total -> digit_tot[d]
__global__ void __launch_bounds__(256)
k_rs_scan_tiles(uint32_t *__restrict__ tile_hist, uint32_t ntiles, uint32_t *__restrict__ digit_tot)
{
    __shared__ uint32_t part[256];
    uint32_t *row = tile_hist + (uint64_t)blockIdx.x * ntiles;
    uint32_t per = (ntiles + 255) / 256;
    uint32_t b = threadIdx.x * per, e = min(b + per, ntiles);
    uint32_t s = 0;
    for (uint32_t i = b; i < e; ++i) s += row[i];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 256 partials
    for (int o = 1; o < 256; o <<= 1) {
        uint32_t v = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - s;
    for (uint32_t i = b; i < e; ++i) { uint32_t v = row[i]; row[i] = run; run += v; }
    if (threadIdx.x == 255) digit_tot[blockIdx.x] = part[255];
}

__global__ void __launch_bounds__(RS_THREADS, 4)
k_rs_scatter(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
             uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, uint64_t n, int shift,
             const uint32_t *__restrict__ tile_hist, const uint32_t *__restrict__ digit_tot, uint32_t ntiles)
{
    __shared__ uint32_t wcnt[RS_WARPS][256];     // per-warp running digit counts
    __shared__ uint32_t dbase[256];              // global start of digit + this tile's offset
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    // exclusive scan of digit totals (256 values) -> global digit starts
    {
        uint32_t v = digit_tot[threadIdx.x];
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (l >= o) x += y; }
        __shared__ uint32_t wsum[RS_WARPS];
        if (l == 31) wsum[w] = x;
        __syncthreads();
        uint32_t off = 0;
        for (int k = 0; k < w; ++k) off += wsum[k];
        dbase[threadIdx.x] = off + x - v + tile_hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x];
    }
    __syncthreads();

    uint64_t kreg[RS_ROUNDS];
    uint32_t vreg[RS_ROUNDS];
    uint16_t rnk[RS_ROUNDS];
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE + (uint64_t)w * (32 * RS_ROUNDS);
    const uint32_t lt = (1u << l) - 1u;
    // all loads first (16 + 16 independent requests in flight), then the ranking rounds
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        uint64_t i = base + r * 32 + l;
        kreg[r] = i < n ? kin[i] : ~0ull;
    }
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        uint64_t i = base + r * 32 + l;
        vreg[r] = i < n ? vin[i] : 0u;
    }
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        uint64_t i = base + r * 32 + l;
        bool valid = i < n;
        uint64_t k = kreg[r];
        uint32_t d = valid ? ((uint32_t)(k >> shift) & 0xFFu) : 256u;
        // lanes holding the same digit: nine ballots (8 digit bits + validity) instead of MATCH.ANY,
        // which serialises over the distinct values of the warp
        uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
        m = valid ? m : ~m;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
            m &= ((d >> b) & 1u) ? bal : ~bal;
        }
        uint32_t before = valid ? wcnt[w][d] : 0;
        __syncwarp();
        if (valid && (m & lt) == 0) wcnt[w][d] = before + __popc(m);
        __syncwarp();
        rnk[r] = (uint16_t)(before + __popc(m & lt));
    }
    __syncthreads();
    // exclusive scan across warps, per digit (thread d owns digit d)
    {
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < RS_WARPS; ++k) { uint32_t v = wcnt[k][threadIdx.x]; wcnt[k][threadIdx.x] = run; run += v; }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        uint64_t i = base + r * 32 + l;
        if (i < n) {
            uint32_t d = (uint32_t)(kreg[r] >> shift) & 0xFFu;
            uint64_t dst = (uint64_t)dbase[d] + wcnt[w][d] + rnk[r];
            kout[dst] = kreg[r];
            vout[dst] = vreg[r];
        }
    }
}


// ---- single-pass-per-digit variant ("Onesweep", Adinets & Merrill 2022) -------
// One read of the keys builds all eight digit histograms; each pass is then ONE
// kernel: a tile ranks its keys locally, publishes its per-digit counts, finds
// its global offsets by decoupled look-back over the tiles before it, and
// scatters.  Per pass that is one read and one write of (key, value) instead
// of the three kernels / extra key read of the classic variant above.
// Tiles are handed out through an atomic counter so a tile only ever waits on
// tiles that are already running.
constexpr uint32_t OS_AGG = 1u << 30, OS_INC = 2u << 30, OS_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(256)
k_os_histogram(const uint64_t *__restrict__ keys, uint64_t n, uint32_t *__restrict__ ghist /* [8][256] */)
{
    __shared__ uint32_t h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
    __syncthreads();
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < 8; ++p) atomicAdd(&h[p][(uint32_t)(k >> (8 * p)) & 0xFFu], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += 256) { uint32_t v = (&h[0][0])[i]; if (v) atomicAdd(&ghist[i], v); }
}

// exclusive scan of each 256-bin histogram in place (one block per pass)
__global__ void __launch_bounds__(256)
k_os_scan_hist(uint32_t *__restrict__ ghist)
{
    __shared__ uint32_t wsum[8];
    uint32_t *row = ghist + blockIdx.x * 256;
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t v = row[threadIdx.x], x = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (l >= o) x += y; }
    if (l == 31) wsum[w] = x;
    __syncthreads();
    uint32_t off = 0;
    for (int k = 0; k < w; ++k) off += wsum[k];
    row[threadIdx.x] = off + x - v;
}

__global__ void __launch_bounds__(RS_THREADS, 4)
k_os_pass(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
          uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, uint64_t n, int shift,
          const uint32_t *__restrict__ gbase /* [256] digit starts of this pass */,
          uint32_t *status /* [ntiles][256], zeroed */, uint32_t *tile_counter)
{
    __shared__ uint32_t wcnt[RS_WARPS][256];
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t s_tile;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;

    uint64_t kreg[RS_ROUNDS];
    uint32_t vreg[RS_ROUNDS];
    uint16_t rnk[RS_ROUNDS];
    const uint64_t base = (uint64_t)tile * RS_TILE + (uint64_t)w * (32 * RS_ROUNDS);
    const uint32_t lt = (1u << l) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) { uint64_t i = base + r * 32 + l; kreg[r] = i < n ? kin[i] : ~0ull; }
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) { uint64_t i = base + r * 32 + l; vreg[r] = i < n ? vin[i] : 0u; }
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const bool valid = base + r * 32 + l < n;
        const uint32_t d = valid ? ((uint32_t)(kreg[r] >> shift) & 0xFFu) : 256u;
        uint32_t m = __ballot_sync(0xFFFFFFFFu, valid);
        m = valid ? m : ~m;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, (d >> b) & 1u);
            m &= ((d >> b) & 1u) ? bal : ~bal;
        }
        const uint32_t before = valid ? wcnt[w][d] : 0;
        __syncwarp();
        if (valid && (m & lt) == 0) wcnt[w][d] = before + __popc(m);
        __syncwarp();
        rnk[r] = (uint16_t)(before + __popc(m & lt));
    }
    __syncthreads();
    {   // thread d owns digit d: tile count, publish, look back, global base
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < RS_WARPS; ++k) { uint32_t v = wcnt[k][d]; wcnt[k][d] = run; run += v; }
        volatile uint32_t *st = status;
        st[(uint64_t)tile * 256 + d] = (tile == 0 ? OS_INC : OS_AGG) | run;
        uint32_t excl = 0;
        for (int64_t k = (int64_t)tile - 1; k >= 0; --k) {
            uint32_t sv;
            do { sv = st[(uint64_t)k * 256 + d]; } while ((sv >> 30) == 0u);
            excl += sv & OS_MASK;
            if ((sv >> 30) == 2u) break;
        }
        if (tile != 0) st[(uint64_t)tile * 256 + d] = OS_INC | (excl + run);
        dbase[d] = gbase[d] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        if (base + r * 32 + l < n) {
            const uint32_t d = (uint32_t)(kreg[r] >> shift) & 0xFFu;
            const uint64_t dst = (uint64_t)dbase[d] + wcnt[w][d] + rnk[r];
            kout[dst] = kreg[r];
            vout[dst] = vreg[r];
        }
    }
}

// ------------------------------------------------------------- hierarchy
__device__ __forceinline__ int delta(const uint64_t *__restrict__ k, int64_t n, int64_t i, int64_t j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t a = k[i], b = k[j];
    if (a != b) return __clzll((long long)(a ^ b));
    return 64 + __clz((int)((uint32_t)i ^ (uint32_t)j));
}

// Karras 2012: one thread per internal node; writes children (unified index
// space), parents and the covered leaf range.
__global__ void __launch_bounds__(256)
k_karras(const uint64_t *__restrict__ keys, int64_t n, BNode *__restrict__ bn,
         int32_t *__restrict__ parent, int2 *__restrict__ range)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int64_t lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int64_t l = 0;
    for (int64_t t = lmax / 2; t >= 1; t /= 2)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int64_t j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int64_t sp = 0;
    for (int64_t t = (l + 1) / 2;; t = (t + 1) / 2) {
        if (delta(keys, n, i, i + (sp + t) * d) > dnode) sp += t;
        if (t <= 1) break;
    }
    int64_t gamma = i + sp * d + (d < 0 ? -1 : 0);
    int64_t first = min(i, j), last = max(i, j);
    int32_t left  = (first == gamma)    ? (int32_t)(n - 1 + gamma)     : (int32_t)gamma;
    int32_t right = (last == gamma + 1) ? (int32_t)(n - 1 + gamma + 1) : (int32_t)(gamma + 1);
    bn[i].left = left; bn[i].right = right;
    parent[left] = (int32_t)i; parent[right] = (int32_t)i;
    if (i == 0) parent[0] = -1;
    range[i] = make_int2((int)first, (int)last);
}

__device__ __forceinline__ uint32_t geom_of(const uint64_t *__restrict__ goff, uint32_t ngeoms, uint64_t t)
{
    uint32_t lo = 0, hi = ngeoms;             // goff[g] <= t < goff[g+1]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (goff[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}

// Leaves in sorted order: padded boxes into the binary node array and the
// 48-byte triangle records the traversal kernels read.
__global__ void __launch_bounds__(256)
k_emit_leaves(const float *__restrict__ verts, const uint32_t *__restrict__ idx, int64_t n,
              const uint32_t *__restrict__ order, const uint64_t *__restrict__ goff, uint32_t ngeoms,
              BuildParams *__restrict__ bp, BNode *__restrict__ bn, TriRec *__restrict__ tris)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    float diag = 0.0f;
    if (i < n) {
        uint64_t t = order[i];
        uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
        float p0[3], p1[3], p2[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { p0[a] = verts[3ull * i0 + a]; p1[a] = verts[3ull * i1 + a]; p2[a] = verts[3ull * i2 + a]; }
        const float pad = bp->pad;
        BNode b;
        b.lox = __fsub_rn(fminf(p0[0], fminf(p1[0], p2[0])), pad);
        b.loy = __fsub_rn(fminf(p0[1], fminf(p1[1], p2[1])), pad);
        b.loz = __fsub_rn(fminf(p0[2], fminf(p1[2], p2[2])), pad);
        b.hix = __fadd_rn(fmaxf(p0[0], fmaxf(p1[0], p2[0])), pad);
        b.hiy = __fadd_rn(fmaxf(p0[1], fmaxf(p1[1], p2[1])), pad);
        b.hiz = __fadd_rn(fmaxf(p0[2], fmaxf(p1[2], p2[2])), pad);
        b.left = -1; b.right = -1;
        bn[n - 1 + i] = b;
        float dx = b.hix - b.lox, dy = b.hiy - b.loy, dz = b.hiz - b.loz;
        diag = sqrtf(dx * dx + dy * dy + dz * dz);
        uint32_t g = ngeoms > 1 ? geom_of(goff, ngeoms, t) : 0u;
        uint32_t prim = (uint32_t)(t - goff[g]);
        TriRec r;
        r.p0 = make_float4(p0[0], p0[1], p0[2], __uint_as_float(prim));
        r.p1 = make_float4(__fsub_rn(p0[0], p1[0]), __fsub_rn(p0[1], p1[1]), __fsub_rn(p0[2], p1[2]), __uint_as_float(g));
        r.p2 = make_float4(__fsub_rn(p2[0], p0[0]), __fsub_rn(p2[1], p0[1]), __fsub_rn(p2[2], p0[2]), 0.0f);
        tris[i] = r;
    }
    // mean leaf size (decides whether the 16-bit node grid is fine enough): one atomic per warp
    for (int o = 16; o > 0; o >>= 1) diag += __shfl_xor_sync(0xFFFFFFFFu, diag, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&bp->leaf_diag_sum, diag);
}

// The 32-byte nodes widen every box by 3 grid cells per side: use them when that is small against a leaf.
__global__ void k_decide_quant(BuildParams *bp, float ntris)
{
    const float max_cell = fmaxf(bp->cell[0], fmaxf(bp->cell[1], bp->cell[2]));
    bp->use_q = 6.0f * max_cell <= 0.15f * (bp->leaf_diag_sum / ntris) ? 1 : 0;
}

// Bottom-up refit: one thread per leaf climbs; the second arrival at a node
// (atomic flag) merges the two child boxes and continues.
__global__ void __launch_bounds__(256)
k_refit(int64_t n, BNode *bn, const int32_t *__restrict__ parent, uint32_t *flags, unsigned long long *counters)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t p = parent[n - 1 + i];
    uint32_t h = 0;                       // height of the subtree just finished (leaf = 0)
    while (p >= 0) {
        __threadfence();
        // the flag carries the first child's height + 1; 0 = nobody arrived yet
        uint32_t other = atomicExch(&flags[p], h + 1u);
        if (other == 0u) return;
        h = max(h, other - 1u) + 1u;
        if (p == 0) counters[2] = h;      // tree height (edges from the root to the deepest leaf)
        const float4 *cl = reinterpret_cast<const float4 *>(&bn[bn[p].left]);
        const float4 *cr = reinterpret_cast<const float4 *>(&bn[bn[p].right]);
        float4 llo = __ldcg(cl), lhi = __ldcg(cl + 1), rlo = __ldcg(cr), rhi = __ldcg(cr + 1);
        bn[p].lox = fminf(llo.x, rlo.x); bn[p].loy = fminf(llo.y, rlo.y); bn[p].loz = fminf(llo.z, rlo.z);
        bn[p].hix = fmaxf(lhi.x, rhi.x); bn[p].hiy = fmaxf(lhi.y, rhi.y); bn[p].hiz = fmaxf(lhi.z, rhi.z);
        p = parent[p];
    }
}

// Collapse subtrees of <= QSMRT_LEAF_MAX triangles into leaves and emit the
// 64-byte traversal nodes (indexed like the binary internal nodes; collapsed
// interior nodes are simply never referenced).
// conservative 16-bit encoding of one box axis: floor/ceil on the grid, widened by 3 cells
__device__ __forceinline__ uint32_t quantise_axis(float lo, float hi, float glo, float inv_cell)
{
    float fl = floorf((lo - glo) * inv_cell) - 3.0f, fh = ceilf((hi - glo) * inv_cell) + 3.0f;
    uint32_t ql = (uint32_t)fminf(fmaxf(fl, 0.0f), 65535.0f), qh = (uint32_t)fminf(fmaxf(fh, 0.0f), 65535.0f);
    return ql | (qh << 16);
}

__device__ __forceinline__ QNode quantise_node(const TNode &o, const BuildParams *__restrict__ bp)
{
    QNode q;
    q.w[0] = quantise_axis(o.a.x, o.a.y, bp->glo[0], bp->inv_cell[0]);
    q.w[1] = quantise_axis(o.a.z, o.a.w, bp->glo[1], bp->inv_cell[1]);
    q.w[2] = quantise_axis(o.c.x, o.c.y, bp->glo[2], bp->inv_cell[2]);
    q.w[3] = quantise_axis(o.b.x, o.b.y, bp->glo[0], bp->inv_cell[0]);
    q.w[4] = quantise_axis(o.b.z, o.b.w, bp->glo[1], bp->inv_cell[1]);
    q.w[5] = quantise_axis(o.c.z, o.c.w, bp->glo[2], bp->inv_cell[2]);
    q.w[6] = (uint32_t)o.d.x; q.w[7] = (uint32_t)o.d.y;
    return q;
}

__device__ __forceinline__ int child_ref(int32_t c, int64_t n, const int2 *__restrict__ range, int leaf_max)
{
    if (c >= n - 1) return ~(int)(((uint32_t)(c - (n - 1)) << 2) | 0u);
    int2 r = range[c];
    int cnt = r.y - r.x + 1;
    if (cnt <= leaf_max) return ~(int)(((uint32_t)r.x << 2) | (uint32_t)(cnt - 1));
    return c;
}

__global__ void __launch_bounds__(256)
k_emit_tnodes(int64_t n, const BNode *__restrict__ bn, const int2 *__restrict__ range,
              TNode *__restrict__ tn, QNode *__restrict__ qn, const BuildParams *__restrict__ bp,
              unsigned long long *__restrict__ counters, int leaf_max)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    bool live = i < n - 1;
    int2 r = live ? range[i] : make_int2(0, 0);
    if (live && i != 0 && (r.y - r.x + 1) <= leaf_max) live = false;     // folded into a leaf above
    // statistics: one atomic per warp, not per thread
    unsigned m = __ballot_sync(0xFFFFFFFFu, live);
    int nleaf = 0;
    if (live) {
        BNode me0 = bn[i];
        nleaf = (child_ref(me0.left, n, range, leaf_max) < 0) + (child_ref(me0.right, n, range, leaf_max) < 0);
    }
    for (int o = 16; o > 0; o >>= 1) nleaf += __shfl_xor_sync(0xFFFFFFFFu, nleaf, o);
    if ((threadIdx.x & 31) == 0 && m) {
        atomicAdd(&counters[0], (unsigned long long)__popc(m));
        atomicAdd(&counters[1], (unsigned long long)nleaf);
    }
    if (!live) return;
    BNode me = bn[i];
    BNode c0 = bn[me.left], c1 = bn[me.right];
    TNode o;
    o.a = make_float4(c0.lox, c0.hix, c0.loy, c0.hiy);
    o.b = make_float4(c1.lox, c1.hix, c1.loy, c1.hiy);
    o.c = make_float4(c0.loz, c0.hiz, c1.loz, c1.hiz);
    int r0 = child_ref(me.left, n, range, leaf_max), r1 = child_ref(me.right, n, range, leaf_max);
    o.d = make_int4(r0, r1, 0, 0);
    tn[i] = o;
    if (bp->use_q) qn[i] = quantise_node(o, bp);
}

// single-triangle scene: one node, one real child, one empty (inverted) box
__global__ void k_emit_single(const BNode *__restrict__ bn, TNode *__restrict__ tn, QNode *__restrict__ qn,
                              const BuildParams *__restrict__ bp, unsigned long long *counters)
{
    BNode c0 = bn[0];
    TNode o;
    o.a = make_float4(c0.lox, c0.hix, c0.loy, c0.hiy);
    o.b = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
    o.c = make_float4(c0.loz, c0.hiz, INFINITY, -INFINITY);
    o.d = make_int4(~0, ~0, 0, 0);
    // second child: empty box, never entered
    tn[0] = o;
    // the quantised twin cannot express an empty box: child 1 repeats child 0 (same leaf twice is harmless)
    TNode dup = o;
    dup.b = o.a; dup.c = make_float4(o.c.x, o.c.y, o.c.x, o.c.y);
    qn[0] = quantise_node(dup, bp);
    counters[0] = 1; counters[1] = 1; counters[2] = 1;
}

} // namespace

// -------------------------------------------------------------- host side
int g_sort_variant = 1;      // 0 classic (3 kernels per pass), 1 onesweep (decoupled look-back)

size_t lbvh_sort_scratch_bytes(uint64_t n)
{
    uint64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    // classic: tile_hist[256][ntiles] + digit_tot[256]; onesweep: ghist[8][256] + status[ntiles][256] + counters[8]
    return (size_t)(ntiles * 256 + 8 * 256 + 256 + 64) * sizeof(uint32_t);
}

int lbvh_radix_sort(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp,
                    uint64_t n, uint32_t *scratch, cudaStream_t st)
{
    if (n == 0) return 0;
    uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint32_t *tile_hist = scratch, *digit_tot = scratch + (uint64_t)ntiles * 256;
    uint64_t *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    if (g_sort_variant == 1 && n >= (8u << 20)) {      // measured: look-back wins from ~8M keys, the 3-kernel pass below that
        uint32_t *status = scratch, *ghist = scratch + (uint64_t)ntiles * 256, *counters = ghist + 8 * 256;
        CUDA_TRY(cudaMemsetAsync(ghist, 0, (8 * 256 + 8) * sizeof(uint32_t), st));
        k_os_histogram<<<(unsigned)std::min<uint64_t>((n + 4095) / 4096, 148 * 8), 256, 0, st>>>(keys, n, ghist);
        k_os_scan_hist<<<8, 256, 0, st>>>(ghist);
        for (int pass = 0; pass < 8; ++pass) {
            CUDA_TRY(cudaMemsetAsync(status, 0, (size_t)ntiles * 256 * sizeof(uint32_t), st));
            k_os_pass<<<ntiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, pass * 8, ghist + pass * 256, status, counters + pass);
            uint64_t *tk = kin; kin = kout; kout = tk;
            uint32_t *tv = vin; vin = vout; vout = tv;
        }
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    for (int pass = 0; pass < 8; ++pass) {
        int shift = pass * 8;
        k_rs_tile_hist<<<ntiles, RS_THREADS, 0, st>>>(kin, n, shift, tile_hist, ntiles);
        k_rs_scan_tiles<<<256, 256, 0, st>>>(tile_hist, ntiles, digit_tot);
        k_rs_scatter<<<ntiles, RS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, tile_hist, digit_tot, ntiles);
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;   // 8 passes: sorted data is back in keys / vals
}

int lbvh_build(const LbvhBuildArgs &A, cudaStream_t st)
{
    const uint64_t n = A.ntris;
    const int B = 256;
    const unsigned gN = (unsigned)((n + B - 1) / B);
    k_init_bounds<<<1, 32, 0, st>>>(A.bounds_ord);
    k_scene_bounds<<<min(gN, 148u * 8u), B, 0, st>>>(A.verts, A.idx, n, A.bounds_ord);
    k_finalize_bounds<<<1, 32, 0, st>>>(A.bounds_ord, A.params);
    k_morton<<<gN, B, 0, st>>>(A.verts, A.idx, n, A.params, A.keys, A.order);
    CUDA_TRY(cudaGetLastError());
    if (A.ev_sort0) CUDA_TRY(cudaEventRecord(A.ev_sort0, st));
    if (lbvh_radix_sort(A.keys, A.keys_tmp, A.order, A.order_tmp, n, A.sort_scratch, st)) return 1;
    if (A.ev_sort1) CUDA_TRY(cudaEventRecord(A.ev_sort1, st));
    k_emit_leaves<<<gN, B, 0, st>>>(A.verts, A.idx, (int64_t)n, A.order, A.geom_offsets, A.ngeoms, A.params, A.bnodes, A.tris);
    k_decide_quant<<<1, 1, 0, st>>>(A.params, (float)n);
    CUDA_TRY(cudaMemsetAsync(A.counters, 0, 3 * sizeof(unsigned long long), st));
    if (n == 1) {
        k_emit_single<<<1, 1, 0, st>>>(A.bnodes, A.tnodes, A.qnodes, A.params, A.counters);
    } else {
        const unsigned gI = (unsigned)((n - 1 + B - 1) / B);
        CUDA_TRY(cudaMemsetAsync(A.flags, 0, (n - 1) * sizeof(uint32_t), st));
        k_karras<<<gI, B, 0, st>>>(A.keys, (int64_t)n, A.bnodes, A.parent, A.range);
        k_refit<<<gN, B, 0, st>>>((int64_t)n, A.bnodes, A.parent, A.flags, A.counters);
        k_emit_tnodes<<<gI, B, 0, st>>>((int64_t)n, A.bnodes, A.range, A.tnodes, A.qnodes, A.params, A.counters, A.leaf_max);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}
