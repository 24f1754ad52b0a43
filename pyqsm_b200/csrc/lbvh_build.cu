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
#include <cmath>
#include <cstring>
#include "common.cuh"
#include "build.h"

namespace {

// ---------------------------------------------------------------- bounds
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
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

// Registration-time statistics of one geometry (qsmrt_add_triangles): the largest vertex index (Embree would read
// out of bounds; we reject) and the bounds over the vertices its triangles reference, in one pass over the index
// array.  out[0..2] running min, out[3..5] running max (ordered-uint encoding), out[6] largest index.  The commit
// then only combines the geometries' bounds on the host: no bounds kernels in the build.
__global__ void __launch_bounds__(256)
k_geometry_stats(const float *__restrict__ verts, uint64_t nverts, const uint32_t *__restrict__ idx, uint64_t n, uint32_t *out)
{
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    uint32_t mx = 0;
#pragma unroll 4
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
        mx = max(mx, max(i0, max(i1, i2)));
        if (i0 < nverts && i1 < nverts && i2 < nverts) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float p0 = verts[3ull * i0 + a], p1 = verts[3ull * i1 + a], p2 = verts[3ull * i2 + a];
                lo[a] = fminf(lo[a], fminf(p0, fminf(p1, p2)));
                hi[a] = fmaxf(hi[a], fmaxf(p0, fmaxf(p1, p2)));
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], o));
        }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    __shared__ uint32_t sb[7];
    if (threadIdx.x < 3) sb[threadIdx.x] = 0xFFFFFFFFu; else if (threadIdx.x < 7) sb[threadIdx.x] = 0u;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&sb[a], f2ord(lo[a])); atomicMax(&sb[3 + a], f2ord(hi[a])); }
        atomicMax(&sb[6], mx);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(&out[threadIdx.x], sb[threadIdx.x]);
    else if (threadIdx.x < 7) atomicMax(&out[threadIdx.x], sb[threadIdx.x]);
}

// ------------------------------------------------------- sliver splitting
// An LBVH keys every primitive by the centre of its box, so long thin triangles at arbitrary tilt -- the 1.5 m x 8 cm
// side triangles of create_cylinder trunks in a cylinder QSM -- get huge, mostly empty, heavily overlapping boxes:
// on the C1 tree one ray through the trunk tested 813 boxes and 239 triangles (oracle counters), and a 1M-ray launch
// spent most of its 0.37 ms waiting for a handful of such rays (profiles/r02_c1_tail.txt).  "Early split clipping"
// (Ernst & Greiner 2007): a triangle whose aspect L^2 / 2A (longest edge L, area A) exceeds split_aspect is cut into
// K = min(split_max, floor(aspect / split_aspect)) slabs across its longest edge and enters the build as K
// REFERENCES, each with the tight box of its slab.  The triangle record is duplicated per reference, so traversal
// is unchanged, and every query already treats two hits of one triangle as one (equal t, equal primitive).
struct __align__(16) RefBox { float lo[3]; uint32_t tri; float hi[3]; uint32_t pad_; };

__device__ __forceinline__ void load_corners(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t t,
                                             float A[3], float B[3], float Cc[3])
{
    const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) { A[a] = verts[3ull * i0 + a]; B[a] = verts[3ull * i1 + a]; Cc[a] = verts[3ull * i2 + a]; }
}

// number of slabs for the triangle (A, B, C); on return the corners are rotated so that AB is the longest edge
__device__ __forceinline__ int split_factor(float A[3], float B[3], float C[3], int split_max, float split_aspect)
{
    if (split_max <= 1) return 1;
    float ab = 0.f, bc = 0.f, ca = 0.f, n[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { ab += (B[a] - A[a]) * (B[a] - A[a]); bc += (C[a] - B[a]) * (C[a] - B[a]); ca += (A[a] - C[a]) * (A[a] - C[a]); }
    if (bc >= ab && bc >= ca) {            // rotate: (A, B, C) <- (B, C, A)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float t_ = A[a]; A[a] = B[a]; B[a] = C[a]; C[a] = t_; }
        ab = bc;
    } else if (ca >= ab && ca >= bc) {     // (A, B, C) <- (C, A, B)
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float t_ = C[a]; C[a] = B[a]; B[a] = A[a]; A[a] = t_; }
        ab = ca;
    }
    n[0] = (B[1] - A[1]) * (C[2] - A[2]) - (B[2] - A[2]) * (C[1] - A[1]);
    n[1] = (B[2] - A[2]) * (C[0] - A[0]) - (B[0] - A[0]) * (C[2] - A[2]);
    n[2] = (B[0] - A[0]) * (C[1] - A[1]) - (B[1] - A[1]) * (C[0] - A[0]);
    const float area2 = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);        // 2 x area
    if (!(area2 > 0.0f) || !(ab > 0.0f)) return 1;
    const float k = floorf(ab / area2 / split_aspect);                          // L^2 / 2A / split_aspect
    return k >= (float)split_max ? split_max : (k >= 2.0f ? (int)k : 1);
}

__global__ void __launch_bounds__(256)
k_split_count(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t n, int split_max, float split_aspect,
              int32_t *__restrict__ counts, unsigned long long *total)
{
    unsigned long long mine = 0;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        float A[3], B[3], C[3];
        load_corners(verts, idx, t, A, B, C);
        const int k = split_factor(A, B, C, split_max, split_aspect);
        if (counts) counts[t] = k;
        mine += (unsigned long long)k;
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
    __shared__ unsigned long long ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long s_ = 0; for (int k = 0; k < 8; ++k) s_ += ws[k]; if (s_) atomicAdd(total, s_); }
}

// slab j of K across the longest edge AB (apex C projects to sc in [0, L]): the cross-section at parameter s joins
// A + s u on AB with the point at s on AC (s <= sc) or CB (s >= sc); the slab is the hull of its two cross-sections
// (plus C when sc lies inside), so its box is the min / max of at most five points
__global__ void __launch_bounds__(256)
k_split_emit(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t n, int split_max, float split_aspect,
             const int64_t *__restrict__ ref_off, RefBox *__restrict__ refs)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    float A[3], B[3], C[3];
    load_corners(verts, idx, t, A, B, C);
    const int64_t o = ref_off[t];
    const int K = (int)(ref_off[t + 1] - o);
    if (K <= 1) {
        RefBox r; r.tri = (uint32_t)t; r.pad_ = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) { r.lo[a] = fminf(A[a], fminf(B[a], C[a])); r.hi[a] = fmaxf(A[a], fmaxf(B[a], C[a])); }
        refs[o] = r;
        return;
    }
    split_factor(A, B, C, split_max, split_aspect);         // rotates the corners: AB is the longest edge
    float u[3], L = 0.f, sc = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) { u[a] = B[a] - A[a]; L += u[a] * u[a]; }
    L = sqrtf(L);
#pragma unroll
    for (int a = 0; a < 3; ++a) { u[a] /= L; sc += (C[a] - A[a]) * u[a]; }
    sc = fminf(fmaxf(sc, 0.0f), L);
    for (int j = 0; j < K; ++j) {
        const float s0 = L * (float)j / (float)K, s1 = L * (float)(j + 1) / (float)K;
        RefBox r; r.tri = (uint32_t)t; r.pad_ = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) { r.lo[a] = INFINITY; r.hi[a] = -INFINITY; }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const float sv = e ? s1 : s0;
            // the cross-section's end on the far boundary: on AC for sv <= sc, on CB beyond
            float f, P0[3], P1[3];
            if (sv <= sc) { f = sc > 0.0f ? sv / sc : 1.0f;
#pragma unroll
                for (int a = 0; a < 3; ++a) { P0[a] = A[a]; P1[a] = C[a]; } }
            else { f = (L - sc) > 0.0f ? (sv - sc) / (L - sc) : 0.0f;
#pragma unroll
                for (int a = 0; a < 3; ++a) { P0[a] = C[a]; P1[a] = B[a]; } }
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float onab = A[a] + sv * u[a], far = P0[a] + f * (P1[a] - P0[a]);
                r.lo[a] = fminf(r.lo[a], fminf(onab, far)); r.hi[a] = fmaxf(r.hi[a], fmaxf(onab, far));
            }
        }
        if (sc > s0 && sc < s1) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { r.lo[a] = fminf(r.lo[a], C[a]); r.hi[a] = fmaxf(r.hi[a], C[a]); }
        }
        // the slab lies inside the triangle: never let rounding push its box outside the triangle's own box, and
        // widen by the slack of the interpolation (a few ulp of the coordinates; the leaf padding comes on top)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float tl = fminf(A[a], fminf(B[a], C[a])), th = fmaxf(A[a], fmaxf(B[a], C[a]));
            const float e = 4.76837158e-07f * fmaxf(fabsf(tl), fabsf(th));      // 2^-21
            r.lo[a] = fmaxf(r.lo[a] - e, tl); r.hi[a] = fminf(r.hi[a] + e, th);
        }
        refs[o + j] = r;
    }
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

// Morton keys + (fused) the digit histograms of all radix passes: the key is histogrammed as it is produced, which
// saves the separate histogram kernel and its re-read of the keys (shift0 / npass as in the sort; ghist zeroed).
__global__ void __launch_bounds__(256)
k_morton(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t n,
         const BuildParams bpv /* by value: finished on the host */, BuildParams *__restrict__ bp_out, float *__restrict__ diag_sum,
         uint64_t *__restrict__ keys, uint32_t *__restrict__ order,
         uint32_t *__restrict__ ghist /* [npass][256], may be null */, int shift0, int npass,
         const RefBox *__restrict__ refs /* non-null: n references (split slivers) instead of n triangles */)
{
    __shared__ float wsum[8];
    __shared__ uint32_t h[8][256];
    if (ghist) { for (int i = threadIdx.x; i < npass * 256; i += 256) (&h[0][0])[i] = 0; __syncthreads(); }
    // the device copy of the parameters the later kernels read: written here instead of by a separate upload
    if (blockIdx.x == 0 && threadIdx.x == 0) *bp_out = bpv;
    const BuildParams *bp = &bpv;
    const float pad = bp->pad;
    float diag = 0.0f;
#pragma unroll 2
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n; t += (uint64_t)gridDim.x * blockDim.x) {
        float lo[3], hi[3];
        if (refs) {
            const RefBox r = refs[t];
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo[a] = r.lo[a]; hi[a] = r.hi[a]; }
        } else tri_bounds(verts, idx, t, lo, hi);
        uint32_t q[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float c = __fmul_rn(__fadd_rn(lo[a], hi[a]), 0.5f);
            float f = __fmul_rn(__fsub_rn(c, bp->slo[a]), bp->scale[a]);
            f = fminf(fmaxf(f, 0.0f), 2097151.0f);
            q[a] = (uint32_t)f;
        }
        const uint64_t key = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
        keys[t] = key;
        order[t] = (uint32_t)t;
        if (ghist) {
            const uint64_t k = key >> shift0;
#pragma unroll
            for (int p = 0; p < 8; ++p) if (p < npass) atomicAdd(&h[p][(uint32_t)(k >> (8 * p)) & 0xFFu], 1u);
        }
        // diagonal of the padded leaf box, as k_leaves_refit_emit will write it
        float dx = __fadd_rn(hi[0], pad) - __fsub_rn(lo[0], pad), dy = __fadd_rn(hi[1], pad) - __fsub_rn(lo[1], pad),
              dz = __fadd_rn(hi[2], pad) - __fsub_rn(lo[2], pad);
        diag += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    // mean leaf size (decides whether the 16-bit node grid is fine enough): one atomic per block -- one per warp
    // of a 10M-triangle launch is 312k atomics on one address, ~0.45 ms of serialised L2 traffic
    for (int o = 16; o > 0; o >>= 1) diag += __shfl_xor_sync(0xFFFFFFFFu, diag, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = diag;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int k = 0; k < 8; ++k) s += wsum[k];
        atomicAdd(diag_sum, s);
    }
    if (ghist)      // wsum's barrier above ordered the shared-memory atomics of all warps
        for (int i = threadIdx.x; i < npass * 256; i += 256) { const uint32_t v = (&h[0][0])[i]; if (v) atomicAdd(&ghist[i], v); }
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
// The tile is first sorted by digit in shared memory, so the global stores of
// a warp are runs of consecutive addresses (one run per digit present) rather
// than 32 unrelated 8-byte stores.  Tiles are handed out through an atomic
// counter so a tile only ever waits on tiles that are already running.
constexpr uint32_t OS_AGG = 1u << 30, OS_INC = 2u << 30, OS_MASK = (1u << 30) - 1u;
constexpr int OS_THREADS = 256;
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_ITEMS = 12;                         // keys per thread
constexpr int OS_CTAS = 5;                           // resident CTAs per SM the kernel is shaped for (registers <= 51, shared <= 45.6 KB)
constexpr int OS_TILE = OS_THREADS * OS_ITEMS;       // 3072 keys: 36 KB staged + 8 KB counters -> 4 CTAs / SM

// After the passes over the top 32 / 40 bits: order every run of keys that agree in those bits by (whole key, index) --
// exactly what the three skipped low passes of the stable sort would have produced -- while copying to the
// final buffers.  Runs are short (two triangles of one leaf, duplicates); a run longer than SORT_MAX_RUN raises
// *overflow and the host repeats the build with all eight passes.
constexpr int SORT_MAX_RUN = 64;
// passes over the top bits of the 63-bit keys: 5 (bits 23 .. 62) in general, 4 (bits 31 .. 62) for scenes of up to
// 4M triangles, whose keys are almost all distinct in their top 32 bits already -- the fix-up orders what is left
// either way, and a run of more than SORT_MAX_RUN keys falls back to the full 8-pass sort
inline int sort_passes(uint64_t n) { return n <= (4ull << 20) ? 4 : 5; }
inline int sort_shift0(uint64_t n) { return 63 - 8 * sort_passes(n); }

__global__ void __launch_bounds__(256)
k_sort_fixup(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
             uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, int64_t n, unsigned long long *overflow, int SORT_TOP_SHIFT)
{
    const int64_t i = blockIdx.x * 256ll + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = kin[i];
    const uint32_t v = vin[i];
    const uint64_t top = k >> SORT_TOP_SHIFT;
    int64_t a = i, b = i;
    while (a > 0 && i - a <= SORT_MAX_RUN && (kin[a - 1] >> SORT_TOP_SHIFT) == top) --a;
    while (b + 1 < n && b - i <= SORT_MAX_RUN && (kin[b + 1] >> SORT_TOP_SHIFT) == top) ++b;
    int64_t dst = i;
    if (i - a > SORT_MAX_RUN || b - i > SORT_MAX_RUN) {
        *overflow = 1ull;
    } else if (a != b) {
        int rank = 0;
        for (int64_t m = a; m <= b; ++m) {
            const uint64_t km = kin[m];
            rank += (km < k) || (km == k && vin[m] < v);
        }
        dst = a + rank;
    }
    kout[dst] = k;
    vout[dst] = v;
}

__global__ void __launch_bounds__(OS_THREADS, OS_CTAS)
k_os_pass(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
          uint64_t *__restrict__ kout, uint32_t *__restrict__ vout, uint64_t n, int shift,
          const uint32_t *__restrict__ ghist /* [256] digit histogram of this pass over all keys */,
          uint32_t *status /* [ntiles][256], zeroed */, uint32_t *tile_counter)
{
    __shared__ uint64_t skey[OS_TILE];
    __shared__ uint32_t sval[OS_TILE];
    __shared__ uint32_t wcnt[OS_WARPS][256];     // per-warp digit counts, then exclusive offsets inside the tile
    __shared__ uint32_t wsum[OS_WARPS], gsum[OS_WARPS];
    __shared__ uint32_t s_tile;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    for (int i = threadIdx.x; i < OS_WARPS * 256; i += OS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint64_t tbase = (uint64_t)tile * OS_TILE;
    const uint32_t tcount = (uint32_t)min((uint64_t)OS_TILE, n - tbase);

    uint64_t kreg[OS_ITEMS];
    uint16_t rnk[OS_ITEMS];
    const uint32_t wbase = (uint32_t)w * (32 * OS_ITEMS);
    const uint32_t lt = (1u << l) - 1u;
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) { uint32_t j = wbase + r * 32 + l; kreg[r] = j < tcount ? kin[tbase + j] : ~0ull; }
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const bool valid = wbase + r * 32 + l < tcount;
        const uint32_t d = valid ? ((uint32_t)(kreg[r] >> shift) & 0xFFu) : 256u;
        // lanes holding the same digit: nine ballots (8 digit bits + validity) instead of MATCH.ANY,
        // which serialises over the distinct values of the warp
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
    uint32_t gofs_d;        // global address of digit d's first key of this tile, minus its offset in the staged tile
    {   // thread d owns digit d: tile count, publish, offsets inside the tile, look back, global base
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < OS_WARPS; ++k) { uint32_t v = wcnt[k][d]; wcnt[k][d] = run; run += v; }
        volatile uint32_t *st = status;
        st[(uint64_t)tile * 256 + d] = (tile == 0 ? OS_INC : OS_AGG) | run;
        // exclusive scan of the tile's digit counts -> where each digit starts in the staged tile
        // ... and, with the same shuffles, of the pass' global digit histogram -> where each digit starts in the output
        const uint32_t gcount = ghist[d];
        uint32_t x = run, gx = gcount;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o), gy = __shfl_up_sync(0xFFFFFFFFu, gx, o);
            if (l >= o) { x += y; gx += gy; }
        }
        if (l == 31) { wsum[w] = x; gsum[w] = gx; }
        __syncthreads();
        uint32_t tstart = x - run, gstart = gx - gcount;
        for (int k = 0; k < w; ++k) { tstart += wsum[k]; gstart += gsum[k]; }
#pragma unroll
        for (int k = 0; k < OS_WARPS; ++k) wcnt[k][d] += tstart;
        // Decoupled look-back, one predecessor per read.  %globaltimer stamps (2M keys, one wave of 652 tiles): a tile
        // lives ~23 us = 7 us load + ranking, 10 us here, 2.3 us staging, 2.3 us scatter.  Measured and NOT faster:
        // 4/8/16 predecessors per round trip (volatile or relaxed.gpu reads), values fetched before the look-back,
        // a ninth warp that publishes a shared-memory histogram early and walks back while the others rank.
        uint32_t excl = 0;
        for (int64_t k = (int64_t)tile - 1; k >= 0; --k) {
            uint32_t sv;
            do { sv = st[(uint64_t)k * 256 + d]; } while ((sv >> 30) == 0u);
            excl += sv & OS_MASK;
            if ((sv >> 30) == 2u) break;
        }
        if (tile != 0) st[(uint64_t)tile * 256 + d] = OS_INC | (excl + run);
        gofs_d = gstart + excl - tstart;
    }
    __syncthreads();
    // stage: keys (and their values, read now) to their place in the digit-sorted tile
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const uint32_t j = wbase + r * 32 + l;
        if (j < tcount) {
            const uint32_t d = (uint32_t)(kreg[r] >> shift) & 0xFFu;
            const uint32_t pos = wcnt[w][d] + rnk[r];
            skey[pos] = kreg[r];
            sval[pos] = vin[tbase + j];
        }
    }
    __syncthreads();
    uint32_t *const gofs = &wcnt[0][0];             // the per-warp offsets are dead: their space takes the digit bases
    gofs[threadIdx.x] = gofs_d;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < OS_ITEMS; ++r) {
        const uint32_t j = r * OS_THREADS + threadIdx.x;
        if (j < tcount) {
            const uint64_t k = skey[j];
            const uint32_t dst = gofs[(uint32_t)(k >> shift) & 0xFFu] + j;      // mod 2^32: gofs may have wrapped below 0
            kout[dst] = k;
            vout[dst] = sval[j];
        }
    }
}

// ------------------------------------------------------------- hierarchy
// Common-prefix length of the augmented keys (key, index) of the adjacent sorted leaves j and j + 1
// (Karras 2012, section 4: ties broken by the index); -1 outside the array.
__device__ __forceinline__ int delta_adjacent(uint64_t ka, uint64_t kb, int64_t j)
{
    if (ka != kb) return __clzll((long long)(ka ^ kb));
    return 64 + __clz((int)((uint32_t)j ^ (uint32_t)(j + 1)));
}

__device__ __forceinline__ uint32_t geom_of(const uint64_t *__restrict__ goff, uint32_t ngeoms, uint64_t t)
{
    uint32_t lo = 0, hi = ngeoms;             // goff[g] <= t < goff[g+1]
    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (goff[mid] <= t) lo = mid; else hi = mid; }
    return lo;
}

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

// The hierarchy, the refit and all node output in ONE kernel, one thread per sorted leaf.
//
// Topology bottom-up (Apetrei 2014, "Fast and simple agglomerative LBVH construction"): a thread holds a
// finished subtree over the sorted leaves [l, r] together with its box and the prefix lengths dl = delta(l-1, l),
// dr = delta(r, r+1) to its two outside neighbours.  The subtree's parent joins it with the neighbour it shares
// the LONGER prefix with (dl == dr is impossible for sorted distinct augmented keys): dr > dl -> it is the left
// child of the node whose split lies between r and r+1, else the right child of the split between l-1 and l.
// That is the same binary radix tree Karras' top-down search produces, and its numbering follows too: Karras
// gives a left child the index of its LAST leaf and a right child the index of its FIRST leaf (root 0), so a
// node learns its own index the moment it knows which side its parent lies on -- no search, no parent / range
// arrays, no separate hierarchy kernel.
//   * warp phase: while the sibling subtree is held by a lane of the same warp (the lane of its first leaf), the
//     left lane takes the sibling's box, far end and far delta with shuffles -- no loads, atomics or fences
//     (about 5 of 6 internal nodes end here);
//   * block phase: a parent whose whole leaf range lies inside this block's RF_BLOCK leaves is joined through
//     shared memory (one 32-bit exchange on the split's slot, the first arrival's box parked beside it).  Whether
//     the range stays inside the block is a property of the PARENT that both children evaluate alike: the parent
//     of split s ends at the first delta below delta(s, s+1) on either side, so a prefix / suffix minimum of the
//     block's deltas answers it in O(1).  With 512 leaves per block ~98 % of all nodes finish in the two phases;
//   * global phase: a 64-bit exchange on the split's flag hands the first arrival's far end, far delta and
//     height to the second, which reads the sibling's box (one 32-byte read from L2) and goes on;
//   * output: the merging threads are few and scattered (3 of 32 lanes in the warp phase, 1-2 later; ncu: 70 % of
//     the kernel's instructions ran at <= 4 lanes when they wrote the nodes themselves), so a merge only parks
//     both child boxes in the shared-memory slot of the node's index; after the last phase every thread writes
//     the node whose index is its own leaf number: the 64-byte traversal node and (when the scene uses them) its
//     32-byte quantised twin.  Subtrees of <= leaf_max triangles collapse into leaves: such a node is never
//     referenced and not written.  The binary nodes (BNode) are only the hand-over medium of the global phase:
//     a thread stores its subtree's box right before the exchange, unless keep_bn asks for the complete array
//     (qsmrt_debug_get_build, the builder-vs-oracle test).
struct RefitOut {
    BNode *bn; TNode *tn; QNode *qn; const BuildParams *bp;
    int64_t n; int leaf_max; bool use_q, keep_bn;
};

#ifndef QSMRT_RF_BLOCK
#define QSMRT_RF_BLOCK 256
#endif
#ifndef QSMRT_RF_MINB
#define QSMRT_RF_MINB 5
#endif
constexpr int RF_BLOCK = QSMRT_RF_BLOCK;      // leaves per block (A/B: make EXTRA=-DQSMRT_RF_BLOCK=128)
constexpr int RF_WARPS = RF_BLOCK / 32;
constexpr int RF_BIG = 1 << 20;          // "no delta here" for the minima (deltas are -1 .. 96)

// A finished merge [l .. g | g+1 .. r] -> node `id`, waiting for the dense output pass at the end of the kernel:
// slot id - (first leaf of the block), structure of arrays so that the pass reads without bank conflicts.
struct RefitQueue {
    float (*f)[RF_BLOCK];       // [12] child boxes: left lo xyz, left hi xyz, right lo xyz, right hi xyz
    int   (*i)[RF_BLOCK];       // [3]  l, g (-1: empty slot), r
    int   first, last;          // leaves of this block
};

// 64-byte traversal node + quantised twin of node `id` = [l .. g | g+1 .. r]
__device__ __forceinline__ void emit_node(const RefitOut &R, int32_t id, int l, int g, int r,
                                          const float4 llo, const float4 lhi, const float4 rlo, const float4 rhi,
                                          unsigned &n_nodes, unsigned &n_leafrefs)
{
    const int32_t left = l == g ? (int32_t)(R.n - 1) + g : g;
    const int32_t right = g + 1 == r ? (int32_t)(R.n - 1) + g + 1 : g + 1;
    const int cl = g - l + 1, cr = r - g;
    const int r0 = cl <= R.leaf_max ? ~(int)(((uint32_t)l << 2) | (uint32_t)(cl - 1)) : left;
    const int r1 = cr <= R.leaf_max ? ~(int)(((uint32_t)(g + 1) << 2) | (uint32_t)(cr - 1)) : right;
    TNode o;
    o.a = make_float4(llo.x, lhi.x, llo.y, lhi.y);
    o.b = make_float4(rlo.x, rhi.x, rlo.y, rhi.y);
    o.c = make_float4(llo.z, lhi.z, rlo.z, rhi.z);
    o.d = make_int4(r0, r1, 0, 0);
    R.tn[id] = o;
    if (R.use_q) R.qn[id] = quantise_node(o, R.bp);
    ++n_nodes; n_leafrefs += (r0 < 0) + (r1 < 0);
}

// Record the merge that made node `id` = [l .. g | g+1 .. r] out of the child boxes (llo, lhi) and (rlo, rhi);
// returns the node's box in mlo / mhi.  The merging threads are few and scattered over their warps, so they only
// park the operands; the node records are computed and stored by all threads together at the end of the kernel.
__device__ __forceinline__ void refit_merge(const RefitOut &R, const RefitQueue &Q, int32_t id, int l, int g, int r,
                                            const float4 llo, const float4 lhi, const float4 rlo, const float4 rhi,
                                            float4 &mlo, float4 &mhi, unsigned &n_nodes, unsigned &n_leafrefs)
{
    if (id == 0 || r - l + 1 > R.leaf_max) {        // smaller subtrees collapse into a leaf: never referenced
        if (id >= Q.first && id <= Q.last) {
            const int k = id - Q.first;
            Q.f[0][k] = llo.x; Q.f[1][k] = llo.y; Q.f[2][k] = llo.z; Q.f[3][k] = lhi.x; Q.f[4][k] = lhi.y; Q.f[5][k] = lhi.z;
            Q.f[6][k] = rlo.x; Q.f[7][k] = rlo.y; Q.f[8][k] = rlo.z; Q.f[9][k] = rhi.x; Q.f[10][k] = rhi.y; Q.f[11][k] = rhi.z;
            Q.i[0][k] = l; Q.i[1][k] = g; Q.i[2][k] = r;
        } else {
            emit_node(R, id, l, g, r, llo, lhi, rlo, rhi, n_nodes, n_leafrefs);     // a node numbered in another block
        }
    }
    mlo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
    mhi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
    if (R.keep_bn) {
        mlo.w = __int_as_float(l == g ? (int32_t)(R.n - 1) + g : g);
        mhi.w = __int_as_float(g + 1 == r ? (int32_t)(R.n - 1) + g + 1 : g + 1);
        reinterpret_cast<float4 *>(&R.bn[id])[0] = mlo;
        reinterpret_cast<float4 *>(&R.bn[id])[1] = mhi;
    }
}

// Global phase of the hierarchy: climb from the finished subtree [l, r] while its sibling has already arrived.
// A 64-bit exchange on the split's flag hands the first arrival's far end, far delta and height to the second,
// which reads the sibling's box (one 32-byte read from L2), records the parent and goes on; the first arrival
// stops.  Nobody ever waits, so the order in which subtrees are processed does not matter.
struct ClimbItem { int4 a; float4 b, c; };       // l, r, dl, dr | box lo, height | box hi

__device__ __forceinline__ void global_climb(const RefitOut &R, const RefitQueue &Q, unsigned long long *flags,
                                             unsigned long long *counters, int l, int r, int dl, int dr, uint32_t h,
                                             float4 mlo, float4 mhi, unsigned &n_nodes, unsigned &n_leafrefs)
{
    const int64_t n = R.n;
    BNode *bn = R.bn;
    const bool keep_bn = R.keep_bn;
    for (;;) {
        const bool go_right = dr > dl;
        const int s = go_right ? r : l - 1;
        if (!keep_bn) {     // park my box where the sibling's thread will look for it
            float4 *me = reinterpret_cast<float4 *>(&bn[l == r ? (int)(n - 1) + l : (go_right ? r : l)]);
            me[0] = mlo; me[1] = mhi;
        }
        // Release exchange (MEMBAR.ALL + ATOMG): my binary node is in L2 before the flag.  __threadfence() would
        // be MEMBAR.SC plus an L1 invalidate (CCTL.IVALL) per level, which the __ldcg sibling read does not need.
        const unsigned long long mine = (unsigned long long)(uint32_t)((go_right ? l : r) + 1)
                                      | ((unsigned long long)h << 32) | ((unsigned long long)(uint32_t)((go_right ? dl : dr) + 1) << 40);
        unsigned long long other;
        asm volatile("atom.release.gpu.global.exch.b64 %0, [%1], %2;" : "=l"(other) : "l"(flags + s), "l"(mine) : "memory");
        if (other == 0ull) break;                 // first to arrive: the sibling's thread takes over
        const int o_end = (int)(uint32_t)(other & 0xFFFFFFFFull) - 1;
        const uint32_t o_h = (uint32_t)(other >> 32) & 0xFFu;
        const int o_d = (int)((other >> 40) & 0xFFu) - 1;
        // sibling's index: a right sibling [s+1, o_end] is a right child (index = its first leaf), a left
        // sibling [o_end, s] a left child (index = its last leaf); single leaves live at n-1+leaf
        const int sib = go_right ? (o_end == s + 1 ? (int)(n - 1) + s + 1 : s + 1) : (o_end == s ? (int)(n - 1) + s : s);
        const float4 *cs = reinterpret_cast<const float4 *>(&bn[sib]);
        const float4 slo = __ldcg(cs), shi = __ldcg(cs + 1);
        h = max(h, o_h) + 1u;
        if (go_right) { r = o_end; dr = o_d; } else { l = o_end; dl = o_d; }
        const bool root = l == 0 && r == (int)(n - 1);
        const int32_t id = root ? 0 : (dr > dl ? r : l);
        if (root) counters[2] = h;
        if (go_right) refit_merge(R, Q, id, l, s, r, mlo, mhi, slo, shi, mlo, mhi, n_nodes, n_leafrefs);
        else          refit_merge(R, Q, id, l, s, r, slo, shi, mlo, mhi, mlo, mhi, n_nodes, n_leafrefs);
        if (root) break;
    }
}

__global__ void __launch_bounds__(RF_BLOCK, QSMRT_RF_MINB)
k_hierarchy_refit_emit(const float *__restrict__ verts, const uint32_t *__restrict__ idx, int64_t n,
                       const uint64_t *__restrict__ keys, uint32_t *order, const RefBox *__restrict__ refs,
                       const uint64_t *__restrict__ goff, uint32_t ngeoms,
                       BuildParams *bp, BNode *bn, TriRec *__restrict__ tris,
                       unsigned long long *flags, TNode *__restrict__ tn, QNode *__restrict__ qn,
                       unsigned long long *counters, int leaf_max, int keep_bn,
                       ClimbItem *__restrict__ work, unsigned *work_count, unsigned work_cap, float quant_frac)
{
    // block phase state: slot k belongs to the split between leaves b0 + k and b0 + k + 1
    __shared__ uint32_t s_flag[RF_BLOCK];
    __shared__ float s_box[2][RF_BLOCK][6];       // [0]: box parked by the left child of the split, [1]: by the right child
    __shared__ int s_pmin[RF_BLOCK], s_smin[RF_BLOCK];
    __shared__ int s_wp[RF_WARPS], s_ws[RF_WARPS];
    __shared__ float s_qf[12][RF_BLOCK];
    __shared__ int s_qi[3][RF_BLOCK];

    const unsigned FULL = 0xFFFFFFFFu;
    const int64_t b0 = blockIdx.x * (int64_t)RF_BLOCK;
    const int64_t i = b0 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w0 = (int)(i - lane);               // first leaf of this warp
    const int bfirst = (int)b0, blast = (int)min(b0 + RF_BLOCK, n) - 1;      // this block's leaves
    unsigned n_nodes = 0, n_leafrefs = 0;
    // The 32-byte nodes widen every box by 3 grid cells per side: use them when that is small against a leaf
    // (mean leaf diagonal from k_morton).  Every thread evaluates the same expression; thread 0 records it.
    const float max_cell = fmaxf(bp->cell[0], fmaxf(bp->cell[1], bp->cell[2]));
    // (mean leaf diagonal: k_morton's sum, kept in the zeroed counter block: counters[5] holds a float)
    const bool use_q = 6.0f * max_cell <= quant_frac * (*reinterpret_cast<const float *>(counters + 5) / (float)n);
    if (i == 0) bp->use_q = use_q ? 1 : 0;
    const RefitOut R{ bn, tn, qn, bp, n, leaf_max, use_q, keep_bn != 0 };
    const RefitQueue Q{ s_qf, s_qi, bfirst, blast };
    float4 mlo = make_float4(0.f, 0.f, 0.f, 0.f), mhi = mlo;      // box of the subtree I hold
    int l = 0, r = 0, dl = -1, dr = -1;
    bool holding = false;                         // I carry a finished subtree nobody has merged yet
    s_qi[1][threadIdx.x] = -1;
    s_flag[threadIdx.x] = 0u;
    if (i < n) {
        uint64_t t = order[i];
        const uint64_t k = keys[i];
        if (i > 0) dl = delta_adjacent(keys[i - 1], k, i - 1);
        if (i + 1 < n) dr = delta_adjacent(k, keys[i + 1], i);
        float blo[3], bhi[3];
        if (refs) {                 // a reference of a split triangle: its slab's box; `order` becomes sorted position -> triangle
            const RefBox rb = refs[t];
            t = rb.tri;
            order[i] = rb.tri;
#pragma unroll
            for (int a = 0; a < 3; ++a) { blo[a] = rb.lo[a]; bhi[a] = rb.hi[a]; }
        }
        uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
        float p0[3], p1[3], p2[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { p0[a] = verts[3ull * i0 + a]; p1[a] = verts[3ull * i1 + a]; p2[a] = verts[3ull * i2 + a]; }
        if (!refs) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { blo[a] = fminf(p0[a], fminf(p1[a], p2[a])); bhi[a] = fmaxf(p0[a], fmaxf(p1[a], p2[a])); }
        }
        const float pad = bp->pad;
        mlo.x = __fsub_rn(blo[0], pad); mlo.y = __fsub_rn(blo[1], pad); mlo.z = __fsub_rn(blo[2], pad);
        mhi.x = __fadd_rn(bhi[0], pad); mhi.y = __fadd_rn(bhi[1], pad); mhi.z = __fadd_rn(bhi[2], pad);
        if (keep_bn) {
            mlo.w = __int_as_float(-1); mhi.w = __int_as_float(-1);
            float4 *leaf = reinterpret_cast<float4 *>(&bn[n - 1 + i]);
            leaf[0] = mlo; leaf[1] = mhi;
        }
        uint32_t g = ngeoms > 1 ? geom_of(goff, ngeoms, t) : 0u;
        uint32_t prim = (uint32_t)(t - goff[g]);
        TriRec rec;
        rec.p0 = make_float4(p0[0], p0[1], p0[2], __uint_as_float(prim));
        rec.p1 = make_float4(__fsub_rn(p0[0], p1[0]), __fsub_rn(p0[1], p1[1]), __fsub_rn(p0[2], p1[2]), __uint_as_float(g));
        rec.p2 = make_float4(__fsub_rn(p2[0], p0[0]), __fsub_rn(p2[1], p0[1]), __fsub_rn(p2[2], p0[2]), 0.0f);
        tris[i] = rec;
        l = r = (int)i;
        holding = n > 1;
    }
    uint32_t h = 0;                               // height of the subtree I hold (leaf = 0)

    // ---- block-wide minima of the deltas: s_pmin[k] = min delta(j, j+1) over j = b0-1 .. b0-1+k,
    //      s_smin[k] = min over j = b0+k .. last leaf of the block (threads past the end contribute nothing)
    {
        int pm = i < n ? dl : RF_BIG, sm = i < n ? dr : RF_BIG;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(FULL, pm, o), b = __shfl_down_sync(FULL, sm, o);
            if (lane >= o) pm = min(pm, a);
            if (lane + o < 32) sm = min(sm, b);
        }
        if (lane == 31) s_wp[warp] = pm;
        if (lane == 0) s_ws[warp] = sm;
        __syncthreads();
        for (int k = 0; k < warp; ++k) pm = min(pm, s_wp[k]);
        for (int k = warp + 1; k < RF_WARPS; ++k) sm = min(sm, s_ws[k]);
        s_pmin[threadIdx.x] = pm; s_smin[threadIdx.x] = sm;
        __syncthreads();
    }

    // ---- warp phase: I am a left child (dr > dl) and the lane of leaf r+1 holds my finished right sibling
    for (;;) {
        const bool go_right = dr > dl;
        const int partner = (holding && go_right && r + 1 - w0 < 32) ? r + 1 - w0 : lane;
        const bool p_hold = __shfl_sync(FULL, (int)holding, partner) != 0;
        const bool p_right = __shfl_sync(FULL, (int)go_right, partner) != 0;
        const int p_r = __shfl_sync(FULL, r, partner);
        const int p_dr = __shfl_sync(FULL, dr, partner);
        const uint32_t p_h = __shfl_sync(FULL, h, partner);
        float4 slo, shi;
        slo.x = __shfl_sync(FULL, mlo.x, partner); slo.y = __shfl_sync(FULL, mlo.y, partner); slo.z = __shfl_sync(FULL, mlo.z, partner);
        shi.x = __shfl_sync(FULL, mhi.x, partner); shi.y = __shfl_sync(FULL, mhi.y, partner); shi.z = __shfl_sync(FULL, mhi.z, partner);
        // the partner lane's subtree starts at its own leaf (r + 1); it is my sibling iff it wants to go left
        const bool take = partner != lane && p_hold && !p_right;
        // a taken lane cannot see its taker (it does not know where its left sibling starts): the takers tell it
        const unsigned consumed = __reduce_or_sync(FULL, take ? (1u << partner) : 0u);
        if (take) {
            const int g = r;
            r = p_r; dr = p_dr;
            h = max(h, p_h) + 1u;
            const bool root = l == 0 && r == (int)(n - 1);
            const int32_t id = root ? 0 : (dr > dl ? r : l);
            if (root) { counters[2] = h; holding = false; }      // tree height (edges from the root to the deepest leaf)
            refit_merge(R, Q, id, l, g, r, mlo, mhi, slo, shi, mlo, mhi, n_nodes, n_leafrefs);
        }
        if ((consumed >> lane) & 1u) holding = false;
        if (consumed == 0u) break;
    }

    // ---- block phase: the parent's whole range lies inside this block -> join through shared memory
    while (holding) {
        const bool go_right = dr > dl;
        const int s = go_right ? r : l - 1;       // the parent's split lies between leaves s and s + 1
        // parent = [l .. first delta below dr right of r] or [first delta below dl left of l .. r]
        const bool local = l >= bfirst && r <= blast &&
                           (go_right ? (r < blast && s_smin[r + 1 - bfirst] < dr) : (l > bfirst && s_pmin[l - 1 - bfirst] < dl));
        if (!local) break;
        const int k = s - bfirst;
        float *mine_box = s_box[go_right ? 0 : 1][k];
        mine_box[0] = mlo.x; mine_box[1] = mlo.y; mine_box[2] = mlo.z; mine_box[3] = mhi.x; mine_box[4] = mhi.y; mine_box[5] = mhi.z;
        // far end (relative, +1 so the word is never 0), height, far delta + 1
        const uint32_t mine = (uint32_t)((go_right ? l : r) - bfirst + 1) | (h << 11) | ((uint32_t)((go_right ? dl : dr) + 1) << 19);
        __threadfence_block();
        const uint32_t other = atomicExch(&s_flag[k], mine);
        if (other == 0u) { holding = false; break; }      // first to arrive: the sibling's thread takes over
        __threadfence_block();
        const volatile float *ob = s_box[go_right ? 1 : 0][k];
        float4 slo, shi;
        slo.x = ob[0]; slo.y = ob[1]; slo.z = ob[2]; shi.x = ob[3]; shi.y = ob[4]; shi.z = ob[5];
        const int o_end = (int)(other & 0x7FFu) - 1 + bfirst;
        const uint32_t o_h = (other >> 11) & 0xFFu;
        const int o_d = (int)((other >> 19) & 0xFFu) - 1;
        h = max(h, o_h) + 1u;
        if (go_right) { r = o_end; dr = o_d; } else { l = o_end; dl = o_d; }
        const bool root = l == 0 && r == (int)(n - 1);
        const int32_t id = root ? 0 : (dr > dl ? r : l);
        if (root) { counters[2] = h; holding = false; }
        if (go_right) refit_merge(R, Q, id, l, s, r, mlo, mhi, slo, shi, mlo, mhi, n_nodes, n_leafrefs);
        else          refit_merge(R, Q, id, l, s, r, slo, shi, mlo, mhi, mlo, mhi, n_nodes, n_leafrefs);
    }

    // ---- hand the subtrees that are still open to the climb kernel (their parents span several blocks);
    //      one atomic per block on the list length (per warp: 62k atomics on one address at 2M triangles)
    {
        __shared__ unsigned s_hold[RF_WARPS], s_base;
        const unsigned hm = __ballot_sync(FULL, holding);
        if (lane == 0) s_hold[warp] = (unsigned)__popc(hm);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
            for (int k = 0; k < RF_WARPS; ++k) { const unsigned c = s_hold[k]; s_hold[k] = tot; tot += c; }
            s_base = tot ? atomicAdd(work_count, tot) : 0u;
        }
        __syncthreads();
        if (holding) {
            const unsigned slot = s_base + s_hold[warp] + __popc(hm & ((1u << lane) - 1u));
            if (slot < work_cap) {
                ClimbItem it;
                it.a = make_int4(l, r, dl, dr);
                it.b = make_float4(mlo.x, mlo.y, mlo.z, __uint_as_float(h));
                it.c = make_float4(mhi.x, mhi.y, mhi.z, 0.0f);
                work[slot] = it;
            } else {
                global_climb(R, Q, flags, counters, l, r, dl, dr, h, mlo, mhi, n_nodes, n_leafrefs);    // list full: climb here
            }
        }
    }

    // ---- output: every parked merge of this block, one node per thread
    __syncthreads();
    {
        const int k = threadIdx.x, g = s_qi[1][k];
        if (g >= 0) {
            const float4 llo = make_float4(s_qf[0][k], s_qf[1][k], s_qf[2][k], 0.0f), lhi = make_float4(s_qf[3][k], s_qf[4][k], s_qf[5][k], 0.0f);
            const float4 rlo = make_float4(s_qf[6][k], s_qf[7][k], s_qf[8][k], 0.0f), rhi = make_float4(s_qf[9][k], s_qf[10][k], s_qf[11][k], 0.0f);
            emit_node(R, (int32_t)i, s_qi[0][k], g, s_qi[2][k], llo, lhi, rlo, rhi, n_nodes, n_leafrefs);
        }
    }
    // statistics: one pair of atomics per block (per warp they were 125k atomics on two addresses at 2M triangles)
    for (int o = 16; o > 0; o >>= 1) {
        n_nodes += __shfl_xor_sync(FULL, n_nodes, o);
        n_leafrefs += __shfl_xor_sync(FULL, n_leafrefs, o);
    }
    __syncthreads();                        // s_wp / s_ws are free again
    if (lane == 0) { s_wp[warp] = (int)n_nodes; s_ws[warp] = (int)n_leafrefs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tn_ = 0, tl_ = 0;
        for (int k = 0; k < RF_WARPS; ++k) { tn_ += (unsigned)s_wp[k]; tl_ += (unsigned)s_ws[k]; }
        if (tn_) {
            atomicAdd(&counters[0], (unsigned long long)tn_);
            atomicAdd(&counters[1], (unsigned long long)tl_);
        }
    }
}

// The climb kernel: one open subtree per thread (grid-stride), small blocks -- a long chain towards the root holds
// one warp, not the 256 leaves' worth of registers and shared memory it would hold inside the kernel above.
constexpr int CL_BLOCK = 64;

__global__ void __launch_bounds__(CL_BLOCK)
k_hierarchy_climb(int64_t n, BuildParams *bp, BNode *bn, unsigned long long *flags, TNode *__restrict__ tn,
                  QNode *__restrict__ qn, unsigned long long *counters, int leaf_max, int keep_bn,
                  const ClimbItem *__restrict__ work, const unsigned *__restrict__ work_count, unsigned work_cap)
{
    const RefitOut R{ bn, tn, qn, bp, n, leaf_max, bp->use_q != 0, keep_bn != 0 };
    const RefitQueue Q{ nullptr, nullptr, 0, -1 };          // no parking here: nodes are written as they are made
    const unsigned count = min(*work_count, work_cap);
    unsigned n_nodes = 0, n_leafrefs = 0;
    for (unsigned w = blockIdx.x * CL_BLOCK + threadIdx.x; w < count; w += gridDim.x * CL_BLOCK) {
        const ClimbItem it = work[w];
        global_climb(R, Q, flags, counters, it.a.x, it.a.y, it.a.z, it.a.w, __float_as_uint(it.b.w),
                     it.b, it.c, n_nodes, n_leafrefs);
    }
    for (int o = 16; o > 0; o >>= 1) {
        n_nodes += __shfl_xor_sync(0xFFFFFFFFu, n_nodes, o);
        n_leafrefs += __shfl_xor_sync(0xFFFFFFFFu, n_leafrefs, o);
    }
    if ((threadIdx.x & 31) == 0 && n_nodes) {
        atomicAdd(&counters[0], (unsigned long long)n_nodes);
        atomicAdd(&counters[1], (unsigned long long)n_leafrefs);
    }
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

size_t lbvh_sort_scratch_bytes(uint64_t n)
{
    uint64_t ntiles = (n + RS_TILE - 1) / RS_TILE, os_tiles = (n + OS_TILE - 1) / OS_TILE;
    // classic: tile_hist[256][ntiles] + digit_tot[256]; onesweep: status[8][os_tiles][256] + ghist[8][256] + counters[8]
    uint64_t classic = ntiles * 256 + 256, onesweep = 8 * os_tiles * 256 + 8 * 256 + 8;
    return (size_t)(std::max(classic, onesweep) + 64) * sizeof(uint32_t);
}

// capacity of the list of subtrees handed from k_hierarchy_refit_emit to k_hierarchy_climb (48-byte items); typically
// ~n/40 are needed, a full list only means the rest climbs inside the first kernel
size_t lbvh_climb_items(uint64_t n) { return (size_t)(n / 4 + 1024); }
size_t lbvh_climb_bytes(uint64_t n) { return lbvh_climb_items(n) * sizeof(ClimbItem); }


// counters (64 B) | flags [n - 1] | sort scratch: one allocation, cleared by one memset at the start of the build
size_t lbvh_zero_block_bytes(uint64_t n) { return 64 + (size_t)n * sizeof(unsigned long long) + lbvh_sort_scratch_bytes(n); }

size_t lbvh_sort_clear_bytes(uint64_t n)
{
    const uint64_t os_tiles = (n + OS_TILE - 1) / OS_TILE;
    return (size_t)(8ull * os_tiles * 256 + 8 * 256 + 8) * sizeof(uint32_t);
}

// keys / vals sorted by key (stable).  `overflow` non-null: only the top bits are sorted (sort_passes(n) passes) and
// k_sort_fixup orders the short runs that agree in them -- fewer trips of 12 B per key through HBM for the same final
// order.  hist_done: the caller zeroed the scratch and k_morton already filled the digit histograms.
int lbvh_radix_sort(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp,
                    uint64_t n, uint32_t *scratch, cudaStream_t st, unsigned long long *overflow, int sort_variant, bool hist_done,
                    int *result_in_tmp)
{
    *result_in_tmp = 0;
    if (n == 0) return 0;
    uint32_t ntiles = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint32_t *tile_hist = scratch, *digit_tot = scratch + (uint64_t)ntiles * 256;
    uint64_t *kin = keys, *kout = keys_tmp;
    uint32_t *vin = vals, *vout = vals_tmp;
    if (sort_variant == 1) {
        if (!hist_done) { qsmrt_set_error("onesweep sort needs the histograms of k_morton"); return 1; }
        const uint32_t os_tiles = (uint32_t)((n + OS_TILE - 1) / OS_TILE);
        // OS_CTAS x 45 KB only fit with the largest shared-memory split (a per-device function attribute; setting it is idempotent)
        CUDA_TRY(cudaFuncSetAttribute(k_os_pass, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        uint32_t *status = scratch, *ghist = scratch + 8ull * os_tiles * 256, *counters = ghist + 8 * 256;
        const int shift0 = overflow ? sort_shift0(n) : 0, npass = overflow ? sort_passes(n) : 8;
        for (int pass = 0; pass < npass; ++pass) {
            k_os_pass<<<os_tiles, OS_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift0 + pass * 8, ghist + pass * 256,
                                                       status + (uint64_t)pass * os_tiles * 256, counters + pass);
            uint64_t *tk = kin; kin = kout; kout = tk;
            uint32_t *tv = vin; vin = vout; vout = tv;
        }
        if (overflow) {         // the fix-up copies while it orders: the result is in whichever buffer pair it wrote
            k_sort_fixup<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(kin, vin, kout, vout, (int64_t)n, overflow, shift0);
            *result_in_tmp = kout == keys_tmp;
        } else *result_in_tmp = kin == keys_tmp;
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

// Scene bounds -> Morton scale, box padding and the 16-bit grid of the quantised nodes.  Host code: the bounds come
// from the geometries' registration-time statistics.  Plain IEEE single arithmetic, the same operations in the same
// order as oracle/qsmrt_oracle.c::orc_commit (the builder parity tests compare keys and padding bit for bit).
void lbvh_finalize_params(const float lo[3], const float hi[3], BuildParams *bp)
{
    memset(bp, 0, sizeof(*bp));
    volatile float m = 0.0f;
    for (int a = 0; a < 3; ++a) {
        volatile float ext = hi[a] - lo[a];
        bp->slo[a] = lo[a]; bp->shi[a] = hi[a];
        bp->scale[a] = ext > 0.0f ? 2097152.0f / ext : 0.0f;
        m = fmaxf(m, fabsf(lo[a])); m = fmaxf(m, fabsf(hi[a])); m = fmaxf(m, ext);
    }
    volatile float pad = m * 7.62939453125e-06f;    // 2^-17, see DESIGN.md "box padding"
    bp->pad = pad > 0.0f ? pad : 1e-30f;
    // quantisation grid: padded bounds plus a 32-cell margin so the +-3-cell widening never clamps
    for (int a = 0; a < 3; ++a) {
        volatile float ext = bp->shi[a] - bp->slo[a];
        volatile float w = ext * 4.8828125e-04f;    // 2^-11 (volatile: no fused multiply-add on the host either)
        volatile float e = bp->pad + w;
        bp->glo[a] = bp->slo[a] - e;
        volatile float two_e = 2.0f * e;
        bp->cell[a] = (ext + two_e) / 65535.0f;
        bp->inv_cell[a] = 1.0f / bp->cell[a];
    }
    bp->leaf_diag_sum = 0.0f;
}

// out7_dev: 7 words of device scratch; out_host[7] receives lo.xyz, hi.xyz (as floats) and the largest index
int lbvh_geometry_stats(const float *verts, uint64_t V, const uint32_t *idx, uint64_t T, uint32_t *out7_dev, float lo[3], float hi[3],
                        uint32_t *max_index, cudaStream_t st)
{
    const uint32_t init[7] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0u };
    uint32_t h[7];
    CUDA_TRY(cudaMemcpyAsync(out7_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (T) k_geometry_stats<<<(unsigned)std::min<uint64_t>((T + 255) / 256, 148 * 8), 256, 0, st>>>(verts, V, idx, T, out7_dev);
    CUDA_TRY(cudaMemcpyAsync(h, out7_dev, sizeof(h), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (int a = 0; a < 3; ++a) {
        const uint32_t ul = h[a], uh = h[3 + a];
        const uint32_t bl = (ul & 0x80000000u) ? (ul & 0x7FFFFFFFu) : ~ul, bh = (uh & 0x80000000u) ? (uh & 0x7FFFFFFFu) : ~uh;
        memcpy(&lo[a], &bl, 4); memcpy(&hi[a], &bh, 4);
    }
    *max_index = h[6];
    return 0;
}

// Sliver splitting, host side: how many references the triangles want in total (*total_dev, zeroed by the caller;
// counts may be null), and their boxes once the counts are scanned into ref_off[T + 1].
size_t lbvh_ref_bytes(uint64_t nrefs) { return (size_t)nrefs * sizeof(RefBox); }

int lbvh_split_count(const float *verts, const uint32_t *idx, uint64_t T, int split_max, float split_aspect, int32_t *counts,
                     unsigned long long *total_dev, cudaStream_t st)
{
    if (T == 0) return 0;
    k_split_count<<<(unsigned)std::min<uint64_t>((T + 255) / 256, 148 * 8), 256, 0, st>>>(verts, idx, T, split_max, split_aspect, counts, total_dev);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbvh_split_emit(const float *verts, const uint32_t *idx, uint64_t T, int split_max, float split_aspect, const int64_t *ref_off,
                    void *refs, cudaStream_t st)
{
    if (T == 0) return 0;
    k_split_emit<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(verts, idx, T, split_max, split_aspect, ref_off, static_cast<RefBox *>(refs));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbvh_build(const LbvhBuildArgs &A, cudaStream_t st)
{
    const uint64_t n = A.refs ? A.nrefs : A.ntris;          // leaves of the tree: references when slivers were split
    const int B = 256;
    const unsigned gN = (unsigned)((n + B - 1) / B);
    // ONE clear for everything the build accumulates into: the counters, the global-phase flags of the hierarchy and
    // the sort's look-back status / histograms / tile counters sit in one block (lbvh_zero_block_bytes); the finished
    // parameters travel to k_morton by value, which also stores the device copy the later kernels read
    const bool onesweep = A.sort_variant == 1;
    const bool top_only = !A.full_sort;
    CUDA_TRY(cudaMemsetAsync(A.counters, 0, 64 + (n > 1 ? n - 1 : 0) * sizeof(unsigned long long) + (onesweep ? lbvh_sort_clear_bytes(n) : 0), st));
    uint32_t *ghist = onesweep ? A.sort_scratch + 8ull * ((n + OS_TILE - 1) / OS_TILE) * 256 : nullptr;
    k_morton<<<min(gN, 148u * 16u), B, 0, st>>>(A.verts, A.idx, n, *A.params_host, A.params, reinterpret_cast<float *>(A.counters + 5),
                                                 A.keys, A.order, ghist,
                                                 top_only ? sort_shift0(n) : 0, top_only ? sort_passes(n) : 8,
                                                 static_cast<const RefBox *>(A.refs));
    CUDA_TRY(cudaGetLastError());
    if (A.ev_sort0) CUDA_TRY(cudaEventRecord(A.ev_sort0, st));
    int in_tmp = 0;
    if (lbvh_radix_sort(A.keys, A.keys_tmp, A.order, A.order_tmp, n, A.sort_scratch, st, top_only ? A.counters + 4 : nullptr,
                        A.sort_variant, onesweep, &in_tmp)) return 1;
    if (A.result_in_tmp) *A.result_in_tmp = in_tmp;
    const uint64_t *keys = in_tmp ? A.keys_tmp : A.keys;        // the caller keeps whichever pair holds the sorted arrays
    uint32_t *order = in_tmp ? A.order_tmp : A.order;
    if (A.ev_sort1) CUDA_TRY(cudaEventRecord(A.ev_sort1, st));
    const int keep = (A.keep_bnodes || n == 1) ? 1 : 0;
    const unsigned work_cap = A.climb_capacity > 0 ? std::min<unsigned>((unsigned)A.climb_capacity, (unsigned)lbvh_climb_items(n))
                                                       : (unsigned)lbvh_climb_items(n);
    unsigned *work_count = reinterpret_cast<unsigned *>(A.counters + 3);
    ClimbItem *work = reinterpret_cast<ClimbItem *>(A.climb_work);
    k_hierarchy_refit_emit<<<(unsigned)((n + RF_BLOCK - 1) / RF_BLOCK), RF_BLOCK, 0, st>>>(
        A.verts, A.idx, (int64_t)n, keys, order, static_cast<const RefBox *>(A.refs), A.geom_offsets, A.ngeoms, A.params, A.bnodes, A.tris, A.flags,
        A.tnodes, A.qnodes, A.counters, A.leaf_max, keep, work, work_count, work_cap, A.quant_frac);
    k_hierarchy_climb<<<std::min((work_cap + CL_BLOCK - 1) / CL_BLOCK, 148u * 32u), CL_BLOCK, 0, st>>>(
        (int64_t)n, A.params, A.bnodes, A.flags, A.tnodes, A.qnodes, A.counters, A.leaf_max, keep, work, work_count, work_cap);
    if (n == 1) k_emit_single<<<1, 1, 0, st>>>(A.bnodes, A.tnodes, A.qnodes, A.params, A.counters);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
