// traverse.cu -- BVH traversal kernels for sm_100a: closest hit (cast_rays),
// all hits (count_intersections / list_intersections) and any hit
// (test_occlusions), plus the ray generators.
//
// Replaces Embree's rtcIntersect1/rtcOccluded1 loops inside Open3D's
// RaycastingScene::CastRays / CountIntersections / ListIntersections /
// TestOcclusions (reference call sites: pyQSM/viz/ray_casting.py:168,223,231,
// 279,319).  Triangle arithmetic: common.cuh::mt_test, the bit-for-bit twin
// of oracle/qsmrt_oracle.c::mt_test.
//
// One thread per ray.  Each thread keeps a short traversal stack in shared
// memory laid out [entry][thread] (bank-conflict free, no local-memory
// traffic on the hot path) and spills to a local array only below depth
// TR_SSTACK.  Nodes are fetched as four 16-byte loads, triangles as three.
#include <algorithm>
#include <cstring>
#include "common.cuh"
#include "traverse.h"

namespace {

constexpr int TR_BLOCK  = 128;
constexpr int TR_SSTACK = 16;     // shared-memory entries per thread
constexpr int TR_LSTACK = 80;     // local spill; 96 total > max LBVH depth (63 key bits + 26 index bits)

struct Stack {
    int *s;                 // &smem[0][threadIdx.x]
    int *loc;               // the kernel's local spill array (kept outside the struct so sp stays in a register)
    int  sp;
    __device__ __forceinline__ void push(int v) {
        if (sp < TR_SSTACK) s[sp * TR_BLOCK] = v; else loc[sp - TR_SSTACK] = v;
        ++sp;
    }
    __device__ __forceinline__ int pop() {
        --sp;
        return sp < TR_SSTACK ? s[sp * TR_BLOCK] : loc[sp - TR_SSTACK];
    }
};

// Conservative slab test (subtract first: the sign is exact; interval widened
// by ~4 ulp).  Returns entry distance; hit iff entry <= exit.
__device__ __forceinline__ bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz,
                                     const Ray &r, float tmax, float &tn)
{
    float x0 = (lox - r.O.x) * r.idx, x1 = (hix - r.O.x) * r.idx;
    float y0 = (loy - r.O.y) * r.idy, y1 = (hiy - r.O.y) * r.idy;
    float z0 = (loz - r.O.z) * r.idz, z1 = (hiz - r.O.z) * r.idz;
    float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tfar = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    tmin *= 0.9999995f;
    tfar *= 1.0000005f;
    tn = tmin;
    return tmin <= tfar;
}

__device__ __forceinline__ void load_node(const TNode *__restrict__ nodes, int i, float4 &a, float4 &b, float4 &c, int4 &d)
{
    const float4 *p = reinterpret_cast<const float4 *>(nodes + i);
    a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2);
    d = __ldg(reinterpret_cast<const int4 *>(p + 3));
}

__device__ __forceinline__ void load_tri(const TriRec *__restrict__ tris, uint32_t i, float4 &p0, float4 &p1, float4 &p2)
{
    const float4 *p = reinterpret_cast<const float4 *>(tris + i);
    p0 = __ldg(p); p1 = __ldg(p + 1); p2 = __ldg(p + 2);
}

// Generic stack traversal.  V::tmax() bounds the interval (shrinks for closest
// hit), V::leaf(first, count) tests triangles and returns true to stop.
template <bool ORDERED, class V>
__device__ __forceinline__ void traverse(const SceneView &sc, const Ray &r, Stack &st, V &vis)
{
    st.sp = 0;
    int cur = 0;                         // root
    for (;;) {
        if (cur >= 0) {
            float4 a, b, c; int4 d;
            load_node(sc.nodes, cur, a, b, c, d);
            float t0, t1;
            const float tm = vis.tmax();
            bool h0 = slab(a.x, a.y, a.z, a.w, c.x, c.y, r, tm, t0);
            bool h1 = slab(b.x, b.y, b.z, b.w, c.z, c.w, r, tm, t1);
            if (h0 & h1) {
                bool swap = ORDERED && (t1 < t0);
                st.push(swap ? d.x : d.y);
                cur = swap ? d.y : d.x;
                continue;
            }
            if (h0) { cur = d.x; continue; }
            if (h1) { cur = d.y; continue; }
        } else {
            uint32_t ref = (uint32_t)~cur;
            if (vis.leaf(ref >> 2, (ref & 3u) + 1u)) return;
        }
        if (st.sp == 0) return;
        cur = st.pop();
    }
}

// ---- slab test of the persistent kernel --------------------------------------
// FMA form (t = plane * (1/d) - o * (1/d), FMNMX3): in position space its rounding
// moves a box face by <= ~1e-7 * max|coord|, far inside the 2^-17 padding every leaf
// box carries, so it stays conservative (checked against brute force in tests/).
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}

__device__ __forceinline__ bool slab_fma(float lox, float hix, float loy, float hiy, float loz, float hiz,
                                         const Ray &r, float tmax, float &tn)
{
    float x0 = fmaf(lox, r.idx, -r.oodx), x1 = fmaf(hix, r.idx, -r.oodx);
    float y0 = fmaf(loy, r.idy, -r.oody), y1 = fmaf(hiy, r.idy, -r.oody);
    float z0 = fmaf(loz, r.idz, -r.oodz), z1 = fmaf(hiz, r.idz, -r.oodz);
    float tmin = fmaxf(fmax3(fminf(x0, x1), fminf(y0, y1), fminf(z0, z1)) * 0.99999905f, 0.0f);      // see SLAB_NEAR / SLAB_FAR
    float tfar = fminf(fmin3(fmaxf(x0, x1), fmaxf(y0, y1), fmaxf(z0, z1)) * 1.00000095f, tmax);
    tn = tmin;
    return tmin <= tfar;
}

constexpr int TR_SENTINEL = 0x7FFFFFFF;

// 2-D tile mapping for image / grid shaped ray batches [rows][row_len]: a warp
// takes an 8 x 4 tile instead of 32 consecutive rays of one row, which keeps
// its rays in the same subtree longer.  Returns false for padding lanes.
__device__ __forceinline__ bool ray_index(uint64_t N, uint32_t row_len, uint64_t &i)
{
    if (row_len == 0) {
        i = blockIdx.x * (uint64_t)TR_BLOCK + threadIdx.x;
        return i < N;
    }
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = blockIdx.x * (uint64_t)(TR_BLOCK / 32) + (threadIdx.x >> 5);
    const uint32_t tiles_x = (row_len + 7u) >> 3;
    uint64_t ty = warp / tiles_x;
    const uint32_t tx = (uint32_t)(warp - ty * tiles_x);
    const uint64_t tile_rows = (N / row_len + 3u) >> 2;
    if (ty >= tile_rows) return false;
    // blocks are scheduled in index order: the middle rows of the grid first, the (cheap) edge rows last -- see centre_out
    ty = (ty & 1u) ? (tile_rows >> 1) - 1u - (ty >> 1) : (tile_rows >> 1) + (ty >> 1);
    const uint32_t x = tx * 8u + (lane & 7u);
    const uint64_t y = ty * 4u + (lane >> 3);
    i = y * row_len + x;
    return x < row_len && i < N;
}

// ------------------------------------------------------------ closest hit
struct ClosestVis {
    const SceneView &sc; const Ray &r;
    float t; uint32_t geom, prim, tri;
    __device__ __forceinline__ float tmax() const { return t; }
    __device__ __forceinline__ bool leaf(uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k) {
            float4 p0, p1, p2;
            load_tri(sc.tris, first + k, p0, p1, p2);
            MtHit h;
            if (mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                float tt = __fdiv_rn(h.T, h.absDen);
                uint32_t pg = __float_as_uint(p1.w), pp = __float_as_uint(p0.w);
                // exact ties go to the lowest (geometry, primitive)
                bool better = (tt < t) | ((tt == t) & ((pg < geom) | ((pg == geom) & (pp < prim))));
                if (better) { t = tt; geom = pg; prim = pp; tri = first + k; }
            }
        }
        return false;
    }
};

__global__ void __launch_bounds__(TR_BLOCK)
k_cast_rays(SceneView sc, const float *__restrict__ rays, uint64_t N, uint32_t row_len,
            float *__restrict__ t_hit, uint32_t *__restrict__ geom, uint32_t *__restrict__ prim,
            float2 *__restrict__ uv, float *__restrict__ nrm)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    uint64_t i;
    if (!ray_index(N, row_len, i)) return;
    Ray r = load_ray(rays, i);
    int spill[TR_LSTACK];
    Stack st; st.s = sstack + threadIdx.x; st.loc = spill;
    ClosestVis vis{ sc, r, INFINITY, QSMRT_INVALID, QSMRT_INVALID, 0u };
    if (sc.ntris) traverse<true>(sc, r, st, vis);
    if (t_hit) t_hit[i] = vis.t;
    if (geom) geom[i] = vis.geom;
    if (prim) prim[i] = vis.prim;
    if (uv || nrm) {
        float u = 0.0f, v = 0.0f, nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (vis.prim != QSMRT_INVALID) {
            float4 p0, p1, p2;
            load_tri(sc.tris, vis.tri, p0, p1, p2);
            MtHit h;
            mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h);
            u = __fdiv_rn(h.U, h.absDen); v = __fdiv_rn(h.V, h.absDen);
            float inv = __fdiv_rn(1.0f, __fsqrt_rn(f3dot(h.Ng, h.Ng)));
            nx = __fmul_rn(h.Ng.x, inv); ny = __fmul_rn(h.Ng.y, inv); nz = __fmul_rn(h.Ng.z, inv);
        }
        if (uv) uv[i] = make_float2(u, v);
        if (nrm) { nrm[3 * i] = nx; nrm[3 * i + 1] = ny; nrm[3 * i + 2] = nz; }
    }
}

// ---- persistent kernel (trace_persistent.cuh): ray slots and outputs ----------
// Evidence for its shape (profiles/README.md): the per-thread loop below spends half
// its issue slots in leaf code with ~2.3 of 32 lanes active, and idle lanes of
// finished rays wait for the slowest ray of the warp.
// Rows of tiles are handed out from the middle of the grid outwards (mid, mid - 1, mid + 1, ...): ray grids are laid
// over the scene's bounding box, so the long rays -- through the crown of a tree, the bulk of a canopy -- sit in
// the middle rows and the cheap ones (misses near the box edge) at the top and bottom.  Row-major order left the
// heaviest rays to start late and finish alone: on the C1 tree (1M rays) the average SM was busy 56 % of the
// launch (ncu sm__cycles_active avg / max, profiles/r02_c1_tail.txt).  tile_rows = number of tile rows.
__device__ __forceinline__ uint64_t centre_out(uint64_t k, uint64_t tile_rows)
{
    const uint64_t mid = tile_rows >> 1;
    return (k & 1u) ? mid - 1u - (k >> 1) : mid + (k >> 1);
}

__device__ __forceinline__ bool ray_index_of_slot(uint64_t slot, uint64_t N, uint32_t row_len, uint64_t &i, uint32_t &x, uint64_t &y,
                                                  bool row_major = false)
{
    if (row_len == 0) { i = slot; return slot < N; }
    const uint32_t lane = (uint32_t)(slot & 31u);
    const uint64_t tile = slot >> 5;
    const uint32_t tiles_x = (row_len + 7u) >> 3;
    // 32-bit division whenever the tile number fits (batches below 2^37 rays): the 64-bit one is ~100 instructions
    uint64_t ty = (tile >> 32) == 0 ? (uint64_t)((uint32_t)tile / tiles_x) : tile / tiles_x;
    const uint32_t tx = (uint32_t)(tile - ty * tiles_x);
    const uint64_t rows = (N >> 32) == 0 ? (uint64_t)((uint32_t)N / row_len) : N / row_len;
    if (!row_major) ty = centre_out(ty, (rows + 3u) >> 2);
    x = tx * 8u + (lane & 7u);
    y = ty * 4u + (lane >> 3);
    i = y * row_len + x;
    return x < row_len && i < N;
}

struct CastOut {
    float *t_hit; uint32_t *geom, *prim; float2 *uv; float *nrm;
};

#include "trace_persistent.cuh"

// ---------------------------------------------------------------- any hit
struct AnyVis {
    const SceneView &sc; const Ray &r; float tnear, tfar; bool hit;
    __device__ __forceinline__ float tmax() const { return tfar; }
    __device__ __forceinline__ bool leaf(uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k) {
            float4 p0, p1, p2;
            load_tri(sc.tris, first + k, p0, p1, p2);
            MtHit h;
            if (mt_test(p0, p1, p2, r.O, r.D, tnear, tfar, h)) { hit = true; return true; }
        }
        return false;
    }
};

__global__ void __launch_bounds__(TR_BLOCK)
k_test_occlusions(SceneView sc, const float *__restrict__ rays, uint64_t N, float tnear, float tfar,
                  uint8_t *__restrict__ out)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    uint64_t i = blockIdx.x * (uint64_t)TR_BLOCK + threadIdx.x;
    if (i >= N) return;
    Ray r = load_ray(rays, i);
    int spill[TR_LSTACK];
    Stack st; st.s = sstack + threadIdx.x; st.loc = spill;
    AnyVis vis{ sc, r, tnear, tfar, false };
    if (sc.ntris) traverse<false>(sc, r, st, vis);
    out[i] = vis.hit ? 1 : 0;
}

// --------------------------------------------------------------- all hits
// count_intersections: Open3D's CountIntersectionsFunc counts a hit unless it
// repeats the previous callback's geometry with an equal t; restated
// order-independently as "distinct (geometry, t) pairs".  Fast path: a small
// per-thread set; a ray with more distinct hits than the set holds falls
// through to next_distinct(), which enumerates the pairs in increasing order
// with one traversal each (exact, no storage, no host round trip).
constexpr int CNT_CAP = 24;

struct CountVis {
    const SceneView &sc; const Ray &r;
    float ts[CNT_CAP]; uint32_t gs[CNT_CAP]; int n; bool overflow;
    __device__ __forceinline__ float tmax() const { return INFINITY; }
    __device__ __forceinline__ bool leaf(uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k) {
            float4 p0, p1, p2;
            load_tri(sc.tris, first + k, p0, p1, p2);
            MtHit h;
            if (mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                float tt = __fdiv_rn(h.T, h.absDen);
                uint32_t pg = __float_as_uint(p1.w);
                bool dup = false;
                for (int q = 0; q < n; ++q) dup |= (ts[q] == tt) & (gs[q] == pg);
                if (!dup) {
                    if (n < CNT_CAP) { ts[n] = tt; gs[n] = pg; ++n; }
                    else { overflow = true; return true; }
                }
            }
        }
        return false;
    }
};

// smallest (t, geom) pair strictly greater than (pt, pg) among accepted hits
struct NextVis {
    const SceneView &sc; const Ray &r;
    float pt; uint32_t pg; bool have_prev;
    float bt; uint32_t bg; bool found;
    __device__ __forceinline__ float tmax() const { return bt; }
    __device__ __forceinline__ bool leaf(uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k) {
            float4 p0, p1, p2;
            load_tri(sc.tris, first + k, p0, p1, p2);
            MtHit h;
            if (mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                float tt = __fdiv_rn(h.T, h.absDen);
                uint32_t g = __float_as_uint(p1.w);
                bool after = !have_prev || (tt > pt) || (tt == pt && g > pg);
                bool better = !found || (tt < bt) || (tt == bt && g < bg);
                if (after && better) { bt = tt; bg = g; found = true; }
            }
        }
        return false;
    }
};

__global__ void __launch_bounds__(TR_BLOCK)
k_count_intersections(SceneView sc, const float *__restrict__ rays, uint64_t N, int32_t *__restrict__ out)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    uint64_t i = blockIdx.x * (uint64_t)TR_BLOCK + threadIdx.x;
    if (i >= N) return;
    Ray r = load_ray(rays, i);
    int spill[TR_LSTACK];
    Stack st; st.s = sstack + threadIdx.x; st.loc = spill;
    int result = 0;
    if (sc.ntris) {
        CountVis vis{ sc, r };
        vis.n = 0; vis.overflow = false;
        traverse<false>(sc, r, st, vis);
        result = vis.n;
        if (vis.overflow) {
            result = 0;
            NextVis nv{ sc, r, 0.0f, 0u, false, INFINITY, 0u, false };
            for (;;) {
                nv.bt = INFINITY; nv.bg = 0u; nv.found = false;
                traverse<true>(sc, r, st, nv);
                if (!nv.found) break;
                ++result;
                nv.pt = nv.bt; nv.pg = nv.bg; nv.have_prev = true;
            }
        }
    }
    out[i] = result;
}

// exact recount of the rays the persistent count kernel marked -1 (more distinct hits than its set holds)
__global__ void __launch_bounds__(TR_BLOCK)
k_count_fix(SceneView sc, const float *__restrict__ rays, uint64_t N, int32_t *__restrict__ out)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    int spill[TR_LSTACK];
    Stack st; st.s = sstack + threadIdx.x; st.loc = spill;
    for (uint64_t i = blockIdx.x * (uint64_t)TR_BLOCK + threadIdx.x; i < N; i += (uint64_t)gridDim.x * TR_BLOCK) {
        if (out[i] >= 0) continue;
        Ray r = load_ray(rays, i);
        int result = 0;
        NextVis nv{ sc, r, 0.0f, 0u, false, INFINITY, 0u, false };
        for (;;) {
            nv.bt = INFINITY; nv.bg = 0u; nv.found = false;
            traverse<true>(sc, r, st, nv);
            if (!nv.found) break;
            ++result;
            nv.pt = nv.bt; nv.pg = nv.bg; nv.have_prev = true;
        }
        out[i] = result;
    }
}

// ---- list_intersections, throughput path -------------------------------------------------------------------
// (1) ONE all-hits traversal (persistent MODE 6 + exact fix-up of the rays with more than CNT_SET hits): the counts,
//     and the distinct hits of every ray with 1 .. CNT_SET of them in a stash (records in retirement order);
// (2) exclusive scan of the counts = ray_splits;
// (3) k_list_finish: one thread per ray moves its records from the stash to the caller's arrays at ray_splits[i],
//     ordered by (t, geometry, primitive), and fills ray_ids; rays with more than CNT_SET hits enumerate them in
//     increasing order instead, one traversal per hit.
struct NextFullVis {      // smallest (t, geom) strictly after (pt, pg); among equal (t, geom) the lowest primitive
    const SceneView &sc; const Ray &r;
    float pt; uint32_t pg; bool have_prev;
    float bt; uint32_t bg, bp; float bu, bv; bool found;
    __device__ __forceinline__ float tmax() const { return bt; }
    __device__ __forceinline__ bool leaf(uint32_t first, uint32_t count) {
        for (uint32_t k = 0; k < count; ++k) {
            float4 p0, p1, p2;
            load_tri(sc.tris, first + k, p0, p1, p2);
            MtHit h;
            if (mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                const float tt = __fdiv_rn(h.T, h.absDen);
                const uint32_t g = __float_as_uint(p1.w), pp = __float_as_uint(p0.w);
                const bool after = !have_prev || (tt > pt) || (tt == pt && g > pg);
                const bool better = !found || (tt < bt) || (tt == bt && (g < bg || (g == bg && pp < bp)));
                if (after && better) {
                    bt = tt; bg = g; bp = pp; found = true;
                    bu = __fdiv_rn(h.U, h.absDen); bv = __fdiv_rn(h.V, h.absDen);
                }
            }
        }
        return false;
    }
};

__global__ void __launch_bounds__(TR_BLOCK)
k_list_finish(SceneView sc, const float *__restrict__ rays, uint64_t N, const int64_t *__restrict__ splits, int max_fast,
              const long long *__restrict__ st_base, const float *__restrict__ st_t, const uint32_t *__restrict__ st_geom,
              const uint32_t *__restrict__ st_prim, const float2 *__restrict__ st_uv,
              int64_t *__restrict__ ray_ids, float *__restrict__ t_hit, uint32_t *__restrict__ geom, uint32_t *__restrict__ prim,
              float2 *__restrict__ uv)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    const uint64_t i = blockIdx.x * (uint64_t)TR_BLOCK + threadIdx.x;
    if (i >= N) return;
    const int64_t o = splits[i];
    const int n = (int)(splits[i + 1] - o);
    if (n == 0) return;
    if (ray_ids) for (int k = 0; k < n; ++k) ray_ids[o + k] = (int64_t)i;
    if (n <= max_fast) {
        // insertion sort of the records the persistent kernel stashed (a handful of hits; keys (t, geom, prim))
        const long long sb = st_base[i];
        for (int a = 0; a < n; ++a) {
            const float xt = st_t[sb + a]; const uint32_t xg = st_geom[sb + a], xp = st_prim[sb + a]; const float2 xu = st_uv[sb + a];
            int b = a;
            for (; b > 0; --b) {
                const float yt = t_hit[o + b - 1]; const uint32_t yg = geom[o + b - 1], yp = prim[o + b - 1];
                const bool less = xt < yt || (xt == yt && (xg < yg || (xg == yg && xp < yp)));
                if (!less) break;
                t_hit[o + b] = yt; geom[o + b] = yg; prim[o + b] = yp; uv[o + b] = uv[o + b - 1];
            }
            t_hit[o + b] = xt; geom[o + b] = xg; prim[o + b] = xp; uv[o + b] = xu;
        }
        return;
    }
    const Ray r = load_ray(rays, i);
    int spill[TR_LSTACK];
    Stack st; st.s = sstack + threadIdx.x; st.loc = spill;
    NextFullVis nv{ sc, r, 0.0f, 0u, false, INFINITY, 0u, 0u, 0.0f, 0.0f, false };
    for (int k = 0; k < n; ++k) {
        nv.bt = INFINITY; nv.bg = 0u; nv.bp = 0u; nv.found = false;
        traverse<true>(sc, r, st, nv);
        if (!nv.found) break;                       // cannot happen: n came from the exact count
        t_hit[o + k] = nv.bt; geom[o + k] = nv.bg; prim[o + k] = nv.bp; uv[o + k] = make_float2(nv.bu, nv.bv);
        nv.pt = nv.bt; nv.pg = nv.bg; nv.have_prev = true;
    }
}

// ------------------------------------------- exclusive scan int32 -> int64
constexpr int SC_THREADS = 256, SC_ITEMS = 16, SC_TILE = SC_THREADS * SC_ITEMS;

__device__ __forceinline__ long long block_excl_scan(long long v, long long *total)
{
    __shared__ long long ws[SC_THREADS / 32];
    const int l = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long x = v;
    for (int o = 1; o < 32; o <<= 1) { long long y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (l >= o) x += y; }
    if (l == 31) ws[w] = x;
    __syncthreads();
    long long off = 0, tot = 0;
    for (int k = 0; k < SC_THREADS / 32; ++k) { if (k < w) off += ws[k]; tot += ws[k]; }
    __syncthreads();
    *total = tot;
    return off + x - v;
}

__global__ void __launch_bounds__(SC_THREADS)
k_scan_tile_sums(const int32_t *__restrict__ in, uint64_t n, long long *__restrict__ tile_sum)
{
    uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
    long long s = 0;
    for (int k = 0; k < SC_ITEMS; ++k) if (base + k < n) s += in[base + k];
    long long tot;
    block_excl_scan(s, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SC_THREADS)
k_scan_tile_offsets(long long *tile_sum, uint32_t ntiles)   // single block, in place -> exclusive
{
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < ntiles; b += SC_THREADS) {
        uint32_t i = b + threadIdx.x;
        long long v = i < ntiles ? tile_sum[i] : 0, tot;
        long long e = block_excl_scan(v, &tot);
        if (i < ntiles) tile_sum[i] = carry + e;
        __syncthreads();
        if (threadIdx.x == 0) carry += tot;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SC_THREADS)
k_scan_final(const int32_t *__restrict__ in, uint64_t n, const long long *__restrict__ tile_off,
             int64_t *__restrict__ out /* n+1 */)
{
    uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
    int32_t v[SC_ITEMS];
    long long s = 0;
    for (int k = 0; k < SC_ITEMS; ++k) { v[k] = base + k < n ? in[base + k] : 0; s += v[k]; }
    long long tot;
    long long run = block_excl_scan(s, &tot) + tile_off[blockIdx.x];
    for (int k = 0; k < SC_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
        if (base + k + 1 == n) out[n] = run;
    }
}

// ---------------------------------------------------------- ray generators
__global__ void __launch_bounds__(256)
k_gen_parallel(float *__restrict__ rays, uint64_t nu, uint64_t nv, f3 o0, f3 du, f3 dv, f3 dir)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nu * nv) return;
    float fu = (float)(i % nu), fv = (float)(i / nu);
    float2 *p = reinterpret_cast<float2 *>(rays + 6 * i);
    float ox = __fmaf_rn(fu, du.x, __fmaf_rn(fv, dv.x, o0.x));
    float oy = __fmaf_rn(fu, du.y, __fmaf_rn(fv, dv.y, o0.y));
    float oz = __fmaf_rn(fu, du.z, __fmaf_rn(fv, dv.z, o0.z));
    p[0] = make_float2(ox, oy); p[1] = make_float2(oz, dir.x); p[2] = make_float2(dir.y, dir.z);
}

struct PinholeArgs { double m[9]; double eye[3]; };

__global__ void __launch_bounds__(256)
k_gen_pinhole(float *__restrict__ rays, uint32_t w, uint32_t h, PinholeArgs pa)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (uint64_t)w * h) return;
    double x = (double)(i % w) + 0.5, y = (double)(i / w) + 0.5;
    float2 *p = reinterpret_cast<float2 *>(rays + 6 * i);
    float dx = (float)(pa.m[0] * x + pa.m[1] * y + pa.m[2]);
    float dy = (float)(pa.m[3] * x + pa.m[4] * y + pa.m[5]);
    float dz = (float)(pa.m[6] * x + pa.m[7] * y + pa.m[8]);
    p[0] = make_float2((float)pa.eye[0], (float)pa.eye[1]);
    p[1] = make_float2((float)pa.eye[2], dx);
    p[2] = make_float2(dy, dz);
}

// hit-primitive marking (ray_casting.py:285-289): flags are OR-ed, benign races
__global__ void __launch_bounds__(256)
k_mark_hits(const uint32_t *__restrict__ geom, const uint32_t *__restrict__ prim, uint64_t N,
            const uint64_t *__restrict__ goff, const uint64_t *__restrict__ voff, uint32_t ngeoms,
            const uint32_t *__restrict__ idx, uint8_t *__restrict__ tri_hit, uint8_t *__restrict__ vert_hit)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t p = prim[i];
    if (p == QSMRT_INVALID) return;
    uint32_t g = geom ? geom[i] : 0u;
    if (g >= ngeoms) return;
    uint64_t t = goff[g] + p;
    if (tri_hit) tri_hit[t] = 1;
    if (vert_hit) { vert_hit[idx[3 * t]] = 1; vert_hit[idx[3 * t + 1]] = 1; vert_hit[idx[3 * t + 2]] = 1; }
    (void)voff;
}

__global__ void __launch_bounds__(256)
k_accumulate_hits(const uint32_t *__restrict__ geom, const uint32_t *__restrict__ prim, uint64_t N,
                  const uint64_t *__restrict__ goff, uint32_t ngeoms, uint32_t *__restrict__ tri_counts)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    uint32_t p = prim[i];
    if (p == QSMRT_INVALID) return;
    uint32_t g = geom ? geom[i] : 0u;
    if (g >= ngeoms) return;
    atomicAdd(&tri_counts[goff[g] + p], 1u);
}

// per-vertex exposure: a vertex inherits the hits of every triangle it is a corner of (scene order; the
// concatenated index array is already rebased into scene vertex numbering)
__global__ void __launch_bounds__(256)
k_vertex_exposure(const uint32_t *__restrict__ idx, uint64_t ntris, const uint32_t *__restrict__ tri_counts, uint32_t *__restrict__ vert_counts)
{
    uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= ntris) return;
    const uint32_t c = tri_counts[t];
    if (c == 0u) return;
    atomicAdd(&vert_counts[idx[3 * t]], c); atomicAdd(&vert_counts[idx[3 * t + 1]], c); atomicAdd(&vert_counts[idx[3 * t + 2]], c);
}

// ------------------------------------------------------------ closest points
// Open3D ComputeClosestPoints / ComputeDistance (rtcPointQuery + ClosestPointFunc;
// reference: compute_signed_distance at pyQSM/viz/ray_casting.py:250,255).
// Bit-for-bit twin of oracle/qsmrt_oracle.c::cp_triangle / cp_one: Ericson's
// region test on a = v0, ab = -e1, ac = e2; smallest squared distance wins,
// exact ties go to the lowest (geometry, primitive).
struct CpBest { float d2; f3 q; float u, v; uint32_t geom, prim, tri; };

__device__ __forceinline__ f3 f3sub(f3 a, f3 b) { return f3{ __fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z) }; }
__device__ __forceinline__ f3 f3add(f3 a, f3 b) { return f3{ __fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z) }; }
__device__ __forceinline__ f3 f3madd(f3 a, float s, f3 b) { return f3{ __fmaf_rn(s, b.x, a.x), __fmaf_rn(s, b.y, a.y), __fmaf_rn(s, b.z, a.z) }; }

__device__ __forceinline__ void cp_triangle(const float4 p0, const float4 p1, const float4 p2, f3 p, f3 &q, float &bu, float &bv)
{
    f3 a = { p0.x, p0.y, p0.z }, ab = { -p1.x, -p1.y, -p1.z }, ac = { p2.x, p2.y, p2.z };
    f3 b = f3add(a, ab), c = f3add(a, ac);
    f3 ap = f3sub(p, a);
    float d1 = f3dot(ab, ap), d2 = f3dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { q = a; bu = 0.0f; bv = 0.0f; return; }
    f3 bp = f3sub(p, b);
    float d3 = f3dot(ab, bp), d4 = f3dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { q = b; bu = 1.0f; bv = 0.0f; return; }
    f3 cp = f3sub(p, c);
    float d5 = f3dot(ab, cp), d6 = f3dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { q = c; bu = 0.0f; bv = 1.0f; return; }
    float vc = __fmaf_rn(d1, d4, -__fmul_rn(d3, d2));
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) { float v = __fdiv_rn(d1, __fsub_rn(d1, d3)); q = f3madd(a, v, ab); bu = v; bv = 0.0f; return; }
    float vb = __fmaf_rn(d5, d2, -__fmul_rn(d1, d6));
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) { float w = __fdiv_rn(d2, __fsub_rn(d2, d6)); q = f3madd(a, w, ac); bu = 0.0f; bv = w; return; }
    float va = __fmaf_rn(d3, d6, -__fmul_rn(d5, d4));
    float e43 = __fsub_rn(d4, d3), e56 = __fsub_rn(d5, d6);
    if (va <= 0.0f && e43 >= 0.0f && e56 >= 0.0f) {
        float w = __fdiv_rn(e43, __fadd_rn(e43, e56));
        q = f3madd(b, w, f3sub(c, b)); bu = __fsub_rn(1.0f, w); bv = w; return;
    }
    float denom = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(va, vb), vc));
    float v = __fmul_rn(vb, denom), w = __fmul_rn(vc, denom);
    q = f3madd(f3madd(a, v, ab), w, ac); bu = v; bv = w;
}

__device__ __forceinline__ float box_dist2(float lox, float hix, float loy, float hiy, float loz, float hiz, f3 p)
{
    float dx = fmaxf(fmaxf(__fsub_rn(lox, p.x), __fsub_rn(p.x, hix)), 0.0f);
    float dy = fmaxf(fmaxf(__fsub_rn(loy, p.y), __fsub_rn(p.y, hiy)), 0.0f);
    float dz = fmaxf(fmaxf(__fsub_rn(loz, p.z), __fsub_rn(p.z, hiz)), 0.0f);
    return __fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz)));
}

__global__ void __launch_bounds__(TR_BLOCK)
k_closest_points(SceneView sc, const float *__restrict__ pts, uint64_t N, float *__restrict__ closest, float *__restrict__ dist,
                 uint32_t *__restrict__ geom, uint32_t *__restrict__ prim, float2 *__restrict__ uv, float *__restrict__ nrm)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    uint64_t i = blockIdx.x * (uint64_t)TR_BLOCK + threadIdx.x;
    if (i >= N) return;
    int spill[TR_LSTACK];
    Stack st; st.s = sstack + threadIdx.x; st.loc = spill; st.sp = 0;
    const f3 p = { pts[3 * i], pts[3 * i + 1], pts[3 * i + 2] };
    CpBest best{ INFINITY, f3{ 0.0f, 0.0f, 0.0f }, 0.0f, 0.0f, QSMRT_INVALID, QSMRT_INVALID, 0u };
    if (sc.ntris) {
        int cur = 0;
        for (;;) {
            if (cur >= 0) {
                float4 a, b, c; int4 d;
                load_node(sc.nodes, cur, a, b, c, d);
                float d0 = box_dist2(a.x, a.y, a.z, a.w, c.x, c.y, p);
                float d1 = box_dist2(b.x, b.y, b.z, b.w, c.z, c.w, p);
                bool h0 = d0 <= best.d2, h1 = d1 <= best.d2;
                if (h0 & h1) {
                    bool swap = d1 < d0;
                    st.push(swap ? d.x : d.y);
                    cur = swap ? d.y : d.x;
                    continue;
                }
                if (h0) { cur = d.x; continue; }
                if (h1) { cur = d.y; continue; }
            } else {
                uint32_t ref = (uint32_t)~cur, first = ref >> 2, count = (ref & 3u) + 1u;
                for (uint32_t k = 0; k < count; ++k) {
                    float4 p0, p1, p2;
                    load_tri(sc.tris, first + k, p0, p1, p2);
                    f3 q; float u, v;
                    cp_triangle(p0, p1, p2, p, q, u, v);
                    f3 df = f3sub(q, p);
                    float d2 = f3dot(df, df);
                    uint32_t pg = __float_as_uint(p1.w), pp = __float_as_uint(p0.w);
                    bool better = (d2 < best.d2) | ((d2 == best.d2) & ((pg < best.geom) | ((pg == best.geom) & (pp < best.prim))));
                    if (better) { best.d2 = d2; best.q = q; best.u = u; best.v = v; best.geom = pg; best.prim = pp; best.tri = first + k; }
                }
            }
            if (st.sp == 0) break;
            cur = st.pop();
        }
    }
    const bool ok = best.prim != QSMRT_INVALID;
    if (closest) { closest[3 * i] = best.q.x; closest[3 * i + 1] = best.q.y; closest[3 * i + 2] = best.q.z; }
    if (dist) dist[i] = ok ? __fsqrt_rn(best.d2) : INFINITY;
    if (geom) geom[i] = best.geom;
    if (prim) prim[i] = best.prim;
    if (uv) uv[i] = make_float2(best.u, best.v);
    if (nrm) {
        float nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (ok) {
            float4 p0, p1, p2;
            load_tri(sc.tris, best.tri, p0, p1, p2);
            f3 Ng = f3cross(f3{ p2.x, p2.y, p2.z }, f3{ p1.x, p1.y, p1.z });
            float inv = __fdiv_rn(1.0f, __fsqrt_rn(f3dot(Ng, Ng)));
            nx = __fmul_rn(Ng.x, inv); ny = __fmul_rn(Ng.y, inv); nz = __fmul_rn(Ng.z, inv);
        }
        nrm[3 * i] = nx; nrm[3 * i + 1] = ny; nrm[3 * i + 2] = nz;
    }
}

// signed distance = distance with the sign of the occupancy (count parity)
__global__ void __launch_bounds__(256)
k_apply_sign(float *__restrict__ dist, const int32_t *__restrict__ counts, uint64_t N)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < N && (counts[i] & 1)) dist[i] = -dist[i];
}

// rays for the occupancy test of Open3D ComputeOccupancy (nsamples == 1): origin = point, direction (1,1,1)
__global__ void __launch_bounds__(256)
k_points_to_rays(const float *__restrict__ pts, float *__restrict__ rays, uint64_t N)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= N) return;
    float2 *p = reinterpret_cast<float2 *>(rays + 6 * i);
    p[0] = make_float2(pts[3 * i], pts[3 * i + 1]); p[1] = make_float2(pts[3 * i + 2], 1.0f); p[2] = make_float2(1.0f, 1.0f);
}

// peel projection bookkeeping: triangles hit in this layer leave the scene; their 3-D and projected areas
// are summed (ray_casting.py:285-301: hit-triangle surface area, 3-D and flattened along the view direction).
// Flags are per TRIANGLE (scene order), not per sorted record: a split sliver has several records.
__global__ void __launch_bounds__(256)
k_peel_update(const float *__restrict__ verts, const uint32_t *__restrict__ idx, uint64_t n, uint8_t *__restrict__ alive,
              uint8_t *__restrict__ hitflag, int32_t *__restrict__ layer_of, int layer, f3 dir, double *__restrict__ sums)
{
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    double cnt = 0.0, a3 = 0.0, ap = 0.0;
    if (t < n && hitflag[t]) {
        hitflag[t] = 0;
        if (alive[t]) {
            alive[t] = 0;
            if (layer_of) layer_of[t] = layer;
            const uint32_t i0 = idx[3 * t], i1 = idx[3 * t + 1], i2 = idx[3 * t + 2];
            const f3 v0 = { verts[3ull * i0], verts[3ull * i0 + 1], verts[3ull * i0 + 2] };
            const f3 e1 = { __fsub_rn(v0.x, verts[3ull * i1]), __fsub_rn(v0.y, verts[3ull * i1 + 1]), __fsub_rn(v0.z, verts[3ull * i1 + 2]) };
            const f3 e2 = { __fsub_rn(verts[3ull * i2], v0.x), __fsub_rn(verts[3ull * i2 + 1], v0.y), __fsub_rn(verts[3ull * i2 + 2], v0.z) };
            f3 Ng = f3cross(e2, e1);          // the triangle record's e2 x e1
            cnt = 1.0;
            a3 = 0.5 * sqrt((double)Ng.x * Ng.x + (double)Ng.y * Ng.y + (double)Ng.z * Ng.z);
            ap = 0.5 * fabs((double)Ng.x * dir.x + (double)Ng.y * dir.y + (double)Ng.z * dir.z);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o); a3 += __shfl_xor_sync(0xFFFFFFFFu, a3, o); ap += __shfl_xor_sync(0xFFFFFFFFu, ap, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0.0) { atomicAdd(&sums[0], cnt); atomicAdd(&sums[1], a3); atomicAdd(&sums[2], ap); }
}

inline unsigned grid_for(uint64_t n, int block) { return (unsigned)((n + block - 1) / block); }

} // namespace

// Closest point, one WARP per query: for the small batches the reference makes (256 random points,
// ray_casting.py:248-250) the per-thread kernel above is one long dependent chain -- 0.6 ms for 256 points, set by the
// slowest query.  Here a warp shares one stack and takes up to 32 entries per round: every lane expands its node
// (two child boxes against the warp's best distance so far) or tests its leaf, the hit children are pushed with
// ballot-compacted offsets, the best distance is min-reduced across the lanes.  The winner is chosen by the same
// rule (distance, then lowest geometry / primitive), so the answer is the one the per-thread kernel gives.
constexpr int CPW_STACK = 2048;          // entries per warp; above CPW_STACK - 256 the warp goes depth-first, which needs <= height (<= 160 here) more
constexpr int CPW_WARPS = TR_BLOCK / 32;

__global__ void __launch_bounds__(TR_BLOCK)
k_closest_points_warp(SceneView sc, const float *__restrict__ pts, uint64_t N, float *__restrict__ closest, float *__restrict__ dist,
                      uint32_t *__restrict__ geom, uint32_t *__restrict__ prim, float2 *__restrict__ uv, float *__restrict__ nrm)
{
    __shared__ int wstack[CPW_WARPS][CPW_STACK];
    const unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t i = blockIdx.x * (uint64_t)CPW_WARPS + w;
    if (i >= N) return;                                  // warp-uniform
    int *stk = wstack[w];
    const f3 p = { pts[3 * i], pts[3 * i + 1], pts[3 * i + 2] };
    CpBest best{ INFINITY, f3{ 0.0f, 0.0f, 0.0f }, 0.0f, 0.0f, QSMRT_INVALID, QSMRT_INVALID, 0u };
    float wbest = INFINITY;                              // min over the lanes' best.d2 (pruning bound)
    int sp = 0;
    if (sc.ntris) { if (lane == 0) stk[0] = 0; sp = 1; }
    __syncwarp();
    while (sp > 0) {
        const int take = sp > CPW_STACK - 256 ? 1 : min(sp, 32);
        const bool have = lane < take;
        const int cur = have ? stk[sp - 1 - lane] : 0;
        sp -= take;
        __syncwarp();
        int c0 = 0, c1 = 0;
        bool h0 = false, h1 = false;
        if (have) {
            if (cur >= 0) {
                float4 a, b, c; int4 d;
                load_node(sc.nodes, cur, a, b, c, d);
                const float d0 = box_dist2(a.x, a.y, a.z, a.w, c.x, c.y, p);
                const float d1 = box_dist2(b.x, b.y, b.z, b.w, c.z, c.w, p);
                h0 = d0 <= wbest; h1 = d1 <= wbest;
                // the nearer child is pushed last, so the next round's lane 0 (top of the stack) descends towards the point
                const bool swap = h0 && h1 && d0 < d1;
                c0 = swap ? d.y : d.x; c1 = swap ? d.x : d.y;
                if (swap) { const bool t_ = h0; h0 = h1; h1 = t_; }
            } else {
                const uint32_t ref = (uint32_t)~cur, first = ref >> 2, count = (ref & 3u) + 1u;
                for (uint32_t k = 0; k < count; ++k) {
                    float4 p0, p1, p2;
                    load_tri(sc.tris, first + k, p0, p1, p2);
                    f3 q; float u, v;
                    cp_triangle(p0, p1, p2, p, q, u, v);
                    const f3 df = f3sub(q, p);
                    const float d2 = f3dot(df, df);
                    const uint32_t pg = __float_as_uint(p1.w), pp = __float_as_uint(p0.w);
                    const bool better = (d2 < best.d2) | ((d2 == best.d2) & ((pg < best.geom) | ((pg == best.geom) & (pp < best.prim))));
                    if (better) { best.d2 = d2; best.q = q; best.u = u; best.v = v; best.geom = pg; best.prim = pp; best.tri = first + k; }
                }
            }
        }
        // tighten the bound, then push the hit children compacted: first every lane's c0, then every lane's c1
        float m = best.d2;
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(FULL, m, o));
        wbest = m;
        const unsigned m0 = __ballot_sync(FULL, h0), m1 = __ballot_sync(FULL, h1);
        const unsigned lt = (1u << lane) - 1u;
        if (h0) stk[sp + __popc(m0 & lt)] = c0;
        if (h1) stk[sp + __popc(m0) + __popc(m1 & lt)] = c1;
        sp += __popc(m0) + __popc(m1);
        __syncwarp();
    }
    // winner: smallest (d2, geometry, primitive) over the lanes
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(FULL, best.d2, o);
        const uint32_t og = __shfl_xor_sync(FULL, best.geom, o), op = __shfl_xor_sync(FULL, best.prim, o), ot = __shfl_xor_sync(FULL, best.tri, o);
        const float oqx = __shfl_xor_sync(FULL, best.q.x, o), oqy = __shfl_xor_sync(FULL, best.q.y, o), oqz = __shfl_xor_sync(FULL, best.q.z, o);
        const float ou = __shfl_xor_sync(FULL, best.u, o), ov = __shfl_xor_sync(FULL, best.v, o);
        const bool better = (od < best.d2) | ((od == best.d2) & ((og < best.geom) | ((og == best.geom) & (op < best.prim))));
        if (better) { best.d2 = od; best.geom = og; best.prim = op; best.tri = ot; best.q = f3{ oqx, oqy, oqz }; best.u = ou; best.v = ov; }
    }
    if (lane != 0) return;
    const bool ok = best.prim != QSMRT_INVALID;
    if (closest) { closest[3 * i] = best.q.x; closest[3 * i + 1] = best.q.y; closest[3 * i + 2] = best.q.z; }
    if (dist) dist[i] = ok ? __fsqrt_rn(best.d2) : INFINITY;
    if (geom) geom[i] = best.geom;
    if (prim) prim[i] = best.prim;
    if (uv) uv[i] = make_float2(best.u, best.v);
    if (nrm) {
        float nx = 0.0f, ny = 0.0f, nz = 0.0f;
        if (ok) {
            float4 p0, p1, p2;
            load_tri(sc.tris, best.tri, p0, p1, p2);
            f3 Ng = f3cross(f3{ p2.x, p2.y, p2.z }, f3{ p1.x, p1.y, p1.z });
            float inv = __fdiv_rn(1.0f, __fsqrt_rn(f3dot(Ng, Ng)));
            nx = __fmul_rn(Ng.x, inv); ny = __fmul_rn(Ng.y, inv); nz = __fmul_rn(Ng.z, inv);
        }
        nrm[3 * i] = nx; nrm[3 * i + 1] = ny; nrm[3 * i + 2] = nz;
    }
}

// --------------------------------------------------------------- launchers
// All launch state is per scene (TrvState, traverse.h): tuning, the ring of work cursors of the persistent
// kernels (launches in flight never share one), the fetch-counter buffer and the occupancy cache.
namespace {
constexpr int CURSOR_RING = 256;

int next_cursor(TrvState &ts, unsigned long long **out, cudaStream_t st)
{
    if (!ts.cursor_ring) CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&ts.cursor_ring), CURSOR_RING * sizeof(unsigned long long)));
    unsigned k = ts.cursor_next++ % CURSOR_RING;
    *out = ts.cursor_ring + k;
    CUDA_TRY(cudaMemsetAsync(*out, 0, sizeof(unsigned long long), st));
    return 0;
}

int stats_buffer(TrvState &ts, unsigned long long **out, cudaStream_t st)
{
    if (!ts.stats) CUDA_TRY(cudaMalloc(reinterpret_cast<void **>(&ts.stats), 16 * sizeof(unsigned long long)));
    CUDA_TRY(cudaMemsetAsync(ts.stats, 0, 16 * sizeof(unsigned long long), st));
    *out = ts.stats;
    return 0;
}

template <int MODE, bool COUNTERS, bool QUANT>
int launch_trace5_q(TrvState &ts, TraceArgs &a, size_t smem, cudaStream_t st)
{
    // occupancy of this instantiation at this stack size: queried once per scene (small launches are latency bound)
    const void *fn = reinterpret_cast<const void *>(&k_trace5<MODE, COUNTERS, QUANT>);
    int per_sm = 0;
    for (const TrvState::Occ &o : ts.occ) if (o.fn == fn && o.smem == smem) per_sm = o.per_sm;
    if (per_sm == 0) {
        if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_trace5<MODE, COUNTERS, QUANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_trace5<MODE, COUNTERS, QUANT>, TR_BLOCK, smem));
        if (per_sm < 1) { qsmrt_set_error("persistent kernel does not fit (smem %zu)", smem); return 1; }
        ts.occ.push_back(TrvState::Occ{ fn, smem, per_sm });
    }
    if (ts.sms == 0) CUDA_TRY(cudaDeviceGetAttribute(&ts.sms, cudaDevAttrMultiProcessorCount, ts.device));
    if (ts.opt.ctas_per_sm > 0) per_sm = std::min(per_sm, ts.opt.ctas_per_sm);
    unsigned g = (unsigned)std::min<uint64_t>((uint64_t)per_sm * ts.sms, (a.nslots + TR_BLOCK - 1) / TR_BLOCK);
    k_trace5<MODE, COUNTERS, QUANT><<<g, TR_BLOCK, smem, st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// common launch of the persistent kernel: tuning, work cursor, occupancy-sized grid
template <int MODE, bool COUNTERS>
int launch_trace5(TrvState &ts, TraceArgs &a, size_t smem, cudaStream_t st)
{
    a.refill = ts.opt.refill; a.want = std::max(1, ts.opt.want); a.tri_min = std::max(1, ts.opt.tri_min);
    a.row_major = ts.opt.tile_order == 0;
    a.node_path = a.sc.node_tex ? ts.opt.node_path : 0;
    if (next_cursor(ts, &a.cursor, st)) return 1;
    if (a.sc.qnodes) return launch_trace5_q<MODE, COUNTERS, true>(ts, a, smem, st);
    return launch_trace5_q<MODE, COUNTERS, false>(ts, a, smem, st);
}

inline size_t stack_bytes(const SceneView &sc) { return (size_t)((int)sc.height + 2) * TR_BLOCK * sizeof(int); }
inline bool use_v5(const TrvState &ts, const SceneView &sc, size_t smem) { return ts.opt.variant == 2 && sc.ntris && smem <= 96 * 1024; }

uint64_t slots_for(uint64_t N, uint32_t row_len)
{
    if (!row_len) return N;
    uint64_t rows = N / row_len;
    return (uint64_t)((row_len + 7u) / 8u) * ((rows + 3) / 4) * 32u;
}
} // namespace

void trv_state_free(TrvState &ts)
{
    if (ts.cursor_ring) cudaFree(ts.cursor_ring);
    if (ts.stats) cudaFree(ts.stats);
    ts.cursor_ring = nullptr; ts.stats = nullptr; ts.occ.clear();
}

int trv_read_counters(TrvState &ts, unsigned long long out[16])
{
    memset(out, 0, 16 * sizeof(unsigned long long));
    if (ts.stats) CUDA_TRY(cudaMemcpy(out, ts.stats, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}

int trv_cast_rays(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, uint32_t row_len, float *t_hit, uint32_t *geom,
                  uint32_t *prim, float *uv, float *nrm, cudaStream_t st)
{
    if (N == 0) return 0;
    unsigned grid = grid_for(N, TR_BLOCK);
    if (row_len) {
        if (N % row_len) { qsmrt_set_error("ray count %llu is not a multiple of the row length %u", (unsigned long long)N, row_len); return 1; }
        uint64_t rows = N / row_len, warps = (uint64_t)((row_len + 7u) / 8u) * ((rows + 3) / 4);
        grid = (unsigned)((warps + TR_BLOCK / 32 - 1) / (TR_BLOCK / 32));
    }
    if (use_v5(ts, sc, stack_bytes(sc))) {
        TraceArgs a{};
        a.sc = sc; a.src.kind = 0; a.src.rays = rays; a.N = N; a.row_len = row_len; a.nslots = slots_for(N, row_len);
        a.out = CastOut{ t_hit, geom, prim, reinterpret_cast<float2 *>(uv), nrm };
        a.tnear = 0.0f; a.tfar = INFINITY; a.depth = (int)sc.height + 2;
        if (ts.opt.counters) {
            if (stats_buffer(ts, &a.stats, st)) return 1;
            if (launch_trace5<0, true>(ts, a, stack_bytes(sc), st)) return 1;
        } else if (launch_trace5<0, false>(ts, a, stack_bytes(sc), st)) return 1;
    } else {
        // per-thread loop: the simple kernel (A/B reference; also used when the LBVH is too deep for the shared-memory stack)
        k_cast_rays<<<grid, TR_BLOCK, 0, st>>>(sc, rays, N, row_len, t_hit, geom, prim, reinterpret_cast<float2 *>(uv), nrm);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_count(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, int32_t *out, uint32_t ngeoms, cudaStream_t st)
{
    if (N == 0) return 0;
    const int depth = (int)sc.height + 2;
    const int set_cap = std::min(std::max(ts.opt.count_set, 4), CNT_SET);
    const size_t smem = (size_t)depth * TR_BLOCK * sizeof(int);
    if (use_v5(ts, sc, smem)) {
        TraceArgs a{};
        a.sc = sc; a.src.kind = 0; a.src.rays = rays; a.N = N; a.row_len = 0; a.nslots = N;
        a.counts = out; a.depth = depth; a.multi_geom = ngeoms > 1; a.set_cap = set_cap;
        if (launch_trace5<2, false>(ts, a, smem, st)) return 1;
        const int sms = ts.sms ? ts.sms : 148;
        k_count_fix<<<(unsigned)std::min<uint64_t>(grid_for(N, TR_BLOCK), (uint64_t)sms * 8), TR_BLOCK, 0, st>>>(sc, rays, N, out);
        CUDA_TRY(cudaGetLastError());
        return 0;
    }
    k_count_intersections<<<grid_for(N, TR_BLOCK), TR_BLOCK, 0, st>>>(sc, rays, N, out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_occluded(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, float tnear, float tfar, uint8_t *out, cudaStream_t st)
{
    if (N == 0) return 0;
    if (use_v5(ts, sc, stack_bytes(sc))) {
        TraceArgs a{};
        a.sc = sc; a.src.kind = 0; a.src.rays = rays; a.N = N; a.row_len = 0; a.nslots = N;
        a.occluded = out; a.tnear = tnear; a.tfar = tfar; a.depth = (int)sc.height + 2;
        if (launch_trace5<1, false>(ts, a, stack_bytes(sc), st)) return 1;
    } else {
        k_test_occlusions<<<grid_for(N, TR_BLOCK), TR_BLOCK, 0, st>>>(sc, rays, N, tnear, tfar, out);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// list_intersections, the traversal: counts[i] = distinct hits of ray i (exact), and the hit records of the rays with
// 1 .. *max_fast_out of them in the stash (stash.cap records; *stash.count says afterwards how many the batch
// wanted: more than the capacity means records were dropped and the caller repeats with room).  *max_fast_out = 0:
// no stash (per-thread kernels; every ray takes k_list_finish's enumeration path).  Asynchronous.
int trv_list_collect(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, uint32_t ngeoms, int32_t *counts,
                     const ListStash &stash, int *max_fast_out, cudaStream_t st)
{
    *max_fast_out = 0;
    if (N == 0 || sc.ntris == 0) return 0;
    const int depth = (int)sc.height + 2;
    const int set_cap = std::min(std::max(ts.opt.count_set, 4), CNT_SET);
    const size_t smem = (size_t)depth * TR_BLOCK * sizeof(int);
    if (!use_v5(ts, sc, smem) || !stash.count) return trv_count(ts, sc, rays, N, counts, ngeoms, st);
    CUDA_TRY(cudaMemsetAsync(stash.count, 0, sizeof(unsigned long long), st));
    TraceArgs a{};
    a.sc = sc; a.src.kind = 0; a.src.rays = rays; a.N = N; a.row_len = 0; a.nslots = N;
    a.counts = counts; a.depth = depth; a.multi_geom = ngeoms > 1; a.set_cap = set_cap;
    a.st_base = stash.base; a.st_t = stash.t; a.st_geom = stash.geom; a.st_prim = stash.prim; a.st_uv = reinterpret_cast<float2 *>(stash.uv);
    a.st_count = stash.count; a.st_cap = stash.cap;
    if (launch_trace5<6, false>(ts, a, smem, st)) return 1;
    const int sms = ts.sms ? ts.sms : 148;
    k_count_fix<<<(unsigned)std::min<uint64_t>(grid_for(N, TR_BLOCK), (uint64_t)sms * 8), TR_BLOCK, 0, st>>>(sc, rays, N, counts);
    CUDA_TRY(cudaGetLastError());
    *max_fast_out = set_cap;
    return 0;
}

// list_intersections, the output: stash -> CSR arrays at the scanned offsets, each ray's segment in order
int trv_list_emit(const SceneView &sc, const float *rays, uint64_t N, const int64_t *splits, const ListStash &stash, int max_fast,
                  int64_t *ray_ids, float *t_hit, uint32_t *geom, uint32_t *prim, float *uv, cudaStream_t st)
{
    if (N == 0 || sc.ntris == 0) return 0;
    k_list_finish<<<grid_for(N, TR_BLOCK), TR_BLOCK, 0, st>>>(sc, rays, N, splits, max_fast, stash.base, stash.t, stash.geom, stash.prim,
                                                            reinterpret_cast<const float2 *>(stash.uv), ray_ids, t_hit, geom, prim,
                                                            reinterpret_cast<float2 *>(uv));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

size_t trv_scan_scratch_bytes(uint64_t n) { return ((n + SC_TILE - 1) / SC_TILE + 1) * sizeof(long long); }

// out[0..n] = exclusive scan of in[0..n), out[n] = total
int trv_exclusive_scan(const int32_t *in, uint64_t n, int64_t *out, void *scratch, cudaStream_t st)
{
    if (n == 0) { CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(int64_t), st)); return 0; }
    uint32_t ntiles = (uint32_t)((n + SC_TILE - 1) / SC_TILE);
    long long *ts = reinterpret_cast<long long *>(scratch);
    k_scan_tile_sums<<<ntiles, SC_THREADS, 0, st>>>(in, n, ts);
    k_scan_tile_offsets<<<1, SC_THREADS, 0, st>>>(ts, ntiles);
    k_scan_final<<<ntiles, SC_THREADS, 0, st>>>(in, n, ts, out);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_gen_parallel(float *rays, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                     const float dv[3], const float dir[3], cudaStream_t st)
{
    uint64_t n = nu * nv;
    if (n == 0) return 0;
    f3 a = { o0[0], o0[1], o0[2] }, b = { du[0], du[1], du[2] }, c = { dv[0], dv[1], dv[2] }, d = { dir[0], dir[1], dir[2] };
    k_gen_parallel<<<grid_for(n, 256), 256, 0, st>>>(rays, nu, nv, a, b, c, d);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_gen_pinhole(float *rays, uint32_t w, uint32_t h, const double minv[9], const double eye[3], cudaStream_t st)
{
    uint64_t n = (uint64_t)w * h;
    if (n == 0) return 0;
    PinholeArgs pa;
    for (int k = 0; k < 9; ++k) pa.m[k] = minv[k];
    for (int k = 0; k < 3; ++k) pa.eye[k] = eye[k];
    k_gen_pinhole<<<grid_for(n, 256), 256, 0, st>>>(rays, w, h, pa);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_mark_hits(const uint32_t *geom, const uint32_t *prim, uint64_t N, const uint64_t *goff,
                  const uint64_t *voff, uint32_t ngeoms, const uint32_t *idx, uint8_t *tri_hit,
                  uint8_t *vert_hit, cudaStream_t st)
{
    if (N == 0) return 0;
    k_mark_hits<<<grid_for(N, 256), 256, 0, st>>>(geom, prim, N, goff, voff, ngeoms, idx, tri_hit, vert_hit);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_accumulate_hits(const uint32_t *geom, const uint32_t *prim, uint64_t N, const uint64_t *goff,
                        uint32_t ngeoms, uint32_t *tri_counts, cudaStream_t st)
{
    if (N == 0) return 0;
    k_accumulate_hits<<<grid_for(N, 256), 256, 0, st>>>(geom, prim, N, goff, ngeoms, tri_counts);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_vertex_exposure(const uint32_t *idx, uint64_t ntris, const uint32_t *tri_counts, uint32_t *vert_counts, cudaStream_t st)
{
    if (ntris == 0) return 0;
    k_vertex_exposure<<<grid_for(ntris, 256), 256, 0, st>>>(idx, ntris, tri_counts, vert_counts);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// Fused drivers: rays generated inside the persistent kernel, results reduced on the device.
int trv_sun_exposure(TrvState &ts, const SceneView &sc, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                     const float dv[3], const float dir[3], const uint64_t *goff, uint32_t *tri_counts, cudaStream_t st)
{
    if (nu * nv == 0 || sc.ntris == 0) return 0;
    if (!use_v5(ts, sc, stack_bytes(sc))) { qsmrt_set_error("sun_exposure needs the persistent kernel (BVH height %u too deep?)", sc.height); return 1; }
    TraceArgs a{};
    a.sc = sc; a.src.kind = 1; a.src.nu = nu;
    a.src.o0 = f3{ o0[0], o0[1], o0[2] }; a.src.du = f3{ du[0], du[1], du[2] };
    a.src.dv = f3{ dv[0], dv[1], dv[2] }; a.src.dir = f3{ dir[0], dir[1], dir[2] };
    a.N = nu * nv; a.row_len = nu >= 8 && nu < (1ull << 32) ? (uint32_t)nu : 0; a.nslots = slots_for(a.N, a.row_len);
    a.accum = tri_counts; a.goff = goff; a.tnear = 0.0f; a.tfar = INFINITY; a.depth = (int)sc.height + 2;
    return launch_trace5<3, false>(ts, a, stack_bytes(sc), st);
}

// a whole sweep of parallel grids (solar angles) in one launch: sweep_dev[n_grids][12] = origin0, du, dv, dir
int trv_sun_exposure_sweep(TrvState &ts, const SceneView &sc, uint32_t n_grids, const float *sweep_dev, uint64_t nu, uint64_t nv,
                           const uint64_t *goff, uint32_t *tri_counts, uint64_t count_stride, cudaStream_t st)
{
    if (n_grids == 0 || nu * nv == 0 || sc.ntris == 0) return 0;
    if (!use_v5(ts, sc, stack_bytes(sc))) { qsmrt_set_error("sun_exposure needs the persistent kernel (BVH height %u too deep?)", sc.height); return 1; }
    TraceArgs a{};
    a.sc = sc; a.src.kind = 3; a.src.nu = nu; a.src.sweep = sweep_dev;
    a.src.per_grid_rays = nu * nv;
    a.row_len = nu >= 8 && nu < (1ull << 32) ? (uint32_t)nu : 0;
    a.src.per_grid_slots = slots_for(a.src.per_grid_rays, a.row_len);
    a.N = a.src.per_grid_rays * n_grids; a.nslots = a.src.per_grid_slots * n_grids;
    a.accum = tri_counts; a.goff = goff; a.accum_stride = count_stride; a.tnear = 0.0f; a.tfar = INFINITY; a.depth = (int)sc.height + 2;
    return launch_trace5<3, false>(ts, a, stack_bytes(sc), st);
}

static RaySource hemisphere_source(const float *points, const float *normals, uint64_t point_base, uint32_t dir_begin,
                                    uint32_t dir_count, uint64_t seed, float offset)
{
    RaySource s{};
    s.kind = 2; s.points = points; s.normals = normals; s.dir_begin = dir_begin; s.dir_count = dir_count;
    s.seed = seed; s.offset = offset; s.point_base = point_base;
    return s;
}

// Morton key of a query point inside the scene's box (10 bits per axis), value = its index: the sky driver works the
// points in this order, so that the rays in flight at any time start in one region of the scene
__global__ void __launch_bounds__(256)
k_point_keys(const float *__restrict__ points, uint64_t n, f3 lo, f3 scale, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    const uint64_t i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= n) return;
    const float c[3] = { (points[3 * i] - lo.x) * scale.x, (points[3 * i + 1] - lo.y) * scale.y, (points[3 * i + 2] - lo.z) * scale.z };
    uint64_t key = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        uint32_t q = (uint32_t)fminf(fmaxf(c[a], 0.0f), 1023.0f);          // NaN -> 0
        q = (q | (q << 16)) & 0x030000FFu; q = (q | (q << 8)) & 0x0300F00Fu; q = (q | (q << 4)) & 0x030C30C3u; q = (q | (q << 2)) & 0x09249249u;
        key |= (uint64_t)q << a;
    }
    keys[i] = key; vals[i] = (uint32_t)i;
}

int trv_point_keys(const float *points, uint64_t n, const float lo[3], const float hi[3], uint64_t *keys, uint32_t *vals, cudaStream_t st)
{
    if (n == 0) return 0;
    f3 l{ lo[0], lo[1], lo[2] }, sc;
    sc.x = hi[0] > lo[0] ? 1024.0f / (hi[0] - lo[0]) : 0.0f; sc.y = hi[1] > lo[1] ? 1024.0f / (hi[1] - lo[1]) : 0.0f;
    sc.z = hi[2] > lo[2] ? 1024.0f / (hi[2] - lo[2]) : 0.0f;
    k_point_keys<<<grid_for(n, 256), 256, 0, st>>>(points, n, l, sc, keys, vals);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_sky_visibility(TrvState &ts, const SceneView &sc, const float *points, const float *normals, uint64_t n_points, uint64_t point_base,
                       uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, const uint32_t *perm, uint32_t *unoccluded,
                       cudaStream_t st)
{
    if (n_points == 0 || dir_count == 0) return 0;
    if (ts.opt.variant != 2 || stack_bytes(sc) > 96 * 1024) { qsmrt_set_error("sky_visibility needs the persistent kernel"); return 1; }
    TraceArgs a{};
    a.sc = sc; a.src = hemisphere_source(points, normals, point_base, dir_begin, dir_count, seed, offset);
    a.src.perm = perm;
    a.N = n_points * (uint64_t)dir_count; a.row_len = 0; a.nslots = a.N;
    a.accum = unoccluded; a.tnear = 0.0f; a.tfar = INFINITY; a.depth = (int)sc.height + 2;
    return launch_trace5<4, false>(ts, a, stack_bytes(sc), st);
}

int trv_gen_hemisphere(float *rays, const float *points, const float *normals, uint64_t n_points, uint64_t point_base,
                       uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, cudaStream_t st)
{
    uint64_t n = n_points * (uint64_t)dir_count;
    if (n == 0) return 0;
    k_gen_hemisphere<<<grid_for(n, 256), 256, 0, st>>>(rays, hemisphere_source(points, normals, point_base, dir_begin, dir_count, seed, offset), n);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_closest_points(TrvState &ts, const SceneView &sc, const float *pts, uint64_t N, float *closest, float *dist, uint32_t *geom,
                       uint32_t *prim, float *uv, float *nrm, cudaStream_t st)
{
    if (N == 0) return 0;
    if (N <= (uint64_t)ts.opt.cp_warp_max && sc.height + 2u <= 160u)       // small batch: one warp per query (latency), else one thread (throughput)
        k_closest_points_warp<<<(unsigned)((N + CPW_WARPS - 1) / CPW_WARPS), TR_BLOCK, 0, st>>>(sc, pts, N, closest, dist, geom, prim, reinterpret_cast<float2 *>(uv), nrm);
    else
    k_closest_points<<<grid_for(N, TR_BLOCK), TR_BLOCK, 0, st>>>(sc, pts, N, closest, dist, geom, prim, reinterpret_cast<float2 *>(uv), nrm);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_points_to_rays(const float *pts, float *rays, uint64_t N, cudaStream_t st)
{
    if (N == 0) return 0;
    k_points_to_rays<<<grid_for(N, 256), 256, 0, st>>>(pts, rays, N);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int trv_apply_sign(float *dist, const int32_t *counts, uint64_t N, cudaStream_t st)
{
    if (N == 0) return 0;
    k_apply_sign<<<grid_for(N, 256), 256, 0, st>>>(dist, counts, N);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// one layer of the peel projection: cast the grid against the triangles still alive, flag the owners of the closest hits
int trv_peel_cast(TrvState &ts, const SceneView &sc, uint64_t nu, uint64_t nv, const float o0[3], const float du[3], const float dv[3],
                  const float dir[3], const uint32_t *order, const uint8_t *alive, uint8_t *hitflag, cudaStream_t st)
{
    if (nu * nv == 0 || sc.ntris == 0) return 0;
    if (!use_v5(ts, sc, stack_bytes(sc))) { qsmrt_set_error("peel projection needs the persistent kernel"); return 1; }
    TraceArgs a{};
    a.sc = sc; a.src.kind = 1; a.src.nu = nu;
    a.src.o0 = f3{ o0[0], o0[1], o0[2] }; a.src.du = f3{ du[0], du[1], du[2] };
    a.src.dv = f3{ dv[0], dv[1], dv[2] }; a.src.dir = f3{ dir[0], dir[1], dir[2] };
    a.N = nu * nv; a.row_len = nu >= 8 && nu < (1ull << 32) ? (uint32_t)nu : 0; a.nslots = slots_for(a.N, a.row_len);
    a.alive = alive; a.hitflag = hitflag; a.order = order; a.tnear = 0.0f; a.tfar = INFINITY; a.depth = (int)sc.height + 2;
    return launch_trace5<5, false>(ts, a, stack_bytes(sc), st);
}

int trv_peel_update(const float *verts, const uint32_t *idx, uint64_t ntris, uint8_t *alive, uint8_t *hitflag, int32_t *layer_of,
                    int layer, const float dir[3], double *sums, cudaStream_t st)
{
    if (ntris == 0) return 0;
    float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    f3 d = len > 0.0f ? f3{ dir[0] / len, dir[1] / len, dir[2] / len } : f3{ 0.0f, 0.0f, 0.0f };
    k_peel_update<<<grid_for(ntris, 256), 256, 0, st>>>(verts, idx, ntris, alive, hitflag, layer_of, layer, d, sums);
    CUDA_TRY(cudaGetLastError());
    return 0;
}
