// trace_persistent.cuh -- the traversal kernel that ships (v5, sm_100a),
// included inside traverse.cu's anonymous namespace.
//
// Persistent warps, warp-uniform control flow (see the v3 notes in
// traverse.cu for the ncu evidence that motivated it), refined by what the
// v3/v4 profiles and the tuning sweep showed:
//   * the whole traversal stack lives in shared memory ([entry][thread],
//     depth = LBVH height + 2 known from the build), so push/pop are single
//     predicated STS/LDS with the stack pointer in a register -- no local
//     spill path, no branches;
//   * nodes are fetched with two 256-bit loads (LDG.E.256, new on sm_100);
//   * the node phase ends once fewer than `want` lanes are still looking for
//     a leaf (and somebody holds triangles), the triangle phase drains one
//     triangle per lane per iteration;
//   * finished lanes are refilled from a global cursor, compacted with
//     ballot/popc, once `refill` lanes are idle.
//
// MODE (what is done with the hits)
//   0  closest hit  -> t_hit / ids / uv / normal            (cast_rays)
//   1  any hit      -> occluded[ray]                        (test_occlusions)
//   2  all hits     -> number of distinct (geometry, t)     (count_intersections)
//                      kept per lane in a shared-memory set [slot][thread]; a ray
//                      with more than CNT_SET distinct hits is marked -1 and
//                      finished exactly by k_count_fix
//   3  closest hit  -> atomicAdd(accum[triangle], 1)        (sun exposure: no 32 B/ray of results)
//   4  any hit      -> atomicAdd(accum[ray / n_dirs], !hit) (sky visibility per query point)
//   5  closest hit among the triangles still alive -> hitflag[sorted triangle] = 1   (peel projection)
//   6  all hits     -> MODE 2's count AND the distinct hits themselves (list_intersections in ONE traversal): the
//                      per-lane set also keeps primitive id and uv (equal (geometry, t) keeps the lowest primitive
//                      id); a retiring warp reserves room for its rays' records in a stash with one atomicAdd
//                      and writes them there; stash_base[ray] says where.  k_list_finish later moves them to
//                      the caller's CSR arrays in (t, geometry, primitive) order.  Rays with more than CNT_SET
//                      distinct hits are marked -1 like MODE 2 and enumerated by the slow path.
// SRC (where rays come from; a uniform runtime switch, only touched at refill)
//   0  rays[N][6] in memory
//   1  parallel grid: origin0 + i*du + j*dv, direction dir  (same arithmetic as k_gen_parallel)
//   2  hemisphere Monte-Carlo about +z from points[n][3] (+ offset along normals[n][3]),
//      direction k of point p from a counter-based hash of (seed, p, k) (same as k_gen_hemisphere)
//   3  a sweep of parallel grids in ONE launch: grid a of `sweep[a][12]` = (origin0, du, dv, dir), all nu x nv;
//      ray index = a * per_grid_rays + the index inside grid a (a solar sweep without 63 kernel tails)

__device__ __forceinline__ void ld256f(const void *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// ---- counter-based sampling of the upper hemisphere (uniform in solid angle)
__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;     // lowbias32
    return x;
}

__device__ __forceinline__ f3 hemisphere_dir(uint64_t seed, uint64_t point, uint32_t k)
{
    uint32_t h0 = hash32((uint32_t)point ^ hash32((uint32_t)(point >> 32) + 0x9E3779B9u) ^ hash32((uint32_t)seed));
    uint32_t a = hash32(h0 + 2u * k + (uint32_t)(seed >> 32));
    uint32_t b = hash32(a ^ (0x85EBCA6Bu + 2u * k + 1u));
    float u1 = (float)(a >> 8) * 5.9604644775390625e-08f;        // [0,1), exact
    float u2 = (float)(b >> 8) * 5.9604644775390625e-08f;
    float z = u1, rr = sqrtf(fmaxf(0.0f, 1.0f - z * z)), phi = 6.2831853071795864f * u2;
    float sn, cs;
    sincosf(phi, &sn, &cs);
    return f3{ rr * cs, rr * sn, z };
}

__device__ __forceinline__ Ray make_ray(f3 O, f3 D)
{
    Ray r;
    r.O = O; r.D = D;
    r.idx = safe_inv(D.x); r.idy = safe_inv(D.y); r.idz = safe_inv(D.z);
    r.oodx = O.x * r.idx; r.oody = O.y * r.idy; r.oodz = O.z * r.idz;
    return r;
}

struct RaySource {
    int kind;                                   // SRC above
    const float *rays;                          // 0
    f3 o0, du, dv, dir; uint64_t nu;            // 1
    const float *points, *normals; uint64_t seed; float offset;                     // 2
    uint64_t point_base;                        // 2: index of points[0] in the sample (hash of (seed, point_base + p, k))
    uint32_t dir_begin, dir_count;              // 2: this launch draws directions [dir_begin, dir_begin + dir_count) of each point
    const uint32_t *perm;                       // 2: the order the points are WORKED in (perm[j] = index of the j-th point; NULL: as stored)
    const float *sweep; uint64_t per_grid_rays, per_grid_slots;                    // 3 (nu as in 1)
};

// 64-bit / 32-bit-sized operands: the 64-bit software division is ~100 instructions, and every refill of the grid
// and Monte-Carlo sources used to run up to four of them per lane (measured: the fused sun sweep 10 % behind
// cast_rays + accumulate_hits for that reason alone); batches below 2^32 rays take the 32-bit path
__device__ __forceinline__ uint64_t fast_div(uint64_t n, uint64_t d)
{
    if (((n | d) >> 32) == 0) return (uint64_t)((uint32_t)n / (uint32_t)d);
    return n / d;
}

// x / y: column and row of ray i in its grid when the batch is walked in 2-D tiles (have_xy), which saves the grid
// sources the division of i by the row length
__device__ __forceinline__ Ray source_ray(const RaySource &S, uint64_t i, uint32_t x = 0, uint64_t y = 0, bool have_xy = false)
{
    if (S.kind == 0) return load_ray(S.rays, i);
    if (S.kind == 1) {
        if (!have_xy) { y = fast_div(i, S.nu); x = (uint32_t)(i - y * S.nu); }
        float fu = (float)x, fv = (float)y;
        f3 O = { __fmaf_rn(fu, S.du.x, __fmaf_rn(fv, S.dv.x, S.o0.x)),
                 __fmaf_rn(fu, S.du.y, __fmaf_rn(fv, S.dv.y, S.o0.y)),
                 __fmaf_rn(fu, S.du.z, __fmaf_rn(fv, S.dv.z, S.o0.z)) };
        return make_ray(O, S.dir);
    }
    if (S.kind == 3) {
        const uint64_t a = fast_div(i, S.per_grid_rays);
        const float4 *g = reinterpret_cast<const float4 *>(S.sweep + 12 * a);      // the same arithmetic as kind 1
        const float4 g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2);          // o0.xyz du.x | du.yz dv.xy | dv.z dir.xyz
        if (!have_xy) { const uint64_t li = i - a * S.per_grid_rays; y = fast_div(li, S.nu); x = (uint32_t)(li - y * S.nu); }
        float fu = (float)x, fv = (float)y;
        f3 O = { __fmaf_rn(fu, g0.w, __fmaf_rn(fv, g1.z, g0.x)),
                 __fmaf_rn(fu, g1.x, __fmaf_rn(fv, g1.w, g0.y)),
                 __fmaf_rn(fu, g1.y, __fmaf_rn(fv, g2.x, g0.z)) };
        return make_ray(O, f3{ g2.y, g2.z, g2.w });
    }
    uint64_t p = fast_div(i, S.dir_count);
    const uint32_t k = S.dir_begin + (uint32_t)(i - p * S.dir_count);
    if (S.perm) p = S.perm[p];
    f3 O = { S.points[3 * p], S.points[3 * p + 1], S.points[3 * p + 2] };
    if (S.normals) {
        O.x = __fmaf_rn(S.offset, S.normals[3 * p], O.x);
        O.y = __fmaf_rn(S.offset, S.normals[3 * p + 1], O.y);
        O.z = __fmaf_rn(S.offset, S.normals[3 * p + 2], O.z);
    }
    return make_ray(O, hemisphere_dir(S.seed, S.point_base + p, k));
}

struct TraceArgs {
    SceneView sc;
    RaySource src;
    uint64_t N; uint32_t row_len; uint64_t nslots;
    CastOut out; uint8_t *occluded; float tnear, tfar;          // MODE 0 / 1
    int32_t *counts;                                            // MODE 2
    long long *st_base; float *st_t; uint32_t *st_geom, *st_prim; float2 *st_uv;     // MODE 6: the stash of hit records,
    unsigned long long *st_count, st_cap;                                            //         records reserved so far, capacity
    uint32_t *accum; const uint64_t *goff; uint64_t accum_stride;   // MODE 3 / 4 (stride: one row of counts per grid of a sweep, 0 = one row)
    const uint8_t *alive; uint8_t *hitflag; const uint32_t *order;  // MODE 5: flags per triangle (scene order) = order[sorted record]
    unsigned long long *cursor, *stats;
    int refill, want, tri_min, node_path;
    int depth;          // stack entries per thread
    int multi_geom;     // > 1 geometry: the set also keeps geometry ids
    int row_major;      // 2-D batches: tile rows in memory order instead of from the middle outwards
    int set_cap;        // MODE 2 / 6: entries of the per-lane hit set in use (<= CNT_SET)
};

// Resident CTAs per SM the kernel is compiled for (register cap 51 at 10, 40 at 12).  Measured (profiles/r02_tuning.txt,
// profiles/r02_kernel_select.txt): 12 gains 1-2 % for cast_rays on the coherent 16M-ray grids of the canopy and loses
// everywhere else -- 3-10 % on cylinder-QSM scenes, 37-75 % on incoherent rays through the 2M-triangle canopy (the
// spills of the 40-register build sit in the refill path and incoherent rays leave them no L1) -- so it is 10 for all.
#ifndef QSMRT_TRACE_MINB
#define QSMRT_TRACE_MINB 10
#endif
template <int MODE, bool QUANT> struct TraceMinB { static constexpr int value = QSMRT_TRACE_MINB; };
// Node steps per phase vote (the vote only decides the phase, so voting less often saves the loop control all 32
// lanes execute, but lanes that finish inside the block idle until its end).  Round 1 (before the sweep source and
// the aggregated atomics): 1 -> 2 +8 %, 2 -> 4 +2 %, 4 -> 8 +2-3 % for cast_rays.  Re-measured in round 2
// (profiles/r02_tuning.txt): cast_rays C2 4 = 8 (4568 vs 4583 Mrays/s) while the short rays of the C1 tree gain
// 5-12 % with 4; the fused kernels, whose lanes retire cheaply and want prompt refills, prefer 2: sun sweep
// 3511 / 3854 / 4086 and sky 2038 / 2184 / 2444 Mrays/s at 8 / 4 / 2.
#ifndef QSMRT_STEPS_M0
#define QSMRT_STEPS_M0 4
#endif
#ifndef QSMRT_STEPS_M3
#define QSMRT_STEPS_M3 2
#endif
#ifndef QSMRT_STEPS_M4
#define QSMRT_STEPS_M4 2
#endif
#ifndef QSMRT_STEPS_OTHER
#define QSMRT_STEPS_OTHER 4
#endif
template <int MODE> struct NodeSteps {
    static constexpr int value = MODE == 0 ? QSMRT_STEPS_M0 : MODE == 3 ? QSMRT_STEPS_M3 : MODE == 4 ? QSMRT_STEPS_M4 : QSMRT_STEPS_OTHER;
};
constexpr int TR_TRI_STEPS = 2;      // triangle tests per phase vote (1: -1.5 %, 3: same)
// The FMA slab form t = plane * (1/d) - o * (1/d) rounds with an error of a few ulp OF THE CONSTANT o/d, which grows
// with the distance of the ray origin from the scene, while the boxes' padding is a fixed 2^-17 of the scene size
// (3 grid cells for the quantised nodes): for origins hundreds of scene sizes away the padding no longer covers it.
// Widening every slab interval by 2^-20 of its own t (8 ulp) keeps the test conservative at any distance; the two
// multiplications per child are immediate-form FMULs on the FMA pipe, which this ALU-bound loop leaves idle.
constexpr float SLAB_NEAR = 0.99999905f, SLAB_FAR = 1.00000095f;
constexpr int CNT_SET = 32;     // largest hit set: distinct (geometry, t) pairs a lane can hold before it defers to the exact slow path (TraceArgs.set_cap <= CNT_SET)

__device__ __forceinline__ void ld256u(const void *p, uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3,
                                       uint32_t &w4, uint32_t &w5, uint32_t &w6, uint32_t &w7)
{
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7)
                 : "l"(p));
}
// 16-bit grid coordinate -> the float 2^23 + q, one PRMT each (no int->float conversion).  Raw prmt.b32: the
// __byte_perm intrinsic masks a run-time selector first (an extra LOP3 per plane pair)
__device__ __forceinline__ float prmt_f(uint32_t w, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0x4B000000u), "r"(sel));
    return __uint_as_float(d);
}
__device__ __forceinline__ float qlo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610)); }
__device__ __forceinline__ float qhi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632)); }

// QUANT: read the 32-byte quantised nodes (one 256-bit load per node step instead of two; the L1 data pipe
// is the binding resource, profiles/README.md).  The ray's slab constants are folded with the grid:
// t = (2^23 + q) * (cell/d) - (2^23 * cell/d - (glo - o)/d), so the slab code is unchanged.
template <int MODE, bool COUNTERS, bool QUANT>
__global__ void __launch_bounds__(TR_BLOCK, (TraceMinB<MODE, QUANT>::value))
k_trace5(const TraceArgs A)
{
    constexpr bool CLOSEST = MODE == 0 || MODE == 3 || MODE == 5;
    constexpr bool ANYHIT = MODE == 1 || MODE == 4;
    extern __shared__ int sstack[];                 // [depth][TR_BLOCK]
    int *const sbase = sstack + threadIdx.x;
#define STACK_PUSH(v) do { *sptr = (v); sptr += TR_BLOCK; } while (0)
#define STACK_POP(dst) do { sptr -= TR_BLOCK; (dst) = *sptr; } while (0)
#define STACK_RESET() do { sbase[0] = TR_SENTINEL; sptr = sbase + TR_BLOCK; } while (0)
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const TNode *__restrict__ nodes = A.sc.nodes;
    const TriRec *__restrict__ tris = A.sc.tris;

    Ray r;
    float best_t = 0.0f; uint32_t best_geom = QSMRT_INVALID, best_prim = QSMRT_INVALID, best_tri = 0u;
    uint64_t ray_i = 0;
    bool have_ray = false, exhausted = false;
    int cur = TR_SENTINEL;
    int *sptr = sbase;                              // next free stack slot (register; stride TR_BLOCK ints)
    // the leaf a lane has parked for the next triangle phase, as its (negative) child reference ~(first << 2 | left - 1);
    // 0 = none (node 0 is the root, never a leaf).  One register and a two-instruction park: the node step used to
    // decode it into a first / end pair with four more ALU-pipe instructions per step, the pipe this loop is bound by.
    int lcur = 0;
    unsigned n_node = 0, n_tri = 0;
    // MODE 2 / 6 hit set in LOCAL memory (L1-cached, thread-interleaved): shared memory then holds only the stack and
    // 10 CTAs fit per SM instead of the 7 a shared-memory set allowed (measured: C3 count 381 -> 474, canopy count
    // 701 -> 781 Mrays/s); the set is touched once per accepted hit
    float tset_l[(MODE == 2 || MODE == 6) ? CNT_SET : 1];
    uint32_t gset_l[(MODE == 2 || MODE == 6) ? CNT_SET : 1];
    uint32_t pset_l[MODE == 6 ? CNT_SET : 1];                   // MODE 6: primitive id and uv of each set entry
    float uset_l[MODE == 6 ? CNT_SET : 1], vset_l[MODE == 6 ? CNT_SET : 1];
#define TSET(q) tset_l[q]
#define GSET(q) gset_l[q]
    int cnt = 0; bool overflow = false;
    uint32_t snx = 0x7610u, sny = 0x7610u, snz = 0x7610u;       // QUANT: per-axis "near plane" byte selectors

#define PARK_LEAF5() do { lcur = cur; STACK_POP(cur); } while (0)
    // after a triangle of the parked leaf has been tested: the next one (reference - 3: first + 1, left - 1), or done
#define NEXT_TRI5() do { if (((uint32_t)~lcur & 3u) != 0u) lcur -= 3; else { lcur = 0; if (cur < 0) PARK_LEAF5(); } } while (0)

    for (;;) {
        // ---- retire finished rays, refill idle lanes
        const bool idle = (cur == TR_SENTINEL) && (lcur == 0);
        const unsigned im = __ballot_sync(FULL, idle);
        if (im == FULL || (!exhausted && __popc(im) >= A.refill)) {
            if (MODE == 3 || MODE == 4) {
                // hit / miss aggregation (north_star item 4): the lanes retiring together are neighbours of a ray tile
                // (MODE 3: they mostly hit the same few triangles) or directions of one query point (MODE 4), so the
                // lanes with equal targets are grouped with MATCH.ANY and each group sends ONE atomicAdd of its popc
                const bool add = idle && have_ray && ((MODE == 3) == (best_prim != QSMRT_INVALID));
                const unsigned am = __ballot_sync(FULL, add);
                if (add) {
                    unsigned long long key;
                    if (MODE == 3) key = (A.goff ? A.goff[best_geom] : 0ull) + best_prim +
                                         (A.src.kind == 3 && A.accum_stride ? fast_div(ray_i, A.src.per_grid_rays) * A.accum_stride : 0ull);
                    else { key = fast_div(ray_i, A.src.dir_count); if (A.src.perm) key = A.src.perm[key]; }
                    const unsigned grp = __match_any_sync(am, key);
                    if ((grp & lt) == 0u) atomicAdd(&A.accum[key], (uint32_t)__popc(grp));
                }
                if (idle) have_ray = false;
            }
            if (MODE == 6) {
                // the rays retiring together reserve their stash records with one atomicAdd (prefix sum over the lanes)
                const bool put = idle && have_ray && !overflow && cnt > 0;
                const int mine = put ? cnt : 0;
                int incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += y; }
                const int total = __shfl_sync(FULL, incl, 31);
                if (total) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(A.st_count, (unsigned long long)total);
                    base = __shfl_sync(FULL, base, 0) + (unsigned long long)(incl - mine);
                    if (put) {
                        A.st_base[ray_i] = (long long)base;
                        if (base + (unsigned long long)mine <= A.st_cap)        // else: the host sees st_count > st_cap and repeats with room
                            for (int q = 0; q < cnt; ++q) {
                                A.st_t[base + q] = TSET(q); A.st_geom[base + q] = GSET(q); A.st_prim[base + q] = pset_l[q];
                                A.st_uv[base + q] = make_float2(uset_l[q], vset_l[q]);
                            }
                    }
                }
            }
            if (idle && have_ray) {
                have_ray = false;
                if (MODE == 0) {
                    if (A.out.t_hit) A.out.t_hit[ray_i] = best_t;
                    if (A.out.geom) A.out.geom[ray_i] = best_geom;
                    if (A.out.prim) A.out.prim[ray_i] = best_prim;
                    if (A.out.uv || A.out.nrm) {
                        float u = 0.0f, v = 0.0f, nx = 0.0f, ny = 0.0f, nz = 0.0f;
                        if (best_prim != QSMRT_INVALID) {
                            float4 p0, p1, p2;
                            load_tri(tris, best_tri, p0, p1, p2);
                            MtHit h;
                            mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h);
                            u = __fdiv_rn(h.U, h.absDen); v = __fdiv_rn(h.V, h.absDen);
                            float inv = __fdiv_rn(1.0f, __fsqrt_rn(f3dot(h.Ng, h.Ng)));
                            nx = __fmul_rn(h.Ng.x, inv); ny = __fmul_rn(h.Ng.y, inv); nz = __fmul_rn(h.Ng.z, inv);
                        }
                        if (A.out.uv) A.out.uv[ray_i] = make_float2(u, v);
                        if (A.out.nrm) { A.out.nrm[3 * ray_i] = nx; A.out.nrm[3 * ray_i + 1] = ny; A.out.nrm[3 * ray_i + 2] = nz; }
                    }
                } else if (MODE == 1) {
                    A.occluded[ray_i] = best_prim != QSMRT_INVALID ? 1 : 0;
                } else if (MODE == 2 || MODE == 6) {
                    A.counts[ray_i] = overflow ? -1 : cnt;           // -1: k_count_fix recounts this ray exactly
                } else if (MODE == 5) {
                    if (best_prim != QSMRT_INVALID) A.hitflag[A.order[best_tri]] = 1;
                }
            }
            if (exhausted) {
                if (im == FULL) break;
            } else {
                const int need = __popc(im);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(A.cursor, (unsigned long long)need);
                base = __shfl_sync(FULL, base, 0);
                exhausted = base + (unsigned long long)need >= A.nslots;
                if (idle) {
                    const uint64_t slot = base + __popc(im & lt);
                    uint64_t i = 0, gy = 0;
                    uint32_t gx = 0;
                    bool ok = slot < A.nslots;
                    if (A.src.kind == 3) {          // sweep: slot -> (grid, slot inside the grid)
                        const uint64_t a = fast_div(slot, A.src.per_grid_slots);
                        ok = ok && ray_index_of_slot(slot - a * A.src.per_grid_slots, A.src.per_grid_rays, A.row_len, i, gx, gy, A.row_major != 0);
                        i += a * A.src.per_grid_rays;
                    } else ok = ok && ray_index_of_slot(slot, A.N, A.row_len, i, gx, gy, A.row_major != 0);
                    if (ok) {
                        r = source_ray(A.src, i, gx, gy, A.row_len != 0);
                        if (QUANT) {
                            const float ax = A.sc.cell[0] * r.idx, ay = A.sc.cell[1] * r.idy, az = A.sc.cell[2] * r.idz;
                            r.oodx = fmaf(8388608.0f, ax, -((A.sc.glo[0] - r.O.x) * r.idx));
                            r.oody = fmaf(8388608.0f, ay, -((A.sc.glo[1] - r.O.y) * r.idy));
                            r.oodz = fmaf(8388608.0f, az, -((A.sc.glo[2] - r.O.z) * r.idz));
                            r.idx = ax; r.idy = ay; r.idz = az;
                            // PRMT selectors that decode the near / far plane of each axis straight from the packed word
                            // (sign of 1/d, not of d: a -0.0 component has a negative reciprocal)
                            snx = ax < 0.0f ? 0x7632u : 0x7610u; sny = ay < 0.0f ? 0x7632u : 0x7610u;
                            snz = az < 0.0f ? 0x7632u : 0x7610u;
                        }
                        ray_i = i; have_ray = true;
                        best_t = ANYHIT ? A.tfar : INFINITY;
                        cnt = 0; overflow = false;
                        best_geom = QSMRT_INVALID; best_prim = QSMRT_INVALID;
                        STACK_RESET();
                        cur = A.sc.ntris ? 0 : TR_SENTINEL;
                        lcur = 0;
                    }
                }
                continue;
            }
        }
        // ---- node phase
        for (;;) {
            const bool inner = (unsigned)cur < (unsigned)TR_SENTINEL;
            const bool parked = lcur != 0;
            const unsigned mw = __ballot_sync(FULL, inner && !parked);
            if (mw == 0u) break;
            if (__popc(mw) < A.want && __popc(__ballot_sync(FULL, parked)) >= A.tri_min) break;
            if (COUNTERS) {
                // lane census of the node phase (stats[2..6]): who sits out and why
                const unsigned m_in = __ballot_sync(FULL, inner);
                const unsigned m_idle = __ballot_sync(FULL, cur == TR_SENTINEL && !parked);
                const unsigned m_leaf2 = __ballot_sync(FULL, parked && cur < 0);
                const unsigned m_done = __ballot_sync(FULL, parked && cur == TR_SENTINEL);
                if (lane == 0) {
                    atomicAdd(&A.stats[2], 1ull); atomicAdd(&A.stats[3], (unsigned long long)__popc(m_in));
                    atomicAdd(&A.stats[4], (unsigned long long)__popc(m_idle));
                    atomicAdd(&A.stats[5], (unsigned long long)__popc(m_leaf2));
                    atomicAdd(&A.stats[6], (unsigned long long)__popc(m_done));
                }
            }
            // NodeSteps<MODE> node steps per vote: the vote only decides the phase, so skipping every other one
            // just delays a phase change by one step and saves the loop-control instructions all 32 lanes execute
#pragma unroll
            for (int rep = 0; rep < NodeSteps<MODE>::value; ++rep) {
            const bool in_ = (unsigned)cur < (unsigned)TR_SENTINEL;
            const bool pk_ = lcur != 0;
            if (in_) {
                int c0, c1;
                float t0, t1;
                bool h0, h1;
                if (QUANT) {
                    // the ray's direction signs pick each axis' entry / exit plane while decoding (PRMT with a
                    // per-lane selector), so the slab test needs no per-axis min/max: 8 FMNMX instead of 24
                    uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
                    ld256u(A.sc.qnodes + cur, w0, w1, w2, w3, w4, w5, w6, w7);
                    const uint32_t sfx = snx ^ 0x0022u, sfy = sny ^ 0x0022u, sfz = snz ^ 0x0022u;
#define QPL(W, S) prmt_f((W), (S))
                    float nx = fmaf(QPL(w0, snx), r.idx, -r.oodx), fx = fmaf(QPL(w0, sfx), r.idx, -r.oodx);
                    float ny = fmaf(QPL(w1, sny), r.idy, -r.oody), fy = fmaf(QPL(w1, sfy), r.idy, -r.oody);
                    float nz = fmaf(QPL(w2, snz), r.idz, -r.oodz), fz = fmaf(QPL(w2, sfz), r.idz, -r.oodz);
                    t0 = fmaxf(fmax3(nx, ny, nz) * SLAB_NEAR, 0.0f);
                    h0 = t0 <= fminf(fmin3(fx, fy, fz) * SLAB_FAR, best_t);
                    nx = fmaf(QPL(w3, snx), r.idx, -r.oodx); fx = fmaf(QPL(w3, sfx), r.idx, -r.oodx);
                    ny = fmaf(QPL(w4, sny), r.idy, -r.oody); fy = fmaf(QPL(w4, sfy), r.idy, -r.oody);
                    nz = fmaf(QPL(w5, snz), r.idz, -r.oodz); fz = fmaf(QPL(w5, sfz), r.idz, -r.oodz);
                    t1 = fmaxf(fmax3(nx, ny, nz) * SLAB_NEAR, 0.0f);
                    h1 = t1 <= fminf(fmin3(fx, fy, fz) * SLAB_FAR, best_t);
#undef QPL
                    c0 = (int)w6; c1 = (int)w7;
                } else {
                    float4 a, b, c, dd;
                    if (A.node_path == 0) {
                        ld256f(nodes + cur, a, b);
                        ld256f(reinterpret_cast<const char *>(nodes + cur) + 32, c, dd);
                    } else if (A.node_path == 1) {         // experiment: all four 16-byte chunks through the TEX path
                        a = tex1Dfetch<float4>(A.sc.node_tex, cur * 4);     b = tex1Dfetch<float4>(A.sc.node_tex, cur * 4 + 1);
                        c = tex1Dfetch<float4>(A.sc.node_tex, cur * 4 + 2); dd = tex1Dfetch<float4>(A.sc.node_tex, cur * 4 + 3);
                    } else {                               // experiment: half through TEX, half through LSU
                        a = tex1Dfetch<float4>(A.sc.node_tex, cur * 4);     b = tex1Dfetch<float4>(A.sc.node_tex, cur * 4 + 1);
                        ld256f(reinterpret_cast<const char *>(nodes + cur) + 32, c, dd);
                    }
                    c0 = __float_as_int(dd.x); c1 = __float_as_int(dd.y);
                    h0 = slab_fma(a.x, a.y, a.z, a.w, c.x, c.y, r, best_t, t0);
                    h1 = slab_fma(b.x, b.y, b.z, b.w, c.z, c.w, r, best_t, t1);
                }
                if (COUNTERS) ++n_node;
                // take child 1 first when child 0 is missed, or both are hit and 1 is nearer
                const bool take1 = !h0 || (CLOSEST && h1 && (t1 < t0));
                int nxt = take1 ? c1 : c0;
                const int other = c0 ^ c1 ^ nxt;        // the child not taken (one LOP3)
                if (h0 && h1) STACK_PUSH(other);
                if (!h0 && !h1) STACK_POP(nxt);
                cur = nxt;
                if (cur < 0 && !pk_) PARK_LEAF5();
            }
            }
        }
        // ---- triangle phase
        for (;;) {
            const bool has = lcur != 0;
            const unsigned mh = __ballot_sync(FULL, has);
            if (mh == 0u) break;
            if (__popc(mh) < A.tri_min && __any_sync(FULL, (unsigned)cur < (unsigned)TR_SENTINEL && !has)) break;
            if (COUNTERS && lane == 0) { atomicAdd(&A.stats[7], 1ull); atomicAdd(&A.stats[8], (unsigned long long)__popc(mh)); }
#pragma unroll
            for (int rep = 0; rep < TR_TRI_STEPS; ++rep)        // most leaves hold two triangles: one vote per leaf
            if (lcur != 0) {
                const uint32_t tri_i = (uint32_t)~lcur >> 2;
                float4 p0, p1, p2;
                load_tri(tris, tri_i, p0, p1, p2);
                MtHit h;
                if (COUNTERS) ++n_tri;
                if (CLOSEST) {
                    if ((MODE != 5 || A.alive[A.order[tri_i]]) && mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                        float tt = __fdiv_rn(h.T, h.absDen);
                        uint32_t pg = __float_as_uint(p1.w), pp = __float_as_uint(p0.w);
                        bool better = (tt < best_t) | ((tt == best_t) & ((pg < best_geom) | ((pg == best_geom) & (pp < best_prim))));
                        if (better) { best_t = tt; best_geom = pg; best_prim = pp; best_tri = tri_i; }
                    }
                    NEXT_TRI5();
                } else if (ANYHIT) {
                    if (mt_test(p0, p1, p2, r.O, r.D, A.tnear, A.tfar, h)) {
                        best_prim = 0u; cur = TR_SENTINEL; lcur = 0;              // occluded: drop the rest
                    } else {
                        NEXT_TRI5();
                    }
                } else {
                    if (mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                        const float tt = __fdiv_rn(h.T, h.absDen);
                        const uint32_t pg = __float_as_uint(p1.w);
                        int at = -1;
                        for (int q = 0; q < cnt; ++q)
                            if (TSET(q) == tt && (!A.multi_geom || GSET(q) == pg)) at = q;
                        if (at < 0) {
                            if (cnt < A.set_cap) {
                                TSET(cnt) = tt; if (A.multi_geom || MODE == 6) GSET(cnt) = pg;
                                if (MODE == 6) pset_l[cnt] = QSMRT_INVALID;                  // so the first record always wins below
                                at = cnt; ++cnt;
                            } else { overflow = true; cur = TR_SENTINEL; lcur = 0; }          // the fix-up kernel recounts this ray
                        }
                        if (MODE == 6 && at >= 0) {
                            const uint32_t pp = __float_as_uint(p0.w);
                            if (pp < pset_l[at]) {                      // survivor of equal (geometry, t): the lowest primitive id
                                pset_l[at] = pp; uset_l[at] = __fdiv_rn(h.U, h.absDen); vset_l[at] = __fdiv_rn(h.V, h.absDen);
                            }
                        }
                    }
                    if (!overflow) {
                        NEXT_TRI5();
                    }
                }
            }
        }
    }
    if (COUNTERS) {
        for (int o = 16; o > 0; o >>= 1) { n_node += __shfl_xor_sync(FULL, n_node, o); n_tri += __shfl_xor_sync(FULL, n_tri, o); }
        if (lane == 0) { atomicAdd(&A.stats[0], (unsigned long long)n_node); atomicAdd(&A.stats[1], (unsigned long long)n_tri); }
    }
#undef PARK_LEAF5
#undef NEXT_TRI5
#undef TSET
#undef GSET
#undef STACK_PUSH
#undef STACK_POP
#undef STACK_RESET
}

// materialise the hemisphere rays of SRC 2 (tests and small batches): rays[n_points * n_dirs][6]
__global__ void __launch_bounds__(256)
k_gen_hemisphere(float *__restrict__ rays, RaySource S, uint64_t n)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r = source_ray(S, i);
    float2 *p = reinterpret_cast<float2 *>(rays + 6 * i);
    p[0] = make_float2(r.O.x, r.O.y); p[1] = make_float2(r.O.z, r.D.x); p[2] = make_float2(r.D.y, r.D.z);
}
