// trace_sched.cuh -- v4 traversal: persistent warps driven by a one-step
// scheduler.  Included inside traverse.cu's anonymous namespace (uses its
// helpers).  sm_100a.
//
// Every iteration the warp votes once on what its lanes need -- a node step
// (slab tests of one 64-byte node), a triangle step (one Moeller-Trumbore
// test), or a retire/refill -- and executes exactly one kind, predicated per
// lane.  Policy: node steps while at least `want_thresh` lanes are still
// looking for their next leaf, otherwise triangle steps while any lane holds
// triangles; idle lanes are refilled from the global cursor once
// `refill_thresh` of them are idle (ballot/popc compaction).  A lane that
// holds a parked leaf may keep descending speculatively (`spec`), which costs
// no issue slots (the warp runs the node step anyway) but does cost L1/L2
// bandwidth -- a tunable, measured in profiles/.
//
// Nodes are fetched with two 256-bit loads (LDG.E.256, new on sm_100),
// triangles (48-byte records, 16-byte aligned) with three 128-bit loads.

struct __align__(32) V8 { float4 lo, hi; };

__device__ __forceinline__ void ld256(const void *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

struct Tuning { int refill_thresh, want_thresh, spec; };

template <int MODE, bool COUNTERS>
__global__ void __launch_bounds__(TR_BLOCK)
k_trace_sched(SceneView sc, const float *__restrict__ rays, uint64_t N, uint32_t row_len, uint64_t nslots,
              CastOut out, uint8_t *__restrict__ occluded, float tnear, float tfar_in,
              unsigned long long *__restrict__ cursor, Tuning tune, unsigned long long *__restrict__ stats)
{
    __shared__ int sstack[TR_SSTACK * TR_BLOCK];
    int *const sbase = sstack + threadIdx.x;
    int loc[TR_LSTACK];
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;

    Ray r;
    float best_t = 0.0f; uint32_t best_geom = QSMRT_INVALID, best_prim = QSMRT_INVALID, best_tri = 0u;
    uint64_t ray_i = 0;
    bool have_ray = false, exhausted = false;
    int cur = TR_SENTINEL, sp = 0;
    uint32_t tri_i = 0, tri_end = 0;
    unsigned n_node = 0, n_tri = 0;          // COUNTERS only

#define PUSH(v) do { if (sp < TR_SSTACK) sbase[sp * TR_BLOCK] = (v); else loc[sp - TR_SSTACK] = (v); ++sp; } while (0)
#define POP(dst) do { --sp; (dst) = sp < TR_SSTACK ? sbase[sp * TR_BLOCK] : loc[sp - TR_SSTACK]; } while (0)
#define PARK_LEAF() do { uint32_t ref_ = (uint32_t)~cur; tri_i = ref_ >> 2; tri_end = tri_i + (ref_ & 3u) + 1u; POP(cur); } while (0)

    for (;;) {
        const bool inner = (unsigned)cur < (unsigned)TR_SENTINEL;
        const bool has = tri_i < tri_end;
        const unsigned m_inner = __ballot_sync(FULL, inner);
        const unsigned m_has = __ballot_sync(FULL, has);
        const unsigned m_idle = ~(m_inner | m_has);           // a lane is never at a leaf without holding it
        if (m_idle == FULL || (!exhausted && __popc(m_idle) >= tune.refill_thresh)) {
            // ---- retire finished rays, refill idle lanes
            const bool idle = (m_idle >> lane) & 1u;
            if (idle && have_ray) {
                have_ray = false;
                if (MODE == 0) {
                    if (out.t_hit) out.t_hit[ray_i] = best_t;
                    if (out.geom) out.geom[ray_i] = best_geom;
                    if (out.prim) out.prim[ray_i] = best_prim;
                    if (out.uv || out.nrm) {
                        float u = 0.0f, v = 0.0f, nx = 0.0f, ny = 0.0f, nz = 0.0f;
                        if (best_prim != QSMRT_INVALID) {
                            float4 p0, p1, p2;
                            load_tri(sc.tris, best_tri, p0, p1, p2);
                            MtHit h;
                            mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h);
                            u = __fdiv_rn(h.U, h.absDen); v = __fdiv_rn(h.V, h.absDen);
                            float inv = __fdiv_rn(1.0f, __fsqrt_rn(f3dot(h.Ng, h.Ng)));
                            nx = __fmul_rn(h.Ng.x, inv); ny = __fmul_rn(h.Ng.y, inv); nz = __fmul_rn(h.Ng.z, inv);
                        }
                        if (out.uv) out.uv[ray_i] = make_float2(u, v);
                        if (out.nrm) { out.nrm[3 * ray_i] = nx; out.nrm[3 * ray_i + 1] = ny; out.nrm[3 * ray_i + 2] = nz; }
                    }
                } else {
                    occluded[ray_i] = best_prim != QSMRT_INVALID ? 1 : 0;
                }
            }
            if (exhausted) {
                if (m_idle == FULL) break;
                // fall through to traversal for the lanes still working
            } else {
                const int need = __popc(m_idle);
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cursor, (unsigned long long)need);
                base = __shfl_sync(FULL, base, 0);
                exhausted = base + (unsigned long long)need >= nslots;
                if (idle) {
                    const uint64_t slot = base + __popc(m_idle & lt);
                    uint64_t i;
                    if (slot < nslots && ray_index_of_slot(slot, N, row_len, i)) {
                        r = load_ray(rays, i);
                        ray_i = i; have_ray = true;
                        best_t = MODE == 0 ? INFINITY : tfar_in;
                        best_geom = QSMRT_INVALID; best_prim = QSMRT_INVALID;
                        sbase[0] = TR_SENTINEL; sp = 1;
                        cur = sc.ntris ? 0 : TR_SENTINEL;
                        tri_i = tri_end = 0;
                    }
                }
                continue;
            }
        }
        const unsigned m_want = m_inner & ~m_has;               // lanes still looking for a leaf
        if (__popc(m_want) >= tune.want_thresh || m_has == 0u) {
            // ---- node step
            if (inner && (tune.spec || !has)) {
                float4 a, b, c, dd;
                ld256(sc.nodes + cur, a, b);
                ld256(reinterpret_cast<const char *>(sc.nodes + cur) + 32, c, dd);
                const int c0 = __float_as_int(dd.x), c1 = __float_as_int(dd.y);
                float t0, t1;
                bool h0 = slab_fma(a.x, a.y, a.z, a.w, c.x, c.y, r, best_t, t0);
                bool h1 = slab_fma(b.x, b.y, b.z, b.w, c.z, c.w, r, best_t, t1);
                if (COUNTERS) ++n_node;
                if (!(h0 | h1)) {
                    POP(cur);
                } else {
                    cur = h0 ? c0 : c1;
                    if (h0 & h1) {
                        int far = c1;
                        if (MODE == 0 && t1 < t0) { far = cur; cur = c1; }
                        PUSH(far);
                    }
                }
                if (cur < 0 && !has) PARK_LEAF();
            }
        } else {
            // ---- triangle step
            if (has) {
                float4 p0, p1, p2;
                load_tri(sc.tris, tri_i, p0, p1, p2);       // 48-byte records are only 16-byte aligned
                MtHit h;
                if (COUNTERS) ++n_tri;
                if (MODE == 0) {
                    if (mt_test(p0, p1, p2, r.O, r.D, 0.0f, INFINITY, h)) {
                        float tt = __fdiv_rn(h.T, h.absDen);
                        uint32_t pg = __float_as_uint(p1.w), pp = __float_as_uint(p0.w);
                        bool better = (tt < best_t) | ((tt == best_t) & ((pg < best_geom) | ((pg == best_geom) & (pp < best_prim))));
                        if (better) { best_t = tt; best_geom = pg; best_prim = pp; best_tri = tri_i; }
                    }
                    ++tri_i;
                    if (tri_i == tri_end && cur < 0) PARK_LEAF();
                } else {
                    if (mt_test(p0, p1, p2, r.O, r.D, tnear, tfar_in, h)) {
                        best_prim = 0u; cur = TR_SENTINEL; tri_i = tri_end;       // occluded: drop the rest
                    } else {
                        ++tri_i;
                        if (tri_i == tri_end && cur < 0) PARK_LEAF();
                    }
                }
            }
        }
    }
    if (COUNTERS) {
        for (int o = 16; o > 0; o >>= 1) { n_node += __shfl_xor_sync(FULL, n_node, o); n_tri += __shfl_xor_sync(FULL, n_tri, o); }
        if (lane == 0) { atomicAdd(&stats[0], (unsigned long long)n_node); atomicAdd(&stats[1], (unsigned long long)n_tri); }
    }
#undef PUSH
#undef POP
#undef PARK_LEAF
}
