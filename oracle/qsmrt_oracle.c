/*
 * qsmrt_oracle.c -- CPU restatement of Open3D/Embree RaycastingScene semantics.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (pyqsm_b200 + libqsmrt.so) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (wischmcj/pyQSM) delegates all ray/mesh
 * arithmetic to open3d.t.geometry.RaycastingScene (pyQSM/viz/ray_casting.py:8,
 * call sites :65-69 :155-169 :218-231 :241-255 :275-279 :316-319).  Open3D
 * (>=0.18, pyproject.toml:38) and Embree are third-party, absent from
 * /root/reference, not installable here, and the reference has no tests or
 * golden vectors.  What follows restates the published behaviour of
 *   Open3D  cpp/open3d/t/geometry/RaycastingScene.cpp   (CastRays,
 *           CountIntersections, ListIntersections, TestOcclusions)
 *   Embree  kernels/geometry/triangle_intersector_moeller.h
 *           (MoellerTrumboreIntersector1, non-robust, no culling)
 * as recorded in SURVEY.md section 8c, and is anchored on the analytic
 * known-answer vectors in tests/golden/.
 *
 * Arithmetic contract (shared bit-for-bit with the CUDA kernels):
 *   all fp32; cross(a,b) = (a.y*b.z - a.z*b.y, ...) as fmaf(a.y,b.z,-(a.z*b.y));
 *   dot(a,b) = fmaf(a.x,b.x, fmaf(a.y,b.y, a.z*b.z));  IEEE divide and sqrt;
 *   compiled with -ffp-contract=off so nothing else is fused.
 *
 * Two search modes produce identical answers:
 *   mode 0  brute force over every triangle (ground truth),
 *   mode 1  the canonical LBVH (63-bit Morton, Karras topology, one triangle
 *           per leaf, 32-byte nodes) traversed near-child-first.  Its fetch
 *           counters define the roofline bytes per ray.
 *   mode 2  a top-down binned-SAH BVH over the same triangles (one per leaf,
 *           same node format, same traversal).  Measurement only: its fetch
 *           counters next to mode 1's say how much tree quality the LBVH
 *           leaves on the table (baseline/canonical_counters.json).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_INVALID_ID 0xFFFFFFFFu

typedef struct { float x, y, z; } v3;

/* 48-byte triangle record: v0, e1 = v0-v1, e2 = v2-v0 (Embree TriangleM), ids */
typedef struct {
    v3 v0; uint32_t prim;
    v3 e1; uint32_t geom;
    v3 e2; uint32_t pad;
} orc_tri;

/* 32-byte canonical node: AABB + two child references.
 * child >= 0: internal node index; child < 0: leaf holding sorted triangle ~child */
typedef struct {
    float lo[3], hi[3];
    int32_t left, right;
} orc_node;

typedef struct {
    /* geometry as added */
    float    *verts;   uint64_t nverts;       /* concatenated */
    uint32_t *idx;     uint64_t ntris;        /* vertex ids rebased into verts */
    uint32_t *tri_geom, *tri_prim;
    uint32_t ngeoms;
    /* committed */
    int       committed;
    orc_tri  *tris;                           /* in input order (brute force) */
    orc_tri  *stris;                          /* Morton-sorted order (BVH) */
    uint64_t *keys;                           /* sorted Morton keys */
    uint32_t *order;                          /* sorted position -> input triangle */
    orc_node *nodes;                          /* ntris-1 internal nodes, root = 0 */
    float    *leaf_lo, *leaf_hi;              /* padded leaf boxes, sorted order */
    int32_t  *parent;                         /* parent of internal node */
    float     slo[3], shi[3];                 /* scene bounds (unpadded) */
    float     pad;                            /* box padding */
    /* mode 2: binned-SAH tree, built on first use */
    orc_tri  *sah_tris; orc_node *sah_nodes; float *sah_leaf_lo, *sah_leaf_hi;
} orc_scene;

/* the arrays a traversal reads: the canonical LBVH (mode 1) or the SAH tree (mode 2) */
typedef struct { const orc_node *nodes; const float *leaf_lo, *leaf_hi; const orc_tri *stris; } orc_tree;
static void build_sah(orc_scene *s);
static inline orc_tree tree_of(const orc_scene *s, int mode)
{
    orc_tree t;
    if (mode == 2) { t.nodes = s->sah_nodes; t.leaf_lo = s->sah_leaf_lo; t.leaf_hi = s->sah_leaf_hi; t.stris = s->sah_tris; }
    else { t.nodes = s->nodes; t.leaf_lo = s->leaf_lo; t.leaf_hi = s->leaf_hi; t.stris = s->stris; }
    return t;
}

/* ------------------------------------------------------------------ math */
static inline v3 v3sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline v3 v3cross(v3 a, v3 b) {
    v3 r;
    r.x = fmaf(a.y, b.z, -(a.z * b.y));
    r.y = fmaf(a.z, b.x, -(a.x * b.z));
    r.z = fmaf(a.x, b.y, -(a.y * b.x));
    return r;
}
static inline float v3dot(v3 a, v3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
static inline float fxor(float f, uint32_t m) {
    uint32_t u; memcpy(&u, &f, 4); u ^= m; memcpy(&f, &u, 4); return f;
}
static inline uint32_t fsignmask(float f) { uint32_t u; memcpy(&u, &f, 4); return u & 0x80000000u; }

typedef struct { float U, V, T, absDen; } mt_hit;

/* Embree MoellerTrumboreIntersector1 (triangle_intersector_moeller.h), edges
 * inclusive, no culling.  Range test: absDen*tnear < T <= absDen*tfar. */
static inline int mt_test(const orc_tri *tr, v3 O, v3 D, float tnear, float tfar, mt_hit *h)
{
    v3 Ng = v3cross(tr->e2, tr->e1);
    v3 C = v3sub(tr->v0, O);
    v3 R = v3cross(C, D);
    float den = v3dot(Ng, D);
    float absDen = fabsf(den);
    uint32_t sgn = fsignmask(den);
    float U = fxor(v3dot(R, tr->e2), sgn);
    float V = fxor(v3dot(R, tr->e1), sgn);
    if (!(den != 0.0f)) return 0;
    if (!(U >= 0.0f)) return 0;
    if (!(V >= 0.0f)) return 0;
    if (!(U + V <= absDen)) return 0;
    float T = fxor(v3dot(Ng, C), sgn);
    if (!(absDen * tnear < T)) return 0;
    if (!(T <= absDen * tfar)) return 0;
    h->U = U; h->V = V; h->T = T; h->absDen = absDen;
    return 1;
}

/* near-edge predicate: plane hit in range and a barycentric within eps of 0,
 * evaluated whether or not the triangle test itself accepted. */
static inline int mt_near_edge(const orc_tri *tr, v3 O, v3 D, float eps)
{
    v3 Ng = v3cross(tr->e2, tr->e1);
    v3 C = v3sub(tr->v0, O);
    v3 R = v3cross(C, D);
    float den = v3dot(Ng, D);
    float absDen = fabsf(den);
    uint32_t sgn = fsignmask(den);
    if (!(den != 0.0f)) return 0;
    float T = fxor(v3dot(Ng, C), sgn);
    if (!(T > 0.0f)) return 0;
    float U = fxor(v3dot(R, tr->e2), sgn);
    float V = fxor(v3dot(R, tr->e1), sgn);
    float W = absDen - U - V;
    float band = eps * absDen;
    float m = fminf(U, fminf(V, W));
    /* inside-or-outside by at most the band, and not far outside on another edge */
    return (m >= -band && m <= band);
}

/* ------------------------------------------------------------- lifecycle */
orc_scene *orc_scene_create(void) { return (orc_scene *)calloc(1, sizeof(orc_scene)); }

static void free_committed(orc_scene *s)
{
    free(s->tris); free(s->stris); free(s->keys); free(s->order); free(s->nodes);
    free(s->leaf_lo); free(s->leaf_hi); free(s->parent);
    free(s->sah_tris); free(s->sah_nodes); free(s->sah_leaf_lo); free(s->sah_leaf_hi);
    s->sah_tris = NULL; s->sah_nodes = NULL; s->sah_leaf_lo = s->sah_leaf_hi = NULL;
    s->tris = s->stris = NULL; s->keys = NULL; s->order = NULL; s->nodes = NULL;
    s->leaf_lo = s->leaf_hi = NULL; s->parent = NULL; s->committed = 0;
}

void orc_scene_destroy(orc_scene *s)
{
    if (!s) return;
    free_committed(s);
    free(s->verts); free(s->idx); free(s->tri_geom); free(s->tri_prim);
    free(s);
}

/* Open3D RaycastingScene::AddTriangles: copies the mesh, returns the geometry
 * id (0, 1, ...).  Returns ORC_INVALID_ID on an out-of-range vertex index. */
uint32_t orc_add_triangles(orc_scene *s, const float *v, uint64_t V, const uint32_t *idx, uint64_t T)
{
    for (uint64_t i = 0; i < 3 * T; ++i) if (idx[i] >= V) return ORC_INVALID_ID;
    free_committed(s);
    s->verts = (float *)realloc(s->verts, sizeof(float) * 3 * (s->nverts + V + 1));
    memcpy(s->verts + 3 * s->nverts, v, sizeof(float) * 3 * V);
    s->idx = (uint32_t *)realloc(s->idx, sizeof(uint32_t) * 3 * (s->ntris + T + 1));
    s->tri_geom = (uint32_t *)realloc(s->tri_geom, sizeof(uint32_t) * (s->ntris + T + 1));
    s->tri_prim = (uint32_t *)realloc(s->tri_prim, sizeof(uint32_t) * (s->ntris + T + 1));
    for (uint64_t t = 0; t < T; ++t) {
        for (int k = 0; k < 3; ++k) s->idx[3 * (s->ntris + t) + k] = idx[3 * t + k] + (uint32_t)s->nverts;
        s->tri_geom[s->ntris + t] = s->ngeoms;
        s->tri_prim[s->ntris + t] = (uint32_t)t;
    }
    s->nverts += V; s->ntris += T;
    return s->ngeoms++;
}

/* --------------------------------------------------------------- LBVH */
static inline uint64_t spread21(uint32_t x)
{
    uint64_t v = x & 0x1FFFFFu;
    v = (v | (v << 32)) & 0x001F00000000FFFFull;
    v = (v | (v << 16)) & 0x001F0000FF0000FFull;
    v = (v | (v << 8))  & 0x100F00F00F00F00Full;
    v = (v | (v << 4))  & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2))  & 0x1249249249249249ull;
    return v;
}

static inline uint32_t quant21(float c, float lo, float scale)
{
    float q = (c - lo) * scale;
    q = fminf(fmaxf(q, 0.0f), 2097151.0f);
    return (uint32_t)q;
}

static inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
static inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }

/* Karras 2012 delta: common-prefix length of keys i and j, index-augmented on ties */
static inline int delta(const uint64_t *k, int64_t n, int64_t i, int64_t j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t a = k[i], b = k[j];
    if (a != b) return clz64(a ^ b);
    return 64 + clz32((uint32_t)i ^ (uint32_t)j);
}

static void radix_sort_pairs(uint64_t *keys, uint32_t *vals, uint64_t n)
{
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (n + 1));
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * (n + 1));
    for (int pass = 0; pass < 8; ++pass) {
        uint64_t hist[257]; memset(hist, 0, sizeof(hist));
        int sh = pass * 8;
        for (uint64_t i = 0; i < n; ++i) hist[((keys[i] >> sh) & 0xFF) + 1]++;
        for (int d = 0; d < 256; ++d) hist[d + 1] += hist[d];
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t p = hist[(keys[i] >> sh) & 0xFF]++;
            k2[p] = keys[i]; v2[p] = vals[i];
        }
        uint64_t *tk = keys; keys = k2; k2 = tk;
        uint32_t *tv = vals; vals = v2; v2 = tv;
    }
    /* 8 passes: result is back in the caller's arrays */
    free(k2); free(v2);
}

static inline void tri_bounds(const orc_scene *s, uint64_t t, float lo[3], float hi[3])
{
    for (int a = 0; a < 3; ++a) {
        float p0 = s->verts[3 * s->idx[3 * t + 0] + a];
        float p1 = s->verts[3 * s->idx[3 * t + 1] + a];
        float p2 = s->verts[3 * s->idx[3 * t + 2] + a];
        lo[a] = fminf(p0, fminf(p1, p2));
        hi[a] = fmaxf(p0, fmaxf(p1, p2));
    }
}

int orc_commit(orc_scene *s)
{
    if (s->committed) return 0;
    free_committed(s);
    uint64_t n = s->ntris;
    s->committed = 1;
    if (n == 0) return 0;
    s->tris  = (orc_tri *)malloc(sizeof(orc_tri) * n);
    s->stris = (orc_tri *)malloc(sizeof(orc_tri) * n);
    s->keys  = (uint64_t *)malloc(sizeof(uint64_t) * n);
    s->order = (uint32_t *)malloc(sizeof(uint32_t) * n);
    s->leaf_lo = (float *)malloc(sizeof(float) * 3 * n);
    s->leaf_hi = (float *)malloc(sizeof(float) * 3 * n);

    float slo[3] = { INFINITY, INFINITY, INFINITY }, shi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (uint64_t t = 0; t < n; ++t) {
        const float *p0 = s->verts + 3 * s->idx[3 * t + 0];
        const float *p1 = s->verts + 3 * s->idx[3 * t + 1];
        const float *p2 = s->verts + 3 * s->idx[3 * t + 2];
        orc_tri *tr = &s->tris[t];
        tr->v0.x = p0[0]; tr->v0.y = p0[1]; tr->v0.z = p0[2];
        tr->e1.x = p0[0] - p1[0]; tr->e1.y = p0[1] - p1[1]; tr->e1.z = p0[2] - p1[2];
        tr->e2.x = p2[0] - p0[0]; tr->e2.y = p2[1] - p0[1]; tr->e2.z = p2[2] - p0[2];
        tr->prim = s->tri_prim[t]; tr->geom = s->tri_geom[t]; tr->pad = 0;
        float lo[3], hi[3]; tri_bounds(s, t, lo, hi);
        for (int a = 0; a < 3; ++a) { slo[a] = fminf(slo[a], lo[a]); shi[a] = fmaxf(shi[a], hi[a]); }
    }
    memcpy(s->slo, slo, sizeof(slo)); memcpy(s->shi, shi, sizeof(shi));
    /* Box padding: 2^-17 of the largest |coordinate| or extent; covers the
     * fp32 slack of the triangle test for origins within a few scene sizes. */
    float m = 0.0f;
    for (int a = 0; a < 3; ++a) {
        m = fmaxf(m, fabsf(slo[a])); m = fmaxf(m, fabsf(shi[a])); m = fmaxf(m, shi[a] - slo[a]);
    }
    s->pad = m * 7.62939453125e-06f;           /* 2^-17 */
    if (!(s->pad > 0.0f)) s->pad = 1e-30f;

    float scale[3];
    for (int a = 0; a < 3; ++a) {
        float ext = shi[a] - slo[a];
        scale[a] = ext > 0.0f ? 2097152.0f / ext : 0.0f;
    }
    for (uint64_t t = 0; t < n; ++t) {
        float lo[3], hi[3]; tri_bounds(s, t, lo, hi);
        uint32_t q[3];
        for (int a = 0; a < 3; ++a) q[a] = quant21((lo[a] + hi[a]) * 0.5f, slo[a], scale[a]);
        s->keys[t] = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
        s->order[t] = (uint32_t)t;
    }
    radix_sort_pairs(s->keys, s->order, n);
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t t = s->order[i];
        s->stris[i] = s->tris[t];
        float lo[3], hi[3]; tri_bounds(s, t, lo, hi);
        for (int a = 0; a < 3; ++a) { s->leaf_lo[3 * i + a] = lo[a] - s->pad; s->leaf_hi[3 * i + a] = hi[a] + s->pad; }
    }
    if (n == 1) return 0;

    int64_t ni = (int64_t)n - 1;
    s->nodes = (orc_node *)malloc(sizeof(orc_node) * ni);
    s->parent = (int32_t *)malloc(sizeof(int32_t) * ni);
    int32_t *leaf_parent = (int32_t *)malloc(sizeof(int32_t) * n);
    s->parent[0] = -1;
    const uint64_t *k = s->keys;
    for (int64_t i = 0; i < ni; ++i) {
        int d = (delta(k, n, i, i + 1) - delta(k, n, i, i - 1)) >= 0 ? 1 : -1;
        int dmin = delta(k, n, i, i - d);
        int64_t lmax = 2;
        while (delta(k, n, i, i + lmax * d) > dmin) lmax *= 2;
        int64_t l = 0;
        for (int64_t t = lmax / 2; t >= 1; t /= 2)
            if (delta(k, n, i, i + (l + t) * d) > dmin) l += t;
        int64_t j = i + l * d;
        int dnode = delta(k, n, i, j);
        int64_t sp = 0;
        for (int64_t t = (l + 1) / 2; ; t = (t + 1) / 2) {
            if (delta(k, n, i, i + (sp + t) * d) > dnode) sp += t;
            if (t <= 1) break;
        }
        int64_t gamma = i + sp * d + (d < 0 ? -1 : 0);
        int64_t first = i < j ? i : j, last = i < j ? j : i;
        orc_node *nd = &s->nodes[i];
        if (first == gamma) { nd->left = ~(int32_t)gamma; leaf_parent[gamma] = (int32_t)i; }
        else { nd->left = (int32_t)gamma; s->parent[gamma] = (int32_t)i; }
        if (last == gamma + 1) { nd->right = ~(int32_t)(gamma + 1); leaf_parent[gamma + 1] = (int32_t)i; }
        else { nd->right = (int32_t)(gamma + 1); s->parent[gamma + 1] = (int32_t)i; }
    }
    /* bottom-up refit: second arrival computes the parent box */
    uint8_t *seen = (uint8_t *)calloc(ni, 1);
    for (uint64_t i = 0; i < n; ++i) {
        int32_t p = leaf_parent[i];
        while (p >= 0) {
            if (!seen[p]) { seen[p] = 1; break; }
            orc_node *nd = &s->nodes[p];
            const float *llo, *lhi, *rlo, *rhi;
            if (nd->left < 0) { llo = s->leaf_lo + 3 * (~nd->left); lhi = s->leaf_hi + 3 * (~nd->left); }
            else { llo = s->nodes[nd->left].lo; lhi = s->nodes[nd->left].hi; }
            if (nd->right < 0) { rlo = s->leaf_lo + 3 * (~nd->right); rhi = s->leaf_hi + 3 * (~nd->right); }
            else { rlo = s->nodes[nd->right].lo; rhi = s->nodes[nd->right].hi; }
            for (int a = 0; a < 3; ++a) { nd->lo[a] = fminf(llo[a], rlo[a]); nd->hi[a] = fmaxf(lhi[a], rhi[a]); }
            p = s->parent[p];
        }
    }
    free(seen); free(leaf_parent);
    return 0;
}

/* builder introspection for the builder-parity tests */
uint64_t orc_num_triangles(const orc_scene *s) { return s->ntris; }
const uint64_t *orc_sorted_keys(const orc_scene *s) { return s->keys; }
const uint32_t *orc_sorted_order(const orc_scene *s) { return s->order; }
const orc_node *orc_nodes(const orc_scene *s) { return s->nodes; }
float orc_box_pad(const orc_scene *s) { return s->pad; }
void orc_scene_bounds(const orc_scene *s, float *lo, float *hi) { memcpy(lo, s->slo, 12); memcpy(hi, s->shi, 12); }

/* ------------------------------------------------ binned-SAH tree (mode 2)
 * Top-down, 32 centroid bins on the axis of the largest centroid extent, split
 * where leftArea*leftCount + rightArea*rightCount is smallest (median split
 * when all centroids coincide), down to one triangle per leaf.  Same padded
 * leaf boxes and node format as the canonical LBVH, so the same traversal and
 * the same counters apply: the difference in node visits is tree quality only. */
#define SAH_BINS 32
typedef struct { orc_scene *s; uint32_t *ids; float *clo, *chi, *cen; int32_t next_node; } sah_ctx;

static inline float half_area(const float *lo, const float *hi)
{
    float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

/* builds the subtree over ids[b, e); returns its child reference and writes its box */
static int32_t sah_rec(sah_ctx *c, int64_t b, int64_t e, float *olo, float *ohi)
{
    orc_scene *s = c->s;
    if (e - b == 1) {
        uint32_t t = c->ids[b];
        s->sah_tris[b] = s->tris[t];
        for (int a = 0; a < 3; ++a) { s->sah_leaf_lo[3 * b + a] = olo[a] = c->clo[3 * t + a]; s->sah_leaf_hi[3 * b + a] = ohi[a] = c->chi[3 * t + a]; }
        return ~(int32_t)b;
    }
    float cmin[3] = { INFINITY, INFINITY, INFINITY }, cmax[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int64_t i = b; i < e; ++i)
        for (int a = 0; a < 3; ++a) { float x = c->cen[3 * c->ids[i] + a]; cmin[a] = fminf(cmin[a], x); cmax[a] = fmaxf(cmax[a], x); }
    int ax = 0;
    for (int a = 1; a < 3; ++a) if (cmax[a] - cmin[a] > cmax[ax] - cmin[ax]) ax = a;
    int64_t mid = (b + e) / 2;
    float ext = cmax[ax] - cmin[ax];
    if (ext > 0.0f) {
        float blo[SAH_BINS][3], bhi[SAH_BINS][3]; int64_t bcnt[SAH_BINS];
        for (int k = 0; k < SAH_BINS; ++k) { bcnt[k] = 0; for (int a = 0; a < 3; ++a) { blo[k][a] = INFINITY; bhi[k][a] = -INFINITY; } }
        float scale = (float)SAH_BINS / ext;
        for (int64_t i = b; i < e; ++i) {
            uint32_t t = c->ids[i];
            int k = (int)((c->cen[3 * t + ax] - cmin[ax]) * scale); if (k >= SAH_BINS) k = SAH_BINS - 1; if (k < 0) k = 0;
            bcnt[k]++;
            for (int a = 0; a < 3; ++a) { blo[k][a] = fminf(blo[k][a], c->clo[3 * t + a]); bhi[k][a] = fmaxf(bhi[k][a], c->chi[3 * t + a]); }
        }
        float rarea[SAH_BINS]; int64_t rcnt[SAH_BINS];
        float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY }; int64_t cnt = 0;
        for (int k = SAH_BINS - 1; k > 0; --k) {
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], blo[k][a]); hi[a] = fmaxf(hi[a], bhi[k][a]); }
            cnt += bcnt[k]; rcnt[k] = cnt; rarea[k] = cnt ? half_area(lo, hi) : 0.0f;
        }
        for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        cnt = 0;
        float best = INFINITY; int bestk = -1;
        for (int k = 0; k < SAH_BINS - 1; ++k) {
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], blo[k][a]); hi[a] = fmaxf(hi[a], bhi[k][a]); }
            cnt += bcnt[k];
            if (cnt == 0 || rcnt[k + 1] == 0) continue;
            float cost = half_area(lo, hi) * (float)cnt + rarea[k + 1] * (float)rcnt[k + 1];
            if (cost < best) { best = cost; bestk = k; }
        }
        if (bestk >= 0) {       /* partition ids[b, e) by bin <= bestk */
            int64_t i = b, j = e - 1;
            while (i <= j) {
                uint32_t t = c->ids[i];
                int k = (int)((c->cen[3 * t + ax] - cmin[ax]) * scale); if (k >= SAH_BINS) k = SAH_BINS - 1; if (k < 0) k = 0;
                if (k <= bestk) ++i; else { c->ids[i] = c->ids[j]; c->ids[j] = t; --j; }
            }
            mid = i;
            if (mid == b || mid == e) mid = (b + e) / 2;
        }
    }
    int32_t me = c->next_node++;
    float llo[3], lhi[3], rlo[3], rhi[3];
    int32_t L = sah_rec(c, b, mid, llo, lhi), R = sah_rec(c, mid, e, rlo, rhi);
    orc_node *nd = &s->sah_nodes[me];
    nd->left = L; nd->right = R;
    for (int a = 0; a < 3; ++a) { nd->lo[a] = olo[a] = fminf(llo[a], rlo[a]); nd->hi[a] = ohi[a] = fmaxf(lhi[a], rhi[a]); }
    return me;
}

static void build_sah(orc_scene *s)
{
#pragma omp critical(orc_sah)
    if (!s->sah_nodes && s->ntris > 1) {
        uint64_t n = s->ntris;
        sah_ctx c; c.s = s; c.next_node = 0;
        c.ids = (uint32_t *)malloc(sizeof(uint32_t) * n);
        c.clo = (float *)malloc(sizeof(float) * 3 * n); c.chi = (float *)malloc(sizeof(float) * 3 * n); c.cen = (float *)malloc(sizeof(float) * 3 * n);
        s->sah_tris = (orc_tri *)malloc(sizeof(orc_tri) * n);
        s->sah_leaf_lo = (float *)malloc(sizeof(float) * 3 * n); s->sah_leaf_hi = (float *)malloc(sizeof(float) * 3 * n);
        orc_node *nodes = (orc_node *)malloc(sizeof(orc_node) * (n - 1));
        for (uint64_t t = 0; t < n; ++t) {
            float lo[3], hi[3]; tri_bounds(s, t, lo, hi);
            c.ids[t] = (uint32_t)t;
            for (int a = 0; a < 3; ++a) { c.clo[3 * t + a] = lo[a] - s->pad; c.chi[3 * t + a] = hi[a] + s->pad; c.cen[3 * t + a] = (lo[a] + hi[a]) * 0.5f; }
        }
        s->sah_nodes = nodes;
        float lo[3], hi[3];
        sah_rec(&c, 0, (int64_t)n, lo, hi);
        free(c.ids); free(c.clo); free(c.chi); free(c.cen);
    }
}

/* ------------------------------------------------------------ ray setup */
typedef struct { v3 O, D; float idx, idy, idz; } ray_t;

static inline float safe_inv(float d)
{
    /* |d| clamped away from zero so 0 * inv never produces NaN */
    const float tiny = 8.271806125530277e-25f;  /* 2^-80 */
    float a = fabsf(d) < tiny ? copysignf(tiny, d) : d;
    return 1.0f / a;
}

static inline ray_t load_ray(const float *r)
{
    ray_t y;
    y.O.x = r[0]; y.O.y = r[1]; y.O.z = r[2];
    y.D.x = r[3]; y.D.y = r[4]; y.D.z = r[5];
    y.idx = safe_inv(y.D.x); y.idy = safe_inv(y.D.y); y.idz = safe_inv(y.D.z);
    return y;
}

/* Conservative slab test: subtract first (sign exact), interval widened by a
 * few ulps.  Returns entry distance in *tn; hit iff entry <= min(exit, tmax). */
static inline int box_test(const float *lo, const float *hi, const ray_t *y, float tmax, float *tn)
{
    float x0 = (lo[0] - y->O.x) * y->idx, x1 = (hi[0] - y->O.x) * y->idx;
    float y0 = (lo[1] - y->O.y) * y->idy, y1 = (hi[1] - y->O.y) * y->idy;
    float z0 = (lo[2] - y->O.z) * y->idz, z1 = (hi[2] - y->O.z) * y->idz;
    float tmin = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    float tfar = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    tmin *= 0.9999995f;              /* 1 - 2^-21: ~4 ulp */
    tfar *= 1.0000005f;
    *tn = tmin;
    return tmin <= tfar;
}

typedef struct { uint64_t nodes, tris; } orc_counters;

/* ------------------------------------------------------------- cast_rays */
typedef struct { float t, u, v; uint32_t geom, prim; v3 Ng; } best_t;

static inline void consider(const orc_tri *tr, const ray_t *y, best_t *b)
{
    mt_hit h;
    if (!mt_test(tr, y->O, y->D, 0.0f, INFINITY, &h)) return;
    float t = h.T / h.absDen;
    /* closest hit; exact ties go to the lowest (geometry, primitive) */
    int better = (t < b->t) ||
                 (t == b->t && (tr->geom < b->geom || (tr->geom == b->geom && tr->prim < b->prim)));
    if (!better) return;
    b->t = t; b->u = h.U / h.absDen; b->v = h.V / h.absDen;
    b->geom = tr->geom; b->prim = tr->prim;
    b->Ng = v3cross(tr->e2, tr->e1);
}

#define ORC_STACK 256

static void cast_one(const orc_scene *s, const ray_t *y, int mode, best_t *b, orc_counters *c)
{
    uint64_t n = s->ntris;
    b->t = INFINITY; b->u = b->v = 0.0f; b->geom = b->prim = ORC_INVALID_ID;
    b->Ng.x = b->Ng.y = b->Ng.z = 0.0f;
    if (n == 0) return;
    if (mode == 0) { for (uint64_t i = 0; i < n; ++i) consider(&s->tris[i], y, b); return; }
    const orc_tree T = tree_of(s, mode);
    float tn;
    if (n == 1) {
        c->nodes++;
        if (box_test(s->leaf_lo, s->leaf_hi, y, b->t, &tn)) { c->tris++; consider(&s->stris[0], y, b); }
        return;
    }
    int32_t stk[ORC_STACK]; float stn[ORC_STACK]; int sp = 0;
    c->nodes++;
    if (!box_test(T.nodes[0].lo, T.nodes[0].hi, y, b->t, &tn)) return;
    int32_t cur = 0;
    for (;;) {
        if (cur >= 0) {
            const orc_node *nd = &T.nodes[cur];
            int32_t ch[2] = { nd->left, nd->right };
            float ctn[2]; int hit[2];
            for (int k = 0; k < 2; ++k) {
                const float *lo = ch[k] < 0 ? T.leaf_lo + 3 * (~ch[k]) : T.nodes[ch[k]].lo;
                const float *hi = ch[k] < 0 ? T.leaf_hi + 3 * (~ch[k]) : T.nodes[ch[k]].hi;
                c->nodes++;
                hit[k] = box_test(lo, hi, y, b->t, &ctn[k]);
            }
            if (hit[0] && hit[1]) {
                int nearc = ctn[1] < ctn[0] ? 1 : 0;
                stk[sp] = ch[1 - nearc]; stn[sp] = ctn[1 - nearc]; ++sp;
                cur = ch[nearc];
                continue;
            }
            if (hit[0]) { cur = ch[0]; continue; }
            if (hit[1]) { cur = ch[1]; continue; }
        } else {
            c->tris++;
            consider(&T.stris[~cur], y, b);
        }
        /* pop, skipping entries the current best already excludes */
        for (;;) {
            if (sp == 0) return;
            --sp;
            if (stn[sp] <= b->t) { cur = stk[sp]; break; }
        }
    }
}

/* Open3D RaycastingScene::CastRays output contract */
void orc_cast_rays(const orc_scene *s, const float *rays, uint64_t N, int mode,
                   float *t_hit, uint32_t *geom, uint32_t *prim, float *uv, float *nrm,
                   orc_counters *counters)
{
    uint64_t cn = 0, ct = 0;
    if (mode == 2 && !s->sah_nodes) build_sah((orc_scene *)s);
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : cn, ct)
    for (int64_t i = 0; i < (int64_t)N; ++i) {
        ray_t y = load_ray(rays + 6 * i);
        best_t b; orc_counters c = { 0, 0 };
        cast_one(s, &y, mode, &b, &c);
        cn += c.nodes; ct += c.tris;
        if (t_hit) t_hit[i] = b.t;
        if (geom) geom[i] = b.geom;
        if (prim) prim[i] = b.prim;
        if (uv) { uv[2 * i] = b.u; uv[2 * i + 1] = b.v; }
        if (nrm) {
            if (b.prim != ORC_INVALID_ID) {
                float len2 = v3dot(b.Ng, b.Ng);
                float inv = 1.0f / sqrtf(len2);
                nrm[3 * i] = b.Ng.x * inv; nrm[3 * i + 1] = b.Ng.y * inv; nrm[3 * i + 2] = b.Ng.z * inv;
            } else { nrm[3 * i] = nrm[3 * i + 1] = nrm[3 * i + 2] = 0.0f; }
        }
    }
    if (counters) { counters->nodes = cn; counters->tris = ct; }
}

/* ------------------------------------------- all-hits (count / list / occl) */
typedef struct { float t, u, v; uint32_t geom, prim; } hit_rec;
typedef struct { hit_rec *h; int n, cap; } hit_list;

static inline void hl_push(hit_list *l, hit_rec r)
{
    if (l->n == l->cap) { l->cap = l->cap ? 2 * l->cap : 16; l->h = (hit_rec *)realloc(l->h, sizeof(hit_rec) * l->cap); }
    l->h[l->n++] = r;
}

static int hit_cmp(const void *a, const void *b)
{
    const hit_rec *x = (const hit_rec *)a, *y = (const hit_rec *)b;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    if (x->geom != y->geom) return x->geom < y->geom ? -1 : 1;
    if (x->prim != y->prim) return x->prim < y->prim ? -1 : 1;
    return 0;
}

static inline void collect(const orc_tri *tr, const ray_t *y, float tnear, float tfar, hit_list *l)
{
    mt_hit h;
    if (!mt_test(tr, y->O, y->D, tnear, tfar, &h)) return;
    hit_rec r = { h.T / h.absDen, h.U / h.absDen, h.V / h.absDen, tr->geom, tr->prim };
    hl_push(l, r);
}

/* every accepted hit on (tnear, tfar]; first_only stops at the first (occlusion) */
static void all_hits_ex(const orc_scene *s, const ray_t *y, int mode, float tnear, float tfar,
                        int first_only, hit_list *l, orc_counters *c, float eps, uint8_t *edge)
{
#define VISIT(TR) do { collect((TR), y, tnear, tfar, l); \
        if (edge && mt_near_edge((TR), y->O, y->D, eps)) *edge |= 1; } while (0)
    uint64_t n = s->ntris;
    l->n = 0;
    if (n == 0) return;
    if (mode == 0) {
        for (uint64_t i = 0; i < n; ++i) { VISIT(&s->tris[i]); if (first_only && l->n) return; }
        return;
    }
    const orc_tree T = tree_of(s, mode);
    float tn;
    if (n == 1) {
        c->nodes++;
        if (box_test(s->leaf_lo, s->leaf_hi, y, tfar, &tn)) { c->tris++; VISIT(&s->stris[0]); }
        return;
    }
    int32_t stk[ORC_STACK]; int sp = 0;
    c->nodes++;
    if (!box_test(T.nodes[0].lo, T.nodes[0].hi, y, tfar, &tn)) return;
    int32_t cur = 0;
    for (;;) {
        if (cur >= 0) {
            const orc_node *nd = &T.nodes[cur];
            int32_t ch[2] = { nd->left, nd->right };
            int hit[2];
            for (int k = 0; k < 2; ++k) {
                const float *lo = ch[k] < 0 ? T.leaf_lo + 3 * (~ch[k]) : T.nodes[ch[k]].lo;
                const float *hi = ch[k] < 0 ? T.leaf_hi + 3 * (~ch[k]) : T.nodes[ch[k]].hi;
                c->nodes++;
                hit[k] = box_test(lo, hi, y, tfar, &tn);
            }
            if (hit[0] && hit[1]) { stk[sp++] = ch[1]; cur = ch[0]; continue; }
            if (hit[0]) { cur = ch[0]; continue; }
            if (hit[1]) { cur = ch[1]; continue; }
        } else {
            c->tris++;
            VISIT(&T.stris[~cur]);
            if (first_only && l->n) return;
        }
        if (sp == 0) return;
        cur = stk[--sp];
    }
#undef VISIT
}

static void all_hits(const orc_scene *s, const ray_t *y, int mode, float tnear, float tfar,
                     int first_only, hit_list *l, orc_counters *c)
{
    all_hits_ex(s, y, mode, tnear, tfar, first_only, l, c, 0.0f, NULL);
}

/* Open3D CountIntersectionsFunc dedup, restated order-independently: per
 * geometry, hits with the same t are one intersection; the survivor is the
 * lowest primitive id.  Input list is sorted by (t, geom, prim) in place. */
static int dedup_hits(hit_list *l)
{
    qsort(l->h, l->n, sizeof(hit_rec), hit_cmp);
    int m = 0;
    for (int i = 0; i < l->n; ++i) {
        if (m > 0 && l->h[m - 1].t == l->h[i].t && l->h[m - 1].geom == l->h[i].geom) continue;
        l->h[m++] = l->h[i];
    }
    l->n = m;
    return m;
}

void orc_count_intersections(const orc_scene *s, const float *rays, uint64_t N, int mode,
                             int32_t *out, orc_counters *counters)
{
    uint64_t cn = 0, ct = 0;
    if (mode == 2 && !s->sah_nodes) build_sah((orc_scene *)s);
#pragma omp parallel reduction(+ : cn, ct)
    {
        hit_list l = { NULL, 0, 0 };
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)N; ++i) {
            ray_t y = load_ray(rays + 6 * i);
            orc_counters c = { 0, 0 };
            all_hits(s, &y, mode, 0.0f, INFINITY, 0, &l, &c);
            out[i] = dedup_hits(&l);
            cn += c.nodes; ct += c.tris;
        }
        free(l.h);
    }
    if (counters) { counters->nodes = cn; counters->tris = ct; }
}

void orc_test_occlusions(const orc_scene *s, const float *rays, uint64_t N, int mode,
                         float tnear, float tfar, uint8_t *out)
{
#pragma omp parallel
    {
        hit_list l = { NULL, 0, 0 };
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)N; ++i) {
            ray_t y = load_ray(rays + 6 * i);
            orc_counters c = { 0, 0 };
            all_hits(s, &y, mode, tnear, tfar, 1, &l, &c);
            out[i] = l.n > 0;
        }
        free(l.h);
    }
}

/* Open3D ListIntersections: CSR of the deduplicated hits; per ray sorted by
 * (t, geom, prim).  Call with ray_splits from orc_count_intersections scanned. */
void orc_list_intersections(const orc_scene *s, const float *rays, uint64_t N, int mode,
                            const int64_t *ray_splits, int64_t *ray_ids, float *t_hit,
                            uint32_t *geom, uint32_t *prim, float *uv)
{
#pragma omp parallel
    {
        hit_list l = { NULL, 0, 0 };
#pragma omp for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)N; ++i) {
            ray_t y = load_ray(rays + 6 * i);
            orc_counters c = { 0, 0 };
            all_hits(s, &y, mode, 0.0f, INFINITY, 0, &l, &c);
            int m = dedup_hits(&l);
            int64_t base = ray_splits[i];
            for (int k = 0; k < m; ++k) {
                ray_ids[base + k] = i; t_hit[base + k] = l.h[k].t;
                geom[base + k] = l.h[k].geom; prim[base + k] = l.h[k].prim;
                uv[2 * (base + k)] = l.h[k].u; uv[2 * (base + k) + 1] = l.h[k].v;
            }
        }
        free(l.h);
    }
}

/* Edge report (north_star: "rays within 1e-6 of a shared triangle edge,
 * counted and reported").  bit0: some triangle plane is crossed in (0,inf)
 * with a barycentric within eps of an edge; bit1: two accepted hits share t. */
void orc_edge_flags(const orc_scene *s, const float *rays, uint64_t N, int mode, float eps, uint8_t *flags)
{
#pragma omp parallel
    {
        hit_list l = { NULL, 0, 0 };
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < (int64_t)N; ++i) {
            ray_t y = load_ray(rays + 6 * i);
            uint8_t f = 0;
            orc_counters c = { 0, 0 };
            all_hits_ex(s, &y, mode, 0.0f, INFINITY, 0, &l, &c, eps, &f);
            qsort(l.h, l.n, sizeof(hit_rec), hit_cmp);
            for (int k = 1; k < l.n; ++k) if (l.h[k].t == l.h[k - 1].t) f |= 2;
            flags[i] = f;
        }
        free(l.h);
    }
}

/* ------------------------------------------------------- closest points
 * Open3D RaycastingScene::ComputeClosestPoints (rtcPointQuery + ClosestPointFunc):
 * closest point on each candidate triangle by the region test of Ericson,
 * "Real-Time Collision Detection" 5.1.5 (the closestPointTriangle of Embree's
 * closest_point tutorial, which Open3D uses), keep the smallest distance.
 * Reference call sites: pyQSM/viz/ray_casting.py:250,255 (compute_signed_distance).
 * Points whose nearest feature is an edge or vertex are equidistant to several
 * triangles; Embree keeps whichever it meets first, here the lowest
 * (geometry, primitive) among exactly equal squared distances wins.
 * Triangle corners come from the 48-byte record: a = v0, ab = -e1, ac = e2. */
typedef struct { float d2; v3 q; float u, v; uint32_t geom, prim; v3 Ng; } cp_best;

static inline v3 v3add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 v3madd(v3 a, float s, v3 b) { v3 r = { fmaf(s, b.x, a.x), fmaf(s, b.y, a.y), fmaf(s, b.z, a.z) }; return r; }

static inline void cp_triangle(const orc_tri *tr, v3 p, v3 *q, float *bu, float *bv)
{
    v3 a = tr->v0;
    v3 ab = { -tr->e1.x, -tr->e1.y, -tr->e1.z }, ac = tr->e2;
    v3 b = v3add(a, ab), c = v3add(a, ac);
    v3 ap = v3sub(p, a);
    float d1 = v3dot(ab, ap), d2 = v3dot(ac, ap);
    if (d1 <= 0.0f && d2 <= 0.0f) { *q = a; *bu = 0.0f; *bv = 0.0f; return; }
    v3 bp = v3sub(p, b);
    float d3 = v3dot(ab, bp), d4 = v3dot(ac, bp);
    if (d3 >= 0.0f && d4 <= d3) { *q = b; *bu = 1.0f; *bv = 0.0f; return; }
    v3 cp = v3sub(p, c);
    float d5 = v3dot(ab, cp), d6 = v3dot(ac, cp);
    if (d6 >= 0.0f && d5 <= d6) { *q = c; *bu = 0.0f; *bv = 1.0f; return; }
    float vc = fmaf(d1, d4, -(d3 * d2));
    if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) { float v = d1 / (d1 - d3); *q = v3madd(a, v, ab); *bu = v; *bv = 0.0f; return; }
    float vb = fmaf(d5, d2, -(d1 * d6));
    if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) { float w = d2 / (d2 - d6); *q = v3madd(a, w, ac); *bu = 0.0f; *bv = w; return; }
    float va = fmaf(d3, d6, -(d5 * d4));
    if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        *q = v3madd(b, w, v3sub(c, b)); *bu = 1.0f - w; *bv = w; return;
    }
    float denom = 1.0f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    *q = v3madd(v3madd(a, v, ab), w, ac); *bu = v; *bv = w;
}

static inline void cp_consider(const orc_tri *tr, v3 p, cp_best *best)
{
    v3 q; float u, v;
    cp_triangle(tr, p, &q, &u, &v);
    v3 d = v3sub(q, p);
    float d2 = v3dot(d, d);
    int better = (d2 < best->d2) ||
                 (d2 == best->d2 && (tr->geom < best->geom || (tr->geom == best->geom && tr->prim < best->prim)));
    if (!better) return;
    best->d2 = d2; best->q = q; best->u = u; best->v = v; best->geom = tr->geom; best->prim = tr->prim;
    best->Ng = v3cross(tr->e2, tr->e1);
}

static inline float box_dist2(const float *lo, const float *hi, v3 p)
{
    float dx = fmaxf(fmaxf(lo[0] - p.x, p.x - hi[0]), 0.0f);
    float dy = fmaxf(fmaxf(lo[1] - p.y, p.y - hi[1]), 0.0f);
    float dz = fmaxf(fmaxf(lo[2] - p.z, p.z - hi[2]), 0.0f);
    return fmaf(dx, dx, fmaf(dy, dy, dz * dz));
}

static void cp_one(const orc_scene *s, v3 p, int mode, cp_best *best)
{
    uint64_t n = s->ntris;
    best->d2 = INFINITY; best->geom = best->prim = ORC_INVALID_ID; best->u = best->v = 0.0f;
    best->q.x = best->q.y = best->q.z = 0.0f; best->Ng = best->q;
    if (n == 0) return;
    if (mode == 0 || n == 1) { for (uint64_t i = 0; i < n; ++i) cp_consider(&s->tris[i], p, best); return; }
    int32_t stk[ORC_STACK]; float sd[ORC_STACK]; int sp = 0;
    int32_t cur = 0;
    for (;;) {
        if (cur >= 0) {
            const orc_node *nd = &s->nodes[cur];
            int32_t ch[2] = { nd->left, nd->right };
            float cd[2];
            for (int k = 0; k < 2; ++k) {
                const float *lo = ch[k] < 0 ? s->leaf_lo + 3 * (~ch[k]) : s->nodes[ch[k]].lo;
                const float *hi = ch[k] < 0 ? s->leaf_hi + 3 * (~ch[k]) : s->nodes[ch[k]].hi;
                cd[k] = box_dist2(lo, hi, p);
            }
            int h0 = cd[0] <= best->d2, h1 = cd[1] <= best->d2;
            if (h0 && h1) {
                int nearc = cd[1] < cd[0] ? 1 : 0;
                stk[sp] = ch[1 - nearc]; sd[sp] = cd[1 - nearc]; ++sp;
                cur = ch[nearc];
                continue;
            }
            if (h0) { cur = ch[0]; continue; }
            if (h1) { cur = ch[1]; continue; }
        } else {
            cp_consider(&s->stris[~cur], p, best);
        }
        for (;;) {
            if (sp == 0) return;
            --sp;
            if (sd[sp] <= best->d2) { cur = stk[sp]; break; }
        }
    }
}

/* points[N][3] -> closest[N][3], distance[N], geometry/primitive ids, uv[N][2], normals[N][3] */
void orc_closest_points(const orc_scene *s, const float *pts, uint64_t N, int mode, float *closest, float *dist,
                        uint32_t *geom, uint32_t *prim, float *uv, float *nrm)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)N; ++i) {
        v3 p = { pts[3 * i], pts[3 * i + 1], pts[3 * i + 2] };
        cp_best b;
        cp_one(s, p, mode, &b);
        int ok = b.prim != ORC_INVALID_ID;
        if (closest) { closest[3 * i] = b.q.x; closest[3 * i + 1] = b.q.y; closest[3 * i + 2] = b.q.z; }
        if (dist) dist[i] = ok ? sqrtf(b.d2) : INFINITY;
        if (geom) geom[i] = b.geom;
        if (prim) prim[i] = b.prim;
        if (uv) { uv[2 * i] = b.u; uv[2 * i + 1] = b.v; }
        if (nrm) {
            if (ok) {
                float inv = 1.0f / sqrtf(v3dot(b.Ng, b.Ng));
                nrm[3 * i] = b.Ng.x * inv; nrm[3 * i + 1] = b.Ng.y * inv; nrm[3 * i + 2] = b.Ng.z * inv;
            } else { nrm[3 * i] = nrm[3 * i + 1] = nrm[3 * i + 2] = 0.0f; }
        }
    }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
