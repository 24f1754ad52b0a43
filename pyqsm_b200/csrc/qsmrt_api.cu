// qsmrt_api.cu -- the scene object and the extern "C" surface of libqsmrt.so
// (include/qsmrt.h).  Each entry point stands in for one method of Open3D's
// RaycastingScene as the reference calls it (pyQSM/viz/ray_casting.py; the
// header cites the line numbers).  No CPU fallback: every path needs a CUDA
// device and fails with an error otherwise.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <map>
#include <unordered_map>
#include <mutex>

#include "../../include/qsmrt.h"
#include "common.cuh"
#include "build.h"
#include "traverse.h"

static thread_local char g_err[512] = "";
constexpr int QSMRT_MAX_DEVICES = 64;

void qsmrt_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define FAIL(...) do { qsmrt_set_error(__VA_ARGS__); return 1; } while (0)

struct Geometry {
    float *verts = nullptr; uint32_t *idx = nullptr;
    uint64_t V = 0, T = 0;
    float lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };   // over the vertices its triangles reference
    uint64_t nrefs = 0;          // references its triangles want when slivers are split (>= T), counted at registration ...
    int ref_max = 0; float ref_aspect = 0.0f;      // ... with these split options
};

struct HostPipe {            // device staging of the *_host entry points: rays in, up to 32 B/ray of results out
    static constexpr int NBUF = 3;
    uint64_t chunk = 0; size_t out_bytes = 0;       // rays per stage; result bytes per ray the staging holds
    float *rays[NBUF] = {}; char *out[NBUF] = {};
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t e_in[NBUF] = {}, e_run[NBUF] = {}, e_out[NBUF] = {};
};

// Builder options of a scene (qsmrt_scene_set_option), read at the next commit.
struct BuildOptions {
    int leaf_max = 2;            // triangles per collapsed leaf; 2 measured best on C2
    int keep_bnodes = 0;         // materialise the complete 32-byte binary node array (builder-vs-oracle test)
    float quant_frac = 0.15f;    // 32-byte nodes when 6 grid cells <= this share of the mean leaf diagonal
    int climb_capacity = 0;      // > 0: cap of the climb work list (test hook)
    int sort_variant = 1;        // 0 classic radix passes, 1 onesweep
    int allow_qnodes = 1;        // traversal may read the 32-byte nodes when the build made them
    int split_max = 8;           // a sliver triangle enters the build as up to this many references (1 = no splitting)
    float split_aspect = 2.0f;   // ... one per `split_aspect` units of its aspect L^2 / 2A
};

struct qsmrt_scene {
    int device = 0;
    std::vector<Geometry> geoms;
    bool committed = false;
    // concatenated mesh (aliases geoms[0] when there is a single geometry)
    float *verts = nullptr; uint32_t *idx = nullptr; bool own_concat = false;
    uint64_t ntris = 0, nverts = 0;
    uint64_t nleaves = 0;                            // leaves of the LBVH: ntris, or the number of references when slivers were split
    uint64_t *goff = nullptr, *voff = nullptr;       // device [ngeoms+1]
    // build products kept for traversal / introspection
    uint64_t *keys = nullptr; uint32_t *order = nullptr;
    BNode *bnodes = nullptr; TNode *tnodes = nullptr; TriRec *tris = nullptr;
    QNode *qnodes = nullptr; bool use_qnodes = false; float glo[3] = {}, cell[3] = {};
    BuildParams *params = nullptr;
    qsmrt_stats stats{};
    BuildOptions bopt;
    TrvState trv;
    cudaTextureObject_t node_tex = 0;
    // list_intersections cache between _count and _fill
    const float *list_rays = nullptr; uint64_t list_n = 0;
    ListStash list_stash; int list_max_fast = 0;     // the hit records _count collected for _fill
    uint64_t host_chunk = 0; int host_ramp = 1;      // QSMRT_OPT_HOST_CHUNK / _RAMP
    HostPipe pipe;
    float *sweep_dev = nullptr; uint32_t sweep_cap = 0;     // per-grid constants of qsmrt_sun_exposure_sweep
    // qsmrt_sky_visibility: Morton order of the query points (keys / values double-buffered for the radix sort)
    struct SkyOrder { uint64_t *keys = nullptr, *keys_tmp = nullptr; uint32_t *vals = nullptr, *vals_tmp = nullptr, *scratch = nullptr; uint64_t cap = 0; } sky;
};

namespace {

// ---- device memory: a small per-device block cache in front of cudaMalloc / cudaFree.
// The reference builds a fresh RaycastingScene in every function (ray_casting.py:65,155,218,241,275,316); a commit
// makes ~25 allocations and as many frees, and on a 50k-triangle tree those driver calls took 6-18 ms around a
// 0.23 ms build.  Freed blocks are kept (up to QSMRT_CACHE_MB, default 1024) and handed out again for requests of
// nearly the same size.  A block is only recycled after cudaDeviceSynchronize(), which is what cudaFree implied:
// no kernel that still reads it can be in flight.
struct BlockCache {
    std::mutex m;
    std::multimap<size_t, void *> idle[QSMRT_MAX_DEVICES];
    std::unordered_map<void *, size_t> size_of;          // every block we handed out or hold
    size_t idle_bytes = 0, cap = 0;
    bool cap_read = false;
    size_t limit()
    {
        if (!cap_read) { const char *e = getenv("QSMRT_CACHE_MB"); cap = (size_t)(e ? atoll(e) : 1024) << 20; cap_read = true; }
        return cap;
    }
    void release_all()
    {
        for (auto &mm : idle) { for (auto &kv : mm) { size_of.erase(kv.second); cudaFree(kv.second); } mm.clear(); }
        idle_bytes = 0;
    }
} g_blocks;

int dmalloc_bytes(void **p, size_t bytes)
{
    *p = nullptr;
    bytes = (std::max<size_t>(bytes, 1) + 511) & ~(size_t)511;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_blocks.m);
    if (dev >= 0 && dev < QSMRT_MAX_DEVICES) {
        auto &mm = g_blocks.idle[dev];
        auto it = mm.lower_bound(bytes);
        if (it != mm.end() && it->first <= bytes + bytes / 4 + (64u << 10)) {
            *p = it->second; g_blocks.idle_bytes -= it->first; mm.erase(it);
            return 0;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {                      // out of memory: give the cached blocks back and try once more
        cudaGetLastError();
        g_blocks.release_all();
        e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) { *p = nullptr; qsmrt_set_error("cudaMalloc(%llu bytes): %s", (unsigned long long)bytes, cudaGetErrorString(e)); return 1; }
    g_blocks.size_of[*p] = bytes;
    return 0;
}

thread_local int tl_frees_synced = 0;
// one device synchronisation for a run of frees (scene teardown, end of a commit) instead of one per block
struct SyncedFrees {
    SyncedFrees() { cudaDeviceSynchronize(); ++tl_frees_synced; }
    ~SyncedFrees() { --tl_frees_synced; }
};

void dfree_bytes(void *p)
{
    if (!p) return;
    if (!tl_frees_synced) cudaDeviceSynchronize();      // cudaFree's implicit guarantee, kept for recycled blocks
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_blocks.m);
    auto it = g_blocks.size_of.find(p);
    if (it != g_blocks.size_of.end() && dev >= 0 && dev < QSMRT_MAX_DEVICES &&
        g_blocks.idle_bytes + it->second <= g_blocks.limit()) {
        g_blocks.idle[dev].emplace(it->second, p);
        g_blocks.idle_bytes += it->second;
        return;
    }
    if (it != g_blocks.size_of.end()) g_blocks.size_of.erase(it);
    cudaFree(p);
}

template <class T> int dmalloc(T **p, uint64_t count)
{
    return dmalloc_bytes(reinterpret_cast<void **>(p), (size_t)count * sizeof(T));
}
template <class T> void dfree(T *&p) { dfree_bytes(p); p = nullptr; }

void free_list_stash(qsmrt_scene *s)
{
    ListStash &ls = s->list_stash;
    s->list_max_fast = 0; ls.cap = 0;
    if (!ls.base && !ls.t && !ls.geom && !ls.prim && !ls.uv && !ls.count) return;
    SyncedFrees batch;
    dfree(ls.base); dfree(ls.t); dfree(ls.geom); dfree(ls.prim); dfree(ls.uv); dfree(ls.count);
}

void free_build(qsmrt_scene *s)
{
    free_list_stash(s);
    if (s->own_concat) { dfree(s->verts); dfree(s->idx); }
    s->verts = nullptr; s->idx = nullptr; s->own_concat = false;
    dfree(s->goff); dfree(s->voff); dfree(s->keys); dfree(s->order);
    dfree(s->bnodes); dfree(s->tnodes); dfree(s->tris); dfree(s->params); dfree(s->qnodes);
    s->use_qnodes = false;
    s->list_rays = nullptr; s->list_n = 0;
    if (s->node_tex) { cudaDestroyTextureObject(s->node_tex); s->node_tex = 0; }
    s->committed = false;
}

void free_pipe(HostPipe &hp)
{
    for (int b = 0; b < HostPipe::NBUF; ++b) {
        dfree(hp.rays[b]); dfree(hp.out[b]);
        if (hp.e_in[b]) cudaEventDestroy(hp.e_in[b]);
        if (hp.e_run[b]) cudaEventDestroy(hp.e_run[b]);
        if (hp.e_out[b]) cudaEventDestroy(hp.e_out[b]);
        hp.e_in[b] = hp.e_run[b] = hp.e_out[b] = nullptr;
    }
    if (hp.s_in) cudaStreamDestroy(hp.s_in);
    if (hp.s_run) cudaStreamDestroy(hp.s_run);
    if (hp.s_out) cudaStreamDestroy(hp.s_out);
    hp.s_in = hp.s_run = hp.s_out = nullptr;
    hp.chunk = 0; hp.out_bytes = 0;
}

// Read-bandwidth probe (qsmrt_util_read_sweep): every thread streams 32-byte chunks of the buffer, grid-stride,
// `reps` times; XOR keeps the loads alive.  A buffer well inside the 126 MB L2 gives the L2 -> SM read peak the
// traversal kernel's node / triangle fetches are measured against, a buffer far above it the HBM read rate.
__global__ void __launch_bounds__(256)
k_read_sweep(const uint4 *__restrict__ buf, uint64_t n32, uint32_t reps, uint32_t *sink)
{
    uint32_t acc = 0;
    for (uint32_t r = 0; r < reps; ++r)
        for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n32; i += (uint64_t)gridDim.x * 256ull) {
            uint32_t w0, w1, w2, w3, w4, w5, w6, w7;
            asm volatile("ld.global.cg.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"        // .cg: L2 only, never an L1 hit
                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3), "=r"(w4), "=r"(w5), "=r"(w6), "=r"(w7) : "l"(buf + 2 * i));
            acc ^= w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7;
        }
    if (acc == 0x9E3779B9u) *sink = acc;        // practically never: the compiler cannot drop the loads
}

__global__ void k_rebase_idx(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n3, uint32_t add)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n3) out[i] = in[i] + add;
}

// Scene files: everything a traversal kernel would follow from a loaded BVH must stay inside the loaded arrays.
// The node array has unused slots (subtrees collapsed into leaves), so the nodes are checked by walking the tree
// from the root one level per launch: child references and leaf ranges in range, the quantised twin naming the same
// children; the host bounds the number of nodes visited (a cycle or a shared subtree exceeds it) and the depth.
__global__ void __launch_bounds__(256)
k_validate_level(const TNode *__restrict__ tn, const QNode *__restrict__ qn, uint64_t nn, uint64_t nleaves,
                 const int *__restrict__ frontier, uint32_t count, int *__restrict__ next, uint32_t *next_count, uint32_t *bad)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= count) return;
    const int node = frontier[i];
    const int c[2] = { tn[node].d.x, tn[node].d.y };
    bool ok = !qn || ((int)qn[node].w[6] == c[0] && (int)qn[node].w[7] == c[1]);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (c[k] >= 0) {
            ok = ok && c[k] != 0 && (uint64_t)c[k] < nn;               // the root is nobody's child
            if (ok) { const uint32_t at = atomicAdd(next_count, 1u); if (at < nn) next[at] = c[k]; }
        } else {
            const uint32_t r = ~(uint32_t)c[k];
            ok = ok && (uint64_t)(r >> 2) + (r & 3u) + 1u <= nleaves;
        }
    }
    if (!ok) *bad = 1u;
}

// ... and the triangle records and the order array, which have no unused slots
__global__ void __launch_bounds__(256)
k_validate_leaves(const TriRec *__restrict__ tris, const uint32_t *__restrict__ order, uint64_t nleaves, uint64_t ntris,
                  const uint64_t *__restrict__ goff, uint32_t ngeoms, uint32_t *bad)
{
    const uint64_t i = blockIdx.x * 256ull + threadIdx.x;
    if (i >= nleaves) return;
    const uint32_t g = __float_as_uint(tris[i].p1.w), p = __float_as_uint(tris[i].p0.w);
    if (!(order[i] < ntris && g < ngeoms && goff[g] + p < goff[g + 1])) *bad = 1u;
}

// QSM cylinder records -> triangle mesh with Open3D's create_cylinder topology (axis z, centred, 2 cap centres +
// (split+1) rings of `res` vertices; 2*res cap + 2*res*split side triangles), rotated from +z onto the record's
// axis (Rodrigues) and translated to its centre: what get_shape(..., shape="cylinder") builds on the CPU
// (pyQSM/geometry/point_cloud_processing.py:266-304) from the records of qsm_generation.py:171-178.
__global__ void __launch_bounds__(128)
k_cylinders(const float *__restrict__ rec, uint64_t n, uint32_t res, uint32_t split, float *__restrict__ verts, uint32_t *__restrict__ idx)
{
    const uint64_t c = blockIdx.x;
    if (c >= n) return;
    const uint32_t V = res * (split + 1) + 2, T = 2 * res + 2 * res * split;
    const float cx = rec[8 * c], cy = rec[8 * c + 1], cz = rec[8 * c + 2];
    float ax = rec[8 * c + 3], ay = rec[8 * c + 4], az = rec[8 * c + 5];
    const float radius = rec[8 * c + 6], height = rec[8 * c + 7];
    float len = sqrtf(ax * ax + ay * ay + az * az);
    if (len > 0.0f) { ax /= len; ay /= len; az /= len; } else { ax = 0.0f; ay = 0.0f; az = 1.0f; }
    // R = I + [v]x + [v]x^2 / (1 + c), v = z x a = (-ay, ax, 0), c = az;  a = -z: rotate pi about x
    float R[9];
    if (az > -0.999999f) {
        const float k = 1.0f / (1.0f + az), vx = -ay, vy = ax;
        R[0] = 1.0f - vy * vy * k; R[1] = vx * vy * k;        R[2] = vy;
        R[3] = vx * vy * k;        R[4] = 1.0f - vx * vx * k; R[5] = -vx;
        R[6] = -vy;                R[7] = vx;                 R[8] = 1.0f - (vx * vx + vy * vy) * k;
    } else {
        R[0] = 1.0f; R[1] = 0.0f; R[2] = 0.0f; R[3] = 0.0f; R[4] = -1.0f; R[5] = 0.0f; R[6] = 0.0f; R[7] = 0.0f; R[8] = -1.0f;
    }
    float *vo = verts + 3ull * V * c;
    for (uint32_t k = threadIdx.x; k < V; k += blockDim.x) {
        float x, y, z;
        if (k == 0) { x = 0.0f; y = 0.0f; z = 0.5f * height; }
        else if (k == 1) { x = 0.0f; y = 0.0f; z = -0.5f * height; }
        else {
            const uint32_t ring = (k - 2) / res, j = (k - 2) % res;
            const float th = 6.2831853071795864f * (float)j / (float)res;
            x = cosf(th) * radius; y = sinf(th) * radius; z = 0.5f * height - (height / (float)split) * (float)ring;
        }
        vo[3 * k]     = R[0] * x + R[1] * y + R[2] * z + cx;
        vo[3 * k + 1] = R[3] * x + R[4] * y + R[5] * z + cy;
        vo[3 * k + 2] = R[6] * x + R[7] * y + R[8] * z + cz;
    }
    uint32_t *to = idx + 3ull * T * c;
    const uint32_t vb = (uint32_t)(V * c);
    for (uint32_t k = threadIdx.x; k < T; k += blockDim.x) {
        uint32_t a, b, d;
        if (k < 2 * res) {                               // caps, interleaved top / bottom like the host generator
            const uint32_t j = k >> 1, j1 = (j + 1) % res;
            if ((k & 1) == 0) { a = 0; b = 2 + j; d = 2 + j1; }
            else { const uint32_t bb = 2 + res * split; a = 1; b = bb + j1; d = bb + j; }
        } else {
            const uint32_t q = k - 2 * res, i = q / (2 * res), r2 = q % (2 * res), j = r2 >> 1, j1 = (j + 1) % res;
            const uint32_t b1 = 2 + res * i, b2 = b1 + res;
            if ((r2 & 1) == 0) { a = b2 + j; b = b1 + j1; d = b1 + j; }
            else { a = b2 + j; b = b2 + j1; d = b1 + j1; }
        }
        to[3 * k] = vb + a; to[3 * k + 1] = vb + b; to[3 * k + 2] = vb + d;
    }
}

// Every ABI entry runs on the scene's device and hands the caller's current device back on return (a process
// that drives several GPUs -- or torch, whose current device we must not move -- never sees it change).
struct DeviceGuard {
    int prev = -1; cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != dev) err = cudaSetDevice(dev);
        if (prev == dev) prev = -1;                 // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define SCENE_ENTER(s) \
    if (!(s)) FAIL("null scene"); \
    DeviceGuard device_guard_((s)->device); \
    if (device_guard_.err != cudaSuccess) FAIL("cudaSetDevice(%d): %s", (s)->device, cudaGetErrorString(device_guard_.err))

int check_rays(const float *rays, uint64_t N)
{
    if (N && !rays) FAIL("rays pointer is null");
    if (reinterpret_cast<uintptr_t>(rays) & 7u) FAIL("rays must be 8-byte aligned");
    return 0;
}

// references the triangles of a geometry want under the scene's current split options (registration-time statistic;
// recounted by the commit only if the options were changed since)
int count_refs(qsmrt_scene *s, Geometry &g)
{
    g.ref_max = s->bopt.split_max; g.ref_aspect = s->bopt.split_aspect; g.nrefs = g.T;
    if (g.T == 0 || s->bopt.split_max <= 1) return 0;
    unsigned long long *d = nullptr, h = 0;
    if (dmalloc(&d, 1)) return 1;
    int rc = cudaMemset(d, 0, sizeof(h)) != cudaSuccess || lbvh_split_count(g.verts, g.idx, g.T, g.ref_max, g.ref_aspect, nullptr, d, nullptr) ||
             cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess;
    dfree(d);
    if (rc) { if (!g_err[0]) qsmrt_set_error("reference count failed: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    g.nrefs = h;
    return 0;
}

SceneView view_of(qsmrt_scene *s)
{
    if (s->trv.opt.node_path != 0 && !s->node_tex && s->tnodes) {
        // float4 texture view of the node array, made on first use: only the TEX-path experiment
        // (QSMRT_OPT_NODE_PATH) reads it, and creating it cost every commit a driver call
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = s->tnodes;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = std::max<uint64_t>(s->nleaves - 1, 1) * sizeof(TNode);
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        if (cudaCreateTextureObject(&s->node_tex, &rd, &td, nullptr) != cudaSuccess) { s->node_tex = 0; cudaGetLastError(); }
    }
    SceneView v;
    v.nodes = s->tnodes; v.tris = s->tris; v.ntris = (uint32_t)s->ntris; v.height = s->stats.bvh_height;
    v.node_tex = s->node_tex;
    v.qnodes = (s->use_qnodes && s->bopt.allow_qnodes) ? s->qnodes : nullptr;
    for (int a = 0; a < 3; ++a) { v.glo[a] = s->glo[a]; v.cell[a] = s->cell[a]; }
    return v;
}

// Scratch of one commit; everything here goes back to the block cache on every exit path.
struct CommitScratch {
    uint64_t *keys_tmp = nullptr; uint32_t *order_tmp = nullptr;
    char *zero_block = nullptr;         // counters | hierarchy flags | sort scratch: cleared by one memset (lbvh_zero_block_bytes)
    uint32_t *climb = nullptr;
    int32_t *split_cnt = nullptr; int64_t *ref_off = nullptr; char *scan_scratch = nullptr, *refs = nullptr;     // sliver splitting
    cudaEvent_t e0 = nullptr, e1 = nullptr, es0 = nullptr, es1 = nullptr;
    ~CommitScratch()
    {
        dfree(keys_tmp); dfree(order_tmp); dfree(zero_block); dfree(climb);
        dfree(split_cnt); dfree(ref_off); dfree(scan_scratch); dfree(refs);
        if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); if (es0) cudaEventDestroy(es0); if (es1) cudaEventDestroy(es1);
    }
};

int build_scene(qsmrt_scene *s, cudaStream_t st, CommitScratch &cs)
{
    const uint32_t G = (uint32_t)s->geoms.size();
    const uint64_t T = s->ntris, V = s->nverts;
    std::vector<uint64_t> goff(G + 1, 0), voff(G + 1, 0);
    { uint64_t t = 0, v = 0; for (uint32_t g = 0; g < G; ++g) { goff[g] = t; voff[g] = v; t += s->geoms[g].T; v += s->geoms[g].V; } goff[G] = t; voff[G] = v; }
    CUDA_TRY(cudaEventCreate(&cs.e0)); CUDA_TRY(cudaEventCreate(&cs.e1));
    CUDA_TRY(cudaEventCreate(&cs.es0)); CUDA_TRY(cudaEventCreate(&cs.es1));
    if (dmalloc(&s->goff, G + 1) || dmalloc(&s->voff, G + 1)) return 1;
    CUDA_TRY(cudaMemcpyAsync(s->goff, goff.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->voff, voff.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));        // goff / voff are locals
    // sliver splitting: the geometries counted their references at registration, so the number of leaves is known here
    uint64_t R = 0;
    for (Geometry &ge : s->geoms) {
        if (ge.ref_max != s->bopt.split_max || ge.ref_aspect != s->bopt.split_aspect) { if (count_refs(s, ge)) return 1; }
        R += ge.nrefs;
    }
    const bool split = R > T && R < (1ull << 29);
    const uint64_t Lv = split ? R : T;
    s->nleaves = Lv;
    // allocate everything before the timed region
    if (dmalloc(&s->keys, Lv) || dmalloc(&cs.keys_tmp, Lv) || dmalloc(&s->order, Lv) || dmalloc(&cs.order_tmp, Lv) ||
        dmalloc(&cs.zero_block, lbvh_zero_block_bytes(Lv)) ||
        dmalloc(&s->params, 1) || dmalloc(&s->bnodes, 2 * Lv - 1) || dmalloc(&s->tris, Lv) ||
        dmalloc(&s->tnodes, std::max<uint64_t>(Lv - 1, 1)) || dmalloc(&s->qnodes, std::max<uint64_t>(Lv - 1, 1)) ||
        dmalloc(&cs.climb, lbvh_climb_bytes(Lv) / sizeof(uint32_t)))
        return 1;
    unsigned long long *const counters = reinterpret_cast<unsigned long long *>(cs.zero_block);
    unsigned long long *const flags = counters + 8;
    uint32_t *const sort_scratch = reinterpret_cast<uint32_t *>(flags + (Lv > 1 ? Lv - 1 : 0));
    if (split && (dmalloc(&cs.split_cnt, T) || dmalloc(&cs.ref_off, T + 1) || dmalloc(&cs.scan_scratch, trv_scan_scratch_bytes(T)) ||
                  dmalloc(&cs.refs, lbvh_ref_bytes(R))))
        return 1;
    if (G == 1) { s->verts = s->geoms[0].verts; s->idx = s->geoms[0].idx; s->own_concat = false; }
    else {
        if (dmalloc(&s->verts, 3 * V) || dmalloc(&s->idx, 3 * T)) return 1;
        s->own_concat = true;
    }

    CUDA_TRY(cudaEventRecord(cs.e0, st));
    if (G > 1) {
        for (uint32_t g = 0; g < G; ++g) {
            const Geometry &ge = s->geoms[g];
            if (ge.V) CUDA_TRY(cudaMemcpyAsync(s->verts + 3 * voff[g], ge.verts, 3 * ge.V * sizeof(float), cudaMemcpyDeviceToDevice, st));
            if (ge.T) k_rebase_idx<<<(unsigned)((3 * ge.T + 255) / 256), 256, 0, st>>>(ge.idx, s->idx + 3 * goff[g], 3 * ge.T, (uint32_t)voff[g]);
        }
    }
    // scene bounds = union of the geometries' registration-time bounds (only geometries with triangles count)
    float slo[3] = { INFINITY, INFINITY, INFINITY }, shi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (const Geometry &ge : s->geoms)
        if (ge.T) for (int a = 0; a < 3; ++a) { slo[a] = fminf(slo[a], ge.lo[a]); shi[a] = fmaxf(shi[a], ge.hi[a]); }
    BuildParams bp_host;
    lbvh_finalize_params(slo, shi, &bp_host);
    if (split) {        // per-triangle slab counts -> offsets -> the references' boxes
        if (lbvh_split_count(s->verts, s->idx, T, s->bopt.split_max, s->bopt.split_aspect, cs.split_cnt, counters + 6, st) ||
            trv_exclusive_scan(cs.split_cnt, T, cs.ref_off, cs.scan_scratch, st) ||
            lbvh_split_emit(s->verts, s->idx, T, s->bopt.split_max, s->bopt.split_aspect, cs.ref_off, cs.refs, st))
            return 1;
    }
    int in_tmp = 0;
    LbvhBuildArgs A{};
    A.refs = split ? cs.refs : nullptr; A.nrefs = Lv;
    A.params_host = &bp_host; A.result_in_tmp = &in_tmp;
    A.leaf_max = s->bopt.leaf_max; A.sort_variant = s->bopt.sort_variant; A.climb_capacity = s->bopt.climb_capacity;
    A.quant_frac = s->bopt.quant_frac;
    A.verts = s->verts; A.idx = s->idx; A.ntris = T; A.geom_offsets = s->goff; A.ngeoms = G;
    A.params = s->params; A.keys = s->keys; A.keys_tmp = cs.keys_tmp;
    A.order = s->order; A.order_tmp = cs.order_tmp; A.sort_scratch = sort_scratch;
    A.bnodes = s->bnodes; A.flags = flags; A.keep_bnodes = s->bopt.keep_bnodes ? 1 : 0; A.climb_work = cs.climb;
    A.qnodes = s->qnodes;
    A.tris = s->tris; A.tnodes = s->tnodes; A.counters = counters; A.ev_sort0 = cs.es0; A.ev_sort1 = cs.es1;
    unsigned long long cnt[5] = {};
    // The sort normally runs over the top 40 key bits plus an exact fix-up of short runs; a scene with a run of more
    // than 64 triangles in one 2^-13 cell (thousands of coincident triangles) reports an overflow and is built again
    // with all eight passes.  Both attempts are inside the timed region.
    for (int attempt = 0; attempt < 2; ++attempt) {
        A.full_sort = attempt;
        if (lbvh_build(A, st)) return 1;
        CUDA_TRY(cudaEventRecord(cs.e1, st));
        CUDA_TRY(cudaEventSynchronize(cs.e1));
        CUDA_TRY(cudaMemcpy(cnt, counters, sizeof(cnt), cudaMemcpyDeviceToHost));
        if (!cnt[4]) break;
    }
    s->stats.full_sort = (uint32_t)A.full_sort;
    if (in_tmp) { std::swap(s->keys, cs.keys_tmp); std::swap(s->order, cs.order_tmp); }     // keep the pair that holds the sorted arrays
    CUDA_TRY(cudaEventElapsedTime(&s->stats.build_ms, cs.e0, cs.e1));
    CUDA_TRY(cudaEventElapsedTime(&s->stats.sort_ms, cs.es0, cs.es1));
    BuildParams bp;
    CUDA_TRY(cudaMemcpy(&bp, s->params, sizeof(bp), cudaMemcpyDeviceToHost));
    for (int a = 0; a < 3; ++a) { s->stats.scene_lo[a] = bp.slo[a]; s->stats.scene_hi[a] = bp.shi[a]; }
    s->stats.box_pad = bp.pad;
    s->stats.num_bvh_nodes = cnt[0]; s->stats.num_bvh_leaves = cnt[1]; s->stats.bvh_height = (uint32_t)cnt[2];
    s->use_qnodes = bp.use_q != 0;                      // decided on the device (k_hierarchy_refit_emit)
    if (!s->use_qnodes) dfree(s->qnodes);
    if (!s->bopt.keep_bnodes) dfree(s->bnodes);         // only the hand-over boxes of the build were in it
    for (int a = 0; a < 3; ++a) { s->glo[a] = bp.glo[a]; s->cell[a] = bp.cell[a]; }
    s->stats.quantised_nodes = s->use_qnodes ? 1u : 0u;
    s->stats.bvh_bytes = cnt[0] * (s->use_qnodes ? sizeof(QNode) : sizeof(TNode)) + Lv * sizeof(TriRec);
    s->stats.num_references = Lv;
    return 0;
}

int do_commit(qsmrt_scene *s, cudaStream_t st, float *build_ms_out)
{
    if (s->committed) { if (build_ms_out) *build_ms_out = s->stats.build_ms; return 0; }
    SyncedFrees batch;          // covers the re-commit teardown here and, after the build's own event wait, the scratch frees
    free_build(s);
    uint64_t T = 0, V = 0;
    for (const Geometry &g : s->geoms) { T += g.T; V += g.V; }
    if (T >= (1ull << 29)) FAIL("scene has %llu triangles; the leaf encoding holds 2^29", (unsigned long long)T);
    if (V >= (1ull << 32)) FAIL("scene has %llu vertices; indices are 32-bit", (unsigned long long)V);
    s->ntris = T; s->nverts = V;
    memset(&s->stats, 0, sizeof(s->stats));
    s->stats.num_triangles = T; s->stats.num_geometries = s->geoms.size(); s->stats.leaf_max = (uint32_t)s->bopt.leaf_max;
    if (T) {
        // `committed` is only set after a complete build: a failed one (out of memory on a large scene, say) frees
        // whatever it had allocated and leaves the scene uncommitted, so the next query builds again instead of
        // traversing null or half-written nodes
        CommitScratch cs;
        if (build_scene(s, st, cs)) {
            cudaDeviceSynchronize();            // kernels of the failed build may still be in flight
            cudaGetLastError();
            free_build(s);
            return 1;
        }
    }
    s->committed = true;
    if (build_ms_out) *build_ms_out = s->stats.build_ms;
    return 0;
}

int ensure_pipe(qsmrt_scene *s, uint64_t chunk, size_t out_bytes)
{
    HostPipe &hp = s->pipe;
    if (hp.chunk >= chunk && hp.out_bytes >= out_bytes) return 0;
    chunk = std::max(chunk, hp.chunk); out_bytes = std::max(out_bytes, hp.out_bytes);
    SyncedFrees batch;
    free_pipe(hp);
    for (int b = 0; b < HostPipe::NBUF; ++b) {
        if (dmalloc(&hp.rays[b], 6 * chunk) || dmalloc(&hp.out[b], chunk * out_bytes)) return 1;
        CUDA_TRY(cudaEventCreateWithFlags(&hp.e_in[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&hp.e_run[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&hp.e_out[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_run, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    hp.chunk = chunk; hp.out_bytes = out_bytes;
    return 0;
}

// One result array of a *_host call: `bytes` per ray, copied back to `host` (NULL: not wanted).  When `keep` is set
// the kernel writes the array for the whole batch there (device memory) instead: the caller fetches it later or never.
struct HostOut { void *host; void *keep; size_t bytes; };

// Host rays in, host results out: the batch is cut into chunks and host->device copy, traversal and device->host
// copy of consecutive chunks overlap on three streams (triple-buffered device staging).  `launch(rays_dev, n, dst[],
// stream)` enqueues the traversal of one chunk; dst[k] is where output k of that chunk goes.
template <int NOUT, class Launch>
int run_host_pipe(qsmrt_scene *s, const float *rays, uint64_t N, const HostOut (&outs)[NOUT], Launch launch)
{
    size_t stage_bytes = 0, back_bytes = 0, off_of[NOUT];
    for (int k = 0; k < NOUT; ++k) {
        off_of[k] = stage_bytes;
        if (outs[k].host && !outs[k].keep) { stage_bytes += (outs[k].bytes + 15) & ~(size_t)15; back_bytes += outs[k].bytes; }
    }
    // rays per pipeline stage (QSMRT_OPT_HOST_CHUNK overrides): 1M when the results are the larger transfer, 2M when the
    // rays are (measured on C2, profiles/r02_tuning.txt: 1.39 Grays/s with all five results, 1.96 -> 2.11 with t_hit + ids)
    const uint64_t chunk_rays = s->host_chunk ? s->host_chunk : back_bytes > 24 ? 1ull << 20 : 2ull << 20;
    const bool ramp = s->host_ramp != 0;
    const uint64_t chunk = std::min<uint64_t>(N, chunk_rays);
    if (ensure_pipe(s, chunk, std::max<size_t>(stage_bytes, 16))) return 1;
    HostPipe &hp = s->pipe;
    // Stage sizes.  The first copy-in and the last copy-out are the only transfers nothing overlaps, so a long batch
    // starts and ends with short stages (chunk/8, /4, /2) and runs full stages in between.
    std::vector<uint64_t> sizes;
    if (ramp && N >= 4 * chunk && chunk >= (1u << 16)) {
        uint64_t used = 0;
        for (uint64_t c = chunk / 8; c < chunk; c *= 2) { sizes.push_back(c); used += 2 * c; }
        const size_t nramp = sizes.size();
        for (uint64_t left = N - used; left; ) { const uint64_t n = std::min(chunk, left); sizes.push_back(n); left -= n; }
        for (size_t k = nramp; k-- > 0; ) sizes.push_back(sizes[k]);
    } else {
        for (uint64_t left = N; left; ) { const uint64_t n = std::min(chunk, left); sizes.push_back(n); left -= n; }
    }
    uint64_t off = 0;
    for (size_t c = 0; c < sizes.size(); off += sizes[c], ++c) {
        const int b = (int)(c % HostPipe::NBUF);
        const uint64_t n = sizes[c];
        if (c >= HostPipe::NBUF) CUDA_TRY(cudaStreamWaitEvent(hp.s_in, hp.e_out[b], 0));   // buffer drained
        CUDA_TRY(cudaMemcpyAsync(hp.rays[b], rays + 6 * off, 6 * n * sizeof(float), cudaMemcpyHostToDevice, hp.s_in));
        CUDA_TRY(cudaEventRecord(hp.e_in[b], hp.s_in));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_run, hp.e_in[b], 0));
        void *dst[NOUT];
        for (int k = 0; k < NOUT; ++k)
            dst[k] = outs[k].keep ? static_cast<char *>(outs[k].keep) + off * outs[k].bytes
                   : outs[k].host ? hp.out[b] + off_of[k] * hp.chunk : nullptr;
        if (launch(hp.rays[b], n, dst, hp.s_run)) return 1;
        CUDA_TRY(cudaEventRecord(hp.e_run[b], hp.s_run));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_out, hp.e_run[b], 0));
        for (int k = 0; k < NOUT; ++k)
            if (outs[k].host && !outs[k].keep)
                CUDA_TRY(cudaMemcpyAsync(static_cast<char *>(outs[k].host) + off * outs[k].bytes, dst[k], n * outs[k].bytes, cudaMemcpyDeviceToHost, hp.s_out));
        CUDA_TRY(cudaEventRecord(hp.e_out[b], hp.s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(hp.s_out));
    CUDA_TRY(cudaStreamSynchronize(hp.s_run));
    CUDA_TRY(cudaStreamSynchronize(hp.s_in));
    return 0;
}

} // namespace

extern "C" {

const char *qsmrt_last_error(void) { return g_err; }
int qsmrt_abi_version(void) { return QSMRT_ABI_VERSION; }

int qsmrt_scene_create(int cuda_device, qsmrt_scene **out)
{
    if (!out) FAIL("null output pointer");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        FAIL("no CUDA device (%s); libqsmrt has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (cuda_device < 0 || cuda_device >= ndev) FAIL("cuda_device %d out of range (0..%d)", cuda_device, ndev - 1);
    DeviceGuard guard(cuda_device);
    if (guard.err != cudaSuccess) FAIL("cudaSetDevice(%d): %s", cuda_device, cudaGetErrorString(guard.err));
    int major = 0, minor = 0;       // two attribute reads, not cudaGetDeviceProperties (2.4 ms per new scene)
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cuda_device));
    CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, cuda_device));
    if (major != 10) FAIL("device %d is sm_%d%d; libqsmrt is built for sm_100a only", cuda_device, major, minor);
    qsmrt_scene *s = new qsmrt_scene();
    s->device = cuda_device;
    s->trv.device = cuda_device;
    *out = s;
    return 0;
}

int qsmrt_scene_destroy(qsmrt_scene *s)
{
    if (!s) return 0;
    DeviceGuard guard(s->device);
    SyncedFrees batch;
    free_build(s);
    free_pipe(s->pipe);
    dfree(s->sweep_dev);
    dfree(s->sky.keys); dfree(s->sky.keys_tmp); dfree(s->sky.vals); dfree(s->sky.vals_tmp); dfree(s->sky.scratch);
    trv_state_free(s->trv);
    for (Geometry &g : s->geoms) { dfree(g.verts); dfree(g.idx); }
    delete s;
    return 0;
}

int qsmrt_add_triangles(qsmrt_scene *s, const float *verts, uint64_t V, const uint32_t *idx, uint64_t T,
                        int on_device, uint32_t *geom_id_out)
{
    SCENE_ENTER(s);
    if ((V && !verts) || (T && !idx)) FAIL("null vertex or index pointer");
    if (V >= (1ull << 32)) FAIL("too many vertices");
    Geometry g;
    g.V = V; g.T = T;
    uint32_t *d_max = nullptr;
    uint32_t maxi = 0;
    // copies and the index check run on the legacy default stream, which orders them after work already enqueued on
    // blocking streams; callers that produce the mesh on a non-blocking stream synchronise it first (the Python
    // front end does)
    auto body = [&]() -> int {
        if (dmalloc(&g.verts, 3 * V) || dmalloc(&g.idx, 3 * T)) return 1;
        const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (V) CUDA_TRY(cudaMemcpy(g.verts, verts, 3 * V * sizeof(float), kind));
        if (T) CUDA_TRY(cudaMemcpy(g.idx, idx, 3 * T * sizeof(uint32_t), kind));
        if (T) {
            // one pass over the index array: the largest index (Embree would read out of bounds; reject instead,
            // SURVEY.md 8b) and the bounds of the referenced vertices, which the commit only has to combine
            if (dmalloc(&d_max, 8)) return 1;
            if (lbvh_geometry_stats(g.verts, V, g.idx, T, d_max, g.lo, g.hi, &maxi, nullptr)) return 1;
            if (maxi >= V) { qsmrt_set_error("triangle index %u out of range (%llu vertices)", maxi, (unsigned long long)V); return 1; }
        }
        return 0;
    };
    const int rc = body() || count_refs(s, g);
    dfree(d_max);
    if (rc) { dfree(g.verts); dfree(g.idx); return 1; }
    if (s->committed || s->verts) free_build(s);
    s->geoms.push_back(g);
    if (geom_id_out) *geom_id_out = (uint32_t)(s->geoms.size() - 1);
    return 0;
}

int qsmrt_add_cylinders(qsmrt_scene *s, const float *records, uint64_t n, uint32_t resolution, uint32_t split,
                        int on_device, uint32_t *geom_id_out)
{
    SCENE_ENTER(s);
    if (n && !records) FAIL("null records pointer");
    if (resolution < 3 || split < 1 || resolution > 4096 || split > 4096) FAIL("resolution must be >= 3 and split >= 1");
    const uint64_t V = (uint64_t)resolution * (split + 1) + 2, T = 2ull * resolution * (1 + split);
    if (n * V >= (1ull << 32)) FAIL("too many cylinder vertices for 32-bit indices");
    Geometry g;
    g.V = n * V; g.T = n * T;
    float *rec_dev = nullptr;
    if (dmalloc(&g.verts, 3 * g.V) || dmalloc(&g.idx, 3 * g.T)) { dfree(g.verts); dfree(g.idx); return 1; }
    if (n) {
        const float *rec = records;
        if (!on_device) {
            if (dmalloc(&rec_dev, 8 * n)) { dfree(g.verts); dfree(g.idx); return 1; }
            if (cudaMemcpy(rec_dev, records, 8 * n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
                dfree(rec_dev); dfree(g.verts); dfree(g.idx);
                FAIL("cylinder records upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
            rec = rec_dev;
        }
        k_cylinders<<<(unsigned)n, 128>>>(rec, n, resolution, split, g.verts, g.idx);
        cudaError_t e = cudaDeviceSynchronize();
        dfree(rec_dev);
        uint32_t *d_stats = nullptr, maxi = 0;
        if (e == cudaSuccess && (dmalloc(&d_stats, 8) || lbvh_geometry_stats(g.verts, g.V, g.idx, g.T, d_stats, g.lo, g.hi, &maxi, nullptr)))
            e = cudaErrorUnknown;
        dfree(d_stats);
        if (e != cudaSuccess) { dfree(g.verts); dfree(g.idx); FAIL("cylinder generation failed: %s", cudaGetErrorString(e)); }
        if (count_refs(s, g)) { dfree(g.verts); dfree(g.idx); return 1; }
    }
    if (s->committed || s->verts) free_build(s);
    s->geoms.push_back(g);
    if (geom_id_out) *geom_id_out = (uint32_t)(s->geoms.size() - 1);
    return 0;
}

int qsmrt_geometry_size(qsmrt_scene *s, uint32_t geom_id, uint64_t *V_out, uint64_t *T_out)
{
    if (!s) FAIL("null scene");
    if (geom_id >= s->geoms.size()) FAIL("geometry id %u out of range", geom_id);
    if (V_out) *V_out = s->geoms[geom_id].V;
    if (T_out) *T_out = s->geoms[geom_id].T;
    return 0;
}

int qsmrt_copy_geometry(qsmrt_scene *s, uint32_t geom_id, float *verts_dev, uint32_t *idx_dev, void *stream)
{
    SCENE_ENTER(s);
    if (geom_id >= s->geoms.size()) FAIL("geometry id %u out of range", geom_id);
    const Geometry &g = s->geoms[geom_id];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (verts_dev && g.V) CUDA_TRY(cudaMemcpyAsync(verts_dev, g.verts, 3 * g.V * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (idx_dev && g.T) CUDA_TRY(cudaMemcpyAsync(idx_dev, g.idx, 3 * g.T * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int qsmrt_commit(qsmrt_scene *s, void *stream, float *build_ms_out)
{
    SCENE_ENTER(s);
    return do_commit(s, static_cast<cudaStream_t>(stream), build_ms_out);
}

int qsmrt_cast_rays(qsmrt_scene *s, const float *rays, uint64_t N, float *t_hit, uint32_t *geom, uint32_t *prim,
                    float *uv, float *nrm, void *stream)
{
    SCENE_ENTER(s);
    if (check_rays(rays, N)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    return trv_cast_rays(s->trv, view_of(s), rays, N, 0, t_hit, geom, prim, uv, nrm, st);
}

int qsmrt_cast_rays_2d(qsmrt_scene *s, const float *rays, uint32_t width, uint64_t height, float *t_hit, uint32_t *geom,
                       uint32_t *prim, float *uv, float *nrm, void *stream)
{
    SCENE_ENTER(s);
    if (check_rays(rays, (uint64_t)width * height)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    return trv_cast_rays(s->trv, view_of(s), rays, (uint64_t)width * height, width, t_hit, geom, prim, uv, nrm, st);
}

int qsmrt_scene_set_option(qsmrt_scene *s, int key, double value)
{
    if (!s) FAIL("null scene");
    const int iv = (int)value;
    BuildOptions &b = s->bopt; TrvOptions &t = s->trv.opt;
    bool rebuild = false;
    switch (key) {
    case QSMRT_OPT_LEAF_MAX:
        if (iv < 1 || iv > QSMRT_LEAF_MAX) FAIL("leaf_max must be in 1..%d", QSMRT_LEAF_MAX);
        rebuild = b.leaf_max != iv; b.leaf_max = iv; break;
    case QSMRT_OPT_KEEP_BINARY_NODES: rebuild = b.keep_bnodes != (iv != 0); b.keep_bnodes = iv != 0; break;
    case QSMRT_OPT_QUANT_THRESHOLD: { const float f = value > 0.0 ? (float)value : 0.15f; rebuild = b.quant_frac != f; b.quant_frac = f; break; }
    case QSMRT_OPT_CLIMB_CAPACITY: rebuild = b.climb_capacity != std::max(iv, 0); b.climb_capacity = std::max(iv, 0); break;
    case QSMRT_OPT_SORT_VARIANT:
        if (iv != 0 && iv != 1) FAIL("sort variant must be 0 (classic) or 1 (onesweep)");
        rebuild = b.sort_variant != iv; b.sort_variant = iv; break;
    case QSMRT_OPT_SPLIT_MAX:
        if (iv < 1 || iv > 64) FAIL("split_max must be in 1..64");
        rebuild = b.split_max != iv; b.split_max = iv; break;
    case QSMRT_OPT_SPLIT_ASPECT: { const float f = value > 0.0 ? (float)value : 2.0f; rebuild = b.split_aspect != f; b.split_aspect = f; break; }
    case QSMRT_OPT_QUANTISED_NODES: b.allow_qnodes = iv != 0; break;
    case QSMRT_OPT_TRAVERSAL_VARIANT:
        if (iv != 1 && iv != 2) FAIL("unknown traversal variant %d (1 = per-thread loop, 2 = persistent kernel)", iv);
        t.variant = iv; break;
    case QSMRT_OPT_REFILL: if (iv < 1 || iv > 32) FAIL("thresholds must be in 1..32"); t.refill = iv; break;
    case QSMRT_OPT_WANT: if (iv < 1 || iv > 32) FAIL("thresholds must be in 1..32"); t.want = iv; break;
    case QSMRT_OPT_TRI_MIN: if (iv < 1 || iv > 32) FAIL("thresholds must be in 1..32"); t.tri_min = iv; break;
    case QSMRT_OPT_COUNTERS: t.counters = iv != 0; break;
    case QSMRT_OPT_NODE_PATH: if (iv < 0 || iv > 2) FAIL("node path must be 0, 1 or 2"); t.node_path = iv; break;
    case QSMRT_OPT_CP_WARP_MAX: t.cp_warp_max = std::max(iv, 0); break;
    case QSMRT_OPT_CTAS_PER_SM: t.ctas_per_sm = std::max(iv, 0); break;
    case QSMRT_OPT_TILE_ORDER: t.tile_order = iv != 0; break;
    case QSMRT_OPT_POINT_ORDER: t.point_order = iv != 0; break;
    case QSMRT_OPT_HOST_CHUNK: if (iv != 0 && iv < 1024) FAIL("host_chunk must be 0 (automatic) or >= 1024 rays"); s->host_chunk = (uint64_t)iv; break;
    case QSMRT_OPT_HOST_RAMP: s->host_ramp = iv != 0; break;
    case QSMRT_OPT_COUNT_SET: if (iv < 4 || iv > 32) FAIL("count_set must be in 4..32"); t.count_set = iv; break;
    default: FAIL("unknown option %d", key);
    }
    if (rebuild && s->committed) { SCENE_ENTER(s); free_build(s); }     // the next query builds with the new option
    return 0;
}

int qsmrt_scene_get_option(qsmrt_scene *s, int key, double *value)
{
    if (!s || !value) FAIL("null pointer");
    const BuildOptions &b = s->bopt; const TrvOptions &t = s->trv.opt;
    switch (key) {
    case QSMRT_OPT_LEAF_MAX: *value = b.leaf_max; break;
    case QSMRT_OPT_KEEP_BINARY_NODES: *value = b.keep_bnodes; break;
    case QSMRT_OPT_QUANT_THRESHOLD: *value = b.quant_frac; break;
    case QSMRT_OPT_CLIMB_CAPACITY: *value = b.climb_capacity; break;
    case QSMRT_OPT_SORT_VARIANT: *value = b.sort_variant; break;
    case QSMRT_OPT_SPLIT_MAX: *value = b.split_max; break;
    case QSMRT_OPT_SPLIT_ASPECT: *value = b.split_aspect; break;
    case QSMRT_OPT_QUANTISED_NODES: *value = b.allow_qnodes; break;
    case QSMRT_OPT_TRAVERSAL_VARIANT: *value = t.variant; break;
    case QSMRT_OPT_REFILL: *value = t.refill; break;
    case QSMRT_OPT_WANT: *value = t.want; break;
    case QSMRT_OPT_TRI_MIN: *value = t.tri_min; break;
    case QSMRT_OPT_COUNTERS: *value = t.counters; break;
    case QSMRT_OPT_NODE_PATH: *value = t.node_path; break;
    case QSMRT_OPT_CP_WARP_MAX: *value = t.cp_warp_max; break;
    case QSMRT_OPT_CTAS_PER_SM: *value = t.ctas_per_sm; break;
    case QSMRT_OPT_TILE_ORDER: *value = t.tile_order; break;
    case QSMRT_OPT_POINT_ORDER: *value = t.point_order; break;
    case QSMRT_OPT_HOST_CHUNK: *value = (double)s->host_chunk; break;
    case QSMRT_OPT_HOST_RAMP: *value = s->host_ramp; break;
    case QSMRT_OPT_COUNT_SET: *value = t.count_set; break;
    default: FAIL("unknown option %d", key);
    }
    return 0;
}

int qsmrt_scene_get_counters(qsmrt_scene *s, uint64_t out[16])
{
    SCENE_ENTER(s);
    if (!out) FAIL("null pointer");
    unsigned long long h[16];
    if (trv_read_counters(s->trv, h)) return 1;
    for (int k = 0; k < 16; ++k) out[k] = h[k];
    return 0;
}

int qsmrt_util_read_sweep(const void *buf_dev, uint64_t bytes, uint32_t reps, uint32_t *sink_dev, void *stream)
{
    if (!buf_dev || !sink_dev) FAIL("null pointer");
    if (reinterpret_cast<uintptr_t>(buf_dev) & 31u) FAIL("buffer must be 32-byte aligned");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    k_read_sweep<<<(unsigned)(sms * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4 *>(buf_dev), bytes / 32, reps, sink_dev);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int qsmrt_release_cached_memory(void)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_blocks.m);
    g_blocks.release_all();
    return 0;
}

// results of cast_rays in ABI order: t_hit, geometry_ids, primitive_ids, primitive_uvs, primitive_normals
static const size_t CAST_BYTES[5] = { 4, 4, 4, 8, 12 };

int qsmrt_cast_rays_host_split(qsmrt_scene *s, const float *rays, uint64_t N, void *const host_out[5], void *const dev_out[5])
{
    SCENE_ENTER(s);
    if (N && !rays) FAIL("rays pointer is null");
    if (!host_out) FAIL("null output table");
    if (do_commit(s, nullptr, nullptr)) return 1;
    if (N == 0) return 0;
    HostOut outs[5];
    for (int k = 0; k < 5; ++k) {
        void *keep = dev_out ? dev_out[k] : nullptr;
        outs[k] = HostOut{ keep ? keep : host_out[k], keep, CAST_BYTES[k] };       // `host` non-null marks the output as wanted
    }
    if (dev_out && dev_out[3] && (reinterpret_cast<uintptr_t>(dev_out[3]) & 7u)) FAIL("primitive_uvs must be 8-byte aligned");
    const SceneView sv = view_of(s);
    return run_host_pipe<5>(s, rays, N, outs, [&](const float *r, uint64_t n, void **dst, cudaStream_t st) {
        return trv_cast_rays(s->trv, sv, r, n, 0, static_cast<float *>(dst[0]), static_cast<uint32_t *>(dst[1]),
                             static_cast<uint32_t *>(dst[2]), static_cast<float *>(dst[3]), static_cast<float *>(dst[4]), st);
    });
}

int qsmrt_cast_rays_host(qsmrt_scene *s, const float *rays, uint64_t N, float *t_hit, uint32_t *geom,
                         uint32_t *prim, float *uv, float *nrm)
{
    void *const host[5] = { t_hit, geom, prim, uv, nrm };
    return qsmrt_cast_rays_host_split(s, rays, N, host, nullptr);
}

int qsmrt_count_intersections_host(qsmrt_scene *s, const float *rays, uint64_t N, int32_t *counts)
{
    SCENE_ENTER(s);
    if (N && (!rays || !counts)) FAIL("null pointer");
    if (do_commit(s, nullptr, nullptr)) return 1;
    if (N == 0) return 0;
    const HostOut outs[1] = { HostOut{ counts, nullptr, sizeof(int32_t) } };
    const SceneView sv = view_of(s);
    const uint32_t ngeoms = (uint32_t)s->geoms.size();
    return run_host_pipe<1>(s, rays, N, outs, [&](const float *r, uint64_t n, void **dst, cudaStream_t st) {
        return trv_count(s->trv, sv, r, n, static_cast<int32_t *>(dst[0]), ngeoms, st);
    });
}

int qsmrt_test_occlusions_host(qsmrt_scene *s, const float *rays, uint64_t N, float tnear, float tfar, uint8_t *out)
{
    SCENE_ENTER(s);
    if (N && (!rays || !out)) FAIL("null pointer");
    if (do_commit(s, nullptr, nullptr)) return 1;
    if (N == 0) return 0;
    const HostOut outs[1] = { HostOut{ out, nullptr, sizeof(uint8_t) } };
    const SceneView sv = view_of(s);
    return run_host_pipe<1>(s, rays, N, outs, [&](const float *r, uint64_t n, void **dst, cudaStream_t st) {
        return trv_occluded(s->trv, sv, r, n, tnear, tfar, static_cast<uint8_t *>(dst[0]), st);
    });
}

int qsmrt_count_intersections(qsmrt_scene *s, const float *rays, uint64_t N, int32_t *counts, void *stream)
{
    SCENE_ENTER(s);
    if (check_rays(rays, N)) return 1;
    if (N && !counts) FAIL("counts pointer is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_count(s->trv, view_of(s), rays, N, counts, (uint32_t)s->geoms.size(), st);
}

int qsmrt_test_occlusions(qsmrt_scene *s, const float *rays, uint64_t N, float tnear, float tfar, uint8_t *out, void *stream)
{
    SCENE_ENTER(s);
    if (check_rays(rays, N)) return 1;
    if (N && !out) FAIL("output pointer is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_occluded(s->trv, view_of(s), rays, N, tnear, tfar, out, st);
}

int qsmrt_list_intersections_count(qsmrt_scene *s, const float *rays, uint64_t N, int64_t *ray_splits,
                                   int64_t *total_out, void *stream)
{
    SCENE_ENTER(s);
    if (check_rays(rays, N)) return 1;
    if (!ray_splits || !total_out) FAIL("null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    s->list_rays = nullptr; s->list_n = 0;
    *total_out = 0;
    if (s->ntris == 0 || N == 0) {
        CUDA_TRY(cudaMemsetAsync(ray_splits, 0, (N + 1) * sizeof(int64_t), st));
        CUDA_TRY(cudaStreamSynchronize(st));
        s->list_rays = rays; s->list_n = N;
        return 0;
    }
    // ONE all-hits traversal: the counts (count_intersections' distinct-(geometry, t) rule), scanned into the CSR
    // offsets, and the hit records themselves in a stash that _fill moves to the caller's arrays.  The stash is sized
    // for 4 hits per ray; a batch with more is traversed once more with exactly the room it needs.
    free_list_stash(s);
    int32_t *cnt = nullptr; char *scratch = nullptr;
    auto body = [&]() -> int {
        if (dmalloc(&cnt, N) || dmalloc(&scratch, trv_scan_scratch_bytes(N))) return 1;
        ListStash &ls = s->list_stash;
        unsigned long long cap = 4 * N + 65536, needed = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            needed = 0;
            if (dmalloc(&ls.base, N) || dmalloc(&ls.t, cap) || dmalloc(&ls.geom, cap) || dmalloc(&ls.prim, cap) ||
                dmalloc(&ls.uv, 2 * cap) || dmalloc(&ls.count, 1)) {
                free_list_stash(s);         // no room for a stash: count only, _fill enumerates the hits of every ray
                g_err[0] = 0;
            } else {
                ls.cap = cap;
                CUDA_TRY(cudaMemsetAsync(ls.count, 0, sizeof(unsigned long long), st));     // also read when no kernel used the stash
            }
            if (trv_list_collect(s->trv, view_of(s), rays, N, (uint32_t)s->geoms.size(), cnt, ls, &s->list_max_fast, st) ||
                trv_exclusive_scan(cnt, N, ray_splits, scratch, st)) return 1;
            CUDA_TRY(cudaMemcpyAsync(total_out, ray_splits + N, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
            if (ls.count) CUDA_TRY(cudaMemcpyAsync(&needed, ls.count, sizeof(needed), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            if (needed <= ls.cap) break;
            free_list_stash(s);
            cap = needed;
        }
        return 0;
    };
    const int rc = body();
    dfree(cnt); dfree(scratch);
    if (rc) { free_list_stash(s); return 1; }
    s->list_rays = rays; s->list_n = N;
    return 0;
}

int qsmrt_list_intersections_fill(qsmrt_scene *s, const float *rays, uint64_t N, const int64_t *ray_splits,
                                  int64_t *ray_ids, float *t_hit, uint32_t *geom, uint32_t *prim, float *uv, void *stream)
{
    SCENE_ENTER(s);
    if (!s->committed || s->list_rays != rays || s->list_n != N)
        FAIL("list_intersections_fill must follow list_intersections_count on the same rays");
    if (!ray_splits) FAIL("null ray_splits");
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    s->list_rays = nullptr; s->list_n = 0;
    if (s->ntris == 0 || N == 0) { free_list_stash(s); return 0; }
    // the output pass sorts each ray's records in the caller's t / geometry / primitive / uv arrays: outputs the
    // caller skipped are staged in scratch
    int64_t total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, ray_splits + N, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (total == 0) { free_list_stash(s); return 0; }
    float *t_tmp = nullptr, *uv_tmp = nullptr; uint32_t *g_tmp = nullptr, *p_tmp = nullptr;
    auto body = [&]() -> int {
        if ((!t_hit && dmalloc(&t_tmp, (uint64_t)total)) || (!geom && dmalloc(&g_tmp, (uint64_t)total)) ||
            (!prim && dmalloc(&p_tmp, (uint64_t)total)) || (!uv && dmalloc(&uv_tmp, 2 * (uint64_t)total))) return 1;
        if (trv_list_emit(view_of(s), rays, N, ray_splits, s->list_stash, s->list_max_fast, ray_ids, t_hit ? t_hit : t_tmp,
                          geom ? geom : g_tmp, prim ? prim : p_tmp, uv ? uv : uv_tmp, st)) return 1;
        CUDA_TRY(cudaStreamSynchronize(st));        // the stash (and any scratch) goes back to the block cache below
        return 0;
    };
    const int rc = body();
    dfree(t_tmp); dfree(g_tmp); dfree(p_tmp); dfree(uv_tmp);
    free_list_stash(s);
    return rc;
}

int qsmrt_gen_parallel_rays(float *rays, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                            const float dv[3], const float dir[3], void *stream)
{
    if (!rays || !o0 || !du || !dv || !dir) FAIL("null pointer");
    if (reinterpret_cast<uintptr_t>(rays) & 7u) FAIL("rays must be 8-byte aligned");
    return trv_gen_parallel(rays, nu, nv, o0, du, dv, dir, static_cast<cudaStream_t>(stream));
}

int qsmrt_gen_pinhole_rays(float *rays, uint32_t w, uint32_t h, const double K[9], const double E[16], void *stream)
{
    if (!rays || !K || !E) FAIL("null pointer");
    if (reinterpret_cast<uintptr_t>(rays) & 7u) FAIL("rays must be 8-byte aligned");
    // Open3D CreateRaysPinhole: C = -R^T t, direction = (K R)^-1 (x+.5, y+.5, 1)
    double R[9], t[3], M[9], inv[9], eye[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R[3 * r + c] = E[4 * r + c]; t[r] = E[4 * r + 3]; }
    for (int c = 0; c < 3; ++c) eye[c] = -(R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2]);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        double a = 0; for (int k = 0; k < 3; ++k) a += K[3 * r + k] * R[3 * k + c];
        M[3 * r + c] = a;
    }
    double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
    if (det == 0.0 || !std::isfinite(det)) FAIL("intrinsic * rotation is singular");
    inv[0] = (M[4] * M[8] - M[5] * M[7]) / det; inv[1] = (M[2] * M[7] - M[1] * M[8]) / det; inv[2] = (M[1] * M[5] - M[2] * M[4]) / det;
    inv[3] = (M[5] * M[6] - M[3] * M[8]) / det; inv[4] = (M[0] * M[8] - M[2] * M[6]) / det; inv[5] = (M[2] * M[3] - M[0] * M[5]) / det;
    inv[6] = (M[3] * M[7] - M[4] * M[6]) / det; inv[7] = (M[1] * M[6] - M[0] * M[7]) / det; inv[8] = (M[0] * M[4] - M[1] * M[3]) / det;
    return trv_gen_pinhole(rays, w, h, inv, eye, static_cast<cudaStream_t>(stream));
}

int qsmrt_mark_hit_primitives(qsmrt_scene *s, const uint32_t *geom, const uint32_t *prim, uint64_t N,
                              uint8_t *tri_hit, uint8_t *vert_hit, void *stream)
{
    SCENE_ENTER(s);
    if (!prim) FAIL("primitive_ids pointer is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (s->ntris == 0) return 0;
    return trv_mark_hits(geom, prim, N, s->goff, s->voff, (uint32_t)s->geoms.size(), s->idx, tri_hit, vert_hit, st);
}

int qsmrt_accumulate_hits(qsmrt_scene *s, const uint32_t *geom, const uint32_t *prim, uint64_t N,
                          uint32_t *tri_counts, void *stream)
{
    SCENE_ENTER(s);
    if (!prim || !tri_counts) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (s->ntris == 0) return 0;
    return trv_accumulate_hits(geom, prim, N, s->goff, (uint32_t)s->geoms.size(), tri_counts, st);
}

int qsmrt_closest_points(qsmrt_scene *s, const float *pts, uint64_t N, float *closest, float *dist, uint32_t *geom,
                         uint32_t *prim, float *uv, float *nrm, void *stream)
{
    SCENE_ENTER(s);
    if (N && !pts) FAIL("query points pointer is null");
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_closest_points(s->trv, view_of(s), pts, N, closest, dist, geom, prim, uv, nrm, st);
}

int qsmrt_signed_distance(qsmrt_scene *s, const float *pts, uint64_t N, float *dist, void *stream)
{
    SCENE_ENTER(s);
    if (N && (!pts || !dist)) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (N == 0) return 0;
    float *rays = nullptr; int32_t *counts = nullptr;
    if (dmalloc(&rays, 6 * N) || dmalloc(&counts, N)) { dfree(rays); dfree(counts); return 1; }
    int rc = trv_closest_points(s->trv, view_of(s), pts, N, nullptr, dist, nullptr, nullptr, nullptr, nullptr, st) ||
             trv_points_to_rays(pts, rays, N, st) ||
             trv_count(s->trv, view_of(s), rays, N, counts, (uint32_t)s->geoms.size(), st) ||
             trv_apply_sign(dist, counts, N, st);
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { qsmrt_set_error("signed_distance: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; }
    dfree(rays); dfree(counts);
    return rc;
}

int qsmrt_sun_exposure(qsmrt_scene *s, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                       const float dv[3], const float dir[3], uint32_t *tri_counts, void *stream)
{
    SCENE_ENTER(s);
    if (!o0 || !du || !dv || !dir || !tri_counts) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_sun_exposure(s->trv, view_of(s), nu, nv, o0, du, dv, dir, s->geoms.size() > 1 ? s->goff : nullptr, tri_counts, st);
}

int qsmrt_sun_exposure_sweep(qsmrt_scene *s, uint32_t n_grids, const float *grids, uint64_t nu, uint64_t nv,
                             uint32_t *tri_counts, uint64_t count_stride, void *stream)
{
    SCENE_ENTER(s);
    if (!grids || !tri_counts) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (n_grids == 0) return 0;
    if (s->sweep_cap < n_grids) {
        if (s->sweep_dev) { SyncedFrees batch; dfree(s->sweep_dev); }
        s->sweep_cap = 0;
        if (dmalloc(&s->sweep_dev, 12ull * std::max<uint32_t>(n_grids, 64))) return 1;
        s->sweep_cap = std::max<uint32_t>(n_grids, 64);
    }
    // `grids` is ordinary host memory: the runtime stages it before the call returns; the copy is ordered on `stream`
    // behind earlier sweeps of this scene, which read the same buffer
    CUDA_TRY(cudaMemcpyAsync(s->sweep_dev, grids, 12ull * n_grids * sizeof(float), cudaMemcpyHostToDevice, st));
    return trv_sun_exposure_sweep(s->trv, view_of(s), n_grids, s->sweep_dev, nu, nv, s->geoms.size() > 1 ? s->goff : nullptr,
                                  tri_counts, count_stride, st);
}

int qsmrt_sky_visibility(qsmrt_scene *s, const float *points, const float *normals, uint64_t n_points, uint64_t point_base,
                         uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count,
                         uint32_t *unoccluded, void *stream)
{
    SCENE_ENTER(s);
    if (!points || !unoccluded) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    // The points are WORKED in Morton order (sorted here with the builder's radix sort; results land at the points'
    // own indices): the rays in flight then start in one region of the scene and share their first nodes -- 2.41 ->
    // 2.64 Grays/s on C5, whose leaf vertices come in random spatial order (profiles/r02_tuning.txt).
    const uint32_t *perm = nullptr;
    if (s->trv.opt.point_order && n_points >= 4096 && n_points < (1ull << 32) && s->ntris) {
        qsmrt_scene::SkyOrder &so = s->sky;
        if (so.cap < n_points) {
            SyncedFrees batch;
            dfree(so.keys); dfree(so.keys_tmp); dfree(so.vals); dfree(so.vals_tmp); dfree(so.scratch); so.cap = 0;
            if (dmalloc(&so.keys, n_points) || dmalloc(&so.keys_tmp, n_points) || dmalloc(&so.vals, n_points) || dmalloc(&so.vals_tmp, n_points) ||
                dmalloc_bytes(reinterpret_cast<void **>(&so.scratch), lbvh_sort_scratch_bytes(n_points))) return 1;
            so.cap = n_points;
        }
        int in_tmp = 0;
        if (trv_point_keys(points, n_points, s->stats.scene_lo, s->stats.scene_hi, so.keys, so.vals, st) ||
            lbvh_radix_sort(so.keys, so.keys_tmp, so.vals, so.vals_tmp, n_points, so.scratch, st, nullptr, 0, false, &in_tmp)) return 1;
        perm = in_tmp ? so.vals_tmp : so.vals;
    }
    return trv_sky_visibility(s->trv, view_of(s), points, normals, n_points, point_base, seed, offset, dir_begin, dir_count, perm, unoccluded, st);
}

int qsmrt_gen_hemisphere_rays(float *rays, const float *points, const float *normals, uint64_t n_points, uint64_t point_base,
                              uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, void *stream)
{
    if (!rays || !points) FAIL("null pointer");
    if (reinterpret_cast<uintptr_t>(rays) & 7u) FAIL("rays must be 8-byte aligned");
    return trv_gen_hemisphere(rays, points, normals, n_points, point_base, seed, offset, dir_begin, dir_count, static_cast<cudaStream_t>(stream));
}

int qsmrt_peel_projection(qsmrt_scene *s, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                          const float dv[3], const float dir[3], int max_layers, int32_t *layer_of,
                          double *layer_stats, int *n_layers_out, void *stream)
{
    SCENE_ENTER(s);
    if (!o0 || !du || !dv || !dir || !layer_stats || !n_layers_out || max_layers < 1) FAIL("bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    *n_layers_out = 0;
    const uint64_t T = s->ntris;
    if (T == 0) return 0;
    uint8_t *alive = nullptr, *hit = nullptr; double *sums = nullptr;
    if (dmalloc(&alive, T) || dmalloc(&hit, T) || dmalloc(&sums, 3)) { dfree(alive); dfree(hit); dfree(sums); return 1; }
    int rc = 0;
    if (cudaMemsetAsync(alive, 1, T, st) != cudaSuccess || cudaMemsetAsync(hit, 0, T, st) != cudaSuccess ||
        (layer_of && cudaMemsetAsync(layer_of, 0xFF, T * sizeof(int32_t), st) != cudaSuccess)) rc = 1;
    SceneView sv = view_of(s);
    for (int layer = 0; !rc && layer < max_layers; ++layer) {
        double h[3] = { 0, 0, 0 };
        if (cudaMemsetAsync(sums, 0, 3 * sizeof(double), st) != cudaSuccess) { rc = 1; break; }
        rc = trv_peel_cast(s->trv, sv, nu, nv, o0, du, dv, dir, s->order, alive, hit, st) ||
             trv_peel_update(s->verts, s->idx, T, alive, hit, layer_of, layer, dir, sums, st);
        if (!rc && (cudaMemcpyAsync(h, sums, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                    cudaStreamSynchronize(st) != cudaSuccess)) rc = 1;
        if (rc || h[0] == 0.0) break;                       // nothing left that the rays can see
        layer_stats[3 * layer] = h[0]; layer_stats[3 * layer + 1] = h[1]; layer_stats[3 * layer + 2] = h[2];
        *n_layers_out = layer + 1;
    }
    if (rc && !g_err[0]) qsmrt_set_error("peel_projection: %s", cudaGetErrorString(cudaGetLastError()));
    dfree(alive); dfree(hit); dfree(sums);
    return rc;
}

// ---- scene files --------------------------------------------------------------------------------------------
// The reference's idiom for built search structures is pickling them next to the data (pyQSM/utils/io.py:44-60,
// tree_isolation.py:114,136).  Here: the geometries as added plus, with QSMRT_SAVE_BVH, the committed LBVH
// (traversal nodes, quantised twin, triangle records, order, keys, grid).  A file without the BVH is rebuilt on the
// first query after loading -- the build is deterministic, so both routes give the identical tree.
namespace {
struct SceneFileHeader {
    char     magic[8];           // "QSMRTSC1"
    uint32_t version, flags;
    uint64_t ngeoms, ntris, nverts, nleaves;
    int32_t  leaf_max, use_q, split_max; float split_aspect;
    float    quant_frac, glo[3], cell[3];
    qsmrt_stats stats;
};
constexpr size_t IO_CHUNK = 64u << 20;

struct FileIO {
    FILE *f = nullptr; void *bounce = nullptr;
    ~FileIO() { if (f) fclose(f); if (bounce) cudaFreeHost(bounce); }
    int open(const char *path, const char *mode)
    {
        f = fopen(path, mode);
        if (!f) { qsmrt_set_error("cannot open %s", path); return 1; }
        CUDA_TRY(cudaMallocHost(&bounce, IO_CHUNK));
        return 0;
    }
    int put(const void *dev, size_t bytes)
    {
        for (size_t o = 0; o < bytes; o += IO_CHUNK) {
            const size_t n = std::min(IO_CHUNK, bytes - o);
            CUDA_TRY(cudaMemcpy(bounce, static_cast<const char *>(dev) + o, n, cudaMemcpyDeviceToHost));
            if (fwrite(bounce, 1, n, f) != n) { qsmrt_set_error("short write"); return 1; }
        }
        return 0;
    }
    int get(void *dev, size_t bytes)
    {
        for (size_t o = 0; o < bytes; o += IO_CHUNK) {
            const size_t n = std::min(IO_CHUNK, bytes - o);
            if (fread(bounce, 1, n, f) != n) { qsmrt_set_error("scene file is truncated"); return 1; }
            CUDA_TRY(cudaMemcpy(static_cast<char *>(dev) + o, bounce, n, cudaMemcpyHostToDevice));
        }
        return 0;
    }
};
} // namespace

int qsmrt_scene_save(qsmrt_scene *s, const char *path, uint32_t flags)
{
    SCENE_ENTER(s);
    if (!path) FAIL("null path");
    const bool with_bvh = (flags & QSMRT_SAVE_BVH) != 0;
    if (with_bvh && do_commit(s, nullptr, nullptr)) return 1;
    CUDA_TRY(cudaDeviceSynchronize());
    FileIO io;
    if (io.open(path, "wb")) return 1;
    SceneFileHeader h{};
    memcpy(h.magic, "QSMRTSC1", 8);
    h.version = 1; h.flags = (with_bvh && s->ntris) ? QSMRT_SAVE_BVH : 0u;
    h.ngeoms = s->geoms.size();
    for (const Geometry &g : s->geoms) { h.ntris += g.T; h.nverts += g.V; }
    h.leaf_max = s->bopt.leaf_max; h.quant_frac = s->bopt.quant_frac;
    h.split_max = s->bopt.split_max; h.split_aspect = s->bopt.split_aspect; h.nleaves = h.flags ? s->nleaves : 0;
    if (h.flags) {
        h.use_q = s->use_qnodes ? 1 : 0; h.stats = s->stats;
        for (int a = 0; a < 3; ++a) { h.glo[a] = s->glo[a]; h.cell[a] = s->cell[a]; }
    }
    if (fwrite(&h, sizeof(h), 1, io.f) != 1) FAIL("short write");
    for (const Geometry &g : s->geoms) {
        const uint64_t vt[2] = { g.V, g.T };
        if (fwrite(vt, sizeof(vt), 1, io.f) != 1) FAIL("short write");
        if (io.put(g.verts, 3 * g.V * sizeof(float)) || io.put(g.idx, 3 * g.T * sizeof(uint32_t))) return 1;
    }
    if (h.flags) {
        const uint64_t T = s->nleaves, NN = std::max<uint64_t>(T - 1, 1);
        if (io.put(s->params, sizeof(BuildParams)) || io.put(s->tnodes, NN * sizeof(TNode)) ||
            (s->use_qnodes && io.put(s->qnodes, NN * sizeof(QNode))) || io.put(s->tris, T * sizeof(TriRec)) ||
            io.put(s->order, T * sizeof(uint32_t)) || io.put(s->keys, T * sizeof(uint64_t)))
            return 1;
    }
    return 0;
}

int qsmrt_scene_load(int cuda_device, const char *path, qsmrt_scene **out)
{
    if (!out || !path) FAIL("null pointer");
    *out = nullptr;
    qsmrt_scene *s = nullptr;
    if (qsmrt_scene_create(cuda_device, &s)) return 1;
    auto body = [&]() -> int {
        SCENE_ENTER(s);
        FileIO io;
        if (io.open(path, "rb")) return 1;
        SceneFileHeader h;
        if (fread(&h, sizeof(h), 1, io.f) != 1 || memcmp(h.magic, "QSMRTSC1", 8) != 0 || h.version != 1)
            FAIL("%s is not a qsmrt scene file (version 1)", path);
        if (h.ntris >= (1ull << 29) || h.nleaves >= (1ull << 29) || h.nverts >= (1ull << 32) || h.leaf_max < 1 || h.leaf_max > QSMRT_LEAF_MAX ||
            h.split_max < 1 || h.split_max > 64 || !(h.split_aspect > 0.0f)) FAIL("corrupt scene file header");
        s->bopt.split_max = h.split_max; s->bopt.split_aspect = h.split_aspect;      // before the geometries count their references
        uint64_t T = 0, V = 0;
        for (uint64_t gi = 0; gi < h.ngeoms; ++gi) {
            uint64_t vt[2];
            if (fread(vt, sizeof(vt), 1, io.f) != 1) FAIL("scene file is truncated");
            if (vt[0] > h.nverts || vt[1] > h.ntris) FAIL("corrupt scene file");
            Geometry g; g.V = vt[0]; g.T = vt[1];
            if (dmalloc(&g.verts, 3 * g.V) || dmalloc(&g.idx, 3 * g.T) ||
                io.get(g.verts, 3 * g.V * sizeof(float)) || io.get(g.idx, 3 * g.T * sizeof(uint32_t))) { dfree(g.verts); dfree(g.idx); return 1; }
            uint32_t *d_stats = nullptr, maxi = 0;
            if (g.T && (dmalloc(&d_stats, 8) || lbvh_geometry_stats(g.verts, g.V, g.idx, g.T, d_stats, g.lo, g.hi, &maxi, nullptr) || maxi >= g.V)) {
                dfree(d_stats); dfree(g.verts); dfree(g.idx);
                FAIL("corrupt scene file (geometry %llu)", (unsigned long long)gi);
            }
            dfree(d_stats);
            if (count_refs(s, g)) { dfree(g.verts); dfree(g.idx); return 1; }
            s->geoms.push_back(g);
            T += g.T; V += g.V;
        }
        if (T != h.ntris || V != h.nverts) FAIL("corrupt scene file (geometry sizes)");
        s->bopt.leaf_max = h.leaf_max; s->bopt.quant_frac = h.quant_frac;
        if (!(h.flags & QSMRT_SAVE_BVH) || T == 0) return 0;        // geometry only: the first query builds
        const uint32_t G = (uint32_t)s->geoms.size();
        const uint64_t Lv = h.nleaves;
        if (Lv < T) FAIL("corrupt scene file (leaf count)");
        const uint64_t NN = std::max<uint64_t>(Lv - 1, 1);
        s->ntris = T; s->nverts = V; s->nleaves = Lv;
        std::vector<uint64_t> goff(G + 1, 0), voff(G + 1, 0);
        { uint64_t t = 0, v = 0; for (uint32_t g = 0; g < G; ++g) { goff[g] = t; voff[g] = v; t += s->geoms[g].T; v += s->geoms[g].V; } goff[G] = t; voff[G] = v; }
        if (dmalloc(&s->goff, G + 1) || dmalloc(&s->voff, G + 1) || dmalloc(&s->params, 1) || dmalloc(&s->tnodes, NN) ||
            (h.use_q && dmalloc(&s->qnodes, NN)) || dmalloc(&s->tris, Lv) || dmalloc(&s->order, Lv) || dmalloc(&s->keys, Lv))
            return 1;
        CUDA_TRY(cudaMemcpy(s->goff, goff.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(s->voff, voff.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
        if (G == 1) { s->verts = s->geoms[0].verts; s->idx = s->geoms[0].idx; s->own_concat = false; }
        else {
            if (dmalloc(&s->verts, 3 * V) || dmalloc(&s->idx, 3 * T)) return 1;
            s->own_concat = true;
            for (uint32_t g = 0; g < G; ++g) {
                const Geometry &ge = s->geoms[g];
                if (ge.V) CUDA_TRY(cudaMemcpy(s->verts + 3 * voff[g], ge.verts, 3 * ge.V * sizeof(float), cudaMemcpyDeviceToDevice));
                if (ge.T) k_rebase_idx<<<(unsigned)((3 * ge.T + 255) / 256), 256>>>(ge.idx, s->idx + 3 * goff[g], 3 * ge.T, (uint32_t)voff[g]);
            }
        }
        if (io.get(s->params, sizeof(BuildParams)) || io.get(s->tnodes, NN * sizeof(TNode)) ||
            (h.use_q && io.get(s->qnodes, NN * sizeof(QNode))) || io.get(s->tris, Lv * sizeof(TriRec)) ||
            io.get(s->order, Lv * sizeof(uint32_t)) || io.get(s->keys, Lv * sizeof(uint64_t)))
            return 1;
        // range check of everything the traversal follows (a file cut short or belonging to other geometry fails here,
        // not inside a kernel), one tree level per launch; the stored height sizes the traversal stack
        if (h.stats.bvh_height > 4096u || h.stats.num_triangles != T) FAIL("corrupt scene file (statistics)");
        {
            int *front[2] = { nullptr, nullptr };
            uint32_t *d_cnt = nullptr, h_cnt[3] = { 0, 0, 0 };                  // next level's size, bad flag, spare
            const QNode *qn = h.use_q ? s->qnodes : nullptr;
            int rc = dmalloc(&front[0], NN) || dmalloc(&front[1], NN) || dmalloc(&d_cnt, 4);
            uint64_t visited = 1;
            uint32_t count = 1, levels = 0;
            if (!rc) rc = cudaMemset(front[0], 0, sizeof(int)) != cudaSuccess || cudaMemset(d_cnt, 0, 4 * sizeof(uint32_t)) != cudaSuccess;
            if (!rc) {
                k_validate_leaves<<<(unsigned)((Lv + 255) / 256), 256>>>(s->tris, s->order, Lv, T, s->goff, G, d_cnt + 1);
                while (count && !h_cnt[1] && visited <= NN && levels <= h.stats.bvh_height + 1u) {
                    k_validate_level<<<(count + 255) / 256, 256>>>(s->tnodes, qn, NN, Lv, front[levels & 1], count, front[(levels + 1) & 1], d_cnt, d_cnt + 1);
                    if (cudaMemcpy(h_cnt, d_cnt, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
                        cudaMemset(d_cnt, 0, sizeof(uint32_t)) != cudaSuccess) { rc = 2; break; }
                    count = h_cnt[0]; visited += count; ++levels;
                }
            }
            dfree(front[0]); dfree(front[1]); dfree(d_cnt);
            if (rc == 1) return 1;
            if (rc) FAIL("scene file validation: %s", cudaGetErrorString(cudaGetLastError()));
            if (h_cnt[1]) FAIL("corrupt scene file (BVH references out of range)");
            if (count || visited > NN) FAIL("corrupt scene file (BVH is not a tree of the stored height)");
        }
        CUDA_TRY(cudaDeviceSynchronize());
        s->use_qnodes = h.use_q != 0;
        s->stats = h.stats;
        for (int a = 0; a < 3; ++a) { s->glo[a] = h.glo[a]; s->cell[a] = h.cell[a]; }
        s->committed = true;
        return 0;
    };
    if (body()) { qsmrt_scene_destroy(s); return 1; }
    *out = s;
    return 0;
}

// per-vertex exposure from per-triangle exposure: vert_counts[v] += tri_counts[t] for the three corners of t
// (ray_casting.py:289-292: hit_tris = triangles[prim_ids]; hit_vert_ids = np.unique(hit_tris))
int qsmrt_vertex_exposure(qsmrt_scene *s, const uint32_t *tri_counts, uint32_t *vert_counts, void *stream)
{
    SCENE_ENTER(s);
    if (!tri_counts || !vert_counts) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (s->ntris == 0) return 0;
    return trv_vertex_exposure(s->idx, s->ntris, tri_counts, vert_counts, st);
}

int qsmrt_get_stats(qsmrt_scene *s, qsmrt_stats *out)
{
    if (!s || !out) FAIL("null pointer");
    *out = s->stats;
    if (!s->committed) {
        uint64_t T = 0; for (const Geometry &g : s->geoms) T += g.T;
        out->num_triangles = T; out->num_geometries = s->geoms.size();
    }
    return 0;
}

int qsmrt_debug_get_build(qsmrt_scene *s, uint64_t *keys, uint32_t *order, void *nodes)
{
    SCENE_ENTER(s);
    if (do_commit(s, nullptr, nullptr)) return 1;
    uint64_t T = s->nleaves;         // = triangles unless slivers were split (then: references, see qsmrt_stats.num_references)
    if (s->ntris == 0) return 0;
    if (keys) CUDA_TRY(cudaMemcpy(keys, s->keys, T * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (order) CUDA_TRY(cudaMemcpy(order, s->order, T * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (nodes && !s->bnodes) FAIL("binary nodes were not kept: set QSMRT_OPT_KEEP_BINARY_NODES before the commit");
    if (nodes) CUDA_TRY(cudaMemcpy(nodes, s->bnodes, (2 * T - 1) * sizeof(BNode), cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"
