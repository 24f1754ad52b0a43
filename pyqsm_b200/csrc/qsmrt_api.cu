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
static int g_allow_qnodes = 1;              // qsmrt_debug_set_quantised_nodes
static bool g_keep_bnodes = false;          // qsmrt_debug_set_keep_binary_nodes
constexpr int QSMRT_MAX_DEVICES = 64;
static int g_leaf_max = 2;                  // triangles per leaf (qsmrt_debug_set_leaf_max); 2 measured best on C2

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
};

struct HostPipe {            // staging for qsmrt_cast_rays_host
    static constexpr int NBUF = 3;
    uint64_t chunk = 0;
    float *rays[NBUF] = {}; float *t[NBUF] = {}; uint32_t *g[NBUF] = {}; uint32_t *p[NBUF] = {};
    float *uv[NBUF] = {}; float *nrm[NBUF] = {};
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t e_in[NBUF] = {}, e_run[NBUF] = {}, e_out[NBUF] = {};
};

struct qsmrt_scene {
    int device = 0;
    std::vector<Geometry> geoms;
    bool committed = false;
    // concatenated mesh (aliases geoms[0] when there is a single geometry)
    float *verts = nullptr; uint32_t *idx = nullptr; bool own_concat = false;
    uint64_t ntris = 0, nverts = 0;
    uint64_t *goff = nullptr, *voff = nullptr;       // device [ngeoms+1]
    // build products kept for traversal / introspection
    uint64_t *keys = nullptr; uint32_t *order = nullptr;
    BNode *bnodes = nullptr; TNode *tnodes = nullptr; TriRec *tris = nullptr;
    QNode *qnodes = nullptr; bool use_qnodes = false; float glo[3] = {}, cell[3] = {};
    BuildParams *params = nullptr;
    qsmrt_stats stats{};
    cudaTextureObject_t node_tex = 0;
    // list_intersections cache between _count and _fill
    const float *list_rays = nullptr; uint64_t list_n = 0;
    int64_t *list_raw_off = nullptr; HitRec *list_raw = nullptr;
    HostPipe pipe;
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

void free_build(qsmrt_scene *s)
{
    if (s->own_concat) { dfree(s->verts); dfree(s->idx); }
    s->verts = nullptr; s->idx = nullptr; s->own_concat = false;
    dfree(s->goff); dfree(s->voff); dfree(s->keys); dfree(s->order);
    dfree(s->bnodes); dfree(s->tnodes); dfree(s->tris); dfree(s->params); dfree(s->qnodes);
    s->use_qnodes = false;
    dfree(s->list_raw_off); dfree(s->list_raw);
    s->list_rays = nullptr; s->list_n = 0;
    if (s->node_tex) { cudaDestroyTextureObject(s->node_tex); s->node_tex = 0; }
    s->committed = false;
}

void free_pipe(HostPipe &hp)
{
    for (int b = 0; b < HostPipe::NBUF; ++b) {
        dfree(hp.rays[b]); dfree(hp.t[b]); dfree(hp.g[b]); dfree(hp.p[b]); dfree(hp.uv[b]); dfree(hp.nrm[b]);
        if (hp.e_in[b]) cudaEventDestroy(hp.e_in[b]);
        if (hp.e_run[b]) cudaEventDestroy(hp.e_run[b]);
        if (hp.e_out[b]) cudaEventDestroy(hp.e_out[b]);
        hp.e_in[b] = hp.e_run[b] = hp.e_out[b] = nullptr;
    }
    if (hp.s_in) cudaStreamDestroy(hp.s_in);
    if (hp.s_run) cudaStreamDestroy(hp.s_run);
    if (hp.s_out) cudaStreamDestroy(hp.s_out);
    hp.s_in = hp.s_run = hp.s_out = nullptr;
    hp.chunk = 0;
}

__global__ void k_rebase_idx(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t n3, uint32_t add)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n3) out[i] = in[i] + add;
}

__global__ void k_max_index(const uint32_t *__restrict__ in, uint64_t n3, uint32_t *out)
{
    uint32_t m = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n3; i += (uint64_t)gridDim.x * blockDim.x) m = max(m, in[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
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

int use_device(qsmrt_scene *s)
{
    if (!s) FAIL("null scene");
    CUDA_TRY(cudaSetDevice(s->device));
    return 0;
}

int check_rays(const float *rays, uint64_t N)
{
    if (N && !rays) FAIL("rays pointer is null");
    if (reinterpret_cast<uintptr_t>(rays) & 7u) FAIL("rays must be 8-byte aligned");
    return 0;
}

SceneView view_of(qsmrt_scene *s)
{
    if (g_trv_node_path != 0 && !s->node_tex && s->tnodes) {
        // float4 texture view of the node array, made on first use: only the TEX-path experiment
        // (qsmrt_debug_set_node_path) reads it, and creating it cost every commit a driver call
        cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = s->tnodes;
        rd.res.linear.desc = cudaCreateChannelDesc<float4>();
        rd.res.linear.sizeInBytes = std::max<uint64_t>(s->ntris - 1, 1) * sizeof(TNode);
        cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
        if (cudaCreateTextureObject(&s->node_tex, &rd, &td, nullptr) != cudaSuccess) { s->node_tex = 0; cudaGetLastError(); }
    }
    SceneView v;
    v.nodes = s->tnodes; v.tris = s->tris; v.ntris = (uint32_t)s->ntris; v.height = s->stats.bvh_height;
    v.node_tex = s->node_tex;
    v.qnodes = (s->use_qnodes && g_allow_qnodes) ? s->qnodes : nullptr;
    for (int a = 0; a < 3; ++a) { v.glo[a] = s->glo[a]; v.cell[a] = s->cell[a]; }
    return v;
}

int do_commit(qsmrt_scene *s, cudaStream_t st, float *build_ms_out)
{
    if (s->committed) { if (build_ms_out) *build_ms_out = s->stats.build_ms; return 0; }
    SyncedFrees batch;          // covers the re-commit teardown here and, after the build's own event wait, the scratch frees
    free_build(s);
    const uint32_t G = (uint32_t)s->geoms.size();
    uint64_t T = 0, V = 0;
    std::vector<uint64_t> goff(G + 1, 0), voff(G + 1, 0);
    for (uint32_t g = 0; g < G; ++g) { goff[g] = T; voff[g] = V; T += s->geoms[g].T; V += s->geoms[g].V; }
    goff[G] = T; voff[G] = V;
    if (T >= (1ull << 29)) FAIL("scene has %llu triangles; the leaf encoding holds 2^29", (unsigned long long)T);
    if (V >= (1ull << 32)) FAIL("scene has %llu vertices; indices are 32-bit", (unsigned long long)V);
    s->ntris = T; s->nverts = V;
    memset(&s->stats, 0, sizeof(s->stats));
    s->stats.num_triangles = T; s->stats.num_geometries = G; s->stats.leaf_max = (uint32_t)g_leaf_max;
    s->committed = true;
    if (T == 0) { if (build_ms_out) *build_ms_out = 0.0f; return 0; }

    cudaEvent_t e0, e1, es0, es1;
    CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
    CUDA_TRY(cudaEventCreate(&es0)); CUDA_TRY(cudaEventCreate(&es1));

    if (dmalloc(&s->goff, G + 1) || dmalloc(&s->voff, G + 1)) return 1;
    CUDA_TRY(cudaMemcpyAsync(s->goff, goff.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->voff, voff.data(), (G + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
    // allocate everything before the timed region
    uint64_t *keys_tmp = nullptr; uint32_t *order_tmp = nullptr, *sort_scratch = nullptr, *bounds = nullptr;
    unsigned long long *flags = nullptr, *counters = nullptr; uint32_t *climb = nullptr;
    if (dmalloc(&s->keys, T) || dmalloc(&keys_tmp, T) || dmalloc(&s->order, T) || dmalloc(&order_tmp, T) ||
        dmalloc(&sort_scratch, lbvh_sort_scratch_bytes(T) / sizeof(uint32_t)) || dmalloc(&bounds, 8) ||
        dmalloc(&s->params, 1) || dmalloc(&s->bnodes, 2 * T - 1) || dmalloc(&flags, T) || dmalloc(&s->tris, T) ||
        dmalloc(&s->tnodes, std::max<uint64_t>(T - 1, 1)) || dmalloc(&s->qnodes, std::max<uint64_t>(T - 1, 1)) ||
        dmalloc(&counters, 5) || dmalloc(&climb, lbvh_climb_bytes(T) / sizeof(uint32_t)))
        return 1;
    if (G == 1) { s->verts = s->geoms[0].verts; s->idx = s->geoms[0].idx; s->own_concat = false; }
    else {
        if (dmalloc(&s->verts, 3 * V) || dmalloc(&s->idx, 3 * T)) return 1;
        s->own_concat = true;
    }

    CUDA_TRY(cudaEventRecord(e0, st));
    if (G > 1) {
        for (uint32_t g = 0; g < G; ++g) {
            const Geometry &ge = s->geoms[g];
            if (ge.V) CUDA_TRY(cudaMemcpyAsync(s->verts + 3 * voff[g], ge.verts, 3 * ge.V * sizeof(float), cudaMemcpyDeviceToDevice, st));
            if (ge.T) k_rebase_idx<<<(unsigned)((3 * ge.T + 255) / 256), 256, 0, st>>>(ge.idx, s->idx + 3 * goff[g], 3 * ge.T, (uint32_t)voff[g]);
        }
    }
    LbvhBuildArgs A{};
    A.leaf_max = g_leaf_max;
    A.verts = s->verts; A.idx = s->idx; A.ntris = T; A.geom_offsets = s->goff; A.ngeoms = G;
    A.bounds_ord = bounds; A.params = s->params; A.keys = s->keys; A.keys_tmp = keys_tmp;
    A.order = s->order; A.order_tmp = order_tmp; A.sort_scratch = sort_scratch;
    A.bnodes = s->bnodes; A.flags = flags; A.keep_bnodes = g_keep_bnodes ? 1 : 0; A.climb_work = climb;
    A.qnodes = s->qnodes;
    A.tris = s->tris; A.tnodes = s->tnodes; A.counters = counters; A.ev_sort0 = es0; A.ev_sort1 = es1;
    int rc = 0;
    unsigned long long cnt[5] = {};
    // The sort normally runs over the top 40 key bits plus an exact fix-up of short runs; a scene with a run of more
    // than 64 triangles in one 2^-13 cell (thousands of coincident triangles) reports an overflow and is built again
    // with all eight passes.  Both attempts are inside the timed region.
    for (int attempt = 0; attempt < 2 && !rc; ++attempt) {
        A.full_sort = attempt;
        rc = lbvh_build(A, st);
        if (rc) break;
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        CUDA_TRY(cudaMemcpy(cnt, counters, sizeof(cnt), cudaMemcpyDeviceToHost));
        if (!cnt[4]) break;
    }
    s->stats.full_sort = (uint32_t)A.full_sort;
    if (!rc) {
        CUDA_TRY(cudaEventElapsedTime(&s->stats.build_ms, e0, e1));
        CUDA_TRY(cudaEventElapsedTime(&s->stats.sort_ms, es0, es1));
        BuildParams bp;
        CUDA_TRY(cudaMemcpy(&bp, s->params, sizeof(bp), cudaMemcpyDeviceToHost));
        for (int a = 0; a < 3; ++a) { s->stats.scene_lo[a] = bp.slo[a]; s->stats.scene_hi[a] = bp.shi[a]; }
        s->stats.box_pad = bp.pad;
        s->stats.num_bvh_nodes = cnt[0]; s->stats.num_bvh_leaves = cnt[1]; s->stats.bvh_height = (uint32_t)cnt[2];
        s->use_qnodes = bp.use_q != 0;                      // decided on the device (k_decide_quant)
        if (!s->use_qnodes) dfree(s->qnodes);
        if (!g_keep_bnodes) dfree(s->bnodes);               // only the hand-over boxes of the build were in it
        for (int a = 0; a < 3; ++a) { s->glo[a] = bp.glo[a]; s->cell[a] = bp.cell[a]; }
        s->stats.quantised_nodes = s->use_qnodes ? 1u : 0u;
        s->stats.bvh_bytes = cnt[0] * (s->use_qnodes ? sizeof(QNode) : sizeof(TNode)) + T * sizeof(TriRec);
    }
    if (rc) cudaDeviceSynchronize();        // a failed build may still have kernels in flight
    dfree(keys_tmp); dfree(order_tmp); dfree(sort_scratch); dfree(bounds); dfree(flags);
    dfree(counters); dfree(climb);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(es0); cudaEventDestroy(es1);
    if (rc) { s->committed = false; return 1; }
    if (build_ms_out) *build_ms_out = s->stats.build_ms;
    return 0;
}

int ensure_pipe(qsmrt_scene *s, uint64_t chunk)
{
    HostPipe &hp = s->pipe;
    if (hp.chunk >= chunk) return 0;
    free_pipe(hp);
    for (int b = 0; b < HostPipe::NBUF; ++b) {
        if (dmalloc(&hp.rays[b], 6 * chunk) || dmalloc(&hp.t[b], chunk) || dmalloc(&hp.g[b], chunk) ||
            dmalloc(&hp.p[b], chunk) || dmalloc(&hp.uv[b], 2 * chunk) || dmalloc(&hp.nrm[b], 3 * chunk))
            return 1;
        CUDA_TRY(cudaEventCreateWithFlags(&hp.e_in[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&hp.e_run[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&hp.e_out[b], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_run, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
    hp.chunk = chunk;
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
    CUDA_TRY(cudaSetDevice(cuda_device));
    int major = 0, minor = 0;       // two attribute reads, not cudaGetDeviceProperties (2.4 ms per new scene)
    CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, cuda_device));
    CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, cuda_device));
    if (major != 10) FAIL("device %d is sm_%d%d; libqsmrt is built for sm_100a only", cuda_device, major, minor);
    qsmrt_scene *s = new qsmrt_scene();
    s->device = cuda_device;
    *out = s;
    return 0;
}

int qsmrt_scene_destroy(qsmrt_scene *s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    SyncedFrees batch;
    free_build(s);
    free_pipe(s->pipe);
    for (Geometry &g : s->geoms) { dfree(g.verts); dfree(g.idx); }
    delete s;
    return 0;
}

int qsmrt_add_triangles(qsmrt_scene *s, const float *verts, uint64_t V, const uint32_t *idx, uint64_t T,
                        int on_device, uint32_t *geom_id_out)
{
    if (use_device(s)) return 1;
    if ((V && !verts) || (T && !idx)) FAIL("null vertex or index pointer");
    if (V >= (1ull << 32)) FAIL("too many vertices");
    Geometry g;
    g.V = V; g.T = T;
    if (dmalloc(&g.verts, 3 * V) || dmalloc(&g.idx, 3 * T)) { dfree(g.verts); dfree(g.idx); return 1; }
    cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (V) CUDA_TRY(cudaMemcpy(g.verts, verts, 3 * V * sizeof(float), kind));
    if (T) CUDA_TRY(cudaMemcpy(g.idx, idx, 3 * T * sizeof(uint32_t), kind));
    // Embree would read out of bounds; reject instead
    uint32_t maxi = 0;
    if (T) {
        uint32_t *d = nullptr;
        if (dmalloc(&d, 1)) { dfree(g.verts); dfree(g.idx); return 1; }
        cudaMemset(d, 0, sizeof(uint32_t));
        k_max_index<<<(unsigned)std::min<uint64_t>((3 * T + 255) / 256, 1184), 256>>>(g.idx, 3 * T, d);
        cudaError_t e = cudaMemcpy(&maxi, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
        dfree(d);
        if (e != cudaSuccess) { dfree(g.verts); dfree(g.idx); FAIL("index check failed: %s", cudaGetErrorString(e)); }
        if (maxi >= V) { dfree(g.verts); dfree(g.idx); FAIL("triangle index %u out of range (%llu vertices)", maxi, (unsigned long long)V); }
    }
    if (s->committed || s->verts) free_build(s);
    s->geoms.push_back(g);
    if (geom_id_out) *geom_id_out = (uint32_t)(s->geoms.size() - 1);
    return 0;
}

int qsmrt_add_cylinders(qsmrt_scene *s, const float *records, uint64_t n, uint32_t resolution, uint32_t split,
                        int on_device, uint32_t *geom_id_out)
{
    if (use_device(s)) return 1;
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
            CUDA_TRY(cudaMemcpy(rec_dev, records, 8 * n * sizeof(float), cudaMemcpyHostToDevice));
            rec = rec_dev;
        }
        k_cylinders<<<(unsigned)n, 128>>>(rec, n, resolution, split, g.verts, g.idx);
        cudaError_t e = cudaDeviceSynchronize();
        dfree(rec_dev);
        if (e != cudaSuccess) { dfree(g.verts); dfree(g.idx); FAIL("cylinder generation failed: %s", cudaGetErrorString(e)); }
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
    if (use_device(s)) return 1;
    if (geom_id >= s->geoms.size()) FAIL("geometry id %u out of range", geom_id);
    const Geometry &g = s->geoms[geom_id];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (verts_dev && g.V) CUDA_TRY(cudaMemcpyAsync(verts_dev, g.verts, 3 * g.V * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (idx_dev && g.T) CUDA_TRY(cudaMemcpyAsync(idx_dev, g.idx, 3 * g.T * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int qsmrt_commit(qsmrt_scene *s, void *stream, float *build_ms_out)
{
    if (use_device(s)) return 1;
    return do_commit(s, static_cast<cudaStream_t>(stream), build_ms_out);
}

int qsmrt_cast_rays(qsmrt_scene *s, const float *rays, uint64_t N, float *t_hit, uint32_t *geom, uint32_t *prim,
                    float *uv, float *nrm, void *stream)
{
    if (use_device(s) || check_rays(rays, N)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    return trv_cast_rays(view_of(s), rays, N, 0, t_hit, geom, prim, uv, nrm, st);
}

int qsmrt_cast_rays_2d(qsmrt_scene *s, const float *rays, uint32_t width, uint64_t height, float *t_hit, uint32_t *geom,
                       uint32_t *prim, float *uv, float *nrm, void *stream)
{
    if (use_device(s) || check_rays(rays, (uint64_t)width * height)) return 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    return trv_cast_rays(view_of(s), rays, (uint64_t)width * height, width, t_hit, geom, prim, uv, nrm, st);
}

int qsmrt_debug_set_variant(int variant)
{
    if (variant != 1 && variant != 2) FAIL("unknown traversal variant %d (1 = per-thread loop, 2 = persistent kernel)", variant);
    g_trv_variant = variant;
    return 0;
}

int qsmrt_debug_set_tuning(int refill_thresh, int want_thresh, int speculate, int counters)
{
    if (refill_thresh < 1 || refill_thresh > 32 || want_thresh < 1 || want_thresh > 32) FAIL("thresholds must be in 1..32");
    g_trv_tuning[0] = refill_thresh; g_trv_tuning[1] = want_thresh; g_trv_tuning[2] = speculate != 0; g_trv_tuning[3] = counters != 0;
    return 0;
}

int qsmrt_debug_get_counters(uint64_t *nodes_out, uint64_t *tris_out)
{
    unsigned long long h[2] = { 0, 0 };
    if (g_trv_stats_dev) CUDA_TRY(cudaMemcpy(h, g_trv_stats_dev, sizeof(h), cudaMemcpyDeviceToHost));
    if (nodes_out) *nodes_out = h[0];
    if (tris_out) *tris_out = h[1];
    return 0;
}

int qsmrt_debug_set_sort(int variant)
{
    if (variant != 0 && variant != 1) FAIL("sort variant must be 0 (classic) or 1 (onesweep)");
    g_sort_variant = variant;
    return 0;
}

int qsmrt_release_cached_memory(void)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_blocks.m);
    g_blocks.release_all();
    return 0;
}

int qsmrt_debug_set_cp_warp_max(int max_points)
{
    g_trv_cp_warp_max = max_points < 0 ? 0 : max_points;
    return 0;
}

int qsmrt_debug_set_quant_threshold(float frac)
{
    g_quant_frac = frac > 0.0f ? frac : 0.15f;
    return 0;
}

int qsmrt_debug_set_climb_capacity(int items)
{
    g_climb_cap_override = items > 0 ? items : 0;
    return 0;
}

int qsmrt_debug_set_keep_binary_nodes(int keep)
{
    g_keep_bnodes = keep != 0;
    return 0;
}

int qsmrt_debug_set_quantised_nodes(int allow)
{
    g_allow_qnodes = allow != 0;
    return 0;
}

int qsmrt_debug_set_node_path(int path)
{
    if (path < 0 || path > 2) FAIL("node path must be 0, 1 or 2");
    g_trv_node_path = path;
    return 0;
}

int qsmrt_debug_set_leaf_max(int leaf_max)
{
    if (leaf_max < 1 || leaf_max > QSMRT_LEAF_MAX) FAIL("leaf_max must be in 1..%d", QSMRT_LEAF_MAX);
    g_leaf_max = leaf_max;
    return 0;
}

int qsmrt_debug_get_census(uint64_t out[16])
{
    if (!out) FAIL("null pointer");
    memset(out, 0, 16 * sizeof(uint64_t));
    if (g_trv_stats_dev) CUDA_TRY(cudaMemcpy(out, g_trv_stats_dev, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return 0;
}

int qsmrt_cast_rays_host(qsmrt_scene *s, const float *rays, uint64_t N, float *t_hit, uint32_t *geom,
                         uint32_t *prim, float *uv, float *nrm)
{
    if (use_device(s)) return 1;
    if (N && !rays) FAIL("rays pointer is null");
    if (do_commit(s, nullptr, nullptr)) return 1;
    if (N == 0) return 0;
    uint64_t chunk_rays = 1ull << 20;                       // QSMRT_HOST_CHUNK overrides (rays per pipeline stage)
    if (const char *e = getenv("QSMRT_HOST_CHUNK")) { long long v = atoll(e); if (v >= 1024) chunk_rays = (uint64_t)v; }
    const uint64_t chunk = std::min<uint64_t>(N, chunk_rays);
    if (ensure_pipe(s, chunk)) return 1;
    HostPipe &hp = s->pipe;
    SceneView sv = view_of(s);
    uint64_t nchunks = (N + chunk - 1) / chunk;
    for (uint64_t c = 0; c < nchunks; ++c) {
        int b = (int)(c % HostPipe::NBUF);
        uint64_t off = c * chunk, n = std::min(chunk, N - off);
        if (c >= HostPipe::NBUF) CUDA_TRY(cudaStreamWaitEvent(hp.s_in, hp.e_out[b], 0));   // buffer drained
        CUDA_TRY(cudaMemcpyAsync(hp.rays[b], rays + 6 * off, 6 * n * sizeof(float), cudaMemcpyHostToDevice, hp.s_in));
        CUDA_TRY(cudaEventRecord(hp.e_in[b], hp.s_in));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_run, hp.e_in[b], 0));
        if (trv_cast_rays(sv, hp.rays[b], n, 0, t_hit ? hp.t[b] : nullptr, geom ? hp.g[b] : nullptr,
                          prim ? hp.p[b] : nullptr, uv ? hp.uv[b] : nullptr, nrm ? hp.nrm[b] : nullptr, hp.s_run))
            return 1;
        CUDA_TRY(cudaEventRecord(hp.e_run[b], hp.s_run));
        CUDA_TRY(cudaStreamWaitEvent(hp.s_out, hp.e_run[b], 0));
        if (t_hit) CUDA_TRY(cudaMemcpyAsync(t_hit + off, hp.t[b], n * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        if (geom) CUDA_TRY(cudaMemcpyAsync(geom + off, hp.g[b], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, hp.s_out));
        if (prim) CUDA_TRY(cudaMemcpyAsync(prim + off, hp.p[b], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, hp.s_out));
        if (uv) CUDA_TRY(cudaMemcpyAsync(uv + 2 * off, hp.uv[b], 2 * n * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        if (nrm) CUDA_TRY(cudaMemcpyAsync(nrm + 3 * off, hp.nrm[b], 3 * n * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        CUDA_TRY(cudaEventRecord(hp.e_out[b], hp.s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(hp.s_out));
    CUDA_TRY(cudaStreamSynchronize(hp.s_run));
    CUDA_TRY(cudaStreamSynchronize(hp.s_in));
    return 0;
}

int qsmrt_count_intersections(qsmrt_scene *s, const float *rays, uint64_t N, int32_t *counts, void *stream)
{
    if (use_device(s) || check_rays(rays, N)) return 1;
    if (N && !counts) FAIL("counts pointer is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_count(view_of(s), rays, N, counts, (uint32_t)s->geoms.size(), st);
}

int qsmrt_test_occlusions(qsmrt_scene *s, const float *rays, uint64_t N, float tnear, float tfar, uint8_t *out, void *stream)
{
    if (use_device(s) || check_rays(rays, N)) return 1;
    if (N && !out) FAIL("output pointer is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_occluded(view_of(s), rays, N, tnear, tfar, out, st);
}

int qsmrt_list_intersections_count(qsmrt_scene *s, const float *rays, uint64_t N, int64_t *ray_splits,
                                   int64_t *total_out, void *stream)
{
    if (use_device(s) || check_rays(rays, N)) return 1;
    if (!ray_splits || !total_out) FAIL("null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    dfree(s->list_raw_off); dfree(s->list_raw);
    s->list_rays = nullptr; s->list_n = 0;
    int32_t *cnt = nullptr; void *scratch = nullptr;
    if (dmalloc(&cnt, N) || dmalloc(&s->list_raw_off, N + 1) ||
        dmalloc(reinterpret_cast<char **>(&scratch), trv_scan_scratch_bytes(N))) { dfree(cnt); return 1; }
    SceneView sv = view_of(s);
    int64_t raw_total = 0;
    int rc = trv_raw_count(sv, rays, N, cnt, st) || trv_exclusive_scan(cnt, N, s->list_raw_off, scratch, st);
    if (!rc && cudaMemcpyAsync(&raw_total, s->list_raw_off + N, sizeof(int64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = 1;
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
    if (!rc) rc = dmalloc(&s->list_raw, (uint64_t)raw_total);
    if (!rc) rc = trv_raw_fill_sort(sv, rays, N, s->list_raw_off, s->list_raw, cnt, st) ||
                  trv_exclusive_scan(cnt, N, ray_splits, scratch, st);
    if (!rc && cudaMemcpyAsync(total_out, ray_splits + N, sizeof(int64_t), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = 1;
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
    dfree(cnt);
    { char *p = reinterpret_cast<char *>(scratch); dfree(p); }
    if (rc) { if (!g_err[0]) qsmrt_set_error("list_intersections failed: %s", cudaGetErrorString(cudaGetLastError())); return 1; }
    s->list_rays = rays; s->list_n = N;
    return 0;
}

int qsmrt_list_intersections_fill(qsmrt_scene *s, const float *rays, uint64_t N, const int64_t *ray_splits,
                                  int64_t *ray_ids, float *t_hit, uint32_t *geom, uint32_t *prim, float *uv, void *stream)
{
    if (use_device(s)) return 1;
    if (!s->list_raw_off || s->list_rays != rays || s->list_n != N)
        FAIL("list_intersections_fill must follow list_intersections_count on the same rays");
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = trv_list_compact(N, s->list_raw_off, s->list_raw, ray_splits, ray_ids, t_hit, geom, prim, uv, st);
    if (!rc) CUDA_TRY(cudaStreamSynchronize(st));
    dfree(s->list_raw_off); dfree(s->list_raw);
    s->list_rays = nullptr; s->list_n = 0;
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
    if (use_device(s)) return 1;
    if (!prim) FAIL("primitive_ids pointer is null");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (s->ntris == 0) return 0;
    return trv_mark_hits(geom, prim, N, s->goff, s->voff, (uint32_t)s->geoms.size(), s->idx, tri_hit, vert_hit, st);
}

int qsmrt_accumulate_hits(qsmrt_scene *s, const uint32_t *geom, const uint32_t *prim, uint64_t N,
                          uint32_t *tri_counts, void *stream)
{
    if (use_device(s)) return 1;
    if (!prim || !tri_counts) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (s->ntris == 0) return 0;
    return trv_accumulate_hits(geom, prim, N, s->goff, (uint32_t)s->geoms.size(), tri_counts, st);
}

int qsmrt_closest_points(qsmrt_scene *s, const float *pts, uint64_t N, float *closest, float *dist, uint32_t *geom,
                         uint32_t *prim, float *uv, float *nrm, void *stream)
{
    if (use_device(s)) return 1;
    if (N && !pts) FAIL("query points pointer is null");
    if (reinterpret_cast<uintptr_t>(uv) & 7u) FAIL("primitive_uvs must be 8-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_closest_points(view_of(s), pts, N, closest, dist, geom, prim, uv, nrm, st);
}

int qsmrt_signed_distance(qsmrt_scene *s, const float *pts, uint64_t N, float *dist, void *stream)
{
    if (use_device(s)) return 1;
    if (N && (!pts || !dist)) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    if (N == 0) return 0;
    float *rays = nullptr; int32_t *counts = nullptr;
    if (dmalloc(&rays, 6 * N) || dmalloc(&counts, N)) { dfree(rays); dfree(counts); return 1; }
    int rc = trv_closest_points(view_of(s), pts, N, nullptr, dist, nullptr, nullptr, nullptr, nullptr, st) ||
             trv_points_to_rays(pts, rays, N, st) ||
             trv_count(view_of(s), rays, N, counts, (uint32_t)s->geoms.size(), st) ||
             trv_apply_sign(dist, counts, N, st);
    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) { qsmrt_set_error("signed_distance: %s", cudaGetErrorString(cudaGetLastError())); rc = 1; }
    dfree(rays); dfree(counts);
    return rc;
}

int qsmrt_sun_exposure(qsmrt_scene *s, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                       const float dv[3], const float dir[3], uint32_t *tri_counts, void *stream)
{
    if (use_device(s)) return 1;
    if (!o0 || !du || !dv || !dir || !tri_counts) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_sun_exposure(view_of(s), nu, nv, o0, du, dv, dir, s->geoms.size() > 1 ? s->goff : nullptr, tri_counts, st);
}

int qsmrt_sky_visibility(qsmrt_scene *s, const float *points, const float *normals, uint64_t n_points,
                         uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count,
                         uint32_t *unoccluded, void *stream)
{
    if (use_device(s)) return 1;
    if (!points || !unoccluded) FAIL("null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (do_commit(s, st, nullptr)) return 1;
    return trv_sky_visibility(view_of(s), points, normals, n_points, seed, offset, dir_begin, dir_count, unoccluded, st);
}

int qsmrt_gen_hemisphere_rays(float *rays, const float *points, const float *normals, uint64_t n_points,
                              uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, void *stream)
{
    if (!rays || !points) FAIL("null pointer");
    if (reinterpret_cast<uintptr_t>(rays) & 7u) FAIL("rays must be 8-byte aligned");
    return trv_gen_hemisphere(rays, points, normals, n_points, seed, offset, dir_begin, dir_count, static_cast<cudaStream_t>(stream));
}

int qsmrt_peel_projection(qsmrt_scene *s, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                          const float dv[3], const float dir[3], int max_layers, int32_t *layer_of,
                          double *layer_stats, int *n_layers_out, void *stream)
{
    if (use_device(s)) return 1;
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
        rc = trv_peel_cast(sv, nu, nv, o0, du, dv, dir, alive, hit, st) ||
             trv_peel_update(sv, s->order, alive, hit, layer_of, layer, dir, sums, st);
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
    if (use_device(s)) return 1;
    if (do_commit(s, nullptr, nullptr)) return 1;
    uint64_t T = s->ntris;
    if (T == 0) return 0;
    if (keys) CUDA_TRY(cudaMemcpy(keys, s->keys, T * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (order) CUDA_TRY(cudaMemcpy(order, s->order, T * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (nodes && !s->bnodes) FAIL("binary nodes were not kept: call qsmrt_debug_set_keep_binary_nodes(1) before the commit");
    if (nodes) CUDA_TRY(cudaMemcpy(nodes, s->bnodes, (2 * T - 1) * sizeof(BNode), cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"
