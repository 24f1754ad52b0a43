// traverse.h -- launchers of the traversal / scan / ray-generator kernels.
#pragma once
#include <vector>
#include "common.cuh"

// Per-scene traversal options (qsmrt_scene_set_option), read at every launch.
struct TrvOptions {
    int variant = 2;        // 1 = one independent loop per thread, 2 = persistent warp-uniform kernel (ships)
    int refill = 12;        // idle lanes that trigger a refill          } tuned on C2
    int want = 16;          // node phase ends below this many searching lanes } (profiles/r01_tuning.txt)
    int tri_min = 1;        // triangle-phase early-exit threshold
    int counters = 0;       // 1: cast_rays launches count the node records / triangles they fetch
    int node_path = 0;      // 0 LSU 256-bit loads, 1 TEX, 2 half/half (L1 data-pipe experiment)
    int cp_warp_max = 16384;// closest-point batches up to this size run one warp per query
    int ctas_per_sm = 0;    // persistent kernel: resident CTAs per SM (0 = as many as fit)
    int point_order = 1;    // sky visibility: 1 = query points worked in Morton order (default), 0 = as stored
    int tile_order = 1;     // 2-D batches: 1 = tile rows from the middle of the grid outwards (default), 0 = memory order
    int count_set = 32;     // count / list: per-lane hit-set entries in shared memory (4..32); rays with more go to the exact slow path
};

// Per-scene launch state: options, the ring of work cursors of the persistent kernels, the counter buffer and
// the occupancy cache.  Owned by the scene, so two scenes (or two threads with their own scenes) never share it.
struct TrvState {
    int device = 0, sms = 0;
    TrvOptions opt;
    unsigned long long *cursor_ring = nullptr; unsigned cursor_next = 0;
    unsigned long long *stats = nullptr;
    struct Occ { const void *fn; size_t smem; int per_sm; };
    std::vector<Occ> occ;
};
void trv_state_free(TrvState &ts);
int trv_read_counters(TrvState &ts, unsigned long long out[16]);
int trv_cast_rays(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, uint32_t row_len, float *t_hit, uint32_t *geom,
                  uint32_t *prim, float *uv, float *nrm, cudaStream_t st);
int trv_count(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, int32_t *out, uint32_t ngeoms, cudaStream_t st);
int trv_occluded(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, float tnear, float tfar, uint8_t *out, cudaStream_t st);
// hit records of one list_intersections batch between its two calls (device memory owned by the caller)
struct ListStash {
    long long *base = nullptr;                  // [N]: first record of ray i
    float *t = nullptr; uint32_t *geom = nullptr, *prim = nullptr; float *uv = nullptr;     // [cap] (uv: [cap][2])
    unsigned long long *count = nullptr, cap = 0;
};
int trv_list_collect(TrvState &ts, const SceneView &sc, const float *rays, uint64_t N, uint32_t ngeoms, int32_t *counts,
                     const ListStash &stash, int *max_fast_out, cudaStream_t st);
int trv_list_emit(const SceneView &sc, const float *rays, uint64_t N, const int64_t *splits, const ListStash &stash, int max_fast,
                  int64_t *ray_ids, float *t_hit, uint32_t *geom, uint32_t *prim, float *uv, cudaStream_t st);
size_t trv_scan_scratch_bytes(uint64_t n);
int trv_exclusive_scan(const int32_t *in, uint64_t n, int64_t *out, void *scratch, cudaStream_t st);
int trv_gen_parallel(float *rays, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                     const float dv[3], const float dir[3], cudaStream_t st);
int trv_gen_pinhole(float *rays, uint32_t w, uint32_t h, const double minv[9], const double eye[3], cudaStream_t st);
int trv_mark_hits(const uint32_t *geom, const uint32_t *prim, uint64_t N, const uint64_t *goff,
                  const uint64_t *voff, uint32_t ngeoms, const uint32_t *idx, uint8_t *tri_hit,
                  uint8_t *vert_hit, cudaStream_t st);
int trv_accumulate_hits(const uint32_t *geom, const uint32_t *prim, uint64_t N, const uint64_t *goff,
                        uint32_t ngeoms, uint32_t *tri_counts, cudaStream_t st);
int trv_vertex_exposure(const uint32_t *idx, uint64_t ntris, const uint32_t *tri_counts, uint32_t *vert_counts, cudaStream_t st);
int trv_sun_exposure(TrvState &ts, const SceneView &sc, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                     const float dv[3], const float dir[3], const uint64_t *goff, uint32_t *tri_counts, cudaStream_t st);
int trv_sun_exposure_sweep(TrvState &ts, const SceneView &sc, uint32_t n_grids, const float *sweep_dev, uint64_t nu, uint64_t nv,
                           const uint64_t *goff, uint32_t *tri_counts, uint64_t count_stride, cudaStream_t st);
int trv_sky_visibility(TrvState &ts, const SceneView &sc, const float *points, const float *normals, uint64_t n_points, uint64_t point_base,
                       uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, const uint32_t *perm, uint32_t *unoccluded,
                       cudaStream_t st);
int trv_point_keys(const float *points, uint64_t n, const float lo[3], const float hi[3], uint64_t *keys, uint32_t *vals, cudaStream_t st);
int trv_gen_hemisphere(float *rays, const float *points, const float *normals, uint64_t n_points, uint64_t point_base,
                       uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, cudaStream_t st);
int trv_closest_points(TrvState &ts, const SceneView &sc, const float *pts, uint64_t N, float *closest, float *dist, uint32_t *geom,
                       uint32_t *prim, float *uv, float *nrm, cudaStream_t st);
int trv_points_to_rays(const float *pts, float *rays, uint64_t N, cudaStream_t st);
int trv_apply_sign(float *dist, const int32_t *counts, uint64_t N, cudaStream_t st);
int trv_peel_cast(TrvState &ts, const SceneView &sc, uint64_t nu, uint64_t nv, const float o0[3], const float du[3], const float dv[3],
                  const float dir[3], const uint32_t *order, const uint8_t *alive, uint8_t *hitflag, cudaStream_t st);
int trv_peel_update(const float *verts, const uint32_t *idx, uint64_t ntris, uint8_t *alive, uint8_t *hitflag, int32_t *layer_of,
                    int layer, const float dir[3], double *sums, cudaStream_t st);
