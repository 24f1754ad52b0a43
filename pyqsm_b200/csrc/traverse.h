// traverse.h -- launchers of the traversal / scan / ray-generator kernels.
#pragma once
#include "common.cuh"

struct HitRec { float t, u, v; uint32_t geom, prim; };   // 20-byte raw hit (list_intersections scratch)

extern int g_trv_variant;
extern int g_trv_node_path;
extern int g_trv_cp_warp_max;
extern int g_trv_tuning[4];
extern unsigned long long *g_trv_stats_dev;
int trv_cast_rays(const SceneView &sc, const float *rays, uint64_t N, uint32_t row_len, float *t_hit, uint32_t *geom,
                  uint32_t *prim, float *uv, float *nrm, cudaStream_t st);
int trv_count(const SceneView &sc, const float *rays, uint64_t N, int32_t *out, uint32_t ngeoms, cudaStream_t st);
int trv_occluded(const SceneView &sc, const float *rays, uint64_t N, float tnear, float tfar, uint8_t *out, cudaStream_t st);
int trv_raw_count(const SceneView &sc, const float *rays, uint64_t N, int32_t *out, cudaStream_t st);
int trv_raw_fill_sort(const SceneView &sc, const float *rays, uint64_t N, const int64_t *raw_off,
                      HitRec *raw, int32_t *cnt, cudaStream_t st);
int trv_list_compact(uint64_t N, const int64_t *raw_off, const HitRec *raw, const int64_t *splits,
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
int trv_sun_exposure(const SceneView &sc, uint64_t nu, uint64_t nv, const float o0[3], const float du[3],
                     const float dv[3], const float dir[3], const uint64_t *goff, uint32_t *tri_counts, cudaStream_t st);
int trv_sky_visibility(const SceneView &sc, const float *points, const float *normals, uint64_t n_points,
                       uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, uint32_t *unoccluded, cudaStream_t st);
int trv_gen_hemisphere(float *rays, const float *points, const float *normals, uint64_t n_points,
                       uint64_t seed, float offset, uint32_t dir_begin, uint32_t dir_count, cudaStream_t st);
int trv_closest_points(const SceneView &sc, const float *pts, uint64_t N, float *closest, float *dist, uint32_t *geom,
                       uint32_t *prim, float *uv, float *nrm, cudaStream_t st);
int trv_points_to_rays(const float *pts, float *rays, uint64_t N, cudaStream_t st);
int trv_apply_sign(float *dist, const int32_t *counts, uint64_t N, cudaStream_t st);
int trv_peel_cast(const SceneView &sc, uint64_t nu, uint64_t nv, const float o0[3], const float du[3], const float dv[3],
                  const float dir[3], const uint8_t *alive, uint8_t *hitflag, cudaStream_t st);
int trv_peel_update(const SceneView &sc, const uint32_t *order, uint8_t *alive, uint8_t *hitflag, int32_t *layer_of,
                    int layer, const float dir[3], double *sums, cudaStream_t st);
