/*
 * qsmrt.h -- C ABI of libqsmrt.so, the B200 (sm_100a) ray/mesh intersection
 * engine behind pyqsm_b200.RaycastingScene.
 *
 * Each entry point replaces one method of open3d.t.geometry.RaycastingScene as
 * it is called from the reference, wischmcj/pyQSM pyQSM/viz/ray_casting.py
 * (the import at :8 binds `rcs`; the call sites are cited per function).
 * Plain pointers and sizes only: no torch, no C++ types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     qsmrt_last_error() then holds a thread-local message.
 *   - "dev" pointers are CUDA device pointers on the scene's device; "host"
 *     pointers are ordinary (ideally pinned) host memory.  Caller owns all
 *     ray/result buffers.  A NULL result pointer skips that output.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Query entry points only enqueue work; they do not synchronise unless
 *     they take host pointers or the documentation says so.
 *   - one scene is used from one thread, and its calls are ordered on one
 *     stream at a time (as with Open3D's scene): some calls keep scratch in
 *     the scene between launches (the work cursor of the persistent kernels,
 *     the sun sweep's grid table, the sky driver's point order, the hit
 *     records between list_intersections_count and _fill).  Different scenes
 *     are independent, on any threads and streams.
 *   - rays are N x 6 float32 rows (ox,oy,oz,dx,dy,dz); directions are used
 *     as given (not normalised), so t is in units of |d|.
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef QSMRT_H
#define QSMRT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSMRT_INVALID_ID 0xFFFFFFFFu /* RaycastingScene.INVALID_ID */
#define QSMRT_ABI_VERSION 2

typedef struct qsmrt_scene qsmrt_scene;

typedef struct qsmrt_stats {
    uint64_t num_triangles;      /* over all geometries */
    uint64_t num_geometries;
    uint64_t num_bvh_nodes;      /* traversal nodes emitted (after leaf collapse) */
    uint64_t num_bvh_leaves;
    uint64_t bvh_bytes;          /* traversal nodes + triangle records */
    float    build_ms;           /* last commit: CUDA-event time of the whole build */
    float    sort_ms;            /* radix sort share of build_ms */
    float    box_pad;            /* absolute AABB padding */
    float    scene_lo[3], scene_hi[3];
    uint32_t leaf_max;           /* triangles per collapsed leaf */
    uint32_t bvh_height;         /* binary LBVH height (bounds the traversal stack) */
    uint32_t quantised_nodes;    /* 1: the persistent kernel reads the 32-byte 16-bit-grid nodes */
    uint32_t full_sort;          /* 1: the last commit fell back to all eight radix passes (a run of > 64 keys equal in their top bits) */
    uint64_t num_references;     /* leaves of the LBVH before collapsing: num_triangles, more when sliver triangles were split */
} qsmrt_stats;

const char *qsmrt_last_error(void);
int qsmrt_abi_version(void);

/* rcs()  -- ray_casting.py:65,155,218,241,275,316 */
int qsmrt_scene_create(int cuda_device, qsmrt_scene **out);
int qsmrt_scene_destroy(qsmrt_scene *scene);
/* Device blocks freed by commits, queries and destroyed scenes are kept (up to QSMRT_CACHE_MB megabytes per
 * process, default 1024) and reused, because the reference builds a new scene per call and cudaMalloc / cudaFree
 * dominated small scenes.  This hands them back to the driver (like torch.cuda.empty_cache()).  Synchronises. */
int qsmrt_release_cached_memory(void);

/* Per-scene options.  Open3D's constructor takes nthreads / device
 * (ray_casting.py:65,...: rcs()); everything else here is tuning and test
 * hooks, stored IN THE SCENE: builder options are read at the next commit (a
 * changed one marks the scene for rebuilding), traversal options at every
 * launch.  No process-global state; scenes with different options can be used
 * from different threads. */
enum qsmrt_option {
    /* builder */
    QSMRT_OPT_LEAF_MAX = 1,            /* triangles per collapsed leaf, 1..4 (default 2) */
    QSMRT_OPT_KEEP_BINARY_NODES = 2,   /* 1: keep the complete 32-byte binary node array for qsmrt_debug_get_build */
    QSMRT_OPT_QUANT_THRESHOLD = 3,     /* 32-byte quantised nodes when 6 grid cells <= value x mean leaf-box diagonal (default 0.15; <= 0 restores it) */
    QSMRT_OPT_CLIMB_CAPACITY = 4,      /* cap of the hierarchy kernel's hand-over list (0 = default); test hook for its overflow path */
    QSMRT_OPT_SORT_VARIANT = 5,        /* 0 = histogram / scan / scatter per radix pass, 1 = one kernel per pass with decoupled look-back (default) */
    QSMRT_OPT_SPLIT_MAX = 6,           /* sliver splitting: a long thin triangle enters the build as up to this many references with tight slab boxes (default 8; 1 = off) */
    QSMRT_OPT_SPLIT_ASPECT = 7,        /* ... one reference per this many units of its aspect L^2 / 2A (default 2.0) */
    /* traversal (results are identical for every setting) */
    QSMRT_OPT_QUANTISED_NODES = 16,    /* 0 forces the 64-byte fp32 nodes even where the 32-byte nodes qualify */
    QSMRT_OPT_TRAVERSAL_VARIANT = 17,  /* 1 = one independent loop per thread (simple reference), 2 = persistent warp-uniform kernel (default) */
    QSMRT_OPT_REFILL = 18,             /* persistent kernel: idle lanes that trigger a refill (default 12) */
    QSMRT_OPT_WANT = 19,               /* ... node phase ends below this many searching lanes (default 16) */
    QSMRT_OPT_TRI_MIN = 20,            /* ... triangle-phase early-exit threshold (default 1) */
    QSMRT_OPT_COUNTERS = 21,           /* 1: cast_rays launches count what they fetch (qsmrt_scene_get_counters) */
    QSMRT_OPT_NODE_PATH = 22,          /* node fetch path: 0 = 256-bit LSU loads (default), 1 = texture, 2 = half and half */
    QSMRT_OPT_CP_WARP_MAX = 23,        /* closest-point batches up to this many queries use one warp per query (default 16384) */
    QSMRT_OPT_CTAS_PER_SM = 24,        /* persistent kernel: cap on resident CTAs per SM (0 = as many as fit, the default) */
    QSMRT_OPT_HOST_CHUNK = 28,         /* *_host calls: rays per stage of the three-stream pipe; 0 (default) = 1M when the results are the larger transfer, 2M when the rays are */
    QSMRT_OPT_HOST_RAMP = 29,          /* *_host calls: 1 (default) = short stages (chunk/8, /4, /2) at both ends of a long batch, 0 = equal stages */
    QSMRT_OPT_POINT_ORDER = 27,        /* qsmrt_sky_visibility: 1 = query points worked in Morton order (default), 0 = as stored */
    QSMRT_OPT_TILE_ORDER = 26,         /* grid-shaped batches: 1 = tile rows handed out from the middle of the grid outwards (default), 0 = in memory order */
    QSMRT_OPT_COUNT_SET = 25           /* count / list_intersections: entries of the per-lane hit set in shared memory, 4..32 (default 32); rays with more distinct hits are finished exactly by the slow path */
};
int qsmrt_scene_set_option(qsmrt_scene *scene, int key, double value);
int qsmrt_scene_get_option(qsmrt_scene *scene, int key, double *value);
/* The 16 counters of the scene's last counted cast_rays launch (QSMRT_OPT_COUNTERS): [0] node records fetched,
 * [1] triangle tests, [2] node-phase iterations, [3..6] lanes per iteration that step / are idle / hold a second
 * leaf / finished descending, [7] triangle-phase iterations, [8] lanes testing a triangle.  Synchronises. */
int qsmrt_scene_get_counters(qsmrt_scene *scene, uint64_t out[16]);

/* Scene files (the reference pickles its built search structures next to the
 * data: pyQSM/utils/io.py:44-60, tree_isolation.py:114,136).  save writes the
 * geometries as added and, with QSMRT_SAVE_BVH, the committed LBVH (traversal
 * nodes, quantised twin, triangle records, order, keys, grid); load creates a
 * new scene from the file -- committed if the file holds the BVH, otherwise
 * built by the first query (the build is deterministic: same tree either way).
 * load range-checks the geometry indices and every reference of a stored BVH
 * (children, leaf ranges, order, primitive ids), so a damaged or mismatched file
 * is an error, not a device fault; it does not defend against a crafted file
 * (e.g. cyclic child references) -- treat scene files like the pickles they replace. */
#define QSMRT_SAVE_BVH 1u
int qsmrt_scene_save(qsmrt_scene *scene, const char *path, uint32_t flags);
int qsmrt_scene_load(int cuda_device, const char *path, qsmrt_scene **out);

/* scene.add_triangles(mesh)  -- ray_casting.py:66,156,219,242,276,317.
 * Copies V x 3 float32 positions and T x 3 uint32 indices (Open3D copies
 * too); rejects an index >= V.  geom_id_out receives 0, 1, ... */
int qsmrt_add_triangles(qsmrt_scene *scene, const float *verts, uint64_t V,
                        const uint32_t *idx, uint64_t T, int ptrs_on_device,
                        uint32_t *geom_id_out);

/* Cylinder-QSM ingestion on the device: n records of 8
 * float32 (centre xyz, axis xyz, radius, height -- the cyl_details of
 * pyQSM/qsm_generation.py:171-178) become ONE geometry of n closed cylinders
 * with Open3D's create_cylinder topology (resolution 20 / split 4 in the
 * reference, point_cloud_processing.py:274-279), rotated from +z onto the
 * axis and moved to the centre as get_shape() does (:266-304).  Primitive id
 * = cylinder * (2*res*(1+split)) + local triangle. */
int qsmrt_add_cylinders(qsmrt_scene *scene, const float *records, uint64_t n, uint32_t resolution,
                        uint32_t split, int ptr_on_device, uint32_t *geom_id_out);
/* Size of / device copy of a registered geometry (vertices V x 3 float32, indices T x 3 uint32). */
int qsmrt_geometry_size(qsmrt_scene *scene, uint32_t geom_id, uint64_t *V_out, uint64_t *T_out);
int qsmrt_copy_geometry(qsmrt_scene *scene, uint32_t geom_id, float *verts_dev, uint32_t *idx_dev, void *stream);

/* Embree rtcCommitScene, which Open3D runs lazily on the first query.  Builds
 * the LBVH (Morton codes, radix sort, Karras hierarchy, refit, leaf collapse).
 * Queries call it implicitly; calling it directly lets the build be timed.
 * build_ms_out (may be NULL) is the device time of the build; synchronises. */
int qsmrt_commit(qsmrt_scene *scene, void *stream, float *build_ms_out);

/* scene.cast_rays(rays)  -- ray_casting.py:223,231,279,319.
 * Closest hit.  t_hit[N] (inf on miss), geometry_ids[N], primitive_ids[N]
 * (QSMRT_INVALID_ID on miss), primitive_uvs[N x 2], primitive_normals[N x 3]
 * (0 on miss). */
int qsmrt_cast_rays(qsmrt_scene *scene, const float *rays_dev, uint64_t N,
                    float *t_hit, uint32_t *geometry_ids, uint32_t *primitive_ids,
                    float *primitive_uvs, float *primitive_normals, void *stream);

/* cast_rays for an image / grid shaped batch rays[height][width][6] (what
 * create_rays_pinhole returns, ray_casting.py:222-223,277-279, and the
 * parallel grids of :159-165): identical results, but warps take 8 x 4 pixel
 * tiles instead of 32 consecutive rays, which keeps them coherent. */
int qsmrt_cast_rays_2d(qsmrt_scene *scene, const float *rays_dev, uint32_t width, uint64_t height,
                       float *t_hit, uint32_t *geometry_ids, uint32_t *primitive_ids,
                       float *primitive_uvs, float *primitive_normals, void *stream);

/* Same call with HOST buffers: chunks the batch and overlaps host->device,
 * traversal and device->host copies on three streams.  Synchronises. */
int qsmrt_cast_rays_host(qsmrt_scene *scene, const float *rays_host, uint64_t N,
                         float *t_hit, uint32_t *geometry_ids, uint32_t *primitive_ids,
                         float *primitive_uvs, float *primitive_normals);

/* cast_rays with HOST rays where every result either goes to host memory
 * (host_out[k]), stays on the device for a later fetch (dev_out[k], a
 * full-length device array; wins over host_out[k]) or is skipped (both NULL).
 * k = 0..4: t_hit, geometry_ids, primitive_ids, primitive_uvs,
 * primitive_normals.  The reference only reads t_hit and primitive_ids
 * (ray_casting.py:280-289,320-322): 8 instead of 32 bytes per ray cross PCIe. */
int qsmrt_cast_rays_host_split(qsmrt_scene *scene, const float *rays_host, uint64_t N,
                               void *const host_out[5], void *const dev_out[5]);

/* scene.count_intersections(rays)  -- the engine under list_intersections
 * (ray_casting.py:168) and compute_occupancy (:69).  counts[N] int32.
 * Never synchronises: a ray with more distinct hits than the on-chip set
 * holds is finished exactly by a second kernel (one traversal per hit). */
int qsmrt_count_intersections(qsmrt_scene *scene, const float *rays_dev, uint64_t N,
                              int32_t *counts, void *stream);

/* scene.test_occlusions(rays, tnear, tfar): out[N] = 1 iff any hit with
 * tnear < t <= tfar. */
int qsmrt_test_occlusions(qsmrt_scene *scene, const float *rays_dev, uint64_t N,
                          float tnear, float tfar, uint8_t *out, void *stream);

/* The two calls above with HOST buffers, through the same three-stream pipe as
 * qsmrt_cast_rays_host.  Synchronise. */
int qsmrt_count_intersections_host(qsmrt_scene *scene, const float *rays_host, uint64_t N, int32_t *counts);
int qsmrt_test_occlusions_host(qsmrt_scene *scene, const float *rays_host, uint64_t N,
                               float tnear, float tfar, uint8_t *out);

/* scene.list_intersections(rays)  -- ray_casting.py:168.  Two phases:
 *   _count  fills ray_splits[N+1] (int64, exclusive scan of the per-ray
 *           counts) and returns the total K in *total_out (synchronises);
 *   _fill   writes the K hits, per ray sorted by (t, geometry, primitive):
 *           ray_ids[K] int64, t_hit[K], geometry_ids[K], primitive_ids[K],
 *           primitive_uvs[K x 2] (synchronises).
 * The rays are traversed ONCE, in _count, which also collects the hit records
 * in device memory owned by the scene (20 bytes per hit + 8 per ray, returned
 * by _fill); _fill only moves them into the caller's arrays and must follow
 * _count on the same, unchanged rays. */
int qsmrt_list_intersections_count(qsmrt_scene *scene, const float *rays_dev, uint64_t N,
                                   int64_t *ray_splits, int64_t *total_out, void *stream);
int qsmrt_list_intersections_fill(qsmrt_scene *scene, const float *rays_dev, uint64_t N,
                                  const int64_t *ray_splits, int64_t *ray_ids, float *t_hit,
                                  uint32_t *geometry_ids, uint32_t *primitive_ids,
                                  float *primitive_uvs, void *stream);

/* On-device generator for the parallel-ray grids of the sun / rain drivers
 * (the pattern of ray_casting.py:159-165): ray(i,j) has origin
 * origin0 + i*du + j*dv (i < nu fastest) and direction dir. */
int qsmrt_gen_parallel_rays(float *rays_dev, uint64_t nu, uint64_t nv,
                            const float origin0[3], const float du[3], const float dv[3],
                            const float dir[3], void *stream);

/* RaycastingScene.create_rays_pinhole(intrinsic, extrinsic, w, h)
 * -- ray_casting.py:222,230,277,318.  rays_dev[h x w x 6]. */
int qsmrt_gen_pinhole_rays(float *rays_dev, uint32_t width_px, uint32_t height_px,
                           const double intrinsic[9], const double extrinsic[16], void *stream);

/* Post-processing of ray_casting.py:285-289 on the device: marks the
 * primitives (and their three vertices) that own a closest hit.
 * tri_hit[T of geometry 0..] / vert_hit[V] are uint8 flags, OR-ed in. */
int qsmrt_mark_hit_primitives(qsmrt_scene *scene, const uint32_t *geometry_ids,
                              const uint32_t *primitive_ids, uint64_t N,
                              uint8_t *tri_hit, uint8_t *vert_hit, void *stream);

/* Per-triangle exposure: tri_counts[t] += number of rays whose closest hit is
 * triangle t (scene order: geometry offsets + primitive id).  The reduction
 * the sun-sweep driver keeps on the device instead of 32 B/ray of results. */
int qsmrt_accumulate_hits(qsmrt_scene *scene, const uint32_t *geometry_ids,
                          const uint32_t *primitive_ids, uint64_t N,
                          uint32_t *tri_counts, void *stream);

/* Per-vertex exposure from per-triangle exposure (ray_casting.py:289-292:
 * hit_tris = triangles[prim_ids]; hit_vert_ids = np.unique(hit_tris)):
 * vert_counts[v] += tri_counts[t] for the three corners of every triangle t
 * (scene order, vertices numbered through the geometries in the order added). */
int qsmrt_vertex_exposure(qsmrt_scene *scene, const uint32_t *tri_counts, uint32_t *vert_counts, void *stream);

/* scene.compute_closest_points / compute_distance  (Open3D; the engine under
 * compute_signed_distance, ray_casting.py:250,255).  query points [N x 3]
 * float32 on the device.  closest[N x 3], distance[N] (inf for an empty
 * scene), ids, uv[N x 2] (u <-> v1, v <-> v2), normals[N x 3]; any NULL. */
int qsmrt_closest_points(qsmrt_scene *scene, const float *points_dev, uint64_t N,
                         float *closest, float *distance, uint32_t *geometry_ids,
                         uint32_t *primitive_ids, float *primitive_uvs,
                         float *primitive_normals, void *stream);

/* scene.compute_signed_distance(query_points)  -- ray_casting.py:250,255:
 * distance[N], negative where compute_occupancy is 1 (odd intersection count
 * along (1,1,1)).  Synchronises (frees its scratch). */
int qsmrt_signed_distance(qsmrt_scene *scene, const float *points_dev, uint64_t N,
                          float *distance, void *stream);

/* Environmental drivers (README.md:127 "sunlight angle, cloud cover and rain
 * angle"; data/notes/methods.md:16,53-55): the rays are generated inside the
 * traversal kernel and the results reduced on the device, so no 24 B/ray of
 * input or 32 B/ray of output ever touches HBM.
 *
 * qsmrt_sun_exposure: the parallel grid of qsmrt_gen_parallel_rays (nu x nv)
 *   is cast and tri_counts[t] += 1 for the triangle owning each closest hit
 *   (scene order).  Equals gen_parallel_rays + cast_rays + accumulate_hits.
 * qsmrt_sky_visibility: for each of n_points query points (origin = point +
 *   offset * normal; normals may be NULL) directions dir_begin ..
 *   dir_begin+dir_count-1 of a seeded uniform sample of the upper hemisphere
 *   (about +z) are tested for occlusion; unoccluded[p] += number of free
 *   directions.  Gap fraction = unoccluded / directions.  Direction k of
 *   point p is a hash of (seed, point_base + p, k): a block of points cut
 *   out of a larger set (a rank's share, a test's subsample) sees the same
 *   directions when point_base is the block's offset in that set.
 * qsmrt_gen_hemisphere_rays materialises exactly those rays
 *   (rays[n_points * dir_count][6]) for inspection and parity tests. */
int qsmrt_sun_exposure(qsmrt_scene *scene, uint64_t nu, uint64_t nv, const float origin0[3],
                       const float du[3], const float dv[3], const float dir[3],
                       uint32_t *tri_counts, void *stream);
/* A whole sweep in ONE launch: n_grids parallel grids, all nu x nv, grid a
 * given by grids_host[a][12] = origin0, du, dv, dir (ordinary host memory).
 * count_stride = 0 adds every grid into tri_counts[T]; otherwise grid a adds
 * into tri_counts[a * count_stride + t].  Same counts as n_grids calls of
 * qsmrt_sun_exposure, without their kernel tails and host round trips.
 * Sweeps of one scene must be issued on one stream. */
int qsmrt_sun_exposure_sweep(qsmrt_scene *scene, uint32_t n_grids, const float *grids_host,
                             uint64_t nu, uint64_t nv, uint32_t *tri_counts, uint64_t count_stride,
                             void *stream);
int qsmrt_sky_visibility(qsmrt_scene *scene, const float *points_dev, const float *normals_dev,
                         uint64_t n_points, uint64_t point_base, uint64_t seed, float offset, uint32_t dir_begin,
                         uint32_t dir_count, uint32_t *unoccluded, void *stream);
int qsmrt_gen_hemisphere_rays(float *rays_dev, const float *points_dev, const float *normals_dev,
                              uint64_t n_points, uint64_t point_base, uint64_t seed, float offset, uint32_t dir_begin,
                              uint32_t dir_count, void *stream);

/* "Raycasting projection" of data/notes/methods.md:53-55 and
 * epiphyte_isolation_methods.md:17 (the surf_2d branch of cast_rays,
 * ray_casting.py:285-301, iterated): cast the parallel grid, sum the area of
 * the triangles that own a closest hit (3-D and flattened along the ray
 * direction), remove them, repeat until the rays see nothing.
 * layer_of[T] (device, scene order, may be NULL) = layer in which each
 * triangle was removed, -1 if never hit; layer_stats (HOST, max_layers x 3)
 * = triangles, 3-D area, projected area per layer.  Synchronises per layer. */
int qsmrt_peel_projection(qsmrt_scene *scene, uint64_t nu, uint64_t nv, const float origin0[3],
                          const float du[3], const float dv[3], const float dir[3], int max_layers,
                          int32_t *layer_of, double *layer_stats_host, int *n_layers_out, void *stream);

int qsmrt_get_stats(qsmrt_scene *scene, qsmrt_stats *out);

/* Measurement utility (bench.py's roofline denominators): reads `bytes` of a
 * device buffer `reps` times with the 256-bit loads the traversal kernel uses,
 * on the current device.  Time it with events: a buffer well inside L2 gives
 * the L2 -> SM read peak, one far above it the HBM read rate. */
int qsmrt_util_read_sweep(const void *buf_dev, uint64_t bytes, uint32_t reps, uint32_t *sink_dev, void *stream);

/* Builder introspection (parity tests of the LBVH builder; host outputs,
 * any may be NULL).  keys[T] uint64 sorted Morton keys; order[T] sorted
 * position -> input triangle; nodes[(2T-1) x 8] float32/int32 words of the
 * 32-byte binary nodes (lo.xyz,left,hi.xyz,right; internal 0..T-2, leaves
 * after).  The builder only materialises the complete binary node array when
 * QSMRT_OPT_KEEP_BINARY_NODES was set at the commit (the product path writes
 * the traversal nodes straight from registers); asking for `nodes` otherwise
 * is an error. */
int qsmrt_debug_get_build(qsmrt_scene *scene, uint64_t *keys, uint32_t *order, void *nodes);

#ifdef __cplusplus
}
#endif
#endif /* QSMRT_H */
