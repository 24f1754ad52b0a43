// build.h -- host-side interface between the scene object and the LBVH builder.
#pragma once
#include "common.cuh"

struct BuildParams {        // device-resident; filled by the bounds kernels
    float slo[3], shi[3];   // scene bounds over referenced vertices
    float scale[3];         // 2^21 / extent per axis (0 for a flat axis)
    float pad;              // absolute AABB padding
    float glo[3], cell[3], inv_cell[3];   // 16-bit quantisation grid of the 32-byte nodes
    float leaf_diag_sum;    // sum of leaf box diagonals (mean leaf size decides whether the grid is fine enough)
    int   use_q;            // decided on the device (k_hierarchy_refit_emit): write / read the 32-byte nodes
};

struct LbvhBuildArgs {
    const float    *verts;          // [V,3] concatenated
    const uint32_t *idx;            // [T,3] rebased into verts
    uint64_t        ntris;
    const void     *refs;           // RefBox[nrefs] when slivers were split (lbvh_split_emit), else null: one leaf per triangle
    uint64_t        nrefs;
    const uint64_t *geom_offsets;   // [ngeoms+1] first triangle of each geometry
    uint32_t        ngeoms;
    int             leaf_max;       // 1..QSMRT_LEAF_MAX triangles per collapsed leaf
    int             sort_variant;   // 0 classic (3 kernels per pass, 8 passes), 1 onesweep (decoupled look-back; ships)
    int             climb_capacity; // > 0: cap of the climb work list (test hook for its overflow path)
    float           quant_frac;     // 6 grid cells <= this share of the mean leaf diagonal -> 32-byte nodes
    // outputs / scratch (device)
    const BuildParams *params_host; // finished on the host (lbvh_finalize_params) from the geometries' bounds
    BuildParams *params;            // device copy; the build adds leaf_diag_sum / use_q
    uint64_t   *keys, *keys_tmp;    // [L]  (L = nrefs when refs, else ntris; likewise below) the sorted arrays end up in (keys, order) or, when *result_in_tmp, in
    uint32_t   *order, *order_tmp;  // [T]  (keys_tmp, order_tmp): the caller keeps that pair and recycles the other
    int        *result_in_tmp;
    uint32_t   *sort_scratch;       // lbvh_sort_scratch_bytes(L); counters | flags | sort_scratch are ONE block (lbvh_zero_block_bytes), in that order
    BNode      *bnodes;             // [2T-1] hand-over boxes of the global phase; the complete binary tree iff keep_bnodes
    int         keep_bnodes;
    unsigned long long *flags;      // [T-1] hand-over slot of each split (k_hierarchy_refit_emit)
    TriRec     *tris;               // [T]
    TNode      *tnodes;             // [max(T-1,1)]
    QNode      *qnodes;             // [max(T-1,1)] quantised twin of tnodes
    void       *climb_work;         // lbvh_climb_bytes(T): subtrees handed to the climb kernel
    unsigned long long *counters;   // [8] (the start of the zero block) nodes emitted, leaves emitted, binary tree height, climb list length, sort overflow
    int         full_sort;          // all eight radix passes (after a sort overflow: a run of > 64 keys equal in their top 40 bits)
    cudaEvent_t ev_sort0, ev_sort1; // optional
};

size_t lbvh_sort_scratch_bytes(uint64_t n);
size_t lbvh_zero_block_bytes(uint64_t n);
size_t lbvh_climb_bytes(uint64_t n);
int lbvh_radix_sort(uint64_t *keys, uint64_t *keys_tmp, uint32_t *vals, uint32_t *vals_tmp,
                    uint64_t n, uint32_t *scratch, cudaStream_t st, unsigned long long *overflow, int sort_variant, bool hist_done,
                    int *result_in_tmp);
size_t lbvh_ref_bytes(uint64_t nrefs);
int lbvh_split_count(const float *verts, const uint32_t *idx, uint64_t T, int split_max, float split_aspect, int32_t *counts,
                     unsigned long long *total_dev, cudaStream_t st);
int lbvh_split_emit(const float *verts, const uint32_t *idx, uint64_t T, int split_max, float split_aspect, const int64_t *ref_off,
                    void *refs, cudaStream_t st);
void lbvh_finalize_params(const float lo[3], const float hi[3], BuildParams *out);
int lbvh_geometry_stats(const float *verts, uint64_t V, const uint32_t *idx, uint64_t T, uint32_t *out7_dev, float lo[3], float hi[3],
                        uint32_t *max_index, cudaStream_t st);
int lbvh_build(const LbvhBuildArgs &args, cudaStream_t st);
