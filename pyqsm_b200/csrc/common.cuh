// common.cuh -- device data layouts and the shared fp32 triangle arithmetic.
// sm_100a only.  The arithmetic here is the bit-for-bit twin of
// oracle/qsmrt_oracle.c (explicit __fmaf_rn in the same association, IEEE
// divide/sqrt); this TU is compiled with -fmad=false so nothing else fuses.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define QSMRT_LEAF_MAX 4           // triangles per collapsed leaf
#define QSMRT_INVALID 0xFFFFFFFFu

// 32-byte binary LBVH node (Karras topology, refit boxes).  Unified index
// space: internal nodes 0..n-2 (root 0), leaf j at (n-1)+j.
struct __align__(16) BNode {
    float lox, loy, loz; int left;   // leaves: left = right = -1
    float hix, hiy, hiz; int right;
};

// 64-byte traversal node: both child boxes in one record (4 x 16-byte loads).
//   a = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//   b = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   c = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//   d = (child0, child1, -, -)   child >= 0: node index
//                                child <  0: leaf, ~child = first<<2 | (count-1)
struct __align__(64) TNode { float4 a, b, c; int4 d; };

// 32-byte quantised traversal node: the same two child boxes on a 16-bit grid
// spanning the padded scene bounds (plane = glo + q * cell), one 256-bit load.
//   w[0..2] = child 0 x,y,z (lo | hi << 16), w[3..5] = child 1, w[6], w[7] = child refs.
// Boxes are widened by 3 cells per side at encode time, which covers the decode
// rounding (t = fma(2^23 + q, cell/d, const), <= ~1.1 cells) -- still conservative.
struct __align__(32) QNode { uint32_t w[8]; };

// 48-byte triangle record (3 x 16-byte loads), Embree TriangleM convention.
struct __align__(16) TriRec {
    float4 p0;   // v0.xyz, primitive id bits
    float4 p1;   // e1 = v0 - v1, geometry id bits
    float4 p2;   // e2 = v2 - v0, unused
};

struct SceneView {
    const TNode  *nodes;
    const TriRec *tris;
    uint32_t      ntris;
    uint32_t      height;      // binary LBVH height: bound on the traversal stack depth
    cudaTextureObject_t node_tex;   // float4 view of nodes (TEX-path experiment), 0 if absent
    const QNode  *qnodes;      // 32-byte quantised nodes, nullptr when the grid is too coarse for this scene
    float         glo[3], cell[3];
};

#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    qsmrt_set_error("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

void qsmrt_set_error(const char *fmt, ...);

struct f3 { float x, y, z; };

__device__ __forceinline__ f3 f3cross(f3 a, f3 b) {
    f3 r;
    r.x = __fmaf_rn(a.y, b.z, -__fmul_rn(a.z, b.y));
    r.y = __fmaf_rn(a.z, b.x, -__fmul_rn(a.x, b.z));
    r.z = __fmaf_rn(a.x, b.y, -__fmul_rn(a.y, b.x));
    return r;
}
__device__ __forceinline__ float f3dot(f3 a, f3 b) {
    return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, __fmul_rn(a.z, b.z)));
}
__device__ __forceinline__ float fxor(float f, uint32_t m) {
    return __uint_as_float(__float_as_uint(f) ^ m);
}

struct MtHit { float U, V, T, absDen; f3 Ng; };

// Embree MoellerTrumboreIntersector1 (edges inclusive, no culling):
// accept iff den != 0, U >= 0, V >= 0, U+V <= |den|, |den|*tnear < T <= |den|*tfar.
__device__ __forceinline__ bool mt_test(const float4 p0, const float4 p1, const float4 p2,
                                        f3 O, f3 D, float tnear, float tfar, MtHit &h)
{
    f3 v0 = { p0.x, p0.y, p0.z }, e1 = { p1.x, p1.y, p1.z }, e2 = { p2.x, p2.y, p2.z };
    f3 Ng = f3cross(e2, e1);
    f3 C = { __fsub_rn(v0.x, O.x), __fsub_rn(v0.y, O.y), __fsub_rn(v0.z, O.z) };
    f3 R = f3cross(C, D);
    float den = f3dot(Ng, D);
    float absDen = fabsf(den);
    uint32_t sgn = __float_as_uint(den) & 0x80000000u;
    float U = fxor(f3dot(R, e2), sgn);
    float V = fxor(f3dot(R, e1), sgn);
    float T = fxor(f3dot(Ng, C), sgn);
    bool ok = (den != 0.0f) & (U >= 0.0f) & (V >= 0.0f) & (__fadd_rn(U, V) <= absDen);
    ok = ok & (__fmul_rn(absDen, tnear) < T) & (T <= __fmul_rn(absDen, tfar));
    h.U = U; h.V = V; h.T = T; h.absDen = absDen; h.Ng = Ng;
    return ok;
}

struct Ray { f3 O, D; float idx, idy, idz, oodx, oody, oodz; };

__device__ __forceinline__ float safe_inv(float d) {
    const float tiny = 8.271806125530277e-25f;   // 2^-80: 0 * inv is never NaN
    float a = fabsf(d) < tiny ? copysignf(tiny, d) : d;
    return __frcp_rn(a);        // correctly rounded, i.e. the same value as 1.0f / a, in fewer instructions
}

__device__ __forceinline__ Ray load_ray(const float *__restrict__ rays, uint64_t i) {
    // 24-byte rows: three 8-byte loads (rows are 8-byte aligned when the base is)
    const float2 *p = reinterpret_cast<const float2 *>(rays + 6 * i);
    float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    Ray r;
    r.O = { a.x, a.y, b.x }; r.D = { b.y, c.x, c.y };
    r.idx = safe_inv(r.D.x); r.idy = safe_inv(r.D.y); r.idz = safe_inv(r.D.z);
    r.oodx = r.O.x * r.idx; r.oody = r.O.y * r.idy; r.oodz = r.O.z * r.idz;
    return r;
}
