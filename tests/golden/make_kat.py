"""Writes tests/golden/kat.json: analytic known-answer vectors for the
RaycastingScene boundary.  Every expected
value is derived by hand / closed form in float64 here -- neither the oracle
nor the CUDA path is involved -- so the file pins both.

    python tests/golden/make_kat.py
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from pyqsm_b200 import synthetic as syn  # noqa: E402  (mesh topology only)

INV = 4294967295
INF = "inf"
kats = []

tri_v = [[0, 0, 0], [1, 0, 0], [1, 1, 0]]
tri_t = [[0, 1, 2]]

# KAT-1: Open3D test_raycasting_scene.py::test_cast_rays (as recalled)
kats.append(dict(name="kat1_cast_rays", mesh=dict(v=tri_v, t=tri_t),
                 rays=[[0.2, 0.1, 1, 0, 0, -1], [10, 10, 10, 1, 0, 0]],
                 cast=dict(t_hit=[1.0, INF], geometry_ids=[0, INV], primitive_ids=[0, INV],
                           primitive_uvs=[[0.1, 0.1], [0, 0]], primitive_normals=[[0, 0, 1], [0, 0, 0]]),
                 count=[1, 0], occluded=[True, False]))
# KAT-4: un-normalised direction scales t
kats.append(dict(name="kat4_unnormalised", mesh=dict(v=tri_v, t=tri_t),
                 rays=[[0.2, 0.1, 1, 0, 0, -2], [0.2, 0.1, 1, 0, 0, -0.25]],
                 cast=dict(t_hit=[0.5, 4.0], geometry_ids=[0, 0], primitive_ids=[0, 0],
                           primitive_uvs=[[0.1, 0.1], [0.1, 0.1]], primitive_normals=[[0, 0, 1], [0, 0, 1]]),
                 count=[1, 1], occluded=[True, True]))
# KAT-5: origin on the triangle plane -> t = 0 is excluded (tnear exclusive)
kats.append(dict(name="kat5_t0_excluded", mesh=dict(v=tri_v, t=tri_t),
                 rays=[[0.2, 0.1, 0, 0, 0, -1], [0.2, 0.1, 0, 0, 0, 1]],
                 cast=dict(t_hit=[INF, INF], geometry_ids=[INV, INV], primitive_ids=[INV, INV],
                           primitive_uvs=[[0, 0], [0, 0]], primitive_normals=[[0, 0, 0], [0, 0, 0]]),
                 count=[0, 0], occluded=[False, False]))
# KAT-6: back face hits (no culling); geometric normal not flipped; ray pointing away misses
kats.append(dict(name="kat6_backface", mesh=dict(v=tri_v, t=tri_t),
                 rays=[[0.2, 0.1, -1, 0, 0, 1], [0.2, 0.1, -1, 0, 0, -1]],
                 cast=dict(t_hit=[1.0, INF], geometry_ids=[0, INV], primitive_ids=[0, INV],
                           primitive_uvs=[[0.1, 0.1], [0, 0]], primitive_normals=[[0, 0, 1], [0, 0, 0]]),
                 count=[1, 0], occluded=[True, False]))

# KAT-2 / KAT-3: unit box; rays 0/1 run through face diagonals (shared edges)
bv, bt = syn.box_mesh()
kats.append(dict(name="kat2_box_count_list", mesh=dict(v=bv.tolist(), t=bt.tolist()),
                 rays=[[0.5, 0.5, -1, 0, 0, 1], [0.5, 0.5, 0.5, 0, 0, 1], [10, 10, 10, 1, 0, 0],
                       [0.25, 0.6, -1, 0, 0, 1]],
                 count=[2, 1, 0, 2], occluded=[True, True, False, True],
                 list=dict(ray_splits=[0, 2, 3, 3, 5], t_hit=[1.0, 2.0, 0.5, 1.0, 2.0]),
                 cast=dict(t_hit=[1.0, 0.5, INF, 1.0])))

# KAT-7: closed cylinder r=1 h=2 (create_cylinder topology, 20-gon)
cv, ct = syn.cylinder_mesh(1.0, 2.0, 20, 4)
s18, c18 = math.sin(math.radians(18)), math.cos(math.radians(18))
x = 1.0 - (0.05 / s18) * (1.0 - c18)           # where y = 0.05 meets the facet (1,0)-(cos18,sin18)
kats.append(dict(name="kat7_cylinder", mesh=dict(v=cv.tolist(), t=ct.tolist()),
                 rays=[[-5, 0.05, 0.3, 1, 0, 0], [0.3, 0.2, -5, 0, 0, 1], [0.3, 0.2, 0, 0, 0, 1], [3, 3, -5, 0, 0, 1]],
                 count=[2, 2, 1, 0], occluded=[True, True, True, False],
                 list=dict(ray_splits=[0, 2, 4, 5, 5], t_hit=[5 - x, 5 + x, 4.0, 6.0, 1.0]),
                 cast=dict(t_hit=[5 - x, 4.0, 1.0, INF])))

# KAT-9: Open3D test_raycasting_scene.py::test_compute_closest_points (as recalled): the triangle of KAT-1, one query
# above its interior and one beyond the vertex (1,1,0); uv convention p = (1-u-v) v0 + u v1 + v v2
kats.append(dict(name="kat9_closest_points", mesh=dict(v=tri_v, t=tri_t),
                 points=[[0.2, 0.1, 1], [10, 10, 10], [0.5, -2.0, 0.0]],
                 closest=dict(points=[[0.2, 0.1, 0], [1, 1, 0], [0.5, 0, 0]], geometry_ids=[0, 0, 0], primitive_ids=[0, 0, 0],
                              primitive_uvs=[[0.1, 0.1], [0, 1], [0.5, 0]], primitive_normals=[[0, 0, 1]] * 3),
                 distance=[1.0, math.sqrt(81 + 81 + 100), 2.0]))
# KAT-10: Open3D test_compute_distance / test_compute_occupancy / test_compute_signed_distance (as recalled): unit box,
# its centre and a point outside one corner; plus a point 0.25 under a face and one over an edge
kats.append(dict(name="kat10_box_distance_occupancy", mesh=dict(v=bv.tolist(), t=bt.tolist()),
                 points=[[0.5, 0.5, 0.5], [-0.5, -0.5, -0.5], [0.5, 0.5, 0.75], [0.5, 1.5, 1.5]],
                 distance=[0.5, math.sqrt(0.75), 0.25, math.sqrt(0.5)], occupancy=[1.0, 0.0, 1.0, 0.0],
                 signed_distance=[-0.5, math.sqrt(0.75), -0.25, math.sqrt(0.5)]))

# KAT-11: Open3D test_raycasting_scene.py::test_test_occlusions (as recalled): the KAT-1 rays with the interval moved
# around the hit at t = 1.  tnear is exclusive, tfar inclusive (Embree: tnear < t <= tfar).
kats.append(dict(name="kat11_occlusion_interval", mesh=dict(v=tri_v, t=tri_t),
                 rays=[[0.2, 0.1, 1, 0, 0, -1], [10, 10, 10, 1, 0, 0]],
                 occlusion_cases=[dict(tnear=0.0, tfar=INF, expect=[True, False]), dict(tnear=0.0, tfar=0.99, expect=[False, False]),
                                  dict(tnear=1.01, tfar=INF, expect=[False, False]), dict(tnear=0.0, tfar=1.0, expect=[True, False]),
                                  dict(tnear=1.0, tfar=INF, expect=[False, False]), dict(tnear=0.5, tfar=2.0, expect=[True, False])]))
# KAT-12: Open3D test_output_shapes (as recalled): results keep the leading shape of the rays, plus 2 / 3 for uv / normals
kats.append(dict(name="kat12_leading_shape", mesh=dict(v=tri_v, t=tri_t), shape=[2, 3],
                 rays=[[0.2, 0.1, 1, 0, 0, -1], [10, 10, 10, 1, 0, 0], [0.9, 0.5, 2, 0, 0, -1],
                       [0.2, 0.1, -3, 0, 0, 1], [0.5, 0.9, 1, 0, 0, -1], [0.6, 0.3, 1, 0, 0, -4]],
                 cast=dict(t_hit=[1.0, INF, 2.0, 3.0, INF, 0.25], primitive_ids=[0, INV, 0, 0, INV, 0], geometry_ids=[0, INV, 0, 0, INV, 0]),
                 count=[1, 0, 1, 1, 0, 1], occluded=[True, False, True, True, False, True]))
# KAT-13: two geometries (add_triangles returns 0 then 1; ray_casting.py:156 keeps that id): the KAT-1 triangle at z = 0
# and a copy at z = -1.  geometry_ids / primitive_ids of closest hits from above and below, both in list_intersections.
tri_v2 = [[0, 0, -1], [1, 0, -1], [1, 1, -1]]
kats.append(dict(name="kat13_two_geometries", mesh=dict(v=tri_v, t=tri_t), mesh2=dict(v=tri_v2, t=tri_t),
                 rays=[[0.2, 0.1, 1, 0, 0, -1], [0.2, 0.1, -3, 0, 0, 1], [0.2, 0.1, -0.5, 0, 0, 1], [0.2, 0.1, -0.5, 0, 0, -1]],
                 cast=dict(t_hit=[1.0, 2.0, 0.5, 0.5], geometry_ids=[0, 1, 0, 1], primitive_ids=[0, 0, 0, 0],
                           primitive_uvs=[[0.1, 0.1]] * 4, primitive_normals=[[0, 0, 1]] * 4),
                 count=[2, 2, 1, 1], occluded=[True, True, True, True],
                 list=dict(ray_splits=[0, 2, 4, 5, 6], t_hit=[1.0, 2.0, 2.0, 3.0, 0.5, 0.5], geometry_ids=[0, 1, 1, 0, 0, 1],
                           primitive_ids=[0] * 6, primitive_uvs=[[0.1, 0.1]] * 6)))
# KAT-14: list_intersections on the unit box with ids and uvs (ray_casting.py:168-180 reads primitive_ids / primitive_uvs):
# the ray at (0.25, 0.6) crosses the z = 0 and z = 1 faces in their second triangles (a, c, d) = prims 1 and 3,
# p = a + u (c - a) + v (d - a) = (u, u + v)  =>  u = 0.25, v = 0.35
kats.append(dict(name="kat14_box_list_ids", mesh=dict(v=bv.tolist(), t=bt.tolist()),
                 rays=[[0.25, 0.6, -1, 0, 0, 1], [0.25, 0.6, 0.5, 0, 0, -2]],
                 list=dict(ray_splits=[0, 2, 3], t_hit=[1.0, 2.0, 0.25], geometry_ids=[0, 0, 0], primitive_ids=[1, 3, 1],
                           primitive_uvs=[[0.25, 0.35]] * 3),
                 cast=dict(t_hit=[1.0, 0.25], primitive_ids=[1, 1], primitive_uvs=[[0.25, 0.35]] * 2, primitive_normals=[[0, 0, 1]] * 2)))

# KAT-8: pinhole camera (Open3D CreateRaysPinhole): fov 90, eye (0,0,-1) looking at the origin
kats.append(dict(name="kat8_pinhole", pinhole=dict(fov_deg=90, center=[0, 0, 0], eye=[0, 0, -1], up=[0, 1, 0],
                                                     width_px=4, height_px=2),
                 # f = 0.5*4/tan(45deg) = 2 ; dir = R^T K^-1 (x+.5, y+.5, 1) with R = I here (up x fwd = +x)
                 expect_dirs=[[[(xx + 0.5 - 2) / 2, (yy + 0.5 - 1) / 2, 1.0] for xx in range(4)] for yy in range(2)],
                 expect_origin=[0, 0, -1]))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
with open(out, "w") as f:
    json.dump(kats, f, indent=1)
print("wrote", out, len(kats), "vectors")
