"""An INDEPENDENT check of the host mirror (square_wheel_b200/host/renderer.cpp: prepare_polygon_surface, draw_polygon, the job list),
which otherwise has no second implementation: everything here is recomputed in numpy / float64 from two things only — the map
as written to a .sqwm file (parsed here with the layout of bsp_map_save_load.rs) and the camera pose — and
compared with what the mirror hands to the GPU.

For every polygon draw command of a frame:
  * the command is matched to its map polygon through Polygon::lightmap_data_offset (unique per lit polygon),
  * each projected vertex (x, y) is un-projected in float64 onto the polygon's WORLD plane with a numpy restatement of
    build_view_matrix / complete_view_matrix (common/matrix.rs:13-110),
  * the map polygon's tex_coord_equation at that world point, divided by 2^mip and shifted by the surface origin, must equal
    what the command's screen-space equations give: tc_eq(x, y) / depth_eq(x, y)  (TexCoordEquation / DepthEquation,
    equations.rs:5-19; the common z_scale of draw_polygon, partial_renderer.rs:4106-4109, cancels in the ratio),
  * 1 / depth_eq(x, y) / z_scale must be the camera-space depth of the point,
  * the surface rect must contain every sampled texture coordinate (partial_renderer.rs:1558-1615) and the job must carry the
    polygon's lightmap size (common/lightmap.rs:19-27) and the mip that calculate_mip can justify (same mip on job and command).
"""
import struct

import numpy as np
import pytest

from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S

W, H = 640, 360


def _parse_sqwm(path):
    raw = open(path, "rb").read()
    assert raw[:4] == b"SqwM" and struct.unpack_from("<I", raw, 4)[0] == 11
    lumps = [struct.unpack_from("<III", raw, 8 + 12 * i) for i in range(32)]
    poly_dt = np.dtype(
        [("first_vertex", "<u4"), ("num_vertices", "<u4"), ("plane", "<f4", 4), ("tce", "<f4", (2, 4)), ("tc_min", "<i4", 2), ("tc_max", "<i4", 2), ("lightmap_data_offset", "<u4"), ("texture", "<u4")]
    )
    assert poly_dt.itemsize == 80 and lumps[2][1] == 80 and lumps[5][1] == 12
    polygons = np.frombuffer(raw, poly_dt, lumps[2][2], lumps[2][0])
    vertices = np.frombuffer(raw, "<f4", lumps[5][2] * 3, lumps[5][0]).reshape(-1, 3).astype(np.float64)
    return polygons, vertices


def _view_matrix(pose, fov, width, height):
    """common/matrix.rs build_view_matrix + complete_view_matrix in float64: world point -> (x * w, y * w, w) with w = depth."""
    x, y, z, azimuth, elevation = [float(v) for v in pose]

    def quat_z(a):
        return np.array([np.cos(a / 2), 0.0, 0.0, np.sin(a / 2)])

    def quat_y(a):
        return np.array([np.cos(a / 2), 0.0, np.sin(a / 2), 0.0])

    def qmul(a, b):
        s1, v1, s2, v2 = a[0], a[1:], b[0], b[1:]
        return np.concatenate([[s1 * s2 - v1 @ v2], s1 * v2 + s2 * v1 + np.cross(v1, v2)])

    def qmat(q):  # rotation matrix of a unit quaternion
        s, (a, b, c) = q[0], q[1:]
        return np.array(
            [[1 - 2 * (b * b + c * c), 2 * (a * b - s * c), 2 * (a * c + s * b)], [2 * (a * b + s * c), 1 - 2 * (a * a + c * c), 2 * (b * c - s * a)], [2 * (a * c - s * b), 2 * (b * c + s * a), 1 - 2 * (a * a + b * b)]]
        )

    rotation = qmul(quat_z(azimuth + np.pi / 2), quat_y(-elevation))
    conj = rotation * np.array([1.0, -1.0, -1.0, -1.0])
    basis_change = np.array([[0.0, -1.0, 0.0], [0.0, 0.0, -1.0], [1.0, 0.0, 0.0]])  # columns (0,0,1), (-1,0,0), (0,-1,0)
    rot = basis_change @ qmat(conj)
    inv_half_fov_tan = 1.0 / np.tan(fov / 2)
    aspect = width / height
    persp = np.diag([inv_half_fov_tan / aspect, inv_half_fov_tan, 1.0])
    resize = np.diag([width / 2, height / 2, 1.0])
    shift = np.eye(3)
    shift[0, 2], shift[1, 2] = width / 2, height / 2
    m3 = shift @ resize @ persp @ rot
    return m3, np.array([x, y, z])


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    sc = S.Scene(seed=5, texture_size_log2=6, num_materials=8)
    path = str(tmp_path_factory.mktemp("map") / "independent.sqwm")
    sc.save_map(path)
    polygons, vertices = _parse_sqwm(path)
    yield sc, polygons, vertices
    sc.close()


def test_commands_and_jobs_against_a_numpy_restatement(world):
    sc, polygons, vertices = world
    by_lightmap = {int(p["lightmap_data_offset"]): i for i, p in enumerate(polygons) if p["lightmap_data_offset"] != 0}
    assert len(by_lightmap) > 0.9 * len(polygons)  # the key is unique
    hr = S.HostRenderer(sc, use_directional_lightmaps=True)
    checked = 0
    worst_tc = worst_depth = 0.0
    try:
        for pose in sc.camera_path(16):
            hf = hr.build_frame(W, H, pose)
            m3, cam = _view_matrix(pose, S.FOV_DEFAULT, W, H)
            inv_m3 = np.linalg.inv(m3)
            jobs_by_offset = {int(j["surface_pixels_offset"]): j for j in hf.jobs}
            assert np.allclose(hf.cam_pos, cam, atol=1e-4)
            for c in hf.cmds:
                assert c["kind"] == N.CMD_POLYGON and c["tex_kind"] == N.TEX_SURFACE
                j = jobs_by_offset[int(c["tex_ref"])]
                if j["lightmap_offset"] == 0:
                    continue
                p = polygons[by_lightmap[int(j["lightmap_offset"])]]
                mip = 4 - int(j["lightmap_scale_log2"])
                assert 0 <= mip <= 4
                # lightmap size of the polygon (common/lightmap.rs:19-27)
                lm = [((int(p["tc_max"][a]) + 15) >> 4) - (int(p["tc_min"][a]) >> 4) + 1 for a in (0, 1)]
                assert [int(x) for x in j["lightmap_size"]] == lm
                n, d = p["plane"][:3].astype(np.float64), float(p["plane"][3])
                pts = hf.polygon_points[int(c["vertex_offset"]) : int(c["vertex_offset"]) + int(c["num_vertices"])]
                xs, ys = pts["x"] / 65536.0, pts["y"] / 65536.0
                de, te = c["depth_eq"].astype(np.float64), c["tc_eq"].astype(np.float64)
                tex_w, tex_h = sc.images[int(j["texture_image"])][1].shape[1], sc.images[int(j["texture_image"])][1].shape[0]
                for x, y in zip(xs, ys):
                    # un-project: the world point P on the plane with m3 @ (P - cam) = w * (x, y, 1)
                    ray = inv_m3 @ np.array([x, y, 1.0])
                    w = (d - n @ cam) / (n @ ray)
                    assert w > 0
                    P = cam + ray * w
                    inv_z = de[0] * x + de[1] * y + de[2]
                    got = np.array([(te[0] * x + te[2] * y + te[4]) / inv_z, (te[1] * x + te[3] * y + te[5]) / inv_z])
                    world_tc = np.array([p["tce"][a][:3].astype(np.float64) @ P + float(p["tce"][a][3]) for a in (0, 1)])
                    # surface texel coordinates: world tc at this mip, relative to the surface origin (tc_start = surface_tc_min here:
                    # no animated texture shift), modulo the texture size when the mirror reduced the origin (tc_reduce, :1598-1615)
                    want = world_tc / (1 << mip) - j["tc_start"].astype(np.float64)
                    diff = got - want
                    diff -= np.round(diff / np.array([tex_w, tex_h])) * np.array([tex_w, tex_h])
                    worst_tc = max(worst_tc, float(np.abs(diff).max()))
                    assert np.abs(diff).max() < 0.05 + 2e-4 * np.abs(want).max(), (got, want)
                    # the surface rect contains the coordinate (half a texel of slack for the clipped outline's rounding)
                    assert -0.6 <= got[0] <= int(j["surface_size"][0]) + 0.6 and -0.6 <= got[1] <= int(j["surface_size"][1]) + 0.6
                    checked += 1
                # depth: depth_eq is 1/z up to ONE power-of-two scale per command (z_scale of draw_polygon)
                ratios = []
                for x, y in zip(xs, ys):
                    ray = inv_m3 @ np.array([x, y, 1.0])
                    w = (d - n @ cam) / (n @ ray)
                    ratios.append((de[0] * x + de[1] * y + de[2]) * w)
                ratios = np.array(ratios)
                scale = np.median(ratios)
                assert abs(np.log2(scale) - round(np.log2(scale))) < 1e-3  # a power of two
                worst_depth = max(worst_depth, float(np.abs(ratios / scale - 1.0).max()))
                assert np.abs(ratios / scale - 1.0).max() < 2e-3
    finally:
        hr.close()
    assert checked > 5000
    print(f"{checked} vertices checked: worst tc error {worst_tc:.4f} texels, worst relative depth error {worst_depth:.2e}")
