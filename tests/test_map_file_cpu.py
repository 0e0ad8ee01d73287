""".sqwm v11 reader / writer of the host mirror (square_wheel_b200/host/map_file.cpp) against the layout the reference writes
(src/common/bsp_map_save_load.rs) and the quantisation it applies (src/common/lightmap_compression.rs)."""
import struct

import numpy as np
import pytest

from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S

KW = dict(seed=5, texture_size_log2=5, num_materials=6, light_grid=1)
# lump index -> element size (bsp_map_save_load.rs:246-264 with the #[repr(C)] sizes of bsp_map_compact.rs / lightmap_compression.rs)
LUMP_ELEMENT_SIZES = {0: 24, 1: 16, 2: 80, 3: 32, 4: 4, 5: 12, 6: 64, 7: 12, 8: 32, 9: 12, 10: 8, 11: 1, 12: 4, 13: 13, 14: 36, 15: 12, 16: 32}


@pytest.fixture(scope="module")
def saved(tmp_path_factory):
    sc = S.Scene(**KW)
    path = str(tmp_path_factory.mktemp("maps") / "synthetic.sqwm")
    sc.save_map(path)
    yield sc, path
    sc.close()


def _header(path):
    raw = open(path, "rb").read()
    ident, version = raw[:4], struct.unpack_from("<I", raw, 4)[0]
    lumps = [struct.unpack_from("<III", raw, 8 + 12 * i) for i in range(32)]
    return raw, ident, version, lumps


def test_header_and_lump_table_follow_the_reference_layout(saved):
    sc, path = saved
    raw, ident, version, lumps = _header(path)
    assert ident == b"SqwM" and version == 11
    offset = 392  # size_of::<BspMapHeader>() = 4 + 4 + 32 * 12
    for i in range(17):
        lump_offset, element_size, count = lumps[i]
        assert element_size == LUMP_ELEMENT_SIZES[i], f"lump {i}"
        assert lump_offset == offset, f"lump {i} is written right after the previous one"
        offset += element_size * count
    assert offset == len(raw)
    assert all(l == (0, 0, 0) for l in lumps[17:])
    info = sc.info
    assert [lumps[i][2] for i in (0, 1, 2, 3, 5)] == [info.num_nodes, info.num_leafs, info.num_polygons, info.num_portals, info.num_vertices]
    assert lumps[12][2] == lumps[13][2] == info.num_lightmap_elements and lumps[14][2] == 1 and lumps[6][2] == info.num_materials
    # the polygon lump is the raw 80-byte #[repr(C)] Polygon array: texture index is the last u32 of every element
    first_polygon_texture = struct.unpack_from("<I", raw, lumps[2][0] + 76)[0]
    assert first_polygon_texture < info.num_materials
    assert raw[lumps[6][0] : lumps[6][0] + 64].rstrip(b"\0") == b"synthetic/material_00" and sc.texture_name(0) == "synthetic/material_00"


def test_round_trip_quantises_like_lightmap_compression(saved):
    sc, path = saved
    other = S.Scene(**KW)
    try:
        before, before_dir = sc.lightmaps.copy(), sc.directional_lightmaps.copy()
        other.load_map(path)
        assert other.info.num_polygons == sc.info.num_polygons and other.info.num_lightmap_elements == sc.info.num_lightmap_elements
        # CompressedColor (:48-66): scale = clamp(ceil(max), 1, 255), components = trunc(c * 255 / scale) -> error < scale / 255, never above
        scale = np.clip(np.ceil(before.max(axis=1)), 1.0, 255.0)[:, None]
        want = np.trunc(np.clip(before * (np.float32(255.0) / scale), 0.0, 255.0)).astype(np.float32) * (scale / np.float32(255.0))
        np.testing.assert_allclose(other.lightmaps, want, rtol=0, atol=1e-6)
        assert np.all(other.lightmaps <= before + 1e-6) and np.all(before - other.lightmaps < scale / 255.0 + 1e-6)
        d = other.directional_lightmaps
        assert np.all(np.abs(d["directional_light_deviation"] - before_dir["directional_light_deviation"]) <= 1.0 / 255.0 + 1e-6)
        vmax = np.clip(np.ceil(np.abs(before_dir["light_direction_vector_scaled"]).max(axis=1)), 1.0, 255.0)[:, None]
        assert np.all(np.abs(d["light_direction_vector_scaled"] - before_dir["light_direction_vector_scaled"]) <= vmax / 127.0 + 1e-6)
        # saving the loaded map again keeps every lump but the quantised light data byte for byte (the scale byte is re-derived
        # from the de-quantised maximum, so light data is NOT idempotent — in the reference either)
        again = path + ".again"
        other.save_map(again)
        raw_a, _, _, lumps_a = _header(path)
        raw_b, _, _, lumps_b = _header(again)
        assert lumps_a == lumps_b
        for i in (0, 1, 2, 3, 4, 5, 6, 14, 15):
            o, es, n = lumps_a[i]
            assert raw_a[o : o + es * n] == raw_b[o : o + es * n], f"lump {i}"
        # the geometry is untouched: the host produces the same jobs and commands from the loaded map
        a, b = S.HostRenderer(sc), S.HostRenderer(other)
        try:
            pose = sc.camera_path(8)[3]
            fa, fb = a.build_frame(320, 180, pose), b.build_frame(320, 180, pose)
            assert fa.jobs.tobytes() == fb.jobs.tobytes() and fa.cmds.tobytes() == fb.cmds.tobytes() and fa.polygon_points.tobytes() == fb.polygon_points.tobytes()
        finally:
            a.close()
            b.close()
    finally:
        other.close()


def test_bad_files_are_rejected_like_the_reference(saved, tmp_path):
    sc, path = saved
    raw = bytearray(open(path, "rb").read())
    other = S.Scene(**KW)
    try:
        bad_id = tmp_path / "bad_id.sqwm"
        bad_id.write_bytes(b"IBSP" + bytes(raw[4:]))
        with pytest.raises(OSError, match="not a Square-Wheel map"):
            other.load_map(str(bad_id))
        bad_version = tmp_path / "bad_version.sqwm"
        bad_version.write_bytes(bytes(raw[:4]) + struct.pack("<I", 10) + bytes(raw[8:]))
        with pytest.raises(OSError, match="incompatible map version"):
            other.load_map(str(bad_version))
        with pytest.raises(OSError):
            other.load_map(str(tmp_path / "missing.sqwm"))
        truncated = tmp_path / "truncated.sqwm"
        truncated.write_bytes(bytes(raw[:300]))
        with pytest.raises(OSError):
            other.load_map(str(truncated))
        # a lump with another element size reads as EMPTY (bsp_map_save_load.rs:292-300): here the light grid samples vanish,
        # the rest of the map still loads
        odd = bytearray(raw)
        struct.pack_into("<I", odd, 8 + 12 * 16 + 4, 31)
        odd_path = tmp_path / "odd_lump.sqwm"
        odd_path.write_bytes(bytes(odd))
        other.load_map(str(odd_path))
        assert other._lib.sqwh_scene_num_light_grid_samples(other._handle) == 0 and other.info.num_polygons == sc.info.num_polygons
        # a map whose polygons refer to materials the scene does not have is refused, the scene keeps its map
        few = S.Scene(seed=5, texture_size_log2=5, num_materials=2)
        try:
            n = few.info.num_polygons
            with pytest.raises(ValueError, match="no material"):
                few.load_map(path)
            assert few.info.num_polygons == n
        finally:
            few.close()
    finally:
        other.close()


def test_light_grid_lights_models_and_sprites(saved):
    """ModelLighting::Default and Sprite::light_scale read the map's light grid (fetch_light_from_grid,
    triangle_models_rendering.rs:574-699): black without a grid, lit and inside the fixture's value range with it."""
    from common import sprites_ahead_of_camera

    lit = S.Scene(seed=5, texture_size_log2=5, num_materials=6, light_grid=1, num_models=3, model_detail=6)
    dark = S.Scene(seed=5, texture_size_log2=5, num_materials=6, light_grid=0, num_models=3, model_detail=6)
    a, b = S.HostRenderer(lit), S.HostRenderer(dark)
    try:
        pose = lit.camera_path(16)[2]
        az = float(pose[3])
        m = np.zeros(3, S.model_entity_dtype)
        for i in range(3):
            m[i]["position"] = pose[:3] + (60.0 + 20.0 * i) * np.array([-np.sin(az), np.cos(az), 0.0], np.float32)
            m[i]["rotation"] = (1.0, 0.0, 0.0, 0.0)
            m[i]["model"] = i
            m[i]["skin"] = i
        sp = sprites_ahead_of_camera(S, pose, 4, num_skins=3, light_scale=1.0)
        sp["light_add"] = 0.0
        fa = a.build_frame(480, 270, pose, models=m, sprites=sp)
        fb = b.build_frame(480, 270, pose, models=m, sprites=sp)
        assert fa.triangle_points is not None and fb.triangle_points is not None and len(fa.triangle_points) == len(fb.triangle_points)
        assert np.all(fb.triangle_points["light"] == 0)
        light = fa.triangle_points["light"].astype(np.float64) / 65536.0
        assert light.max() > 0.3 and light.max() < 4.0 and np.count_nonzero(light) > 0.9 * light.size
    finally:
        a.close()
        b.close()
        lit.close()
        dark.close()
