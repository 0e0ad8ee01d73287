"""Map compiler + lightmapper of the host mirror (square_wheel_b200/host/map_compiler.cpp, lightmapper.cpp,
restating /root/reference/src/common/{map_file_q1, map_polygonizer, map_csg, bsp_builder, bsp_map_compact_conversion, map_lights,
light_trace, lightmaps_builder}.rs).  There is no reference output to compare with (no Rust toolchain: parity UNPINNED), so the
checks are the structural properties a compiled map must have for the renderer, closed forms of the direct-light formulas, and the
oracle drawing gap-free frames of the result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import map_compiler as mc  # noqa: E402

from square_wheel_b200 import _native as N  # noqa: E402
from square_wheel_b200 import scene as S  # noqa: E402


def box_brush(lo, hi, texture="wall", textures=None):
    """A Quake-format brush of six axis-aligned planes, three points each, wound so that (v0 - v1) x (v2 - v1) points out of the
    brush (map_polygonizer.rs:294-306 keeps the half-space dot(vec, p) <= dist)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    lines = ["{"]
    for axis in range(3):
        for sign in (-1, 1):
            n = np.zeros(3)
            n[axis] = sign
            u, v = np.zeros(3), np.zeros(3)
            u[(axis + 1) % 3] = 1
            v[(axis + 2) % 3] = 1
            base = np.where(n > 0, hi, lo) * np.abs(n) + lo * (1 - np.abs(n))
            p1 = base
            p0, p2 = base + u * 16, base + v * 16
            if np.dot(np.cross(p0 - p1, p2 - p1), n) < 0:
                p0, p2 = p2, p0
            tex = (textures or {}).get((axis, sign), texture)
            pts = " ".join("( %g %g %g )" % tuple(p) for p in (p0, p1, p2))
            lines.append(f"{pts} {tex} 0 0 0 1 1")
    lines.append("}")
    return "\n".join(lines)


def two_rooms_map(door_entity=True):
    """Two 256 x 256 x 128 rooms joined by a 64 x 96 doorway in the 32-unit wall between them, walls as slabs (like the demo map)."""
    b = []
    b.append(box_brush((-16, -16, -16), (560, 272, 0), "floor"))
    b.append(box_brush((-16, -16, 128), (560, 272, 144), "ceiling"))
    b.append(box_brush((-16, -16, 0), (0, 272, 128)))
    b.append(box_brush((544, -16, 0), (560, 272, 128)))
    b.append(box_brush((0, -16, 0), (544, 0, 128)))
    b.append(box_brush((0, 256, 0), (544, 272, 128)))
    b.append(box_brush((256, 0, 0), (288, 96, 128), "divider"))
    b.append(box_brush((256, 160, 0), (288, 256, 128), "divider"))
    b.append(box_brush((256, 96, 96), (288, 160, 128), "divider"))
    ents = ['{\n"classname" "worldspawn"\n// a comment\n' + "\n".join(b) + "\n}"]
    ents.append('{\n"classname" "info_player_start"\n"origin" "64 128 32"\n}')
    ents.append('{\n"classname" "light"\n"origin" "128 128 96"\n"light" "200"\n}')
    ents.append('{\n"classname" "info_teleport_destination"\n"origin" "416 128 32"\n}')
    if door_entity:
        ents.append('{\n"classname" "func_door"\n' + box_brush((400, 100, 0), (432, 132, 64), "crate") + "\n}")
    return "// Game: Generic\n" + "\n".join(ents) + "\n"


MATERIALS = {"floor": {}, "ceiling": {}, "wall": {}, "divider": {"normal_map": "x.png", "roughness": 0.3}, "crate": {"is_metal": True, "roughness": 0.2}}


@pytest.fixture(scope="module")
def compiled(tmp_path_factory):
    d = tmp_path_factory.mktemp("mapc")
    table = mc.materials_table(MATERIALS)
    path = str(d / "two_rooms.sqwm")
    stats = mc.compile_map(two_rooms_map(), table, path, perform_csg=True, advanced=True)
    lit = str(d / "two_rooms_lit.sqwm")
    lstats = mc.light_map(path, table, lit, sample_grid_size=2, ambient_light=0.125)
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    sc.load_compiled_map(lit, table, 6, 3)
    yield sc, stats, lstats, table, lit
    sc.close()


def _planes_ok(polys, verts, eps=1.0 / 16.0):
    for p in polys:
        v = verts[p["first_vertex"] : p["first_vertex"] + p["num_vertices"]]
        n = np.array(p["plane"][:3], np.float64)
        assert np.all(np.abs(v @ n - p["plane"][3]) <= eps * np.linalg.norm(n) + 1e-3)


def test_parser_rejects_broken_text(tmp_path):
    table = mc.materials_table(MATERIALS)
    good = two_rooms_map()
    for broken in (good.replace("(", "[", 1), good[: len(good) // 2], "hello"):
        with pytest.raises(RuntimeError):
            mc.compile_map(broken, table, str(tmp_path / "x.sqwm"))
    with pytest.raises(RuntimeError):
        mc.compile_map("", table, str(tmp_path / "x.sqwm"))  # no entities


def test_compiled_map_structure(compiled):
    sc, stats, _, _, _ = compiled
    polys, verts, _flags = sc.map_geometry()
    assert stats["initial_polygons"] == 9 * 6
    # CSG removes what lies inside neighbouring slabs; the BSP keeps only leafs reachable from the entities (the two rooms, the doorway)
    assert stats["polygons_after_csg"] < stats["initial_polygons"] + 20
    assert stats["leafs"] >= 3 and stats["leafs_removed"] >= 1 and stats["portals"] >= 2 and stats["submodels"] == 1
    assert sc.info.num_polygons == stats["polygons"] and sc.info.num_leafs == stats["leafs"]
    _planes_ok(polys, verts)
    # every lightmap within the splitter's limit (bsp_builder.rs:1282: 17, "actual size may be a little greater")
    for p in polys:
        for a in range(2):
            size = ((int(p["tex_coord_max"][a]) + 15) >> 4) - (int(p["tex_coord_min"][a]) >> 4) + 1
            assert 2 <= size <= 18
    # no outward-facing polygon survives: every polygon's plane keeps the player start or the other room's centre in front ... of
    # the faces that bound the rooms: here simply none of the outer faces of the shell (x = -16, x = 560, ...) may be left
    n = polys["plane"][:, :3]
    d = polys["plane"][:, 3]
    outer = [(np.array([-1, 0, 0]), 16), (np.array([1, 0, 0]), 560), (np.array([0, -1, 0]), 16), (np.array([0, 1, 0]), 272), (np.array([0, 0, -1]), 16), (np.array([0, 0, 1]), 144)]
    for on, od in outer:
        unit = n / np.linalg.norm(n, axis=1, keepdims=True)
        hit = (np.abs(unit @ on - 1.0) < 1e-6) & (np.abs(d / np.linalg.norm(n, axis=1) - od) < 1e-3)
        assert not hit.any()


def test_leafs_are_convex_and_portals_symmetric(compiled, tmp_path):
    sc, _, _, _, lit = compiled
    # read the lumps back through the .sqwm file itself (bsp_map_save_load.rs layout)
    raw = open(lit, "rb").read()
    assert raw[:4] == b"SqwM" and int.from_bytes(raw[4:8], "little") == 11
    lumps = np.frombuffer(raw, np.uint32, 96, 8).reshape(32, 3)

    def lump(i, dtype):
        o, es, n = (int(x) for x in lumps[i])
        assert es == np.dtype(dtype).itemsize
        return np.frombuffer(raw, dtype, n, o)

    leafs = lump(1, np.dtype([("first_polygon", "<u4"), ("num_polygons", "<u4"), ("first_leaf_portal", "<u4"), ("num_leaf_portals", "<u4")]))
    polys = lump(2, N.map_polygon_dtype)
    portals = lump(3, np.dtype([("leafs", "<u4", 2), ("plane", "<f4", 4), ("first_vertex", "<u4"), ("num_vertices", "<u4")]))
    leafs_portals = lump(4, np.uint32)
    verts = lump(5, np.dtype(("<f4", 3)))
    for li, leaf in enumerate(leafs):
        lp = polys[leaf["first_polygon"] : leaf["first_polygon"] + leaf["num_polygons"]]
        for p in lp:  # a leaf is convex: everything in it lies in front of (or on) each of its polygons' planes
            nrm = np.array(p["plane"][:3], np.float64)
            for q in lp:
                v = verts[q["first_vertex"] : q["first_vertex"] + q["num_vertices"]].astype(np.float64)
                assert np.all(v @ nrm - p["plane"][3] >= -np.linalg.norm(nrm) / 16.0 - 1e-3)
        mine = leafs_portals[leaf["first_leaf_portal"] : leaf["first_leaf_portal"] + leaf["num_leaf_portals"]]
        for pi in mine:
            assert li in portals[pi]["leafs"]
    for pi, p in enumerate(portals):
        assert p["num_vertices"] >= 3 and p["leafs"][0] != p["leafs"][1]
        for l in p["leafs"]:
            leaf = leafs[l]
            assert pi in leafs_portals[leaf["first_leaf_portal"] : leaf["first_leaf_portal"] + leaf["num_leaf_portals"]]
        v = verts[p["first_vertex"] : p["first_vertex"] + p["num_vertices"]].astype(np.float64)
        nrm = np.array(p["plane"][:3], np.float64)
        assert np.all(np.abs(v @ nrm - p["plane"][3]) <= np.linalg.norm(nrm) / 8.0)
    # the doorway is found: some portal lies on x = 256 or x = 288 inside the 64 x 96 opening
    door = [p for p in portals if abs(abs(p["plane"][0]) / np.linalg.norm(p["plane"][:3]) - 1) < 1e-6 and min(abs(abs(p["plane"][3] / np.linalg.norm(p["plane"][:3])) - x) for x in (256, 288)) < 1e-3]
    assert door
    for p in door:
        v = verts[p["first_vertex"] : p["first_vertex"] + p["num_vertices"]]
        assert v[:, 1].min() >= 96 - 0.1 and v[:, 1].max() <= 160 + 0.1 and v[:, 2].max() <= 96 + 0.1


def test_direct_light_closed_forms(compiled):
    sc, _, lstats, _, _ = compiled
    assert lstats["point_lights"] == 1 and lstats["sun_lights"] == 0
    polys, verts, _ = sc.map_geometry()
    lm = sc.lightmaps
    dlm = sc.directional_lightmaps
    assert len(lm) == lstats["lightmap_texels"] == len(dlm)
    ambient = 0.125
    color = 200.0 * 32.0  # map_lights.rs: intensity * MAP_LIGHTS_SCALE
    # floor of room A right below the light (128, 128, 96): cos = 1, distance 96 (minus the 1/16 texel shift)
    best = 0.0
    best_dir = None
    for p in polys:
        nrm = np.array(p["plane"][:3])
        if not (nrm[2] > 0 and abs(p["plane"][3] / np.linalg.norm(nrm)) < 1e-3):
            continue  # not a floor polygon (z = 0, facing up)
        w = ((int(p["tex_coord_max"][0]) + 15) >> 4) - (int(p["tex_coord_min"][0]) >> 4) + 1
        h = ((int(p["tex_coord_max"][1]) + 15) >> 4) - (int(p["tex_coord_min"][1]) >> 4) + 1
        o = int(p["lightmap_data_offset"])
        assert o != 0
        block = lm[o : o + w * h]
        k = int(np.argmax(block[:, 0]))
        if block[k, 0] > best:
            best = float(block[k, 0])
            best_dir = dlm[o + k]
    expected = ambient + color / (96.0 - 1.0 / 16.0) ** 2
    assert abs(best - expected) / expected < 0.02  # (8-bit quantisation of the file format + the texel grid not being centred under the light)
    # directional element of that texel: the light comes from straight above in texel space, nearly no deviation, white
    d = np.array(best_dir["light_direction_vector_scaled"])
    assert d[2] > 0 and abs(d[0]) < 0.15 * d[2] and abs(d[1]) < 0.15 * d[2]
    assert best_dir["directional_light_deviation"] < 0.05
    np.testing.assert_allclose(best_dir["directional_light_color"], [1, 1, 1], atol=0.02)
    np.testing.assert_allclose(best_dir["ambient_light"], [ambient] * 3, atol=0.01)
    # the divider's faces towards room B look away from the light: exactly the ambient term; so does everything the wall shadows
    for p in polys:
        nrm = np.array(p["plane"][:3])
        if nrm[0] > 0 and abs(p["plane"][3] / np.linalg.norm(nrm) - 288) < 1e-3:
            w = ((int(p["tex_coord_max"][0]) + 15) >> 4) - (int(p["tex_coord_min"][0]) >> 4) + 1
            h = ((int(p["tex_coord_max"][1]) + 15) >> 4) - (int(p["tex_coord_min"][1]) >> 4) + 1
            o = int(p["lightmap_data_offset"])
            np.testing.assert_allclose(lm[o : o + w * h], ambient, atol=1.0 / 255.0)  # (the file stores 8-bit components)
    # room B's far wall (x = 544) is lit only through the doorway: some texels lit, most in shadow
    lit_count = total = 0
    for p in polys:
        nrm = np.array(p["plane"][:3])
        if nrm[0] < 0 and abs(-p["plane"][3] / np.linalg.norm(nrm) - 544) < 1e-3:
            w = ((int(p["tex_coord_max"][0]) + 15) >> 4) - (int(p["tex_coord_min"][0]) >> 4) + 1
            h = ((int(p["tex_coord_max"][1]) + 15) >> 4) - (int(p["tex_coord_min"][1]) >> 4) + 1
            o = int(p["lightmap_data_offset"])
            block = lm[o : o + w * h, 0]
            lit_count += int((block > ambient + 0.01).sum())
            total += block.size
    assert 0 < lit_count < total * 0.5


def test_light_grid_of_the_direct_light(compiled):
    """prepare_light_grid / calculate_light_grid / compress_light_grid (lightmaps_builder.rs:1614-1834) with the direct light: cell
    alignment, column compression, and more light next to the lamp than behind the dividing wall."""
    sc, _, lstats, _, lit = compiled
    raw = open(lit, "rb").read()
    lumps = np.frombuffer(raw, np.uint32, 96, 8).reshape(32, 3)
    o, es, n = (int(x) for x in lumps[14])
    assert (es, n) == (36, 1)
    cell = np.frombuffer(raw, "<f4", 3, o)
    start = np.frombuffer(raw, "<f4", 3, o + 12)
    size = np.frombuffer(raw, "<u4", 3, o + 24)
    np.testing.assert_array_equal(cell, [64, 64, 64])
    np.testing.assert_array_equal(start % 64, 0)  # aligned to the cell size
    assert np.all(start <= [0, 0, 0]) and np.all(start + (size - 1) * cell >= [544, 256, 128])  # the bounding box of the map's vertices (the rooms' inner faces)
    o, es, n = (int(x) for x in lumps[15])
    assert es == 12 and n == size[0] * size[1]
    cols = np.frombuffer(raw, np.dtype([("start_z", "<u4"), ("first_sample", "<u4"), ("num_samples", "<u4")]), n, o)
    o, es, n = (int(x) for x in lumps[16])
    assert es == 32 and n == lstats["light_grid_samples"] == int(cols["num_samples"].sum()) and n == sc._lib.sqwh_scene_num_light_grid_samples(sc._handle)
    samples = np.frombuffer(raw, np.uint8, n * 32, o).reshape(n, 32)

    def cube_light(x, y, z):  # sum of the six de-quantised cube sides' first component + the directional part's length
        gx, gy, gz = (int(round((v - s) / c)) for v, s, c in zip((x, y, z), start, cell))
        col = cols[gx + gy * size[0]]
        assert col["start_z"] <= gz < col["start_z"] + col["num_samples"]
        e = samples[col["first_sample"] + gz - col["start_z"]]
        cube = sum(float(e[4 * s]) * float(e[4 * s + 3]) / 255.0 for s in range(6))
        vec = e[24:27].view(np.int8).astype(np.float32) * float(e[27]) / 127.0
        return cube + float(np.linalg.norm(vec))

    near, far = cube_light(128, 128, 64), cube_light(448, 64, 64)
    assert near > 4.0 * far and far > 0.0  # the far point sees only the ambient term (the wall shadows the lamp)


def test_oracle_draws_the_compiled_map_without_gaps(compiled, oracle_lib):
    sc, _, _, _, _ = compiled
    poses = sc.camera_path(6)
    assert len(poses) == 6
    hr = S.HostRenderer(sc, clear_background=False)
    try:
        for pose in poses:
            assert 0 < pose[0] < 544 and 0 < pose[1] < 256 and 0 < pose[2] < 128  # the path stays inside the rooms (through the doorway)
            hf = hr.build_frame(320, 200, pose, color_bits=32)
            assert len(hf.cmds) > 0
            oracle_lib.reset_assert_failures()
            surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
            fb = np.zeros((200, 320), np.uint32)
            oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
            assert oracle_lib.assert_failures() == 0
            assert int((fb == 0).sum()) == 0  # closed rooms: every pixel is covered by some polygon
    finally:
        hr.close()


def test_compile_is_deterministic_and_csg_only_removes(tmp_path):
    table = mc.materials_table(MATERIALS)
    a, b, c = (str(tmp_path / n) for n in ("a.sqwm", "b.sqwm", "c.sqwm"))
    sa = mc.compile_map(two_rooms_map(), table, a, perform_csg=True, advanced=True)
    sb = mc.compile_map(two_rooms_map(), table, b, perform_csg=True, advanced=True)
    assert open(a, "rb").read() == open(b, "rb").read() and sa == sb
    sc_ = mc.compile_map(two_rooms_map(), table, c, perform_csg=False, advanced=False)
    assert sc_["polygons_after_csg"] == sc_["initial_polygons"] == sa["initial_polygons"]
    assert sc_["leafs"] >= 3


@pytest.mark.skipif(not os.path.exists(S.DEMO2_MAP_PATH), reason="assets/_ref/demo2.sqwm is built by __graft_entry__.build() where /root/reference exists")
def test_demo2_map_loads_and_draws(oracle_lib):
    sc = S.demo2_scene(texture_size_log2=6)
    try:
        assert sc.info.num_polygons > 10000 and sc.info.num_leafs > 2000 and sc.info.num_portals > 5000
        names = {sc.texture_name(i) for i in range(sc.info.num_materials)}
        assert {"bricks", "limestone_bricks", "sky"} <= names and "__TB_empty" not in names  # bsp = false faces are gone
        poses = sc.camera_path(64)
        hr = S.HostRenderer(sc)
        try:
            for k in (0, 20, 41):
                hf = hr.build_frame(480, 270, poses[k], color_bits=32)
                assert len(hf.cmds) > 50
                oracle_lib.reset_assert_failures()
                surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
                fb = np.zeros((270, 480), np.uint32)
                oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
                assert oracle_lib.assert_failures() == 0
        finally:
            hr.close()
    finally:
        sc.close()


def test_load_compiled_map_errors_and_game_submodels(compiled, tmp_path):
    """Scene.load_compiled_map reports what bsp_map_save_load.rs::load_map reports (missing file, bad id, bad version); the inline
    models the reference's game does not draw as such are hidden by HostRenderer.game_submodel_entities (world_spawn.rs:33-55, :331-434)."""
    _, _, _, table, lit = compiled
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        with pytest.raises(OSError):
            sc.load_compiled_map(str(tmp_path / "missing.sqwm"), table)
        bad = tmp_path / "bad.sqwm"
        raw = bytearray(open(lit, "rb").read())
        bad.write_bytes(b"XXXX" + bytes(raw[4:]))
        with pytest.raises(OSError, match="bad id"):
            sc.load_compiled_map(str(bad), table)
        raw[4:8] = (10).to_bytes(4, "little")
        bad.write_bytes(bytes(raw))
        with pytest.raises(OSError, match="version"):
            sc.load_compiled_map(str(bad), table)
    finally:
        sc.close()
    # a map with a trigger volume, a mirror and a door: only the door is shown
    text = two_rooms_map(door_entity=True)
    text += '{\n"classname" "trigger_teleport"\n' + box_brush((300, 40, 0), (340, 80, 64), "wall") + "\n}\n"
    text += '{\n"classname" "func_mirror"\n' + box_brush((500, 100, 20), (504, 160, 100), "wall") + "\n}\n"
    path = str(tmp_path / "ents.sqwm")
    stats = mc.compile_map(text, table, path, perform_csg=True, advanced=True)
    assert stats["submodels"] == 3
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        sc.load_compiled_map(path, table, 5, 1)
        hr = S.HostRenderer(sc)
        try:
            names = [(hr._lib.sqwh_scene_submodel_classname(sc._handle, i) or b"").decode() for i in range(3)]
            assert names == ["func_door", "trigger_teleport", "func_mirror"]
            assert list(hr.default_submodel_entities()["present"]) == [1, 1, 1]
            assert list(hr.game_submodel_entities()["present"]) == [1, 0, 0]
            pose = sc.camera_path(4)[0]
            shown = hr.build_frame(320, 200, pose, submodels=hr.default_submodel_entities())
            hidden = hr.build_frame(320, 200, pose, submodels=hr.game_submodel_entities())
            assert len(hidden.cmds) <= len(shown.cmds)
        finally:
            hr.close()
    finally:
        sc.close()


def _polygon_area(v):
    v = np.asarray(v, np.float64)
    c = np.zeros(3)
    for i in range(1, len(v) - 1):
        c += np.cross(v[i] - v[0], v[i + 1] - v[0])
    return 0.5 * float(np.linalg.norm(c))


def _order_independent_summary(path):
    """The facts tools/rust_golden/src/bin/map_golden.rs prints for the reference's compiler, read from a .sqwm file."""
    raw = open(path, "rb").read()
    lumps = np.frombuffer(raw, np.uint32, 96, 8).reshape(32, 3)

    def lump(i, dtype):
        o, es, n = (int(x) for x in lumps[i])
        return np.frombuffer(raw, dtype, n, o) if es == np.dtype(dtype).itemsize else np.zeros(0, dtype)

    polys = lump(2, N.map_polygon_dtype)
    portals = lump(3, np.dtype([("leafs", "<u4", 2), ("plane", "<f4", 4), ("first_vertex", "<u4"), ("num_vertices", "<u4")]))
    verts = lump(5, np.dtype(("<f4", 3)))
    areas = sorted(_polygon_area(verts[p["first_vertex"] : p["first_vertex"] + p["num_vertices"]]) for p in polys)
    portal_areas = sorted(_polygon_area(verts[p["first_vertex"] : p["first_vertex"] + p["num_vertices"]]) for p in portals)
    return {"nodes": int(lumps[0][2]), "leafs": int(lumps[1][2]), "polygons": len(polys), "portals": len(portals), "submodels": int(lumps[7][2]),
            "lightmap_texels": int(lumps[12][2]), "polygon_areas_sorted": areas, "portal_areas_sorted": portal_areas}


def test_fixture_file_is_current():
    """tests/golden/two_rooms.map (input of the reference-side emitter) is what two_rooms_map() writes."""
    with open(os.path.join(ROOT, "tests", "golden", "two_rooms.map"), "r", encoding="utf-8") as f:
        assert f.read() == two_rooms_map()


GOLDEN_MAP = os.path.join(ROOT, "tests", "golden", "reference_map_compiler.json")


@pytest.mark.skipif(not os.path.exists(GOLDEN_MAP), reason="tests/golden/reference_map_compiler.json is emitted by tools/rust_golden (map_golden) where a Rust toolchain exists")
def test_compiler_against_the_reference_summary(tmp_path):
    """Pins the C++ map compiler against the reference's own (order-independent facts: counts and the sorted areas of polygons and portals)."""
    import json

    golden = json.load(open(GOLDEN_MAP))
    table = mc.materials_table({})
    for name, csg, adv in (("csg_advanced", True, True), ("plain", False, False)):
        path = str(tmp_path / f"{name}.sqwm")
        st = mc.compile_map(two_rooms_map(), table, path, perform_csg=csg, advanced=adv)
        want, got = golden[name], _order_independent_summary(path)
        assert st["initial_polygons"] == want["initial_polygons"] and st["polygons_after_csg"] == want["polygons_after_csg"]
        for key in ("nodes", "leafs", "polygons", "portals", "submodels", "lightmap_texels"):
            assert got[key] == want[key], key
        np.testing.assert_allclose(got["polygon_areas_sorted"], want["polygon_areas_sorted"], rtol=1e-5, atol=1e-2)
        np.testing.assert_allclose(got["portal_areas_sorted"], want["portal_areas_sorted"], rtol=1e-5, atol=1e-2)


def test_order_independent_summary_is_self_consistent(compiled):
    """(the reader used for the reference comparison agrees with the compiler's own statistics)"""
    _, stats, _, _, lit = compiled
    got = _order_independent_summary(lit)
    assert (got["nodes"], got["leafs"], got["polygons"], got["portals"], got["submodels"]) == (stats["nodes"], stats["leafs"], stats["polygons"], stats["portals"], stats["submodels"])
    # the rooms' inner surface: two 256 x 256 x 128 rooms + the doorway's reveals, minus the door opening on both sides
    inner = 2 * (2 * 256 * 256 + 4 * 256 * 128) - 2 * 64 * 96 + 32 * (2 * 96 + 64)
    world = sum(got["polygon_areas_sorted"]) - (2 * 32 * 32 + 4 * 32 * 64)  # the door submodel's six faces
    assert abs(world - inner) / inner < 0.02


def _lightmap_blocks(sc, select):
    """(polygon, lightmap block [h, w, 3]) of the polygons `select` accepts."""
    polys, _verts, _ = sc.map_geometry()
    for p in polys:
        if not select(p):
            continue
        w = ((int(p["tex_coord_max"][0]) + 15) >> 4) - (int(p["tex_coord_min"][0]) >> 4) + 1
        h = ((int(p["tex_coord_max"][1]) + 15) >> 4) - (int(p["tex_coord_min"][1]) >> 4) + 1
        o = int(p["lightmap_data_offset"])
        if o:
            yield p, sc.lightmaps[o : o + w * h].reshape(h, w, 3)


def test_sun_through_a_sky_ceiling_and_a_spot_light(tmp_path):
    """map_lights.rs: the sun of worldspawn (`_sun_mangle`, `_sunlight`, `_sunlight_color`, :160-246) reaches a texel only through a
    sky polygon (light_trace.rs get_sun_shadow_factor :162-176) and a `light` with a `target` is a spot light with half-angle
    `angle` / 2 (:120-140).  A courtyard with a sky ceiling and a roofed half: sunlit floor = colour x cos(angle of incidence), the
    floor under the roof gets no sun; the spot light lights its cone on the roofed floor and nothing outside it."""
    b = []
    b.append(box_brush((-16, -16, -16), (528, 272, 0), "floor"))
    b.append(box_brush((-16, -16, 256), (272, 272, 272), "sky"))        # open half: sky ceiling high up
    b.append(box_brush((256, -16, 128), (528, 272, 272), "ceiling"))    # roofed half: solid block down to z = 128
    b.append(box_brush((-16, -16, 0), (0, 272, 256)))
    b.append(box_brush((512, -16, 0), (528, 272, 128)))
    b.append(box_brush((0, -16, 0), (512, 0, 256)))
    b.append(box_brush((0, 256, 0), (512, 272, 256)))
    world = '{\n"classname" "worldspawn"\n"_sun_mangle" "0 -90 0"\n"_sunlight" "2"\n"_sunlight_color" "255 128 0"\n' + "\n".join(b) + "\n}"
    ents = [world, '{\n"classname" "info_player_start"\n"origin" "64 128 32"\n}']
    ents.append('{\n"classname" "light"\n"origin" "448 128 120"\n"light" "100"\n"target" "t1"\n"angle" "60"\n}')
    ents.append('{\n"classname" "info_null"\n"targetname" "t1"\n"origin" "448 128 0"\n}')
    text = "\n".join(ents) + "\n"
    materials = dict(MATERIALS, sky={"light": False, "special_effect": {"Skybox": {}}})
    table = mc.materials_table(materials)
    path = str(tmp_path / "yard.sqwm")
    mc.compile_map(text, table, path, perform_csg=True, advanced=False)
    st = mc.light_map(path, table, path, sample_grid_size=1)
    assert st["sun_lights"] == 1 and st["point_lights"] == 1
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        sc.load_compiled_map(path, table, 5, 1)
        _polys, verts, _ = sc.map_geometry()
        sun = np.array([2.0 * 32.0 * 1.0, 2.0 * 32.0 * 128.0 / 255.0, 0.0])  # straight down (pitch -90): cos = 1 on the floor; colour 255 128 0
        sunlit = shaded = 0
        spot_hit = spot_miss = False

        def is_floor(p):
            return p["plane"][2] > 0 and abs(p["plane"][3]) < 1e-3

        for p, block in _lightmap_blocks(sc, is_floor):
            # world x / y of the texel columns / rows: the floor's texture axes are the QBSP ones (u = x, v = -y, scale 1, no offset)
            assert tuple(np.round(p["tex_coord_equation"][0][:3], 6)) == (1.0, 0.0, 0.0) and tuple(np.round(p["tex_coord_equation"][1][:3], 6)) == (0.0, -1.0, 0.0)
            xs = ((int(p["tex_coord_min"][0]) >> 4) + np.arange(block.shape[1])) * 16.0
            ys = -((int(p["tex_coord_min"][1]) >> 4) + np.arange(block.shape[0])) * 16.0
            inside = (ys > 16) & (ys < 240)
            for col, x in enumerate(xs):
                texels = block[inside, col]
                if len(texels) == 0:
                    continue
                if 16 < x < 224:  # the open half, away from the roof's edge
                    np.testing.assert_allclose(texels, np.broadcast_to(sun, texels.shape), rtol=0.02, atol=0.3)
                    sunlit += len(texels)
                elif 288 < x < 496:  # under the roof: no sun (red / green stay what the white spot light gives: equal to blue)
                    np.testing.assert_allclose(texels[:, 0], texels[:, 2], atol=0.02)
                    shaded += len(texels)
                    d = np.hypot(x - 448.0, ys[inside] - 128.0)
                    if (d < 40).any():
                        spot_hit = spot_hit or float(texels[d < 40].min()) > 0.05
                    if (d > 100).any():  # outside the 30 degree half-angle cone (radius 120 tan 30 = 69 units on the floor)
                        spot_miss = True
                        assert float(texels[d > 100].max()) < 1e-3
        assert sunlit > 0 and shaded > 0 and spot_hit and spot_miss
        # right below the spot light: 100 * 32 / 120^2 (the blue channel sees only the white spot light)
        under = 100.0 * 32.0 / (120.0 - 1.0 / 16.0) ** 2
        spot_best = max(float(block[..., 2].max()) for _p, block in _lightmap_blocks(sc, is_floor))
        assert abs(spot_best - under) / under < 0.05
    finally:
        sc.close()


def test_pvs_of_a_zig_zag_of_rooms(tmp_path):
    """pvs.rs calculate_visibility_matrix: three rooms in a row whose doorways are at opposite ends of the dividing walls.  The
    matrix is symmetric with a true diagonal; neighbouring rooms see each other through their doorway; the far corner of room C is
    hidden from the far corner of room A (no straight line passes both doorways).  Lighting with --use-pvs gives the same lightmaps
    as the contribution cut-off here (the lamp reaches every leaf it can see; what the PVS prunes is shadowed anyway)."""
    # rooms A x in [0, 256], B [384, 640], C [768, 1024], all y in [0, 512]; the walls between them are 128 thick with 48-wide
    # doorways (tunnels) at opposite ends: whatever passes tunnel A|B has |dy/dx| <= 48 / 128 seen from that tunnel's mouth, while
    # tunnel B|C lies at dy/dx >= 384 / 384 from there (the PVS is conservative: it unions the views from a portal's vertices)
    b = [box_brush((-16, -16, -16), (1040, 528, 0), "floor"), box_brush((-16, -16, 128), (1040, 528, 144), "ceiling"),
         box_brush((-16, -16, 0), (0, 528, 128)), box_brush((1024, -16, 0), (1040, 528, 128)), box_brush((0, -16, 0), (1024, 0, 128)), box_brush((0, 512, 0), (1024, 528, 128)),
         box_brush((256, 0, 0), (384, 16, 128), "divider"), box_brush((256, 64, 0), (384, 512, 128), "divider"), box_brush((256, 16, 96), (384, 64, 128), "divider"),
         box_brush((640, 0, 0), (768, 448, 128), "divider"), box_brush((640, 496, 0), (768, 512, 128), "divider"), box_brush((640, 448, 96), (768, 496, 128), "divider")]
    ents = ['{\n"classname" "worldspawn"\n' + "\n".join(b) + "\n}", '{\n"classname" "info_player_start"\n"origin" "64 128 32"\n}',
            '{\n"classname" "light"\n"origin" "128 128 96"\n"light" "200"\n}']
    table = mc.materials_table(MATERIALS)
    path = str(tmp_path / "zigzag.sqwm")
    mc.compile_map("\n".join(ents) + "\n", table, path, perform_csg=True, advanced=True)
    vis = mc.leafs_visibility(path)
    n = len(vis)
    assert n >= 5 and (vis == vis.T).all() and vis.diagonal().all()
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        sc.load_compiled_map(path, table, 5, 1)
        leaf_at = lambda p: _leaf_of_point(path, p)
        a_far, b_mid, c_far = leaf_at((16, 496, 64)), leaf_at((512, 256, 64)), leaf_at((1008, 16, 64))
        a_door, c_door = leaf_at((320, 40, 48)), leaf_at((704, 472, 48))  # inside the two tunnels
        assert vis[a_far, a_door] and vis[a_door, b_mid] and vis[b_mid, c_door] and vis[c_door, c_far]  # through one doorway each
        assert not vis[a_far, c_far]                               # two doorways that do not line up
        assert vis.sum() < n * n                                   # (not everything sees everything)
    finally:
        sc.close()
    lit_a, lit_b = str(tmp_path / "a.sqwm"), str(tmp_path / "b.sqwm")
    mc.light_map(path, table, lit_a, sample_grid_size=1, ambient_light=0.125)
    mc.light_map(path, table, lit_b, sample_grid_size=1, ambient_light=0.125, use_pvs=True)
    assert open(lit_a, "rb").read() == open(lit_b, "rb").read()


def _leaf_of_point(path, point):
    """bsp_map_compact.rs get_leaf_for_point on the file's node lump."""
    raw = open(path, "rb").read()
    lumps = np.frombuffer(raw, np.uint32, 96, 8).reshape(32, 3)
    o, es, cnt = (int(x) for x in lumps[0])
    nodes = np.frombuffer(raw, np.dtype([("children", "<u4", 2), ("plane", "<f4", 4)]), cnt, o)
    index = cnt - 1 if cnt else 1 << 31
    while index < (1 << 31):
        nd = nodes[index]
        index = int(nd["children"][0] if float(np.dot(nd["plane"][:3], point)) >= float(nd["plane"][3]) else nd["children"][1])
    return index - (1 << 31)


def test_csg_cuts_holes_for_a_pillar(tmp_path, oracle_lib):
    """map_csg.rs make_hole_in_polygon (:162-328): a pillar from floor to ceiling in the middle of a room leaves its footprint as a
    hole in both (pieces cut along the texture axes), and its own bottom / top faces vanish.  Total surface = the closed form; the
    oracle still draws gap-free frames (no crack around the hole)."""
    b = [box_brush((-16, -16, -16), (272, 272, 0), "floor"), box_brush((-16, -16, 128), (272, 272, 144), "ceiling"),
         box_brush((-16, -16, 0), (0, 272, 128)), box_brush((256, -16, 0), (272, 272, 128)), box_brush((0, -16, 0), (256, 0, 128)), box_brush((0, 256, 0), (256, 272, 128)),
         box_brush((96, 112, 0), (144, 160, 128), "divider")]  # the pillar: 48 x 48 footprint
    ents = ['{\n"classname" "worldspawn"\n' + "\n".join(b) + "\n}", '{\n"classname" "info_player_start"\n"origin" "32 32 32"\n}', '{\n"classname" "light"\n"origin" "200 64 100"\n"light" "150"\n}']
    table = mc.materials_table(MATERIALS)
    path = str(tmp_path / "pillar.sqwm")
    st = mc.compile_map("\n".join(ents) + "\n", table, path, perform_csg=True, advanced=True)
    assert st["polygons_after_csg"] > 6 + 4  # the floor and the ceiling are in several pieces now
    got = _order_independent_summary(path)
    want = 2 * (256 * 256 - 48 * 48) + 4 * 256 * 128 + 4 * 48 * 128
    assert abs(sum(got["polygon_areas_sorted"]) - want) / want < 1e-3
    mc.light_map(path, table, path, sample_grid_size=1, ambient_light=0.2)
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        sc.load_compiled_map(path, table, 5, 1)
        hr = S.HostRenderer(sc, clear_background=False)
        try:
            for pose in ([32.0, 32.0, 64.0, -0.8, -0.3], [220.0, 220.0, 40.0, 2.4, 0.4], [128.0, 40.0, 100.0, 0.0, -0.9]):
                hf = hr.build_frame(320, 200, np.array(pose, np.float32), color_bits=32)
                oracle_lib.reset_assert_failures()
                surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
                fb = np.zeros((200, 320), np.uint32)
                oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
                assert oracle_lib.assert_failures() == 0 and int((fb == 0).sum()) == 0
        finally:
            hr.close()
    finally:
        sc.close()


def plane_brush(planes, texture="wall"):
    """A brush from (outward normal, point on the plane) pairs: three points per plane wound like box_brush's."""
    lines = ["{"]
    for n, p in planes:
        n, p = np.asarray(n, float), np.asarray(p, float)
        t1 = np.cross(n, [0.0, 0.0, 1.0] if abs(n[2]) < 0.9 else [1.0, 0.0, 0.0])
        t1 /= np.linalg.norm(t1)
        t2 = np.cross(n, t1)
        p0, p1, p2 = p + t1 * 32, p, p + t2 * 32
        if np.dot(np.cross(p0 - p1, p2 - p1), n) < 0:
            p0, p2 = p2, p0
        lines.append(" ".join("( %.6f %.6f %.6f )" % tuple(q) for q in (p0, p1, p2)) + f" {texture} 0 0 0 1 1")
    lines.append("}")
    return "\n".join(lines)


def test_sloped_brush_in_a_room(tmp_path, oracle_lib):
    """Planes that are not axis-aligned: a ramp (wedge) against a wall.  Its slope and its two triangular sides are what remains of
    it after CSG (bottom and back lie in the floor and the wall), the floor and the wall lose the covered rectangles; QBSP texture
    axes pick the floor projection for the slope (map_polygonizer.rs:447-497).  Closed-form area, gap-free frames."""
    ramp = plane_brush([((0, 0, -1), (0, 0, 0)), ((1, 0, 0), (256, 0, 0)), ((0, -1, 0), (0, 64, 0)), ((0, 1, 0), (0, 192, 0)),
                        ((-64.0, 0.0, 128.0), (128, 0, 0))], "divider")  # rises from (128, *, 0) to (256, *, 64)
    b = [box_brush((-16, -16, -16), (272, 272, 0), "floor"), box_brush((-16, -16, 128), (272, 272, 144), "ceiling"),
         box_brush((-16, -16, 0), (0, 272, 128)), box_brush((256, -16, 0), (272, 272, 128)), box_brush((0, -16, 0), (256, 0, 128)), box_brush((0, 256, 0), (256, 272, 128)), ramp]
    ents = ['{\n"classname" "worldspawn"\n' + "\n".join(b) + "\n}", '{\n"classname" "info_player_start"\n"origin" "32 32 32"\n}', '{\n"classname" "light"\n"origin" "64 128 100"\n"light" "150"\n}']
    table = mc.materials_table(MATERIALS)
    path = str(tmp_path / "ramp.sqwm")
    mc.compile_map("\n".join(ents) + "\n", table, path, perform_csg=True, advanced=True)
    got = _order_independent_summary(path)
    slope = 128.0 * np.hypot(128.0, 64.0)
    sides = 2 * 0.5 * 128.0 * 64.0
    want = 2 * 256 * 256 + 4 * 256 * 128 - 128 * 128 - 128 * 64 + slope + sides  # floor and wall lose the ramp's bottom and back
    assert abs(sum(got["polygon_areas_sorted"]) - want) / want < 2e-3
    mc.light_map(path, table, path, sample_grid_size=1, ambient_light=0.2)
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        sc.load_compiled_map(path, table, 5, 1)
        polys, _verts, _ = sc.map_geometry()
        sloped = [p for p in polys if abs(p["plane"][0]) > 1e-3 and abs(p["plane"][2]) > 1e-3]
        assert sloped
        for p in sloped:  # the slope is textured like a floor: u along x, v along -y
            assert abs(p["tex_coord_equation"][0][0] - 1.0) < 1e-6 and abs(p["tex_coord_equation"][1][1] + 1.0) < 1e-6
            o = int(p["lightmap_data_offset"])
            assert o != 0 and float(sc.lightmaps[o : o + 4].max()) > 0.2  # lit by the lamp it faces
        hr = S.HostRenderer(sc, clear_background=False)
        try:
            for pose in ([32.0, 32.0, 64.0, -0.8, -0.3], [40.0, 128.0, 40.0, -1.57, 0.0], [200.0, 30.0, 110.0, 0.3, -0.8]):
                hf = hr.build_frame(320, 200, np.array(pose, np.float32), color_bits=32)
                oracle_lib.reset_assert_failures()
                surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, sc.images, sc.lightmaps, sc.directional_lightmaps, hf.total_surface_pixels)
                fb = np.zeros((200, 320), np.uint32)
                oracle_lib.draw(32, fb, hf.cmds, hf.polygon_points, None, surf, sc.images)
                assert oracle_lib.assert_failures() == 0 and int((fb == 0).sum()) == 0
        finally:
            hr.close()
    finally:
        sc.close()


def test_material_flags_steer_the_compiler(tmp_path):
    """material.rs flags read by the tools: `bsp: false` faces are dropped (bsp_builder.rs:555-582: the demo marks hidden faces with
    __TB_empty), `twosided` polygons are doubled with the opposite winding, `blocks_view: false` brushes neither cut other brushes
    (map_csg.rs:50-66) nor close portals (bsp_builder.rs:717-737): a fence across the doorway leaves room B reachable and visible."""
    materials = dict(MATERIALS, fence={"twosided": True, "blocks_view": False, "blending_mode": "AlphaTest", "shadow": False}, hidden={"bsp": False})
    table = mc.materials_table(materials)
    base = two_rooms_map(door_entity=False)
    fence = box_brush((268, 96, 0), (276, 160, 96), "fence", textures={(1, -1): "hidden", (1, 1): "hidden", (2, -1): "hidden", (2, 1): "hidden"})
    text = base.replace('// a comment\n', '// a comment\n' + fence + "\n", 1)
    assert text != base
    a, b = str(tmp_path / "plain.sqwm"), str(tmp_path / "fence.sqwm")
    s0 = mc.compile_map(base, table, a, perform_csg=True, advanced=True)
    s1 = mc.compile_map(text, table, b, perform_csg=True, advanced=True)
    g0, g1 = _order_independent_summary(a), _order_independent_summary(b)
    # the fence adds its two 64 x 96 faces, each twice (both windings); its four `hidden` faces are gone; nothing else changes in area
    added = sum(g1["polygon_areas_sorted"]) - sum(g0["polygon_areas_sorted"])
    assert abs(added - 4 * 64 * 96) < 1.0
    assert s1["leafs"] >= s0["leafs"] and g1["portals"] >= g0["portals"]
    sc = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
    try:
        sc.load_compiled_map(b, table, 5, 1)
        names = [sc.texture_name(i) for i in range(sc.info.num_materials)]
        assert "fence" in names and "hidden" not in names
        polys, verts, _ = sc.map_geometry()
        fi = names.index("fence")
        normals = sorted(round(float(p["plane"][0] / np.linalg.norm(p["plane"][:3]))) for p in polys if p["texture"] == fi)
        assert normals.count(1) == normals.count(-1) and len(normals) >= 4  # every fence polygon exists with both facings
        # room B stays reachable (the camera path still runs through the doorway) and potentially visible from room A
        poses = sc.camera_path(8)
        assert poses[:, 0].max() > 300
        vis = mc.leafs_visibility(b)
        assert vis[_leaf_of_point(b, (64, 128, 32)), _leaf_of_point(b, (416, 128, 32))]
    finally:
        sc.close()
    # the fence lets light through (shadow: false -> opacity 1, light_trace.rs:15-40): room B's floor behind it is lit as without it
    la, lb = str(tmp_path / "la.sqwm"), str(tmp_path / "lb.sqwm")
    mc.light_map(a, table, la, sample_grid_size=1)
    mc.light_map(b, table, lb, sample_grid_size=1)

    def floor_light_in_b(path):
        sc2 = S.Scene(num_materials=4, texture_size_log2=4, grid=(4, 4, 1), num_rooms=1)
        try:
            sc2.load_compiled_map(path, table, 5, 1)
            _p, v2, _ = sc2.map_geometry()
            total = 0.0
            for p, block in _lightmap_blocks(sc2, lambda q: q["plane"][2] > 0 and abs(q["plane"][3]) < 1e-3):
                if v2[p["first_vertex"] : p["first_vertex"] + p["num_vertices"]][:, 0].min() >= 288:
                    total += float(block[..., 0].sum())
            return total
        finally:
            sc2.close()

    fa, fb_ = floor_light_in_b(la), floor_light_in_b(lb)
    assert fa > 1.0 and abs(fa - fb_) / fa < 0.05
