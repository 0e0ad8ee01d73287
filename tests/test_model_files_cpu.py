"""MD3 / IQM loaders of the host mirror (square_wheel_b200/host/model_file.cpp <- triangle_model_md3.rs, triangle_model_iqm.rs):
files written by an independent Python writer (tests/model_files.py) must load into the arrays an independent numpy decode gives,
the reference's rejection cases must be rejected, loaded models must go through the unchanged model pipeline, and the one real
model file that ships with the reference (iqm/demo/mrfixit.iqm, when the reference tree is present) must reproduce the per-frame
bounding boxes its exporter stored — a known answer that comes from third-party data, not from this repository."""
import os

import numpy as np
import pytest

import model_files as MF
from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S

MRFIXIT = "/root/reference/iqm/demo/mrfixit.iqm"


@pytest.fixture()
def scene():
    sc = S.Scene(seed=5, texture_size_log2=5, num_materials=4, num_models=2, model_detail=4)
    yield sc
    sc.close()


def _md3_fixture(rng, num_frames=3, with_tags=True):
    nv, nt = 12, 16
    meshes = []
    for mi, nf in enumerate((num_frames, 1)):
        meshes.append(dict(
            name=f"mesh{mi}", shader=f"models/skin{mi}.tga", triangles=rng.integers(0, nv, (nt, 3)), tex_coords=rng.uniform(0, 1, (nv, 2)),
            origins=rng.integers(-2000, 2000, (nf, nv, 3)), normals=rng.integers(-32768, 32767, (nf, nv))))
    bbox = np.concatenate([rng.uniform(-40, -10, (num_frames, 3)), rng.uniform(10, 40, (num_frames, 3))], 1)
    tags = None
    if with_tags:
        tags = [[(f"tag_{t}", rng.uniform(-5, 5, 3), rng.uniform(-1, 1, 9)) for t in range(2)] for _ in range(num_frames)]
    return meshes, bbox, tags


def test_md3_loads_like_an_independent_decode(scene):
    rng = np.random.default_rng(11)
    meshes, bbox, tags = _md3_fixture(rng)
    index = scene.load_model(MF.write_md3(meshes, bbox, tags))
    assert index == 2 and scene.info.num_models == 3
    mi = scene.model_info(index)
    assert (mi.num_frames, mi.num_meshes, mi.num_bones) == (3, 2, 2)
    # mesh 0: three frames -> VertexAnimated; positions are i16 / 64, normals the latitude / longitude code
    a = scene.model_arrays(index, 0)
    np.testing.assert_array_equal(a["triangles"], np.asarray(meshes[0]["triangles"], np.uint16))
    np.testing.assert_array_equal(a["constant"], np.asarray(meshes[0]["tex_coords"], np.float32))
    want_pos = (np.asarray(meshes[0]["origins"], np.float32) / 64.0).reshape(-1, 3)
    np.testing.assert_array_equal(a["variable"][:, :3], want_pos)
    np.testing.assert_allclose(a["variable"][:, 3:], MF.md3_normal(np.asarray(meshes[0]["normals"]).reshape(-1)), atol=1e-6)
    assert len(a["non_animated"]) == 0 and len(a["skeleton"]) == 0
    assert a["names"][:2] == ["mesh0", "models/skin0.tga"] and a["names"][2:4] == ["tag_0", "tag_1"]
    np.testing.assert_array_equal(a["tc_shift"], [-0.5, -0.5])
    np.testing.assert_array_equal(a["frames_bbox"], bbox.astype(np.float32))
    assert (a["bone_parents"] == 0xFFFFFFFF).all()
    # tags -> one bone matrix per (frame, tag): translation * rotation, the nine floats being the columns of the rotation
    fb = a["frame_bones"].reshape(3, 2, 4, 4)  # [frame][tag][column][row]
    for f in range(3):
        for t in range(2):
            _, origin, rot = tags[f][t]
            np.testing.assert_allclose(fb[f, t, :3, :3], np.asarray(rot, np.float32).reshape(3, 3), rtol=0, atol=0)
            np.testing.assert_allclose(fb[f, t, 3, :3], np.asarray(origin, np.float32), rtol=0, atol=0)
            np.testing.assert_array_equal(fb[f, t, :, 3], [0, 0, 0, 1])
    # mesh 1: a single frame with as many vertices as texture coordinates -> NonAnimated
    b = scene.model_arrays(index, 1)
    assert len(b["non_animated"]) == 12 and len(b["variable"]) == 0
    np.testing.assert_array_equal(b["non_animated"][:, :3], (np.asarray(meshes[1]["origins"], np.float32) / 64.0).reshape(-1, 3))
    np.testing.assert_array_equal(b["non_animated"][:, 6:], np.asarray(meshes[1]["tex_coords"], np.float32))


def test_md3_rejections_follow_the_reference(scene):
    rng = np.random.default_rng(12)
    meshes, bbox, tags = _md3_fixture(rng)
    good = MF.write_md3(meshes, bbox, tags)
    with pytest.raises(ValueError, match="not a valid MD3|unknown model format"):
        scene.load_model(MF.write_md3(meshes, bbox, tags, ident=b"IDP2"))
    with pytest.raises(ValueError, match="version"):
        scene.load_model(MF.write_md3(meshes, bbox, tags, version=14))
    with pytest.raises(ValueError, match="header"):
        scene.load_model(good[:50])
    with pytest.raises(ValueError, match="truncated"):
        scene.load_model(good[:-40])
    with pytest.raises(ValueError, match="unknown model format"):
        scene.load_model(b"not a model at all")
    # a mesh with a wrong ident is dropped, the others are kept (triangle_model_md3.rs:124-128)
    meshes[0]["ident"] = b"XXXX"
    idx = scene.load_model(MF.write_md3(meshes, bbox, tags))
    assert scene.model_info(idx).num_meshes == 1 and scene.model_arrays(idx, 0)["names"][0] == "mesh1"
    assert scene.info.num_models == 3  # the failed loads appended nothing


_iqm_fixture = MF.tube_iqm_file


def test_iqm_loads_like_the_published_formulas(scene):
    rng = np.random.default_rng(21)
    model, data, mats, bounds = _iqm_fixture(rng)
    index = scene.load_model(data)
    mi = scene.model_info(index)
    assert (mi.num_frames, mi.num_meshes, mi.num_bones, mi.kind) == (5, 1, 3, 2)
    a = scene.model_arrays(index)
    np.testing.assert_array_equal(a["triangles"], np.asarray(model["triangles"], np.uint16))
    np.testing.assert_array_equal(a["skeleton"]["position"], np.asarray(model["positions"], np.float32))
    np.testing.assert_array_equal(a["skeleton"]["normal"], np.asarray(model["normals"], np.float32))
    np.testing.assert_array_equal(a["skeleton"]["tex_coord"], np.asarray(model["tex_coords"], np.float32))
    np.testing.assert_array_equal(a["skeleton"]["bone_index"], np.asarray(model["bone_indexes"], np.uint8))
    np.testing.assert_array_equal(a["skeleton"]["weight"], np.asarray(model["bone_weights"], np.uint8))
    np.testing.assert_array_equal(a["bone_parents"], [0xFFFFFFFF, 0, 1])
    np.testing.assert_array_equal(a["frames_bbox"], bounds[:, :6])
    assert a["names"] == ["body", "skin.png", "root", "bone1", "bone2", "idle"]
    assert a["animations"][0].tolist() == (0, 5, 10.0, 1)
    got = a["frame_bones"].reshape(5, 3, 4, 4).transpose(0, 1, 3, 2).astype(np.float64)  # -> [row][column]
    np.testing.assert_allclose(got, mats, rtol=0, atol=2e-5)


def test_iqm_rejections_follow_the_reference(scene):
    rng = np.random.default_rng(22)
    model = MF.tube_model(rng)
    with pytest.raises(ValueError, match="not a valid IQM|unknown model format"):
        scene.load_model(MF.write_iqm(**model, magic=b"INTERQUAKEMODEX\0")[0])
    with pytest.raises(ValueError, match="version"):
        scene.load_model(MF.write_iqm(**model, version=1)[0])
    with pytest.raises(ValueError, match="no meshes"):
        scene.load_model(MF.write_iqm(**model, mesh_names=())[0])
    with pytest.raises(ValueError, match="joints/poses mismatch"):
        scene.load_model(MF.write_iqm(**model, num_poses=2)[0])
    good = MF.write_iqm(**model)[0]
    with pytest.raises(ValueError, match="truncated"):
        scene.load_model(good[:-6])
    assert scene.info.num_models == 2


def test_loaded_models_go_through_the_model_pipeline(scene, oracle_lib):
    """An MD3 and an IQM file in front of the camera: the mirror animates, lights, clips and sorts them into SQW_CMD_TRIANGLE commands
    that satisfy the rasteriser's preconditions (the oracle's ported debug_asserts) and cover pixels."""
    rng = np.random.default_rng(31)
    md3 = MF.barrel_md3_file(rng)
    _, iqm, _, _ = _iqm_fixture(rng)
    i_md3, i_iqm = scene.load_model(md3), scene.load_model(iqm)
    hr = S.HostRenderer(scene)
    try:
        pose = scene.camera_path(16)[3]
        fwd = np.array([-np.sin(pose[3]), np.cos(pose[3]), 0.0], np.float32)
        models = np.zeros(2, S.model_entity_dtype)
        for k, (mi, frames) in enumerate(((i_md3, (0, 1)), (i_iqm, (1, 3)))):
            models[k]["position"] = pose[:3] + 60.0 * fwd + np.array([0, 0, -20.0 + 25.0 * k], np.float32)
            models[k]["rotation"] = (1, 0, 0, 0)
            models[k]["frames"] = frames
            models[k]["lerp"] = 0.25
            models[k]["model"] = mi
            models[k]["skin"] = k
            models[k]["lighting"] = S.MODEL_LIGHTING_CONSTANT
            models[k]["light"] = (1.0, 0.9, 0.8)
        W, H = 320, 200
        hf = hr.build_frame(W, H, pose, models=models, color_bits=32)
        tri = hf.cmds[hf.cmds["kind"] == N.CMD_TRIANGLE]
        assert len(tri) > 40
        oracle_lib.reset_assert_failures()
        surf = oracle_lib.build_surfaces(32, hf.jobs, None, None, hf.cam_pos, scene.images, scene.lightmaps, scene.directional_lightmaps, hf.total_surface_pixels)
        with_models = np.zeros((H, W), np.uint32)
        oracle_lib.draw(32, with_models, hf.cmds, hf.polygon_points, hf.triangle_points, surf, scene.images)
        assert oracle_lib.assert_failures() == 0
        hf0 = hr.build_frame(W, H, pose, models=None, color_bits=32)
        surf0 = oracle_lib.build_surfaces(32, hf0.jobs, None, None, hf0.cam_pos, scene.images, scene.lightmaps, scene.directional_lightmaps, hf0.total_surface_pixels)
        without = np.zeros((H, W), np.uint32)
        oracle_lib.draw(32, without, hf0.cmds, hf0.polygon_points, hf0.triangle_points, surf0, scene.images)
        assert (with_models != without).sum() > 300  # the two models really are in the picture
    finally:
        hr.close()


@pytest.mark.skipif(not os.path.exists(MRFIXIT), reason="the reference tree (with its one IQM model) is not on this machine")
def test_mrfixit_iqm_reproduces_the_bounds_its_exporter_stored(scene):
    """Known answer from third-party data: the IQM compiler stores, per frame, the bounding box of the skinned vertices.  Skinning
    the loaded vertices with the loaded frame matrices (numpy, float64, parents concatenated at run time like the reference's
    animate_and_transform_vertices) must land on those boxes for all 101 frames — that pins the channel decode, the quaternion
    matrix, translate * scale * rotate and parent_base * pose * inverse_base."""
    index = scene.load_model(MRFIXIT)
    mi = scene.model_info(index)
    assert (mi.num_frames, mi.num_meshes, mi.num_vertices, mi.num_triangles, mi.kind, mi.num_bones) == (101, 1, 1861, 2988, 2, 75)
    a = scene.model_arrays(index)
    assert a["names"][:5] == ["Body", "Body.tga", "Root", "SpineTwist", "Torso"]
    assert a["animations"][0].tolist() == (0, 101, 24.0, 0)
    sk, parents = a["skeleton"], a["bone_parents"]
    nb = len(parents)
    assert ((parents < np.arange(nb)) | (parents >= nb)).all() and parents[0] >= nb  # parents come before their children; roots have none
    assert (sk["weight"].astype(int).sum(1) == 255).all() and a["triangles"].max() < len(sk)
    fb = a["frame_bones"].reshape(-1, nb, 4, 4).transpose(0, 1, 3, 2).astype(np.float64)  # [frame][bone][row][column]
    pos = np.concatenate([sk["position"].astype(np.float64), np.ones((len(sk), 1))], 1)
    w = sk["weight"].astype(np.float64) / 255.0
    worst = 0.0
    for f in range(fb.shape[0]):
        absolute = np.zeros((nb, 4, 4))
        for b in range(nb):
            absolute[b] = absolute[parents[b]] @ fb[f, b] if parents[b] < nb else fb[f, b]
        out = np.zeros((len(sk), 4))
        for k in range(4):
            out += w[:, k:k + 1] * np.einsum("vrc,vc->vr", absolute[sk["bone_index"][:, k]], pos)
        box = a["frames_bbox"][f]
        worst = max(worst, np.abs(out[:, :3].min(0) - box[:3]).max(), np.abs(out[:, :3].max(0) - box[3:]).max())
    assert worst < 0.02, worst  # model height is 8 units; measured 0.008
