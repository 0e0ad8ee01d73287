"""Triangle-model preparation on the GPU: sqw_models_prepare against the host mirror of
PartialRenderer::build_dynamic_models_buffers (partial_renderer.rs:1122-1190) — animate_and_transform_triangle_mesh_vertices
(triangle_models_rendering.rs:7-203), reject_triangle_model_back_faces (:327-345), sort_model_triangles (:347-364).

The host mirror (host/models.cpp) and the kernels (csrc/model_kernels.cu) evaluate the same float expressions in the same order
without fused multiply-add, so vertices, visible-triangle counts and triangle order are compared BIT FOR BIT, for the three
VertexData kinds: non-animated, vertex-animated (MD3 class), skeleton-animated (IQM class, a bone chain)."""
import numpy as np
import pytest

from square_wheel_b200 import _native as N
from square_wheel_b200 import scene as S
from square_wheel_b200.device import Context
from test_frames_gpu import _lights_along_path, _models_ahead_of_camera

pytestmark = pytest.mark.gpu


def _instances_from(records, mesh_ids):
    inst = np.zeros(len(records), N.model_instance_dtype)
    for k in ("frames", "lerp", "model_matrix", "model_view_matrix", "tc_scale", "tc_shift", "light_cube", "light_vector_scaled", "light_color"):
        inst[k] = records[k]
    inst["mesh"] = [mesh_ids[(int(r["model"]), int(r["mesh"]))] for r in records]
    inst["vertex_offset"] = records["vertices_offset"]
    inst["triangle_offset"] = records["triangles_offset"]
    return inst


@pytest.mark.parametrize("detail", [9, 27])
def test_models_prepare_matches_the_host_mirror_bit_for_bit(detail):
    sc = S.Scene(seed=7, texture_size_log2=6, num_materials=8, num_models=6, model_detail=detail, light_grid=1)
    ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0)
    hr = S.HostRenderer(sc, use_directional_lightmaps=True)
    try:
        mesh_ids = {}
        kinds = set()
        for m in range(int(sc.info.num_models)):
            for mesh in range(sc.model_info(m).num_meshes):
                arrays = sc.model_arrays(m, mesh)
                mesh_ids[(m, mesh)] = ctx.upload_model_mesh(arrays)
                kinds.add(0 if len(arrays["non_animated"]) else (1 if len(arrays["constant"]) else 2))
        assert kinds == {0, 1, 2}
        view = ctx.create_view(64, 64, 64)
        path = sc.camera_path(64)
        total_triangles = total_visible = 0
        for k in (0, 9, 31, 50):
            models = _models_ahead_of_camera(sc, path[k], k)
            lights = _lights_along_path(path, 4, 5)
            lights["shadow_type"] = N.SHADOW_NONE
            lights["position"] = path[k][:3] + np.array([[20, 30, 10], [-40, 60, 0], [10, 90, -20], [0, 10, 30]], np.float32)
            hr.build_frame(640, 360, path[k], lights=lights, models=models)
            rec, want_v, want_t = hr.dynamic_meshes()
            assert len(rec) >= 4  # the models really are in view
            view.models_prepare(_instances_from(rec, mesh_ids))
            got_v, got_t, got_n = view.models_read(len(want_v), len(want_t), len(rec))
            assert got_n.tolist() == rec["num_visible_triangles"].tolist(), f"frame {k}: visible triangle counts"
            for r in rec:
                vo, nv, to, nt = int(r["vertices_offset"]), int(r["num_vertices"]), int(r["triangles_offset"]), int(r["num_visible_triangles"])
                a = got_v[vo : vo + nv].view(np.uint32).reshape(nv, 8)
                b = want_v[vo : vo + nv].view(np.uint32).reshape(nv, 8)
                bad = np.flatnonzero((a != b).any(axis=1))
                assert bad.size == 0, f"frame {k} model {int(r['model'])}: {bad.size} of {nv} vertices differ, first {bad[0]}: {got_v[vo + bad[0]]} vs {want_v[vo + bad[0]]}"
                assert np.array_equal(got_t[to : to + nt], want_t[to : to + nt]), f"frame {k} model {int(r['model'])}: triangle order"
                total_triangles += int(r["num_triangles"])
                total_visible += nt
        assert 0 < total_visible < total_triangles  # back faces were rejected, front faces kept
        view.close()
    finally:
        hr.close()
        ctx.close()
        sc.close()


def test_models_prepare_rejects_bad_references():
    from square_wheel_b200.device import SqwError

    sc = S.Scene(seed=3, texture_size_log2=5, num_materials=4, num_models=3, model_detail=5)
    ctx = Context(sc.lightmaps, sc.directional_lightmaps, device=0)
    try:
        mid = ctx.upload_model_mesh(sc.model_arrays(1, 0))
        view = ctx.create_view(32, 32, 32)
        inst = np.zeros(1, N.model_instance_dtype)
        inst["mesh"] = mid + 5
        with pytest.raises(SqwError):
            view.models_prepare(inst)
        inst["mesh"] = mid
        inst["frames"] = (0, 10000)
        with pytest.raises(SqwError):
            view.models_prepare(inst)
        bad = dict(sc.model_arrays(1, 0))
        bad["triangles"] = bad["triangles"].copy()
        bad["triangles"][0, 0] = 65535
        with pytest.raises(SqwError):
            ctx.upload_model_mesh(bad)
        view.close()
    finally:
        ctx.close()
        sc.close()
