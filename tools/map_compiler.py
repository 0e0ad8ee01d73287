#!/usr/bin/env python3
"""map_compiler + lightmapper of the host mirror, with the reference tools' command lines
(/root/reference/src/map_compiler/main.rs:9-36, src/lightmapper/main.rs) so that demo/build_demo_map_2.sh reads the same:

    python tools/map_compiler.py compile -i demo2.map -o demo2.sqwm --materials-dir materials --perform-csg \
        --perform-advanced-splitter-plane-selection
    python tools/map_compiler.py light -i demo2.sqwm -o demo2.sqwm --materials-dir materials --sample-grid-size 2

The work is done by libsqw_host.so (square_wheel_b200/host/map_compiler.cpp, lightmapper.cpp); this file only turns the
materials/*.json of the reference (common/material.rs:7-98, serde defaults) into the table those read.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

BLENDING = {"None": 0, "Average": 1, "Additive": 2, "AlphaTest": 3, "AlphaBlend": 4}


def load_materials(materials_dir: str | None) -> dict:
    """material.rs load_materials: every *.json of the directory, one object of name -> material each."""
    out: dict = {}
    if not materials_dir:
        return out
    for path in sorted(glob.glob(os.path.join(materials_dir, "*.json"))):
        with open(path, "r", encoding="utf-8") as f:
            out.update(json.load(f))
    return out


def materials_table(materials: dict) -> str:
    lines = []
    for name, m in materials.items():
        if any(c.isspace() for c in name):
            continue
        special = m.get("special_effect", "None")
        skybox = isinstance(special, dict) and "Skybox" in special
        e = m.get("emissive_light", [0.0, 0.0, 0.0])
        layer = m.get("emissive_layer") or {}
        ll = layer.get("light", [0.0, 0.0, 0.0])
        lines.append(
            " ".join(
                str(x)
                for x in [
                    name,
                    int(m.get("bsp", True)),
                    int(m.get("twosided", False)),
                    int(m.get("blocks_view", True)),
                    int(m.get("light", True)),
                    int(m.get("shadow", True)),
                    int(m.get("draw", True)),
                    int(m.get("decals", True)),
                    int(skybox),
                    BLENDING[m.get("blending_mode", "None")],
                    int(m.get("normal_map") is not None),
                    float(m.get("roughness", 1.0)),
                    int(m.get("roughness_map") is not None),
                    int(m.get("is_metal", False)),
                    float(e[0]),
                    float(e[1]),
                    float(e[2]),
                    int(bool(layer)),
                    float(ll[0]),
                    float(ll[1]),
                    float(ll[2]),
                ]
            )
        )
    return "\n".join(lines) + "\n"


def _lib():
    from square_wheel_b200 import scene

    l = scene.lib()
    l.sqwh_map_compile.restype = C.c_int
    l.sqwh_map_compile.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint32), C.c_char_p, C.c_uint32]
    l.sqwh_map_light.restype = C.c_int
    l.sqwh_map_light.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.c_float, C.c_float, C.c_int32, C.c_float, C.c_char_p, C.POINTER(C.c_uint32)]
    return l


def leafs_visibility(map_path: str):
    """pvs.rs calculate_visibility_matrix of a .sqwm file as a bool array [num_leafs, num_leafs]."""
    import numpy as np

    l = _lib()
    l.sqwh_map_pvs.restype = C.c_int64
    l.sqwh_map_pvs.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64]
    n = l.sqwh_map_pvs(os.fsencode(map_path), None, 0)
    if n < 0:
        raise OSError(f"sqwh_map_pvs({map_path}): {-n}")
    m = np.zeros((n, n), np.uint8)
    l.sqwh_map_pvs(os.fsencode(map_path), m.ctypes.data_as(C.c_void_p), m.size)
    return m.astype(bool)


COMPILE_STATS = ["initial_polygons", "polygons_after_csg", "leafs_initial", "leafs_removed", "tree_depth", "nodes", "leafs", "polygons", "portals", "submodels", "lightmap_texels"]


def compile_map(map_text: str, table: str, out_path: str, perform_csg=False, advanced=False) -> dict:
    stats = (C.c_uint32 * 11)()
    log = C.create_string_buffer(1 << 16)
    flags = (1 if perform_csg else 0) | (2 if advanced else 0)
    r = _lib().sqwh_map_compile(map_text.encode("utf-8"), table.encode("utf-8"), flags, os.fsencode(out_path), stats, log, len(log))
    if r != 0:
        raise RuntimeError(f"map compile failed ({r}): {log.value.decode(errors='replace')[:400]}")
    d = dict(zip(COMPILE_STATS, [int(x) for x in stats]))
    d["log"] = log.value.decode(errors="replace")
    return d


def light_map(in_path: str, table: str, out_path: str, sample_grid_size=1, light_scale=1.0, ambient_light=0.0, directional=True, light_cutoff=0.0, use_pvs=False) -> dict:
    stats = (C.c_uint32 * 4)()
    r = _lib().sqwh_map_light(os.fsencode(in_path), table.encode("utf-8"), sample_grid_size, light_scale, ambient_light, int(directional) | (2 if use_pvs else 0), light_cutoff, os.fsencode(out_path), stats)
    if r != 0:
        raise RuntimeError(f"lightmapper failed ({r})")
    return dict(zip(["point_lights", "sun_lights", "lightmap_texels", "light_grid_samples"], [int(x) for x in stats]))


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    c = sub.add_parser("compile")
    c.add_argument("-i", dest="input", required=True)
    c.add_argument("-o", dest="output", required=True)
    c.add_argument("--materials-dir")
    c.add_argument("--textures-dir")  # accepted for command-line parity; only the quake4 format reads texture sizes
    c.add_argument("--perform-csg", action="store_true")
    c.add_argument("--perform-advanced-splitter-plane-selection", action="store_true")
    c.add_argument("--print-stats", action="store_true")
    l = sub.add_parser("light")
    l.add_argument("-i", dest="input", required=True)
    l.add_argument("-o", dest="output", required=True)
    l.add_argument("--materials-dir")
    l.add_argument("--textures-dir")
    l.add_argument("--sample-grid-size", type=int, default=1)
    l.add_argument("--light-scale", type=float, default=1.0)
    l.add_argument("--ambient-light", type=float, default=0.0)
    l.add_argument("--no-directional-lightmap", action="store_true")
    l.add_argument("--light-cutoff", type=float, default=0.0)
    l.add_argument("--use-pvs", action="store_true", help="pick a polygon's lights through the leaf-to-leaf visibility matrix (pvs.rs) like the reference, not by the contribution cut-off")
    a = ap.parse_args()
    table = materials_table(load_materials(a.materials_dir))
    t0 = time.time()
    if a.cmd == "compile":
        with open(a.input, "r", encoding="utf-8", errors="replace") as f:
            text = f.read()
        st = compile_map(text, table, a.output, a.perform_csg, a.perform_advanced_splitter_plane_selection)
        log = st.pop("log")
        if log:
            sys.stderr.write(log)
        print(json.dumps({"seconds": round(time.time() - t0, 2), **st}))
    else:
        st = light_map(a.input, table, a.output, a.sample_grid_size, a.light_scale, a.ambient_light, not a.no_directional_lightmap, a.light_cutoff, a.use_pvs)
        print(json.dumps({"seconds": round(time.time() - t0, 2), **st}))


if __name__ == "__main__":
    main()
