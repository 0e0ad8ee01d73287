"""Writers of small MD3 / IQM v2 model files from numpy arrays, and an independent float64 evaluation of the IQM frame matrices
(the published IQM formulas: baseframe / inversebaseframe / frame = parent_base * pose * inverse_base), for the loader tests.
Nothing here shares code with square_wheel_b200/host/model_file.cpp."""
import struct

import numpy as np

MAX_QPATH = 64


def _qpath(s: str, n: int = MAX_QPATH) -> bytes:
    b = s.encode()
    assert len(b) < n
    return b + b"\0" * (n - len(b))


# ---- MD3 ---------------------------------------------------------------------------------------------------
def write_md3(meshes, frames_bbox, tags=None, version=15, ident=b"IDP3") -> bytes:
    """meshes: list of dict(name, shader, triangles u32[t][3], tex_coords f32[v][2], origins i16[frames][v][3], normals i16[frames][v]);
    frames_bbox: f32[frames][6]; tags: list over frames of list of (name, origin[3], rotation[9])."""
    num_frames = len(frames_bbox)
    tags = tags or [[] for _ in range(num_frames)]
    num_tags = len(tags[0])
    header_size = 108
    frames_blob = b"".join(struct.pack("<3f3f3ff16s", *bb[:3], *bb[3:], 0.0, 0.0, 0.0, 1.0, b"frame") for bb in frames_bbox)
    tags_blob = b"".join(_qpath(n) + struct.pack("<3f9f", *o, *r) for frame in tags for (n, o, r) in frame)
    mesh_blobs = []
    for m in meshes:
        tri = np.asarray(m["triangles"], "<u4")
        tc = np.asarray(m["tex_coords"], "<f4")
        org = np.asarray(m["origins"], "<i2")
        nrm = np.asarray(m["normals"], "<i2")
        nf, nv = org.shape[0], org.shape[1]
        verts = np.zeros((nf, nv, 4), "<i2")
        verts[:, :, :3] = org
        verts[:, :, 3] = nrm
        shaders = _qpath(m.get("shader", "")) + struct.pack("<I", 0) if m.get("shader") is not None else b""
        ofs_tri = 108
        ofs_shaders = ofs_tri + tri.nbytes
        ofs_tc = ofs_shaders + len(shaders)
        ofs_verts = ofs_tc + tc.nbytes
        end = ofs_verts + verts.nbytes
        hdr = m.get("ident", b"IDP3") + _qpath(m["name"]) + struct.pack("<I4I5I", 0, nf, 1 if shaders else 0, nv, len(tri), ofs_tri, ofs_shaders, ofs_tc, ofs_verts, end)
        assert len(hdr) == 108
        mesh_blobs.append(hdr + tri.tobytes() + shaders + tc.tobytes() + verts.tobytes())
    ofs_frames = header_size
    ofs_tags = ofs_frames + len(frames_blob)
    ofs_meshes = ofs_tags + len(tags_blob)
    end = ofs_meshes + sum(len(b) for b in mesh_blobs)
    hdr = ident + struct.pack("<I", version) + _qpath("model") + struct.pack("<I4I4I", 0, num_frames, num_tags, len(meshes), 0, ofs_frames, ofs_tags, ofs_meshes, end)
    assert len(hdr) == header_size
    return hdr + frames_blob + tags_blob + b"".join(mesh_blobs)


def md3_normal(code: np.ndarray) -> np.ndarray:
    """Quake III latitude / longitude normal: low byte = pitch, high byte = yaw, 256 steps per turn."""
    c = np.asarray(code).astype(np.int64) & 0xFFFF
    pitch = (c & 255) * (2.0 * np.pi / 256.0)
    yaw = ((c >> 8) & 255) * (2.0 * np.pi / 256.0)
    return np.stack([np.sin(pitch) * np.cos(yaw), np.sin(pitch) * np.sin(yaw), np.cos(pitch)], -1)


# ---- IQM ---------------------------------------------------------------------------------------------------
IQM_POSITION, IQM_TEXCOORD, IQM_NORMAL, IQM_TANGENT, IQM_BLENDINDEXES, IQM_BLENDWEIGHTS = 0, 1, 2, 3, 4, 5
IQM_UBYTE, IQM_FLOAT = 1, 7


def quat_matrix(q):
    """Rotation matrix of the unit quaternion (x, y, z, w), float64."""
    x, y, z, w = [float(v) for v in q]
    n = np.sqrt(x * x + y * y + z * z + w * w)
    x, y, z, w = x / n, y / n, z / n, w / n
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def joint_matrix(translate, rotate_xyzw, scale):
    """translate * scale * rotate, the order the reference composes (triangle_model_iqm.rs get_bone_matrix)."""
    m = np.eye(4)
    m[:3, :3] = np.diag(np.asarray(scale, np.float64)) @ quat_matrix(rotate_xyzw)
    m[:3, 3] = translate
    return m


def write_iqm(positions, normals, tex_coords, bone_indexes, bone_weights, triangles, joints, frames, bounds=None, anims=(("idle", 0, None, 10.0, 1),),
              mesh_names=(("body", "skin.png"),), version=2, magic=b"INTERQUAKEMODEL\0", num_poses=None):
    """joints: list of (name, parent (-1 = none), translate[3], rotate xyzw[4], scale[3]); frames: f32[num_frames][num_joints][10] channel values
    (tx ty tz qx qy qz qw sx sy sz) quantised to u16 per channel like the IQM compiler does.  Returns (bytes, decoded channel values f32[frames][joints][10])."""
    positions = np.asarray(positions, "<f4")
    nv = len(positions)
    frames = np.asarray(frames, np.float64)
    nf, nj = frames.shape[0], len(joints)
    text = bytearray(b"\0")

    def name(s):
        o = len(text)
        text.extend(s.encode() + b"\0")
        return o

    meshes_blob = b"".join(struct.pack("<6I", name(n), name(mat), 0, nv, 0, len(triangles)) for n, mat in mesh_names)
    joints_blob = b"".join(struct.pack("<Ii3f4f3f", name(n), p, *t, *r, *s) for (n, p, t, r, s) in joints)
    # poses: offset = channel minimum, scale = range / 65535, a channel is animated when it varies
    lo, hi = frames.min(0), frames.max(0)
    scale = np.where(hi > lo, (hi - lo) / 65535.0, 0.0).astype(np.float32)
    offset = lo.astype(np.float32)
    poses_blob = b""
    frame_values = []
    decoded = np.zeros((nf, nj, 10), np.float32)
    masks = []
    for j in range(nj):
        mask = 0
        for c in range(10):
            if scale[j, c] != 0:
                mask |= 1 << c
        masks.append(mask)
        poses_blob += struct.pack("<iI10f10f", joints[j][1], mask, *offset[j], *scale[j])
    for f in range(nf):
        for j in range(nj):
            for c in range(10):
                if masks[j] & (1 << c):
                    q = int(np.clip(np.rint((frames[f, j, c] - float(offset[j, c])) / float(scale[j, c])), 0, 65535))
                    frame_values.append(q)
                    decoded[f, j, c] = np.float32(offset[j, c] + np.float32(q) * scale[j, c])
                else:
                    decoded[f, j, c] = offset[j, c]
    frames_blob = np.asarray(frame_values, "<u2").tobytes()
    anims_blob = b"".join(struct.pack("<3IfI", name(n), first, nf if count is None else count, fps, flags) for (n, first, count, fps, flags) in anims)
    arrays = [
        (IQM_POSITION, IQM_FLOAT, 3, positions.tobytes()),
        (IQM_TEXCOORD, IQM_FLOAT, 2, np.asarray(tex_coords, "<f4").tobytes()),
        (IQM_NORMAL, IQM_FLOAT, 3, np.asarray(normals, "<f4").tobytes()),
        (IQM_TANGENT, IQM_FLOAT, 4, np.zeros((nv, 4), "<f4").tobytes()),  # ignored by the loader
        (IQM_BLENDINDEXES, IQM_UBYTE, 4, np.asarray(bone_indexes, "u1").tobytes()),
        (IQM_BLENDWEIGHTS, IQM_UBYTE, 4, np.asarray(bone_weights, "u1").tobytes()),
    ]
    tri_blob = np.asarray(triangles, "<u4").tobytes()
    if bounds is None:
        bounds = np.zeros((nf, 8), "<f4")
    bounds_blob = np.asarray(bounds, "<f4").tobytes()

    header_size = 124
    pos = header_size
    layout = {}

    def place(key, blob, align=4):
        nonlocal pos
        pos = (pos + align - 1) // align * align
        layout[key] = pos
        pos += len(blob)

    while len(text) % 4:
        text.append(0)
    place("text", bytes(text))
    place("meshes", meshes_blob)
    place("vertexarrays", b"\0" * (20 * len(arrays)))
    for i, a in enumerate(arrays):
        place(("va", i), a[3])
    place("triangles", tri_blob)
    place("joints", joints_blob)
    place("poses", poses_blob)
    place("anims", anims_blob)
    place("bounds", bounds_blob)
    place("frames", frames_blob)  # last: the loader reads channel data from here to the end of the file
    total = pos
    va_blob = b"".join(struct.pack("<5I", t, 0, fmt, size, layout[("va", i)]) for i, (t, fmt, size, _) in enumerate(arrays))
    npos = nj if num_poses is None else num_poses
    hdr = magic + struct.pack(
        "<27I", version, total, 0, len(text), layout["text"], len(mesh_names), layout["meshes"], len(arrays), nv, layout["vertexarrays"], len(triangles), layout["triangles"], 0,
        nj, layout["joints"], npos, layout["poses"], len(anims), layout["anims"], nf, len(frame_values) // max(nf, 1), layout["frames"], layout["bounds"] if nf else 0, 0, 0, 0, 0)
    assert len(hdr) == header_size
    out = bytearray(total)
    out[:header_size] = hdr
    for key, blob in [("text", bytes(text)), ("meshes", meshes_blob), ("vertexarrays", va_blob), ("triangles", tri_blob), ("joints", joints_blob), ("poses", poses_blob),
                      ("anims", anims_blob), ("bounds", bounds_blob), ("frames", frames_blob)] + [(("va", i), a[3]) for i, a in enumerate(arrays)]:
        out[layout[key]:layout[key] + len(blob)] = blob
    return bytes(out), decoded


def iqm_frame_matrices(joints, decoded):
    """float64 frame matrices [frames][joints][4][4] (row-major) from the decoded channel values: parent_base * (pose * inverse_base)."""
    nj = len(joints)
    base, inv = [None] * nj, [None] * nj
    for j, (_, parent, t, r, s) in enumerate(joints):
        m = joint_matrix(t, r, s)
        base[j] = base[parent] @ m if parent >= 0 else m
        inv[j] = np.linalg.inv(base[j])
    out = np.zeros((decoded.shape[0], nj, 4, 4))
    for f in range(decoded.shape[0]):
        for j, (_, parent, *_rest) in enumerate(joints):
            c = decoded[f, j].astype(np.float64)
            m = joint_matrix(c[0:3], c[3:7], c[7:10]) @ inv[j]
            out[f, j] = base[parent] @ m if parent >= 0 else m
    return out


def tube_model(rng, rings=7, segments=8, num_joints=3, num_frames=5, length=48.0, radius=7.0):
    """A bending tube along +z with a joint chain: arrays for write_iqm."""
    positions, normals, tcs, bi, bw, tris = [], [], [], [], [], []
    for r in range(rings):
        z = length * r / (rings - 1)
        jf = (num_joints - 1) * r / (rings - 1)
        j0 = min(int(jf), num_joints - 2)
        w1 = int(round(255 * (jf - j0)))
        for s in range(segments):
            a = 2 * np.pi * s / segments
            positions.append((radius * np.cos(a), radius * np.sin(a), z))
            normals.append((np.cos(a), np.sin(a), 0.0))
            tcs.append((s / segments, r / (rings - 1)))
            bi.append((j0, j0 + 1, 0, 0))
            bw.append((255 - w1, w1, 0, 0))
    for r in range(rings - 1):
        for s in range(segments):
            a, b = r * segments + s, r * segments + (s + 1) % segments
            c, d = a + segments, b + segments
            tris += [(a, b, c), (b, d, c)]
    seg = length / (num_joints - 1)
    joints = [("root", -1, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0))]
    for j in range(1, num_joints):
        joints.append((f"bone{j}", j - 1, (0.0, 0.0, seg), (0.0, 0.0, 0.0, 1.0), (1.0, 1.0, 1.0)))
    frames = np.zeros((num_frames, num_joints, 10))
    for f in range(num_frames):
        for j in range(num_joints):
            ang = 0.35 * np.sin(0.9 * f + j) * (j > 0)
            axis = np.array([np.cos(0.5 * j), np.sin(0.5 * j), 0.0])
            frames[f, j, 0:3] = np.asarray(joints[j][2]) + (rng.uniform(-0.5, 0.5, 3) if j == 0 else 0.0)
            frames[f, j, 3:6] = axis * np.sin(ang / 2)
            frames[f, j, 6] = np.cos(ang / 2)
            frames[f, j, 7:10] = 1.0 + (0.1 * np.sin(f) if j == 1 else 0.0)
    return dict(positions=positions, normals=normals, tex_coords=tcs, bone_indexes=bi, bone_weights=bw, triangles=tris, joints=joints, frames=frames)


def skinned_positions(model, mats):
    """float64 positions [frames][v][3] of the tube under hierarchical frame matrices (the run-time parent concatenation of
    triangle_models_rendering.rs:118-133)."""
    pos = np.concatenate([np.asarray(model["positions"], np.float64), np.ones((len(model["positions"]), 1))], 1)
    bi, bw = np.asarray(model["bone_indexes"]), np.asarray(model["bone_weights"], np.float64) / 255.0
    out = np.zeros((mats.shape[0], len(pos), 3))
    for f in range(mats.shape[0]):
        absolute = []
        for j, (_, parent, *_r) in enumerate(model["joints"]):
            absolute.append(absolute[parent] @ mats[f, j] if parent >= 0 else mats[f, j])
        absolute = np.asarray(absolute)
        for k in range(4):
            out[f] += bw[:, k:k + 1] * np.einsum("vrc,vc->vr", absolute[bi[:, k]], pos)[:, :3]
    return out


def tube_iqm_file(rng, **kw):
    """(arrays, file bytes, float64 frame matrices, per-frame bounds) of a bending tube with correct bounding boxes."""
    model = tube_model(rng, **kw)
    _, decoded = write_iqm(**model)
    mats = iqm_frame_matrices(model["joints"], decoded)
    skinned = skinned_positions(model, mats)
    bounds = np.zeros((len(mats), 8), np.float32)
    bounds[:, :3] = skinned.min(1) - 0.01
    bounds[:, 3:6] = skinned.max(1) + 0.01
    data, decoded = write_iqm(**model, bounds=bounds)
    return model, data, mats, bounds


def barrel_md3_file(rng, num_frames=2):
    """A two-frame vertex-animated barrel (one mesh, 24 vertices, 24 triangles) as MD3 bytes."""
    nv = 24
    h = nv // 2
    ang = np.linspace(0, 2 * np.pi, h, endpoint=False)
    ring = np.stack([np.cos(ang), np.sin(ang)], 1)
    origins = np.zeros((num_frames, nv, 3))
    for f in range(num_frames):
        origins[f, :h] = np.concatenate([ring * (600 + 200 * f), np.full((h, 1), -900.0)], 1)
        origins[f, h:] = np.concatenate([ring * (500 - 100 * f), np.full((h, 1), 900.0)], 1)
    tris = []
    for s in range(h):
        a, b = s, (s + 1) % h
        tris += [(a, b, a + h), (b, b + h, a + h)]
    normals = rng.integers(0, 65535, (num_frames, nv)).astype(np.uint16).view(np.int16)
    return write_md3([dict(name="barrel", shader="skin", triangles=tris, tex_coords=rng.uniform(0, 1, (nv, 2)), origins=origins, normals=normals)],
                     np.array([[-13, -13, -15, 13, 13, 15]] * num_frames, np.float32))
