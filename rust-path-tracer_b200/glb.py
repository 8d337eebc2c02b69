"""Minimal binary-glTF (.glb) scene import for the tracing hot path's input buffers.

Host-side input producer.  The reference imports scenes through assimp
(`World::from_path`, src/asset.rs:55-133); assimp does not exist in this image, so this module
restates what that import produces for the constructs the shipped scenes use (triangle
primitives, float attributes, node TRS / matrix hierarchy, pbrMetallicRoughness factors and
textures):

* world-space bake through the node hierarchy, position swizzle ``(x, z, y)`` and the winding
  swap ``(f0, f2, f1)`` with the material index in ``.w`` (src/asset.rs:101-106);
* normals / tangents rotated by the node rotation after dividing by the node scale, same
  swizzle (src/asset.rs:108-115);
* assimp's glTF2 conventions: V flipped to ``1 - v``; a default material (baseColor 1,
  metallic 1, roughness 1) appended after the file's materials and used by primitives without
  one; ``metallicFactor`` / ``roughnessFactor`` default to 1;
* the material mapping of src/asset.rs:135-175: ``$clr.diffuse`` -> albedo,
  ``$clr.emissive`` x 15 -> emissive, scalar factors splatted.

Not reproduced (none changes a rendered value): JoinIdenticalVertices and
ImproveCacheLocality (vertex / face re-ordering only).
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field

import numpy as np

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}

MATERIAL_DTYPE = np.dtype(
    [
        ("emissive", "<f4", 4),
        ("albedo", "<f4", 4),
        ("roughness", "<f4", 4),
        ("metallic", "<f4", 4),
        ("normals", "<f4", 4),
        ("has_albedo_texture", "<u4"),
        ("has_metallic_texture", "<u4"),
        ("has_roughness_texture", "<u4"),
        ("has_normal_texture", "<u4"),
    ]
)
assert MATERIAL_DTYPE.itemsize == 96

VERTEX_DTYPE = np.dtype([("vertex", "<f4", 4), ("normal", "<f4", 4), ("tangent", "<f4", 4), ("uv0", "<f4", 2), ("uv1", "<f4", 2)])
assert VERTEX_DTYPE.itemsize == 64


@dataclass
class BakedScene:
    """Flat, pre-BVH scene arrays — what `walk_node_graph` + the material loop leave behind."""

    vertices: np.ndarray  # (V,4) f32, w = 1
    normals: np.ndarray  # (V,4) f32, w = 0
    tangents: np.ndarray  # (V,4) f32, w = 0
    uvs: np.ndarray  # (V,2) f32
    indices: np.ndarray  # (T,4) u32: i0,i1,i2,material
    materials: np.ndarray  # (M,) MATERIAL_DTYPE
    # decoded texture images per material: {"albedo"|"metallic"|"roughness"|"normals": HxWx4 u8}
    textures: list = field(default_factory=list)

    def save(self, path: str) -> None:
        np.savez_compressed(
            path,
            vertices=self.vertices,
            normals=self.normals,
            tangents=self.tangents,
            uvs=self.uvs,
            indices=self.indices,
            materials=self.materials.view(np.uint8).reshape(-1, 96),
        )

    def save_rptw(self, path: str, atlas: np.ndarray | None = None) -> None:
        """Flat little-endian container read by the C++ host (host/trace.cpp, World::from_path)."""
        aw, ah = (0, 0) if atlas is None else (atlas.shape[1], atlas.shape[0])
        with open(path, "wb") as f:
            f.write(b"RPTW0001")
            f.write(np.array([len(self.vertices), len(self.indices), len(self.materials), aw, ah], "<u4").tobytes())
            for arr, dt in ((self.vertices, "<f4"), (self.normals, "<f4"), (self.tangents, "<f4"), (self.uvs, "<f4"), (self.indices, "<u4")):
                f.write(np.ascontiguousarray(arr, dt).tobytes())
            f.write(np.ascontiguousarray(self.materials).tobytes())
            if atlas is not None:
                f.write(np.ascontiguousarray(atlas, np.uint8).tobytes())

    @staticmethod
    def load(path: str) -> "BakedScene":
        z = np.load(path)
        mats = np.ascontiguousarray(z["materials"]).view(MATERIAL_DTYPE).reshape(-1)
        return BakedScene(z["vertices"], z["normals"], z["tangents"], z["uvs"], z["indices"], mats.copy())


def _read_chunks(blob: bytes):
    magic, version, _length = struct.unpack_from("<III", blob, 0)
    if magic != 0x46546C67 or version != 2:
        raise ValueError("not a glTF 2.0 binary file")
    off, doc, binary = 12, None, b""
    while off < len(blob):
        clen, ctype = struct.unpack_from("<II", blob, off)
        body = blob[off + 8 : off + 8 + clen]
        if ctype == 0x4E4F534A:
            doc = json.loads(body)
        elif ctype == 0x004E4942:
            binary = body
        off += 8 + clen
    if doc is None:
        raise ValueError("glb has no JSON chunk")
    return doc, binary


def _accessor(doc, binary, idx) -> np.ndarray:
    acc = doc["accessors"][idx]
    if "sparse" in acc:
        raise NotImplementedError("sparse accessors")
    view = doc["bufferViews"][acc["bufferView"]]
    dt = np.dtype(_COMPONENT[acc["componentType"]])
    ncomp = _NCOMP[acc["type"]]
    start = view.get("byteOffset", 0) + acc.get("byteOffset", 0)
    stride = view.get("byteStride") or dt.itemsize * ncomp
    count = acc["count"]
    if stride == dt.itemsize * ncomp:
        arr = np.frombuffer(binary, dt, count * ncomp, start).reshape(count, ncomp)
    else:
        raw = np.frombuffer(binary, np.uint8, offset=start)
        arr = np.lib.stride_tricks.as_strided(raw, (count, dt.itemsize * ncomp), (stride, 1)).copy().view(dt).reshape(count, ncomp)
    if acc.get("normalized") and dt != np.float32:
        info = np.iinfo(dt)
        arr = np.maximum(arr.astype(np.float32) / np.float32(info.max), np.float32(-1.0))
    return arr


def _node_matrix(node) -> np.ndarray:
    """Local transform as a 4x4 float32 matrix (column-vector convention: p' = M @ p)."""
    if "matrix" in node:
        return np.asarray(node["matrix"], np.float32).reshape(4, 4).T.copy()
    f = np.float32
    t = np.asarray(node.get("translation", (0, 0, 0)), f)
    x, y, z, w = (f(c) for c in node.get("rotation", (0, 0, 0, 1)))
    s = np.asarray(node.get("scale", (1, 1, 1)), f)
    one, two = f(1), f(2)
    rot = np.array(
        [
            [one - two * (y * y + z * z), two * (x * y - z * w), two * (x * z + y * w)],
            [two * (x * y + z * w), one - two * (x * x + z * z), two * (y * z - x * w)],
            [two * (x * z - y * w), two * (y * z + x * w), one - two * (x * x + y * y)],
        ],
        f,
    )
    m = np.eye(4, dtype=f)
    m[:3, :3] = rot * s[None, :]
    m[:3, 3] = t
    return m


def _scale_rotation(m: np.ndarray):
    """glam `Mat4::to_scale_rotation_translation`: per-axis lengths (x signed by det) + rotation."""
    f = np.float32
    det = f(np.linalg.det(m[:3, :3].astype(np.float64)))
    sx = f(np.linalg.norm(m[:3, 0])) * (f(-1) if det < 0 else f(1))
    sy = f(np.linalg.norm(m[:3, 1]))
    sz = f(np.linalg.norm(m[:3, 2]))
    scale = np.array([sx, sy, sz], f)
    rot = (m[:3, :3] / scale[None, :]).astype(f)
    return scale, rot


def _normalize_rows(a: np.ndarray) -> np.ndarray:
    a = a.astype(np.float32)
    length = np.sqrt((a * a).sum(axis=1, dtype=np.float32)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (a * (np.float32(1.0) / length)[:, None]).astype(np.float32)


def _tangents(pos, uv, nrm, faces) -> np.ndarray:
    """Per-vertex tangents from UV derivatives (stand-in for assimp CalculateTangentSpace)."""
    tan = np.zeros((len(pos), 3), np.float64)
    p0, p1, p2 = (pos[faces[:, k]].astype(np.float64) for k in range(3))
    w0, w1, w2 = (uv[faces[:, k]].astype(np.float64) for k in range(3))
    e1, e2 = p1 - p0, p2 - p0
    d1, d2 = w1 - w0, w2 - w0
    det = d1[:, 0] * d2[:, 1] - d2[:, 0] * d1[:, 1]
    det = np.where(np.abs(det) < 1e-20, 1.0, det)
    t = (e1 * d2[:, 1:2] - e2 * d1[:, 1:2]) / det[:, None]
    for k in range(3):
        np.add.at(tan, faces[:, k], t)
    n = nrm.astype(np.float64)
    tan = tan - n * (tan * n).sum(axis=1, keepdims=True)
    ln = np.linalg.norm(tan, axis=1, keepdims=True)
    tan = np.where(ln > 1e-20, tan / np.where(ln > 1e-20, ln, 1.0), 0.0)
    return tan.astype(np.float32)


def _decode_image(doc, binary, image_index):
    from io import BytesIO

    from PIL import Image

    img = doc["images"][image_index]
    view = doc["bufferViews"][img["bufferView"]]
    start = view.get("byteOffset", 0)
    data = binary[start : start + view["byteLength"]]
    return np.asarray(Image.open(BytesIO(data)).convert("RGBA"), np.uint8)


def load_glb(path: str) -> BakedScene:
    doc, binary = _read_chunks(open(path, "rb").read())
    verts, nrms, tans, uvs, tris = [], [], [], [], []
    n_file_materials = len(doc.get("materials", []))
    vertex_count = 0

    def walk(node_index: int, parent: np.ndarray):
        nonlocal vertex_count
        node = doc["nodes"][node_index]
        trs = (parent @ _node_matrix(node)).astype(np.float32)
        if "mesh" in node:
            scale, rot = _scale_rotation(trs)
            for prim in doc["meshes"][node["mesh"]]["primitives"]:
                if prim.get("mode", 4) != 4:
                    continue  # SortByPrimitiveType + the `assert_eq!(f.0.len(), 3)`: triangles only
                attrs = prim["attributes"]
                pos = _accessor(doc, binary, attrs["POSITION"]).astype(np.float32)
                n = len(pos)
                if "indices" in prim:
                    faces = _accessor(doc, binary, prim["indices"]).astype(np.uint32).reshape(-1, 3)
                else:
                    faces = np.arange(n, dtype=np.uint32).reshape(-1, 3)
                hom = np.concatenate([pos, np.ones((n, 1), np.float32)], axis=1)
                world = (hom @ trs.T).astype(np.float32)
                verts.append(np.stack([world[:, 0], world[:, 2], world[:, 1], np.ones(n, np.float32)], axis=1))
                if "NORMAL" in attrs:
                    nraw = _accessor(doc, binary, attrs["NORMAL"]).astype(np.float32)
                else:  # GenerateSmoothNormals: area-weighted vertex normals
                    fn = np.cross(pos[faces[:, 1]] - pos[faces[:, 0]], pos[faces[:, 2]] - pos[faces[:, 0]])
                    nraw = np.zeros_like(pos)
                    for k in range(3):
                        np.add.at(nraw, faces[:, k], fn)
                    nraw = _normalize_rows(nraw)
                if "TEXCOORD_0" in attrs:
                    uv = _accessor(doc, binary, attrs["TEXCOORD_0"]).astype(np.float32)
                    uv = np.stack([uv[:, 0], np.float32(1.0) - uv[:, 1]], axis=1)
                else:
                    uv = np.zeros((n, 2), np.float32)
                traw = _tangents(pos, uv, nraw, faces)
                for raw, out in ((nraw, nrms), (traw, tans)):
                    w = _normalize_rows((raw / scale[None, :]).astype(np.float32) @ rot.T)
                    out.append(np.stack([w[:, 0], w[:, 2], w[:, 1], np.zeros(n, np.float32)], axis=1))
                uvs.append(uv)
                mat = prim.get("material", n_file_materials)
                off = np.uint32(vertex_count)
                tris.append(
                    np.stack([faces[:, 0] + off, faces[:, 2] + off, faces[:, 1] + off, np.full(len(faces), mat, np.uint32)], axis=1)
                )
                vertex_count += n
        for child in node.get("children", []):
            walk(child, trs)

    scene = doc["scenes"][doc.get("scene", 0)]
    for root in scene["nodes"]:
        walk(root, np.eye(4, dtype=np.float32))

    materials = np.zeros(n_file_materials + 1, MATERIAL_DTYPE)
    textures = [dict() for _ in range(n_file_materials + 1)]
    for i in range(n_file_materials + 1):
        m = doc["materials"][i] if i < n_file_materials else {}
        pbr = m.get("pbrMetallicRoughness", {})
        materials[i]["albedo"] = np.asarray(pbr.get("baseColorFactor", (1, 1, 1, 1)), np.float32)
        em = np.asarray(list(m.get("emissiveFactor", (0, 0, 0))) + [1.0], np.float32)
        materials[i]["emissive"] = em * np.float32(15.0)  # src/asset.rs:165-168
        materials[i]["metallic"] = np.float32(pbr.get("metallicFactor", 1.0))
        materials[i]["roughness"] = np.float32(pbr.get("roughnessFactor", 1.0))

        def tex(info):
            if info is None or "textures" not in doc:
                return None
            src = doc["textures"][info["index"]].get("source")
            return None if src is None else _decode_image(doc, binary, src)

        img = tex(pbr.get("baseColorTexture"))
        if img is not None:
            textures[i]["albedo"] = img
        img = tex(pbr.get("metallicRoughnessTexture"))
        if img is not None:  # assimp exposes the one glTF image as both METALNESS and DIFFUSE_ROUGHNESS
            textures[i]["metallic"] = img
            textures[i]["roughness"] = img
        img = tex(m.get("normalTexture"))
        if img is not None:
            textures[i]["normals"] = img

    return BakedScene(
        np.concatenate(verts).astype(np.float32),
        np.concatenate(nrms).astype(np.float32),
        np.concatenate(tans).astype(np.float32),
        np.concatenate(uvs).astype(np.float32),
        np.concatenate(tris).astype(np.uint32),
        materials,
        textures,
    )
