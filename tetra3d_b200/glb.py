"""A small GLB (binary glTF 2.0) reader that produces a `SceneDesc` the way the reference's loader would build its scene.

Follows gltf.go (LoadGLTFData) for the parts that feed the geometry stage:
  materials   BackfaceCulling = !doubleSided (gltf.go:330); the t3d* extras (colour, shadeless, lighting / transparency /
              sort-related switches, gltf.go:361-485)
  meshes      one MeshPart per primitive (gltf.go:583-730): POSITION, NORMAL, TEXCOORD_0 with V' = -(V - 1) (gltf.go:616),
              COLOR_0 as bytes / 255 through ConvertTosRGB (gltf.go:657-671, color.go:196-218), WEIGHTS_0 / JOINTS_0 keeping only
              the influences whose weight is > 0, in order (gltf.go:692-700); morph targets as shape keys holding the glTF
              deltas (gltf.go:749-792, mesh.go:796-822)
  nodes       TRS -> SetLocalPosition / Scale / Rotation(Quaternion.ToMatrix4()) (gltf.go:1407-1409, quaternion.go:135-184);
              KHR_lights_punctual: sun energy = intensity, point / other energy = intensity / 80, colours through
              ConvertTosRGB (gltf.go:1130-1152); cameras (gltf.go:1083-1096)
  skins       joints in skin order, inverse bind matrices (gltf.go:1476-1500), the armature = the node named like the skin
  animations  one channel per target node name, one track per TRS path, keyframe times / values as stored (gltf.go:844-957)
  world       t3dWorlds__ ambient colour (ConvertTosRGB) and energy (gltf.go:1556-1571)

The reference reads accessors through github.com/qmuntal/gltf v0.28.1 (go.mod:7; not vendored): its `modeler` returns a MAT4
accessor as [][4][4]float32 indexed [row][column] of the mathematical matrix, and gltf.go:1492-1497 stores row i as COLUMN i
of its own row-vector Matrix4 -- i.e. Tetra3D's matrix is the accessor's 16 floats read as a row-major 4x4 (the glTF file holds
them column-major).  node.Matrix (gltf.go:1389-1395) is handled the same way.

Only what the two reference assets under examples/ need is implemented (float / ubyte / ushort / uint accessors, sparse
accessors, embedded buffer); anything else raises.
"""
from __future__ import annotations

import json
import struct
from typing import Dict, List, Optional

import numpy as np

from .scenedesc import (LIGHT_AMBIENT, LIGHT_DIRECTIONAL, LIGHT_POINT, AnimationDesc, CameraDesc, LightDesc, MaterialDesc, MeshDesc, ModelDesc,
                        NodeDesc, PartDesc, PoseDesc, SceneDesc, SkeletonDesc, TrackDesc, WorldDesc, TRACK_POSITION, TRACK_ROTATION, TRACK_SCALE)

F32 = np.float32
_COMP = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4, "MAT4": 16}
_INTERP = {"LINEAR": 0, "STEP": 1, "CUBICSPLINE": 2}


def m4_mult(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Matrix4.Mult (matrix_amd64.s:9-26, matrix_none.go:10-28): float32, left to right, never fused."""
    a = a.astype(F32); b = b.astype(F32)
    out = np.zeros((4, 4), F32)
    for i in range(4):
        for j in range(4):
            out[i, j] = F32(F32(F32(a[i, 0] * b[0, j]) + F32(a[i, 1] * b[1, j])) + F32(a[i, 2] * b[2, j])) + F32(a[i, 3] * b[3, j])
    return out


def quat_to_matrix(q) -> np.ndarray:
    """Quaternion.ToMatrix4 (quaternion.go:135-184)."""
    x, y, z, w = (F32(v) for v in q)
    m1 = np.array([[w, z, -y, x], [-z, w, x, y], [y, -x, w, z], [-x, -y, -z, w]], F32)
    m2 = np.array([[w, z, -y, -x], [-z, w, x, -y], [y, -x, w, -z], [x, y, z, w]], F32)
    return m4_mult(m1, m2)


def to_srgb(c: np.ndarray) -> np.ndarray:
    """Color4.ConvertTosRGB (color.go:196-218) on the first three components: math32.Pow is float32(math.Pow(float64, float64))."""
    c = np.array(c, F32)
    for i in range(3):
        v = c[..., i]
        lo = v <= F32(0.0031308)
        p = np.power(v.astype(np.float64), np.float64(F32(1 / 2.4))).astype(F32)        # math32.Pow(float32, float32)
        hi = (F32(1.055) * p - F32(0.055)).astype(F32)                                  # untyped constants: float32 arithmetic in Go
        c[..., i] = np.where(lo, (v * F32(12.92)).astype(F32), hi)
    return c


class GLB:
    def __init__(self, data: bytes):
        magic, version, length = struct.unpack_from("<III", data, 0)
        if magic != 0x46546C67 or version != 2:
            raise ValueError("not a GLB 2.0 container")
        off = 12
        self.json = None
        self.bin = b""
        while off < length:
            clen, ctype = struct.unpack_from("<II", data, off)
            body = data[off + 8: off + 8 + clen]
            if ctype == 0x4E4F534A:
                self.json = json.loads(body)
            elif ctype == 0x004E4942:
                self.bin = bytes(body)
            off += 8 + clen
        if self.json is None:
            raise ValueError("GLB without a JSON chunk")

    def _view(self, bv_index: int, byte_offset: int, dt, n: int, count: int) -> np.ndarray:
        bv = self.json["bufferViews"][bv_index]
        if bv.get("buffer", 0) != 0:
            raise NotImplementedError("external buffers")
        start = bv.get("byteOffset", 0) + byte_offset
        stride = bv.get("byteStride", 0) or dt.itemsize * n
        if not count:
            return np.zeros((0, n), dt)
        buf = np.frombuffer(self.bin, dtype=np.uint8, count=(count - 1) * stride + dt.itemsize * n, offset=start)
        if stride == dt.itemsize * n:
            return np.array(buf.view(dt).reshape(count, n))
        return np.stack([buf[i * stride: i * stride + dt.itemsize * n].view(dt) for i in range(count)])

    def accessor(self, index: int, raw: bool = False) -> np.ndarray:
        """Accessor `index` as a (count, components) array; sparse accessors are expanded; normalized integers become
        floats unless `raw`."""
        a = self.json["accessors"][index]
        dt = np.dtype(_COMP[a["componentType"]]).newbyteorder("<")
        n = _NCOMP[a["type"]]
        count = a["count"]
        out = self._view(a["bufferView"], a.get("byteOffset", 0), dt, n, count) if "bufferView" in a else np.zeros((count, n), dt)
        if "sparse" in a:
            sp = a["sparse"]
            idt = np.dtype(_COMP[sp["indices"]["componentType"]]).newbyteorder("<")
            ii = self._view(sp["indices"]["bufferView"], sp["indices"].get("byteOffset", 0), idt, 1, sp["count"]).reshape(-1)
            vv = self._view(sp["values"]["bufferView"], sp["values"].get("byteOffset", 0), dt, n, sp["count"])
            out[ii.astype(np.int64)] = vv
        if a.get("normalized") and dt.kind in "ui" and not raw:
            out = out.astype(F32) / F32(np.iinfo(dt).max)
        return out


def _extras(obj) -> dict:
    e = obj.get("extras")
    return e if isinstance(e, dict) else {}


def load_glb(data: bytes, camera: Optional[CameraDesc] = None, playheads: Optional[Dict[str, float]] = None,
             animation_for: Optional[Dict[str, str]] = None, shape_key_weights: Optional[Dict[str, List[float]]] = None,
             name: str = "glb") -> SceneDesc:
    """The default scene of a GLB as a SceneDesc.

    camera:             used when given; else the file's first camera node; else a camera at (0, 0, 10).
    animation_for:      armature (skin) name -> animation name played on it; playheads: armature name -> seconds.  Armatures
                        without an entry stay in their rest pose.
    shape_key_weights:  mesh name -> weights (default: the file's mesh.weights).
    """
    g = GLB(data)
    J = g.json
    playheads = playheads or {}
    animation_for = animation_for or {}
    shape_key_weights = shape_key_weights or {}

    # ---- materials (gltf.go:325-485) ----
    materials: List[MaterialDesc] = []
    for gm in J.get("materials", []):
        m = MaterialDesc(backface_culling=not gm.get("doubleSided", False))
        ex = _extras(gm)
        if "t3dMaterialColor__" in ex:
            m.color = tuple(float(F32(v)) for v in ex["t3dMaterialColor__"])
        if "t3dMaterialShadeless__" in ex:
            m.shadeless = ex["t3dMaterialShadeless__"] > 0.5
        if "t3dLightVolumeShadingMode__" in ex:
            m.light_volume_shading_mode = int(ex["t3dLightVolumeShadingMode__"])
        if "t3dBillboardEnabled__" in ex:
            m.billboard_enabled = ex["t3dBillboardEnabled__"] > 0
        if "t3dRenderOrder__" in ex:
            m.render_order = int(ex["t3dRenderOrder__"])
        if "t3dMaterialLightingMode__" in ex:
            m.lighting_mode = int(ex["t3dMaterialLightingMode__"])
        if "t3dTransparencyMode__" in ex:
            m.transparency_mode = int(ex["t3dTransparencyMode__"])
        if "t3dVisible__" in ex:
            m.visible = ex["t3dVisible__"] > 0
        pbr = gm.get("pbrMetallicRoughness") or {}
        if "baseColorTexture" in pbr:       # the texture's pixels stay with the rasteriser; the geometry stage only needs its size
            m.use_texture = True
            m.tex_w, m.tex_h = 16, 16
        materials.append(m)

    # ---- meshes (gltf.go:547-840) ----
    meshes: List[Optional[MeshDesc]] = []
    for gmesh in J.get("meshes", []):
        pos, nrm, uv, col, wgt, bon, ninf, idx, parts = [], [], [], [], [], [], [], [], []
        kpos: List[List[np.ndarray]] = []
        knrm: List[List[np.ndarray]] = []
        have_col = have_skin = False
        vbase = 0
        for prim in gmesh["primitives"]:
            at = prim["attributes"]
            p = g.accessor(at["POSITION"]).astype(F32)
            n = len(p)
            pos.append(p)
            nrm.append(g.accessor(at["NORMAL"]).astype(F32) if "NORMAL" in at else np.zeros((n, 3), F32))
            if "TEXCOORD_0" in at:
                t = g.accessor(at["TEXCOORD_0"]).astype(F32)
                t[:, 1] = -(t[:, 1] - F32(1))                          # gltf.go:616
                uv.append(t)
            else:
                uv.append(np.zeros((n, 2), F32))
            if "COLOR_0" in at:
                a = J["accessors"][at["COLOR_0"]]
                c = g.accessor(at["COLOR_0"], raw=True)
                # modeler.ReadColor hands out [4]uint8: bytes as they are, ushort >> 8, float * 255 truncated (qmuntal/gltf's
                # conversion restated from its published source; it only shapes this test input), then / 255 and sRGB
                if a["componentType"] == 5126:
                    c8 = np.clip(c.astype(np.float64) * 255.0, 0, 255).astype(np.uint8)
                elif a["componentType"] == 5123:
                    c8 = (c.astype(np.uint32) >> 8).astype(np.uint8)
                else:
                    c8 = c.astype(np.uint8)
                if c8.shape[1] == 3:
                    c8 = np.concatenate([c8, np.full((n, 1), 255, np.uint8)], axis=1)
                col.append(to_srgb((c8.astype(F32) / F32(255)).astype(F32)))
                have_col = True
            else:
                col.append(np.ones((n, 4), F32))
            if "WEIGHTS_0" in at:
                w = g.accessor(at["WEIGHTS_0"]).astype(F32)
                b = g.accessor(at["JOINTS_0"]).astype(np.uint16)
                W = np.zeros((n, 4), F32); B = np.zeros((n, 4), np.uint16); K = np.zeros(n, np.uint8)
                for v in range(n):                                      # gltf.go:692-700: only weights > 0, in order
                    k = 0
                    for i in range(4):
                        if w[v, i] > 0:
                            W[v, k] = w[v, i]; B[v, k] = b[v, i]; k += 1
                    K[v] = k
                wgt.append(W); bon.append(B); ninf.append(K)
                have_skin = True
            else:
                wgt.append(np.zeros((n, 4), F32)); bon.append(np.zeros((n, 4), np.uint16)); ninf.append(np.zeros(n, np.uint8))
            ix = g.accessor(prim["indices"]).astype(np.uint32).reshape(-1, 3) + np.uint32(vbase)
            t0 = sum(len(x) for x in idx)
            idx.append(ix)
            mat = materials[prim["material"]] if "material" in prim else None
            parts.append(PartDesc(vbase, vbase + n, t0, t0 + len(ix) - 1, mat))
            for ti, tgt in enumerate(prim.get("targets", [])):
                while len(kpos) <= ti:
                    kpos.append([]); knrm.append([])
                kpos[ti].append(g.accessor(tgt["POSITION"]).astype(F32) if "POSITION" in tgt else np.zeros((n, 3), F32))
                knrm[ti].append(g.accessor(tgt["NORMAL"]).astype(F32) if "NORMAL" in tgt else np.zeros((n, 3), F32))
            vbase += n
        kw = None
        if kpos:
            kw = np.array(shape_key_weights.get(gmesh.get("name", ""), gmesh.get("weights", [0.0] * len(kpos))), F32)
        meshes.append(MeshDesc(np.concatenate(pos), np.concatenate(nrm), np.concatenate(uv), np.concatenate(idx), parts,
                               colors=np.concatenate(col) if have_col else None, active_color_channel=0 if have_col else -1,
                               weights=np.concatenate(wgt) if have_skin else None, bones=np.concatenate(bon) if have_skin else None,
                               n_influences=np.concatenate(ninf) if have_skin else None,
                               key_positions=np.stack([np.concatenate(k) for k in kpos]) if kpos else None,
                               key_normals=np.stack([np.concatenate(k) for k in knrm]) if kpos else None,
                               key_weights=kw, name=gmesh.get("name", "mesh")))

    # ---- nodes (gltf.go:1039-1411) ----
    gnodes = J.get("nodes", [])
    nodes: List[NodeDesc] = []
    for gn in gnodes:
        if "matrix" in gn:
            raise NotImplementedError("node.matrix")
        nd = NodeDesc(position=tuple(float(F32(v)) for v in gn.get("translation", (0, 0, 0))),
                      scale=tuple(float(F32(v)) for v in gn.get("scale", (1, 1, 1))),
                      rotation=quat_to_matrix(gn.get("rotation", (0, 0, 0, 1))), name=gn.get("name", "node"))
        nodes.append(nd)
    for i, gn in enumerate(gnodes):
        for c in gn.get("children", []):
            nodes[c].parent = nodes[i]

    # ---- animations (gltf.go:844-957) ----
    animations: Dict[str, AnimationDesc] = {}
    for ga in J.get("animations", []):
        channels: Dict[str, dict] = {}
        length = 0.0
        for ch in ga["channels"]:
            smp = ga["samplers"][ch["sampler"]]
            tname = gnodes[ch["target"]["node"]].get("name", "") if "node" in ch["target"] else "root"
            path = ch["target"]["path"]
            kind = {"translation": TRACK_POSITION, "scale": TRACK_SCALE, "rotation": TRACK_ROTATION}.get(path)
            if kind is None:
                continue                                                # morph-weight channels are not tracks (gltf.go has no branch for them)
            times = g.accessor(smp["input"]).astype(F32).reshape(-1)
            vals = g.accessor(smp["output"]).astype(F32)
            interp = _INTERP[smp.get("interpolation", "LINEAR")]
            channels.setdefault(tname, {})[kind] = TrackDesc(times, vals[:len(times)], interp)
            if len(times):
                length = max(length, float(times.max()))
        animations[ga.get("name", "")] = AnimationDesc(ga.get("name", ""), length, channels)

    # ---- skins (gltf.go:1452-1511) ----
    skeletons: Dict[int, tuple] = {}
    for si, gs in enumerate(J.get("skins", [])):
        joints = list(gs["joints"])
        ibm = g.accessor(gs["inverseBindMatrices"]).astype(F32).reshape(-1, 16)     # see the module docstring: already Tetra3D's layout
        parent_of = {}
        for i, gn in enumerate(gnodes):
            for c in gn.get("children", []):
                parent_of[c] = i
        # parents first (the device walks the hierarchy level by level); remap the joint indices of the meshes accordingly
        order = sorted(range(len(joints)), key=lambda k: _depth(joints[k], parent_of, set(joints)))
        remap = np.zeros(len(joints), np.uint16)
        for new, old in enumerate(order):
            remap[old] = new
        jn = [joints[k] for k in order]
        par = np.array([jn.index(parent_of[j]) if parent_of.get(j) in jn else -1 for j in jn], np.int32)
        sk = SkeletonDesc([gnodes[j].get("name", "") for j in jn], par, ibm[order],
                          np.array([nodes[j].position for j in jn], F32), np.array([nodes[j].scale for j in jn], F32),
                          np.array([nodes[j].rotation.reshape(16) for j in jn], F32))
        root = next((nodes[i] for i, gn in enumerate(gnodes) if gn.get("name") == gs.get("name")), None)   # model.SkinRoot, gltf.go:1469-1474
        skeletons[si] = (sk, remap, root, gs.get("name", ""))

    # ---- world (gltf.go:1556-1571) ----
    scene_index = J.get("scene", 0)
    gscene = J["scenes"][scene_index]
    world = WorldDesc()
    sx = _extras(gscene)
    wname = sx.get("t3dCurrentWorld__")
    wprops = (sx.get("t3dWorlds__") or {}).get(wname) if wname else None
    if wprops:
        if "ambient color" in wprops:
            world.ambient_color = tuple(float(v) for v in to_srgb(np.array(list(wprops["ambient color"][:3]) + [1.0], F32))[:3])
        if "ambient energy" in wprops:
            world.ambient_energy = float(F32(wprops["ambient energy"]))
        world.fog_on = wprops.get("fog mode", "OFF") != "OFF"

    # ---- the scene tree, depth first (RenderNodes visits models in this order; scene lights likewise) ----
    models: List[ModelDesc] = []
    lights: List[LightDesc] = []
    cam_node = None
    remapped = set()
    glights = ((J.get("extensions") or {}).get("KHR_lights_punctual") or {}).get("lights", [])

    def visit(i: int):
        nonlocal cam_node
        gn = gnodes[i]
        ex = _extras(gn)
        if ex.get("t3dAddToSceneTree__", 1) == 0:
            return
        if "mesh" in gn:
            mesh = meshes[gn["mesh"]]
            md = ModelDesc(mesh, nodes[i], name=gn.get("name", "model"))
            if "t3dObjectColor__" in ex:
                md.color = tuple(float(v) for v in to_srgb(np.array(ex["t3dObjectColor__"], F32)))
            if "skin" in gn:
                sk, remap, root, sname = skeletons[gn["skin"]]
                if id(mesh) not in remapped:
                    mesh.bones = remap[mesh.bones].astype(np.uint16)
                    remapped.add(id(mesh))
                an = animations.get(animation_for.get(sname, ""))
                md.pose = PoseDesc(sk, an, float(playheads.get(sname, 0.0)), root if root is not None else NodeDesc())
            models.append(md)
        elif "camera" in gn and cam_node is None:
            cam_node = i
        elif "KHR_lights_punctual" in (gn.get("extensions") or {}):
            ld = glights[gn["extensions"]["KHR_lights_punctual"]["light"]]
            colour = to_srgb(np.array(list(ld.get("color", (1, 1, 1))) + [1.0], F32))[:3]
            inten = ld.get("intensity", 1.0)
            if ld["type"] == "directional":
                lights.append(LightDesc(LIGHT_DIRECTIONAL, nodes[i], tuple(float(v) for v in colour), float(F32(inten)), name=gn.get("name", "sun")))
            elif ld["type"] == "point":
                rng = ld.get("range")
                lights.append(LightDesc(LIGHT_POINT, nodes[i], tuple(float(v) for v in colour), float(F32(inten) / F32(80)),
                                        range=float(F32(rng)) if rng is not None and np.isfinite(rng) else 0.0, name=gn.get("name", "point")))
            else:
                lights.append(LightDesc(LIGHT_AMBIENT, nodes[i], tuple(float(v) for v in colour), float(F32(inten) / F32(80)), name=gn.get("name", "ambient")))
        for c in gn.get("children", []):
            visit(c)

    for r in gscene.get("nodes", []):
        visit(r)

    if camera is None:
        if cam_node is not None:
            gc = J["cameras"][gnodes[cam_node]["camera"]]
            camera = CameraDesc(node=nodes[cam_node])
            if "perspective" in gc:
                p = gc["perspective"]
                camera.near = float(F32(p["znear"])); camera.far = float(F32(p.get("zfar", 100.0)))
                camera.field_of_view = float(F32(np.degrees(np.float64(F32(p["yfov"])))))        # math32.ToDegrees
                camera.perspective = True
            else:
                o = gc["orthographic"]
                camera.near = float(F32(o["znear"])); camera.far = float(F32(o["zfar"])); camera.ortho_scale = float(F32(o["xmag"] * 2))
                camera.perspective = False
        else:
            camera = CameraDesc()
    return SceneDesc(models, lights, camera, world, name=name)


def _depth(j: int, parent_of: dict, joint_set: set) -> int:
    d = 0
    while parent_of.get(j) in joint_set:
        j = parent_of[j]; d += 1
    return d
