#!/usr/bin/env python
"""Turns the reference's own example assets into committed parity-test inputs.

    python tests/golden/make_asset_fixtures.py        (run where /root/reference exists)

Reads /root/reference/examples/animation/assets/animations.glb (a 6-joint skinned mesh with its 18-channel ArmatureAction, a
cube with two morph targets, vertex-coloured meshes, a point light) and /root/reference/examples/lighting/assets/lighting.glb
(3 point lights + a sun, a 2916-vertex ground plane, double-sided materials, a 2-joint skinned cube with its animation, a model
without material, the file's own camera) through tetra3d_b200.glb -- which follows gltf.go's load rules -- and stores the
resulting SceneDescs (the INPUTS of the geometry stage: vertex arrays, skin weights as the loader keeps them, morph deltas,
keyframes, node transforms, lights) as gzip-pickled fixtures under tests/golden/assets/, next to a JSON of oracle output digests.
/root/reference does not travel to the GPU box; these fixtures do.  tests/test_assets.py checks the oracle and the CUDA path
against each other on them, and the oracle against the digests.
"""
import gzip
import hashlib
import json
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np

from tetra3d_b200 import glb
from tetra3d_b200.scenedesc import CameraDesc, NodeDesc

REF = "/root/reference/examples"
OUT = os.path.join(HERE, "assets")


def look_from(pos, pitch_deg=0.0, yaw_deg=0.0, **kw):
    """A camera at `pos`, yawed about Y then pitched about its X (Matrix4.Rotated semantics are not needed: any rotation is data)."""
    p, y = np.radians(pitch_deg), np.radians(yaw_deg)
    ry = np.array([[np.cos(y), 0, -np.sin(y), 0], [0, 1, 0, 0], [np.sin(y), 0, np.cos(y), 0], [0, 0, 0, 1]], np.float32)
    rx = np.array([[1, 0, 0, 0], [0, np.cos(p), np.sin(p), 0], [0, -np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)
    return CameraDesc(node=NodeDesc(position=tuple(float(v) for v in pos), rotation=(rx @ ry).astype(np.float32)), **kw)


def scenes():
    anim = open(os.path.join(REF, "animation/assets/animations.glb"), "rb").read()
    light = open(os.path.join(REF, "lighting/assets/lighting.glb"), "rb").read()
    out = {}
    for i, ph in enumerate((0.0, 0.7, 1.9)):
        out["animations_t%d" % i] = glb.load_glb(anim, camera=look_from((5.0, 3.0, 16.0), pitch_deg=-8.0, far=200.0),
                                                 animation_for={"Armature": "ArmatureAction"}, playheads={"Armature": ph},
                                                 shape_key_weights={"DiamondCube": [0.35 * i, 1.0 - 0.4 * i]}, name="animations_t%d" % i)
    out["animations_rest"] = glb.load_glb(anim, camera=look_from((5.0, 3.0, 16.0), pitch_deg=-8.0, far=200.0), name="animations_rest")
    out["lighting_filecam"] = glb.load_glb(light, animation_for={"Armature": "ArmatureAction"}, playheads={"Armature": 0.4}, name="lighting_filecam")
    out["lighting_wide"] = glb.load_glb(light, camera=look_from((2.0, 14.0, 34.0), pitch_deg=-22.0, far=200.0, field_of_view=55.0),
                                        animation_for={"Armature": "ArmatureAction"}, playheads={"Armature": 1.1}, name="lighting_wide")
    nodepth = glb.load_glb(light, camera=look_from((2.0, 14.0, 34.0), pitch_deg=-22.0, far=200.0, field_of_view=55.0, render_depth=False),
                           animation_for={"Armature": "ArmatureAction"}, playheads={"Armature": 0.0}, name="lighting_nodepth")
    out["lighting_nodepth"] = nodepth          # RenderDepth off: the solids are sorted by distance (camera.go:1460-1469)
    out.update(example_scenes())
    return out


# Every other GLB under /root/reference/examples, each once: with the file's own camera when it sees something, else framed from
# the scene's bounds.  What they add over the two files above: vertex colours and several materials per mesh (shapes, baking,
# blendmodes), transparent / alpha-clip materials with render orders (transparency, blendmodes), orthographic file cameras
# (orthographic), 50-model scenes (animatedTextures, autobatch), a 2-D skinned sprite (skeletal2d), world / fog settings exported
# from Blender (worlds), sun + point + cube-light mixes (baking, rays, sectors).
EXAMPLES = ["3dSprite/assets/scene.glb", "animatedTextures/assets/animatedTextures.glb", "animation_queueing/assets/anim_queueing.glb",
            "autoBatch/autobatch.glb", "baking/assets/baking.glb", "blendmodes/blendmodes.glb", "cameraTest/cameraTest.glb",
            "collections/assets/collections.glb", "collections/assets/cone.glb", "colliders/colliders.glb", "engine/engine.glb",
            "grids/grids.glb", "grids2/grids.glb", "lightgroups/lighting.glb", "lightvolume/lightvolume.glb", "offscreen/offscreen.glb",
            "orthographic/orthographic.glb", "particles/particles.glb", "paths/paths.glb", "properties/properties.glb", "rays/rays.glb",
            "sectors/sectors.glb", "shapes/shapes.glb", "skeletal2d/skeletal2d.glb", "text/text.glb", "transparency/transparency.glb",
            "vertexSelection/vertexSelection.glb", "worlds/worlds.glb"]


def frame_camera(sc):
    """A camera that sees the whole scene: above and in front of the bounds of every model's world-space vertices."""
    from oracle import binding as O
    lo, hi = np.full(3, np.inf), np.full(3, -np.inf)
    for m in sc.models:
        w = O.node_world_transform(m.node).reshape(4, 4).astype(np.float64)
        p = np.concatenate([m.mesh.positions.astype(np.float64), np.ones((len(m.mesh.positions), 1))], axis=1) @ w
        lo, hi = np.minimum(lo, p[:, :3].min(axis=0)), np.maximum(hi, p[:, :3].max(axis=0))
    c, r = (lo + hi) / 2, max(float(np.linalg.norm(hi - lo)) / 2, 1.0)
    pos = np.round(c + np.array([0.15 * r, 0.6 * r, 1.5 * r]), 2)
    d = c - pos
    pitch = np.degrees(np.arctan2(d[1], np.hypot(d[0], d[2])))
    yaw = np.degrees(np.arctan2(-d[0], -d[2]))
    return look_from(pos, pitch_deg=round(float(pitch), 1), yaw_deg=round(float(-yaw), 1), far=float(np.ceil(4 * r + 10)))


def example_scenes():
    from oracle import binding as O
    out = {}
    for rel in EXAMPLES:
        name = "ex_" + rel.split("/")[0].lower() + ("_" + os.path.basename(rel)[:-4].lower() if rel.startswith("collections") else "")
        data = open(os.path.join(REF, rel), "rb").read()
        sc = glb.load_glb(data, name=name)
        if O.render_scene(sc).n_vertices < 3 * max(sc.total_triangles() // 20, 1):       # the file's camera (or the default) sees next to nothing
            sc = glb.load_glb(data, camera=frame_camera(sc), name=name)
        out[name] = sc
    return out


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def summarize(r):
    return {"n_vertices": int(r.n_vertices), "render_counts": [int(x["vertex_count"]) for x in r.renders],
            "tri_ids_sha256": digest(r.tri_ids), "color_sha256": digest(r.color_list), "depth_sha256": digest(r.depth_list)}


def main():
    from oracle import binding as O
    os.makedirs(OUT, exist_ok=True)
    digests = {}
    for name, sc in scenes().items():
        with gzip.GzipFile(os.path.join(OUT, name + ".pkl.gz"), "wb", mtime=0) as f:
            pickle.dump(sc, f, protocol=4)
        digests[name] = summarize(O.render_scene(sc))
        print(name, sc.total_triangles(), "triangles ->", digests[name]["n_vertices"], "vertices emitted")
    json.dump(digests, open(os.path.join(OUT, "digests.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
