"""Materialise a `scenedesc.SceneDesc` as host-mirror objects (Scene / Model / Mesh / Camera ...)."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from . import _host as t3d
from .scenedesc import (LIGHT_AMBIENT, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_VOLUME, MaterialDesc, MeshDesc,
                        NodeDesc, SceneDesc)


def _material(md: Optional[MaterialDesc]):
    if md is None:
        return None
    m = t3d.NewMaterial("material")
    m.Color = list(map(float, md.color))
    m.BackfaceCulling = bool(md.backface_culling)
    m.TriangleSortMode = int(md.triangle_sort_mode)
    m.Shadeless = bool(md.shadeless)
    m.LightingMode = int(md.lighting_mode)
    m.LightVolumeShadingMode = int(md.light_volume_shading_mode)
    m.BillboardEnabled = bool(md.billboard_enabled)
    m.BillboardLockX = bool(md.billboard_lock_x)
    m.BillboardLockY = bool(md.billboard_lock_y)
    m.BillboardLockZ = bool(md.billboard_lock_z)
    m.BillboardUpDirection = int(md.billboard_up_direction)
    m.BillboardedDepthMode = int(md.billboarded_depth_mode)
    m.TransparencyMode = int(md.transparency_mode)
    m.BlendSourceOver = bool(md.blend_source_over)
    m.Visible = bool(md.visible)
    m.RenderOrder = int(md.render_order)
    m.UseTexture = bool(md.use_texture)
    m.TextureWidth = int(md.tex_w)
    m.TextureHeight = int(md.tex_h)
    return m


def build_mesh(md: MeshDesc):
    mesh = t3d.NewMesh(md.name)
    mesh.SetVertices(md.positions, md.normals, md.uvs)
    if md.colors is not None and md.active_color_channel >= 0:
        mesh.SetVertexColors(md.colors, int(md.active_color_channel))
    if md.weights is not None:
        mesh.SetSkin(md.weights, md.bones, md.n_influences)
    mesh.SetTriangles(md.indices)
    if md.key_weights is not None:
        mesh.SetShapeKeys(md.key_positions, md.key_normals, md.key_weights)
    for p in md.parts:
        mesh.AddMeshPart(_material(p.material), int(p.vertex_start), int(p.vertex_end), int(p.tri_start), int(p.tri_end))
    mesh.UpdateDimensions()
    if md.auto_subdivide:
        mesh.SetAutoSubdivision(md.tri_centers, np.asarray(md.tri_sub_base, np.uint32), md.tri_sub_level, md.tri_n_levels, md.tri_max_span, md.tri_normal)
    return mesh


def _apply_node(obj, nd: NodeDesc):
    obj.SetLocalPosition(*map(float, nd.position))
    obj.SetLocalScale(*map(float, nd.scale))
    obj.SetLocalRotation(np.asarray(nd.rotation, dtype=np.float32))


def build_animation(ad):
    """An Animation of the host mirror from an AnimationDesc (channels keyed by bone name, animation.go:204-222)."""
    anim = t3d.NewAnimation(ad.name)
    anim.SetLength(float(ad.length))
    anim.SetLoopMode(int(ad.loop_mode))
    for ch_name, tracks in ad.channels.items():
        ch = anim.AddChannel(ch_name)
        for ttype, td in tracks.items():
            tr = ch.AddTrack(int(ttype))
            tr.Interpolation = int(td.interpolation)
            for t, v in zip(td.times, td.values):
                tr.AddKeyframe(float(t), [float(x) for x in (v if int(ttype) == 2 else v[:3])])
    return anim


def build_scene(desc: SceneDesc, device=None) -> Tuple["t3d.Scene", "t3d.Camera", Dict[int, object]]:
    """Returns (scene, camera, objects) where objects maps id(desc object) -> host object.

    Parentless nodes become children of scene.Root in description order (models first, then the camera,
    then lights), so RenderNodes' depth-first traversal (camera.go:1200-1210) sees models and lights in
    the order the description lists them.  A NodeDesc used as `parent` becomes a plain Node.
    """
    scene = t3d.NewScene(desc.name)
    if len(desc.autosubdivision_levels):
        scene.SetAutoSubdivisionLevels(np.asarray(list(desc.autosubdivision_levels), np.float32).reshape(-1))
    objs: Dict[int, object] = {}
    meshes: Dict[int, object] = {}
    group_nodes: Dict[int, object] = {}

    def attach(obj, nd: NodeDesc):
        _apply_node(obj, nd)
        if nd.parent is None:
            scene.Root.AddChildren(obj)
        else:
            pid = id(nd.parent)
            if pid not in group_nodes:
                g = t3d.Node(nd.parent.name)
                attach(g, nd.parent)
                group_nodes[pid] = g
            group_nodes[pid].AddChildren(obj)

    w = desc.world
    if w.present:
        world = scene.World
        world.LightingOn = bool(w.lighting_on)
        world.FogOn = bool(w.fog_on)
        world.FogMode = int(w.fog_mode)
        world.FogColor = list(map(float, w.fog_color))
        world.FogRange = list(map(float, w.fog_range))
        amb = t3d.NewAmbientLight("ambient light", *map(float, w.ambient_color), float(w.ambient_energy))
        amb.SetVisible(bool(w.ambient_visible))
        world.AmbientLight = amb
    else:
        scene.World = None

    lights: Dict[int, object] = {}

    def make_light(ld):
        if id(ld) in lights:
            return lights[id(ld)]
        c = list(map(float, ld.color))
        if ld.type == LIGHT_AMBIENT:
            l = t3d.NewAmbientLight(ld.name, c[0], c[1], c[2], float(ld.energy))
        elif ld.type == LIGHT_POINT:
            l = t3d.NewPointLight(ld.name, c[0], c[1], c[2], float(ld.energy))
            l.SetRange(float(ld.range))
        elif ld.type == LIGHT_DIRECTIONAL:
            l = t3d.NewDirectionalLight(ld.name, c[0], c[1], c[2], float(ld.energy))
        elif ld.type == LIGHT_VOLUME:
            l = t3d.LightVolume(ld.name)
            l.SetColor([c[0], c[1], c[2], 1.0])
            l.SetEnergy(float(ld.energy))
            v = ld.volume
            l.SetGrid(np.asarray(v.dim_min, np.float32), np.asarray(v.dim_max, np.float32), np.asarray(v.cell_size, np.float32),
                      [int(x) for x in v.cell_count], np.asarray(v.colors, np.float32))
            l.LightVolumeSetOffset(*map(float, v.offset))
            l.LightVolumeSetLightBias(*map(float, v.light_bias))
        else:
            raise ValueError("unknown light type %r" % (ld.type,))
        l.SetVisible(bool(ld.visible))
        if ld.in_scene:
            attach(l, ld.node)
        else:
            _apply_node(l, ld.node)
        lights[id(ld)] = l
        objs[id(ld)] = l
        if ld.type == LIGHT_VOLUME and ld.volume.cell_lights:
            for (x, y, z) in sorted(ld.volume.cell_lights):
                l.CellSetUseLights(int(x), int(y), int(z), *[make_light(nl) for nl in ld.volume.cell_lights[(x, y, z)]])
        return l

    groups: Dict[int, object] = {}
    skeletons: Dict[int, object] = {}
    animations: Dict[int, object] = {}

    def build_skeleton(sd):
        # a template armature: bone Nodes with the rest transforms, parented as the joint list says
        joints = []
        for b, name in enumerate(sd.names):
            n = t3d.Node(name)
            n.SetLocalPosition(*map(float, sd.rest_position[b]))
            n.SetLocalScale(*map(float, sd.rest_scale[b]))
            n.SetLocalRotation(np.asarray(sd.rest_rotation[b], np.float32).reshape(16))
            joints.append(n)
        for b, p in enumerate(sd.parent):
            if p >= 0:
                joints[int(p)].AddChildren(joints[b])
        return t3d.NewSkeletonFromJoints(joints, np.asarray(sd.inverse_bind, np.float32).reshape(-1))

    models = []
    for md in desc.models:
        if id(md.mesh) not in meshes:
            meshes[id(md.mesh)] = build_mesh(md.mesh)
            objs[id(md.mesh)] = meshes[id(md.mesh)]
        model = t3d.NewModel(md.name, meshes[id(md.mesh)])
        model.Color = list(map(float, md.color))
        model.Shadeless = bool(md.shadeless)
        model.FrustumCulling = bool(md.frustum_culling)
        model.SetVisible(bool(md.visible))
        if md.bone_influences is not None:
            model.SetBoneInfluences(np.asarray(md.bone_influences, np.float32).reshape(-1))
        if md.pose is not None:
            pose = md.pose
            if id(pose.skeleton) not in skeletons:
                skeletons[id(pose.skeleton)] = build_skeleton(pose.skeleton)
            armature = t3d.Node(md.name + " armature")
            attach(armature, pose.root)
            player = None
            if pose.animation is not None:
                if id(pose.animation) not in animations:
                    animations[id(pose.animation)] = build_animation(pose.animation)
                player = t3d.NewAnimationPlayer(armature)
                if pose.blend_ms is not None:
                    # drive the player the way a game would: play the previous animation with BlendTime set, Play() the next
                    # one (the blend starts: animation.go:548-554), let blend_ms milliseconds of a stepped clock pass
                    clock = [0]
                    player.SetClock(lambda clock=clock: clock[0])
                    player.BlendTime = float(pose.blend_time)
                    if pose.blend_from is not None:
                        if id(pose.blend_from) not in animations:
                            animations[id(pose.blend_from)] = build_animation(pose.blend_from)
                        player.Play(animations[id(pose.blend_from)])
                        player.SetPlayhead(float(pose.from_playhead))
                    player.Play(animations[id(pose.animation)])
                    clock[0] = int(pose.blend_ms)
                    player.SetPlayhead(float(pose.playhead))
                else:
                    player.Play(animations[id(pose.animation)])
                    player.SetPlayhead(float(pose.playhead))
                objs[id(pose)] = player
            model.SetSkeleton(skeletons[id(pose.skeleton)], armature, player)
        attach(model, md.node)
        objs[id(md)] = model
        models.append((md, model))

    cam_d = desc.camera
    cam = t3d.NewCamera("camera", int(cam_d.width), int(cam_d.height))
    cam.SetFieldOfView(float(cam_d.field_of_view))
    cam.SetNear(float(cam_d.near))
    cam.SetFar(float(cam_d.far))
    cam.SetPerspective(bool(cam_d.perspective))
    cam.SetOrthoScale(float(cam_d.ortho_scale))
    cam.DepthMargin = float(cam_d.depth_margin)
    cam.VertexSnapping = float(cam_d.vertex_snapping)
    cam.RenderDepth = bool(cam_d.render_depth)
    cam.RenderNormals = bool(cam_d.render_normals)
    cam.PerspectiveCorrectedTextureMapping = bool(cam_d.perspective_corrected_texture_mapping)
    cam.MaxLightCount = int(cam_d.max_light_count)
    attach(cam, cam_d.node)
    if cam_d.max_light_focus is not None:
        focus = t3d.Node("max light count focus")
        attach(focus, cam_d.max_light_focus)
        cam.MaxLightCountFocusPoint = focus
    if device is not None:
        cam.Device = device

    for ld in desc.lights:
        make_light(ld)

    for md, model in models:
        if md.light_group is not None:
            if id(md.light_group) not in groups:
                g = t3d.NewLightGroup(*[make_light(l) for l in md.light_group.lights])
                g.Active = bool(md.light_group.active)
                groups[id(md.light_group)] = g
            model.LightGroup = groups[id(md.light_group)]
        if md.batch_owner is not None:
            objs[id(md.batch_owner)].DynamicBatchAdd(int(md.batch_owner_part), model)

    return scene, cam, objs
