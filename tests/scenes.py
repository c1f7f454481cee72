"""Small seeded scenes for parity tests (shared by the CPU-oracle tests and the GPU tests)."""
from __future__ import annotations

import math

import numpy as np

from tetra3d_b200 import workloads as W
from tetra3d_b200.scenedesc import (LIGHT_AMBIENT, LIGHT_DIRECTIONAL, LIGHT_POINT, LIGHT_VOLUME, CameraDesc, LightDesc,
                                    LightGroupDesc, MaterialDesc, MeshDesc, ModelDesc, NodeDesc, PartDesc, SceneDesc, VolumeDesc,
                                    WorldDesc)

F32 = np.float32


def single_triangle(z=0.0, flip=False, backface=True) -> MeshDesc:
    pos = np.array([[-1, -1, z], [1, -1, z], [0, 1, z]], F32)
    idx = np.array([[0, 2, 1]] if flip else [[0, 1, 2]], np.uint32)
    nrm = np.tile(np.array([[0, 0, 1]], F32), (3, 1))
    uv = np.array([[0, 0], [1, 0], [0.5, 1]], F32)
    return MeshDesc(pos, nrm, uv, idx, [PartDesc(0, 3, 0, 0, MaterialDesc(backface_culling=backface))])


def random_soup(n_tris=500, seed=0, extent=6.0, material=None, colors=False, n_parts=1) -> MeshDesc:
    """Random, partly shared-vertex triangle soup around the origin."""
    rng = np.random.default_rng(seed)
    n_verts = max(8, n_tris // 2 * 3 // 2)
    n_verts = (n_verts + n_parts - 1) // n_parts * n_parts
    pos = rng.uniform(-extent, extent, (n_verts, 3)).astype(F32)
    nrm = rng.normal(size=(n_verts, 3)).astype(F32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    uv = rng.uniform(0, 1, (n_verts, 2)).astype(F32)
    per_v = n_verts // n_parts
    per_t = n_tris // n_parts
    idx_l, parts = [], []
    for p in range(n_parts):
        lo = p * per_v
        tri = lo + rng.integers(0, per_v, (per_t, 3))
        # no repeated index inside a triangle (the reference double-lights such triangles)
        bad = (tri[:, 0] == tri[:, 1]) | (tri[:, 1] == tri[:, 2]) | (tri[:, 0] == tri[:, 2])
        tri[bad] = lo + np.array([0, 1, 2])
        idx_l.append(tri)
        mat = material if material is not None else MaterialDesc(backface_culling=bool(p % 2 == 0), triangle_sort_mode=p % 3)
        parts.append(PartDesc(lo, lo + per_v, p * per_t, (p + 1) * per_t - 1, mat))
    idx = np.concatenate(idx_l).astype(np.uint32)
    col = rng.uniform(0, 1, (n_verts, 4)).astype(F32) if colors else None
    return MeshDesc(pos, nrm.astype(F32), uv, idx, parts, colors=col, active_color_channel=0 if colors else -1)


def random_rotation(rng) -> np.ndarray:
    return W.rotation_matrix(rng.normal(size=3), float(rng.uniform(0, 2 * np.pi)))


def lights_mixed(rng, ranged=False) -> list:
    ls = [LightDesc(LIGHT_POINT, NodeDesc(position=tuple(float(x) for x in rng.uniform(-8, 8, 3))), color=tuple(float(x) for x in rng.uniform(0.2, 1, 3)),
                    energy=float(rng.uniform(1, 6)), range=float(rng.uniform(4, 12)) if ranged and i % 2 == 0 else 0.0, name="p%d" % i) for i in range(3)]
    ls.append(LightDesc(LIGHT_DIRECTIONAL, NodeDesc(rotation=random_rotation(rng)), color=(1.0, 0.9, 0.8), energy=0.8, name="sun"))
    ls.append(LightDesc(LIGHT_AMBIENT, NodeDesc(), color=(0.2, 0.3, 0.4), energy=0.5, name="amb"))
    return ls


def soup_scene(seed=0, n_models=4, n_tris=600, lit=True, ranged=False, cam: CameraDesc = None, world: WorldDesc = None, colors=False,
               n_parts=3, material=None, scale_range=(0.5, 1.5), parent=False) -> SceneDesc:
    rng = np.random.default_rng(seed)
    mesh = random_soup(n_tris, seed + 100, colors=colors, n_parts=n_parts, material=material)
    models = []
    group = NodeDesc(position=(1.0, -0.5, 0.25), rotation=random_rotation(rng), scale=(1.0, 1.0, 1.0), name="group") if parent else None
    for i in range(n_models):
        s = rng.uniform(*scale_range, 3)
        nd = NodeDesc(position=tuple(float(x) for x in rng.uniform(-5, 5, 3)), scale=tuple(float(x) for x in s), rotation=random_rotation(rng),
                      parent=group if (parent and i % 2 == 0) else None)
        models.append(ModelDesc(mesh, nd, color=tuple(float(x) for x in rng.uniform(0.5, 1, 4)) if i % 2 else (1, 1, 1, 1),
                                frustum_culling=bool(i % 3), name="m%d" % i))
    if parent:   # keep description order == depth-first order: children of the group node first
        models = [m for m in models if m.node.parent is not None] + [m for m in models if m.node.parent is None]
    lights = lights_mixed(rng, ranged) if lit else []
    cam = cam if cam is not None else CameraDesc(node=NodeDesc(position=(0.5, 1.0, 14.0), rotation=W.rotation_matrix((0, 1, 0), 0.05)))
    return SceneDesc(models, lights, cam, world if world is not None else WorldDesc(ambient_energy=0.1), name="soup%d" % seed)


def tile_boundary_scene(seed=0, sizes=(1, 2, 1023, 1024, 1025, 7, 2048, 2049, 300, 300, 300, 300, 3000, 5, 1, 8200, 640)) -> SceneDesc:
    """One model per entry of `sizes` (triangles in its single MeshPart): draws on both sides of the 1024-slot
    triangle tile of the device pipeline, packed several to a tile, filling a tile exactly, and spanning several."""
    rng = np.random.default_rng(seed)
    models = []
    for k, n in enumerate(sizes):
        mat = MaterialDesc(backface_culling=bool(k % 2), triangle_sort_mode=k % 3)
        mesh = random_soup(n, seed + 10 * k, material=mat, n_parts=1, colors=bool(k % 4 == 0))
        nd = NodeDesc(position=tuple(float(x) for x in rng.uniform(-5, 5, 3)), scale=tuple(float(x) for x in rng.uniform(0.5, 1.5, 3)),
                      rotation=random_rotation(rng))
        models.append(ModelDesc(mesh, nd, frustum_culling=False, name="t%d_%d" % (k, n)))
    cam = CameraDesc(node=NodeDesc(position=(0.5, 1.0, 14.0), rotation=W.rotation_matrix((0, 1, 0), 0.05)))
    return SceneDesc(models, lights_mixed(rng, ranged=True), cam, WorldDesc(ambient_energy=0.1), name="tiles%d" % seed)


def fuzz_scene(seed: int) -> SceneDesc:
    """A random frame: random camera / world switches, and a random mix of static soups of random sizes (packed,
    medium and large draws), skinned characters, billboards, a dynamic batch and transparent parts."""
    rng = np.random.default_rng(1000 + seed)
    models = []
    n_static = int(rng.integers(1, 7))
    for k in range(n_static):
        n_tris = int(rng.choice([1, 3, 40, 700, 1024, 1500, 2300, 5000, 9000]))
        n_parts = int(rng.choice([1, 1, 2, 3])) if n_tris >= 40 else 1
        mat = None if rng.random() < 0.6 else MaterialDesc(backface_culling=bool(rng.integers(0, 2)), triangle_sort_mode=int(rng.integers(0, 3)),
                                                           lighting_mode=int(rng.integers(0, 3)), shadeless=bool(rng.random() < 0.2),
                                                           transparency_mode=int(rng.choice([0, 0, 3])), use_texture=bool(rng.integers(0, 2)), tex_w=32, tex_h=16)
        mesh = random_soup(n_tris, seed * 100 + k, material=mat, n_parts=n_parts, colors=bool(rng.integers(0, 2)), extent=float(rng.uniform(2, 8)))
        for inst in range(int(rng.integers(1, 4))):
            nd = NodeDesc(position=tuple(float(x) for x in rng.uniform(-6, 6, 3)), scale=tuple(float(x) for x in rng.uniform(0.4, 1.6, 3)),
                          rotation=random_rotation(rng))
            models.append(ModelDesc(mesh, nd, color=tuple(float(x) for x in rng.uniform(0.3, 1, 4)), frustum_culling=bool(rng.integers(0, 2)),
                                    shadeless=bool(rng.random() < 0.1), name="f%d_%d" % (k, inst)))
    if rng.random() < 0.5:
        sk = skinned_scene(seed, n=int(rng.integers(1, 5)))
        models.extend(sk.models)
    if rng.random() < 0.5:
        bb = billboard_scene(seed, locks=tuple(bool(x) for x in rng.integers(0, 2, 3)), camera_up=bool(rng.integers(0, 2)), unbillboarded=bool(rng.integers(0, 2)))
        models.extend(bb.models[: int(rng.integers(1, 12))])
    if rng.random() < 0.4:
        bs = batch_scene(seed, n_children=int(rng.integers(1, 60)), transparent=bool(rng.integers(0, 2)))
        models.extend(bs.models)
    lights = lights_mixed(rng, ranged=bool(rng.integers(0, 2))) if rng.random() < 0.85 else []
    persp = bool(rng.random() < 0.8)
    cam = CameraDesc(node=NodeDesc(position=(float(rng.uniform(-2, 2)), float(rng.uniform(-1, 3)), float(rng.uniform(6, 16))),
                                   rotation=W.rotation_matrix((0, 1, 0), float(rng.uniform(-0.2, 0.2)))),
                     perspective=persp, ortho_scale=float(rng.uniform(8, 24)), far=float(rng.choice([20.0, 100.0])),
                     vertex_snapping=float(rng.choice([0.0, 0.0, 8.0])), render_depth=bool(rng.random() < 0.7),
                     render_normals=bool(rng.random() < 0.3), perspective_corrected_texture_mapping=bool(rng.random() < 0.3))
    world = WorldDesc(present=bool(rng.random() < 0.9), lighting_on=bool(rng.random() < 0.9), fog_on=bool(rng.integers(0, 2)),
                      fog_mode=int(rng.integers(0, 2)), fog_color=(0.3, 0.5, 0.7, 1.0), fog_range=(0.1, 0.9), ambient_energy=float(rng.uniform(0, 0.4)))
    return SceneDesc(models, lights, cam, world, name="fuzz%d" % seed)


def skinned_scene(seed=0, n=6, rings=9, sides=10) -> SceneDesc:
    rng = np.random.default_rng(seed)
    mesh, bone_y = W.tube_character_mesh(rings=rings, sides=sides, n_bones=5)
    # make some weights exactly 0 / 1 and vary the influence count to reach every branch of skinVertex
    w = mesh.weights.copy(); n_inf = mesh.n_influences.copy()
    w[::7, 0] = 1.0; w[::7, 1] = 0.0
    w[3::11, 0] = 0.0; w[3::11, 1] = 1.0
    n_inf[5::13] = 1
    extra = rng.uniform(0, 0.2, len(w)).astype(F32)
    w[1::5, 2] = extra[1::5]; n_inf[1::5] = 3
    mesh.bones[1::5, 2] = (mesh.bones[1::5, 0] + 1) % 5
    mesh.weights = w; mesh.n_influences = n_inf
    models = []
    ls = [LightDesc(LIGHT_POINT, NodeDesc(position=(2.0, 3.0, 4.0)), color=(1, 0.8, 0.6), energy=4.0, range=30.0, in_scene=False),
          LightDesc(LIGHT_DIRECTIONAL, NodeDesc(rotation=W.rotation_matrix((1, 0, 0), 0.7)), color=(0.5, 0.6, 1.0), energy=1.0, in_scene=False)]
    group = LightGroupDesc(ls)
    for i in range(n):
        bi = W.chain_bone_influences(bone_y, (2.0 * (i - n / 2), -1.0, 0.0), float(rng.uniform(0, 6)), 0.3)
        models.append(ModelDesc(mesh, NodeDesc(), bone_influences=bi, light_group=group if i % 2 == 0 else None, frustum_culling=False, name="char%d" % i))
    lights = [LightDesc(LIGHT_POINT, NodeDesc(position=(-3.0, 2.0, 5.0)), color=(0.9, 0.9, 0.5), energy=3.0, name="scene_point")]
    return SceneDesc(models, lights, CameraDesc(node=NodeDesc(position=(0.0, 1.0, 9.0))), WorldDesc(ambient_energy=0.15), name="skinned")


def shape_key_scene(seed=0) -> SceneDesc:
    rng = np.random.default_rng(seed)
    mesh = random_soup(200, seed, n_parts=1, material=MaterialDesc(backface_culling=False))
    k = 2
    mesh.key_positions = rng.normal(0, 0.3, (k, mesh.n_verts, 3)).astype(F32)
    mesh.key_normals = rng.normal(0, 0.1, (k, mesh.n_verts, 3)).astype(F32)
    mesh.key_weights = np.array([0.35, 0.8], F32)
    return SceneDesc([ModelDesc(mesh, NodeDesc(position=(0.0, 0.0, -2.0)))], lights_mixed(rng), CameraDesc(), WorldDesc(), name="shapekeys")


def billboard_scene(seed=0, locks=(False, False, False), camera_up=False, unbillboarded=False) -> SceneDesc:
    rng = np.random.default_rng(seed)
    mat = MaterialDesc(billboard_enabled=True, billboard_lock_x=locks[0], billboard_lock_y=locks[1], billboard_lock_z=locks[2],
                       billboard_up_direction=1 if camera_up else 0, billboarded_depth_mode=1 if unbillboarded else 0, backface_culling=False,
                       shadeless=True)
    colors = rng.uniform(0.1, 1, (4, 4)).astype(F32)
    mesh = W.plane_mesh(2, 2, material=mat, colors=colors)
    models = [ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-6, 6, 3)), scale=(float(rng.uniform(0.3, 1.5)),) * 3,
                                       rotation=random_rotation(rng)), name="bb%d" % i) for i in range(12)]
    cam = CameraDesc(node=NodeDesc(position=(1.0, 2.0, 12.0), rotation=W.rotation_matrix((1, 0, 0), -0.15)))
    return SceneDesc(models, [], cam, WorldDesc(), name="billboards")


def batch_scene(seed=0, n_children=40, transparent=True) -> SceneDesc:
    """A dynamic batch owner with particle-like children (camera.go:1398-1453, 2075-2100)."""
    rng = np.random.default_rng(seed)
    child_mat = MaterialDesc(billboard_enabled=True, transparency_mode=3 if transparent else 0, shadeless=True, backface_culling=False)
    colors = rng.uniform(0.1, 1, (4, 4)).astype(F32)
    child_mesh = W.plane_mesh(1, 1, material=child_mat, colors=colors)
    owner_mesh = W.plane_mesh(1, 1, material=MaterialDesc(transparency_mode=3 if transparent else 0, shadeless=True))
    owner = ModelDesc(owner_mesh, NodeDesc(position=(0.0, 0.0, -3.0)), frustum_culling=False, name="owner")
    models = [owner]
    for i in range(n_children):
        models.append(ModelDesc(child_mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-5, 5, 3)), scale=(float(rng.uniform(0.2, 1)),) * 3),
                                color=(1.0, float(rng.uniform(0, 1)), 0.5, 0.7), batch_owner=owner, frustum_culling=bool(i % 2), name="c%d" % i))
    solid = ModelDesc(W.cube_mesh(1, 1, 1), NodeDesc(position=(2.0, 0.0, 0.0)), name="solid")
    models.append(solid)
    return SceneDesc(models, [], CameraDesc(), WorldDesc(), name="batch")


def batch_frustum_scene(seed=0, n_children=60, spread=40.0, perspective=True, parent=False, billboard=True) -> SceneDesc:
    """A dynamic batch whose children are instances of one another (one mesh, one material, one parent, no rotation) with
    FrustumCulling on, spread far wider than the view: most of them fail Camera.SphereInFrustum (camera.go:2088-2093), which the
    device evaluates for such batches (t3d_submit_instances with bounds).  Non-uniform scales, mirrored ones included."""
    rng = np.random.default_rng(seed)
    child_mat = MaterialDesc(billboard_enabled=billboard, transparency_mode=3, shadeless=True, backface_culling=False)
    child_mesh = W.cube_mesh(1.0, 0.6, 0.8, material=child_mat) if not billboard else W.plane_mesh(1, 1, material=child_mat, colors=rng.uniform(0.1, 1, (4, 4)).astype(F32))
    owner_mesh = W.plane_mesh(1, 1, material=MaterialDesc(transparency_mode=3, shadeless=True))
    owner = ModelDesc(owner_mesh, NodeDesc(position=(0.0, 0.0, -3.0)), frustum_culling=False, name="owner")
    group = NodeDesc(position=(1.5, -0.5, 0.5), scale=(1.0, 1.25, 1.0), rotation=W.rotation_matrix((0, 1, 0), 0.4), name="group") if parent else None
    models = [owner]
    for i in range(n_children):
        s = rng.uniform(0.3, 2.0, 3)
        if i % 5 == 0:
            s[0] = -s[0]
        models.append(ModelDesc(child_mesh, NodeDesc(position=(float(rng.uniform(-spread, spread)), float(rng.uniform(-spread / 2, spread / 2)), float(rng.uniform(-60, 20))),
                                                     scale=tuple(float(x) for x in s), parent=group),
                                color=(1.0, float(rng.uniform(0, 1)), 0.5, 0.7), batch_owner=owner, frustum_culling=True, name="c%d" % i))
    cam = CameraDesc(perspective=perspective, ortho_scale=30.0) if not perspective else CameraDesc()
    return SceneDesc(models, [], cam, WorldDesc(), name="batchfrustum")


def volume_scene(seed=0, mode=0) -> SceneDesc:
    rng = np.random.default_rng(seed)
    cx, cy, cz = 5, 4, 6
    vol = VolumeDesc(dim_min=(-6, -5, -6), dim_max=(6, 5, 6), cell_size=(2.4, 2.5, 2.0), cell_count=(cx, cy, cz),
                     colors=rng.uniform(0, 1, (cx * cy * cz, 3)).astype(F32), offset=(0.1, -0.2, 0.05), light_bias=(0.0, 0.1, 0.0))
    mat = MaterialDesc(light_volume_shading_mode=mode, backface_culling=False)
    mesh = random_soup(300, seed, extent=3.0, n_parts=1, material=mat)
    models = [ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-2, 2, 3)), rotation=random_rotation(rng)), name="v%d" % i) for i in range(3)]
    lights = [LightDesc(LIGHT_VOLUME, NodeDesc(), color=(1.0, 0.9, 0.7), energy=1.5, volume=vol, name="volume"),
              LightDesc(LIGHT_POINT, NodeDesc(position=(0.0, 4.0, 3.0)), energy=2.0)]
    return SceneDesc(models, lights, CameraDesc(), WorldDesc(), name="volume")


def volume_cell_lights_scene(seed=0, mode=0, skinned=False, n_extra=0) -> SceneDesc:
    """LightVolumeCell.useLights (light.go:652-654, 1508-1514): a third of the cells carry their own lights -- ranged and
    unranged point lights, a sun, an ambient light, one of them also in the scene list, some cells with several -- that
    light a vertex before the cell colour does."""
    sc = volume_scene(seed, mode)
    rng = np.random.default_rng(900 + seed)
    vol = sc.lights[0].volume
    cx, cy, cz = vol.cell_count
    pool = [LightDesc(LIGHT_POINT, NodeDesc(position=(2.0, 3.0, 1.0)), color=(1.0, 0.3, 0.2), energy=3.0, in_scene=False, name="cell_point"),
            LightDesc(LIGHT_POINT, NodeDesc(position=(-2.5, -1.0, 2.0)), color=(0.2, 0.4, 1.0), energy=2.0, range=6.0, in_scene=False, name="cell_ranged"),
            LightDesc(LIGHT_DIRECTIONAL, NodeDesc(rotation=W.rotation_matrix((1, 0, 0), 0.9)), color=(0.6, 0.9, 0.5), energy=0.8, in_scene=False, name="cell_sun"),
            LightDesc(LIGHT_AMBIENT, NodeDesc(), color=(0.3, 0.1, 0.3), energy=0.5, in_scene=False, name="cell_ambient"),
            sc.lights[1]]                           # the scene's own point light: it then lights such vertices twice, as in the reference
    for k in range(n_extra):        # many more lights than a light list's first allocation holds
        pool.append(LightDesc(LIGHT_POINT, NodeDesc(position=tuple(float(x) for x in rng.uniform(-5, 5, 3))), color=tuple(float(x) for x in rng.uniform(0.1, 1, 3)),
                              energy=float(rng.uniform(0.5, 2.0)), range=float(rng.choice([0.0, 5.0])), in_scene=False, name="cell_extra%d" % k))
    cell_lights = {}
    for z in range(cz):
        for y in range(cy):
            for x in range(cx):
                if rng.random() < 0.35:
                    k = int(rng.integers(1, 4))
                    cell_lights[(x, y, z)] = [pool[int(i)] for i in rng.choice(len(pool), size=k, replace=False)]
    vol.cell_lights = cell_lights
    if skinned:
        sk = skinned_scene(seed, n=2)
        for m in sk.models:
            m.light_group = None
        sc.models.extend(sk.models)
    sc.name = "volume_cell_lights"
    return sc


def subdivision_scene(seed=0, with_levels=True) -> SceneDesc:
    """Auto-subdivision (mesh.go:447-485, 1758-1959): meshes with pre-built subdivision levels at several distances from the
    camera, so that base triangles, level 0, 1 and 2 are all chosen somewhere; sizes that land in the packed, the one-block and
    the segmented cull paths; two MeshParts; vertex colours; next to an ordinary mesh."""
    from tetra3d_b200.scenedesc import auto_subdivide
    rng = np.random.default_rng(700 + seed)
    levels = [(900.0, 2.5), (225.0, 1.2), (49.0, 0.5)]
    cube = auto_subdivide(W.cube_mesh(2, 2, 2, material=MaterialDesc(backface_culling=True, triangle_sort_mode=seed % 3)), levels, 3)
    soup = auto_subdivide(random_soup(60, seed, n_parts=2, colors=True, extent=4.0, material=MaterialDesc(backface_culling=False)), levels, 2)
    grid = auto_subdivide(W.grid_mesh(24, 16, rows_per_part=8, seed=seed, material=MaterialDesc(backface_culling=True)), levels, 2)
    models = []
    for i, z in enumerate([6.0, 2.0, -4.0, -12.0, -30.0]):
        models.append(ModelDesc(cube, NodeDesc(position=(float(rng.uniform(-3, 3)), float(rng.uniform(-1, 1)), z), rotation=random_rotation(rng)), name="cube%d" % i))
    models.append(ModelDesc(soup, NodeDesc(position=(1.0, 0.5, 0.0), scale=(1.2, 0.8, 1.0), rotation=random_rotation(rng)), name="soup"))
    models.append(ModelDesc(grid, NodeDesc(position=(0.0, -2.0, -5.0), scale=(0.6, 1.0, 0.6)), frustum_culling=False, name="grid"))
    big = auto_subdivide(W.grid_mesh(24, 16, rows_per_part=16, seed=seed + 2, material=MaterialDesc(backface_culling=False, triangle_sort_mode=1)), levels, 2)
    models.append(ModelDesc(big, NodeDesc(position=(0.0, 3.0, -8.0), scale=(0.5, 1.0, 0.5), rotation=W.rotation_matrix((1, 0, 0), 0.4)), frustum_culling=False, name="grid_big"))
    models.append(ModelDesc(random_soup(200, seed + 1, n_parts=1), NodeDesc(position=(-2.0, 1.0, 1.0)), name="plain"))
    return SceneDesc(models, lights_mixed(rng, ranged=True), CameraDesc(node=NodeDesc(position=(0.0, 1.0, 10.0))), WorldDesc(ambient_energy=0.2),
                     name="subdivision", autosubdivision_levels=levels if with_levels else ())


def transparent_scene(seed=0) -> SceneDesc:
    rng = np.random.default_rng(seed)
    models = []
    for i in range(8):
        mat = MaterialDesc(transparency_mode=3 if i % 2 else 0, render_order=int(rng.integers(0, 3)), backface_culling=False,
                           color=(1, 1, 1, 0.5 if i % 3 == 0 else 1.0), triangle_sort_mode=i % 3)
        mesh = W.cube_mesh(1, 1.5, 1, material=mat)
        models.append(ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-4, 4, 3)), rotation=random_rotation(rng)),
                                color=(1, 1, 1, 1.0 if i % 4 else 0.6), name="t%d" % i))
    # fully transparent model: ProcessVertices early-out (model.go:604)
    models.append(ModelDesc(W.cube_mesh(1, 1, 1, material=MaterialDesc(transparency_mode=0)), NodeDesc(), color=(1, 1, 1, 0.0), name="ghost"))
    # invisible model and invisible material
    models.append(ModelDesc(W.cube_mesh(), NodeDesc(position=(1.0, 1.0, 1.0)), visible=False, name="hidden"))
    models.append(ModelDesc(W.cube_mesh(material=MaterialDesc(visible=False)), NodeDesc(position=(-1.0, 1.0, 1.0)), name="hiddenmat"))
    # no material at all
    nomat = W.cube_mesh(); nomat.parts[0].material = None
    models.append(ModelDesc(nomat, NodeDesc(position=(0.0, -2.0, 0.0)), name="nomat"))
    return SceneDesc(models, lights_mixed(rng), CameraDesc(), WorldDesc(), name="transparent")


def render_order_scene(seed=0, n=45) -> SceneDesc:
    """More than 20 solids with mixed RenderOrder under RenderDepth = false: the solids comparator of camera.go:1460-1469 is not
    a strict weak ordering, so the draw order is defined by Go's sort.SliceStable alone (blocks of 20 + SymMerge)."""
    rng = np.random.default_rng(seed)
    models = []
    for i in range(n):
        mat = MaterialDesc(render_order=int(rng.integers(-2, 3)), backface_culling=bool(i % 2), triangle_sort_mode=i % 3)
        mesh = W.cube_mesh(0.6, 0.8, 0.6, material=mat)
        models.append(ModelDesc(mesh, NodeDesc(position=tuple(float(x) for x in rng.uniform(-7, 7, 3)), rotation=random_rotation(rng)), name="ro%d" % i))
    sc = SceneDesc(models, lights_mixed(rng), CameraDesc(render_depth=False), WorldDesc(), name="render_order%d" % seed)
    return sc


def max_light_count_scene(seed=0, n_lights=7, max_count=3, focus=False) -> SceneDesc:
    """Camera.MaxLightCount (camera.go:1474-1497): more visible lights than the cap, two ambient lights (the world's
    and one in the tree: the comparator is inconsistent for them), a sun, point lights at scattered distances."""
    rng = np.random.default_rng(seed)
    sc = soup_scene(seed, n_models=3, n_tris=300)
    ls = []
    for i in range(n_lights):
        kind = LIGHT_POINT if i % 5 not in (1, 3) else (LIGHT_AMBIENT if i % 5 == 1 else LIGHT_DIRECTIONAL)
        ls.append(LightDesc(kind, NodeDesc(position=tuple(float(x) for x in rng.uniform(-12, 12, 3)), rotation=random_rotation(rng)),
                            color=tuple(float(x) for x in rng.uniform(0.2, 1, 3)), energy=float(rng.uniform(1, 5)),
                            range=float(rng.uniform(6, 20)) if i % 2 else 0.0, name="ml%d" % i))
    sc.lights = ls
    sc.camera.max_light_count = max_count
    if focus:
        sc.camera.max_light_focus = NodeDesc(position=(4.0, 1.0, -3.0), name="focus")
    sc.name = "maxlights%d" % seed
    return sc


ALL_SMALL = {
    "soup_lit": lambda: soup_scene(1),
    "max_light_count": lambda: max_light_count_scene(51),
    "max_light_count_focus": lambda: max_light_count_scene(52, n_lights=9, max_count=4, focus=True),
    "max_light_count_many": lambda: max_light_count_scene(53, n_lights=47, max_count=6),
    "soup_unlit_colors": lambda: soup_scene(2, lit=False, colors=True),
    "soup_ranged": lambda: soup_scene(3, ranged=True),
    "soup_parent": lambda: soup_scene(4, parent=True),
    "soup_ortho": lambda: soup_scene(5, cam=CameraDesc(perspective=False, ortho_scale=16.0)),
    "soup_snap_pctm_normals": lambda: soup_scene(6, cam=CameraDesc(vertex_snapping=8.0, perspective_corrected_texture_mapping=True, render_normals=True),
                                                 material=MaterialDesc(use_texture=True, tex_w=64, tex_h=32)),
    "soup_fog_add": lambda: soup_scene(7, cam=CameraDesc(render_depth=False), world=WorldDesc(fog_mode=0, fog_color=(0.3, 0.2, 0.1, 1), fog_range=(0.1, 0.9))),
    "soup_fog_sub": lambda: soup_scene(8, cam=CameraDesc(render_depth=False), world=WorldDesc(fog_mode=1, fog_color=(0.9, 0.8, 0.7, 1))),
    "soup_fixed_normals": lambda: soup_scene(9, material=MaterialDesc(lighting_mode=1, backface_culling=False)),
    "soup_double_sided": lambda: soup_scene(10, material=MaterialDesc(lighting_mode=2, backface_culling=False, triangle_sort_mode=1)),
    "soup_lighting_off": lambda: soup_scene(11, world=WorldDesc(lighting_on=False)),
    "soup_no_world": lambda: soup_scene(12, world=WorldDesc(present=False)),
    "skinned": lambda: skinned_scene(1),
    "shape_keys": lambda: shape_key_scene(1),
    "billboard": lambda: billboard_scene(1),
    "billboard_locks": lambda: billboard_scene(2, locks=(True, False, True), camera_up=True),
    "billboard_lock_y_unbill": lambda: billboard_scene(3, locks=(False, True, False), unbillboarded=True),
    "batch_transparent": lambda: batch_scene(1),
    "batch_solid": lambda: batch_scene(2, transparent=False),
    "volume_mode0": lambda: volume_scene(1, 0),
    "volume_mode1": lambda: volume_scene(2, 1),
    "volume_mode2": lambda: volume_scene(3, 2),
    "volume_off": lambda: volume_scene(4, 3),
    "subdivision": lambda: subdivision_scene(0),
    "subdivision_no_levels": lambda: subdivision_scene(1, with_levels=False),
    "posed": lambda: posed_scene(0),
    "posed_chain": lambda: posed_scene(3, n_bones=8, chain=True),
    "volume_cell_lights": lambda: volume_cell_lights_scene(5, 0),
    "volume_cell_lights_normal": lambda: volume_cell_lights_scene(6, 1, skinned=True),
    "volume_cell_lights_object": lambda: volume_cell_lights_scene(7, 2),
    "transparent": lambda: transparent_scene(1),
    "render_order": lambda: render_order_scene(1),
    "render_order_many": lambda: render_order_scene(2, n=130),
    "stress_small": lambda: SceneDesc([ModelDesc(W.cube_mesh(2, 2, 2, copies=20, material=MaterialDesc(shadeless=True, use_texture=True, tex_w=16, tex_h=16)),
                                                  NodeDesc(position=(4.0 * x, 0.0, 0.0)), name="m%d" % x) for x in range(6)], [], CameraDesc(), WorldDesc()),
}


def random_skeleton(rng, n_bones: int, chain: bool = False) -> "SkeletonDesc":
    """A random bone tree (parents first; several roots possible) with random rest transforms and inverse binds."""
    from tetra3d_b200.scenedesc import SkeletonDesc
    parent = np.full(n_bones, -1, np.int32)
    for b in range(1, n_bones):
        parent[b] = b - 1 if chain else int(rng.integers(-1, b)) if rng.random() < 0.9 else -1
    inv = np.zeros((n_bones, 16), F32); rot = np.zeros((n_bones, 16), F32)
    for b in range(n_bones):
        m = random_rotation(rng).astype(np.float64).reshape(4, 4)
        m[3, :3] = rng.uniform(-0.5, 0.5, 3)
        inv[b] = m.astype(F32).reshape(16)
        rot[b] = random_rotation(rng).reshape(16)
    return SkeletonDesc(names=["b%d" % b for b in range(n_bones)], parent=parent, inverse_bind=inv,
                        rest_position=rng.uniform(-0.3, 0.3, (n_bones, 3)).astype(F32), rest_scale=rng.uniform(0.8, 1.2, (n_bones, 3)).astype(F32),
                        rest_rotation=rot)


def random_animation(rng, sk, length: float = 3.0) -> "AnimationDesc":
    """Random tracks: some bones without a channel, tracks of 1..12 keys, repeated key times, stepped tracks, opposite-hemisphere
    and nearly equal quaternion pairs (both Slerp branches), one all-zero quaternion (SetLocalRotation ignores it)."""
    from tetra3d_b200.scenedesc import AnimationDesc, TrackDesc
    channels = {}
    for b, name in enumerate(sk.names):
        if rng.random() < 0.2:
            continue
        tracks = {}
        for t in range(3):
            if rng.random() < 0.25:
                continue
            k = int(rng.integers(1, 13))
            times = np.sort(rng.uniform(0.0, length, k)).astype(F32)
            if k > 3 and rng.random() < 0.3:
                times[2] = times[1]
            if t == 2:
                v = rng.standard_normal((k, 4))
                v /= np.linalg.norm(v, axis=1, keepdims=True)
                for i in range(1, k):
                    if rng.random() < 0.3:
                        v[i] = v[i - 1] + 1e-4 * rng.standard_normal(4)      # cosOmega > 0.9999: the linear branch
                if k > 2 and rng.random() < 0.15:
                    v[1] = 0.0
                v = v.astype(F32)
            elif t == 1:
                v = np.concatenate([rng.uniform(0.7, 1.3, (k, 3)), np.zeros((k, 1))], axis=1).astype(F32)
            else:
                v = np.concatenate([rng.uniform(-0.4, 0.4, (k, 3)), np.zeros((k, 1))], axis=1).astype(F32)
            tracks[t] = TrackDesc(times, v, interpolation=int(rng.choice([0, 0, 1, 2])))
        channels[name] = tracks
    return AnimationDesc("random", length, channels)


def posed_scene(seed=0, n_models=6, n_bones=5, chain=False, playheads=None, blends=None) -> SceneDesc:
    """Skinned characters whose bones come from a skeleton + animation sampled at a playhead (PoseDesc): two skeletons, two
    animations and a rest-pose armature in one frame, next to a model with explicit bone matrices and a static one.
    blends: milliseconds into a one-second blend, cycled over the models (None entries: not blending); a blending player
    comes from the skeleton's other animation, frozen at a random playhead (every fourth: from nothing)."""
    from tetra3d_b200.scenedesc import PoseDesc
    rng = np.random.default_rng(500 + seed)
    mesh, bone_y = W.tube_character_mesh(rings=9, sides=10, n_bones=n_bones)
    sks = [random_skeleton(rng, n_bones, chain=chain), W.chain_skeleton(bone_y)]
    anims = [[random_animation(rng, sks[0]), random_animation(rng, sks[0], length=1.5)], [W.chain_animation(sks[1]), random_animation(rng, sks[1])]]
    models = []
    for i in range(n_models):
        k = i % 2
        anim = None if i == 3 else anims[k][(i // 2) % 2]
        ph = float(playheads[i % len(playheads)]) if playheads is not None else float(rng.uniform(-0.2, 3.3))
        root = NodeDesc(position=(1.5 * (i - n_models / 2), -1.0, 0.0), rotation=random_rotation(rng) if i % 3 == 0 else W.rotation_matrix((0, 1, 0), 0.0),
                        scale=(1.0, 1.0 + 0.1 * (i % 2), 1.0))
        pose = PoseDesc(sks[k], anim, ph, root)
        if blends is not None and anim is not None and blends[i % len(blends)] is not None:
            pose.blend_ms = int(blends[i % len(blends)])
            pose.blend_from = None if i % 4 == 2 else anims[k][1 - (i // 2) % 2]
            pose.from_playhead = float(rng.uniform(-0.2, 3.3))
        models.append(ModelDesc(mesh, NodeDesc(), pose=pose, frustum_culling=False, name="posed%d" % i))
    bi = W.chain_bone_influences(bone_y, (0.0, 1.5, 0.0), 0.7, 0.3)
    models.append(ModelDesc(mesh, NodeDesc(), bone_influences=bi, frustum_culling=False, name="explicit"))
    models.append(ModelDesc(random_soup(300, seed, n_parts=1), NodeDesc(position=(0.0, 0.0, -2.0)), name="static"))
    lights = [LightDesc(LIGHT_POINT, NodeDesc(position=(-3.0, 2.0, 5.0)), color=(0.9, 0.9, 0.5), energy=3.0, name="scene_point"),
              LightDesc(LIGHT_DIRECTIONAL, NodeDesc(rotation=W.rotation_matrix((1, 0, 0), 0.7)), color=(0.5, 0.6, 1.0), energy=1.0, name="sun")]
    return SceneDesc(models, lights, CameraDesc(node=NodeDesc(position=(0.0, 1.0, 9.0))), WorldDesc(ambient_energy=0.15), name="posed%d" % seed)
