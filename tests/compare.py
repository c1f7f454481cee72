"""GPU (host mirror -> C ABI -> CUDA kernels) versus CPU oracle on the same SceneDesc."""
from __future__ import annotations

import numpy as np

import tetra3d_b200 as t3d
from oracle import binding as O

# tolerances of BASELINE.json's north_star
CLIP_REL_TOL = 1e-4     # transformed positions, relative, clip space
COLOR_TOL = 1.0 / 255   # lit colours


def render_gpu(desc, device, shard=None):
    scene, cam, objs = t3d.build_scene(desc, device)
    if shard is not None:
        cam.SetShard(*shard)
    cam.Clear()
    cam.RenderScene(scene)
    return scene, cam, objs


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def same_bits(a, b):
    """Bit-exact, except that any NaN equals any NaN (payload bits are not part of the contract)."""
    a = np.ascontiguousarray(a, dtype=np.float32); b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


# largest |colour - oracle colour| seen with the default (approximate-reciprocal) lighting, per scene name; test_gpu_parity
# prints / writes it at the end of the session
FAST_COLOR_MAXDIFF = {}


def assert_parity(desc, device, check_capture=True, exact=True, modes=("fast", "exact")):
    """Bit-exact: post-cull triangle set, sorted order, per-render counts, min/max depth, index list,
    positions and depth colours -- in both lighting modes of the library.  Colours: within 1/255 with the default
    lighting (single-instruction reciprocals), bit-exact with t3d_set_exact_lighting (unless `exact` is False)."""
    ref = O.render_scene(desc, capture=check_capture)
    out = None
    try:
        for mode in modes:
            device.SetExactLighting(mode == "exact")
            out = _check_frame(desc, device, ref, check_capture, exact and mode == "exact", mode)
    finally:
        try:
            device.SetExactLighting(False)
        except Exception:       # (must not mask the failure that brought us here)
            pass
    return out


def _check_frame(desc, device, ref, check_capture, exact, mode):
    scene, cam, objs = render_gpu(desc, device)
    names = [m.name for m in desc.models]

    renders = cam.Renders()
    assert len(renders) == len(ref.renders), "render() call count %d vs oracle %d" % (len(renders), len(ref.renders))
    for g, r in zip(renders, ref.renders):
        assert g["model"] == names[r["model"]] and g["part"] == r["part"] and g["flush"] == r["flush"], (g, r)
        assert g["vertex_count"] == r["vertex_count"], ("survivor count", g, r)
        assert g["first_vertex"] == r["first_vertex"], (g, r)
        assert g["total_tris"] == r["total_tris"], (g, r)
        if r["vertex_count"]:
            assert bits(g["min_depth"]) == bits(r["min_depth"]) and bits(g["max_depth"]) == bits(r["max_depth"]), ("min/max depth", g, r)
    flushes = cam.Flushes()
    assert len(flushes) == len(ref.flushes)
    for g, r in zip(flushes, ref.flushes):
        assert g["model"] == names[r["model"]] and g["part"] == r["part"]
        assert g["first_vertex"] == r["first_vertex"] and g["vertex_count"] == r["vertex_count"], (g, r)
        idx = cam.IndexList(r["vertex_count"])
        assert np.array_equal(idx, ref.index_list[r["first_vertex"]:r["first_vertex"] + r["vertex_count"]])

    assert cam.VertexCount() == ref.n_vertices
    tri = cam.TriangleIDs()
    assert np.array_equal(tri, ref.tri_ids), "sorted triangle order differs (first at %s)" % (np.flatnonzero(tri != ref.tri_ids)[:5],)

    col = cam.ColorVertexList(); dep = cam.DepthVertexList()
    # DstX, DstY, SrcX, SrcY and Custom0..3 come from clip positions only: bit-exact
    for name, a, b in (("color", col, ref.color_list), ("depth", dep, ref.depth_list)):
        for cols, what in (((0, 1), "Dst"), ((2, 3), "Src"), ((8, 9, 10, 11), "Custom")):
            ga, rb = a[:, cols], b[:, cols]
            assert same_bits(ga, rb), "%s list %s not bit-exact: max abs diff %g" % (name, what, np.nanmax(np.abs(ga - rb)))
    assert same_bits(dep[:, 4:8], ref.depth_list[:, 4:8]), "depth colours not bit-exact"
    cdiff = np.abs(col[:, 4:8].astype(np.float64) - ref.color_list[:, 4:8].astype(np.float64))
    assert (not cdiff.size) or np.nanmax(cdiff) <= COLOR_TOL, "colours differ by %g > 1/255" % np.nanmax(cdiff)
    if mode == "fast" and cdiff.size:
        key = getattr(desc, "name", None) or "scene"
        FAST_COLOR_MAXDIFF[key] = max(FAST_COLOR_MAXDIFF.get(key, 0.0), float(np.nanmax(cdiff)))
    if exact:
        assert same_bits(col[:, 4:8], ref.color_list[:, 4:8]), "colours not bit-exact: max diff %g" % (np.nanmax(cdiff) if cdiff.size else 0)
    if desc.camera.render_normals:
        nor = cam.NormalVertexList()
        assert same_bits(nor, ref.normal_list), "normal list not bit-exact"

    dbg = cam.DebugInfo()
    for k in ("totalParts", "drawnParts", "batchedParts", "totalTris", "drawnTris", "lightCount", "activeLightCount"):
        assert dbg[k] == ref.debug[k], (k, dbg[k], ref.debug[k])

    if check_capture:
        # per-vertex intermediates: clip positions bit-exact (and therefore <= 1e-4 relative), lights <= 1/255
        off = 0
        for g, r in zip(renders, ref.renders):
            md = desc.models[r["model"]]
            part = md.mesh.parts[r["part"]]
            n = part.vertex_end - part.vertex_start
            rc = ref.cap_clip[off:off + n]; rl = ref.cap_light[off:off + n]
            off += n
            if g["draw_index"] < 0 or n == 0:
                continue
            used = np.unique(md.mesh.indices[part.tri_start:part.tri_end + 1])
            if used.size == 0:
                continue
            lo, hi = int(used.min()), int(used.max()) + 1
            gc, gl = device.ReadDrawVertices(g["draw_index"], lo, hi)
            sel = used - lo
            rsel = used - part.vertex_start
            ok = (rsel >= 0) & (rsel < n)
            sel, rsel = sel[ok], rsel[ok]
            a, b = gc[sel], rc[rsel]
            assert same_bits(a, b), "clip positions not bit-exact (max rel %g)" % np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1e-30))
            lit = rl[rsel][:, 3] == 1.0     # vertices the oracle's lighting loop touched
            if lit.any():
                ld = np.abs(gl[sel][lit][:, :3].astype(np.float64) - rl[rsel][lit][:, :3].astype(np.float64))
                assert np.nanmax(ld) <= COLOR_TOL, "vertex lights differ by %g" % np.nanmax(ld)
    return cam, ref
