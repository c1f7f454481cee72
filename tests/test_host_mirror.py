"""Host mirror (C++ / pybind11) against the oracle for everything that stays on the host: node transforms,
camera matrices, mesh-derived data.  No GPU needed."""
import numpy as np

import scenes
import tetra3d_b200 as t3d
from oracle import binding as O
from tetra3d_b200 import workloads as W

F32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, F32).view(np.uint32)


def test_node_transform_matches_oracle_bitwise():
    rng = np.random.default_rng(0)
    parent = t3d.Node("parent")
    parent.SetLocalPosition(1, 2, 3); parent.SetLocalScale(1.5, 0.5, 2.0); parent.SetLocalRotation(W.rotation_matrix((1, 2, 3), 0.7))
    child = t3d.Node("child")
    r = W.rotation_matrix(rng.normal(size=3), 1.1)
    child.SetLocalPosition(-4, 0.5, 2); child.SetLocalScale(2, 2, 0.25); child.SetLocalRotation(r)
    parent.AddChildren(child)
    nodes = (O.orc_node * 2)()
    nodes[0].position[:] = [1, 2, 3]; nodes[0].scale[:] = [1.5, 0.5, 2.0]; nodes[0].rotation[:] = list(W.rotation_matrix((1, 2, 3), 0.7).reshape(16)); nodes[0].parent = -1
    nodes[1].position[:] = [-4, 0.5, 2]; nodes[1].scale[:] = [2, 2, 0.25]; nodes[1].rotation[:] = list(r.reshape(16)); nodes[1].parent = 0
    out = np.zeros(16, F32)
    O.lib().orc_node_transform(nodes, 1, out.ctypes.data_as(O.f32p))
    assert np.array_equal(bits(child.Transform()), bits(out.reshape(4, 4)))           # node.go:572-608


def test_matrix_helpers_match_oracle_bitwise():
    rng = np.random.default_rng(1)
    for _ in range(20):
        a = rng.normal(size=(4, 4)).astype(F32); b = rng.normal(size=(4, 4)).astype(F32)
        assert np.array_equal(bits(t3d.Matrix4Mult(a, b)), bits(O.mat4_mult(a, b)))
        assert np.array_equal(bits(t3d.Matrix4Inverted(a)), bits(O.mat4_inverted(a)))
    assert np.array_equal(bits(t3d.NewMatrix4Rotate(0.3, 1, -2, 0.77)), bits(O.mat4_rotate(0.3, 1, -2, 0.77)))


def test_camera_matrices_match_oracle():
    cam = t3d.NewCamera("cam", 640, 360)
    cam.SetLocalPosition(0.5, 1.0, 14.0); cam.SetLocalRotation(W.rotation_matrix((0, 1, 0), 0.05))
    assert np.array_equal(bits(cam.Projection()), bits(O.mat4_perspective(60, 0.1, 100, 640, 360)))   # matrix.go:588-604
    view = O.mat4_mult(O.mat4_translate(-0.5, -1.0, -14.0), O.mat4_transposed(W.rotation_matrix((0, 1, 0), 0.05)))
    assert np.array_equal(bits(cam.ViewMatrix()), bits(view))                                        # camera.go:638-648


def test_cube_mesh_matches_reference_tables():
    m = t3d.NewCubeMesh(2, 2, 2)
    d = W.cube_mesh(2, 2, 2)
    assert m.VertexCount() == 24 and m.TriangleCount() == 12                                          # mesh.go:1251-1342
    assert np.array_equal(m.Positions(), d.positions) and np.array_equal(m.Indices(), d.indices) and np.array_equal(m.UVs(), d.uvs)
    assert np.allclose(m.Normals(), d.normals)
    assert np.array_equal(bits(m.Centers()), bits(O.triangle_centers(d.positions, d.indices)))       # mesh.go:1980-1982
    assert np.isclose(m.MaxSpan(), np.sqrt(12.0))                                                    # mesh.go:41-43


def test_build_scene_preserves_traversal_order():
    sc = scenes.soup_scene(4, parent=True)
    scene, cam, objs = t3d.build_scene(sc)
    # depth-first order of RenderNodes (camera.go:1200-1210) == description order
    order = []
    def walk(n):
        for c in n.Children():
            order.append(c.Name()); walk(c)
    walk(scene.Root)
    models = [n for n in order if n.startswith("m")]
    assert models == [m.name for m in sc.models]
