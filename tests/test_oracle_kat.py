"""Pins the oracle's math against the only assertion-bearing test of the reference (matrix_test.go:36-55)
and against an independent numpy float32 evaluation of the cited formulas."""
import math
import struct

import numpy as np
import pytest

from oracle import binding as O

F32 = np.float32


def np_mult(a, b):
    """matrix_none.go:6-32 / matrix_amd64.s: ((a0*b0 + a1*b1) + a2*b2) + a3*b3 in float32, no fusion."""
    a = a.astype(F32); b = b.astype(F32)
    out = np.zeros((4, 4), F32)
    for i in range(4):
        for j in range(4):
            acc = F32(a[i, 0] * b[0, j])
            for k in range(1, 4):
                acc = F32(acc + F32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def test_matrix_inversion_known_answer():
    """TestMatrixInversion (matrix_test.go:36-55): M * M.Inverted() must be IsIdentity() for the four matrices.
    The fourth has a zero Z scale: its inverse is non-finite and Matrix4.Equals (matrix.go:511-522) lets NaN
    differences through, which is how the reference's own test passes; the oracle reproduces that."""
    mats = [
        O.mat4_rotate(0, 1, 0, 0.1),
        O.mat4_translate(-10, 0.1, 3232.1976),
        O.mat4_scale(10, 0.1, -0.45),
        O.mat4_mult(O.mat4_mult(O.mat4_translate(-1, -1, -1), O.mat4_rotate(1, 0, 0.1, 0.334)), O.mat4_scale(10, 1, 0)),
    ]
    ident = O.mat4_identity()
    for i, m in enumerate(mats):
        prod = O.mat4_mult(m, O.mat4_inverted(m))
        assert O.mat4_equals(prod, ident), "matrix #%d: M * M.Inverted() is not identity" % i
    # the first three are genuinely invertible: check against float64
    for m in mats[:3]:
        inv = O.mat4_inverted(m)
        assert np.allclose(inv.astype(np.float64), np.linalg.inv(m.astype(np.float64)), rtol=1e-5, atol=1e-6)
    assert not np.isfinite(O.mat4_inverted(mats[3])).all()


def test_mult_matches_float32_evaluation_order():
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = rng.normal(size=(4, 4)).astype(F32) * F32(10)
        b = rng.normal(size=(4, 4)).astype(F32)
        got = O.mat4_mult(a, b)
        assert np.array_equal(got.view(np.uint32), np_mult(a, b).view(np.uint32))


def test_transposed_and_translate_scale_builders():
    m = np.arange(16, dtype=F32).reshape(4, 4)
    assert np.array_equal(O.mat4_transposed(m), m.T)
    t = O.mat4_translate(1, 2, 3)
    assert np.array_equal(t[3], np.array([1, 2, 3, 1], F32)) and np.array_equal(t[:3, :3], np.eye(3, dtype=F32))   # matrix.go:81-87
    s = O.mat4_scale(2, 3, 4)
    assert np.array_equal(np.diag(s), np.array([2, 3, 4, 1], F32))


def test_rotate_is_counter_clockwise_about_axis():
    # matrix.go:100-128: rotating +X about +Y by 90 degrees; vectors are rows (v * M)
    r = O.mat4_rotate(0, 1, 0, np.pi / 2)
    v = np.array([1, 0, 0, 0], F32) @ r
    assert np.allclose(v[:3], [0, 0, -1], atol=1e-6)


def test_decompose_roundtrip():
    rng = np.random.default_rng(1)
    for _ in range(20):
        s = rng.uniform(0.5, 2, 3).astype(F32)
        r = O.mat4_rotate(*rng.normal(size=3), float(rng.uniform(0, 6)))
        p = rng.uniform(-5, 5, 3).astype(F32)
        m = O.mat4_mult(O.mat4_mult(O.mat4_scale(*s), r), O.mat4_translate(*p))        # node.go:581-583
        pos, sc, rot = O.mat4_decompose(m)
        assert np.array_equal(pos, p)
        assert np.allclose(sc, s, rtol=1e-5)
        assert np.allclose(rot, r, atol=1e-5)


def test_perspective_matrix_layout():
    # matrix.go:588-604, row-vector convention: clip.W = -z_view * (2 f n)/(f-n) ... row 2 column 3
    m = O.mat4_perspective(60, 0.1, 100, 640, 360)
    t = np.tan(np.float32(60) * np.float32(np.pi) / 360)
    assert np.isclose(m[1, 1], 0.1 / t, rtol=1e-6)
    assert np.isclose(m[0, 0], 0.1 / (t * 640 / 360), rtol=1e-6)
    assert np.isclose(m[2, 2], -(100.1) / (99.9), rtol=1e-6) and np.isclose(m[2, 3], -(2 * 100 * 0.1) / 99.9, rtol=1e-6)
    assert m[3, 2] == -1 and m[3, 3] == 0


def test_lookat_degenerate_cases():
    assert np.array_equal(O.mat4_lookat((1, 2, 3), (1, 2, 3), (0, 1, 0)), np.eye(4, dtype=F32))      # matrix.go:791-793
    m = O.mat4_lookat((0, 0, 0), (0, 5, 0), (0, 1, 0))                                              # z == up -> WorldRight
    assert np.allclose(m[2, :3], [0, 1, 0]) and np.isfinite(m).all()


def test_triangle_centers_formula():
    pos = np.array([[0, 0, 0], [3, 0, 0], [0, 3, 3], [1, 1, 1]], F32)
    idx = np.array([[0, 1, 2], [3, 2, 1]], np.uint32)
    c = O.triangle_centers(pos, idx)
    exp = np.array([[(F32(0) + F32(3) + F32(0)) / F32(3), 1, 1], [(F32(1) + F32(0) + F32(3)) / F32(3), (F32(1) + F32(3) + F32(0)) / F32(3), (F32(1) + F32(3) + F32(0)) / F32(3)]], F32)
    assert np.array_equal(c, exp)                                                                    # mesh.go:1980-1982


@pytest.mark.parametrize("depth,mn,mx,mode,expected", [
    (5.0, 5.0, 5.0, 0, 0),          # rangeDiff == 0 -> +0.001 (sortingTriangles.go:43-46): (0/0.001)*512 = 0
    (0.0, 0.0, 10.0, 0, 0),
    (10.0, 0.0, 10.0, 0, 511),      # clamp to binCount-1 (sortingTriangles.go:54)
    (5.0, 0.0, 10.0, 0, 256),
    (9.99, 0.0, 10.0, 1, 511),
    (5.0, 0.0, 10.0, 2, 0),         # TriangleSortModeNone -> bin 0 (sortingTriangles.go:52)
])
def test_bin_index(depth, mn, mx, mode, expected):
    assert O.bin_index(depth, mn, mx, mode) == expected


def test_go_stable_sort_restatement_is_a_stable_sort():
    """Camera.MaxLightCount sorts with sort.SliceStable and an inconsistent comparator (camera.go:1484-1495), so the
    oracle restates Go's algorithm (insertion sort on blocks of 20 + SymMerge).  With a consistent comparator the
    answer is unique: it must equal numpy's stable argsort for every size around the block boundaries."""
    import ctypes
    import numpy as np
    from oracle import binding as O
    lib = O.lib()
    lib.orc_test_go_stable_sort.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]
    lib.orc_test_go_stable_sort.restype = None
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) + [79, 80, 81, 159, 160, 161, 500, 1000]:
        for distinct in (3, 10, 10 ** 6):          # many ties -> stability matters
            keys = rng.integers(0, distinct, n).astype(np.float32)
            perm = np.zeros(max(n, 1), np.int32)
            lib.orc_test_go_stable_sort(n, keys.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), perm.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
            assert np.array_equal(perm[:n], np.argsort(keys, kind="stable")), (n, distinct)

# ---- the camera's trigonometry: Go's pure-Go math.Tan / math.Cos / math.Atan (Cephes-derived, src/math/tan.go, sin.go, atan.go)
# against the C library's, which the oracle and the host mirror call for the projection matrix (matrix.go:592) and the frustum
# sphere factors (camera.go:660-666).  The two differ by up to 2 ulp in float64; what matters is the value after math32's
# conversion to float32.
def f64(bits): return struct.unpack('<d', struct.pack('<Q', bits))[0]
PI4A, PI4B, PI4C = f64(0x3fe921fb40000000), f64(0x3e64442d00000000), f64(0x3ce8469898cc5170)
M4PI = 4/math.pi  # Go: 4 / Pi constant (exact constant arithmetic then rounded to float64)
SIN = [f64(0x3de5d8fd1fd19ccd), f64(0xbe5ae5e5a9291f5d), f64(0x3ec71de3567d48a1), f64(0xbf2a01a019bfdf03), f64(0x3f8111111110f7d0), f64(0xbfc5555555555548)]
COS = [f64(0xbda8fa49a0861a9b), f64(0x3e21ee9d7b4e3f05), f64(0xbe927e4f7eac4bc6), f64(0x3efa01a019c844f5), f64(0xbf56c16c16c14f91), f64(0x3fa555555555554b)]
TANP = [-1.30936939181383777646e4, 1.15351664838587416140e6, -1.79565251976484877988e7]
TANQ = [1.0, 1.36812963470692954678e4, -1.32089234440210967447e6, 2.50083801823357915839e7, -5.38695755929454629881e7]
def go_tan(x):
    sign = False
    if x < 0: x = -x; sign = True
    j = int(x * M4PI); y = float(j)
    if j & 1: j += 1; y += 1
    z = ((x - y*PI4A) - y*PI4B) - y*PI4C
    zz = z*z
    if zz > 1e-14:
        y = z + z*(zz*(((TANP[0]*zz)+TANP[1])*zz+TANP[2])/((((zz+TANQ[1])*zz+TANQ[2])*zz+TANQ[3])*zz+TANQ[4]))
    else:
        y = z
    if j & 2: y = -1/y
    return -y if sign else y
def go_cos(x):
    sign = False
    x = abs(x)
    j = int(x * M4PI); y = float(j)
    if j & 1: j += 1; y += 1
    j &= 7
    z = ((x - y*PI4A) - y*PI4B) - y*PI4C
    if j > 3: j -= 4; sign = not sign
    if j > 1: sign = not sign
    zz = z*z
    if j == 1 or j == 2:
        y = z + z*zz*((((((SIN[0]*zz)+SIN[1])*zz+SIN[2])*zz+SIN[3])*zz+SIN[4])*zz+SIN[5])
    else:
        y = 1.0 - 0.5*zz + zz*zz*((((((COS[0]*zz)+COS[1])*zz+COS[2])*zz+COS[3])*zz+COS[4])*zz+COS[5])
    return -y if sign else y


def test_libm_trigonometry_equals_go_for_every_camera_angle():
    import math
    from oracle import binding as O
    F = np.float32
    worst = 0.0
    for fov10 in range(5, 1795, 5):                      # fields of view 0.5 .. 179 degrees in half-degree steps
        fov = F(fov10 / 10)
        for a in (float(fov * F(math.pi) / F(360)), float(fov * F(3.1415) / F(360))):          # matrix.go:592, camera.go:660
            t_go, t_c, c_go, c_c = go_tan(a), math.tan(a), go_cos(a), math.cos(a)
            worst = max(worst, abs(t_go - t_c) / math.ulp(t_c), abs(c_go - c_c) / math.ulp(c_c))
            assert F(t_go) == F(t_c) and F(c_go) == F(c_c), (fov, a)
            for aspect in (16 / 9, 4 / 3, 1.0, 21 / 9, 2.0):
                x = float(F(t_c) * F(aspect))
                at_go, at_c = O.go_atan2(x, 1.0), math.atan(x)
                assert F(at_go) == F(at_c) and F(go_cos(float(F(at_go)))) == F(math.cos(float(F(at_c)))), (fov, aspect)
    assert worst <= 4.0          # a mistyped Cephes coefficient would be off by many orders of magnitude more

