"""A second restatement of the hot path, written separately from oracle/t3d_oracle.c and structured like the Go source.

TEST INFRASTRUCTURE ONLY (tests/test_literal_restatement.py).  Where the oracle is flat C over arrays, this file keeps the
reference's own shape: `process_vertices` walks `meshPart.forEachTri` and fills `globalVertexTransforms` /
`globalMeshAlteredVertexPositions` exactly as model.go:600-995 does, `SortingTriangleBucket` is the 512-bin bucket of
sortingTriangles.go with its bins as Python lists, the lights are objects with `beginRender` / `beginModel` / `Light` /
`VertexLight` (light.go:110-127, 175-269, 322-372), and `render` is the closure of camera.go:1507-1827 (lighting loop over
the bucket, then the emit loop).  Arithmetic is numpy float32 scalars, one operation per Go operation, left to right.

What it covers: static and skinned models (skinVertex, shape keys), billboards with their locks and the unbillboarded depth mode,
the children of dynamic batches (each is its own render(rp)), auto-subdivision (handleSubdivision with closestPointOnTri);
no host callbacks; perspective
and orthographic cameras; RenderDepth on or off (fog add / sub); vertex colours; texture sizes; perspective-corrected texture
mapping; ambient, point (ranged and unranged), directional and LightVolume lights (all shading modes, cells with lights of their own),
all three lighting modes, LightGroups; every sort mode.
What it borrows from the oracle's binding: Triangle.Center (load-time data) and, for posed armatures, the bone matrices
(tests/test_pose_oracle.py has their own literal check).  Matrix4.Mult / Inverted / Decompose, Node.Transform, the view and
projection matrices and everything else are restated here.
"""
import math

import numpy as np

from oracle import binding as O

F = np.float32
MAXF = F(np.finfo(np.float32).max)


# ---- vector.go / matrix.go --------------------------------------------------------------------------------------------
def sqrt32(x):                      # math32.Sqrt: float32(math.Sqrt(float64(x)))
    return F(math.sqrt(float(x))) if x >= 0 else F(np.nan)


def magnitude(v):                   # vector.go:191-193
    return sqrt32(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def unit(v):                        # vector.go:258-266 (also Vector4.Unit, vector.go:1869-1877: W is left alone)
    l = magnitude(v)
    if l < F(1e-8) or l == F(1):
        return v
    out = list(v)
    out[0], out[1], out[2] = v[0] / l, v[1] / l, v[2] / l
    return out


def dot(a, b):                      # vector.go:526-528
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def distance_squared_to(a, b):      # vector.go:242-247
    x, y, z = a[0] - b[0], a[1] - b[1], a[2] - b[2]
    return x * x + y * y + z * z


def clamp(v, lo, hi):               # math32.Clamp
    if v < lo:
        return lo
    if v > hi:
        return hi
    return v


def mult_vec(m, v):                 # matrix.go:619-627
    return [m[0][0] * v[0] + m[1][0] * v[1] + m[2][0] * v[2] + m[3][0],
            m[0][1] * v[0] + m[1][1] * v[1] + m[2][1] * v[2] + m[3][1],
            m[0][2] * v[0] + m[1][2] * v[1] + m[2][2] * v[2] + m[3][2]]


def mat(a):
    return np.asarray(a, F).reshape(4, 4)


def mat4_mult(a, b):                # matrix_amd64.s (amd64) == matrix_none.go: ret[i][j] = ((a[i][0]*b[0][j] + a[i][1]*b[1][j]) + a[i][2]*b[2][j]) + a[i][3]*b[3][j]
    a, b = mat(a), mat(b)
    out = np.zeros((4, 4), F)
    for i in range(4):
        for j in range(4):
            out[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j] + a[i][3] * b[3][j]
    return out


def mat4_inverted(m):               # matrix.go:255-307, derived rather than typed off: the cofactor expansion the Go code spells out
    """Every entry (r, c) of the source is  det * +-(m[s][k0]*A0 - m[s][k1]*A1 + m[s][k2]*A2)  where s = 1 for c == 0 and 0 otherwise,
    k0 < k1 < k2 are the columns other than r, each A is the 2x2 minor over the two remaining columns and the two rows other than
    c and s -- A(c0, c1, r0, r1) = m[r0][c0]*m[r1][c1] - m[r0][c1]*m[r1][c0], the source's Ac0c1r0r1 -- and the sign is (-1)^(r+c),
    applied as a negation before the multiplication by det.  det is the same expression for column 0, folded left to right."""
    m = mat(m)

    def minor(c0, c1, r0, r1):
        return m[r0][c0] * m[r1][c1] - m[r0][c1] * m[r1][c0]

    def inner(r, c):
        s_ = 1 if c == 0 else 0
        ks = [k for k in range(4) if k != r]
        rows = [q for q in range(4) if q not in (c, s_)]
        terms = []
        for k in ks:
            cols = [q for q in ks if q != k]
            terms.append(m[s_][k] * minor(cols[0], cols[1], rows[0], rows[1]))
        return terms[0] - terms[1] + terms[2]

    det = m[0][0] * inner(0, 0) - m[0][1] * inner(1, 0) + m[0][2] * inner(2, 0) - m[0][3] * inner(3, 0)
    det = F(1) / det
    out = np.zeros((4, 4), F)
    for r in range(4):
        for c in range(4):
            v = inner(r, c)
            out[r][c] = det * (v if (r + c) % 2 == 0 else -v)
    return out


def mat4_scale(x, y, z):            # matrix.go:90-96
    m = np.eye(4, dtype=F)
    m[0][0], m[1][1], m[2][2] = F(x), F(y), F(z)
    return m


def mat4_translate(x, y, z):        # matrix.go:81-87
    m = np.eye(4, dtype=F)
    m[3][0], m[3][1], m[3][2] = F(x), F(y), F(z)
    return m


def mat4_is_identity(m):            # matrix.go:511-529
    eps = F(0.0001)
    ident = np.eye(4, dtype=F)
    return not any(abs(m[i][j] - ident[i][j]) > eps for i in range(4) for j in range(4))


def node_transform(node):           # node.go:572-583: Scale x rotation x Translate [x parent.Transform()]
    t = mat4_scale(*[F(c) for c in node.scale])
    t = mat4_mult(t, mat(node.rotation))
    t = mat4_mult(t, mat4_translate(*[F(c) for c in node.position]))
    if node.parent is not None:
        t = mat4_mult(t, node_transform(node.parent))
    return t


def perspective(fovy, near, far, view_w, view_h):           # matrix.go:588-604; math32.Tan = float32(math.Tan(float64(x)))
    fovy, near, far = F(fovy), F(near), F(far)
    aspect = F(view_w) / F(view_h)
    t = F(math.tan(float(fovy * F(math.pi) / F(360))))
    b = -t
    r = t * aspect
    l = -r
    z = F(0)
    return np.array([[(F(2) * near) / (r - l), z, (r + l) / (r - l), z],
                     [z, (F(2) * near) / (t - b), (t + b) / (t - b), z],
                     [z, z, -((far + near) / (far - near)), -((F(2) * far * near) / (far - near))],
                     [z, z, F(-1), z]], F)


def go_round(x):                    # math32.Round = float32(math.Round(float64(x))): half away from zero
    x = float(x)
    if math.isnan(x) or math.isinf(x):
        return F(x)
    return F(math.copysign(math.floor(abs(x) + 0.5), x))


def orthographic(near, far, right, left, top, bottom):      # matrix.go:609-616
    near, far, right, left, top, bottom = F(near), F(far), F(right), F(left), F(top), F(bottom)
    m = np.zeros((4, 4), F)
    m[0][0] = F(2) / (right - left)
    m[1][1] = F(2) / (top - bottom)
    m[2][2] = F(-2) / (far - near)
    m[3][3] = F(1)
    return m


def transposed(m):                  # matrix.go:218-230
    return mat(m).T.copy()


def decompose_rotation(m):          # matrix.go:202-208
    r = np.eye(4, dtype=F)
    for i in range(3):
        r[i] = unit([m[i][0], m[i][1], m[i][2], m[i][3]])
    return r


def decompose(m):                   # matrix.go:181-196
    position = [m[3][0], m[3][1], m[3][2]]
    rotation = decompose_rotation(m)
    inn = mat4_mult(m, transposed(rotation))
    scale = [magnitude(inn[0]), magnitude(inn[1]), magnitude(inn[2])]
    return position, scale, rotation


def vec_equals(a, b):               # vector.go:415-425
    eps = F(1e-4)
    return not (abs(a[0] - b[0]) > eps or abs(a[1] - b[1]) > eps or abs(a[2] - b[2]) > eps)


def cross(v, o):                    # vector.go:164-179
    z = v[0] * o[1] - o[0] * v[1]
    y = v[2] * o[0] - o[2] * v[0]
    x = v[1] * o[2] - o[1] * v[2]
    return [x, y, z]


def look_at(frm, to, up):           # matrix.go:788-815
    if vec_equals(frm, to):
        return np.eye(4, dtype=F)
    z = unit([to[0] - frm[0], to[1] - frm[1], to[2] - frm[2]])
    up = unit(list(up))
    if vec_equals(z, up) or vec_equals(z, [-up[0], -up[1], -up[2]]):
        if not vec_equals(up, [F(1), F(0), F(0)]):
            up = [F(1), F(0), F(0)]             # WorldRight
        else:
            up = [F(0), F(0), F(1)]             # WorldBackward
    x = unit(cross(up, z))
    y = cross(z, x)
    return np.array([[x[0], x[1], x[2], 0], [y[0], y[1], y[2], 0], [z[0], z[1], z[2], 0], [0, 0, 0, 1]], F)


def up_vector(m):                   # matrix.go:161-167
    return unit([m[1][0], m[1][1], m[1][2]])


def forward(m):                     # matrix.go:170-176
    return unit([m[2][0], m[2][1], m[2][2]])


def world_rotation(node):           # node.go:916-938: decompose only when an ancestor is rotated
    parent = node.parent
    while parent is not None:
        if not mat4_is_identity(mat(parent.rotation)):
            return decompose_rotation(node_transform(node))
        parent = parent.parent
    return mat(node.rotation)


def world_position(node):           # node.go:737-739
    t = node_transform(node)
    return [t[3][0], t[3][1], t[3][2]]


# ---- sortingTriangles.go -----------------------------------------------------------------------------------------------
SORT_BACK_TO_FRONT, SORT_FRONT_TO_BACK, SORT_NONE = 0, 1, 2


class SortingTriangleBucket:
    def __init__(self, bin_count=512):
        self.bins = [[] for _ in range(bin_count)]
        self.unset = []
        self.sort_mode = SORT_BACK_TO_FRONT

    def clear(self):
        for b in self.bins:
            del b[:]
        del self.unset[:]

    def add_triangle(self, tri, depth):
        self.unset.append((tri, depth))

    def sort(self, min_range, max_range):                   # sortingTriangles.go:38-64
        bin_count = len(self.bins)
        range_diff = max_range - min_range
        if range_diff == 0:
            max_range = max_range + F(0.001)
            range_diff = range_diff + F(0.001)
        for tri, d in self.unset:
            target = 0
            if self.sort_mode != SORT_NONE and bin_count > 1:
                depth = (d - min_range) / range_diff * F(bin_count)
                t = clamp(depth, F(0), F(bin_count - 1))
                target = int(t)
            self.bins[target].append(tri)

    def for_each(self, fn):                                 # sortingTriangles.go:93-125
        if not self.unset:
            return
        order = range(len(self.bins) - 1, -1, -1) if self.sort_mode == SORT_BACK_TO_FRONT else range(len(self.bins))
        i = 0
        for b in order:
            for tri in self.bins[b]:
                fn(i, tri)
                i += 1


# ---- light.go ------------------------------------------------------------------------------------------------------------
LIGHTING_DEFAULT, LIGHTING_FIXED_NORMALS, LIGHTING_DOUBLE_SIDED = 0, 1, 2


class Globals:
    """tetra3d.go:21-40: the scratch arrays ProcessVertices fills and the lights / emit loop read."""

    def __init__(self, n):
        self.transforms = [[F(0)] * 4 for _ in range(n)]
        self.altered_positions = [[F(0)] * 3 for _ in range(n)]
        self.altered_normals = [[F(0)] * 3 for _ in range(n)]
        self.vertex_lights = [[F(0), F(0), F(0), F(1)] for _ in range(n)]
        self.depth_unbillboarded = [F(0)] * n


class AmbientLight:
    def __init__(self, color, energy):
        self.color, self.energy = [F(c) for c in color], F(energy)

    def begin_render(self):                                 # light.go:110-112
        self.result = [self.color[0] * self.energy, self.color[1] * self.energy, self.color[2] * self.energy]

    def begin_model(self, model):
        pass

    def vertex_light(self, index, tri, model, g):           # light.go:116-120
        c = g.vertex_lights[index]
        c[0] = c[0] + self.result[0]; c[1] = c[1] + self.result[1]; c[2] = c[2] + self.result[2]

    def light(self, tri, model, g):
        for vi in tri.indices:
            self.vertex_light(vi, tri, model, g)


class PointLight:
    def __init__(self, desc):
        self.desc = desc
        self.color, self.energy = [F(c) for c in desc.color], F(desc.energy)
        self.light_range = F(desc.range)
        self.light_range_squared = self.light_range * self.light_range

    def world_position(self):
        return world_position(self.desc.node)

    def begin_render(self):
        pass

    def begin_model(self, model):                           # light.go:177-191
        if model.skinned:
            self.working_position = self.world_position()
            return
        pos, sca, rot = decompose(mat4_inverted(model.transform))
        rp = mult_vec(rot, self.world_position())
        inv = [F(1) / sca[0], F(1) / sca[1], F(1) / sca[2]]
        self.working_position = [rp[0] + pos[0] * inv[0], rp[1] + pos[1] * inv[1], rp[2] + pos[2] * inv[2]]

    def vertex_light(self, index, tri, model, g):           # light.go:193-258
        vert_pos, vert_normal = g.altered_positions[index], g.altered_normals[index]
        distance = distance_squared_to(self.working_position, vert_pos)
        if self.light_range > 0:
            if distance > self.light_range_squared:
                return
        else:
            if F(1) / distance * self.energy < F(0.001):
                return
        light_vec = unit([self.working_position[0] - vert_pos[0], self.working_position[1] - vert_pos[1], self.working_position[2] - vert_pos[2]])
        mat_ = tri.part.material
        if mat_ is not None and mat_.lighting_mode == LIGHTING_FIXED_NORMALS:
            vert_normal = light_vec
        diffuse = dot(vert_normal, light_vec)
        if mat_ is not None and mat_.lighting_mode == LIGHTING_DOUBLE_SIDED:
            diffuse = F(abs(diffuse))
        if diffuse > 0:
            diffuse_factor = diffuse * (F(1.0) / (F(1.0) + (F(0.1) * distance))) * F(2)
            if self.light_range_squared > 0:
                dist_clamp = clamp((self.light_range_squared - distance) / distance, F(0), F(1))
                diffuse_factor = diffuse_factor * dist_clamp
            c = g.vertex_lights[index]
            c[0] = c[0] + self.color[0] * diffuse_factor * self.energy
            c[1] = c[1] + self.color[1] * diffuse_factor * self.energy
            c[2] = c[2] + self.color[2] * diffuse_factor * self.energy

    def light(self, tri, model, g):
        for vi in tri.indices:
            self.vertex_light(vi, tri, model, g)


class DirectionalLight:
    def __init__(self, desc):
        self.desc = desc
        self.color, self.energy = [F(c) for c in desc.color], F(desc.energy)

    def begin_render(self):                                 # light.go:322-324
        self.working_forward = forward(world_rotation(self.desc.node))

    def begin_model(self, model):                           # light.go:326-330
        if not model.skinned:
            self.working_model_rotation = transposed(mat4_inverted(model.world_rotation))

    def vertex_light(self, index, tri, model, g):           # light.go:332-360
        mat_ = tri.part.material
        if mat_ is not None and mat_.lighting_mode == LIGHTING_FIXED_NORMALS:
            normal = self.working_forward
        elif model.skinned:
            normal = g.altered_normals[index]
        else:
            normal = mult_vec(self.working_model_rotation, g.altered_normals[index])
        diffuse_factor = dot(normal, self.working_forward)
        if mat_ is not None and mat_.lighting_mode == LIGHTING_DOUBLE_SIDED:
            diffuse_factor = F(abs(diffuse_factor))
        if diffuse_factor <= 0:
            return
        c = g.vertex_lights[index]
        c[0] = c[0] + self.color[0] * diffuse_factor * self.energy
        c[1] = c[1] + self.color[1] * diffuse_factor * self.energy
        c[2] = c[2] + self.color[2] * diffuse_factor * self.energy

    def light(self, tri, model, g):
        for vi in tri.indices:
            self.vertex_light(vi, tri, model, g)


LV_PER_VERTEX_WITHOUT_NORMAL, LV_PER_VERTEX_WITH_NORMAL, LV_PER_OBJECT, LV_OFF = 0, 1, 2, 3


class LightVolume:
    """light.go:376-396, 567-680, 717-728, 1267-1272: a grid of cell colours in world space; cells may carry lights of their own."""

    def __init__(self, desc):
        v = desc.volume
        self.color, self.energy = [F(c) for c in desc.color], F(desc.energy)
        self.dim_min, self.dim_max = [F(c) for c in v.dim_min], [F(c) for c in v.dim_max]
        self.cell_size = [F(c) for c in v.cell_size]
        self.cell_count = [int(c) for c in v.cell_count]
        self.colors = np.asarray(v.colors, F).reshape(-1, 3)
        self.offset, self.light_bias = [F(c) for c in v.offset], [F(c) for c in v.light_bias]
        self.cell_lights = {}
        self.use_lights_render_set = []
        for key, lights in (v.cell_lights or {}).items():
            objs = [make_light(ld) for ld in lights]
            self.cell_lights[tuple(int(k) for k in key)] = objs
            self.use_lights_render_set.extend(objs)

    def begin_render(self):
        for l in self.use_lights_render_set:
            l.begin_render()

    def begin_model(self, model):
        for l in self.use_lights_render_set:
            l.begin_model(model)

    def convert_to_xyz_indices(self, w):                    # light.go:717-724
        out = []
        for k in range(3):
            out.append(int(go_round((w[k] - self.offset[k] - self.dim_min[k] - (self.cell_size[k] / F(2))) / self.cell_size[k])))
        return out

    def get_light_cell(self, x, y, z):                      # light.go:1267-1272: (colour, lights) or None
        cx, cy, cz = self.cell_count
        if x < 0 or y < 0 or z < 0 or x >= cx or y >= cy or z >= cz:
            return None
        return self.colors[x + (y * cx) + (z * cx * cy)], self.cell_lights.get((x, y, z), [])

    def contains(self, p):                                  # mesh.go:82-84
        return all(p[k] >= self.dim_min[k] for k in range(3)) and all(p[k] <= self.dim_max[k] for k in range(3))

    def vertex_light(self, vi, tri, model, g):              # light.go:583-663
        object_shading_mode = 0
        cell_color = [F(0), F(0), F(0)]
        transform = model.transform
        cell_size = [self.cell_size[0] / F(2), self.cell_size[1] / F(2), self.cell_size[2] / F(2)]
        mat_ = tri.part.material
        if mat_ is not None:
            object_shading_mode = mat_.light_volume_shading_mode
        if mat_ is not None and mat_.light_volume_shading_mode == LV_PER_OBJECT:
            x, y, z = self.convert_to_xyz_indices([model.transform[3][0], model.transform[3][1], model.transform[3][2]])
            cell = self.get_light_cell(x, y, z)
            if cell is not None:
                cell_color = [cell[0][0], cell[0][1], cell[0][2]]
        else:
            vert_pos = [F(0), F(0), F(0)]
            vert_pos_with_normal = [F(0), F(0), F(0)]
            ap, an = g.altered_positions[vi], g.altered_normals[vi]
            if model.skinned:
                if object_shading_mode == LV_PER_VERTEX_WITH_NORMAL:
                    vert_pos_with_normal = [vert_pos[k] + an[k] * cell_size[k] + self.light_bias[k] for k in range(3)]
                else:
                    vert_pos = [ap[k] + self.light_bias[k] for k in range(3)]
            else:
                if object_shading_mode == LV_PER_VERTEX_WITH_NORMAL:
                    t = mult_vec(transform, [ap[k] + an[k] * cell_size[k] for k in range(3)])
                    vert_pos_with_normal = [t[k] + self.light_bias[k] for k in range(3)]
                else:
                    t = mult_vec(transform, ap)
                    vert_pos = [t[k] + self.light_bias[k] for k in range(3)]
            if not self.contains(vert_pos):                 # (the unshifted position even in the with-normal mode, where it is still zero)
                return
            cell_color = [F(0), F(0), F(0)]
            if object_shading_mode == LV_PER_VERTEX_WITH_NORMAL:
                x, y, z = self.convert_to_xyz_indices(vert_pos_with_normal)
            else:
                x, y, z = self.convert_to_xyz_indices(vert_pos)
            cell = self.get_light_cell(x, y, z)
            if cell is not None:
                cell_color = [cell[0][0], cell[0][1], cell[0][2]]
                for l in cell[1]:
                    l.vertex_light(vi, tri, model, g)
        c = g.vertex_lights[vi]
        c[0] = c[0] + self.color[0] * cell_color[0] * self.energy
        c[1] = c[1] + self.color[1] * cell_color[1] * self.energy
        c[2] = c[2] + self.color[2] * cell_color[2] * self.energy

    def light(self, tri, model, g):                         # light.go:667-680
        mat_ = tri.part.material
        if mat_ is not None and mat_.light_volume_shading_mode == LV_OFF:
            return
        for vi in tri.indices:
            self.vertex_light(vi, tri, model, g)


def make_light(desc):
    if desc.type == 0:
        return AmbientLight(desc.color, desc.energy)
    if desc.type == 1:
        return PointLight(desc)
    if desc.type == 2:
        return DirectionalLight(desc)
    return LightVolume(desc)


# ---- mesh.go / model.go ----------------------------------------------------------------------------------------------------
class Triangle:
    def __init__(self, tid, indices, center, part):
        self.id, self.indices, self.center, self.part = tid, [int(i) for i in indices], [F(c) for c in center], part
        self.visible = True
        self.is_base = True
        self.subdivision_levels = []        # base triangles: [[Triangle, ...] per level] (mesh.go:1758-1864 builds them at load time)
        self.normal, self.max_span = None, None


# ---- colliderTriangles.go:177-182, 206-281 (closestPointOnTri, as Triangle.handleSubdivision uses it) ----------------------
def sub(a, b):
    return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]


def closest_point_on_line(point, start, end):
    diff = sub(end, start)
    dot_a = dot(sub(point, start), diff)
    dot_b = dot(diff, diff)
    d = dot_a / dot_b
    if d > 1:
        d = F(1)
    elif d < 0:
        d = F(0)
    return [start[0] + diff[0] * d, start[1] + diff[1] * d, start[2] + diff[2] * d]


def point_inside_triangle(point, v0, v1, v2):
    ca, ba, pa = sub(v2, v0), sub(v1, v0), sub(point, v0)
    dot00, dot01, dot02 = dot(ca, ca), dot(ca, ba), dot(ca, pa)
    dot11, dot12 = dot(ba, ba), dot(ba, pa)
    inv_denom = F(1.0) / ((dot00 * dot11) - (dot01 * dot01))
    u = ((dot11 * dot02) - (dot01 * dot12)) * inv_denom
    v = ((dot00 * dot12) - (dot01 * dot02)) * inv_denom
    return u, v


def closest_point_on_tri(point, v0, v1, v2, normal):
    distance = dot(normal, v0)
    dist = dot(normal, point) - distance
    plane_point = [point[0] - normal[0] * dist, point[1] - normal[1] * dist, point[2] - normal[2] * dist]
    u, v = point_inside_triangle(plane_point, v0, v1, v2)
    if u >= 0 and v >= 0 and u + v < 1:
        return plane_point
    ab = closest_point_on_line(point, v0, v1)
    bc = closest_point_on_line(point, v1, v2)
    ca = closest_point_on_line(point, v2, v0)
    closest, closest_dist = ab, distance_squared_to(point, ab)
    bc_dist, ca_dist = distance_squared_to(point, bc), distance_squared_to(point, ca)
    if bc_dist < closest_dist:
        closest, closest_dist = bc, bc_dist
    if ca_dist < closest_dist:
        closest = ca
    return closest


def disable_subdivision(t):                                 # mesh.go:1866-1875
    if not t.visible:
        t.visible = True
        for level in t.subdivision_levels:
            for s_ in level:
                s_.visible = False


def handle_subdivision(t, camera_pos, mesh, levels):        # mesh.go:1877-1935
    t.visible = False
    if len(t.subdivision_levels) == 0:
        t.visible = True
        return
    for level in t.subdivision_levels:
        for s_ in level:
            s_.visible = False
    active, closest_distance = -1, MAXF
    P = mesh.positions
    a, b, c = t.indices
    dist = distance_squared_to(camera_pos, closest_point_on_tri(camera_pos, list(P[a]), list(P[b]), list(P[c]), t.normal))
    for i, (distance_squared, _max_size) in enumerate(levels):
        if dist < F(distance_squared) and F(distance_squared) < closest_distance:
            closest_distance = F(distance_squared)
            active = i
    if active == -1:
        t.visible = True
    else:
        ms, sub_index = t.max_span, -1
        while ms > F(levels[active][1]):
            ms = ms / F(2)
            sub_index += 1
        if sub_index >= len(t.subdivision_levels):
            sub_index = len(t.subdivision_levels) - 1
        if sub_index >= 0:
            for s_ in t.subdivision_levels[sub_index]:
                s_.visible = True
        else:
            t.visible = True


class Model:
    def __init__(self, desc):
        self.desc = desc
        self.transform = node_transform(desc.node)
        self.world_rotation = world_rotation(desc.node)
        self.color = [F(c) for c in desc.color]
        mesh = desc.mesh
        centers = mesh.tri_centers if mesh.auto_subdivide else O.triangle_centers(mesh.positions, mesh.indices)
        self.parts = []
        for p in mesh.parts:
            tris = [Triangle(t, mesh.indices[t], centers[t], p) for t in range(p.tri_start, p.tri_end + 1)]
            if mesh.auto_subdivide:     # the entries behind a base triangle are its sub-triangles, level by level (MeshPart.forEachTri(true, ..))
                by_slot = {tr.id: tr for tr in tris}
                for tr in tris:
                    lvl = int(mesh.tri_sub_level[tr.id])
                    if lvl < 0:
                        tr.normal = [F(x) for x in mesh.tri_normal[tr.id]]
                        tr.max_span = F(mesh.tri_max_span[tr.id])
                        tr.subdivision_levels = [[] for _ in range(int(mesh.tri_n_levels[tr.id]))]
                    else:
                        tr.is_base = False
                        tr.visible = False
                        by_slot[int(mesh.tri_sub_base[tr.id])].subdivision_levels[lvl].append(tr)
            self.parts.append(tris)
        # skinned models (model.go:33-40): boneInfluence of every joint, given directly or posed from a skeleton + animation
        self.bone_influences = None
        if desc.bone_influences is not None:
            self.bone_influences = np.asarray(desc.bone_influences, F).reshape(-1, 4, 4)
        elif desc.pose is not None:
            p = desc.pose
            if p.blend_ms is not None:
                bi = O.pose_bones_blended(p.skeleton, p.blend_from, p.from_playhead, p.animation, p.playhead, p.blend_fraction(), O.node_world_transform(p.root))
            else:
                bi = O.pose_bones(p.skeleton, p.animation, p.playhead, O.node_world_transform(p.root))
            self.bone_influences = bi.reshape(-1, 4, 4)
        self.skinned = self.bone_influences is not None
        lo, hi = mesh.positions.min(axis=0), mesh.positions.max(axis=0)         # Dimensions (mesh.go:718-752): bounds of the vertices
        self.max_span = magnitude([hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]])

    def is_transparent(self, part):                         # model.go:1274-1284 (TransparencyMode: 0 auto, 1 opaque, 2 alpha clip, 3 transparent)
        m = part.material
        if m is not None:
            mat_transparent = m.transparency_mode == 3 or not m.blend_source_over or (m.transparency_mode == 0 and F(m.color[3]) < F(0.999))
            model_transparent = m.transparency_mode != 1 and self.color[3] < F(0.999)
            return mat_transparent or model_transparent
        return self.color[3] < F(0.999)


def vertex_position_with_shape_keys(mesh, vi):             # mesh.go:796-807
    out = [mesh.positions[vi][0], mesh.positions[vi][1], mesh.positions[vi][2]]
    if mesh.key_weights is not None and len(mesh.key_weights) > 0:
        for k in range(len(mesh.key_weights)):
            w = mesh.key_weights[k]
            out[0] = out[0] + mesh.key_positions[k][vi][0] * w
            out[1] = out[1] + mesh.key_positions[k][vi][1] * w
            out[2] = out[2] + mesh.key_positions[k][vi][2] * w
    return out


def vertex_normal_with_shape_keys(mesh, vi):               # mesh.go:811-822
    out = [mesh.normals[vi][0], mesh.normals[vi][1], mesh.normals[vi][2]]
    if mesh.key_weights is not None and len(mesh.key_weights) > 0:
        for k in range(len(mesh.key_weights)):
            w = mesh.key_weights[k]
            out[0] = out[0] + mesh.key_normals[k][vi][0] * w
            out[1] = out[1] + mesh.key_normals[k][vi][1] * w
            out[2] = out[2] + mesh.key_normals[k][vi][2] * w
    return out


def skin_vertex(model, vi):                                 # model.go:559-596
    mesh = model.desc.mesh
    skin = np.zeros((4, 4), F)                              # skinMatrix.Clear()
    for b in range(int(mesh.n_influences[vi])):             # model.bones[vertID]: one entry per influence the loader kept
        weight = F(mesh.weights[vi][b])
        if weight == 0:
            continue
        influence = model.bone_influences[int(mesh.bones[vi][b])]
        if weight == 1:
            skin = influence.copy()
            break
        skin = skin + influence * weight                    # Add(Scale(weight)): element by element
    vert_out = mult_vec(skin, vertex_position_with_shape_keys(mesh, vi))
    skin = skin.copy()
    skin[3][0] = F(0); skin[3][1] = F(0); skin[3][2] = F(0); skin[3][3] = F(1)
    normal = mult_vec(skin, vertex_normal_with_shape_keys(mesh, vi))
    return vert_out, normal


class Camera:
    def __init__(self, desc):
        self.desc = desc
        self.near, self.far = F(desc.near), F(desc.far)
        self.position = world_position(desc.node)
        cam_pos = [-self.position[0], -self.position[1], -self.position[2]]                    # camera.go:638-648
        self.view = mat4_mult(mat4_translate(*cam_pos), transposed(world_rotation(desc.node)))
        if desc.perspective:                                                                   # camera.go:651-680
            self.projection = perspective(desc.field_of_view, desc.near, desc.far, desc.width, desc.height)
        else:
            asr = F(desc.height) / F(desc.width)
            s = F(desc.ortho_scale)
            self.projection = orthographic(desc.near, desc.far, F(1) * s, F(-1) * s, asr * s, -asr * s)
        self.vp = mat4_mult(self.view, self.projection)                                      # camera.go:1329
        self.camera_up = up_vector(world_rotation(desc.node))                                  # camera.go:1085-1088 (Camera.Clear)


def process_vertices(model, part_index, camera, bucket, g, autosubdivision_levels=(), process_only_visible=True):
    """model.go:600-995 with processOnlyVisible = true, for a model without billboarding or callbacks.
    Returns False for the early-out of model.go:604-606."""
    bucket.clear()
    part = model.desc.mesh.parts[part_index]
    if process_only_visible and ((model.color[3] == 0 and model.is_transparent(part)) or not model.desc.visible):
        return False
    mesh = model.desc.mesh
    model_transform = model.transform
    cam_pos = camera.position
    p, s, r = decompose(mat4_inverted(model_transform))                                      # model.go:627-628
    rc = mult_vec(r, cam_pos)
    inv = [F(1) / s[0], F(1) / s[1], F(1) / s[2]]
    inverted_cam_pos = [rc[0] + p[0] * inv[0], rc[1] + p[1] * inv[1], rc[2] + p[2] * inv[2]]
    mat_ = part.material
    base = model_transform
    if mat_ is not None and mat_.billboard_enabled:                                            # model.go:634-695
        up = [F(0), F(1), F(0)]
        if mat_.billboard_up_direction == 1:                                                   # BillboardYUpCameraY
            up = camera.camera_up
        model_world_position = [model_transform[3][0], model_transform[3][1], model_transform[3][2]]
        lookat = look_at(model_world_position, cam_pos, up)
        if mat_.billboard_lock_x:
            lookat[0] = [1, 0, 0, 0]
            row = list(lookat[1]); row[0] = F(0); lookat[1] = unit(row)
            row = list(lookat[2]); row[0] = F(0); lookat[2] = unit(row)
        if mat_.billboard_lock_y:
            lookat[1] = [0, 1, 0, 0]
            row = list(lookat[0]); row[1] = F(0); lookat[0] = unit(row)
            row = list(lookat[2]); row[1] = F(0); lookat[2] = unit(row)
        if mat_.billboard_lock_z:
            lookat[2] = [0, 0, 1, 0]
            row = list(lookat[0]); row[2] = F(0); lookat[0] = unit(row)
            row = list(lookat[1]); row[2] = F(0); lookat[1] = unit(row)
        bp, bs, br = decompose(base)
        base = mat4_mult(mat4_mult(br, mat4_scale(bs[0], bs[1], bs[2])), lookat)
        base[3] = [bp[0], bp[1], bp[2], F(1)]
    mvp = mat4_mult(base, camera.vp)
    unbillboarded = mat_ is not None and mat_.billboarded_depth_mode == 1                      # model.go:699-705
    unaltered_mvp = mat4_mult(model_transform, camera.vp) if unbillboarded else None
    min_depth, max_depth = MAXF, -MAXF
    snapping = F(camera.desc.vertex_snapping)
    if not model.skinned:                                                                      # model.go:727-741
        for tri in model.parts[part_index]:
            if not tri.is_base:
                continue
            if mesh.auto_subdivide and len(autosubdivision_levels) > 0:
                handle_subdivision(tri, inverted_cam_pos, mesh, autosubdivision_levels)
            else:
                disable_subdivision(tri)
    for tri in model.parts[part_index]:
        if mesh.auto_subdivide and not tri.visible:                                            # model.go:753-755
            continue
        for vi in tri.indices:
            if model.skinned:                                                                  # model.go:760-776: world space, vpMatrix
                vert, normal = skin_vertex(model, vi)
                m = camera.vp
            else:                                                                              # model.go:778-800
                vert = vertex_position_with_shape_keys(mesh, vi)
                normal = vertex_normal_with_shape_keys(mesh, vi)
                m = mvp
            g.altered_positions[vi] = vert
            g.altered_normals[vi] = normal
            t = g.transforms[vi]
            t[0] = m[0][0] * vert[0] + m[1][0] * vert[1] + m[2][0] * vert[2] + m[3][0]
            t[1] = m[0][1] * vert[0] + m[1][1] * vert[1] + m[2][1] * vert[2] + m[3][1]
            t[2] = m[0][2] * vert[0] + m[1][2] * vert[1] + m[2][2] * vert[2] + m[3][2]
            t[3] = m[0][3] * vert[0] + m[1][3] * vert[1] + m[2][3] * vert[2] + m[3][3]
            if unbillboarded and not model.skinned:                                            # model.go:802-805
                u = unaltered_mvp
                g.depth_unbillboarded[vi] = u[0][2] * vert[0] + u[1][2] * vert[1] + u[2][2] * vert[2] + u[3][2]
            if snapping > 0:                                                                   # model.go:812-817
                for k in range(3):
                    t[k] = go_round(t[k] * snapping) / snapping
        out_of_bounds = True
        tv = []
        for vi in tri.indices:
            w = g.transforms[vi][3]
            if w < 0:
                w = F(0.000001)
            tv.append((g.transforms[vi][0] / w, g.transforms[vi][1] / w))
            if g.transforms[vi][2] + F(1) >= camera.near and g.transforms[vi][2] < camera.far:
                out_of_bounds = False
        if process_only_visible and out_of_bounds:
            continue
        h = F(0.5)
        if process_only_visible and ((tv[0][0] < -h and tv[1][0] < -h and tv[2][0] < -h) or (tv[0][0] > h and tv[1][0] > h and tv[2][0] > h) or \
           (tv[0][1] < -h and tv[1][1] < -h and tv[2][1] < -h) or (tv[0][1] > h and tv[1][1] > h and tv[2][1] > h)):
            continue
        if mat_ is not None and mat_.backface_culling:                                         # model.go:893-916
            n0x, n0y = tv[0][0] - tv[1][0], tv[0][1] - tv[1][1]
            n1x, n1y = tv[1][0] - tv[2][0], tv[1][1] - tv[2][1]
            nor = (n0x * n1y) - (n1x * n0y)
            if process_only_visible and nor < 0:
                continue
        if model.skinned:                                                                      # model.go:839-856, 934-940
            cx = cy = cz = F(0)
            for vi in tri.indices:
                cx = cx + g.altered_positions[vi][0]; cy = cy + g.altered_positions[vi][1]; cz = cz + g.altered_positions[vi][2]
            dx, dy, dz = cam_pos[0] - (cx / F(3)), cam_pos[1] - (cy / F(3)), cam_pos[2] - (cz / F(3))
        else:                                                                                  # model.go:942-948
            dx, dy, dz = inverted_cam_pos[0] - tri.center[0], inverted_cam_pos[1] - tri.center[1], inverted_cam_pos[2] - tri.center[2]
        depth = dx * dx + dy * dy + dz * dz
        if process_only_visible and (np.isnan(depth) or np.isinf(depth)):
            continue
        if depth < min_depth:
            min_depth = depth
        if depth > max_depth:
            max_depth = depth
        bucket.add_triangle(tri, depth)
    bucket.sort(min_depth, max_depth)
    return True


def render_scene(desc):
    """Camera.Render's light set-up (camera.go:1275-1303) and the render closure (camera.go:1507-1827) for every visible
    (model, part); returns {(model index, part index): (triangle ids in emitted order, colour records (n, 12), depth records (n, 12))}."""
    cam = Camera(desc.camera)
    world = desc.world
    scene_lights = []
    if world.present and world.lighting_on and world.ambient_visible:
        amb = AmbientLight(world.ambient_color, world.ambient_energy)
        amb.begin_render()
        scene_lights.append(amb)
    for ld in desc.lights:
        if not ld.in_scene:
            continue
        if (not world.present or world.lighting_on) and ld.visible:
            l = make_light(ld)
            l.begin_render()
            scene_lights.append(l)
    if desc.camera.max_light_count > 0:                                                        # camera.go:1474-1497
        focus = desc.camera.max_light_focus
        pos = world_position(focus) if focus is not None else cam.position

        def light_pos(l):
            return l.world_position() if isinstance(l, PointLight) else (world_position(l.desc.node) if hasattr(l, "desc") else [F(0), F(0), F(0)])

        def less_light(i, j):
            if isinstance(i, AmbientLight):
                return True
            elif isinstance(j, AmbientLight):
                return False
            elif isinstance(i, DirectionalLight):
                return True
            elif isinstance(j, DirectionalLight):
                return False
            return distance_squared_to(pos, light_pos(i)) < distance_squared_to(pos, light_pos(j))

        scene_lights = slice_stable(scene_lights, less_light)
        scene_lights = scene_lights[:min(desc.camera.max_light_count, len(scene_lights))]
    cam_w, cam_h = desc.camera.width, desc.camera.height
    depth_margin = (cam.far - cam.near) * F(desc.camera.depth_margin)                          # camera.go:1339
    cam_spread = cam.far - cam.near + (depth_margin * F(2))                                    # camera.go:1341
    bucket = SortingTriangleBucket()
    out = {}
    for mi, md in enumerate(desc.models):
        if not md.visible:
            continue
        model = Model(md)
        mesh = md.mesh
        g = Globals(mesh.n_verts)
        for pi, part in enumerate(mesh.parts):
            mat_ = part.material
            if mat_ is not None and not mat_.visible:
                continue
            lighting = False                                                                   # camera.go:1521-1528
            if world.present:
                if mat_ is not None:
                    lighting = world.lighting_on and not mat_.shadeless and not md.shadeless
                else:
                    lighting = world.lighting_on and not md.shadeless
            bucket.sort_mode = SORT_BACK_TO_FRONT if mat_ is None else mat_.triangle_sort_mode
            if not process_vertices(model, pi, cam, bucket, g, desc.autosubdivision_levels):
                out[(mi, pi)] = ([], np.zeros((0, 12), F), np.zeros((0, 12), F))
                continue
            src_w = src_h = F(0)
            if mat_ is not None and mat_.use_texture and mat_.tex_w > 0:
                src_w, src_h = F(mat_.tex_w), F(mat_.tex_h)
            lights = scene_lights
            if lighting:
                if md.light_group is not None and md.light_group.active:                       # camera.go:1573-1579
                    lights = [make_light(ld) for ld in md.light_group.lights]
                    for l in lights:
                        l.begin_render()
                for l in lights:
                    l.begin_model(model)
            model_pos = [model.transform[3][0], model.transform[3][1], model.transform[3][2]]
            mp_color = list(model.color)                                                       # camera.go:1597-1601
            if mat_ is not None:
                mp_color = [mp_color[k] * F(mat_.color[k]) for k in range(4)]
            half_w, half_h = F(cam_w) / F(2), F(cam_h) / F(2)
            if lighting:                                                                       # camera.go:1610-1641

                def light_tri(i, tri):
                    for vi in tri.indices:
                        g.vertex_lights[vi] = [F(0), F(0), F(0), F(1)]
                    for l in lights:
                        if isinstance(l, PointLight) and l.light_range > 0:
                            dist = model.max_span + l.light_range
                            if distance_squared_to(model_pos, l.world_position()) > dist * dist:
                                continue
                        l.light(tri, model, g)

                bucket.for_each(light_tri)
            tri_ids, col_rows, dep_rows = [], [], []

            def emit_tri(i, tri):                                                              # camera.go:1646-1820
                tri_ids.append(tri.id)
                for vi in tri.indices:
                    w = g.transforms[vi][3]
                    if not desc.camera.perspective:
                        w = F(1.0)
                    if w < 0:
                        w = F(0.001)
                    target = g.transforms[vi]
                    dx = (target[0] / w) * F(cam_w) + half_w
                    dy = (target[1] / -w) * F(cam_h) + half_h
                    custom0 = F(0)
                    if desc.camera.perspective_corrected_texture_mapping:
                        uv_u = (mesh.uvs[vi][0] / w) * src_w
                        uv_v = ((F(1) - mesh.uvs[vi][1]) / w) * src_h
                        custom0 = F(1.0) / w
                    else:
                        uv_u = mesh.uvs[vi][0] * src_w
                        uv_v = (F(1) - mesh.uvs[vi][1]) * src_h
                    if mesh.active_color_channel >= 0 and mesh.colors is not None:
                        col = [mesh.colors[vi][k] * mp_color[k] for k in range(4)]
                    else:
                        col = list(mp_color)
                    if lighting:
                        col[0] = col[0] * g.vertex_lights[vi][0]
                        col[1] = col[1] * g.vertex_lights[vi][1]
                        col[2] = col[2] * g.vertex_lights[vi][2]
                    dcol = [F(0)] * 4
                    if desc.camera.render_depth:                                               # camera.go:1744-1782
                        if mat_ is not None and mat_.billboarded_depth_mode == 1:
                            vertex_depth = g.depth_unbillboarded[vi]
                        else:
                            vertex_depth = g.transforms[vi][2]
                        depth = (vertex_depth + depth_margin) / cam_spread
                        if depth < 0:
                            depth = F(0)
                        elif depth > 1:
                            depth = F(1)
                        dcol = [depth, depth, depth, F(1)]
                    elif world.present and world.fog_on:                                       # camera.go:1784-1814
                        depth = (g.transforms[vi][2] + depth_margin) / cam_spread
                        if depth < 0:
                            depth = F(0)
                        elif depth > 1:
                            depth = F(1)
                        depth = F(world.fog_range[0]) + ((F(world.fog_range[1]) - F(world.fog_range[0])) * F(1) - depth)
                        fc = [F(c) for c in world.fog_color]
                        if world.fog_mode == 0:
                            col[0] = col[0] + fc[0] * depth; col[1] = col[1] + fc[1] * depth; col[2] = col[2] + fc[2] * depth
                        elif world.fog_mode == 1:
                            col[0] = col[0] * (fc[0] * depth); col[1] = col[1] * (fc[1] * depth); col[2] = col[2] * (fc[2] * depth)
                    col_rows.append([dx, dy, uv_u, uv_v] + col + [custom0, F(0), F(0), F(0)])
                    dep_rows.append([dx, dy, uv_u, uv_v] + dcol + [custom0, F(0), F(0), F(0)])

            bucket.for_each(emit_tri)
            out[(mi, pi)] = (tri_ids, np.asarray(col_rows, F).reshape(-1, 12), np.asarray(dep_rows, F).reshape(-1, 12))
    return out


# ---- camera.go:933-989, model.go:142-182, camera.go:1343-1469, 2050-2055: which render pairs there are, and in which order ----------
def frustum_sphere(model):                                  # Model.onTransformUpdate (model.go:142-182)
    position, scale, rotation = decompose(model.transform)
    mesh = model.desc.mesh
    lo, hi = mesh.positions.min(axis=0), mesh.positions.max(axis=0)
    center = [F(0), F(0), F(0)]
    if not model.skinned:
        center = [(hi[0] + lo[0]) / F(2), (hi[1] + lo[1]) / F(2), (hi[2] + lo[2]) / F(2)]           # Dimensions.Center (mesh.go:46-52)
    center = mult_vec(rotation, center)
    position = [position[0] + center[0] * scale[0], position[1] + center[1] * scale[1], position[2] + center[2] * scale[2]]
    mn = [lo[0] * scale[0], lo[1] * scale[1], lo[2] * scale[2]]
    mx = [hi[0] * scale[0], hi[1] * scale[1], hi[2] * scale[2]]
    radius = magnitude([mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]]) / F(2)
    return position, radius


def sphere_in_frustum(cam, position, radius):               # camera.go:933-983
    d_ = cam.desc
    rot = world_rotation(d_.node)
    cam_forward, cam_up = forward(rot), up_vector(rot)
    cam_right = unit([rot[0][0], rot[0][1], rot[0][2]])                                             # Matrix4.Right
    diff = sub(position, cam.position)
    pc_z = dot(diff, [-cam_forward[0], -cam_forward[1], -cam_forward[2]])
    if pc_z > cam.far + radius or pc_z < cam.near - radius:
        return False
    aspect = F(d_.width) / F(d_.height)
    if d_.perspective:
        angle = F(d_.field_of_view) * F(3.1415) / F(360)                                            # camera.go:660-666
        tang = F(math.tan(float(angle)))
        factor_y = F(1.0) / F(math.cos(float(angle)))
        factor_x = F(1.0) / F(math.cos(float(F(math.atan(float(tang * aspect))))))
        d = factor_y * radius
        pc_z = pc_z * tang
        pc_y = dot(diff, cam_up)
        if pc_y > pc_z + d or pc_y < -pc_z - d:
            return False
        pc_z = pc_z * aspect
        d = factor_x * radius
        pc_x = dot(diff, cam_right)
        if pc_x > pc_z + d or pc_x < -pc_z - d:
            return False
    else:
        width = F(d_.ortho_scale)
        height = width / aspect
        pc_y = dot(diff, cam_up)
        if -height / F(2) - radius > pc_y or pc_y > height / F(2) + radius:
            return False
        pc_x = dot(diff, cam_right)
        if -width / F(2) - radius > pc_x or pc_x > width / F(2) + radius:
            return False
    return True


def render_order(desc):
    """The (model index, part index) of every render(rp) call in the order Camera.Render makes them.  Dynamic batches
    (camera.go:1398-1453, 2075-2100): the owner's MeshPart groups are visited in ascending part order (a Go map in the reference:
    the contract fixes the order), the children far to near (sort.Slice; distinct distances assumed), each child's frustum test
    applied, every MeshPart of a child rendered."""
    cam = Camera(desc.camera)
    solids, transparents = [], []
    models = [Model(md) for md in desc.models]
    children = {}
    for ci, md in enumerate(desc.models):
        if md.batch_owner is not None:
            oi = [k for k, o in enumerate(desc.models) if o is md.batch_owner][0]
            children.setdefault((oi, md.batch_owner_part), []).append(ci)

    def in_frustum(model):
        pos, radius = frustum_sphere(model)
        return sphere_in_frustum(cam, pos, radius)

    for mi, md in enumerate(desc.models):
        if not md.visible or md.batch_owner is not None:
            continue
        model = models[mi]
        model_pos = [model.transform[3][0], model.transform[3][1], model.transform[3][2]]
        batcher = any(k[0] == mi for k in children)
        if not batcher:
            if md.frustum_culling and not in_frustum(model):
                continue
            for pi, part in enumerate(md.mesh.parts):
                mat_ = part.material
                if mat_ is not None and not mat_.visible:
                    continue
                depth = distance_squared_to(cam.position, model_pos)
                order = mat_.render_order if mat_ is not None else 0
                (transparents if model.is_transparent(part) else solids).append((order, depth, mi, pi, None))
        else:
            transparent = False
            for pi in sorted(k[1] for k in children if k[0] == mi):
                part = md.mesh.parts[pi]
                if part.material is not None and not part.material.visible:
                    continue
                for ci in children[(mi, pi)]:
                    cd = desc.models[ci]
                    if not cd.visible:
                        continue
                    if not transparent:
                        for cp in cd.mesh.parts:
                            if (cp.material is None or cp.material.visible) and models[ci].is_transparent(cp):
                                transparent = True
                                break
                depth = F(0)
                if transparent or not desc.camera.render_depth:
                    depth = distance_squared_to(cam.position, model_pos)
                order = part.material.render_order if part.material is not None else 0
                (transparents if transparent else solids).append((order, depth, mi, pi, children[(mi, pi)]))
    if not desc.camera.render_depth:                                                                 # camera.go:1460-1469 (not an ordering when orders differ:
        def less(i, j):                                                                              #  the outcome is whatever sort.SliceStable's algorithm makes of it)
            if i[0] < j[0]:
                return True
            return i[1] > j[1]
        solids = slice_stable(solids, less)
    transparents = sorted(transparents, key=lambda r: (r[0], -float(r[1])))                          # camera.go:2050-2055
    out = []
    for _, _, mi, pi, kids in solids + transparents:
        if kids is None:
            out.append((mi, pi))
            continue
        dist = {ci: distance_squared_to(cam.position, [models[ci].transform[3][0], models[ci].transform[3][1], models[ci].transform[3][2]]) for ci in kids}
        for ci in sorted(kids, key=lambda c: -float(dist[c])):                                       # camera.go:2078-2080
            cd = desc.models[ci]
            if not cd.visible:
                continue
            if cd.frustum_culling and not in_frustum(models[ci]):
                continue
            for cp in range(len(cd.mesh.parts)):
                out.append((ci, cp))
    return out


# ---- package sort: sort.SliceStable (the algorithm of Go's sort.Stable: insertion sort on blocks of 20, then SymMerge) -----------
def go_stable_sort(n, less, swap):
    """stable_func of Go's sort package over an abstract (less, swap) pair, as sort.SliceStable drives it."""

    def insertion_sort(a, b):
        for i in range(a + 1, b):
            j = i
            while j > a and less(j, j - 1):
                swap(j, j - 1)
                j -= 1

    def swap_range(a, b, count):
        for i in range(count):
            swap(a + i, b + i)

    def rotate(a, m, b):
        i, j = m - a, b - m
        while i != j:
            if i > j:
                swap_range(m - i, m, j)
                i -= j
            else:
                swap_range(m - i, m + j - i, i)
                j -= i
        swap_range(m - i, m, i)

    def sym_merge(a, m, b):
        if m - a == 1:
            i, j = m, b
            while i < j:
                h = (i + j) >> 1
                if less(h, a):
                    i = h + 1
                else:
                    j = h
            for k in range(a, i - 1):
                swap(k, k + 1)
            return
        if b - m == 1:
            i, j = a, m
            while i < j:
                h = (i + j) >> 1
                if not less(m, h):
                    i = h + 1
                else:
                    j = h
            for k in range(m, i, -1):
                swap(k, k - 1)
            return
        mid = (a + b) >> 1
        nn = mid + m
        if m > mid:
            start, r = nn - b, mid
        else:
            start, r = a, m
        p = nn - 1
        while start < r:
            c = (start + r) >> 1
            if not less(p - c, c):
                start = c + 1
            else:
                r = c
        end = nn - start
        if start < m and m < end:
            rotate(start, m, end)
        if a < start and start < mid:
            sym_merge(a, start, mid)
        if mid < end and end < b:
            sym_merge(mid, end, b)

    block = 20
    a, b = 0, block
    while b <= n:
        insertion_sort(a, b)
        a = b
        b += block
    insertion_sort(a, n)
    while block < n:
        a, b = 0, 2 * block
        while b <= n:
            sym_merge(a, a + block, b)
            a = b
            b += 2 * block
        m = a + block
        if m < n:
            sym_merge(a, m, n)
        block *= 2


def slice_stable(items, less_items):
    """sort.SliceStable(items, less): returns the sorted list."""
    items = list(items)

    def swap(i, j):
        items[i], items[j] = items[j], items[i]

    go_stable_sort(len(items), lambda i, j: less_items(items[i], items[j]), swap)
    return items


# ---- animation.go:63-164, quaternion.go:26-59, 86-104, 132-178, node.go:572-594: the bone poses ------------------------------------
def _key_pair(times, time):
    """The keyframes around `time` as the loops of animation.go:77-86 / 130-139 find them: (first, last) indices."""
    first = last = None
    for k in range(len(times)):
        if times[k] < time:
            first = k
        else:
            last = k
            break
    return first, last


def value_as_vector(track, time):                           # animation.go:63-112: (value, exists)
    time = F(time)
    times, values = track.times, track.values
    if len(times) == 0:
        return None, False
    if time <= times[0]:
        return [values[0][0], values[0][1], values[0][2]], True
    if time >= times[-1]:
        return [values[-1][0], values[-1][1], values[-1][2]], True
    first, last = _key_pair(times, time)
    if time == times[first]:
        return list(values[first][:3]), True
    if time == times[last]:
        return list(values[last][:3]), True
    fd, ld = values[first], values[last]
    if track.interpolation == 1:                            # InterpolationConstant
        return [fd[0], fd[1], fd[2]], True
    t = (time - times[first]) / (times[last] - times[first])
    return [fd[0] + ((ld[0] - fd[0]) * t), fd[1] + ((ld[1] - fd[1]) * t), fd[2] + ((ld[2] - fd[2]) * t)], True      # Vector3.Lerp


def quat_slerp(q, end, percent):                            # quaternion.go:26-59 (x y z w); math32.Sqrt / Atan2 / Sin go through Go's float64 functions
    q, end, percent = [F(c) for c in q], [F(c) for c in end], F(percent)
    cos_omega = q[0] * end[0] + q[1] * end[1] + q[2] * end[2] + q[3] * end[3]
    if cos_omega < 0:
        end = [end[0] * F(-1), end[1] * F(-1), end[2] * F(-1), end[3] * F(-1)]
        cos_omega = cos_omega * F(-1)
    if cos_omega > F(0.9999):
        k0, k1 = F(1) - percent, percent
    else:
        sin_omega = sqrt32(F(1) - (cos_omega * cos_omega))
        omega = F(O.go_atan2(float(sin_omega), float(cos_omega)))
        one_over = F(1) / sin_omega
        k0 = F(O.go_sin(float((F(1) - percent) * omega))) * one_over
        k1 = F(O.go_sin(float(percent * omega))) * one_over
    w = q[3] * k0 + end[3] * k1
    x = q[0] * k0 + end[0] * k1
    y = q[1] * k0 + end[1] * k1
    z = q[2] * k0 + end[2] * k1
    return [x, y, z, w]


def quat_unit(q):                                           # quaternion.go:86-104
    m = sqrt32((q[0] * q[0]) + (q[1] * q[1]) + (q[2] * q[2]) + (q[3] * q[3]))
    return [q[0] / m, q[1] / m, q[2] / m, q[3] / m]


def quat_to_matrix4(q):                                     # quaternion.go:132-178
    x, y, z, w = [F(c) for c in q]
    m1 = np.array([[w, z, -y, x], [-z, w, x, y], [y, -x, w, z], [-x, -y, -z, w]], F)
    m2 = np.array([[w, z, -y, -x], [-z, w, x, -y], [y, -x, w, -z], [x, y, z, w]], F)
    return mat4_mult(m1, m2)


def value_as_quaternion(track, time):                       # animation.go:116-164
    time = F(time)
    times, values = track.times, track.values
    if len(times) == 0:
        return None, False
    if time <= times[0]:
        return list(values[0]), True
    if time >= times[-1]:
        return list(values[-1]), True
    first, last = _key_pair(times, time)
    if time == times[first]:
        return list(values[first]), True
    if time == times[last]:
        return list(values[last]), True
    if track.interpolation == 1:
        return list(values[first]), True
    t = (time - times[first]) / (times[last] - times[first])
    return quat_slerp(values[first], values[last], t), True


def pose_bones(sk, anim, playhead, root16):
    """AnimationPlayer.updateValues + forceUpdate without blending (animation.go:674-722, 939-983) on an armature at rest, then every
    bone's Transform() and boneInfluence (node.go:572-594): (n_bones, 16)."""
    n = len(sk.parent)
    root = mat(root16)
    world = [None] * n
    out = np.zeros((n, 16), F)
    for b in range(n):
        pos = [F(c) for c in sk.rest_position[b]]
        scl = [F(c) for c in sk.rest_scale[b]]
        rot = mat(sk.rest_rotation[b])
        channel = anim.channels.get(sk.names[b]) if anim is not None else None
        if channel:
            if 0 in channel:
                v, exists = value_as_vector(channel[0], playhead)
                if exists:
                    pos = v
            if 1 in channel:
                v, exists = value_as_vector(channel[1], playhead)
                if exists:
                    scl = v
            if 2 in channel:
                q, exists = value_as_quaternion(channel[2], playhead)
                if exists:
                    r = quat_to_matrix4(q)
                    if r.any():                             # SetLocalRotation ignores an all-zero matrix (node.go:906-912)
                        rot = r
        t = mat4_scale(scl[0], scl[1], scl[2])
        t = mat4_mult(t, rot)
        t = mat4_mult(t, mat4_translate(pos[0], pos[1], pos[2]))
        t = mat4_mult(t, world[int(sk.parent[b])] if sk.parent[b] >= 0 else root)
        world[b] = t
        out[b] = mat4_mult(mat(sk.inverse_bind[b]), t).reshape(16)
    return out


# ---- model.go:1203-1268 ------------------------------------------------------------------------------------------------------------
def bake_lighting(desc, model_desc, light_descs, rgb=None):
    """Model.BakeLighting into a colour channel that starts as `rgb` (default: ones, a channel that did not exist): (n_verts, 3).
    Every visible light of the argument list plus the world's ambient light lights the model: ProcessVertices with
    processOnlyVisible = false, then, light by light, light.Light(tri) for every triangle of the part -- so a vertex receives a
    light once per corner that refers to it."""
    cam = Camera(desc.camera)
    model = Model(model_desc)
    mesh = model_desc.mesh
    g = Globals(mesh.n_verts)
    init = np.ones((mesh.n_verts, 3), F) if rgb is None else np.asarray(rgb, F).reshape(-1, 3)
    g.vertex_lights = [[init[i][0], init[i][1], init[i][2], F(1)] for i in range(mesh.n_verts)]      # VertexColors[targetChannel]
    lights = [make_light(ld) for ld in light_descs if ld.visible]
    if desc.world.present and desc.world.ambient_visible:                                            # model.Scene().World.AmbientLight
        lights.append(AmbientLight(desc.world.ambient_color, desc.world.ambient_energy))
    for l in lights:
        l.begin_render()
        l.begin_model(model)
    bucket = SortingTriangleBucket()
    for pi, part in enumerate(mesh.parts):
        process_vertices(model, pi, cam, bucket, g, (), process_only_visible=False)
        if part.material is not None and part.material.shadeless:
            for v in range(part.vertex_start, part.vertex_end):
                g.vertex_lights[v][0] = g.vertex_lights[v][1] = g.vertex_lights[v][2] = F(1)
            continue
        for v in range(part.vertex_start, part.vertex_end):
            g.vertex_lights[v][0] = g.vertex_lights[v][1] = g.vertex_lights[v][2] = F(0)
        for l in lights:
            for tri in model.parts[pi]:                                                              # forEachTri(true, ..)
                l.light(tri, model, g)
    return np.asarray([c[:3] for c in g.vertex_lights], F)


# ---- model.go:1040-1124 (the pass over the model's own triangles) ------------------------------------------------------------------
def vec_angle(v, other):                                    # vector.go:468-478; math32.Acos goes through Go's float64 math.Acos
    eps = F(1e-4)
    if not (abs(other[0]) > eps or abs(other[1]) > eps or abs(other[2]) > eps):
        return F(0)
    d = dot(v, other)
    d = d / ((v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * (other[0] * other[0] + other[1] * other[1] + other[2] * other[2]))
    d = clamp(d, F(-1), F(1))
    return F(O.go_acos(float(d)))


def bake_ao(mesh, tri_normals, tri_max_span, occlusion_angle, occlusion_color, colors4=None, target_parts=None, target_vertices=None):
    """Model.BakeAO without OtherModels: VertexColors[TargetChannel] after the bake, (n_verts, 4)."""
    n = mesh.n_verts
    colors = np.ones((n, 4), F) if colors4 is None else np.asarray(colors4, F).reshape(n, 4).copy()
    centers = O.triangle_centers(mesh.positions, mesh.indices)
    part_of = {}
    for pi, p in enumerate(mesh.parts):
        for t in range(p.tri_start, p.tri_end + 1):
            part_of[t] = pi
    P = mesh.positions
    oc = [F(c) for c in occlusion_color]
    angle_limit = F(occlusion_angle)
    nt = mesh.n_tris
    for ti in range(nt):
        if target_parts is not None and len(target_parts) > 0 and part_of.get(ti) not in target_parts:
            continue
        tri = [int(i) for i in mesh.indices[ti]]
        ao = [F(0), F(0), F(0)]
        for oi in range(nt):
            if oi == ti:
                continue
            other = [int(i) for i in mesh.indices[oi]]
            span = max(F(tri_max_span[ti]), F(tri_max_span[oi])) * F(0.66)
            if distance_squared_to(centers[ti], centers[oi]) > span * span:
                continue
            if vec_angle(tri_normals[ti], tri_normals[oi]) < angle_limit:
                continue
            shared, count = [-1, -1, -1], 0                 # SharesVertexPositions, mesh.go:2050-2081
            for index in range(3):
                for other_index in range(3):
                    if vec_equals(P[tri[index]], P[other[other_index]]):
                        shared[index] = other_index
                        count += 1
                        break
            if count > 0:
                for k in range(3):
                    if shared[k] >= 0:
                        ao[k] = F(1)
            if ao[0] == 1 and ao[1] == 1 and ao[2] == 1:
                break
        for i in range(3):
            p = clamp(ao[i], F(0), F(1))
            vi = tri[i]
            if target_vertices is not None and len(target_vertices) > 0 and vi not in target_vertices:
                continue
            for k in range(4):
                colors[vi][k] = colors[vi][k] + (oc[k] - colors[vi][k]) * p
    return colors

