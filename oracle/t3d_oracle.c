/*
 * t3d_oracle.c -- CPU restatement of the Tetra3D geometry stage.  TEST INFRASTRUCTURE ONLY
 * (see t3d_oracle.h for the rules and the "parity unpinned" note).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * float32 arithmetic, evaluated left to right, multiply and add never fused -- what Go's gc
 * compiler emits on amd64 (GOAMD64=v1) and what matrix_amd64.s does with MULPS/ADDPS.
 *
 * Every function cites the reference lines it follows (paths relative to the reference tree).
 */
#include "t3d_oracle.h"

#include <math.h>
#include <float.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------ */
/* math32 (math32/math32.go)                                                                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct { float X, Y, Z; } vec3;
typedef struct { float X, Y, Z, W; } vec4;
typedef struct { float m[4][4]; } mat4;

static char g_err[256];
const char* orc_last_error(void) { return g_err; }

/* math32.Sqrt: float32(math.Sqrt(float64(x))) (math32.go:125-128) == correctly rounded sqrtf. */
static inline float m32_sqrt(float x) { return (float)sqrt((double)x); }
/* math32.Round: half away from zero (math32.go:114-116). */
static inline float m32_round(float x) { return (float)round((double)x); }
static inline float m32_abs(float x) { return fabsf(x); }
/* math32.Clamp (math32.go:39-46). */
static inline float m32_clamp(float v, float lo, float hi) {
    if (v < lo) return lo;
    else if (v > hi) return hi;
    return v;
}
static inline float m32_tan(float x) { return (float)tan((double)x); }
static inline float m32_cos(float x) { return (float)cos((double)x); }
static inline float m32_sin(float x) { return (float)sin((double)x); }
static inline float m32_atan(float x) { return (float)atan((double)x); }
#define M32_PI ((float)3.14159265358979323846)

/* ------------------------------------------------------------------------------------------ */
/* Vector3 / Vector4 (vector.go)                                                                */
/* ------------------------------------------------------------------------------------------ */

static inline vec3 v3(float x, float y, float z) { vec3 v = {x, y, z}; return v; }
static inline vec3 v3_sub(vec3 a, vec3 b) { return v3(a.X - b.X, a.Y - b.Y, a.Z - b.Z); }
static inline vec3 v3_add(vec3 a, vec3 b) { return v3(a.X + b.X, a.Y + b.Y, a.Z + b.Z); }
static inline vec3 v3_mult(vec3 a, vec3 b) { return v3(a.X * b.X, a.Y * b.Y, a.Z * b.Z); } /* vector.go:248-253 */
static inline vec3 v3_invert(vec3 a) { return v3(-a.X, -a.Y, -a.Z); }
/* vector.go:526-528 */
static inline float v3_dot(vec3 a, vec3 b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
/* vector.go:191-193 */
static inline float v3_magnitude(vec3 v) { return m32_sqrt(v.X * v.X + v.Y * v.Y + v.Z * v.Z); }
/* vector.go:240-245 */
static inline float v3_dist2(vec3 a, vec3 b) {
    a.X -= b.X; a.Y -= b.Y; a.Z -= b.Z;
    return a.X * a.X + a.Y * a.Y + a.Z * a.Z;
}
/* vector.go:258-266 */
static inline vec3 v3_unit(vec3 v) {
    float l = v3_magnitude(v);
    if (l < 1e-8f || l == 1) return v;
    v.X = v.X / l; v.Y = v.Y / l; v.Z = v.Z / l;
    return v;
}
/* vector.go:1869-1877: W untouched */
static inline vec4 v4_unit(vec4 v) {
    float l = m32_sqrt(v.X * v.X + v.Y * v.Y + v.Z * v.Z);
    if (l < 1e-8f || l == 1) return v;
    v.X = v.X / l; v.Y = v.Y / l; v.Z = v.Z / l;
    return v;
}
/* vector.go:164-180 */
static inline vec3 v3_cross(vec3 vec, vec3 other) {
    float ogY = vec.Y, ogZ = vec.Z;
    vec.Z = vec.X * other.Y - other.X * vec.Y;
    vec.Y = ogZ * other.X - other.Z * vec.X;
    vec.X = ogY * other.Z - other.Y * ogZ;
    return vec;
}
/* vector.go:415-425 */
static inline int v3_equals(vec3 a, vec3 b) {
    const float eps = 1e-4f;
    if (m32_abs(a.X - b.X) > eps || m32_abs(a.Y - b.Y) > eps || m32_abs(a.Z - b.Z) > eps) return 0;
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Matrix4 (matrix.go, matrix_amd64.s, matrix_none.go)                                          */
/* ------------------------------------------------------------------------------------------ */

static mat4 m4_identity(void) {
    mat4 r; memset(&r, 0, sizeof r);
    r.m[0][0] = r.m[1][1] = r.m[2][2] = r.m[3][3] = 1;
    return r;
}
/* matrix_amd64.s:4-81 (row i = a[i][0]*B0 + a[i][1]*B1 + a[i][2]*B2 + a[i][3]*B3, mul then add,
 * k ascending) == matrix_none.go:6-32. */
static mat4 m4_mult(mat4 a, mat4 b) {
    mat4 r;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
/* matrix.go:218-230 */
static mat4 m4_transposed(mat4 a) {
    mat4 r;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) r.m[i][j] = a.m[j][i];
    return r;
}
static mat4 m4_translate(float x, float y, float z) { mat4 r = m4_identity(); r.m[3][0] = x; r.m[3][1] = y; r.m[3][2] = z; return r; }
static mat4 m4_scale(float x, float y, float z) { mat4 r = m4_identity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }
/* matrix.go:100-128 */
static mat4 m4_rotate(float x, float y, float z, float angle) {
    if (x == 0 && y == 0 && z == 0) y = 1;
    mat4 mat = m4_identity();
    vec3 v = v3_unit(v3(x, y, z));
    float s = m32_sin(angle), c = m32_cos(angle), m = 1 - c;
    mat.m[0][0] = m * v.X * v.X + c;
    mat.m[0][1] = m * v.X * v.Y + v.Z * s;
    mat.m[0][2] = m * v.Z * v.X - v.Y * s;
    mat.m[1][0] = m * v.X * v.Y - v.Z * s;
    mat.m[1][1] = m * v.Y * v.Y + c;
    mat.m[1][2] = m * v.Y * v.Z + v.X * s;
    mat.m[2][0] = m * v.Z * v.X + v.Y * s;
    mat.m[2][1] = m * v.Y * v.Z - v.X * s;
    mat.m[2][2] = m * v.Z * v.Z + c;
    return mat;
}
static inline vec4 m4_row(const mat4* a, int i) { vec4 v = {a->m[i][0], a->m[i][1], a->m[i][2], a->m[i][3]}; return v; }
static inline void m4_set_row(mat4* a, int i, vec4 v) { a->m[i][0] = v.X; a->m[i][1] = v.Y; a->m[i][2] = v.Z; a->m[i][3] = v.W; }
/* matrix.go:619-627 */
static inline vec3 m4_multvec(const mat4* M, vec3 v) {
    vec3 r;
    r.X = M->m[0][0] * v.X + M->m[1][0] * v.Y + M->m[2][0] * v.Z + M->m[3][0];
    r.Y = M->m[0][1] * v.X + M->m[1][1] * v.Y + M->m[2][1] * v.Z + M->m[3][1];
    r.Z = M->m[0][2] * v.X + M->m[1][2] * v.Y + M->m[2][2] * v.Z + M->m[3][2];
    return r;
}
/* matrix.go:255-307 */
static mat4 m4_inverted(mat4 a) {
#define M(i, j) a.m[i][j]
    float A2323 = M(2,2) * M(3,3) - M(2,3) * M(3,2);
    float A1323 = M(2,1) * M(3,3) - M(2,3) * M(3,1);
    float A1223 = M(2,1) * M(3,2) - M(2,2) * M(3,1);
    float A0323 = M(2,0) * M(3,3) - M(2,3) * M(3,0);
    float A0223 = M(2,0) * M(3,2) - M(2,2) * M(3,0);
    float A0123 = M(2,0) * M(3,1) - M(2,1) * M(3,0);
    float A2313 = M(1,2) * M(3,3) - M(1,3) * M(3,2);
    float A1313 = M(1,1) * M(3,3) - M(1,3) * M(3,1);
    float A1213 = M(1,1) * M(3,2) - M(1,2) * M(3,1);
    float A2312 = M(1,2) * M(2,3) - M(1,3) * M(2,2);
    float A1312 = M(1,1) * M(2,3) - M(1,3) * M(2,1);
    float A1212 = M(1,1) * M(2,2) - M(1,2) * M(2,1);
    float A0313 = M(1,0) * M(3,3) - M(1,3) * M(3,0);
    float A0213 = M(1,0) * M(3,2) - M(1,2) * M(3,0);
    float A0312 = M(1,0) * M(2,3) - M(1,3) * M(2,0);
    float A0212 = M(1,0) * M(2,2) - M(1,2) * M(2,0);
    float A0113 = M(1,0) * M(3,1) - M(1,1) * M(3,0);
    float A0112 = M(1,0) * M(2,1) - M(1,1) * M(2,0);

    float det = M(0,0) * (M(1,1) * A2323 - M(1,2) * A1323 + M(1,3) * A1223) -
                M(0,1) * (M(1,0) * A2323 - M(1,2) * A0323 + M(1,3) * A0223) +
                M(0,2) * (M(1,0) * A1323 - M(1,1) * A0323 + M(1,3) * A0123) -
                M(0,3) * (M(1,0) * A1223 - M(1,1) * A0223 + M(1,2) * A0123);
    det = 1 / det;

    mat4 m = m4_identity();
    m.m[0][0] = det * (M(1,1) * A2323 - M(1,2) * A1323 + M(1,3) * A1223);
    m.m[0][1] = det * -(M(0,1) * A2323 - M(0,2) * A1323 + M(0,3) * A1223);
    m.m[0][2] = det * (M(0,1) * A2313 - M(0,2) * A1313 + M(0,3) * A1213);
    m.m[0][3] = det * -(M(0,1) * A2312 - M(0,2) * A1312 + M(0,3) * A1212);
    m.m[1][0] = det * -(M(1,0) * A2323 - M(1,2) * A0323 + M(1,3) * A0223);
    m.m[1][1] = det * (M(0,0) * A2323 - M(0,2) * A0323 + M(0,3) * A0223);
    m.m[1][2] = det * -(M(0,0) * A2313 - M(0,2) * A0313 + M(0,3) * A0213);
    m.m[1][3] = det * (M(0,0) * A2312 - M(0,2) * A0312 + M(0,3) * A0212);
    m.m[2][0] = det * (M(1,0) * A1323 - M(1,1) * A0323 + M(1,3) * A0123);
    m.m[2][1] = det * -(M(0,0) * A1323 - M(0,1) * A0323 + M(0,3) * A0123);
    m.m[2][2] = det * (M(0,0) * A1313 - M(0,1) * A0313 + M(0,3) * A0113);
    m.m[2][3] = det * -(M(0,0) * A1312 - M(0,1) * A0312 + M(0,3) * A0112);
    m.m[3][0] = det * -(M(1,0) * A1223 - M(1,1) * A0223 + M(1,2) * A0123);
    m.m[3][1] = det * (M(0,0) * A1223 - M(0,1) * A0223 + M(0,2) * A0123);
    m.m[3][2] = det * -(M(0,0) * A1213 - M(0,1) * A0213 + M(0,2) * A0113);
    m.m[3][3] = det * (M(0,0) * A1212 - M(0,1) * A0212 + M(0,2) * A0112);
#undef M
    return m;
}
/* matrix.go:198-205 */
static mat4 m4_decompose_rotation(const mat4* a) {
    mat4 rot = m4_identity();
    m4_set_row(&rot, 0, v4_unit(m4_row(a, 0)));
    m4_set_row(&rot, 1, v4_unit(m4_row(a, 1)));
    m4_set_row(&rot, 2, v4_unit(m4_row(a, 2)));
    return rot;
}
/* matrix.go:181-196 */
static void m4_decompose(const mat4* a, vec3* position, vec3* scale, mat4* rotation) {
    *position = v3(a->m[3][0], a->m[3][1], a->m[3][2]);
    *rotation = m4_decompose_rotation(a);
    mat4 in = m4_mult(*a, m4_transposed(*rotation));
    scale->X = m32_sqrt(in.m[0][0] * in.m[0][0] + in.m[0][1] * in.m[0][1] + in.m[0][2] * in.m[0][2]);
    scale->Y = m32_sqrt(in.m[1][0] * in.m[1][0] + in.m[1][1] * in.m[1][1] + in.m[1][2] * in.m[1][2]);
    scale->Z = m32_sqrt(in.m[2][0] * in.m[2][0] + in.m[2][1] * in.m[2][1] + in.m[2][2] * in.m[2][2]);
}
/* matrix.go:511-522: NaN differences compare false, so a NaN matrix "equals" anything. */
static int m4_equals(const mat4* a, const mat4* b) {
    const float eps = 0.0001f;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
        if (m32_abs(a->m[i][j] - b->m[i][j]) > eps) return 0;
    return 1;
}
static int m4_is_identity(const mat4* a) { mat4 id = m4_identity(); return m4_equals(a, &id); }
/* matrix.go:788-815 */
static mat4 m4_lookat(vec3 from, vec3 to, vec3 up) {
    const vec3 WorldRight = {1, 0, 0}, WorldBackward = {0, 0, 1};
    if (v3_equals(from, to)) return m4_identity();
    vec3 z = v3_unit(v3_sub(to, from));
    up = v3_unit(up);
    if (v3_equals(z, up) || v3_equals(z, v3_invert(up))) {
        if (!v3_equals(up, WorldRight)) up = WorldRight;
        else up = WorldBackward;
    }
    vec3 x = v3_unit(v3_cross(up, z));
    vec3 y = v3_cross(z, x);
    mat4 r = m4_identity();
    r.m[0][0] = x.X; r.m[0][1] = x.Y; r.m[0][2] = x.Z; r.m[0][3] = 0;
    r.m[1][0] = y.X; r.m[1][1] = y.Y; r.m[1][2] = y.Z; r.m[1][3] = 0;
    r.m[2][0] = z.X; r.m[2][1] = z.Y; r.m[2][2] = z.Z; r.m[2][3] = 0;
    return r;
}
/* matrix.go:588-604 */
static mat4 m4_perspective(float fovy, float near_, float far_, float viewWidth, float viewHeight) {
    float aspect = viewWidth / viewHeight;
    float t = m32_tan(fovy * M32_PI / 360);
    float b = -t;
    float r = t * aspect;
    float l = -r;
    mat4 m; memset(&m, 0, sizeof m);
    m.m[0][0] = (2 * near_) / (r - l); m.m[0][2] = (r + l) / (r - l);
    m.m[1][1] = (2 * near_) / (t - b); m.m[1][2] = (t + b) / (t - b);
    m.m[2][2] = -((far_ + near_) / (far_ - near_)); m.m[2][3] = -((2 * far_ * near_) / (far_ - near_));
    m.m[3][2] = -1;
    return m;
}
/* matrix.go:609-616 */
static mat4 m4_orthographic(float near_, float far_, float right, float left, float top, float bottom) {
    mat4 m; memset(&m, 0, sizeof m);
    m.m[0][0] = 2 / (right - left);
    m.m[1][1] = 2 / (top - bottom);
    m.m[2][2] = -2 / (far_ - near_);
    m.m[3][3] = 1;
    return m;
}

static mat4 m4_load(const float* p) { mat4 r; memcpy(&r, p, 64); return r; }
static void m4_store(const mat4* a, float* p) { memcpy(p, a, 64); }

/* exported math for the known-answer tests */
void orc_mat4_identity(float* m) { mat4 r = m4_identity(); m4_store(&r, m); }
void orc_mat4_mult(const float* a, const float* b, float* out) { mat4 r = m4_mult(m4_load(a), m4_load(b)); m4_store(&r, out); }
void orc_mat4_inverted(const float* m, float* out) { mat4 r = m4_inverted(m4_load(m)); m4_store(&r, out); }
void orc_mat4_transposed(const float* m, float* out) { mat4 r = m4_transposed(m4_load(m)); m4_store(&r, out); }
void orc_mat4_rotate(float x, float y, float z, float angle, float* out) { mat4 r = m4_rotate(x, y, z, angle); m4_store(&r, out); }
void orc_mat4_translate(float x, float y, float z, float* out) { mat4 r = m4_translate(x, y, z); m4_store(&r, out); }
void orc_mat4_scale(float x, float y, float z, float* out) { mat4 r = m4_scale(x, y, z); m4_store(&r, out); }
int  orc_mat4_equals(const float* a, const float* b) { mat4 x = m4_load(a), y = m4_load(b); return m4_equals(&x, &y); }
void orc_mat4_decompose(const float* m, float* pos3, float* scale3, float* rot16) {
    mat4 a = m4_load(m), r; vec3 p, s;
    m4_decompose(&a, &p, &s, &r);
    pos3[0] = p.X; pos3[1] = p.Y; pos3[2] = p.Z;
    scale3[0] = s.X; scale3[1] = s.Y; scale3[2] = s.Z;
    m4_store(&r, rot16);
}
void orc_mat4_lookat(const float* f, const float* t, const float* u, float* out) {
    mat4 r = m4_lookat(v3(f[0], f[1], f[2]), v3(t[0], t[1], t[2]), v3(u[0], u[1], u[2])); m4_store(&r, out);
}
void orc_mat4_perspective(float fovy, float n, float f, float w, float h, float* out) { mat4 r = m4_perspective(fovy, n, f, w, h); m4_store(&r, out); }
void orc_mat4_orthographic(float n, float f, float r_, float l, float t, float b, float* out) { mat4 r = m4_orthographic(n, f, r_, l, t, b); m4_store(&r, out); }

/* ------------------------------------------------------------------------------------------ */
/* Node (node.go)                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* node.go:572-608 (without the cache): S * R * T [* parent.Transform()] */
static mat4 node_transform(const orc_node* nodes, int idx) {
    const orc_node* n = &nodes[idx];
    mat4 t = m4_scale(n->scale[0], n->scale[1], n->scale[2]);
    t = m4_mult(t, m4_load(n->rotation));
    t = m4_mult(t, m4_translate(n->position[0], n->position[1], n->position[2]));
    if (n->parent >= 0) t = m4_mult(t, node_transform(nodes, n->parent));
    return t;
}
void orc_node_transform(const orc_node* nodes, int32_t idx, float* out16) { mat4 t = node_transform(nodes, idx); m4_store(&t, out16); }

/* matrix.go:170-176 */
static vec3 m4_forward(const mat4* a) { return v3_unit(v3(a->m[2][0], a->m[2][1], a->m[2][2])); }
static vec3 m4_right(const mat4* a) { return v3_unit(v3(a->m[0][0], a->m[0][1], a->m[0][2])); }
static vec3 m4_up(const mat4* a) { return v3_unit(v3(a->m[1][0], a->m[1][1], a->m[1][2])); }

/* ------------------------------------------------------------------------------------------ */
/* Mesh-derived data                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* mesh.go:1980-1982 */
void orc_triangle_centers(const float* pos, const uint32_t* idx, int32_t n_tris, float* c) {
    for (int t = 0; t < n_tris; t++) {
        const float* a = pos + 3 * (size_t)idx[3 * t], *b = pos + 3 * (size_t)idx[3 * t + 1], *d = pos + 3 * (size_t)idx[3 * t + 2];
        c[3 * t + 0] = (a[0] + b[0] + d[0]) / 3;
        c[3 * t + 1] = (a[1] + b[1] + d[1]) / 3;
        c[3 * t + 2] = (a[2] + b[2] + d[2]) / 3;
    }
}

typedef struct { vec3 Min, Max; } dims_t;
/* mesh.go:719-751 (newEmptyDimensions mesh.go:15-22) */
static dims_t mesh_dimensions(const orc_mesh* m) {
    dims_t d = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
    for (int i = 0; i < m->n_verts; i++) {
        float x = m->positions[3 * i], y = m->positions[3 * i + 1], z = m->positions[3 * i + 2];
        if (d.Min.X > x) d.Min.X = x;
        if (d.Min.Y > y) d.Min.Y = y;
        if (d.Min.Z > z) d.Min.Z = z;
        if (d.Max.X < x) d.Max.X = x;
        if (d.Max.Y < y) d.Max.Y = y;
        if (d.Max.Z < z) d.Max.Z = z;
    }
    return d;
}
/* mesh.go:41-43 */
static float dims_max_span(dims_t d) { return v3_magnitude(v3_sub(d.Max, d.Min)); }
/* mesh.go:46-52 */
static vec3 dims_center(dims_t d) { return v3((d.Max.X + d.Min.X) / 2, (d.Max.Y + d.Min.Y) / 2, (d.Max.Z + d.Min.Z) / 2); }

/* ------------------------------------------------------------------------------------------ */
/* Render state: the reference's package-level globals (tetra3d.go:21-35, sortingTriangles.go)   */
/* ------------------------------------------------------------------------------------------ */

typedef struct { int tri; float depth; } sorting_tri;   /* sortingTriangles.go:9-12 */

typedef struct {
    const orc_scene* sc;
    /* Node.cachedTransform (node.go:576-589): every node's world transform is computed once per frame */
    mat4* nodeT; uint8_t* nodeTValid;
    /* per-mesh derived */
    float**  centers;     /* Triangle.Center */
    dims_t*  dims;
    /* camera-derived (camera.go:1085-1088, 638-679) */
    mat4 camRot, view, proj, vp;
    vec3 camPos, camForward, camRight, camUp;
    float sphereFactorTang, sphereFactorY, sphereFactorX;
    /* scratch sized to the largest mesh (tetra3d.go:30-35) */
    vec4* vertexTransforms;          /* globalVertexTransforms */
    vec3* transformedNormals;        /* globalVertexTransformedNormals */
    float* depthUnbillboarded;       /* globalVertexDepthUnbillboarded */
    vec3* alteredPositions;          /* globalMeshAlteredVertexPositions */
    vec3* alteredNormals;            /* globalMeshAlteredVertexNormals */
    vec4* vertexLights;              /* mesh.vertexLights.colors */
    /* globalSortingTriangleBucket */
    int sortMode;
    sorting_tri* unsetTris; int unsetTriIndex;
    int* sortedTris; int sortedCount;       /* the ForEach order */
    int* binCount;                           /* 512 */
    uint16_t* binOf;
    /* per-light working state (light.go:83, 143, 292-293) */
    float (*ambResult)[3];
    vec3* workingPosition;
    vec3* workingForward;
    mat4* workingModelRotation;
    /* active light list */
    int* sceneLights; int nSceneLights;
    /* output */
    orc_result* res;
    int64_t cap_vertices;    /* capacity of the lists in vertices */
    int64_t cap_captured;
    int capture;
    int vertexListIndex;     /* within the current flush */
    int64_t flushBase;       /* offset of the current flush in the concatenated lists */
    int curFlush;
} rstate;

/* cached Node.Transform() / WorldPosition() / WorldRotation() (node.go:572-608, 737-739, 916-938) */
static mat4 S_transform(rstate* S, int idx) {
    if (!S->nodeTValid[idx]) {
        const orc_node* n = &S->sc->nodes[idx];
        mat4 t = m4_scale(n->scale[0], n->scale[1], n->scale[2]);
        t = m4_mult(t, m4_load(n->rotation));
        t = m4_mult(t, m4_translate(n->position[0], n->position[1], n->position[2]));
        if (n->parent >= 0) t = m4_mult(t, S_transform(S, n->parent));
        S->nodeT[idx] = t; S->nodeTValid[idx] = 1;
    }
    return S->nodeT[idx];
}
static vec3 S_world_position(rstate* S, int idx) {
    mat4 t = S_transform(S, idx);
    return v3(t.m[3][0], t.m[3][1], t.m[3][2]);
}
static mat4 S_world_rotation(rstate* S, int idx) {
    const orc_node* nodes = S->sc->nodes;
    int need = 0;
    int p = nodes[idx].parent;
    while (p >= 0) {
        mat4 rot = m4_load(nodes[p].rotation);
        if (!m4_is_identity(&rot)) { need = 1; break; }
        p = nodes[p].parent;
    }
    if (need) { mat4 t = S_transform(S, idx); return m4_decompose_rotation(&t); }
    return m4_load(nodes[idx].rotation);
}

static double now_s(void) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void ensure_lists(rstate* S, int64_t need) {
    if (need <= S->cap_vertices) return;
    int64_t cap = S->cap_vertices ? S->cap_vertices : 30000;   /* tetra3d.go:19 */
    while (cap < need) cap *= 2;                               /* tetra3d.go:41-90 doubles */
    orc_result* r = S->res;
    r->color_list = realloc(r->color_list, (size_t)cap * sizeof(orc_vertex));
    r->depth_list = realloc(r->depth_list, (size_t)cap * sizeof(orc_vertex));
    r->normal_list = realloc(r->normal_list, (size_t)cap * sizeof(orc_vertex));
    r->index_list = realloc(r->index_list, (size_t)cap * sizeof(uint16_t));
    r->tri_ids = realloc(r->tri_ids, (size_t)(cap / 3 + 1) * sizeof(uint32_t));
    /* fresh ebiten.Vertex{} entries are zero (tetra3d.go:52-62); fields the emit loop does not
     * write in the current camera configuration therefore read as zero. */
    memset(r->color_list + S->cap_vertices, 0, (size_t)(cap - S->cap_vertices) * sizeof(orc_vertex));
    memset(r->depth_list + S->cap_vertices, 0, (size_t)(cap - S->cap_vertices) * sizeof(orc_vertex));
    memset(r->normal_list + S->cap_vertices, 0, (size_t)(cap - S->cap_vertices) * sizeof(orc_vertex));
    memset(r->index_list + S->cap_vertices, 0, (size_t)(cap - S->cap_vertices) * sizeof(uint16_t));
    S->cap_vertices = cap;
}

/* ------------------------------------------------------------------------------------------ */
/* Camera (camera.go)                                                                           */
/* ------------------------------------------------------------------------------------------ */

static float cam_aspect(const orc_camera* c) { return (float)c->width / (float)c->height; }  /* camera.go:992-995 */

static void camera_setup(rstate* S) {
    const orc_scene* sc = S->sc; const orc_camera* c = &sc->camera;
    S->camPos = S_world_position(S, c->node);
    S->camRot = S_world_rotation(S, c->node);
    /* camera.go:1085-1088 (ClearWithColor caches these) */
    S->camForward = m4_forward(&S->camRot);
    S->camRight = m4_right(&S->camRot);
    S->camUp = m4_up(&S->camRot);
    /* camera.go:638-648 */
    vec3 ip = v3_invert(S->camPos);
    S->view = m4_mult(m4_translate(ip.X, ip.Y, ip.Z), m4_transposed(S->camRot));
    /* camera.go:659-665; the 3.1415 literal is the reference's */
    float angle = c->field_of_view * 3.1415f / 360;
    S->sphereFactorTang = m32_tan(angle);
    S->sphereFactorY = 1.0f / m32_cos(angle);
    S->sphereFactorX = 1.0f / m32_cos(m32_atan(S->sphereFactorTang * cam_aspect(c)));
    /* camera.go:667-676 */
    if (c->perspective) {
        S->proj = m4_perspective(c->field_of_view, c->near_plane, c->far_plane, (float)c->width, (float)c->height);
    } else {
        float asr = (float)c->height / (float)c->width;
        S->proj = m4_orthographic(c->near_plane, c->far_plane, 1 * c->ortho_scale, -1 * c->ortho_scale,
                                  asr * c->ortho_scale, -asr * c->ortho_scale);
    }
    S->vp = m4_mult(S->view, S->proj);   /* camera.go:1329 */
}

/* camera.go:933-983 */
static int sphere_in_frustum(const rstate* S, vec3 position, float radius) {
    const orc_camera* c = &S->sc->camera;
    vec3 diff = v3_sub(position, S->camPos);
    float pcZ = v3_dot(diff, v3_invert(S->camForward));
    if (pcZ > c->far_plane + radius || pcZ < c->near_plane - radius) return 0;
    if (c->perspective) {
        float d = S->sphereFactorY * radius;
        pcZ *= S->sphereFactorTang;
        float pcY = v3_dot(diff, S->camUp);
        if (pcY > pcZ + d || pcY < -pcZ - d) return 0;
        pcZ *= cam_aspect(c);
        d = S->sphereFactorX * radius;
        float pcX = v3_dot(diff, S->camRight);
        if (pcX > pcZ + d || pcX < -pcZ - d) return 0;
    } else {
        float width = c->ortho_scale;
        float height = width / cam_aspect(c);
        float pcY = v3_dot(diff, S->camUp);
        if (-height / 2 - radius > pcY || pcY > height / 2 + radius) return 0;
        float pcX = v3_dot(diff, S->camRight);
        if (-width / 2 - radius > pcX || pcX > width / 2 + radius) return 0;
    }
    return 1;
}

/* model.go:142-182 -> frustumCullingSphere position/radius; camera.go:986-989 */
static int model_in_frustum(rstate* S, int mi) {
    const orc_model* model = &S->sc->models[mi];
    mat4 transform = S_transform(S, model->node);
    vec3 position, scale; mat4 rotation;
    m4_decompose(&transform, &position, &scale, &rotation);
    vec3 center = {0, 0, 0};
    dims_t dim = S->dims[model->mesh];
    if (!model->skinned) center = dims_center(dim);
    center = m4_multvec(&rotation, center);
    position.X += center.X * scale.X;
    position.Y += center.Y * scale.Y;
    position.Z += center.Z * scale.Z;
    dim.Min.X *= scale.X; dim.Min.Y *= scale.Y; dim.Min.Z *= scale.Z;
    dim.Max.X *= scale.X; dim.Max.Y *= scale.Y; dim.Max.Z *= scale.Z;
    float radius = dims_max_span(dim) / 2;
    return sphere_in_frustum(S, position, radius);
}

/* model.go:1274-1282 */
static int is_transparent(const orc_model* model, const orc_part* part) {
    const orc_material* mat = &part->material;
    if (mat->present) {
        int matT = mat->transparency_mode == ORC_TRANSPARENCY_TRANSPARENT || !mat->blend_source_over ||
                   (mat->transparency_mode == ORC_TRANSPARENCY_AUTO && mat->color[3] < 0.999f);
        int modelT = mat->transparency_mode != ORC_TRANSPARENCY_OPAQUE && model->color[3] < 0.999f;
        return matT || modelT;
    }
    return model->color[3] < 0.999f;
}
static int part_is_visible(const orc_part* p) { return p->material.present ? p->material.visible : 1; }  /* mesh.go:2402-2407 */

/* ------------------------------------------------------------------------------------------ */
/* sortingTriangleBucket (sortingTriangles.go)                                                  */
/* ------------------------------------------------------------------------------------------ */

#define ORC_BINS 512     /* sortingTriangles.go:166-168 */

/* sortingTriangles.go:41-56 for one triangle */
int orc_bin_index(float depth, float minRange, float maxRange, int sort_mode) {
    float rangeDiff = maxRange - minRange;
    if (rangeDiff == 0) { maxRange += 0.001f; rangeDiff += 0.001f; }
    (void)maxRange;
    int targetBin = 0;
    if (sort_mode != ORC_SORT_NONE && ORC_BINS > 1) {
        float d = (depth - minRange) / rangeDiff * (float)ORC_BINS;
        float t = m32_clamp(d, 0, (float)(ORC_BINS - 1));
        targetBin = (int)t;
    }
    return targetBin;
}

/* sortingTriangles.go:38-64 (Sort) + 93-125 (ForEach order), producing S->sortedTris. */
static void bucket_sort(rstate* S, float minRange, float maxRange) {
    int n = S->unsetTriIndex;
    memset(S->binCount, 0, ORC_BINS * sizeof(int));
    for (int i = 0; i < n; i++) {
        int b = orc_bin_index(S->unsetTris[i].depth, minRange, maxRange, S->sortMode);
        S->binOf[i] = (uint16_t)b;
        S->binCount[b]++;
    }
    /* ForEach: BackToFront walks bins 511..0, otherwise 0..511; inside a bin, arrival order. */
    int start[ORC_BINS];
    int acc = 0;
    if (S->sortMode == ORC_SORT_BACK_TO_FRONT) {
        for (int b = ORC_BINS - 1; b >= 0; b--) { start[b] = acc; acc += S->binCount[b]; }
    } else {
        for (int b = 0; b < ORC_BINS; b++) { start[b] = acc; acc += S->binCount[b]; }
    }
    for (int i = 0; i < n; i++) S->sortedTris[start[S->binOf[i]]++] = S->unsetTris[i].tri;
    S->sortedCount = n;
}

/* ------------------------------------------------------------------------------------------ */
/* Model.skinVertex / Model.ProcessVertices (model.go:559-995)                                  */
/* ------------------------------------------------------------------------------------------ */

/* mesh.go:796-822 */
static vec3 vertex_with_shape_keys(const orc_mesh* mesh, const float* base, const float* keys, int v) {
    vec3 out = v3(base[3 * v], base[3 * v + 1], base[3 * v + 2]);
    if (mesh->n_shape_keys > 0) {
        for (int k = 0; k < mesh->n_shape_keys; k++) {
            const float* kp = keys + ((size_t)k * mesh->n_verts + v) * 3;
            float w = mesh->key_weights[k];
            out.X += kp[0] * w;
            out.Y += kp[1] * w;
            out.Z += kp[2] * w;
        }
    }
    return out;
}

/* model.go:559-596 */
static void skin_vertex(const orc_mesh* mesh, const orc_model* model, int vertID, vec3* vertOut, vec3* normalOut) {
    mat4 skin; memset(&skin, 0, sizeof skin);                      /* skinMatrix.Clear() */
    int n = mesh->n_influences ? mesh->n_influences[vertID] : 4;
    for (int boneIndex = 0; boneIndex < n; boneIndex++) {
        float weightPerc = mesh->weights[4 * vertID + boneIndex];
        if (weightPerc == 0) continue;
        mat4 infl = m4_load(model->bone_influences + 16 * (size_t)mesh->bones[4 * vertID + boneIndex]);
        if (weightPerc == 1) { skin = infl; break; }
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
            float scaled = infl.m[i][j] * weightPerc;               /* Matrix4.Scale, matrix.go:651-661 */
            skin.m[i][j] += scaled;                                 /* Matrix4.Add,   matrix.go:639-649 */
        }
    }
    *vertOut = m4_multvec(&skin, vertex_with_shape_keys(mesh, mesh->positions, mesh->key_positions, vertID));
    skin.m[3][0] = 0; skin.m[3][1] = 0; skin.m[3][2] = 0; skin.m[3][3] = 1;
    *normalOut = m4_multvec(&skin, vertex_with_shape_keys(mesh, mesh->normals, mesh->key_normals, vertID));
}

/* model.go:600-995; returns min/max depth through pointers. */
/* colliderTriangles.go:177-182, 206-281: closestPointOnTri and its helpers */
static vec3 plane_closest_point(vec3 normal, float distance, vec3 point) {
    float dist = v3_dot(normal, point) - distance;
    return v3_sub(point, v3(normal.X * dist, normal.Y * dist, normal.Z * dist));
}
static vec3 closest_point_on_line(vec3 point, vec3 start, vec3 end) {
    vec3 diff = v3_sub(end, start);
    float dotA = v3_dot(v3_sub(point, start), diff);
    float dotB = v3_dot(diff, diff);
    float d = dotA / dotB;
    if (d > 1) d = 1; else if (d < 0) d = 0;
    return v3_add(start, v3(diff.X * d, diff.Y * d, diff.Z * d));
}
static int point_inside_triangle(vec3 point, vec3 v0, vec3 v1, vec3 v2) {
    vec3 ca = v3_sub(v2, v0), ba = v3_sub(v1, v0), pa = v3_sub(point, v0);
    float dot00 = v3_dot(ca, ca), dot01 = v3_dot(ca, ba), dot02 = v3_dot(ca, pa), dot11 = v3_dot(ba, ba), dot12 = v3_dot(ba, pa);
    float invDenom = 1.0f / ((dot00 * dot11) - (dot01 * dot01));
    float u = ((dot11 * dot02) - (dot01 * dot12)) * invDenom;
    float v = ((dot00 * dot12) - (dot01 * dot02)) * invDenom;
    return (u >= 0) && (v >= 0) && (u + v < 1);
}
static vec3 closest_point_on_tri(vec3 point, vec3 v0, vec3 v1, vec3 v2, vec3 normal) {
    float distance = v3_dot(normal, v0);
    vec3 planePoint = plane_closest_point(normal, distance, point);
    if (point_inside_triangle(planePoint, v0, v1, v2)) return planePoint;
    vec3 ab = closest_point_on_line(point, v0, v1), bc = closest_point_on_line(point, v1, v2), ca = closest_point_on_line(point, v2, v0);
    vec3 closest = ab;
    float closestDist = v3_dist2(point, ab);
    float bcDist = v3_dist2(point, bc), caDist = v3_dist2(point, ca);
    if (bcDist < closestDist) { closest = bc; closestDist = bcDist; }
    if (caDist < closestDist) closest = ca;
    return closest;
}

/* Triangle.handleSubdivision (mesh.go:1877-1959) for base triangle `ti`: which of {the base, one of its levels} is visible.
 * Returns -1 = the base triangle itself, else the index of the visible level. */
static int subdivision_choice(const orc_scene* sc, const orc_mesh* mesh, int ti, vec3 cameraPos) {
    const int nLevels = mesh->tri_n_levels[ti];
    if (nLevels == 0) return -1;
    const uint32_t* tri = mesh->indices + 3 * (size_t)ti;
    const float* P = mesh->positions;
    vec3 v0 = v3(P[3 * tri[0]], P[3 * tri[0] + 1], P[3 * tri[0] + 2]);
    vec3 v1 = v3(P[3 * tri[1]], P[3 * tri[1] + 1], P[3 * tri[1] + 2]);
    vec3 v2 = v3(P[3 * tri[2]], P[3 * tri[2] + 1], P[3 * tri[2] + 2]);
    vec3 nrm = v3(mesh->tri_normal[3 * ti], mesh->tri_normal[3 * ti + 1], mesh->tri_normal[3 * ti + 2]);
    float dist = v3_dist2(cameraPos, closest_point_on_tri(cameraPos, v0, v1, v2, nrm));
    int active = -1;
    float closestDistance = FLT_MAX;
    for (int i = 0; i < sc->n_subdivision_levels; i++) {
        float d2 = sc->subdivision_levels[2 * i];
        if (dist < d2 && d2 < closestDistance) { closestDistance = d2; active = i; }
    }
    if (active == -1) return -1;
    float ms = mesh->tri_max_span[ti];
    int subIndex = -1;
    while (ms > sc->subdivision_levels[2 * active + 1] && subIndex < 200) { ms /= 2; subIndex++; }   /* bounded: an infinite MaxSpan hangs the reference */
    if (subIndex >= nLevels) subIndex = nLevels - 1;
    return subIndex;
}

static void process_vertices(rstate* S, int mi, int pi, float* outMin, float* outMax) {
    const orc_scene* sc = S->sc;
    const orc_model* model = &sc->models[mi];
    const orc_mesh* mesh = &sc->meshes[model->mesh];
    const orc_part* part = &mesh->parts[pi];
    const orc_material* mat = &part->material;
    const orc_camera* camera = &sc->camera;
    const mat4 vpMatrix = S->vp;

    S->unsetTriIndex = 0;                                   /* bucket.Clear(), model.go:602 */
    S->sortedCount = 0;
    *outMin = FLT_MAX; *outMax = -FLT_MAX;

    if ((model->color[3] == 0 && is_transparent(model, part)) || !model->visible) return;   /* model.go:604 */

    mat4 modelTransform = S_transform(S, model->node);
    mat4 base = modelTransform;
    vec3 camPos = S->camPos;

    /* model.go:629-630 */
    vec3 p, s; mat4 r;
    { mat4 inv = m4_inverted(modelTransform); m4_decompose(&inv, &p, &s, &r); }
    vec3 invertedCamPos = v3_add(m4_multvec(&r, camPos), v3_mult(p, v3(1 / s.X, 1 / s.Y, 1 / s.Z)));

    /* model.go:634-693 */
    if (mat->present && mat->billboard_enabled) {
        vec3 up = {0, 1, 0};
        if (mat->billboard_up_direction == ORC_BILLBOARD_Y_UP_CAMERA_Y) up = S->camUp;
        mat4 lookat = m4_lookat(v3(modelTransform.m[3][0], modelTransform.m[3][1], modelTransform.m[3][2]), camPos, up);
        if (mat->billboard_lock_x) {
            vec4 e = {1, 0, 0, 0}; m4_set_row(&lookat, 0, e);
            vec4 row = m4_row(&lookat, 1); row.X = 0; m4_set_row(&lookat, 1, v4_unit(row));
            row = m4_row(&lookat, 2); row.X = 0; m4_set_row(&lookat, 2, v4_unit(row));
        }
        if (mat->billboard_lock_y) {
            vec4 e = {0, 1, 0, 0}; m4_set_row(&lookat, 1, e);
            vec4 row = m4_row(&lookat, 0); row.Y = 0; m4_set_row(&lookat, 0, v4_unit(row));
            row = m4_row(&lookat, 2); row.Y = 0; m4_set_row(&lookat, 2, v4_unit(row));
        }
        if (mat->billboard_lock_z) {
            vec4 e = {0, 0, 1, 0}; m4_set_row(&lookat, 2, e);
            vec4 row = m4_row(&lookat, 0); row.Z = 0; m4_set_row(&lookat, 0, v4_unit(row));
            row = m4_row(&lookat, 1); row.Z = 0; m4_set_row(&lookat, 1, v4_unit(row));
        }
        vec3 bp, bs; mat4 br;
        m4_decompose(&base, &bp, &bs, &br);
        base = m4_mult(m4_mult(br, m4_scale(bs.X, bs.Y, bs.Z)), lookat);
        vec4 row3 = {bp.X, bp.Y, bp.Z, 1}; m4_set_row(&base, 3, row3);
    }

    mat4 mvp = m4_mult(base, vpMatrix);                     /* model.go:695 */

    mat4 unalteredMVP = m4_identity();
    int unbillboarded = 0;
    if (mat->present && mat->billboarded_depth_mode == ORC_DEPTH_MODE_UNBILLBOARDED) {   /* model.go:700-703 */
        unbillboarded = 1;
        unalteredMVP = m4_mult(modelTransform, vpMatrix);
    }

    mat4 mvJustRForNormals = m4_identity();
    if (camera->render_normals) {                            /* model.go:705-708 */
        mat4 mv = m4_mult(modelTransform, S->view);
        mvJustRForNormals = m4_decompose_rotation(&mv);
    }

    float minDepth = FLT_MAX, maxDepth = -FLT_MAX;           /* model.go:710-711 */
    const float camNear = camera->near_plane, camFar = camera->far_plane;
    const int modelSkinned = model->skinned;
    const int vertexSnappingOn = camera->vertex_snapping > 0;
    const int renderNormals = camera->render_normals;
    const float* centers = S->centers[model->mesh];

    int subChoice = -1;                                      /* of the base triangle the loop is in (model.go:726-739) */
    for (int ti = part->triangle_start; ti <= part->triangle_end; ti++) {   /* forEachTri, mesh.go:2238-2256 */
        const uint32_t* tri = mesh->indices + 3 * (size_t)ti;
        vec3 vert, normal;

        if (mesh->auto_subdivide) {
            /* Triangle.visible as handleSubdivision leaves it (mesh.go:1877-1959).  Skinned models skip that block
             * (model.go:726) and see whatever the last static draw of the mesh left: the contract is base triangles only,
             * as with an empty level list (disableSubdivision). */
            const int level = mesh->tri_sub_level[ti];
            if (level < 0) subChoice = (!modelSkinned && sc->n_subdivision_levels > 0) ? subdivision_choice(sc, mesh, ti, invertedCamPos) : -1;
            if (level != subChoice) continue;                /* model.go:753-755 */
        }

        for (int vi = 0; vi < 3; vi++) {                     /* model.go:757-822 */
            int vertexIndex = (int)tri[vi];
            vec4* T = &S->vertexTransforms[vertexIndex];
            if (modelSkinned) {
                skin_vertex(mesh, model, vertexIndex, &vert, &normal);
                S->alteredPositions[vertexIndex] = vert;
                S->alteredNormals[vertexIndex] = normal;
                T->X = vpMatrix.m[0][0] * vert.X + vpMatrix.m[1][0] * vert.Y + vpMatrix.m[2][0] * vert.Z + vpMatrix.m[3][0];
                T->Y = vpMatrix.m[0][1] * vert.X + vpMatrix.m[1][1] * vert.Y + vpMatrix.m[2][1] * vert.Z + vpMatrix.m[3][1];
                T->Z = vpMatrix.m[0][2] * vert.X + vpMatrix.m[1][2] * vert.Y + vpMatrix.m[2][2] * vert.Z + vpMatrix.m[3][2];
                T->W = vpMatrix.m[0][3] * vert.X + vpMatrix.m[1][3] * vert.Y + vpMatrix.m[2][3] * vert.Z + vpMatrix.m[3][3];
            } else {
                vert = vertex_with_shape_keys(mesh, mesh->positions, mesh->key_positions, vertexIndex);
                normal = vertex_with_shape_keys(mesh, mesh->normals, mesh->key_normals, vertexIndex);
                S->alteredPositions[vertexIndex] = vert;
                S->alteredNormals[vertexIndex] = normal;
                T->X = mvp.m[0][0] * vert.X + mvp.m[1][0] * vert.Y + mvp.m[2][0] * vert.Z + mvp.m[3][0];
                T->Y = mvp.m[0][1] * vert.X + mvp.m[1][1] * vert.Y + mvp.m[2][1] * vert.Z + mvp.m[3][1];
                T->Z = mvp.m[0][2] * vert.X + mvp.m[1][2] * vert.Y + mvp.m[2][2] * vert.Z + mvp.m[3][2];
                T->W = mvp.m[0][3] * vert.X + mvp.m[1][3] * vert.Y + mvp.m[2][3] * vert.Z + mvp.m[3][3];
                if (unbillboarded)
                    S->depthUnbillboarded[vertexIndex] = unalteredMVP.m[0][2] * vert.X + unalteredMVP.m[1][2] * vert.Y +
                                                         unalteredMVP.m[2][2] * vert.Z + unalteredMVP.m[3][2];
            }
            if (vertexSnappingOn) {                          /* model.go:807-812 */
                float vs = camera->vertex_snapping;
                T->X = m32_round(T->X * vs) / vs;
                T->Y = m32_round(T->Y * vs) / vs;
                T->Z = m32_round(T->Z * vs) / vs;
            }
            if (renderNormals) {                             /* model.go:814-818 (uses vert, as the reference does) */
                vec3* N = &S->transformedNormals[vertexIndex];
                N->X = mvJustRForNormals.m[0][0] * vert.X + mvJustRForNormals.m[1][0] * vert.Y + mvJustRForNormals.m[2][0] * vert.Z + mvJustRForNormals.m[3][0];
                N->Y = mvJustRForNormals.m[0][1] * vert.X + mvJustRForNormals.m[1][1] * vert.Y + mvJustRForNormals.m[2][1] * vert.Z + mvJustRForNormals.m[3][1];
                N->Z = mvJustRForNormals.m[0][2] * vert.X + mvJustRForNormals.m[1][2] * vert.Y + mvJustRForNormals.m[2][2] * vert.Z + mvJustRForNormals.m[3][2];
            }
        }

        vec3 skinnedTriCenter = {0, 0, 0};
        vec3 tv[3];
        int outOfBounds = 1;

        for (int i = 0; i < 3; i++) {                        /* model.go:857-881 */
            int vidx = (int)tri[i];
            if (modelSkinned) {
                skinnedTriCenter.X += S->alteredPositions[vidx].X;
                skinnedTriCenter.Y += S->alteredPositions[vidx].Y;
                skinnedTriCenter.Z += S->alteredPositions[vidx].Z;
            }
            float w = S->vertexTransforms[vidx].W;
            if (w < 0) w = 0.000001f;
            tv[i].X = S->vertexTransforms[vidx].X / w;
            tv[i].Y = S->vertexTransforms[vidx].Y / w;
            if (S->vertexTransforms[vidx].Z + 1 >= camNear && S->vertexTransforms[vidx].Z < camFar) outOfBounds = 0;
        }

        if (outOfBounds) continue;                           /* model.go:883-886 */

        if ((tv[0].X < -0.5f && tv[1].X < -0.5f && tv[2].X < -0.5f) ||      /* model.go:890-898 */
            (tv[0].X > 0.5f && tv[1].X > 0.5f && tv[2].X > 0.5f) ||
            (tv[0].Y < -0.5f && tv[1].Y < -0.5f && tv[2].Y < -0.5f) ||
            (tv[0].Y > 0.5f && tv[1].Y > 0.5f && tv[2].Y > 0.5f))
            continue;

        if (mat->present && mat->backface_culling) {          /* model.go:902-925 */
            float n0x = tv[0].X - tv[1].X;
            float n0y = tv[0].Y - tv[1].Y;
            float n1x = tv[1].X - tv[2].X;
            float n1y = tv[1].Y - tv[2].Y;
            float nor = (n0x * n1y) - (n1x * n0y);
            if (nor < 0) continue;
        }

        float depth;                                          /* model.go:933-955 */
        if (modelSkinned) {
            float dx = camPos.X - (skinnedTriCenter.X / 3);
            float dy = camPos.Y - (skinnedTriCenter.Y / 3);
            float dz = camPos.Z - (skinnedTriCenter.Z / 3);
            depth = dx * dx + dy * dy + dz * dz;
        } else {
            float dx = invertedCamPos.X - centers[3 * ti + 0];
            float dy = invertedCamPos.Y - centers[3 * ti + 1];
            float dz = invertedCamPos.Z - centers[3 * ti + 2];
            depth = dx * dx + dy * dy + dz * dz;
        }

        if (isnan(depth) || isinf(depth)) continue;           /* model.go:957-960 */

        if (depth < minDepth) minDepth = depth;               /* model.go:962-967 */
        if (depth > maxDepth) maxDepth = depth;

        S->unsetTris[S->unsetTriIndex].tri = ti;              /* AddTriangle, sortingTriangles.go:32-36 */
        S->unsetTris[S->unsetTriIndex].depth = depth;
        S->unsetTriIndex++;
    }

    bucket_sort(S, minDepth, maxDepth);                       /* model.go:989 */
    *outMin = minDepth; *outMax = maxDepth;
}

/* ------------------------------------------------------------------------------------------ */
/* Lights (light.go)                                                                            */
/* ------------------------------------------------------------------------------------------ */

/* beginRender: light.go:110-112 (ambient), 173 (point), 322-324 (sun), 567-572 (volume) */
static int lv_cell_count(const orc_light_volume* lv) { return lv->cell_count[0] * lv->cell_count[1] * lv->cell_count[2]; }

static void light_begin_render(rstate* S, int li) {
    const orc_light* l = &S->sc->lights[li];
    if (l->type == ORC_LIGHT_VOLUME) {                        /* light.go:565-570: every light some cell uses */
        const orc_light_volume* lv = &S->sc->volumes[l->volume];
        if (lv->cell_light_start)
            for (int k = 0; k < lv->cell_light_start[lv_cell_count(lv)]; k++)
                if (S->sc->lights[lv->cell_light_index[k]].type != ORC_LIGHT_VOLUME) light_begin_render(S, lv->cell_light_index[k]);
        return;
    }
    if (l->type == ORC_LIGHT_AMBIENT) {
        S->ambResult[li][0] = l->color[0] * l->energy;
        S->ambResult[li][1] = l->color[1] * l->energy;
        S->ambResult[li][2] = l->color[2] * l->energy;
    } else if (l->type == ORC_LIGHT_DIRECTIONAL) {
        mat4 rot = S_world_rotation(S, l->node);
        S->workingForward[li] = m4_forward(&rot);
    }
}

/* beginModel: light.go:175-189 (point), 326-330 (sun) */
static void light_begin_model(rstate* S, int li, int mi) {
    const orc_light* l = &S->sc->lights[li];
    const orc_model* model = &S->sc->models[mi];
    if (l->type == ORC_LIGHT_VOLUME) {                        /* light.go:574-580 */
        const orc_light_volume* lv = &S->sc->volumes[l->volume];
        if (lv->cell_light_start)
            for (int k = 0; k < lv->cell_light_start[lv_cell_count(lv)]; k++)
                if (S->sc->lights[lv->cell_light_index[k]].type != ORC_LIGHT_VOLUME) light_begin_model(S, lv->cell_light_index[k], mi);
        return;
    }
    if (l->type == ORC_LIGHT_POINT) {
        mat4 inv = m4_inverted(S_transform(S, model->node));
        vec3 pos, sca; mat4 rot;
        m4_decompose(&inv, &pos, &sca, &rot);
        vec3 lp = S_world_position(S, l->node);
        if (model->skinned) S->workingPosition[li] = lp;
        else S->workingPosition[li] = v3_add(m4_multvec(&rot, lp), v3_mult(pos, v3(1 / sca.X, 1 / sca.Y, 1 / sca.Z)));
    } else if (l->type == ORC_LIGHT_DIRECTIONAL) {
        if (!model->skinned)
            S->workingModelRotation[li] = m4_transposed(m4_inverted(S_world_rotation(S, model->node)));
    }
}

/* mesh.go:82-84 */
static int dims_contains(const float* mn, const float* mx, vec3 p) {
    return p.X >= mn[0] && p.Y >= mn[1] && p.Z >= mn[2] && p.X <= mx[0] && p.Y <= mx[1] && p.Z <= mx[2];
}
/* light.go:717-724 */
static void lv_indices(const orc_light_volume* l, vec3 w, int* x, int* y, int* z) {
    *x = (int)m32_round(((w.X - l->offset[0] - l->dim_min[0] - (l->cell_size[0] / 2)) / l->cell_size[0]));
    *y = (int)m32_round(((w.Y - l->offset[1] - l->dim_min[1] - (l->cell_size[1] / 2)) / l->cell_size[1]));
    *z = (int)m32_round(((w.Z - l->offset[2] - l->dim_min[2] - (l->cell_size[2] / 2)) / l->cell_size[2]));
}
/* light.go:1267-1272 */
static const float* lv_cell(const orc_light_volume* l, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= l->cell_count[0] || y >= l->cell_count[1] || z >= l->cell_count[2]) return NULL;
    return l->cell_colors + 3 * ((size_t)x + ((size_t)y * l->cell_count[0]) + ((size_t)z * l->cell_count[0] * l->cell_count[1]));
}

/* VertexLight for every light type. */
static void vertex_light(rstate* S, int li, int index, int mi, const orc_material* mat) {
    const orc_light* l = &S->sc->lights[li];
    const orc_model* model = &S->sc->models[mi];
    vec4* target = &S->vertexLights[index];
    switch (l->type) {
    case ORC_LIGHT_AMBIENT:                                   /* light.go:116-120 */
        target->X += S->ambResult[li][0];
        target->Y += S->ambResult[li][1];
        target->Z += S->ambResult[li][2];
        break;
    case ORC_LIGHT_POINT: {                                   /* light.go:191-258 */
        vec3 vertPos = S->alteredPositions[index];
        vec3 vertNormal = S->alteredNormals[index];
        float lightRangeSquared = l->range * l->range;        /* light.go:277-280 */
        float distance = v3_dist2(S->workingPosition[li], vertPos);
        if (l->range > 0) {
            if (distance > lightRangeSquared) return;
        } else {
            if (1 / distance * l->energy < 0.001f) return;
        }
        vec3 lightVec = v3_unit(v3_sub(S->workingPosition[li], vertPos));
        if (mat->present && mat->lighting_mode == ORC_LIGHTING_FIXED_NORMALS) vertNormal = lightVec;
        float diffuse = v3_dot(vertNormal, lightVec);
        if (mat->present && mat->lighting_mode == ORC_LIGHTING_DOUBLE_SIDED) diffuse = m32_abs(diffuse);
        if (diffuse > 0) {
            float diffuseFactor = diffuse * (1.0f / (1.0f + (0.1f * distance))) * 2;
            if (lightRangeSquared > 0) {
                float distClamp = m32_clamp((lightRangeSquared - distance) / distance, 0, 1);
                diffuseFactor *= distClamp;
            }
            target->X += l->color[0] * diffuseFactor * l->energy;
            target->Y += l->color[1] * diffuseFactor * l->energy;
            target->Z += l->color[2] * diffuseFactor * l->energy;
        }
        break;
    }
    case ORC_LIGHT_DIRECTIONAL: {                             /* light.go:332-360 */
        vec3 normal;
        if (mat->present && mat->lighting_mode == ORC_LIGHTING_FIXED_NORMALS) normal = S->workingForward[li];
        else if (model->skinned) normal = S->alteredNormals[index];
        else normal = m4_multvec(&S->workingModelRotation[li], S->alteredNormals[index]);
        float diffuseFactor = v3_dot(normal, S->workingForward[li]);
        if (mat->present && mat->lighting_mode == ORC_LIGHTING_DOUBLE_SIDED) diffuseFactor = m32_abs(diffuseFactor);
        if (diffuseFactor <= 0) return;
        target->X += l->color[0] * diffuseFactor * l->energy;
        target->Y += l->color[1] * diffuseFactor * l->energy;
        target->Z += l->color[2] * diffuseFactor * l->energy;
        break;
    }
    case ORC_LIGHT_VOLUME: {                                  /* light.go:583-663 */
        const orc_light_volume* lv = &S->sc->volumes[l->volume];
        int objectShadingMode = 0;
        float cellColor[3] = {0, 0, 0};
        mat4 transform = S_transform(S, model->node);
        vec3 cellSize = v3(lv->cell_size[0] / 2, lv->cell_size[1] / 2, lv->cell_size[2] / 2);
        vec3 bias = v3(lv->light_bias[0], lv->light_bias[1], lv->light_bias[2]);
        if (mat->present) objectShadingMode = mat->light_volume_shading_mode;
        if (mat->present && mat->light_volume_shading_mode == ORC_LV_PER_OBJECT) {
            int x, y, z;
            lv_indices(lv, v3(transform.m[3][0], transform.m[3][1], transform.m[3][2]), &x, &y, &z);
            const float* cell = lv_cell(lv, x, y, z);
            if (cell) { cellColor[0] = cell[0]; cellColor[1] = cell[1]; cellColor[2] = cell[2]; }
        } else {
            vec3 vertPos = {0, 0, 0}, vertPosWithNormal = {0, 0, 0};
            if (model->skinned) {
                if (objectShadingMode == ORC_LV_PER_VERTEX_WITH_NORMAL)
                    vertPosWithNormal = v3_add(v3_add(vertPos, v3_mult(S->alteredNormals[index], cellSize)), bias);
                else
                    vertPos = v3_add(S->alteredPositions[index], bias);
            } else {
                if (objectShadingMode == ORC_LV_PER_VERTEX_WITH_NORMAL)
                    vertPosWithNormal = v3_add(m4_multvec(&transform, v3_add(S->alteredPositions[index], v3_mult(S->alteredNormals[index], cellSize))), bias);
                else
                    vertPos = v3_add(m4_multvec(&transform, S->alteredPositions[index]), bias);
            }
            if (!dims_contains(lv->dim_min, lv->dim_max, vertPos)) return;
            int x, y, z;
            if (objectShadingMode == ORC_LV_PER_VERTEX_WITH_NORMAL) lv_indices(lv, vertPosWithNormal, &x, &y, &z);
            else lv_indices(lv, vertPos, &x, &y, &z);
            const float* cell = lv_cell(lv, x, y, z);
            if (cell) {
                cellColor[0] = cell[0]; cellColor[1] = cell[1]; cellColor[2] = cell[2];
                if (lv->cell_light_start) {                   /* light.go:652-654: the cell's own lights, before the cell colour */
                    const size_t c = (size_t)(cell - lv->cell_colors) / 3;
                    for (int k = lv->cell_light_start[c]; k < lv->cell_light_start[c + 1]; k++)
                        if (S->sc->lights[lv->cell_light_index[k]].type != ORC_LIGHT_VOLUME)   /* a volume nested in a volume: not modelled */
                            vertex_light(S, lv->cell_light_index[k], index, mi, mat);
                }
            }
        }
        target->X += l->color[0] * cellColor[0] * l->energy;
        target->Y += l->color[1] * cellColor[1] * l->energy;
        target->Z += l->color[2] * cellColor[2] * l->energy;
        break;
    }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Camera.Render: render() and flush() closures (camera.go:1507-2048)                           */
/* ------------------------------------------------------------------------------------------ */

static void add_render_rec(rstate* S, int mi, int pi, int64_t first, int count, int total, float mn, float mx) {
    orc_result* r = S->res;
    r->renders = realloc(r->renders, (size_t)(r->n_renders + 1) * sizeof(orc_render_rec));
    orc_render_rec* rec = &r->renders[r->n_renders++];
    rec->model = mi; rec->part = pi; rec->flush = S->curFlush; rec->first_vertex = first; rec->vertex_count = count;
    rec->total_tris = total; rec->min_depth = mn; rec->max_depth = mx;
}

static void capture_part(rstate* S, const orc_part* part) {
    orc_result* r = S->res;
    int n = part->vertex_index_end - part->vertex_index_start;
    if (r->n_captured + n > S->cap_captured) {
        int64_t cap = S->cap_captured ? S->cap_captured : 4096;
        while (cap < r->n_captured + n) cap *= 2;
        r->cap_clip = realloc(r->cap_clip, (size_t)cap * 16);
        r->cap_light = realloc(r->cap_light, (size_t)cap * 16);
        S->cap_captured = cap;
    }
    memcpy(r->cap_clip + 4 * r->n_captured, S->vertexTransforms + part->vertex_index_start, (size_t)n * 16);
    memcpy(r->cap_light + 4 * r->n_captured, S->vertexLights + part->vertex_index_start, (size_t)n * 16);
    r->n_captured += n;
}

/* camera.go:1507-1827 */
static void render_pair(rstate* S, int mi, int pi) {
    const orc_scene* sc = S->sc;
    const orc_model* model = &sc->models[mi];
    const orc_mesh* mesh = &sc->meshes[model->mesh];
    const orc_part* meshPart = &mesh->parts[pi];
    const orc_material* mat = &meshPart->material;
    const orc_camera* camera = &sc->camera;
    const orc_world* world = &sc->world;
    orc_result* R = S->res;

    int lighting = 0;                                          /* camera.go:1521-1528 */
    if (world->present) {
        if (mat->present) lighting = world->lighting_on && !mat->shadeless && !model->shadeless;
        else lighting = world->lighting_on && !model->shadeless;
    }

    int totalTris = meshPart->triangle_end - meshPart->triangle_start + 1;
    R->total_tris += totalTris;                                /* camera.go:1532 */
    if (model->batch_owner >= 0) R->batched_parts++;

    S->sortMode = ORC_SORT_BACK_TO_FRONT;                      /* camera.go:1540-1543 */
    if (mat->present) S->sortMode = mat->triangle_sort_mode;

    int startingVertexListIndex = S->vertexListIndex;          /* camera.go:1545 */

    if (S->capture) {   /* sentinels so tests can tell which vertices this render() touched */
        for (int v = meshPart->vertex_index_start; v < meshPart->vertex_index_end; v++) {
            vec4 nanv = {NAN, NAN, NAN, NAN}; vec4 neg = {-1, -1, -1, -1};
            S->vertexTransforms[v] = nanv; S->vertexLights[v] = neg;
        }
    }
    double t0 = now_s();
    float mn, mx;
    process_vertices(S, mi, pi, &mn, &mx);                     /* camera.go:1553 */
    double t1 = now_s();
    R->t_process += t1 - t0;

    /* ProcessVertices leaves vertexListIndex advanced by 3 per surviving triangle. */
    int afterPV = startingVertexListIndex + 3 * S->sortedCount;
    if (afterPV == 0) {                                         /* camera.go:1556-1558 */
        add_render_rec(S, mi, pi, S->flushBase + startingVertexListIndex, 0, totalTris, mn, mx);
        if (S->capture) capture_part(S, meshPart);
        return;
    }

    float srcW = 0, srcH = 0;                                   /* camera.go:1560-1566 */
    if (mat->present && mat->use_texture) { srcW = mat->tex_w; srcH = mat->tex_h; }

    /* LightGroup override (camera.go:1571-1588) */
    int* ogSceneLights = S->sceneLights; int ogN = S->nSceneLights;
    if (lighting) {
        if (model->light_group >= 0 && sc->light_groups[model->light_group].active) {
            const orc_light_group* lg = &sc->light_groups[model->light_group];
            S->sceneLights = (int*)lg->lights; S->nSceneLights = lg->n_lights;
            for (int i = 0; i < S->nSceneLights; i++) light_begin_render(S, S->sceneLights[i]);
        }
        for (int i = 0; i < S->nSceneLights; i++) light_begin_model(S, S->sceneLights[i], mi);
    }

    float maxSpan = dims_max_span(S->dims[model->mesh]);        /* camera.go:1592 */
    vec3 modelPos = S_world_position(S, model->node);

    float mpColor[4] = {model->color[0], model->color[1], model->color[2], model->color[3]};   /* camera.go:1597-1601 */
    if (mat->present) { mpColor[0] *= mat->color[0]; mpColor[1] *= mat->color[1]; mpColor[2] *= mat->color[2]; mpColor[3] *= mat->color[3]; }

    const int camWidth = camera->width, camHeight = camera->height;
    float halfCamWidth = (float)camWidth / 2, halfCamHeight = (float)camHeight / 2;

    float depthMarginPercentage = (camera->far_plane - camera->near_plane) * camera->depth_margin;   /* camera.go:1339-1341 */
    float camSpread = camera->far_plane - camera->near_plane + (depthMarginPercentage * 2);

    S->vertexListIndex = startingVertexListIndex;              /* camera.go:1608 */
    ensure_lists(S, S->flushBase + startingVertexListIndex + 3 * (int64_t)S->sortedCount);

    double t2 = now_s();
    if (lighting) {                                            /* camera.go:1610-1644 */
        for (int k = 0; k < S->sortedCount; k++) {
            const uint32_t* tri = mesh->indices + 3 * (size_t)S->sortedTris[k];
            for (int vi = 0; vi < 3; vi++) {
                vec4* c = &S->vertexLights[tri[vi]];
                c->X = 0; c->Y = 0; c->Z = 0; c->W = 1;
            }
            for (int i = 0; i < S->nSceneLights; i++) {
                int li = S->sceneLights[i];
                const orc_light* light = &sc->lights[li];
                if (light->type == ORC_LIGHT_POINT && light->range > 0) {      /* camera.go:1625-1629 */
                    float dist = maxSpan + light->range;
                    if (v3_dist2(modelPos, S_world_position(S, light->node)) > dist * dist) continue;
                }
                if (light->type == ORC_LIGHT_VOLUME && mat->present && mat->light_volume_shading_mode == ORC_LV_OFF) continue;  /* light.go:671-673 */
                for (int vi = 0; vi < 3; vi++) vertex_light(S, li, (int)tri[vi], mi, mat);
            }
        }
    }
    double t3 = now_s();
    R->t_light += t3 - t2;

    for (int k = 0; k < S->sortedCount; k++) {                 /* camera.go:1646-1820 */
        int ti = S->sortedTris[k];
        const uint32_t* tri = mesh->indices + 3 * (size_t)ti;
        R->tri_ids[(S->flushBase + S->vertexListIndex) / 3] = (uint32_t)ti;
        for (int vi = 0; vi < 3; vi++) {
            int vertexIndex = (int)tri[vi];
            int64_t o = S->flushBase + S->vertexListIndex;
            orc_vertex* cv = &R->color_list[o];
            orc_vertex* dv = &R->depth_list[o];
            orc_vertex* nv = &R->normal_list[o];

            float w = S->vertexTransforms[vertexIndex].W;
            if (!camera->perspective) w = 1.0f;
            if (w < 0) w = 0.001f;
            vec4 target = S->vertexTransforms[vertexIndex];
            float dx = (target.X / w) * (float)camWidth + halfCamWidth;
            float dy = (target.Y / -w) * (float)camHeight + halfCamHeight;
            cv->DstX = dx; cv->DstY = dy; dv->DstX = dx; dv->DstY = dy;

            float uvU, uvV;
            float u = mesh->uvs[2 * vertexIndex], v = mesh->uvs[2 * vertexIndex + 1];
            if (camera->perspective_corrected_texture_mapping) {
                uvU = (u / w) * srcW;
                uvV = ((1 - v) / w) * srcH;
            } else {
                uvU = u * srcW;
                uvV = (1 - v) * srcH;
            }
            cv->SrcX = uvU; cv->SrcY = uvV;
            if (camera->perspective_corrected_texture_mapping) {
                float d = 1.0f / w;
                cv->Custom0 = d; dv->Custom0 = d;
            }
            dv->SrcX = uvU; dv->SrcY = uvV;

            if (camera->render_normals) {                      /* camera.go:1710-1722 */
                nv->DstX = dx; nv->DstY = dy; nv->SrcX = uvU; nv->SrcY = uvV;
                nv->ColorR = S->transformedNormals[vertexIndex].X * 0.5f + 0.5f;
                nv->ColorG = S->transformedNormals[vertexIndex].Y * 0.5f + 0.5f;
                nv->ColorB = S->transformedNormals[vertexIndex].Z * 0.5f + 0.5f;
            }

            if (mesh->active_color_channel >= 0 && mesh->colors) {   /* camera.go:1726-1736 */
                cv->ColorR = mesh->colors[4 * vertexIndex + 0] * mpColor[0];
                cv->ColorG = mesh->colors[4 * vertexIndex + 1] * mpColor[1];
                cv->ColorB = mesh->colors[4 * vertexIndex + 2] * mpColor[2];
                cv->ColorA = mesh->colors[4 * vertexIndex + 3] * mpColor[3];
            } else {
                cv->ColorR = mpColor[0]; cv->ColorG = mpColor[1]; cv->ColorB = mpColor[2]; cv->ColorA = mpColor[3];
            }

            if (lighting) {                                    /* camera.go:1738-1742 */
                cv->ColorR *= S->vertexLights[vertexIndex].X;
                cv->ColorG *= S->vertexLights[vertexIndex].Y;
                cv->ColorB *= S->vertexLights[vertexIndex].Z;
            }

            if (camera->render_depth) {                        /* camera.go:1744-1782 */
                float vertexDepth;
                if (mat->present && mat->billboarded_depth_mode == ORC_DEPTH_MODE_UNBILLBOARDED) vertexDepth = S->depthUnbillboarded[vertexIndex];
                else vertexDepth = S->vertexTransforms[vertexIndex].Z;
                float depth = (vertexDepth + depthMarginPercentage) / camSpread;
                if (depth < 0) depth = 0; else if (depth > 1) depth = 1;
                dv->ColorR = depth; dv->ColorG = depth; dv->ColorB = depth; dv->ColorA = 1;
            } else if (world->present && world->fog_on) {       /* camera.go:1784-1814 */
                float vertexDepth = S->vertexTransforms[vertexIndex].Z;
                float depth = (vertexDepth + depthMarginPercentage) / camSpread;
                if (depth < 0) depth = 0; else if (depth > 1) depth = 1;
                depth = world->fog_range[0] + ((world->fog_range[1] - world->fog_range[0]) * 1 - depth);
                if (world->fog_mode == ORC_FOG_ADD) {
                    cv->ColorR += world->fog_color[0] * depth;
                    cv->ColorG += world->fog_color[1] * depth;
                    cv->ColorB += world->fog_color[2] * depth;
                } else if (world->fog_mode == ORC_FOG_SUB) {
                    cv->ColorR *= world->fog_color[0] * depth;
                    cv->ColorG *= world->fog_color[1] * depth;
                    cv->ColorB *= world->fog_color[2] * depth;
                }
            }

            R->index_list[o] = (uint16_t)S->vertexListIndex;   /* camera.go:1816 */
            S->vertexListIndex++;
        }
    }
    R->t_emit += now_s() - t3;

    add_render_rec(S, mi, pi, S->flushBase + startingVertexListIndex, 3 * S->sortedCount, totalTris, mn, mx);
    if (S->capture) capture_part(S, meshPart);

    S->sceneLights = ogSceneLights; S->nSceneLights = ogN;     /* camera.go:1826 */
}

/* camera.go:1829-2048, minus the Ebitengine draw calls */
static void flush_pair(rstate* S, int mi, int pi) {
    orc_result* R = S->res;
    R->flushes = realloc(R->flushes, (size_t)(R->n_flushes + 1) * sizeof(orc_flush_rec));
    orc_flush_rec* f = &R->flushes[R->n_flushes++];
    f->model = mi; f->part = pi; f->first_vertex = S->flushBase; f->vertex_count = S->vertexListIndex;
    if (S->vertexListIndex != 0) {
        R->drawn_tris += S->vertexListIndex / 3;               /* camera.go:2041-2044 */
        R->drawn_parts++;
    }
    S->flushBase += S->vertexListIndex;
    S->vertexListIndex = 0;
    S->curFlush++;
}

typedef struct { int model, part; float depth; int order; } render_pair_t;   /* camera.go:1249-1254 */

/* stable insertion sort with a Go-style less(i, j) */
static void stable_sort_pairs(render_pair_t* a, int n, int (*less)(const render_pair_t*, const render_pair_t*)) {
    for (int i = 1; i < n; i++) {
        render_pair_t x = a[i]; int j = i - 1;
        while (j >= 0 && less(&x, &a[j])) { a[j + 1] = a[j]; j--; }
        a[j + 1] = x;
    }
}
/* camera.go:1462-1467 */
static int less_solids(const render_pair_t* i, const render_pair_t* j) {
    if (i->order < j->order) return 1;
    return i->depth > j->depth;
}
/* camera.go:2050-2055 */
static int less_transparents(const render_pair_t* i, const render_pair_t* j) {
    if (i->order == j->order) return i->depth > j->depth;
    return i->order < j->order;
}

typedef struct { int model; float dist; } child_t;

/* ---- sort.SliceStable as the Go standard library implements it ---------------------------------------------
 * The Go toolchain (and with it the standard library source) is absent from this image; this restates the
 * algorithm of package sort (zsortfunc.go: stable_func, insertionSort_func, symMerge_func, rotate_func,
 * swapRange_func; unchanged since Go 1.6 apart from renames): insertion sort on blocks of 20, then bottom-up
 * SymMerge.  Needed because the reference feeds it an inconsistent comparator (camera.go:1484-1495), so only
 * the exact sequence of Less / Swap calls defines the result. */
typedef int (*gs_less_fn)(void*, int, int);
typedef void (*gs_swap_fn)(void*, int, int);
static void gs_insertion(int a, int b, gs_less_fn less, gs_swap_fn swap, void* c) {
    for (int i = a + 1; i < b; i++)
        for (int j = i; j > a && less(c, j, j - 1); j--) swap(c, j, j - 1);
}
static void gs_swap_range(int a, int b, int n, gs_swap_fn swap, void* c) { for (int i = 0; i < n; i++) swap(c, a + i, b + i); }
static void gs_rotate(int a, int m, int b, gs_swap_fn swap, void* c) {
    int i = m - a, j = b - m;
    while (i != j) {
        if (i > j) { gs_swap_range(m - i, m, j, swap, c); i -= j; }
        else { gs_swap_range(m - i, m + j - i, i, swap, c); j -= i; }
    }
    gs_swap_range(m - i, m, i, swap, c);
}
static void gs_sym_merge(int a, int m, int b, gs_less_fn less, gs_swap_fn swap, void* c) {
    if (m - a == 1) {
        int i = m, j = b;
        while (i < j) { int h = (int)(((unsigned)i + (unsigned)j) >> 1); if (less(c, h, a)) i = h + 1; else j = h; }
        for (int k = a; k < i - 1; k++) swap(c, k, k + 1);
        return;
    }
    if (b - m == 1) {
        int i = a, j = m;
        while (i < j) { int h = (int)(((unsigned)i + (unsigned)j) >> 1); if (!less(c, m, h)) i = h + 1; else j = h; }
        for (int k = m; k > i; k--) swap(c, k, k - 1);
        return;
    }
    int mid = (int)(((unsigned)a + (unsigned)b) >> 1);
    int n = mid + m;
    int start, r;
    if (m > mid) { start = n - b; r = mid; } else { start = a; r = m; }
    int p = n - 1;
    while (start < r) {
        int cc = (int)(((unsigned)start + (unsigned)r) >> 1);
        if (!less(c, p - cc, cc)) start = cc + 1; else r = cc;
    }
    int end = n - start;
    if (start < m && m < end) gs_rotate(start, m, end, swap, c);
    if (a < start && start < mid) gs_sym_merge(a, start, mid, less, swap, c);
    if (mid < end && end < b) gs_sym_merge(mid, end, b, less, swap, c);
}
static void go_stable_sort(int n, gs_less_fn less, gs_swap_fn swap, void* c) {
    int blockSize = 20;
    int a = 0, b = blockSize;
    while (b <= n) { gs_insertion(a, b, less, swap, c); a = b; b += blockSize; }
    gs_insertion(a, n, less, swap, c);
    while (blockSize < n) {
        a = 0; b = 2 * blockSize;
        while (b <= n) { gs_sym_merge(a, a + blockSize, b, less, swap, c); a = b; b += 2 * blockSize; }
        int m = a + blockSize;
        if (m < n) gs_sym_merge(a, m, n, less, swap, c);
        blockSize *= 2;
    }
}
/* test hook: stable argsort of float keys with go_stable_sort (a consistent comparator, so the answer is unique) */
typedef struct { const float* keys; int* perm; } gst_ctx;
static int gst_less(void* cv, int i, int j) { gst_ctx* c = (gst_ctx*)cv; return c->keys[c->perm[i]] < c->keys[c->perm[j]]; }
static void gst_swap(void* cv, int i, int j) { gst_ctx* c = (gst_ctx*)cv; int t = c->perm[i]; c->perm[i] = c->perm[j]; c->perm[j] = t; }
void orc_test_go_stable_sort(int n, const float* keys, int* perm) {
    gst_ctx c; c.keys = keys; c.perm = perm;
    for (int i = 0; i < n; i++) perm[i] = i;
    go_stable_sort(n, gst_less, gst_swap, &c);
}

/* The same with an arbitrary "less" relation given as an n x n table over the ORIGINAL positions: less(i, j) = table[perm[i] * n + perm[j]].
 * Lets a test compare this restatement of package sort's algorithm with another one on comparators that are not orderings at all. */
typedef struct { const uint8_t* table; int n; int* perm; } gtt_ctx;
static int gtt_less(void* cv, int i, int j) { gtt_ctx* c = (gtt_ctx*)cv; return c->table[(size_t)c->perm[i] * c->n + c->perm[j]] != 0; }
static void gtt_swap(void* cv, int i, int j) { gtt_ctx* c = (gtt_ctx*)cv; int t = c->perm[i]; c->perm[i] = c->perm[j]; c->perm[j] = t; }
void orc_test_go_stable_sort_table(int n, const uint8_t* table, int* perm) {
    gtt_ctx c; c.table = table; c.n = n; c.perm = perm;
    for (int i = 0; i < n; i++) perm[i] = i;
    go_stable_sort(n, gtt_less, gtt_swap, &c);
}

/* camera.go:1460-1469 through sort.SliceStable */
typedef struct { render_pair_t* a; } rp_ctx;
static int rp_less_solids(void* cv, int i, int j) { rp_ctx* c = (rp_ctx*)cv; return less_solids(&c->a[i], &c->a[j]); }
static void rp_swap(void* cv, int i, int j) { rp_ctx* c = (rp_ctx*)cv; render_pair_t t = c->a[i]; c->a[i] = c->a[j]; c->a[j] = t; }

/* the comparator of camera.go:1484-1495 over the sceneLights index list */
typedef struct { rstate* S; int* ids; vec3 pos; } mls_ctx;
static int mls_less(void* cv, int i, int j) {
    mls_ctx* c = (mls_ctx*)cv;
    const orc_light* li = &c->S->sc->lights[c->ids[i]];
    const orc_light* lj = &c->S->sc->lights[c->ids[j]];
    if (li->type == ORC_LIGHT_AMBIENT) return 1;
    else if (lj->type == ORC_LIGHT_AMBIENT) return 0;
    else if (li->type == ORC_LIGHT_DIRECTIONAL) return 1;
    else if (lj->type == ORC_LIGHT_DIRECTIONAL) return 0;
    vec3 pi = li->node >= 0 ? S_world_position(c->S, li->node) : v3(0, 0, 0);
    vec3 pj = lj->node >= 0 ? S_world_position(c->S, lj->node) : v3(0, 0, 0);
    return v3_dist2(c->pos, pi) < v3_dist2(c->pos, pj);
}
static void mls_swap(void* cv, int i, int j) { mls_ctx* c = (mls_ctx*)cv; int t = c->ids[i]; c->ids[i] = c->ids[j]; c->ids[j] = t; }

int orc_render_scene(const orc_scene* sc, int capture, orc_result** out) {
    g_err[0] = 0;
    rstate S; memset(&S, 0, sizeof S);
    S.sc = sc; S.capture = capture;
    S.res = calloc(1, sizeof(orc_result));

    S.nodeT = calloc((size_t)(sc->n_nodes > 0 ? sc->n_nodes : 1), sizeof(mat4));
    S.nodeTValid = calloc((size_t)(sc->n_nodes > 0 ? sc->n_nodes : 1), 1);
    int maxVerts = 1, maxTris = 1;
    S.centers = calloc((size_t)sc->n_meshes, sizeof(float*));
    S.dims = calloc((size_t)sc->n_meshes, sizeof(dims_t));
    for (int i = 0; i < sc->n_meshes; i++) {
        const orc_mesh* m = &sc->meshes[i];
        if (m->n_verts > maxVerts) maxVerts = m->n_verts;
        if (m->n_tris > maxTris) maxTris = m->n_tris;
        S.centers[i] = malloc((size_t)(m->n_tris > 0 ? m->n_tris : 1) * 12);
        if (m->auto_subdivide && m->tri_centers) memcpy(S.centers[i], m->tri_centers, (size_t)m->n_tris * 12);
        else orc_triangle_centers(m->positions, m->indices, m->n_tris, S.centers[i]);
        S.dims[i] = mesh_dimensions(m);
    }
    S.vertexTransforms = calloc((size_t)maxVerts, sizeof(vec4));
    S.transformedNormals = calloc((size_t)maxVerts, sizeof(vec3));
    S.depthUnbillboarded = calloc((size_t)maxVerts, sizeof(float));
    S.alteredPositions = calloc((size_t)maxVerts, sizeof(vec3));
    S.alteredNormals = calloc((size_t)maxVerts, sizeof(vec3));
    S.vertexLights = calloc((size_t)maxVerts, sizeof(vec4));
    S.unsetTris = calloc((size_t)maxTris, sizeof(sorting_tri));
    S.sortedTris = calloc((size_t)maxTris, sizeof(int));
    S.binOf = calloc((size_t)maxTris, sizeof(uint16_t));
    S.binCount = calloc(ORC_BINS, sizeof(int));
    int nl = sc->n_lights > 0 ? sc->n_lights : 1;
    S.ambResult = calloc((size_t)nl, sizeof(float[3]));
    S.workingPosition = calloc((size_t)nl, sizeof(vec3));
    S.workingForward = calloc((size_t)nl, sizeof(vec3));
    S.workingModelRotation = calloc((size_t)nl, sizeof(mat4));
    {   /* the reference's display lists persist across frames (tetra3d.go:21-24); size them once, outside the
         * timed region, for the worst case so that the per-frame time excludes allocation */
        int64_t worst = 0;
        for (int i = 0; i < sc->n_models; i++) if (sc->models[i].mesh >= 0) worst += 3 * (int64_t)sc->meshes[sc->models[i].mesh].n_tris;
        ensure_lists(&S, worst > 30000 ? worst : 30000);
    }
    double tStart = now_s();   /* everything above is load-time work in the reference (mesh.go:1972, 719) */

    orc_result* R = S.res;
    const orc_world* world = &sc->world;

    camera_setup(&S);

    /* sceneLights (camera.go:1275-1303).  Light index n_lights-1 convention: the world ambient light is
     * described by orc_world and materialised here as an extra light slot appended to a private copy. */
    orc_light* lights_copy = malloc((size_t)(sc->n_lights + 1) * sizeof(orc_light));
    if (sc->n_lights) memcpy(lights_copy, sc->lights, (size_t)sc->n_lights * sizeof(orc_light));
    int worldAmbient = sc->n_lights;
    memset(&lights_copy[worldAmbient], 0, sizeof(orc_light));
    lights_copy[worldAmbient].type = ORC_LIGHT_AMBIENT;
    lights_copy[worldAmbient].node = -1;
    memcpy(lights_copy[worldAmbient].color, world->ambient_color, 12);
    lights_copy[worldAmbient].energy = world->ambient_energy;
    lights_copy[worldAmbient].visible = world->ambient_visible;
    orc_scene sc2 = *sc; sc2.lights = lights_copy; sc2.n_lights = sc->n_lights + 1;
    S.sc = &sc2;
    free(S.ambResult); S.ambResult = calloc((size_t)sc2.n_lights, sizeof(float[3]));
    free(S.workingPosition); S.workingPosition = calloc((size_t)sc2.n_lights, sizeof(vec3));
    free(S.workingForward); S.workingForward = calloc((size_t)sc2.n_lights, sizeof(vec3));
    free(S.workingModelRotation); S.workingModelRotation = calloc((size_t)sc2.n_lights, sizeof(mat4));
    sc = &sc2;

    int* sceneLights = malloc((size_t)(sc->n_scene_lights + 1) * sizeof(int));
    int nSceneLights = 0;
    if (world->present) {
        R->light_count++;
        if (world->lighting_on && world->ambient_visible) {
            light_begin_render(&S, worldAmbient);
            sceneLights[nSceneLights++] = worldAmbient;
        }
    }
    for (int i = 0; i < sc->n_scene_lights; i++) {
        int li = sc->scene_lights[i];
        R->light_count++;
        if ((!world->present || world->lighting_on) && sc->lights[li].visible) {
            light_begin_render(&S, li);
            sceneLights[nSceneLights++] = li;
        }
    }
    /* Camera.MaxLightCount (camera.go:1474-1497): sort.SliceStable with the reference's comparator, then truncate.
     * The comparator is not a strict weak ordering (two ambient lights each sort before the other), so the result is
     * whatever Go's algorithm leaves: go_stable_sort restates it. */
    if (sc->camera.max_light_count > 0) {
        mls_ctx mc;
        mc.S = &S; mc.ids = sceneLights;
        mc.pos = sc->camera.max_light_focus_node >= 0 ? S_world_position(&S, sc->camera.max_light_focus_node) : S_world_position(&S, sc->camera.node);
        go_stable_sort(nSceneLights, mls_less, mls_swap, &mc);
        if (sc->camera.max_light_count < nSceneLights) nSceneLights = sc->camera.max_light_count;
    }
    S.sceneLights = sceneLights; S.nSceneLights = nSceneLights;
    for (int i = 0; i < nSceneLights; i++)                     /* camera.go:1499-1505 */
        if (world->present && world->lighting_on && sc->lights[sceneLights[i]].visible) R->active_light_count++;

    /* build solids / transparents (camera.go:1343-1457) */
    render_pair_t* solids = malloc((size_t)(sc->n_models * 8 + 8) * sizeof(render_pair_t)); int nSolids = 0, capS = sc->n_models * 8 + 8;
    render_pair_t* transparents = malloc((size_t)(sc->n_models * 8 + 8) * sizeof(render_pair_t)); int nTrans = 0, capT = sc->n_models * 8 + 8;
    /* which models own dynamic batches */
    uint8_t* isBatcher = calloc((size_t)sc->n_models + 1, 1);
    for (int i = 0; i < sc->n_models; i++) if (sc->models[i].batch_owner >= 0) isBatcher[sc->models[i].batch_owner] = 1;

    for (int mi = 0; mi < sc->n_models; mi++) {
        const orc_model* model = &sc->models[mi];
        if (!model->visible || model->batch_owner >= 0) continue;
        const orc_mesh* mesh = model->mesh >= 0 ? &sc->meshes[model->mesh] : NULL;
        if (!isBatcher[mi]) {
            if (mesh) R->total_parts += mesh->n_parts;          /* camera.go:1350-1358 */
            if (model->frustum_culling && !model_in_frustum(&S, mi)) continue;
            if (!mesh) continue;
            vec3 mpos = S_world_position(&S, model->node);
            for (int pi = 0; pi < mesh->n_parts; pi++) {
                const orc_part* mp = &mesh->parts[pi];
                if (!part_is_visible(mp)) continue;
                render_pair_t rp = {mi, pi, v3_dist2(S.camPos, mpos), 0};
                if (mp->material.present) rp.order = mp->material.render_order;
                if (is_transparent(model, mp)) {
                    if (nTrans == capT) { capT *= 2; transparents = realloc(transparents, (size_t)capT * sizeof(render_pair_t)); }
                    transparents[nTrans++] = rp;
                } else {
                    if (nSolids == capS) { capS *= 2; solids = realloc(solids, (size_t)capS * sizeof(render_pair_t)); }
                    solids[nSolids++] = rp;
                }
            }
        } else {
            /* camera.go:1398-1451.  The reference ranges over a Go map here (unordered); the contract
             * fixes ascending MeshPart index of the owner. */
            int transparent = 0;
            for (int pi = 0; pi < mesh->n_parts; pi++) {
                const orc_part* meshPart = &mesh->parts[pi];
                int nChildren = 0;
                for (int ci = 0; ci < sc->n_models; ci++)
                    if (sc->models[ci].batch_owner == mi && sc->models[ci].batch_owner_part == pi) nChildren++;
                if (nChildren == 0) continue;                   /* no map entry */
                if (!part_is_visible(meshPart)) continue;
                for (int ci = 0; ci < sc->n_models; ci++) {
                    const orc_model* child = &sc->models[ci];
                    if (child->batch_owner != mi || child->batch_owner_part != pi) continue;
                    if (!child->visible) continue;
                    if (!transparent) {
                        const orc_mesh* cm = &sc->meshes[child->mesh];
                        for (int cp = 0; cp < cm->n_parts; cp++)
                            if (part_is_visible(&cm->parts[cp]) && is_transparent(child, &cm->parts[cp])) { transparent = 1; break; }
                    }
                }
                render_pair_t rp = {mi, pi, 0, 0};
                if (transparent || !sc->camera.render_depth) rp.depth = v3_dist2(S.camPos, S_world_position(&S, model->node));
                if (meshPart->material.present) rp.order = meshPart->material.render_order;
                if (transparent) {
                    if (nTrans == capT) { capT *= 2; transparents = realloc(transparents, (size_t)capT * sizeof(render_pair_t)); }
                    transparents[nTrans++] = rp;
                } else {
                    if (nSolids == capS) { capS *= 2; solids = realloc(solids, (size_t)capS * sizeof(render_pair_t)); }
                    solids[nSolids++] = rp;
                }
                R->total_parts += nChildren;                    /* camera.go:1447-1449 */
            }
        }
    }

    /* camera.go:1460-1469: the comparator is not a strict weak ordering when RenderOrder values differ, so the result is
     * defined by Go's sort.SliceStable alone: go_stable_sort restates it */
    if (!sc->camera.render_depth) { rp_ctx rc; rc.a = solids; go_stable_sort(nSolids, rp_less_solids, rp_swap, &rc); }
    stable_sort_pairs(transparents, nTrans, less_transparents);                         /* camera.go:2050-2055 */

    child_t* kids = malloc((size_t)(sc->n_models + 1) * sizeof(child_t));
    for (int pass = 0; pass < 2; pass++) {                      /* camera.go:2057-2109 */
        render_pair_t* list = pass == 0 ? solids : transparents;
        int n = pass == 0 ? nSolids : nTrans;
        for (int k = 0; k < n; k++) {
            render_pair_t pair = list[k];
            if (!sc->models[pair.model].visible) continue;
            if (isBatcher[pair.model]) {
                /* camera.go:2075-2100.  sort.Slice is not stable; the contract breaks ties by original order. */
                int nk = 0;
                for (int ci = 0; ci < sc->n_models; ci++) {
                    const orc_model* child = &sc->models[ci];
                    if (child->batch_owner != pair.model || child->batch_owner_part != pair.part) continue;
                    kids[nk].model = ci;
                    kids[nk].dist = v3_dist2(S.camPos, S_world_position(&S, child->node));   /* node.go:1720-1722 */
                    nk++;
                }
                for (int i = 1; i < nk; i++) {                  /* stable, descending distance */
                    child_t x = kids[i]; int j = i - 1;
                    while (j >= 0 && x.dist > kids[j].dist) { kids[j + 1] = kids[j]; j--; }
                    kids[j + 1] = x;
                }
                for (int i = 0; i < nk; i++) {
                    const orc_model* merged = &sc->models[kids[i].model];
                    if (!merged->visible) continue;
                    if (merged->frustum_culling && !model_in_frustum(&S, kids[i].model)) continue;
                    const orc_mesh* cm = &sc->meshes[merged->mesh];
                    for (int cp = 0; cp < cm->n_parts; cp++) render_pair(&S, kids[i].model, cp);
                }
                flush_pair(&S, pair.model, pair.part);
            } else {
                render_pair(&S, pair.model, pair.part);
                flush_pair(&S, pair.model, pair.part);
            }
        }
    }

    R->n_vertices = S.flushBase;
    R->t_total = now_s() - tStart;

    free(kids); free(solids); free(transparents); free(isBatcher); free(sceneLights); free(lights_copy);
    for (int i = 0; i < sc->n_meshes; i++) free(S.centers[i]);
    free(S.centers); free(S.dims); free(S.nodeT); free(S.nodeTValid);
    free(S.vertexTransforms); free(S.transformedNormals); free(S.depthUnbillboarded);
    free(S.alteredPositions); free(S.alteredNormals); free(S.vertexLights);
    free(S.unsetTris); free(S.sortedTris); free(S.binOf); free(S.binCount);
    free(S.ambResult); free(S.workingPosition); free(S.workingForward); free(S.workingModelRotation);
    *out = R;
    return ORC_OK;
}

/* ================================================================================================
 * Model.BakeLighting (model.go:1203-1268)
 * ================================================================================================ */
/* rgb: [n_verts][3] of the model's mesh, in: VertexColors[targetChannel] before the call, out: after it (alpha is untouched by
 * the reference).  light_ids: the `lights` argument, in order; the scene's World.AmbientLight is appended (model.go:1216-1218). */
int orc_bake_lighting(const orc_scene* sc, int32_t mi, const int32_t* light_ids, int32_t n_ids, float* rgb) {
    g_err[0] = 0;
    if (mi < 0 || mi >= sc->n_models || sc->models[mi].mesh < 0) { snprintf(g_err, sizeof g_err, "orc_bake_lighting: bad model"); return ORC_E_INVALID; }
    rstate S; memset(&S, 0, sizeof S);
    S.nodeT = calloc((size_t)(sc->n_nodes > 0 ? sc->n_nodes : 1), sizeof(mat4));
    S.nodeTValid = calloc((size_t)(sc->n_nodes > 0 ? sc->n_nodes : 1), 1);
    int maxVerts = 1, maxTris = 1;
    S.centers = calloc((size_t)sc->n_meshes, sizeof(float*));
    S.dims = calloc((size_t)sc->n_meshes, sizeof(dims_t));
    for (int i = 0; i < sc->n_meshes; i++) {
        const orc_mesh* m = &sc->meshes[i];
        if (m->n_verts > maxVerts) maxVerts = m->n_verts;
        if (m->n_tris > maxTris) maxTris = m->n_tris;
        S.centers[i] = malloc((size_t)(m->n_tris > 0 ? m->n_tris : 1) * 12);
        if (m->auto_subdivide && m->tri_centers) memcpy(S.centers[i], m->tri_centers, (size_t)m->n_tris * 12);
        else orc_triangle_centers(m->positions, m->indices, m->n_tris, S.centers[i]);
        S.dims[i] = mesh_dimensions(m);
    }
    S.vertexTransforms = calloc((size_t)maxVerts, sizeof(vec4));
    S.transformedNormals = calloc((size_t)maxVerts, sizeof(vec3));
    S.depthUnbillboarded = calloc((size_t)maxVerts, sizeof(float));
    S.alteredPositions = calloc((size_t)maxVerts, sizeof(vec3));
    S.alteredNormals = calloc((size_t)maxVerts, sizeof(vec3));
    S.vertexLights = calloc((size_t)maxVerts, sizeof(vec4));        /* stands in for VertexColors[targetChannel] */
    S.unsetTris = calloc((size_t)maxTris, sizeof(sorting_tri));
    S.sortedTris = calloc((size_t)maxTris, sizeof(int));
    S.binOf = calloc((size_t)maxTris, sizeof(uint16_t));
    S.binCount = calloc(ORC_BINS, sizeof(int));
    /* the world's ambient light as one more light slot, as in orc_render_scene */
    orc_light* lights_copy = malloc((size_t)(sc->n_lights + 1) * sizeof(orc_light));
    if (sc->n_lights) memcpy(lights_copy, sc->lights, (size_t)sc->n_lights * sizeof(orc_light));
    const int worldAmbient = sc->n_lights;
    memset(&lights_copy[worldAmbient], 0, sizeof(orc_light));
    lights_copy[worldAmbient].type = ORC_LIGHT_AMBIENT; lights_copy[worldAmbient].node = -1;
    memcpy(lights_copy[worldAmbient].color, sc->world.ambient_color, 12);
    lights_copy[worldAmbient].energy = sc->world.ambient_energy; lights_copy[worldAmbient].visible = sc->world.ambient_visible;
    orc_scene sc2 = *sc; sc2.lights = lights_copy; sc2.n_lights = sc->n_lights + 1;
    S.sc = &sc2; sc = &sc2;
    S.ambResult = calloc((size_t)sc->n_lights, sizeof(float[3]));
    S.workingPosition = calloc((size_t)sc->n_lights, sizeof(vec3));
    S.workingForward = calloc((size_t)sc->n_lights, sizeof(vec3));
    S.workingModelRotation = calloc((size_t)sc->n_lights, sizeof(mat4));
    S.res = calloc(1, sizeof(orc_result));
    camera_setup(&S);

    int* lightsInput = malloc((size_t)(n_ids + 1) * sizeof(int));   /* model.go:1212-1218 */
    int nIn = 0;
    for (int i = 0; i < n_ids; i++) lightsInput[nIn++] = light_ids[i];
    if (sc->world.present) lightsInput[nIn++] = worldAmbient;
    for (int i = 0; i < nIn; i++)                                   /* model.go:1220-1235 */
        if (sc->lights[lightsInput[i]].visible) { light_begin_render(&S, lightsInput[i]); light_begin_model(&S, lightsInput[i], mi); }

    const orc_model* model = &sc->models[mi];
    const orc_mesh* mesh = &sc->meshes[model->mesh];
    for (int v = 0; v < mesh->n_verts; v++) { S.vertexLights[v].X = rgb[3 * v]; S.vertexLights[v].Y = rgb[3 * v + 1]; S.vertexLights[v].Z = rgb[3 * v + 2]; }
    for (int pi = 0; pi < mesh->n_parts; pi++) {                    /* model.go:1237-1266 */
        const orc_part* part = &mesh->parts[pi];
        const orc_material* mat = &part->material;
        float mn, mx;
        process_vertices(&S, mi, pi, &mn, &mx);                     /* ProcessVertices(vp, camera, mp, false, nil) */
        if (mat->present && mat->shadeless) {
            for (int v = part->vertex_index_start; v < part->vertex_index_end; v++) { S.vertexLights[v].X = 1; S.vertexLights[v].Y = 1; S.vertexLights[v].Z = 1; }
            continue;
        }
        for (int v = part->vertex_index_start; v < part->vertex_index_end; v++) { S.vertexLights[v].X = 0; S.vertexLights[v].Y = 0; S.vertexLights[v].Z = 0; }
        for (int i = 0; i < nIn; i++) {
            if (!sc->lights[lightsInput[i]].visible) continue;
            for (int ti = part->triangle_start; ti <= part->triangle_end; ti++) {     /* forEachTri(true, ...): light.Light(tri, ...) */
                const uint32_t* tri = mesh->indices + 3 * (size_t)ti;
                for (int vi = 0; vi < 3; vi++) vertex_light(&S, lightsInput[i], (int)tri[vi], mi, mat);   /* light.go:260-269 */
            }
        }
    }
    for (int v = 0; v < mesh->n_verts; v++) { rgb[3 * v] = S.vertexLights[v].X; rgb[3 * v + 1] = S.vertexLights[v].Y; rgb[3 * v + 2] = S.vertexLights[v].Z; }

    free(lightsInput); free(lights_copy); free(S.res);
    for (int i = 0; i < sc->n_meshes; i++) free(S.centers[i]);
    free(S.centers); free(S.dims); free(S.nodeT); free(S.nodeTValid);
    free(S.vertexTransforms); free(S.transformedNormals); free(S.depthUnbillboarded);
    free(S.alteredPositions); free(S.alteredNormals); free(S.vertexLights);
    free(S.unsetTris); free(S.sortedTris); free(S.binOf); free(S.binCount);
    free(S.ambResult); free(S.workingPosition); free(S.workingForward); free(S.workingModelRotation);
    return ORC_OK;
}

void orc_free_result(orc_result* r) {
    if (!r) return;
    free(r->renders); free(r->flushes); free(r->color_list); free(r->depth_list); free(r->normal_list);
    free(r->index_list); free(r->tri_ids); free(r->cap_clip); free(r->cap_light);
    free(r);
}

/* ================================================================================================
 * Bone poses: animation.go, quaternion.go, node.go:572-594
 * ================================================================================================ */

/* Go's math.Sin for amd64 (pure Go, src/math/sin.go; Cephes sinl.c coefficients).  Arguments reach it from Slerp in
 * [0, pi], far below the Payne-Hanek threshold (reduceThreshold = 1 << 29), which is therefore not restated: larger
 * arguments fall back to libm.  The hexadecimal literals are the bit patterns Go's source lists next to each constant. */
static const double GO_SIN[6] = {
    1.58962301576546568060e-10,   /* 0x3de5d8fd1fd19ccd */
    -2.50507477628578072866e-8,   /* 0xbe5ae5e5a9291f5d */
    2.75573136213857245213e-6,    /* 0x3ec71de3567d48a1 */
    -1.98412698295895385996e-4,   /* 0xbf2a01a019bfdf03 */
    8.33333333332211858878e-3,    /* 0x3f8111111110f7d0 */
    -1.66666666666666307295e-1,   /* 0xbfc5555555555548 */
};
static const double GO_COS[6] = {
    -1.13585365213876817300e-11,  /* 0xbda8fa49a0861a9b */
    2.08757008419747316778e-9,    /* 0x3e21ee9d7b4e3f05 */
    -2.75573141792967388112e-7,   /* 0xbe927e4f7eac4bc6 */
    2.48015872888517045348e-5,    /* 0x3efa01a019c844f5 */
    -1.38888888888730564116e-3,   /* 0xbf56c16c16c14f91 */
    4.16666666666665929218e-2,    /* 0x3fa555555555554b */
};
#define GO_PI 3.14159265358979323846264338327950288419716939937510582097494459
double orc_go_sin(double x) {
    const double PI4A = 7.85398125648498535156e-1;   /* 0x3fe921fb40000000 */
    const double PI4B = 3.77489470793079817668e-8;   /* 0x3e64442d00000000 */
    const double PI4C = 2.69515142907905952645e-15;  /* 0x3ce8469898cc5170 */
    if (x == 0 || isnan(x)) return x;
    if (isinf(x)) return NAN;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    if (x >= 536870912.0) return sign ? -sin(x) : sin(x);     /* reduceThreshold: trigReduce is not restated */
    uint64_t j = (uint64_t)(x * (4 / GO_PI));        /* integer part of x / (Pi/4) */
    double y = (double)j;
    if (j & 1) { j++; y++; }                         /* map zeros to origin */
    j &= 7;                                          /* octant modulo 2 Pi */
    double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;   /* extended precision modular arithmetic */
    if (j > 3) { sign = !sign; j -= 4; }
    const double zz = z * z;
    if (j == 1 || j == 2)
        y = 1.0 - 0.5 * zz + zz * zz * ((((((GO_COS[0] * zz) + GO_COS[1]) * zz + GO_COS[2]) * zz + GO_COS[3]) * zz + GO_COS[4]) * zz + GO_COS[5]);
    else
        y = z + z * zz * ((((((GO_SIN[0] * zz) + GO_SIN[1]) * zz + GO_SIN[2]) * zz + GO_SIN[3]) * zz + GO_SIN[4]) * zz + GO_SIN[5]);
    return sign ? -y : y;
}

/* Go's math.Atan (src/math/atan.go: xatan / satan, Cephes atan.c). */
static double go_xatan(double x) {
    const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01, P2 = -7.500855792314704667340e+01,
                 P3 = -1.228866684490136173410e+02, P4 = -6.485021904942025371773e+01;
    const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02, Q2 = +4.328810604912902668951e+02,
                 Q3 = +4.853903996359136964868e+02, Q4 = +1.945506571482613964425e+02;
    double z = x * x;
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
    z = x * z + x;
    return z;
}
static double go_satan(double x) {
    const double Morebits = 6.123233995736765886130e-17;   /* pi/2 = PIO2 + Morebits */
    const double Tan3pio8 = 2.41421356237309504880;        /* tan(3*pi/8) */
    if (x <= 0.66) return go_xatan(x);
    if (x > Tan3pio8) return GO_PI / 2 - go_xatan(1 / x) + Morebits;
    return GO_PI / 4 + go_xatan((x - 1) / (x + 1)) + 0.5 * Morebits;
}
static double go_atan(double x) {
    if (x == 0) return x;
    if (x > 0) return go_satan(x);
    return -go_satan(-x);
}
double orc_go_atan2(double y, double x) {              /* src/math/atan2.go */
    if (isnan(y) || isnan(x)) return NAN;
    if (y == 0) {
        if (x >= 0 && !signbit(x)) return copysign(0, y);
        return copysign(GO_PI, y);
    }
    if (x == 0) return copysign(GO_PI / 2, y);
    if (isinf(x)) {
        if (x > 0) return isinf(y) ? copysign(GO_PI / 4, y) : copysign(0, y);
        return isinf(y) ? copysign(3 * GO_PI / 4, y) : copysign(GO_PI, y);
    }
    if (isinf(y)) return copysign(GO_PI / 2, y);
    const double q = go_atan(y / x);
    if (x < 0) return q <= 0 ? q + GO_PI : q - GO_PI;
    return q;
}

typedef struct { float X, Y, Z, W; } quat;

static quat quat_slerp(quat a, quat end, float percent) {          /* quaternion.go:26-59 */
    float cosOmega = a.X * end.X + a.Y * end.Y + a.Z * end.Z + a.W * end.W;      /* Dot, quaternion.go:82-84 */
    if (cosOmega < 0) { end.W *= -1; end.X *= -1; end.Y *= -1; end.Z *= -1; cosOmega *= -1; }
    float k0, k1;
    if (cosOmega > 0.9999f) { k0 = 1 - percent; k1 = percent; }
    else {
        const float sinOmega = m32_sqrt(1 - (cosOmega * cosOmega));
        const float omega = (float)orc_go_atan2((double)sinOmega, (double)cosOmega);
        const float oneOverSinOmega = 1 / sinOmega;
        k0 = (float)orc_go_sin((double)((1 - percent) * omega)) * oneOverSinOmega;
        k1 = (float)orc_go_sin((double)(percent * omega)) * oneOverSinOmega;
    }
    quat r;
    r.W = a.W * k0 + end.W * k1;
    r.X = a.X * k0 + end.X * k1;
    r.Y = a.Y * k0 + end.Y * k1;
    r.Z = a.Z * k0 + end.Z * k1;
    return r;
}

static mat4 quat_to_matrix(quat q) {                               /* quaternion.go:132-178 */
    mat4 m1 = m4_identity(), m2 = m4_identity();
    m1.m[0][0] = q.W;  m1.m[0][1] = q.Z;  m1.m[0][2] = -q.Y; m1.m[0][3] = q.X;
    m1.m[1][0] = -q.Z; m1.m[1][1] = q.W;  m1.m[1][2] = q.X;  m1.m[1][3] = q.Y;
    m1.m[2][0] = q.Y;  m1.m[2][1] = -q.X; m1.m[2][2] = q.W;  m1.m[2][3] = q.Z;
    m1.m[3][0] = -q.X; m1.m[3][1] = -q.Y; m1.m[3][2] = -q.Z; m1.m[3][3] = q.W;
    m2.m[0][0] = q.W;  m2.m[0][1] = q.Z;  m2.m[0][2] = -q.Y; m2.m[0][3] = -q.X;
    m2.m[1][0] = -q.Z; m2.m[1][1] = q.W;  m2.m[1][2] = q.X;  m2.m[1][3] = -q.Y;
    m2.m[2][0] = q.Y;  m2.m[2][1] = -q.X; m2.m[2][2] = q.W;  m2.m[2][3] = -q.Z;
    m2.m[3][0] = q.X;  m2.m[3][1] = q.Y;  m2.m[3][2] = q.Z;  m2.m[3][3] = q.W;
    return m4_mult(m1, m2);
}

void orc_quat_slerp(const float* a4, const float* b4, float percent, float* out4) {
    quat a = {a4[0], a4[1], a4[2], a4[3]}, b = {b4[0], b4[1], b4[2], b4[3]};
    quat r = quat_slerp(a, b, percent);
    out4[0] = r.X; out4[1] = r.Y; out4[2] = r.Z; out4[3] = r.W;
}
void orc_quat_to_matrix(const float* q4, float* out16) {
    quat q = {q4[0], q4[1], q4[2], q4[3]};
    mat4 m = quat_to_matrix(q);
    m4_store(&m, out16);
}

/* The keyframe pair around `time` (animation.go:63-90, 116-141): returns 0 = no keyframes, 1 = a single value (*a),
 * 2 = interpolate between *a and *b. */
static int track_keys(const orc_animation* an, const orc_track* tr, float time, int* a, int* b) {
    if (tr->n_keys == 0) return 0;
    const float* T = an->key_times + tr->first_key;
    if (time <= T[0]) { *a = tr->first_key; return 1; }
    if (time >= T[tr->n_keys - 1]) { *a = tr->first_key + tr->n_keys - 1; return 1; }
    int first = -1, last = -1;
    for (int k = 0; k < tr->n_keys; k++) {
        if (T[k] < time) first = k;
        else { last = k; break; }
    }
    if (time == T[first]) { *a = tr->first_key + first; return 1; }
    if (time == T[last]) { *a = tr->first_key + last; return 1; }
    *a = tr->first_key + first; *b = tr->first_key + last;
    if (tr->interpolation == 1) return 1;                           /* InterpolationConstant: the earlier keyframe */
    return 2;
}

static int track_vector(const orc_animation* an, const orc_track* tr, float time, vec3* out) {      /* animation.go:63-112 */
    int a = 0, b = 0;
    const int k = track_keys(an, tr, time, &a, &b);
    if (k == 0) return 0;
    const float* fa = an->key_values + 4 * (size_t)a;
    vec3 fd = v3(fa[0], fa[1], fa[2]);
    if (k == 2) {
        const float* fb = an->key_values + 4 * (size_t)b;
        const float t = (time - an->key_times[a]) / (an->key_times[b] - an->key_times[a]);
        fd.X = fd.X + ((fb[0] - fd.X) * t);                          /* Vector3.Lerp, vector.go:559-565 */
        fd.Y = fd.Y + ((fb[1] - fd.Y) * t);
        fd.Z = fd.Z + ((fb[2] - fd.Z) * t);
    }
    *out = fd;
    return 1;
}

static int track_quaternion(const orc_animation* an, const orc_track* tr, float time, quat* out) {  /* animation.go:116-164 */
    int a = 0, b = 0;
    const int k = track_keys(an, tr, time, &a, &b);
    if (k == 0) return 0;
    const float* fa = an->key_values + 4 * (size_t)a;
    quat fd = {fa[0], fa[1], fa[2], fa[3]};
    if (k == 2) {
        const float* fb = an->key_values + 4 * (size_t)b;
        quat ld = {fb[0], fb[1], fb[2], fb[3]};
        const float t = (time - an->key_times[a]) / (an->key_times[b] - an->key_times[a]);
        fd = quat_slerp(fd, ld, t);
    }
    *out = fd;
    return 1;
}

static int m4_is_zero(const mat4* a) {                              /* matrix.go:700-709 */
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) if (a->m[i][j] != 0) return 0;
    return 1;
}

/* Quaternion.Unit (quaternion.go:97-104): four divisions by Magnitude (quaternion.go:86-95) */
static quat quat_unit(quat q) {
    const float m = m32_sqrt((q.X * q.X) + (q.Y * q.Y) + (q.Z * q.Z) + (q.W * q.W));
    q.X /= m; q.Y /= m; q.Z /= m; q.W /= m;
    return q;
}

/* What AnimationPlayer.updateValues records for one node (AnimationValues, animation.go:674-722). */
typedef struct bone_sample { vec3 p, s; quat q; int P, S, R; } bone_sample;

static int has_channel(const orc_animation* an, int b) {      /* a channel exists for the bone: any of its tracks holds keyframes */
    if (!an || b >= an->n_bones) return 0;
    return an->tracks[3 * b].n_keys > 0 || an->tracks[3 * b + 1].n_keys > 0 || an->tracks[3 * b + 2].n_keys > 0;
}
static bone_sample sample_bone(const orc_animation* an, int b, float time) {
    bone_sample v; memset(&v, 0, sizeof v);
    if (!an || b >= an->n_bones) return v;
    v.P = track_vector(an, &an->tracks[3 * b + 0], time, &v.p);
    v.S = track_vector(an, &an->tracks[3 * b + 1], time, &v.s);
    v.R = track_quaternion(an, &an->tracks[3 * b + 2], time, &v.q);
    return v;
}
/* forceUpdate's blend of one node (animation.go:859-897): start = prevAnimatedProperties[node], props = this frame's samples. */
static bone_sample blend_sample(bone_sample start, bone_sample props, float bp) {
    bone_sample r = props;
    if (start.P) {
        if (props.P) { r.p.X = start.p.X + ((props.p.X - start.p.X) * bp); r.p.Y = start.p.Y + ((props.p.Y - start.p.Y) * bp); r.p.Z = start.p.Z + ((props.p.Z - start.p.Z) * bp); }   /* :863-866 */
        else { r.p = start.p; r.P = 1; }
    }
    if (start.S) {
        if (props.S) { r.s.X = start.s.X + ((props.s.X - start.s.X) * bp); r.s.Y = start.s.Y + ((props.s.Y - start.s.Y) * bp); r.s.Z = start.s.Z + ((props.s.Z - start.s.Z) * bp); }   /* :875-878 */
        else { r.s = start.s; r.S = 1; }
    }
    if (start.R) {
        if (props.R) r.q = quat_unit(quat_slerp(start.q, props.q, bp));                                                                                                       /* :887-889 */
        else { r.q = start.q; r.R = 1; }
    }
    return r;
}

/* One bone's local position / scale / rotation after AnimationPlayer.forceUpdate (animation.go:837-985), for a player whose
 * state is the chain of animations levels[0 .. n): levels[n - 1] is playing; every level i >= 1 was Play()ed with BlendTime > 0
 * while level i - 1's targets were in currentProperties, and -- for i < n - 1 -- was itself interrupted by the next Play() before
 * its blend ended, with playheads[i] / bps[i] the values of its last update.  The reference keeps two things per node, and so
 * does this fold: what currentProperties holds (`cur`: defined only for nodes the level's animation has a channel for, because
 * assignChannels clears the map after every Play(), animation.go:618-626) and what the Node itself holds (every property a level
 * set stays until a later level sets it again).  Level i's targets are blend(cur, samples of level i, bps[i]) where cur is defined,
 * the plain samples where it is not (prevExists false, animation.go:843).  With one level this is the plain pose; with two, a blend
 * between two animations. */
static void pose_local_chain(const orc_skeleton* sk, int b, int n, const orc_animation* const* levels, const float* playheads, const float* bps,
                             vec3* pos, vec3* scl, mat4* rot, int32_t* set3 /* NULL, or: which of the three a level of the chain set */) {
    if (set3) set3[0] = set3[1] = set3[2] = 0;
    *pos = v3(sk->rest_position[3 * b], sk->rest_position[3 * b + 1], sk->rest_position[3 * b + 2]);
    *scl = v3(sk->rest_scale[3 * b], sk->rest_scale[3 * b + 1], sk->rest_scale[3 * b + 2]);
    *rot = m4_load(sk->rest_rotation + 16 * (size_t)b);
    bone_sample cur; memset(&cur, 0, sizeof cur);
    int curDefined = 0;
    for (int k = 0; k < n; k++) {
        if (!has_channel(levels[k], b)) { curDefined = 0; continue; }
        bone_sample v = sample_bone(levels[k], b, playheads[k]);
        if (k > 0 && curDefined) v = blend_sample(cur, v, bps[k]);
        cur = v; curDefined = 1;                                /* animation.go:909-937 */
        if (v.P) { *pos = v.p; if (set3) set3[0] = 1; }         /* animation.go:939-983 without relative motion */
        if (v.S) { *scl = v.s; if (set3) set3[1] = 1; }
        if (v.R) {
            const mat4 r = quat_to_matrix(v.q);
            if (!m4_is_zero(&r)) { *rot = r; if (set3) set3[2] = 1; }   /* SetLocalRotation ignores an all-zero matrix (node.go:906-912) */
        }
    }
}

static void pose_local(const orc_skeleton* sk, int b, const orc_animation* an, float playhead,
                       const orc_animation* from, float from_playhead, float bp, vec3* pos, vec3* scl, mat4* rot) {
    const orc_animation* levels[2] = {from, an};
    const float playheads[2] = {from_playhead, playhead}, bps[2] = {0, bp};
    pose_local_chain(sk, b, 2, levels, playheads, bps, pos, scl, rot, NULL);
}

static void pose_bones(const orc_skeleton* sk, const orc_animation* an, float playhead, const orc_animation* from, float from_playhead,
                       float bp, const float* root16, float* out) {
    const int n = sk->n_bones;
    mat4* world = (mat4*)malloc(sizeof(mat4) * (size_t)(n > 0 ? n : 1));
    const mat4 root = m4_load(root16);
    for (int b = 0; b < n; b++) {
        /* AnimationPlayer.updateValues + forceUpdate (animation.go:674-722, 837-980): tracks that exist overwrite the
         * node's local position / scale / rotation */
        vec3 pos, scl; mat4 rot;
        pose_local(sk, b, an, playhead, from, from_playhead, bp, &pos, &scl, &rot);
        /* Node.Transform (node.go:572-594) */
        mat4 t = m4_scale(scl.X, scl.Y, scl.Z);
        t = m4_mult(t, rot);
        t = m4_mult(t, m4_translate(pos.X, pos.Y, pos.Z));
        t = m4_mult(t, sk->parent[b] >= 0 ? world[sk->parent[b]] : root);
        world[b] = t;
        const mat4 infl = m4_mult(m4_load(sk->inverse_bind + 16 * (size_t)b), t);
        m4_store(&infl, out + 16 * (size_t)b);
    }
    free(world);
}

void orc_pose_bones(const orc_skeleton* sk, const orc_animation* an, float playhead, const float* root16, float* out) {
    pose_bones(sk, an, playhead, NULL, 0, 0, root16, out);
}

/* A player whose blends were interrupted by further Play() calls: see pose_local_chain. */
void orc_pose_bones_chain(const orc_skeleton* sk, int n_levels, const orc_animation* const* levels, const float* playheads, const float* bps,
                          const float* root16, float* out) {
    const int n = sk->n_bones;
    mat4* world = (mat4*)malloc(sizeof(mat4) * (size_t)(n > 0 ? n : 1));
    const mat4 root = m4_load(root16);
    for (int b = 0; b < n; b++) {
        vec3 pos, scl; mat4 rot;
        pose_local_chain(sk, b, n_levels, levels, playheads, bps, &pos, &scl, &rot, NULL);
        mat4 t = m4_scale(scl.X, scl.Y, scl.Z);                 /* Node.Transform (node.go:572-594) */
        t = m4_mult(t, rot);
        t = m4_mult(t, m4_translate(pos.X, pos.Y, pos.Z));
        t = m4_mult(t, sk->parent[b] >= 0 ? world[sk->parent[b]] : root);
        world[b] = t;
        const mat4 infl = m4_mult(m4_load(sk->inverse_bind + 16 * (size_t)b), t);
        m4_store(&infl, out + 16 * (size_t)b);
    }
    free(world);
}
/* The same fold, stopping at the bones' local values: position[3], scale[3], rotation[16] per bone (what the Node holds), and
 * set3[3] per bone: whether a level of the chain set the property (otherwise the value is the skeleton's rest value, where the
 * reference's Node keeps whatever an animation before the chain left there). */
void orc_pose_locals_chain(const orc_skeleton* sk, int n_levels, const orc_animation* const* levels, const float* playheads, const float* bps,
                           float* pos3, float* scale3, float* rot16, int32_t* set3) {
    for (int b = 0; b < sk->n_bones; b++) {
        vec3 pos, scl; mat4 rot;
        pose_local_chain(sk, b, n_levels, levels, playheads, bps, &pos, &scl, &rot, set3 ? set3 + 3 * b : NULL);
        pos3[3 * b] = pos.X; pos3[3 * b + 1] = pos.Y; pos3[3 * b + 2] = pos.Z;
        scale3[3 * b] = scl.X; scale3[3 * b + 1] = scl.Y; scale3[3 * b + 2] = scl.Z;
        m4_store(&rot, rot16 + 16 * (size_t)b);
    }
}
/* One bone's three tracks sampled at `time` (AnimationValues of animation.go:674-722): flags3 = {PositionExists, ScaleExists,
 * RotationExists}, then position, scale, quaternion (x y z w). */
void orc_sample_bone(const orc_animation* an, int bone, float time, int32_t* flags3, float* pos3, float* scale3, float* quat4) {
    const bone_sample v = sample_bone(an, bone, time);
    flags3[0] = v.P; flags3[1] = v.S; flags3[2] = v.R;
    pos3[0] = v.p.X; pos3[1] = v.p.Y; pos3[2] = v.p.Z;
    scale3[0] = v.s.X; scale3[1] = v.s.Y; scale3[2] = v.s.Z;
    quat4[0] = v.q.X; quat4[1] = v.q.Y; quat4[2] = v.q.Z; quat4[3] = v.q.W;
}

/* A player in the middle of a blend (BlendTime > 0, animation.go:852-907): `from` (may be NULL: the previous state animated
 * nothing) sampled at from_playhead is the start, `an` at `playhead` the target, bp the elapsed fraction. */
void orc_pose_bones_blended(const orc_skeleton* sk, const orc_animation* from, float from_playhead, const orc_animation* an, float playhead,
                            float bp, const float* root16, float* out) {
    pose_bones(sk, an, playhead, from, from_playhead, bp, root16, out);
}

/* ================================================================================================
 * Model.BakeAO, the pass over the model's own triangles (model.go:1040-1124)
 * ================================================================================================ */
/* Go's math.Asin / math.Acos for amd64 (pure Go, src/math/asin.go): Asin through satan (atan.go, restated above). */
static double go_asin(double x) {
    if (x == 0) return x;
    int sign = 0;
    if (x < 0) { x = -x; sign = 1; }
    if (x > 1) return NAN;
    double temp = sqrt(1 - x * x);
    if (x > 0.7) temp = 1.57079632679489661923132169163975144 /* Pi/2 */ - go_satan(temp / x);
    else temp = go_satan(x / temp);
    return sign ? -temp : temp;
}
static double go_acos(double x) { return 1.57079632679489661923132169163975144 - go_asin(x); }
double orc_go_acos(double x) { return go_acos(x); }

/* Vector3.Angle (vector.go:468-478) */
static float v3_angle(vec3 vec, vec3 other) {
    const float eps = 1e-4f;
    if (!(m32_abs(other.X) > eps || m32_abs(other.Y) > eps || m32_abs(other.Z) > eps)) return 0;   /* other.IsZero(), vector.go:428-438 */
    float d = v3_dot(vec, other);
    d /= (vec.X * vec.X + vec.Y * vec.Y + vec.Z * vec.Z) * (other.X * other.X + other.Y * other.Y + other.Z * other.Z);
    d = m32_clamp(d, -1, 1);
    return (float)go_acos((double)d);                     /* math32.Acos */
}

/* colors4: VertexColors[TargetChannel], [n_verts][4], in / out.  target_parts: one byte per MeshPart (NULL = all, an empty
 * TargetMeshParts); target_vertices: one byte per vertex (NULL = all, an empty TargetVertices). */
int orc_bake_ao(const orc_mesh* mesh, const float* tri_normals3, const float* tri_max_span, const uint8_t* target_parts,
                const uint8_t* target_vertices, float occlusion_angle, const float* occlusion_color4, float* colors4) {
    const int nt = mesh->n_tris;
    float* centers = malloc((size_t)(nt > 0 ? nt : 1) * 12);
    orc_triangle_centers(mesh->positions, mesh->indices, nt, centers);
    int* partOf = malloc((size_t)(nt > 0 ? nt : 1) * sizeof(int));
    for (int t = 0; t < nt; t++) partOf[t] = -1;
    for (int p = 0; p < mesh->n_parts; p++)
        for (int t = mesh->parts[p].triangle_start; t <= mesh->parts[p].triangle_end && t < nt; t++) partOf[t] = p;
    const vec3* P = (const vec3*)mesh->positions;
    for (int ti = 0; ti < nt; ti++) {                         /* model.go:1052 */
        if (target_parts && !(partOf[ti] >= 0 && target_parts[partOf[ti]])) continue;      /* model.go:1054-1066 */
        const uint32_t* tri = mesh->indices + 3 * (size_t)ti;
        float ao[3] = {0, 0, 0};
        const vec3 triCenter = v3(centers[3 * ti], centers[3 * ti + 1], centers[3 * ti + 2]);
        const vec3 triNormal = v3(tri_normals3[3 * ti], tri_normals3[3 * ti + 1], tri_normals3[3 * ti + 2]);
        for (int oi = 0; oi < nt; oi++) {                     /* model.go:1070 */
            if (oi == ti) continue;
            const uint32_t* other = mesh->indices + 3 * (size_t)oi;
            float span = (tri_max_span[ti] > tri_max_span[oi] ? tri_max_span[ti] : tri_max_span[oi]) * 0.66f;
            if (v3_dist2(triCenter, v3(centers[3 * oi], centers[3 * oi + 1], centers[3 * oi + 2])) > span * span) continue;
            float angle = v3_angle(triNormal, v3(tri_normals3[3 * oi], tri_normals3[3 * oi + 1], tri_normals3[3 * oi + 2]));
            if (angle < occlusion_angle) continue;
            /* SharesVertexPositions, mesh.go:2050-2081 */
            int shared[3] = {-1, -1, -1}, count = 0;
            for (int index = 0; index < 3; index++)
                for (int otherIndex = 0; otherIndex < 3; otherIndex++)
                    if (v3_equals(P[tri[index]], P[other[otherIndex]])) { shared[index] = otherIndex; count++; break; }
            if (count > 0) { if (shared[0] >= 0) ao[0] = 1; if (shared[1] >= 0) ao[1] = 1; if (shared[2] >= 0) ao[2] = 1; }
            if (ao[0] == 1 && ao[1] == 1 && ao[2] == 1) break;
        }
        for (int i = 0; i < 3; i++) {                         /* model.go:1101-1116 */
            float p = m32_clamp(ao[i], 0, 1);
            int vertexIndex = (int)tri[i];
            if (target_vertices && !target_vertices[vertexIndex]) continue;
            float* c = colors4 + 4 * (size_t)vertexIndex;
            c[0] += (occlusion_color4[0] - c[0]) * p;
            c[1] += (occlusion_color4[1] - c[1]) * p;
            c[2] += (occlusion_color4[2] - c[2]) * p;
            c[3] += (occlusion_color4[3] - c[3]) * p;
        }
    }
    free(centers); free(partOf);
    return ORC_OK;
}
