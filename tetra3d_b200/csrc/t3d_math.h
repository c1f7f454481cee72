// t3d_math.h -- float32 vector / matrix math shared by the device kernels and the host mirror.
//
// Operand order follows the reference so that results are bit-identical to the Go path when built
// without multiply-add contraction (nvcc -fmad=false, g++ -ffp-contract=off): matrix_amd64.s:4-81 uses
// separate MULPS/ADDPS and Go/amd64 (GOAMD64=v1) never fuses.
#pragma once

#include <math.h>
#include <stdint.h>
#include <float.h>

#ifdef __CUDACC__
#define T3D_HD __host__ __device__ __forceinline__
#define T3D_UNROLL _Pragma("unroll")
#else
#define T3D_HD inline
#define T3D_UNROLL
#endif

namespace t3d {

struct Vec3 { float X, Y, Z; };
struct Mat4 { float m[4][4]; };

// ---- Vector3 (vector.go) ----------------------------------------------------------------------
T3D_HD Vec3 v3(float x, float y, float z) { Vec3 v; v.X = x; v.Y = y; v.Z = z; return v; }
T3D_HD Vec3 v3_sub(Vec3 a, Vec3 b) { return v3(a.X - b.X, a.Y - b.Y, a.Z - b.Z); }
T3D_HD Vec3 v3_add(Vec3 a, Vec3 b) { return v3(a.X + b.X, a.Y + b.Y, a.Z + b.Z); }
T3D_HD Vec3 v3_mult(Vec3 a, Vec3 b) { return v3(a.X * b.X, a.Y * b.Y, a.Z * b.Z); }
T3D_HD Vec3 v3_invert(Vec3 a) { return v3(-a.X, -a.Y, -a.Z); }
T3D_HD float v3_dot(Vec3 a, Vec3 b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }          // vector.go:526
T3D_HD float v3_magnitude(Vec3 v) { return sqrtf(v.X * v.X + v.Y * v.Y + v.Z * v.Z); }     // vector.go:191
T3D_HD float v3_dist2(Vec3 a, Vec3 b) {                                                    // vector.go:240
    a.X -= b.X; a.Y -= b.Y; a.Z -= b.Z;
    return a.X * a.X + a.Y * a.Y + a.Z * a.Z;
}
T3D_HD Vec3 v3_unit(Vec3 v) {                                                              // vector.go:258
    float l = v3_magnitude(v);
    if (l < 1e-8f || l == 1.0f) return v;
    v.X = v.X / l; v.Y = v.Y / l; v.Z = v.Z / l;
    return v;
}
T3D_HD Vec3 v3_cross(Vec3 vec, Vec3 other) {                                               // vector.go:164
    float ogY = vec.Y, ogZ = vec.Z;
    vec.Z = vec.X * other.Y - other.X * vec.Y;
    vec.Y = ogZ * other.X - other.Z * vec.X;
    vec.X = ogY * other.Z - other.Y * ogZ;
    return vec;
}
T3D_HD bool v3_equals(Vec3 a, Vec3 b) {                                                    // vector.go:415
    const float eps = 1e-4f;
    return !(fabsf(a.X - b.X) > eps || fabsf(a.Y - b.Y) > eps || fabsf(a.Z - b.Z) > eps);
}
T3D_HD float clampf(float v, float lo, float hi) {                                         // math32.go:39
    if (v < lo) return lo;
    else if (v > hi) return hi;
    return v;
}

// ---- Matrix4 (matrix.go) ----------------------------------------------------------------------
T3D_HD Mat4 m4_identity() {
    Mat4 r;
T3D_UNROLL
    for (int i = 0; i < 4; i++)
T3D_UNROLL
        for (int j = 0; j < 4; j++) r.m[i][j] = (i == j) ? 1.0f : 0.0f;
    return r;
}
T3D_HD Mat4 m4_load(const float* p) {
    Mat4 r;
T3D_UNROLL
    for (int i = 0; i < 16; i++) r.m[i >> 2][i & 3] = p[i];
    return r;
}
T3D_HD void m4_store(const Mat4& a, float* p) {
T3D_UNROLL
    for (int i = 0; i < 16; i++) p[i] = a.m[i >> 2][i & 3];
}
// matrix_amd64.s:4-81 / matrix_none.go:6-32
T3D_HD Mat4 m4_mult(const Mat4& a, const Mat4& b) {
    Mat4 r;
T3D_UNROLL
    for (int i = 0; i < 4; i++)
T3D_UNROLL
        for (int j = 0; j < 4; j++)
            r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j] + a.m[i][3] * b.m[3][j];
    return r;
}
T3D_HD Mat4 m4_transposed(const Mat4& a) {                                                 // matrix.go:218
    Mat4 r;
T3D_UNROLL
    for (int i = 0; i < 4; i++)
T3D_UNROLL
        for (int j = 0; j < 4; j++) r.m[i][j] = a.m[j][i];
    return r;
}
T3D_HD Mat4 m4_scale(float x, float y, float z) { Mat4 r = m4_identity(); r.m[0][0] = x; r.m[1][1] = y; r.m[2][2] = z; return r; }
T3D_HD Vec3 m4_multvec(const Mat4& M, Vec3 v) {                                            // matrix.go:619
    Vec3 r;
    r.X = M.m[0][0] * v.X + M.m[1][0] * v.Y + M.m[2][0] * v.Z + M.m[3][0];
    r.Y = M.m[0][1] * v.X + M.m[1][1] * v.Y + M.m[2][1] * v.Z + M.m[3][1];
    r.Z = M.m[0][2] * v.X + M.m[1][2] * v.Y + M.m[2][2] * v.Z + M.m[3][2];
    return r;
}
// Vector4.Unit applied to a matrix row (vector.go:1869): W untouched.
T3D_HD void m4_unit_row(Mat4& a, int i) {
    float l = sqrtf(a.m[i][0] * a.m[i][0] + a.m[i][1] * a.m[i][1] + a.m[i][2] * a.m[i][2]);
    if (l < 1e-8f || l == 1.0f) return;
    a.m[i][0] = a.m[i][0] / l; a.m[i][1] = a.m[i][1] / l; a.m[i][2] = a.m[i][2] / l;
}
T3D_HD Mat4 m4_decompose_rotation(const Mat4& a) {                                         // matrix.go:198
    Mat4 rot = m4_identity();
T3D_UNROLL
    for (int i = 0; i < 3; i++) {
T3D_UNROLL
        for (int j = 0; j < 4; j++) rot.m[i][j] = a.m[i][j];
        m4_unit_row(rot, i);
    }
    return rot;
}
T3D_HD void m4_decompose(const Mat4& a, Vec3& position, Vec3& scale, Mat4& rotation) {     // matrix.go:181
    position = v3(a.m[3][0], a.m[3][1], a.m[3][2]);
    rotation = m4_decompose_rotation(a);
    Mat4 in = m4_mult(a, m4_transposed(rotation));
    scale.X = sqrtf(in.m[0][0] * in.m[0][0] + in.m[0][1] * in.m[0][1] + in.m[0][2] * in.m[0][2]);
    scale.Y = sqrtf(in.m[1][0] * in.m[1][0] + in.m[1][1] * in.m[1][1] + in.m[1][2] * in.m[1][2]);
    scale.Z = sqrtf(in.m[2][0] * in.m[2][0] + in.m[2][1] * in.m[2][1] + in.m[2][2] * in.m[2][2]);
}
T3D_HD Mat4 m4_inverted(const Mat4& a) {                                                   // matrix.go:255
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
    det = 1.0f / det;
    Mat4 m;
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
T3D_HD Mat4 m4_lookat(Vec3 from, Vec3 to, Vec3 up) {                                       // matrix.go:788
    if (v3_equals(from, to)) return m4_identity();
    Vec3 z = v3_unit(v3_sub(to, from));
    up = v3_unit(up);
    if (v3_equals(z, up) || v3_equals(z, v3_invert(up))) {
        if (!v3_equals(up, v3(1, 0, 0))) up = v3(1, 0, 0);
        else up = v3(0, 0, 1);
    }
    Vec3 x = v3_unit(v3_cross(up, z));
    Vec3 y = v3_cross(z, x);
    Mat4 r = m4_identity();
    r.m[0][0] = x.X; r.m[0][1] = x.Y; r.m[0][2] = x.Z;
    r.m[1][0] = y.X; r.m[1][1] = y.Y; r.m[1][2] = y.Z;
    r.m[2][0] = z.X; r.m[2][1] = z.Y; r.m[2][2] = z.Z;
    return r;
}


// Model.onTransformUpdate's culling sphere (model.go:142-182) from the model's transform and its mesh's Dimensions
// {Center, Min, Max}, then Camera.SphereInFrustum (camera.go:933-983).  Host (frames not shipped as runs) and device
// (k_prep_draws) share this one sequence of operations.
struct FrustumTerms { Vec3 camPos, forward, right, up; float nearP, farP, sphereFactorY, sphereFactorX, sphereFactorTang, aspect, orthoScale; bool perspective; };
T3D_HD bool m4_instance_in_frustum(const Mat4& transform, const float* bounds9, const FrustumTerms& F) {
    Vec3 position, scale; Mat4 rotation;
    m4_decompose(transform, position, scale, rotation);
    Vec3 center = m4_multvec(rotation, v3(bounds9[0], bounds9[1], bounds9[2]));
    position.X += center.X * scale.X; position.Y += center.Y * scale.Y; position.Z += center.Z * scale.Z;
    const Vec3 mn = v3(bounds9[3] * scale.X, bounds9[4] * scale.Y, bounds9[5] * scale.Z);
    const Vec3 mx = v3(bounds9[6] * scale.X, bounds9[7] * scale.Y, bounds9[8] * scale.Z);
    const float radius = v3_magnitude(v3_sub(mx, mn)) / 2;                  // Dimensions.MaxSpan() / 2
    const Vec3 diff = v3_sub(position, F.camPos);
    float pcZ = v3_dot(diff, v3_invert(F.forward));
    if (pcZ > F.farP + radius || pcZ < F.nearP - radius) return false;
    if (F.perspective) {
        float d = F.sphereFactorY * radius;
        pcZ *= F.sphereFactorTang;
        const float pcY = v3_dot(diff, F.up);
        if (pcY > pcZ + d || pcY < -pcZ - d) return false;
        pcZ *= F.aspect;
        d = F.sphereFactorX * radius;
        const float pcX = v3_dot(diff, F.right);
        if (pcX > pcZ + d || pcX < -pcZ - d) return false;
    } else {
        const float w = F.orthoScale, h = w / F.aspect;
        const float pcY = v3_dot(diff, F.up);
        if (-h / 2 - radius > pcY || pcY > h / 2 + radius) return false;
        const float pcX = v3_dot(diff, F.right);
        if (-w / 2 - radius > pcX || pcX > w / 2 + radius) return false;
    }
    return true;
}

}  // namespace t3d
