// t3d_bake.cu -- Model.BakeAO, the pass over the model's own triangles (model.go:1040-1124), on the device.
//
// The reference tests every triangle of the target MeshParts against every triangle of the mesh -- close enough (centres within
// 0.66 x the larger MaxSpan), bent enough (angle between the face normals at least OcclusionAngle), sharing vertex POSITIONS --
// and marks the corners it shares; then it walks the triangles once more and pulls each corner's colour towards OcclusionColor
// by the mark, in place, so a vertex that several triangles refer to is pulled once per corner, in triangle order.  Here:
//   k_ao_marks   one thread per triangle: the O(T^2) scan (every thread of a warp reads the same other triangle: broadcast loads);
//                the marks are ORs, so the reference's early exit is only an optimisation and the result does not depend on order
//   k_ao_apply   one thread per vertex: replays the pulls of its corners in triangle order (an O(V x T) scan of the index planes,
//                again broadcast) -- the same sequence of float32 operations the reference applies to that vertex
// math32.Acos is float32(math.Acos(float64(x))): Go's pure-Go asin.go (Pi/2 - Asin; Asin through satan, t3d_pose.cuh).
#include "t3d_internal.h"
#include "t3d_pose.cuh"
#include "../../include/t3dgeo.h"

namespace t3d {

__device__ inline double go_asin(double x) {        // src/math/asin.go
    if (x == 0.0) return x;
    bool sign = false;
    if (x < 0.0) { x = -x; sign = true; }
    if (x > 1.0) return f64(0x7ff8000000000001ull);
    const double PI_2 = f64(0x3ff921fb54442d18ull);
    double temp = sqrt(1.0 - x * x);
    if (x > 0.7) temp = PI_2 - go_satan(temp / x);
    else temp = go_satan(x / temp);
    return sign ? -temp : temp;
}
__device__ inline double go_acos(double x) { return f64(0x3ff921fb54442d18ull) - go_asin(x); }

// Vector3.Angle (vector.go:468-478)
__device__ inline float v3_angle(Vec3 vec, Vec3 other) {
    const float eps = 1e-4f;
    if (!(fabsf(other.X) > eps || fabsf(other.Y) > eps || fabsf(other.Z) > eps)) return 0.0f;       // other.IsZero()
    float d = v3_dot(vec, other);
    d /= (vec.X * vec.X + vec.Y * vec.Y + vec.Z * vec.Z) * (other.X * other.X + other.Y * other.Y + other.Z * other.Z);
    d = clampf(d, -1.0f, 1.0f);
    return (float)go_acos((double)d);
}

__global__ void __launch_bounds__(256)
k_ao_marks(const float* __restrict__ pos, uint32_t vstride, const uint32_t* __restrict__ idx, const float* __restrict__ ctr, uint32_t tstride, uint32_t nTris,
           const float* __restrict__ triNormal, const float* __restrict__ triSpan, const uint8_t* __restrict__ triTarget, float occlusionAngle,
           uint8_t* __restrict__ marks) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= nTris) return;
    if (triTarget && !triTarget[i]) { marks[i] = 0; return; }
    const Vec3 ci = v3(ctr[i], ctr[tstride + i], ctr[2 * (size_t)tstride + i]);
    const Vec3 ni = v3(triNormal[3 * (size_t)i], triNormal[3 * (size_t)i + 1], triNormal[3 * (size_t)i + 2]);
    const float si = triSpan[i];
    Vec3 P[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { const uint32_t v = idx[k * (size_t)tstride + i]; P[k] = v3(pos[v], pos[vstride + v], pos[2 * (size_t)vstride + v]); }
    uint32_t ao = 0;
    for (uint32_t j = 0; j < nTris && ao != 7u; j++) {
        if (j == i) continue;
        const float span = fmaxf(si, triSpan[j]) * 0.66f;                   // model.go:1071
        const Vec3 cj = v3(ctr[j], ctr[tstride + j], ctr[2 * (size_t)tstride + j]);
        if (v3_dist2(ci, cj) > span * span) continue;
        const Vec3 nj = v3(triNormal[3 * (size_t)j], triNormal[3 * (size_t)j + 1], triNormal[3 * (size_t)j + 2]);
        if (v3_angle(ni, nj) < occlusionAngle) continue;                    // model.go:1077-1080
        Vec3 Q[3];
#pragma unroll
        for (int k = 0; k < 3; k++) { const uint32_t v = idx[k * (size_t)tstride + j]; Q[k] = v3(pos[v], pos[vstride + v], pos[2 * (size_t)vstride + v]); }
#pragma unroll
        for (int a = 0; a < 3; a++)                                         // SharesVertexPositions, mesh.go:2050-2081
            if (v3_equals(P[a], Q[0]) || v3_equals(P[a], Q[1]) || v3_equals(P[a], Q[2])) ao |= 1u << a;
    }
    marks[i] = (uint8_t)ao;
}

__global__ void __launch_bounds__(256)
k_ao_apply(const uint32_t* __restrict__ idx, uint32_t tstride, uint32_t nTris, uint32_t nVerts, const uint8_t* __restrict__ triTarget,
           const uint8_t* __restrict__ marks, const uint8_t* __restrict__ vertTarget, float4 oc, float4* __restrict__ colors) {
    const uint32_t v = blockIdx.x * 256u + threadIdx.x;
    if (v >= nVerts) return;
    if (vertTarget && !vertTarget[v]) return;                               // model.go:1106-1108
    float4 c = colors[v];
    for (uint32_t t = 0; t < nTris; t++) {
        if (triTarget && !triTarget[t]) continue;
        const uint32_t m = marks[t];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (idx[k * (size_t)tstride + t] != v) continue;
            const float p = clampf((m >> k) & 1u ? 1.0f : 0.0f, 0.0f, 1.0f);   // model.go:1102
            c.x += (oc.x - c.x) * p; c.y += (oc.y - c.y) * p; c.z += (oc.z - c.z) * p; c.w += (oc.w - c.w) * p;
        }
    }
    colors[v] = c;
}

void launch_ao(cudaStream_t s, const float* pos, uint32_t vstride, const uint32_t* idx, const float* ctr, uint32_t tstride, uint32_t nTris, uint32_t nVerts,
               const float* triNormal, const float* triSpan, const uint8_t* triTarget, const uint8_t* vertTarget, float occlusionAngle,
               const float* occlusionColor4, uint8_t* marks, float4* colors) {
    if (!nTris || !nVerts) return;
    launch_kernel(k_ao_marks, (nTris + 255) / 256, 256, 0, s, false, pos, vstride, idx, ctr, tstride, nTris, triNormal, triSpan, triTarget, occlusionAngle, marks);
    launch_kernel(k_ao_apply, (nVerts + 255) / 256, 256, 0, s, false, idx, tstride, nTris, nVerts, triTarget, (const uint8_t*)marks, vertTarget,
                  make_float4(occlusionColor4[0], occlusionColor4[1], occlusionColor4[2], occlusionColor4[3]), colors);
}

}  // namespace t3d
