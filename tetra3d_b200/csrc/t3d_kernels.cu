// t3d_kernels.cu -- sm_100a kernels of the Tetra3D geometry stage.
//
//   k_prep_draws  one thread per draw      per-draw uniforms of Model.ProcessVertices (model.go:616-708) and
//                                           light.beginModel (light.go:175-189, 326-330)
//   k_vertex      4 vertices per thread     fused skin + transform + light (model.go:559-596, 757-822;
//                                           light.go:116-120, 191-258, 332-360, 583-663)
//   k_cull        one thread per triangle   near/far, side, backface cull + depth key + per-draw min/max
//                                           (model.go:855-969)
//   k_keys        one thread per triangle   512-bin key (sortingTriangles.go:38-64)
//   k_emit        one thread per corner     colour/depth/normal vertex lists (camera.go:1646-1820)
//
// Built with -fmad=false: no multiply-add contraction anywhere (see t3d_device.cuh).
#include "t3d_device.cuh"
#include "t3d_internal.h"
#include "../../include/t3dgeo.h"

namespace t3d {

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------

// Largest d in [lo, hi] with base[d] <= slot.  lo/hi come from the per-tile table the host builds,
// so for draws larger than a tile the loop body never runs.
__device__ __forceinline__ uint32_t find_draw(const uint32_t* __restrict__ base, uint32_t lo, uint32_t hi, uint32_t slot) {
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(&base[mid]) <= slot) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------------------------
// K0: per-draw uniforms
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_prep_draws(DrawRec* __restrict__ draws, uint32_t nDraws, const __grid_constant__ FrameConst fc,
             const t3d_light* __restrict__ lightsIn, LightU* __restrict__ lightsU, DrawStat* __restrict__ stats) {
    uint32_t d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nDraws) return;
    DrawRec& R = draws[d];

    const uint32_t flags = R.flags;
    const bool skinned = flags & T3D_DRAW_SKINNED;
    Mat4 modelT = m4_load(R.transform);
    Mat4 vp = m4_load(fc.vp);
    Vec3 camPos = v3(fc.camPos[0], fc.camPos[1], fc.camPos[2]);

    // model.go:629-630
    Vec3 p, s; Mat4 r;
    { Mat4 inv = m4_inverted(modelT); m4_decompose(inv, p, s, r); }
    Vec3 invPS = v3_mult(p, v3(1.0f / s.X, 1.0f / s.Y, 1.0f / s.Z));
    Vec3 invCam = v3_add(m4_multvec(r, camPos), invPS);
    R.invCam[0] = invCam.X; R.invCam[1] = invCam.Y; R.invCam[2] = invCam.Z;

    Mat4 base = modelT;
    if (flags & T3D_DRAW_BILLBOARD) {                                   // model.go:634-693
        Vec3 up = v3(0, 1, 0);
        if (flags & T3D_DRAW_BILLBOARD_CAMERA_UP) up = v3(fc.camUp[0], fc.camUp[1], fc.camUp[2]);
        Mat4 lookat = m4_lookat(v3(modelT.m[3][0], modelT.m[3][1], modelT.m[3][2]), camPos, up);
        if (flags & T3D_DRAW_BILLBOARD_LOCK_X) {
            lookat.m[0][0] = 1; lookat.m[0][1] = 0; lookat.m[0][2] = 0; lookat.m[0][3] = 0;
            lookat.m[1][0] = 0; m4_unit_row(lookat, 1);
            lookat.m[2][0] = 0; m4_unit_row(lookat, 2);
        }
        if (flags & T3D_DRAW_BILLBOARD_LOCK_Y) {
            lookat.m[1][0] = 0; lookat.m[1][1] = 1; lookat.m[1][2] = 0; lookat.m[1][3] = 0;
            lookat.m[0][1] = 0; m4_unit_row(lookat, 0);
            lookat.m[2][1] = 0; m4_unit_row(lookat, 2);
        }
        if (flags & T3D_DRAW_BILLBOARD_LOCK_Z) {
            lookat.m[2][0] = 0; lookat.m[2][1] = 0; lookat.m[2][2] = 1; lookat.m[2][3] = 0;
            lookat.m[0][2] = 0; m4_unit_row(lookat, 0);
            lookat.m[1][2] = 0; m4_unit_row(lookat, 1);
        }
        Vec3 bp, bs; Mat4 br;
        m4_decompose(base, bp, bs, br);
        base = m4_mult(m4_mult(br, m4_scale(bs.X, bs.Y, bs.Z)), lookat);
        base.m[3][0] = bp.X; base.m[3][1] = bp.Y; base.m[3][2] = bp.Z; base.m[3][3] = 1;
    }
    Mat4 mvp = m4_mult(base, vp);                                       // model.go:695
    m4_store(mvp, R.mvp);

    if (flags & T3D_DRAW_DEPTH_UNBILLBOARDED) {                         // model.go:700-703
        Mat4 u = m4_mult(modelT, vp);
        R.unbill[0] = u.m[0][2]; R.unbill[1] = u.m[1][2]; R.unbill[2] = u.m[2][2]; R.unbill[3] = u.m[3][2];
    }
    if (fc.camFlags & T3D_CAM_RENDER_NORMALS) {                         // model.go:705-708
        Mat4 mv = m4_mult(modelT, m4_load(fc.view));
        Mat4 rot = m4_decompose_rotation(mv);
#pragma unroll
        for (int c = 0; c < 3; c++)
#pragma unroll
            for (int rr = 0; rr < 4; rr++) R.nrmRot[c * 4 + rr] = rot.m[rr][c];
    }

    // Lights: beginModel products + per-draw skip (camera.go:1625-1629, light.go:671-673).
    const uint32_t nL = (flags & T3D_DRAW_LIT) ? R.lightCount : 0;
    if (nL) {
        Mat4 sunRot = m4_identity();
        bool haveSun = false;
        Vec3 modelPos = v3(modelT.m[3][0], modelT.m[3][1], modelT.m[3][2]);
        const float maxSpan = R.maxSpan;
        for (uint32_t i = 0; i < nL; i++) {
            const t3d_light L = lightsIn[R.lightIn + i];
            LightU U;
            U.type = L.type;
            U.color[0] = L.color[0]; U.color[1] = L.color[1]; U.color[2] = L.color[2];
            U.energy = L.energy;
            U.wpos[0] = U.wpos[1] = U.wpos[2] = 0;
            U.range = L.range; U.range2 = L.range * L.range;             // light.go:277-280
            U.fwd[0] = L.forward[0]; U.fwd[1] = L.forward[1]; U.fwd[2] = L.forward[2];
            U.volume = L.volume; U.pad[0] = U.pad[1] = 0;
            if (L.type == T3D_LIGHT_POINT) {
                Vec3 lp = v3(L.position[0], L.position[1], L.position[2]);
                if (L.range > 0) {                                       // camera.go:1625-1629
                    float dist = maxSpan + L.range;
                    if (v3_dist2(modelPos, lp) > dist * dist) U.type = LIGHT_SKIP;
                }
                Vec3 wp = skinned ? lp : v3_add(m4_multvec(r, lp), invPS);   // light.go:183-187
                U.wpos[0] = wp.X; U.wpos[1] = wp.Y; U.wpos[2] = wp.Z;
            } else if (L.type == T3D_LIGHT_DIRECTIONAL) {
                if (!skinned && !haveSun) {                              // light.go:326-330
                    sunRot = m4_transposed(m4_inverted(m4_load(R.worldRot)));
                    haveSun = true;
                }
            } else if (L.type == T3D_LIGHT_VOLUME) {
                if ((flags & T3D_DRAW_HAS_MATERIAL) && R.lvMode == T3D_LV_OFF) U.type = LIGHT_SKIP;   // light.go:671-673
            }
            lightsU[R.lightU + i] = U;
        }
        // The sun's workingModelRotation is per model; park it in worldRot (in place).
        if (haveSun) m4_store(sunRot, R.worldRot);
    }

    DrawStat st; st.minBits = 0x7F7FFFFFu; st.maxBits = 0u; st.survivors = 0; st.pad = 0;   // model.go:710-711
    stats[d] = st;
}

// ------------------------------------------------------------------------------------------------
// K1: fused skin + transform + light, 4 consecutive vertices per thread
// ------------------------------------------------------------------------------------------------

struct Vtx4 { float x[4], y[4], z[4]; };

// 4 packed float3 (48 B, 16B-aligned) -> SoA registers
__device__ __forceinline__ void load_vec3x4(const float* p, Vtx4& o) {
    float4 a = ldg4(p), b = ldg4(p + 4), c = ldg4(p + 8);
    o.x[0] = a.x; o.y[0] = a.y; o.z[0] = a.z;
    o.x[1] = a.w; o.y[1] = b.x; o.z[1] = b.y;
    o.x[2] = b.z; o.y[2] = b.w; o.z[2] = c.x;
    o.x[3] = c.y; o.y[3] = c.z; o.z[3] = c.w;
}

// LightVolume.VertexLight (light.go:583-663); rare, kept out of line.
__device__ __noinline__ void light_volume(const DrawRec& R, const LightU& U, const VolumeU* __restrict__ vols,
                                          Vec3 apos, Vec3 anrm, bool skinned, float& lr, float& lg, float& lb) {
    const VolumeU V = vols[U.volume];
    float cr = 0, cg = 0, cb = 0;
    const bool hasMat = R.flags & T3D_DRAW_HAS_MATERIAL;
    const int mode = hasMat ? R.lvMode : 0;
    Mat4 transform = m4_load(R.transform);
    Vec3 cellHalf = v3(V.cellSize[0] / 2, V.cellSize[1] / 2, V.cellSize[2] / 2);
    Vec3 bias = v3(V.bias[0], V.bias[1], V.bias[2]);
    Vec3 lookup;
    if (hasMat && mode == T3D_LV_PER_OBJECT) {
        lookup = v3(transform.m[3][0], transform.m[3][1], transform.m[3][2]);
    } else {
        Vec3 vertPos = v3(0, 0, 0), vertPosWithNormal = v3(0, 0, 0);
        if (skinned) {
            if (mode == T3D_LV_PER_VERTEX_WITH_NORMAL) vertPosWithNormal = v3_add(v3_add(vertPos, v3_mult(anrm, cellHalf)), bias);
            else vertPos = v3_add(apos, bias);
        } else {
            if (mode == T3D_LV_PER_VERTEX_WITH_NORMAL) vertPosWithNormal = v3_add(m4_multvec(transform, v3_add(apos, v3_mult(anrm, cellHalf))), bias);
            else vertPos = v3_add(m4_multvec(transform, apos), bias);
        }
        // Dimensions.Contains on vertPos (mesh.go:82-84) -- zero in WithNormal mode, as in the reference
        if (!(vertPos.X >= V.dimMin[0] && vertPos.Y >= V.dimMin[1] && vertPos.Z >= V.dimMin[2] &&
              vertPos.X <= V.dimMax[0] && vertPos.Y <= V.dimMax[1] && vertPos.Z <= V.dimMax[2])) return;
        lookup = (mode == T3D_LV_PER_VERTEX_WITH_NORMAL) ? vertPosWithNormal : vertPos;
    }
    // light.go:717-724, 1267-1272
    int x = (int)roundf(((lookup.X - V.offset[0] - V.dimMin[0] - (V.cellSize[0] / 2)) / V.cellSize[0]));
    int y = (int)roundf(((lookup.Y - V.offset[1] - V.dimMin[1] - (V.cellSize[1] / 2)) / V.cellSize[1]));
    int z = (int)roundf(((lookup.Z - V.offset[2] - V.dimMin[2] - (V.cellSize[2] / 2)) / V.cellSize[2]));
    if (!(x < 0 || y < 0 || z < 0 || x >= V.cellCount[0] || y >= V.cellCount[1] || z >= V.cellCount[2])) {
        const float* c = V.cells + 3 * ((size_t)x + ((size_t)y * V.cellCount[0]) + ((size_t)z * V.cellCount[0] * V.cellCount[1]));
        cr = c[0]; cg = c[1]; cb = c[2];
    }
    lr += U.color[0] * cr * U.energy;
    lg += U.color[1] * cg * U.energy;
    lb += U.color[2] * cb * U.energy;
}

// Model.skinVertex (model.go:559-596) for one vertex.
__device__ __forceinline__ void skin_one(const float* __restrict__ bones, float4 w, ushort4 j, int n,
                                         Vec3 p, Vec3 nIn, Vec3& pOut, Vec3& nOut) {
    float K[16];
#pragma unroll
    for (int i = 0; i < 16; i++) K[i] = 0.0f;
    const float ws[4] = {w.x, w.y, w.z, w.w};
    const uint16_t js[4] = {j.x, j.y, j.z, j.w};
#pragma unroll
    for (int b = 0; b < 4; b++) {
        if (b >= n) break;
        float weightPerc = ws[b];
        if (weightPerc == 0.0f) continue;
        const float4* B = reinterpret_cast<const float4*>(bones + 16 * (size_t)js[b]);
        float4 r0 = __ldg(B), r1 = __ldg(B + 1), r2 = __ldg(B + 2), r3 = __ldg(B + 3);
        const float I[16] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, r2.y, r2.z, r2.w, r3.x, r3.y, r3.z, r3.w};
        if (weightPerc == 1.0f) {
#pragma unroll
            for (int i = 0; i < 16; i++) K[i] = I[i];
            break;
        }
#pragma unroll
        for (int i = 0; i < 16; i++) K[i] = K[i] + I[i] * weightPerc;      // Scale then Add (matrix.go:639-661)
    }
    pOut.X = K[0] * p.X + K[4] * p.Y + K[8] * p.Z + K[12];                 // MultVec, matrix.go:619
    pOut.Y = K[1] * p.X + K[5] * p.Y + K[9] * p.Z + K[13];
    pOut.Z = K[2] * p.X + K[6] * p.Y + K[10] * p.Z + K[14];
    // skinMatrix[3] = {0,0,0,1} (model.go:587-590)
    nOut.X = K[0] * nIn.X + K[4] * nIn.Y + K[8] * nIn.Z + 0.0f;
    nOut.Y = K[1] * nIn.X + K[5] * nIn.Y + K[9] * nIn.Z + 0.0f;
    nOut.Z = K[2] * nIn.X + K[6] * nIn.Y + K[10] * nIn.Z + 0.0f;
}

__global__ void __launch_bounds__(256)
k_vertex(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ vbase, const uint32_t* __restrict__ tileFirst,
         uint32_t totalGroups, const __grid_constant__ FrameConst fc,
         const LightU* __restrict__ lightsU, const float* __restrict__ bones, const VolumeU* __restrict__ vols,
         float4* __restrict__ clipOut, float4* __restrict__ lightOut, float4* __restrict__ wposOut,
         float* __restrict__ unbillOut, float4* __restrict__ tnrmOut) {
    const uint32_t g = blockIdx.x * 256u + threadIdx.x;
    if (g >= totalGroups) return;
    const uint32_t slot = g * 4u;
    const uint32_t d = find_draw(vbase, __ldg(&tileFirst[blockIdx.x]), __ldg(&tileFirst[blockIdx.x + 1]), slot);
    const DrawRec& R = draws[d];
    const uint32_t flags = R.flags;
    const bool skinned = flags & T3D_DRAW_SKINNED;
    const bool lit = (flags & T3D_DRAW_LIT) != 0;
    const uint32_t mv = R.a0 + (slot - R.vbase);      // first of 4 mesh vertices

    Vtx4 P, N;
    load_vec3x4(R.pos + 3 * (size_t)mv, P);
    const bool needN = lit || skinned;
    if (needN) load_vec3x4(R.nrm + 3 * (size_t)mv, N);
    else {
#pragma unroll
        for (int k = 0; k < 4; k++) { N.x[k] = 0; N.y[k] = 0; N.z[k] = 0; }
    }

    // shape keys (mesh.go:796-822)
    for (uint32_t k = 0; k < R.nKeys; k++) {
        const float w = __ldg(&R.keyW[k]);
        Vtx4 KP; load_vec3x4(R.keyPos + ((size_t)k * R.keyStride + mv) * 3, KP);
#pragma unroll
        for (int i = 0; i < 4; i++) { P.x[i] += KP.x[i] * w; P.y[i] += KP.y[i] * w; P.z[i] += KP.z[i] * w; }
        if (needN) {
            Vtx4 KN; load_vec3x4(R.keyNrm + ((size_t)k * R.keyStride + mv) * 3, KN);
#pragma unroll
            for (int i = 0; i < 4; i++) { N.x[i] += KN.x[i] * w; N.y[i] += KN.y[i] * w; N.z[i] += KN.z[i] * w; }
        }
    }

    // skinning (model.go:760-775): positions and normals become world-space
    if (skinned) {
        const uint32_t ninf4 = R.ninf ? __ldg(reinterpret_cast<const uint32_t*>(R.ninf + mv)) : 0x04040404u;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 w = ldg4(R.wgt + 4 * (size_t)(mv + i));
            ushort4 j = __ldg(reinterpret_cast<const ushort4*>(R.jnt + 4 * (size_t)(mv + i)));
            int n = (ninf4 >> (8 * i)) & 0xFF;
            Vec3 po, no;
            skin_one(bones + 16 * (size_t)R.boneOffset, w, j, n, v3(P.x[i], P.y[i], P.z[i]), v3(N.x[i], N.y[i], N.z[i]), po, no);
            P.x[i] = po.X; P.y[i] = po.Y; P.z[i] = po.Z;
            N.x[i] = no.X; N.y[i] = no.Y; N.z[i] = no.Z;
            wposOut[slot + i] = make_float4(po.X, po.Y, po.Z, 0.0f);
        }
    }

    // clip = [x y z 1] * M (model.go:772-775 / 795-798)
    {
        const float* Mp = skinned ? fc.vp : R.mvp;
        float4 m0, m1, m2, m3;
        if (skinned) {
            m0 = make_float4(fc.vp[0], fc.vp[1], fc.vp[2], fc.vp[3]);
            m1 = make_float4(fc.vp[4], fc.vp[5], fc.vp[6], fc.vp[7]);
            m2 = make_float4(fc.vp[8], fc.vp[9], fc.vp[10], fc.vp[11]);
            m3 = make_float4(fc.vp[12], fc.vp[13], fc.vp[14], fc.vp[15]);
        } else {
            m0 = ldg4(Mp); m1 = ldg4(Mp + 4); m2 = ldg4(Mp + 8); m3 = ldg4(Mp + 12);
        }
        const float snap = fc.snapping;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 c;
            c.x = m0.x * P.x[i] + m1.x * P.y[i] + m2.x * P.z[i] + m3.x;
            c.y = m0.y * P.x[i] + m1.y * P.y[i] + m2.y * P.z[i] + m3.y;
            c.z = m0.z * P.x[i] + m1.z * P.y[i] + m2.z * P.z[i] + m3.z;
            c.w = m0.w * P.x[i] + m1.w * P.y[i] + m2.w * P.z[i] + m3.w;
            if (snap > 0) {                                             // model.go:807-812
                c.x = roundf(c.x * snap) / snap;
                c.y = roundf(c.y * snap) / snap;
                c.z = roundf(c.z * snap) / snap;
            }
            clipOut[slot + i] = c;
        }
        if (!skinned && (flags & T3D_DRAW_DEPTH_UNBILLBOARDED)) {       // model.go:800-803
            float4 u = ldg4(R.unbill);
            float4 o;
            o.x = u.x * P.x[0] + u.y * P.y[0] + u.z * P.z[0] + u.w;
            o.y = u.x * P.x[1] + u.y * P.y[1] + u.z * P.z[1] + u.w;
            o.z = u.x * P.x[2] + u.y * P.y[2] + u.z * P.z[2] + u.w;
            o.w = u.x * P.x[3] + u.y * P.y[3] + u.z * P.z[3] + u.w;
            *reinterpret_cast<float4*>(unbillOut + slot) = o;
        }
        if (fc.camFlags & T3D_CAM_RENDER_NORMALS) {                     // model.go:814-818 (uses vert)
            float4 c0 = ldg4(R.nrmRot), c1 = ldg4(R.nrmRot + 4), c2 = ldg4(R.nrmRot + 8);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float4 t;
                t.x = c0.x * P.x[i] + c0.y * P.y[i] + c0.z * P.z[i] + c0.w;
                t.y = c1.x * P.x[i] + c1.y * P.y[i] + c1.z * P.z[i] + c1.w;
                t.z = c2.x * P.x[i] + c2.y * P.y[i] + c2.z * P.z[i] + c2.w;
                t.w = 0.0f;
                tnrmOut[slot + i] = t;
            }
        }
    }

    if (!lit) return;

    // per-vertex lighting (camera.go:1610-1644): start at (0,0,0,1), add every light of the draw's list
    float lr[4] = {0, 0, 0, 0}, lg[4] = {0, 0, 0, 0}, lb[4] = {0, 0, 0, 0};
    const uint8_t lightingMode = (flags & T3D_DRAW_HAS_MATERIAL) ? R.lightingMode : (uint8_t)T3D_LIGHTING_DEFAULT;
    const uint32_t nL = R.lightCount;
    const float4* LUp = reinterpret_cast<const float4*>(lightsU + R.lightU);
    for (uint32_t li = 0; li < nL; li++) {
        const float4 q0 = __ldg(LUp + 4 * li), q1 = __ldg(LUp + 4 * li + 1), q2 = __ldg(LUp + 4 * li + 2);
        const uint32_t type = __float_as_uint(q0.x);
        const float cR = q0.y, cG = q0.z, cB = q0.w, energy = q1.x;
        if (type == T3D_LIGHT_AMBIENT) {                                // light.go:110-120
            const float ar = cR * energy, ag = cG * energy, ab = cB * energy;
#pragma unroll
            for (int i = 0; i < 4; i++) { lr[i] += ar; lg[i] += ag; lb[i] += ab; }
        } else if (type == T3D_LIGHT_POINT) {                           // light.go:191-258
            const Vec3 wp = v3(q1.y, q1.z, q1.w);
            const float range = q2.x, range2 = q2.y;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                Vec3 vertPos = v3(P.x[i], P.y[i], P.z[i]);
                Vec3 vertNormal = v3(N.x[i], N.y[i], N.z[i]);
                float distance = v3_dist2(wp, vertPos);
                if (range > 0) { if (distance > range2) continue; }
                else { if (1.0f / distance * energy < 0.001f) continue; }
                Vec3 lightVec = v3_unit(v3_sub(wp, vertPos));
                if (lightingMode == T3D_LIGHTING_FIXED_NORMALS) vertNormal = lightVec;
                float diffuse = v3_dot(vertNormal, lightVec);
                if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuse = fabsf(diffuse);
                if (diffuse > 0) {
                    float diffuseFactor = diffuse * (1.0f / (1.0f + (0.1f * distance))) * 2.0f;
                    if (range2 > 0) diffuseFactor *= clampf((range2 - distance) / distance, 0.0f, 1.0f);
                    lr[i] += cR * diffuseFactor * energy;
                    lg[i] += cG * diffuseFactor * energy;
                    lb[i] += cB * diffuseFactor * energy;
                }
            }
        } else if (type == T3D_LIGHT_DIRECTIONAL) {                     // light.go:332-360
            const float4 q3 = __ldg(LUp + 4 * li + 3);
            const Vec3 fwd = v3(q2.z, q2.w, q3.x);
            float4 s0, s1, s2, s3;
            if (!skinned) { s0 = ldg4(R.worldRot); s1 = ldg4(R.worldRot + 4); s2 = ldg4(R.worldRot + 8); s3 = ldg4(R.worldRot + 12); }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                Vec3 normal;
                if (lightingMode == T3D_LIGHTING_FIXED_NORMALS) normal = fwd;
                else if (skinned) normal = v3(N.x[i], N.y[i], N.z[i]);
                else {
                    normal.X = s0.x * N.x[i] + s1.x * N.y[i] + s2.x * N.z[i] + s3.x;
                    normal.Y = s0.y * N.x[i] + s1.y * N.y[i] + s2.y * N.z[i] + s3.y;
                    normal.Z = s0.z * N.x[i] + s1.z * N.y[i] + s2.z * N.z[i] + s3.z;
                }
                float diffuseFactor = v3_dot(normal, fwd);
                if (lightingMode == T3D_LIGHTING_DOUBLE_SIDED) diffuseFactor = fabsf(diffuseFactor);
                if (diffuseFactor <= 0) continue;
                lr[i] += cR * diffuseFactor * energy;
                lg[i] += cG * diffuseFactor * energy;
                lb[i] += cB * diffuseFactor * energy;
            }
        } else if (type == T3D_LIGHT_VOLUME) {
            const LightU U = lightsU[R.lightU + li];
#pragma unroll 1
            for (int i = 0; i < 4; i++)
                light_volume(R, U, vols, v3(P.x[i], P.y[i], P.z[i]), v3(N.x[i], N.y[i], N.z[i]), skinned, lr[i], lg[i], lb[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) lightOut[slot + i] = make_float4(lr[i], lg[i], lb[i], 1.0f);
}

// ------------------------------------------------------------------------------------------------
// K2: cull + depth key + per-draw min/max (model.go:855-969)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cull(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ tbase, const uint32_t* __restrict__ tileFirst,
       uint32_t totalTris, const __grid_constant__ FrameConst fc,
       const float4* __restrict__ clip, const float4* __restrict__ wpos,
       uint32_t* __restrict__ tdepth, uint32_t* __restrict__ tdraw, DrawStat* __restrict__ stats) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    const bool active = t < totalTris;
    uint32_t d = 0;
    bool keep = false;
    float depth = 0.0f;
    if (active) {
        d = find_draw(tbase, __ldg(&tileFirst[blockIdx.x]), __ldg(&tileFirst[blockIdx.x + 1]), t);
        const DrawRec& R = draws[d];
        const uint32_t mt = R.t0 + (t - R.tbase);
        const uint32_t i0 = __ldg(&R.idx[3 * (size_t)mt]), i1 = __ldg(&R.idx[3 * (size_t)mt + 1]), i2 = __ldg(&R.idx[3 * (size_t)mt + 2]);
        const uint32_t vb = R.vbase - R.a0;
        const float4 c0 = __ldg(&clip[vb + i0]), c1 = __ldg(&clip[vb + i1]), c2 = __ldg(&clip[vb + i2]);

        float w0 = c0.w, w1 = c1.w, w2 = c2.w;                          // model.go:866-875
        if (w0 < 0) w0 = 0.000001f;
        if (w1 < 0) w1 = 0.000001f;
        if (w2 < 0) w2 = 0.000001f;
        const float x0 = c0.x / w0, y0 = c0.y / w0;
        const float x1 = c1.x / w1, y1 = c1.y / w1;
        const float x2 = c2.x / w2, y2 = c2.y / w2;
        const float nearP = fc.nearP, farP = fc.farP;
        const bool inRange = (c0.z + 1.0f >= nearP && c0.z < farP) || (c1.z + 1.0f >= nearP && c1.z < farP) ||
                             (c2.z + 1.0f >= nearP && c2.z < farP);    // model.go:877-879
        keep = inRange;
        if (keep && ((x0 < -0.5f && x1 < -0.5f && x2 < -0.5f) || (x0 > 0.5f && x1 > 0.5f && x2 > 0.5f) ||
                     (y0 < -0.5f && y1 < -0.5f && y2 < -0.5f) || (y0 > 0.5f && y1 > 0.5f && y2 > 0.5f)))
            keep = false;                                               // model.go:891-894
        if (keep && (R.flags & T3D_DRAW_HAS_MATERIAL) && (R.flags & T3D_DRAW_BACKFACE_CULLING)) {   // model.go:902-925
            const float n0x = x0 - x1, n0y = y0 - y1, n1x = x1 - x2, n1y = y1 - y2;
            const float nor = (n0x * n1y) - (n1x * n0y);
            if (nor < 0) keep = false;
        }
        if (keep) {                                                     // model.go:933-955
            if (R.flags & T3D_DRAW_SKINNED) {
                const float4 p0 = __ldg(&wpos[vb + i0]), p1 = __ldg(&wpos[vb + i1]), p2 = __ldg(&wpos[vb + i2]);
                float sx = 0.0f, sy = 0.0f, sz = 0.0f;
                sx += p0.x; sy += p0.y; sz += p0.z;
                sx += p1.x; sy += p1.y; sz += p1.z;
                sx += p2.x; sy += p2.y; sz += p2.z;
                const float dx = fc.camPos[0] - (sx / 3.0f), dy = fc.camPos[1] - (sy / 3.0f), dz = fc.camPos[2] - (sz / 3.0f);
                depth = dx * dx + dy * dy + dz * dz;
            } else {
                const float* C = R.ctr + 3 * (size_t)mt;
                const float dx = R.invCam[0] - __ldg(C), dy = R.invCam[1] - __ldg(C + 1), dz = R.invCam[2] - __ldg(C + 2);
                depth = dx * dx + dy * dy + dz * dz;
            }
            if (isnan(depth) || isinf(depth)) keep = false;            // model.go:957-960
        }
        tdepth[t] = keep ? __float_as_uint(depth) : KEY_CULLED;
        tdraw[t] = d;
    }

    // per-draw min / max / survivor count, aggregated per warp when the warp sits inside one draw
    const unsigned full = 0xFFFFFFFFu;
    const uint32_t d0 = __shfl_sync(full, d, 0);
    const bool uniform = __all_sync(full, (!active) || d == d0);
    const uint32_t bits = __float_as_uint(depth);
    if (uniform) {
        const unsigned km = __ballot_sync(full, keep);
        if (km) {
            uint32_t mn = keep ? bits : 0xFFFFFFFFu, mx = keep ? bits : 0u;
            mn = __reduce_min_sync(full, mn);
            mx = __reduce_max_sync(full, mx);
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&stats[d0].minBits, mn);
                atomicMax(&stats[d0].maxBits, mx);
                atomicAdd(&stats[d0].survivors, (uint32_t)__popc(km));
            }
        }
    } else if (keep) {
        atomicMin(&stats[d].minBits, bits);
        atomicMax(&stats[d].maxBits, bits);
        atomicAdd(&stats[d].survivors, 1u);
    }
}

// ------------------------------------------------------------------------------------------------
// K3: 512-bin key (sortingTriangles.go:38-64); BackToFront visits bins 511..0 (sortingTriangles.go:101)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_keys(const DrawRec* __restrict__ draws, const DrawStat* __restrict__ stats, const uint32_t* __restrict__ tdepth,
       const uint32_t* __restrict__ tdraw, uint32_t totalTris, uint32_t* __restrict__ keys) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    if (t >= totalTris) return;
    const uint32_t bits = tdepth[t];
    uint32_t key = KEY_CULLED;
    if (bits != KEY_CULLED) {
        const uint32_t d = tdraw[t];
        const uint32_t sortMode = draws[d].sortMode;
        const float minRange = __uint_as_float(stats[d].minBits);
        const float maxRange = __uint_as_float(stats[d].maxBits);
        float rangeDiff = maxRange - minRange;
        if (rangeDiff == 0) rangeDiff += 0.001f;
        uint32_t bin = 0;
        if (sortMode != T3D_SORT_NONE) {
            float dd = (__uint_as_float(bits) - minRange) / rangeDiff * (float)BIN_COUNT;
            float tt = clampf(dd, 0.0f, (float)(BIN_COUNT - 1));
            bin = (uint32_t)(int)tt;
        }
        const uint32_t binKey = (sortMode == T3D_SORT_BACK_TO_FRONT) ? (BIN_COUNT - 1 - bin) : bin;
        key = (d << BIN_BITS) | binKey;
    }
    keys[t] = key;
}

// ------------------------------------------------------------------------------------------------
// K4: emit (camera.go:1646-1820), one thread per output vertex
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_emit(const DrawRec* __restrict__ draws, const uint32_t* __restrict__ sortedKeys, const uint32_t* __restrict__ sortedTris,
       const uint32_t* __restrict__ nSurvivors, const __grid_constant__ FrameConst fc,
       const float4* __restrict__ clip, const float4* __restrict__ light, const float* __restrict__ unbill,
       const float4* __restrict__ tnrm,
       float4* __restrict__ colorList, float4* __restrict__ depthList, float4* __restrict__ normalList,
       uint32_t* __restrict__ triIds) {
    const uint32_t o = blockIdx.x * 256u + threadIdx.x;
    const uint32_t n = *nSurvivors;
    if (o >= 3u * n) return;
    const uint32_t k = o / 3u, c = o - 3u * k;
    const uint32_t key = __ldg(&sortedKeys[k]);
    const uint32_t t = __ldg(&sortedTris[k]);
    const uint32_t d = key >> BIN_BITS;
    const DrawRec& R = draws[d];
    const uint32_t mt = R.t0 + (t - R.tbase);
    const uint32_t vi = __ldg(&R.idx[3 * (size_t)mt + c]);
    const uint32_t slot = R.vbase + (vi - R.a0);
    const float4 T = __ldg(&clip[slot]);
    if (c == 0) triIds[k] = mt;

    const bool persp = fc.camFlags & T3D_CAM_PERSPECTIVE;
    const bool pcuv = fc.camFlags & T3D_CAM_PERSP_CORRECT_UV;
    float w = T.w;                                                      // camera.go:1655-1666
    if (!persp) w = 1.0f;
    if (w < 0) w = 0.001f;
    const float dx = (T.x / w) * fc.camW + fc.halfW;                    // camera.go:1674-1675
    const float dy = (T.y / -w) * fc.camH + fc.halfH;

    const float2 uv = __ldg(reinterpret_cast<const float2*>(R.uv) + vi);
    float uvU, uvV, custom0 = 0.0f;
    if (pcuv) {                                                         // camera.go:1689-1705
        uvU = (uv.x / w) * R.srcW;
        uvV = ((1.0f - uv.y) / w) * R.srcH;
        custom0 = 1.0f / w;
    } else {
        uvU = uv.x * R.srcW;
        uvV = (1.0f - uv.y) * R.srcH;
    }

    float4 col = ldg4(R.color);                                         // camera.go:1726-1736
    if (R.col) {
        const float4 vc = ldg4(R.col + 4 * (size_t)vi);
        col.x = vc.x * col.x; col.y = vc.y * col.y; col.z = vc.z * col.z; col.w = vc.w * col.w;
    }
    if (R.flags & T3D_DRAW_LIT) {                                       // camera.go:1738-1742
        const float4 L = __ldg(&light[slot]);
        col.x *= L.x; col.y *= L.y; col.z *= L.z;
    }

    float4 dcol = make_float4(0, 0, 0, 0);
    if (fc.camFlags & T3D_CAM_RENDER_DEPTH) {                           // camera.go:1744-1782
        float vertexDepth = T.z;
        if ((R.flags & T3D_DRAW_HAS_MATERIAL) && (R.flags & T3D_DRAW_DEPTH_UNBILLBOARDED)) vertexDepth = __ldg(&unbill[slot]);
        float depth = (vertexDepth + fc.depthMargin) / fc.camSpread;
        if (depth < 0) depth = 0; else if (depth > 1) depth = 1;
        dcol = make_float4(depth, depth, depth, 1.0f);
    } else if (fc.worldPresent && fc.fogOn) {                           // camera.go:1784-1814
        float depth = (T.z + fc.depthMargin) / fc.camSpread;
        if (depth < 0) depth = 0; else if (depth > 1) depth = 1;
        depth = fc.fogRange[0] + ((fc.fogRange[1] - fc.fogRange[0]) * 1.0f - depth);
        if (fc.fogMode == T3D_FOG_ADD) {
            col.x += fc.fogColor[0] * depth; col.y += fc.fogColor[1] * depth; col.z += fc.fogColor[2] * depth;
        } else if (fc.fogMode == T3D_FOG_SUB) {
            col.x *= fc.fogColor[0] * depth; col.y *= fc.fogColor[1] * depth; col.z *= fc.fogColor[2] * depth;
        }
    }

    const float4 a = make_float4(dx, dy, uvU, uvV);
    const float4 cust = make_float4(custom0, 0.0f, 0.0f, 0.0f);
    float4* C = colorList + 3 * (size_t)o;
    C[0] = a; C[1] = col; C[2] = cust;
    float4* D = depthList + 3 * (size_t)o;
    D[0] = a; D[1] = dcol; D[2] = cust;
    if (fc.camFlags & T3D_CAM_RENDER_NORMALS) {                         // camera.go:1710-1722
        const float4 tn = __ldg(&tnrm[slot]);
        float4* Nn = normalList + 3 * (size_t)o;
        Nn[0] = a;
        Nn[1] = make_float4(tn.x * 0.5f + 0.5f, tn.y * 0.5f + 0.5f, tn.z * 0.5f + 0.5f, 0.0f);
        Nn[2] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
}

// ------------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------------
static inline uint32_t cdiv(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

void launch_prep_draws(cudaStream_t s, DrawRec* draws, uint32_t nDraws, const FrameConst& fc,
                       const t3d_light* lightsIn, LightU* lightsU, DrawStat* stats) {
    if (!nDraws) return;
    k_prep_draws<<<cdiv(nDraws, 128), 128, 0, s>>>(draws, nDraws, fc, lightsIn, lightsU, stats);
}

void launch_vertex(cudaStream_t s, const DrawRec* draws, const uint32_t* vbase, const uint32_t* tileFirst,
                   uint32_t totalGroups, const FrameConst& fc, const LightU* lightsU, const float* bones,
                   const VolumeU* vols, float4* clip, float4* light, float4* wpos, float* unbill, float4* tnrm) {
    if (!totalGroups) return;
    k_vertex<<<cdiv(totalGroups, 256), 256, 0, s>>>(draws, vbase, tileFirst, totalGroups, fc, lightsU, bones, vols,
                                                    clip, light, wpos, unbill, tnrm);
}

void launch_cull(cudaStream_t s, const DrawRec* draws, const uint32_t* tbase, const uint32_t* tileFirst,
                 uint32_t totalTris, const FrameConst& fc, const float4* clip, const float4* wpos,
                 uint32_t* tdepth, uint32_t* tdraw, DrawStat* stats) {
    if (!totalTris) return;
    k_cull<<<cdiv(totalTris, 256), 256, 0, s>>>(draws, tbase, tileFirst, totalTris, fc, clip, wpos, tdepth, tdraw, stats);
}

void launch_keys(cudaStream_t s, const DrawRec* draws, const DrawStat* stats, const uint32_t* tdepth,
                 const uint32_t* tdraw, uint32_t totalTris, uint32_t* keys) {
    if (!totalTris) return;
    k_keys<<<cdiv(totalTris, 256), 256, 0, s>>>(draws, stats, tdepth, tdraw, totalTris, keys);
}

void launch_emit(cudaStream_t s, const DrawRec* draws, const uint32_t* sortedKeys, const uint32_t* sortedTris,
                 const uint32_t* nSurvivors, uint32_t maxTris, const FrameConst& fc, const float4* clip,
                 const float4* light, const float* unbill, const float4* tnrm, float4* colorList,
                 float4* depthList, float4* normalList, uint32_t* triIds) {
    if (!maxTris) return;
    k_emit<<<cdiv(3ull * maxTris, 256), 256, 0, s>>>(draws, sortedKeys, sortedTris, nSurvivors, fc, clip, light, unbill,
                                                     tnrm, colorList, depthList, normalList, triIds);
}

}  // namespace t3d
