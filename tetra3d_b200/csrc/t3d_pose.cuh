// t3d_pose.cuh -- device code of the bone-pose evaluation (see t3d_pose.cu for what it replaces); included by
// t3d_pose.cu (stand-alone k_pose_bones) and t3d_kernels.cu (pose blocks riding at the end of k_prep_draws' grid).
#pragma once

#include "t3d_device.cuh"

namespace t3d {

static const int POSE_THREADS = 256;
__host__ __device__ inline uint32_t pose_armatures_per_block(uint32_t nBones) { return nBones >= (uint32_t)POSE_THREADS ? 1u : (uint32_t)POSE_THREADS / nBones; }
// shared memory of a pose block: local + world matrices (16 planes each) and the block's pose records
__host__ __device__ inline size_t pose_smem_bytes(uint32_t nBones) {
    const uint32_t ipb = pose_armatures_per_block(nBones);
    return ((size_t)32 * ipb * nBones + (size_t)20 * ipb) * sizeof(float);
}

// ---- Go's math.Sin / math.Atan2 (src/math/sin.go, atan.go, atan2.go) -------------------------------------------
// Constants as IEEE-754 bit patterns (the hexadecimal values Go's source lists next to each literal).
__device__ __forceinline__ double f64(unsigned long long bits) { return __longlong_as_double((long long)bits); }

__device__ inline double go_sin(double x) {
    const double PI4A = f64(0x3fe921fb40000000ull), PI4B = f64(0x3e64442d00000000ull), PI4C = f64(0x3ce8469898cc5170ull);
    const double S0 = f64(0x3de5d8fd1fd19ccdull), S1 = f64(0xbe5ae5e5a9291f5dull), S2 = f64(0x3ec71de3567d48a1ull),
                 S3 = f64(0xbf2a01a019bfdf03ull), S4 = f64(0x3f8111111110f7d0ull), S5 = f64(0xbfc5555555555548ull);
    const double C0 = f64(0xbda8fa49a0861a9bull), C1 = f64(0x3e21ee9d7b4e3f05ull), C2 = f64(0xbe927e4f7eac4bc6ull),
                 C3 = f64(0x3efa01a019c844f5ull), C4 = f64(0xbf56c16c16c14f91ull), C5 = f64(0x3fa555555555554bull);
    const double FOUR_OVER_PI = f64(0x3ff45f306dc9c883ull);
    if (x == 0.0 || isnan(x)) return x;
    if (isinf(x)) return f64(0x7ff8000000000001ull);
    bool sign = false;
    if (x < 0.0) { x = -x; sign = true; }
    // Slerp only passes angles in [0, pi]: the Payne-Hanek branch (x >= 2^29) is not restated
    unsigned long long j = (unsigned long long)(x * FOUR_OVER_PI);
    double y = (double)j;
    if (j & 1ull) { j++; y += 1.0; }
    j &= 7ull;
    const double z = ((x - y * PI4A) - y * PI4B) - y * PI4C;
    if (j > 3ull) { sign = !sign; j -= 4ull; }
    const double zz = z * z;
    if (j == 1ull || j == 2ull) y = 1.0 - 0.5 * zz + zz * zz * ((((((C0 * zz) + C1) * zz + C2) * zz + C3) * zz + C4) * zz + C5);
    else y = z + z * zz * ((((((S0 * zz) + S1) * zz + S2) * zz + S3) * zz + S4) * zz + S5);
    return sign ? -y : y;
}

__device__ inline double go_xatan(double x) {
    const double P0 = -8.750608600031904122785e-01, P1 = -1.615753718733365076637e+01, P2 = -7.500855792314704667340e+01,
                 P3 = -1.228866684490136173410e+02, P4 = -6.485021904942025371773e+01;
    const double Q0 = +2.485846490142306297962e+01, Q1 = +1.650270098316988542046e+02, Q2 = +4.328810604912902668951e+02,
                 Q3 = +4.853903996359136964868e+02, Q4 = +1.945506571482613964425e+02;
    double z = x * x;
    z = z * ((((P0 * z + P1) * z + P2) * z + P3) * z + P4) / (((((z + Q0) * z + Q1) * z + Q2) * z + Q3) * z + Q4);
    z = x * z + x;
    return z;
}
__device__ inline double go_satan(double x) {
    const double Morebits = 6.123233995736765886130e-17, Tan3pio8 = 2.41421356237309504880;
    const double PI_2 = f64(0x3ff921fb54442d18ull), PI_4 = f64(0x3fe921fb54442d18ull);
    if (x <= 0.66) return go_xatan(x);
    if (x > Tan3pio8) return PI_2 - go_xatan(1.0 / x) + Morebits;
    return PI_4 + go_xatan((x - 1.0) / (x + 1.0)) + 0.5 * Morebits;
}
__device__ inline double go_atan(double x) {
    if (x == 0.0) return x;
    if (x > 0.0) return go_satan(x);
    return -go_satan(-x);
}
__device__ inline double go_atan2(double y, double x) {
    const double PI = f64(0x400921fb54442d18ull), PI_2 = f64(0x3ff921fb54442d18ull), PI_4 = f64(0x3fe921fb54442d18ull),
                 PI3_4 = f64(0x4002d97c7f3321d2ull);
    if (isnan(y) || isnan(x)) return f64(0x7ff8000000000001ull);
    if (y == 0.0) {
        if (x >= 0.0 && !signbit(x)) return copysign(0.0, y);
        return copysign(PI, y);
    }
    if (x == 0.0) return copysign(PI_2, y);
    if (isinf(x)) {
        if (x > 0.0) return isinf(y) ? copysign(PI_4, y) : copysign(0.0, y);
        return isinf(y) ? copysign(PI3_4, y) : copysign(PI, y);
    }
    if (isinf(y)) return copysign(PI_2, y);
    const double q = go_atan(y / x);
    if (x < 0.0) return q <= 0.0 ? q + PI : q - PI;
    return q;
}

struct Quat { float X, Y, Z, W; };

__device__ inline Quat quat_slerp(Quat a, Quat end, float percent) {          // quaternion.go:26-59
    float cosOmega = a.X * end.X + a.Y * end.Y + a.Z * end.Z + a.W * end.W;
    if (cosOmega < 0.0f) { end.W *= -1.0f; end.X *= -1.0f; end.Y *= -1.0f; end.Z *= -1.0f; cosOmega *= -1.0f; }
    float k0, k1;
    if (cosOmega > 0.9999f) { k0 = 1.0f - percent; k1 = percent; }
    else {
        const float sinOmega = sqrtf(1.0f - (cosOmega * cosOmega));       // math32.Sqrt: float32(math.Sqrt(float64)) == correctly rounded sqrtf
        const float omega = (float)go_atan2((double)sinOmega, (double)cosOmega);
        const float oneOverSinOmega = 1.0f / sinOmega;
        k0 = (float)go_sin((double)((1.0f - percent) * omega)) * oneOverSinOmega;
        k1 = (float)go_sin((double)(percent * omega)) * oneOverSinOmega;
    }
    Quat r;
    r.W = a.W * k0 + end.W * k1;
    r.X = a.X * k0 + end.X * k1;
    r.Y = a.Y * k0 + end.Y * k1;
    r.Z = a.Z * k0 + end.Z * k1;
    return r;
}

__device__ inline Mat4 quat_to_matrix(Quat q) {                                // quaternion.go:132-178
    Mat4 m1, m2;
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

// The keyframes around `time` (animation.go:63-90, 116-141).  Returns 0: no keyframes, 1: the value of key a,
// 2: interpolate keys a -> b.  Key times are non-decreasing inside a track (checked at upload), so the reference's
// linear scan "last key before, first key not before" is a lower bound.
__device__ inline int track_keys(const t3d_track tr, const float* __restrict__ times, float time, uint32_t& a, uint32_t& b) {
    if (tr.n_keys == 0) return 0;
    const float* T = times + tr.first_key;
    if (time <= T[0]) { a = tr.first_key; return 1; }
    if (time >= T[tr.n_keys - 1]) { a = tr.first_key + tr.n_keys - 1; return 1; }
    uint32_t lo = 0, hi = tr.n_keys - 1;                // T[lo] < time <= T[hi] holds throughout (NaN playheads end up at key 0)
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (T[mid] < time) lo = mid; else hi = mid; }
    if (!(T[lo] < time)) { a = tr.first_key; return 1; }
    if (time == T[hi]) { a = tr.first_key + hi; return 1; }
    a = tr.first_key + lo; b = tr.first_key + hi;
    return tr.interpolation == T3D_INTERPOLATION_CONSTANT ? 1 : 2;
}

__device__ inline Quat quat_unit(Quat q) {                                     // quaternion.go:86-104
    const float m = sqrtf((q.X * q.X) + (q.Y * q.Y) + (q.Z * q.Z) + (q.W * q.W));
    q.X /= m; q.Y /= m; q.Z /= m; q.W /= m;
    return q;
}

// What AnimationPlayer.updateValues records for one bone (AnimationValues, animation.go:674-722): the three tracks of the
// bone's channel sampled at `time`, and which of them exist.
struct BoneSample { float px, py, pz, sx, sy, sz; Quat q; bool P, S, R; };
__device__ inline BoneSample sample_bone(const t3d_track* __restrict__ tracks, const float* __restrict__ keyTimes,
                                         const float4* __restrict__ keyValues, uint32_t bone, float time) {
    BoneSample o;
    o.px = o.py = o.pz = o.sx = o.sy = o.sz = 0.0f; o.q.X = o.q.Y = o.q.Z = o.q.W = 0.0f;
    uint32_t a = 0, b = 0;
    int k = track_keys(tracks[3 * bone + 0], keyTimes, time, a, b);
    o.P = k != 0;
    if (k) {                                                    // ValueAsVector + Vector3.Lerp (vector.go:559-565)
        const float4 fa = keyValues[a];
        o.px = fa.x; o.py = fa.y; o.pz = fa.z;
        if (k == 2) {
            const float4 fb = keyValues[b];
            const float t = (time - keyTimes[a]) / (keyTimes[b] - keyTimes[a]);
            o.px = o.px + ((fb.x - o.px) * t); o.py = o.py + ((fb.y - o.py) * t); o.pz = o.pz + ((fb.z - o.pz) * t);
        }
    }
    k = track_keys(tracks[3 * bone + 1], keyTimes, time, a, b);
    o.S = k != 0;
    if (k) {
        const float4 fa = keyValues[a];
        o.sx = fa.x; o.sy = fa.y; o.sz = fa.z;
        if (k == 2) {
            const float4 fb = keyValues[b];
            const float t = (time - keyTimes[a]) / (keyTimes[b] - keyTimes[a]);
            o.sx = o.sx + ((fb.x - o.sx) * t); o.sy = o.sy + ((fb.y - o.sy) * t); o.sz = o.sz + ((fb.z - o.sz) * t);
        }
    }
    k = track_keys(tracks[3 * bone + 2], keyTimes, time, a, b);
    o.R = k != 0;
    if (k) {                                                    // ValueAsQuaternion
        const float4 fa = keyValues[a];
        o.q.X = fa.x; o.q.Y = fa.y; o.q.Z = fa.z; o.q.W = fa.w;
        if (k == 2) {
            const float4 fb = keyValues[b];
            Quat e; e.X = fb.x; e.Y = fb.y; e.Z = fb.z; e.W = fb.w;
            const float t = (time - keyTimes[a]) / (keyTimes[b] - keyTimes[a]);
            o.q = quat_slerp(o.q, e, t);
        }
    }
    return o;
}

// forceUpdate's blend of one node (animation.go:859-897): f = prevAnimatedProperties[node], v = this frame's samples.
__device__ inline BoneSample blend_sample(const BoneSample f, BoneSample v, const float bp) {
    if (f.P) {                                              // animation.go:861-872
        if (v.P) { v.px = f.px + ((v.px - f.px) * bp); v.py = f.py + ((v.py - f.py) * bp); v.pz = f.pz + ((v.pz - f.pz) * bp); }
        else { v.px = f.px; v.py = f.py; v.pz = f.pz; v.P = true; }
    }
    if (f.S) {                                              // animation.go:874-884
        if (v.S) { v.sx = f.sx + ((v.sx - f.sx) * bp); v.sy = f.sy + ((v.sy - f.sy) * bp); v.sz = f.sz + ((v.sz - f.sz) * bp); }
        else { v.sx = f.sx; v.sy = f.sy; v.sz = f.sz; v.S = true; }
    }
    if (f.R) {                                              // animation.go:886-896
        if (v.R) v.q = quat_unit(quat_slerp(f.q, v.q, bp));
        else { v.q = f.q; v.R = true; }
    }
    return v;
}
// SetLocalRotation(q.ToMatrix4()) (animation.go:976, node.go:906-912): an all-zero matrix is ignored
__device__ inline void apply_rotation(const Quat q, Mat4& rot) {
    const Mat4 r = quat_to_matrix(q);
    bool zero = true;
#pragma unroll
    for (int i = 0; i < 16; i++) zero = zero && (r.m[i >> 2][i & 3] == 0.0f);
    if (!zero) rot = r;
}

// pose_block modes
static const int POSE_PLAIN = 0, POSE_BLEND = 1, POSE_CHAIN = 2;
// shared memory of a block evaluating t3d_pose_chain records (24 floats per armature)
__host__ __device__ inline size_t pose_chain_smem_bytes(uint32_t nBones) {
    const uint32_t ipb = pose_armatures_per_block(nBones);
    return ((size_t)32 * ipb * nBones + (size_t)24 * ipb) * sizeof(float);
}

// The armatures [block * ipb, (block + 1) * ipb) of one batch, by the POSE_THREADS threads of a block (k_pose_bones, or the
// pose blocks at the start of k_prep_draws' grid).  POSE_BLEND: the players are in the middle of a blend between two animations
// (animation.go:852-907); POSE_CHAIN: of blends interrupted by further Play() calls (pb.poses then points at t3d_pose_chain records).
template <int MODE>
__device__ inline void pose_block(const PoseBatch& pb, const PoseBlend& bl, const PoseChain& ch, const uint32_t block, float* poseSmem) {
    constexpr uint32_t PS = MODE == POSE_CHAIN ? 24u : 20u;             // floats per pose record
    // shared: local / world matrices as 16 planes of `slots` floats each (a thread's matrix is a column: no bank
    // conflicts), then the block's pose records
    const uint32_t nb = pb.nBones;
    const uint32_t ipb = nb >= (uint32_t)POSE_THREADS ? 1u : (uint32_t)POSE_THREADS / nb;      // armatures per block
    const uint32_t slots = ipb * nb;
    float* sLocal = poseSmem;
    float* sWorld = poseSmem + 16 * (size_t)slots;
    float* sPose = sWorld + 16 * (size_t)slots;            // [ipb][PS]
    const uint32_t inst0 = block * ipb;
    const uint32_t nInst = min(ipb, pb.nPoses - inst0);
    {   // this block's pose records, straight from the pinned frame block (one PCIe round trip)
        const float* src = reinterpret_cast<const float*>(pb.poses) + (size_t)inst0 * PS;
        for (uint32_t i = threadIdx.x; i < nInst * PS; i += POSE_THREADS) sPose[i] = src[i];
    }
    __syncthreads();
    const uint32_t live = nInst * nb;
    for (uint32_t s = threadIdx.x; s < live; s += POSE_THREADS) {
        const uint32_t inst = s / nb, bone = s - inst * nb;
        float px = pb.restPos[3 * bone], py = pb.restPos[3 * bone + 1], pz = pb.restPos[3 * bone + 2];
        float sx = pb.restScale[3 * bone], sy = pb.restScale[3 * bone + 1], sz = pb.restScale[3 * bone + 2];
        Mat4 rot = m4_load(pb.restRot + 16 * (size_t)bone);
        if (MODE == POSE_CHAIN) {
            // the fold of the reference's two maps over the animations played since the last finished blend (t3dgeo.h): `cur` is
            // currentProperties[node], px .. rot what the Node holds
            BoneSample cur; cur.P = cur.S = cur.R = false;
            bool curDefined = false;
            for (uint32_t k = 0; k < ch.nLevels; k++) {
                const PoseChainLevel& L = ch.level[k];
                bool has = false;
                if (bone < L.bones) has = L.tracks[3 * bone].n_keys != 0 || L.tracks[3 * bone + 1].n_keys != 0 || L.tracks[3 * bone + 2].n_keys != 0;
                if (!has) { curDefined = false; continue; }
                BoneSample v = sample_bone(L.tracks, L.keyTimes, L.keyValues, bone, sPose[inst * PS + 16 + k]);
                if (k > 0 && curDefined) v = blend_sample(cur, v, sPose[inst * PS + 20 + k]);
                cur = v; curDefined = true;
                if (v.P) { px = v.px; py = v.py; pz = v.pz; }
                if (v.S) { sx = v.sx; sy = v.sy; sz = v.sz; }
                if (v.R) apply_rotation(v.q, rot);
            }
        } else {
            const float time = sPose[inst * PS + 16];
            BoneSample v;
            v.P = v.S = v.R = false;
            if (bone < pb.animBones) v = sample_bone(pb.tracks, pb.keyTimes, pb.keyValues, bone, time);
            if (MODE == POSE_BLEND) {
                if (bone < bl.fromBones)                                // prevAnimatedProperties[node] (animation.go:859)
                    v = blend_sample(sample_bone(bl.tracks, bl.keyTimes, bl.keyValues, bone, sPose[inst * PS + 17]), v, sPose[inst * PS + 18]);
            }
            if (v.P) { px = v.px; py = v.py; pz = v.pz; }
            if (v.S) { sx = v.sx; sy = v.sy; sz = v.sz; }
            if (v.R) apply_rotation(v.q, rot);
        }
        // Node.Transform (node.go:577-579): Scale x rotation x Translate
        Mat4 t = m4_mult(m4_scale(sx, sy, sz), rot);
        Mat4 tr = m4_identity(); tr.m[3][0] = px; tr.m[3][1] = py; tr.m[3][2] = pz;
        t = m4_mult(t, tr);
#pragma unroll
        for (int i = 0; i < 16; i++) sLocal[i * slots + s] = t.m[i >> 2][i & 3];
    }
    __syncthreads();
    // hierarchy, one depth level at a time: world = local x parent's world (node.go:581-583)
    for (uint32_t level = 0; level <= pb.maxDepth; level++) {
        for (uint32_t s = threadIdx.x; s < live; s += POSE_THREADS) {
            const uint32_t inst = s / nb, bone = s - inst * nb;
            if (pb.depth[bone] != level) continue;
            Mat4 l, p;
#pragma unroll
            for (int i = 0; i < 16; i++) l.m[i >> 2][i & 3] = sLocal[i * slots + s];
            const int32_t par = pb.parent[bone];
            if (par < 0) {
#pragma unroll
                for (int i = 0; i < 16; i++) p.m[i >> 2][i & 3] = sPose[inst * PS + i];
            } else {
                const uint32_t ps = inst * nb + (uint32_t)par;
#pragma unroll
                for (int i = 0; i < 16; i++) p.m[i >> 2][i & 3] = sWorld[i * slots + ps];
            }
            const Mat4 w = m4_mult(l, p);
#pragma unroll
            for (int i = 0; i < 16; i++) sWorld[i * slots + s] = w.m[i >> 2][i & 3];
        }
        __syncthreads();
    }
    // boneInfluence = inverseBindMatrix x Transform (node.go:592-594)
    for (uint32_t s = threadIdx.x; s < live; s += POSE_THREADS) {
        const uint32_t inst = s / nb, bone = s - inst * nb;
        Mat4 w;
#pragma unroll
        for (int i = 0; i < 16; i++) w.m[i >> 2][i & 3] = sWorld[i * slots + s];
        const Mat4 r = m4_mult(m4_load(pb.invBind + 16 * (size_t)bone), w);
        float4* o = reinterpret_cast<float4*>(pb.out + 16 * ((size_t)(inst0 + inst) * nb + bone));
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = make_float4(r.m[i][0], r.m[i][1], r.m[i][2], r.m[i][3]);
    }
}


}  // namespace t3d
