// traverse.cuh -- per-ray closest-hit traversal of the LBVH + the float32 intersection spec.
//
// Intersection spec (bit-exact with oracle/lrc_oracle.c::mt_f32; every rounding is spelled out with
// _rn intrinsics so that neither nvcc's FMA contraction nor --use_fast_math can change a bit):
//     p = d x e2 ; det = e1 . p ; tv = o - v0 ; U = tv . p ; q = tv x e1 ; V = d . q ; W = e2 . q
//     cross component  = fma(a, b, -(c * e))         dot = fma(x, x', fma(y, y', z * z'))
//     det < 0 -> negate det, U, V, W
//     hit <=> det > 0 && U >= 0 && V >= 0 && (U + V) <= det && W >= 0 ;  t = W / det  (IEEE division)
// Closest hit: smallest t, ties -> smallest original triangle id.  Two-sided, t >= 0, t in units of |d|
// (Embree contract behind reference raycast_engine_cpu.py:51; the reference passes directions as given).
#pragma once
#include "common.cuh"

__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz)
{
    return __fmaf_rn(ax, bx, __fmaf_rn(ay, by, __fmul_rn(az, bz)));
}

__device__ __forceinline__ bool mt_hit(float ox, float oy, float oz, float dx, float dy, float dz, const float4 v0,
                                       const float4 e1, const float4 e2, float& t)
{
    float px = __fmaf_rn(dy, e2.z, -__fmul_rn(dz, e2.y));
    float py = __fmaf_rn(dz, e2.x, -__fmul_rn(dx, e2.z));
    float pz = __fmaf_rn(dx, e2.y, -__fmul_rn(dy, e2.x));
    float det = dot3_rn(e1.x, e1.y, e1.z, px, py, pz);
    float tx = __fsub_rn(ox, v0.x), ty = __fsub_rn(oy, v0.y), tz = __fsub_rn(oz, v0.z);
    float U = dot3_rn(tx, ty, tz, px, py, pz);
    float qx = __fmaf_rn(ty, e1.z, -__fmul_rn(tz, e1.y));
    float qy = __fmaf_rn(tz, e1.x, -__fmul_rn(tx, e1.z));
    float qz = __fmaf_rn(tx, e1.y, -__fmul_rn(ty, e1.x));
    float V = dot3_rn(dx, dy, dz, qx, qy, qz);
    float W = dot3_rn(e2.x, e2.y, e2.z, qx, qy, qz);
    if (det < 0.f) { det = -det; U = -U; V = -V; W = -W; }
    if (det > 0.f && U >= 0.f && V >= 0.f && __fadd_rn(U, V) <= det && W >= 0.f) {
        float tt = __fdiv_rn(W, det);
        if (tt < LRC_INF) { t = tt; return true; }
    }
    return false;
}

__device__ __forceinline__ float safe_inv(float d)
{
    const float eps = 1e-20f;
    d = fabsf(d) < eps ? copysignf(eps, d) : d;
    return __frcp_rn(d);
}

// Slab test of one child box against the ray in (inv, ood = o * inv) form; conservative because every
// leaf box carries an absolute pad that dwarfs the rounding of these six FMAs (see bvh_build.cu).
__device__ __forceinline__ bool slab(float lox, float loy, float loz, float hix, float hiy, float hiz, float ix, float iy,
                                     float iz, float oox, float ooy, float ooz, float tmax, float& tnear)
{
    float ax = fmaf(lox, ix, -oox), bx = fmaf(hix, ix, -oox);
    float ay = fmaf(loy, iy, -ooy), by = fmaf(hiy, iy, -ooy);
    float az = fmaf(loz, iz, -ooz), bz = fmaf(hiz, iz, -ooz);
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), tmax));
    tnear = t0;
    return t0 <= t1;
}

// Stack-based closest-hit traversal.  nodes/tris layout: see bvh_build.cu.
template <bool COUNT>
__device__ __forceinline__ void trace_ray(const float4* __restrict__ nodes, const float4* __restrict__ tris, float ox,
                                          float oy, float oz, float dx, float dy, float dz, float& best_t,
                                          uint32_t& best_id, unsigned& n_nodes, unsigned& n_tris)
{
    best_t = LRC_INF;
    best_id = LRC_MISS_ID;
    const float ix = safe_inv(dx), iy = safe_inv(dy), iz = safe_inv(dz);
    const float oox = ox * ix, ooy = oy * iy, ooz = oz * iz;
    int stack[LRC_STACK_DEPTH];
    int sp = 0;
    int cur = 0;
    for (;;) {
        if (cur >= 0) {
            const float4* np = nodes + 4 * (int64_t)cur;
            const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
            if (COUNT) ++n_nodes;
            float t0, t1;
            const bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, ix, iy, iz, oox, ooy, ooz, best_t, t0);
            const bool h1 = slab(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, ix, iy, iz, oox, ooy, ooz, best_t, t1);
            const int l0 = __float_as_int(n3.x), l1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swp = t1 < t0;          // nearer child first, farther one on the stack
                cur = swp ? l1 : l0;
                stack[sp++] = swp ? l0 : l1;
                continue;
            }
            if (h0 || h1) { cur = h0 ? l0 : l1; continue; }
        } else {
            const float4* tp = tris + 3 * (int64_t)(~cur);
            const float4 v0 = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
            if (COUNT) ++n_tris;
            float t;
            if (mt_hit(ox, oy, oz, dx, dy, dz, v0, e1, e2, t)) {
                const uint32_t id = __float_as_uint(v0.w);
                if (t < best_t || (t == best_t && id < best_id)) { best_t = t; best_id = id; }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
}
