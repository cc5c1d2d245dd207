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
//
// Node record (64 B, see bvh_build.cu): child boxes are stored as CENTRE / HALF-EXTENT, half-extents
// rounded up.  The slab test then is, per axis,
//     tc = fma(c, inv, -o*inv) ; tnear = fma(-h, |inv|, tc) ; tfar = fma(h, |inv|, tc)
// i.e. 9 FFMA + 4 FMNMX(3) per box instead of 6 FFMA + 10 FMNMX for the min/max form: the min/max
// selection by ray direction is folded into |inv|.  On sm_100 FMNMX issues on the ALU pipe, which ncu
// showed to be the busiest pipe of the min/max form (profiles/r01_*), while the FMA pipe had headroom.
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

// per-ray constants of the slab test
struct RaySlab {
    float ix, iy, iz;      // 1/d (clamped away from 0)
    float ax, ay, az;      // |1/d|
    float nx, ny, nz;      // -o/d
};

__device__ __forceinline__ RaySlab make_slab(float ox, float oy, float oz, float dx, float dy, float dz)
{
    RaySlab s;
    s.ix = safe_inv(dx); s.iy = safe_inv(dy); s.iz = safe_inv(dz);
    s.ax = fabsf(s.ix); s.ay = fabsf(s.iy); s.az = fabsf(s.iz);
    s.nx = -(ox * s.ix); s.ny = -(oy * s.iy); s.nz = -(oz * s.iz);
    // keep the three products in registers: without this ptxas re-multiplies them at every node
    asm volatile("" : "+f"(s.nx), "+f"(s.ny), "+f"(s.nz));
    return s;
}

// centre/half-extent slab test; conservative because every leaf box carries an absolute pad that dwarfs the
// rounding of these nine FMAs and half-extents are rounded up (see bvh_build.cu).
__device__ __forceinline__ bool slab_ch(float cx, float cy, float cz, float hx, float hy, float hz, const RaySlab& s,
                                        float tmax, float& tnear)
{
    const float tcx = fmaf(cx, s.ix, s.nx), tcy = fmaf(cy, s.iy, s.ny), tcz = fmaf(cz, s.iz, s.nz);
    const float n0 = fmaf(-hx, s.ax, tcx), n1 = fmaf(-hy, s.ay, tcy), n2 = fmaf(-hz, s.az, tcz);
    const float f0 = fmaf(hx, s.ax, tcx), f1 = fmaf(hy, s.ay, tcy), f2 = fmaf(hz, s.az, tcz);
    const float t0 = fmaxf(fmaxf(n0, n1), fmaxf(n2, 0.f));
    const float t1 = fminf(fminf(f0, f1), fminf(f2, tmax));
    tnear = t0;
    return t0 <= t1;
}

// 256-bit global loads (sm_100+: LDG.E.256): a 64 B node record is two of them instead of 3 x 128 + 1 x 64 bits, a 48 B
// triangle record one 256 + one 128.  Halves the LSU instructions and L1 wavefronts per record; the record arrays are
// cudaMalloc'ed (256 B aligned) and records are 64 / 48 B, so a node is always 32 B aligned and a triangle 16 B aligned
// (its first 32 bytes are fetched as 2 x 128 when the slot is odd).
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const void* p)
{
    F8 r;
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
        : "l"(p));
    return r;
}

#define LRC_SENTINEL ((int)0x80000000)   // a "leaf" link no tree contains: ~0x80000000 = 0x7fffffff slots

// A leaf link is ~(first slot | (count - 1) << 28): up to 8 Morton-consecutive triangle records (lrc option "leaf_size").
__device__ __forceinline__ void leaf_test(const float4* __restrict__ tris, int link, float ox, float oy, float oz, float dx,
                                          float dy, float dz, float& best_t, uint32_t& best_id, unsigned& n_tris)
{
    const unsigned x = ~(unsigned)link;
    unsigned slot = x & 0x0fffffffu;
    const unsigned last = slot + (x >> 28);
    // slot indices (two 32-bit registers) are carried across the intersection test, not pointers -- the kernel is
    // register-bound; peeling the first triangle out of the loop costs 11 more registers and 3 % (measured)
#pragma unroll 1
    do {
        const float4* tp = tris + 3 * (int64_t)slot;
        const float4 v0 = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
        float t;
        ++n_tris;
        if (mt_hit(ox, oy, oz, dx, dy, dz, v0, e1, e2, t)) {
            const uint32_t id = __float_as_uint(v0.w);
            if (t < best_t || (t == best_t && id < best_id)) { best_t = t; best_id = id; }
        }
    } while (slot++ != last);
}

// ---- 32-byte nodes: 16-bit boxes on the scene's cell grid (VARIANT bit 5) ----------------------------------------------
// k_trace's ceiling is the L1 -> register return path: a 64 B float record is 56 B x 32 lanes = 14 cycles of data per
// node and warp.  The quantised record halves that.  Plane coordinate x = o + q * s, so along one axis
//     t(q) = (o + q*s - org) * inv = q * A + B          A = s * inv,  B = (o - org) * inv
// A 16-bit q becomes a float without a conversion instruction: PRMT splices it into the mantissa of 2^23
// (bits 0x4B000000 | q  ==  8388608 + q exactly), and the bias is folded into the constant:
//     t(q) = fma(8388608 + q, A, B - 8388608 * A)
// One PRMT + one FFMA per plane; which half-word is the near plane depends only on the ray's direction sign, so the PRMT
// selectors are per-ray constants.  Error budget: the folded constant is rounded once at magnitude 2^23 * |A|, i.e. by at
// most |A| / 2 = half a cell in t; everything else is ~1e-7 relative.  The builder widens every box by a full cell per
// side (bvh_build.cu quant_axis), so culling stays conservative and results stay bit-identical to the float format.
struct RaySlabQ {
    float ax, ay, az;          // A
    float bx, by, bz;          // B - 2^23 * A
    unsigned sx, sy, sz;       // PRMT selector of the NEAR plane's half-word (0x7610 = low, 0x7632 = high)
};

__device__ __forceinline__ RaySlabQ make_slab_q(float ox, float oy, float oz, float dx, float dy, float dz, const NodeQ& q)
{
    RaySlabQ s;
    const float ix = safe_inv(dx), iy = safe_inv(dy), iz = safe_inv(dz);
    s.ax = __fmul_rn(q.s[0], ix); s.ay = __fmul_rn(q.s[1], iy); s.az = __fmul_rn(q.s[2], iz);
    s.bx = __fmaf_rn(-8388608.f, s.ax, __fmul_rn(__fsub_rn(q.o[0], ox), ix));
    s.by = __fmaf_rn(-8388608.f, s.ay, __fmul_rn(__fsub_rn(q.o[1], oy), iy));
    s.bz = __fmaf_rn(-8388608.f, s.az, __fmul_rn(__fsub_rn(q.o[2], oz), iz));
    s.sx = ix >= 0.f ? 0x7610u : 0x7632u;
    s.sy = iy >= 0.f ? 0x7610u : 0x7632u;
    s.sz = iz >= 0.f ? 0x7610u : 0x7632u;
    return s;
}

__device__ __forceinline__ float q2f(unsigned w, unsigned sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)); }

__device__ __forceinline__ bool slab_q(unsigned wx, unsigned wy, unsigned wz, const RaySlabQ& s, float tmax, float& tnear)
{
    const float n0 = fmaf(q2f(wx, s.sx), s.ax, s.bx), f0 = fmaf(q2f(wx, s.sx ^ 0x22u), s.ax, s.bx);
    const float n1 = fmaf(q2f(wy, s.sy), s.ay, s.by), f1 = fmaf(q2f(wy, s.sy ^ 0x22u), s.ay, s.by);
    const float n2 = fmaf(q2f(wz, s.sz), s.az, s.bz), f2 = fmaf(q2f(wz, s.sz ^ 0x22u), s.az, s.bz);
    const float t0 = fmaxf(fmaxf(n0, n1), fmaxf(n2, 0.f));
    const float t1 = fminf(fminf(f0, f1), fminf(f2, tmax));
    tnear = t0;
    return t0 <= t1;
}

// Traversal stack.
// StackLocal: per-thread array (local memory: lanes with different depths touch different 128 B lines).  An empty stack
// pops the sentinel without touching memory.
// StackShared (VARIANT bit 4): the first `levels` entries live in shared memory as [level][thread] -- one conflict-free
// wavefront per access whatever the lanes' depths -- deeper entries spill to a local array.  Measured slower than the
// local stack at every depth (profiles/r01c_smem_stack_sweep.jsonl); selectable, default off.
#define LRC_SS_STRIDE 128      // threads per block of the traversal kernel (upper bound)
// (The stack pointer is a separate scalar on purpose: as a member next to the dynamically indexed array it is demoted to
// local memory with it -- measured +8 % kernel time.)
struct StackLocal {
    int a[LRC_STACK_DEPTH];
    __device__ __forceinline__ void push(int& sp, int v) { a[sp++] = v; }
    __device__ __forceinline__ int pop(int& sp) { return sp > 0 ? a[--sp] : LRC_SENTINEL; }
};
// StackCull (VARIANT bit 6): every entry carries the distance at which the ray enters that child's box; an entry whose
// distance already exceeds the best hit is dropped at pop time without fetching its node record.
struct StackCull {
    int a[LRC_STACK_DEPTH];
    float t[LRC_STACK_DEPTH];
    __device__ __forceinline__ void push(int& sp, int v, float tn) { a[sp] = v; t[sp] = tn; ++sp; }
    __device__ __forceinline__ int pop(int& sp, float best_t)
    {
        while (sp > 0) {
            --sp;
            if (t[sp] <= best_t) return a[sp];
        }
        return LRC_SENTINEL;
    }
};
struct StackShared {
    int a[LRC_STACK_DEPTH];
    __device__ __forceinline__ void push(int& sp, int v, int* sm, int levels) { if (sp < levels) sm[sp * LRC_SS_STRIDE] = v; else a[sp - levels] = v; ++sp; }
    __device__ __forceinline__ int pop(int& sp, int* sm, int levels)
    {
        if (sp == 0) return LRC_SENTINEL;
        --sp;
        return sp < levels ? sm[sp * LRC_SS_STRIDE] : a[sp - levels];
    }
};
// uniform access for both policies: sm / levels are ignored by the local stack
template <class STACK> struct StackOps;
template <> struct StackOps<StackLocal> {
    static __device__ __forceinline__ void push(StackLocal& st, int& sp, int v, float, int*, int) { st.push(sp, v); }
    static __device__ __forceinline__ int pop(StackLocal& st, int& sp, float, int*, int) { return st.pop(sp); }
};
template <> struct StackOps<StackCull> {
    static __device__ __forceinline__ void push(StackCull& st, int& sp, int v, float tn, int*, int) { st.push(sp, v, tn); }
    static __device__ __forceinline__ int pop(StackCull& st, int& sp, float best_t, int*, int) { return st.pop(sp, best_t); }
};
template <> struct StackOps<StackShared> {
    static __device__ __forceinline__ void push(StackShared& st, int& sp, int v, float, int* sm, int levels) { st.push(sp, v, sm, levels); }
    static __device__ __forceinline__ int pop(StackShared& st, int& sp, float, int* sm, int levels) { return st.pop(sp, sm, levels); }
};

// One step at an inner node: returns the next link (child to descend into, or a popped entry, or the sentinel).
// Top of the tree staged in shared memory (VARIANT bit 3): `s_top` holds the first top_n node records in HEAP order
// (entry h has its children at 2h+1 / 2h+2), copied there by every block; a link >= LRC_TOP_BASE addresses that table.
// Records are unmodified copies, so a child link is redirected into the table on the fly when its heap slot exists.
#define LRC_TOP_BASE 0x40000000   // node ids are < 2^28 (lrc_set_mesh)

template <bool COUNT, bool WIDE, bool TOP, class STACK>
__device__ __forceinline__ int inner_step(const float4* __restrict__ nodes, const float4* s_top, int top_n, int cur,
                                          const RaySlab& s, float best_t, STACK& stack, int& sp, int* sm, int levels,
                                          unsigned& n_nodes)
{
    float4 n0, n1, n2;
    int l0, l1;
    if (TOP && cur >= LRC_TOP_BASE) {
        const int h = cur - LRC_TOP_BASE;
        const float4* np = s_top + 4 * h;
        n0 = np[0]; n1 = np[1]; n2 = np[2];
        const float2 n3 = *reinterpret_cast<const float2*>(np + 3);
        l0 = __float_as_int(n3.x); l1 = __float_as_int(n3.y);
        const int c0 = 2 * h + 1;
        if (l0 >= 0 && c0 < top_n) l0 = LRC_TOP_BASE + c0;
        if (l1 >= 0 && c0 + 1 < top_n) l1 = LRC_TOP_BASE + c0 + 1;
    } else {
        const float4* np = nodes + 4 * (int64_t)cur;
        float2 n3;
        if (WIDE) {
            const F8 lo = ldg256(np), hi = ldg256(np + 2);
            n0 = lo.a; n1 = lo.b; n2 = hi.a; n3 = make_float2(hi.b.x, hi.b.y);
        } else {
            n0 = __ldg(np + 0); n1 = __ldg(np + 1); n2 = __ldg(np + 2);
            n3 = __ldg(reinterpret_cast<const float2*>(np + 3));
        }
        l0 = __float_as_int(n3.x); l1 = __float_as_int(n3.y);
    }
    if (COUNT) ++n_nodes;
    float t0, t1;
    const bool h0 = slab_ch(n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, s, best_t, t0);
    const bool h1 = slab_ch(n1.z, n1.w, n2.x, n2.y, n2.z, n2.w, s, best_t, t1);
    if (h0 && h1) {
        const bool swp = t1 < t0;              // nearer child first, farther one on the stack
        StackOps<STACK>::push(stack, sp, swp ? l0 : l1, swp ? t0 : t1, sm, levels);
        return swp ? l1 : l0;
    }
    if (h0) return l0;
    if (h1) return l1;
    return StackOps<STACK>::pop(stack, sp, best_t, sm, levels);
}

template <bool COUNT, class STACK>
__device__ __forceinline__ int inner_step_q(const float4* __restrict__ nodes, int cur, const RaySlabQ& s, float best_t,
                                            STACK& stack, int& sp, int* sm, int levels, unsigned& n_nodes)
{
    const uint4* np = reinterpret_cast<const uint4*>(nodes) + 2 * (int64_t)cur;
    const uint4 a = __ldg(np), b = __ldg(np + 1);
    if (COUNT) ++n_nodes;
    float t0, t1;
    const bool h0 = slab_q(a.x, a.y, a.z, s, best_t, t0);
    const bool h1 = slab_q(b.x, b.y, b.z, s, best_t, t1);
    const int l0 = (int)a.w, l1 = (int)b.w;
    if (h0 && h1) {
        const bool swp = t1 < t0;
        StackOps<STACK>::push(stack, sp, swp ? l0 : l1, swp ? t0 : t1, sm, levels);
        return swp ? l1 : l0;
    }
    if (h0) return l0;
    if (h1) return l1;
    return StackOps<STACK>::pop(stack, sp, best_t, sm, levels);
}

// ---- paired 64 B nodes + packed FP32 FMA (node format 2) -------------------------------------------------------------
// k_trace is bound by issue slots (ncu: 71 % of issue cycles busy, FMA pipe 30 %): 18 of the ~45 instructions of one inner
// step are the scalar FFMAs of the two slab tests.  sm_100 has a two-wide FP32 FMA (PTX fma.rn.f32x2 -> SASS FFMA2): one
// issue slot, two fused multiply-adds on an aligned register pair.  The paired record (bvh_build.cu write_node_p) puts the
// values that meet the same pair of ray constants next to each other, so one inner step is 9 FFMA2 instead of 18 FFMA --
// each lane-wise result is the same fused operation as before, so culling decisions (and node counts) are unchanged.
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c)
{
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
}

struct RaySlabP {
    float2 ixy, izz;       // 1/d
    float2 axy, azz;       // |1/d|
    float2 nxy, nzz;       // -o/d
};

__device__ __forceinline__ RaySlabP make_slab_p(float ox, float oy, float oz, float dx, float dy, float dz)
{
    const RaySlab r = make_slab(ox, oy, oz, dx, dy, dz);
    RaySlabP s;
    s.ixy = make_float2(r.ix, r.iy); s.izz = make_float2(r.iz, r.iz);
    s.axy = make_float2(r.ax, r.ay); s.azz = make_float2(r.az, r.az);
    s.nxy = make_float2(r.nx, r.ny); s.nzz = make_float2(r.nz, r.nz);
    return s;
}

// bring the record a stack entry points at (node or first triangle of a leaf) towards L1 while the near side is walked
__device__ __forceinline__ void prefetch_link(const float4* __restrict__ nodes, const float4* __restrict__ tris, int link)
{
    const float4* p = link >= 0 ? nodes + 4 * (int64_t)link : tris + 3 * (int64_t)((~(unsigned)link) & 0x0fffffffu);
    asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
}

// TEXM: which quarters of the record come through the texture pipe instead of the LSU (bit 0: Z and links, bit 1: A and B).
// ncu: the LSU data pipe and its register write-back are the busiest units of k_trace (77 % / 67 %) while the texture
// pipe of the same L1TEX unit idles; a linear float4 texture over the node array lets the two halves of a record return
// through different pipes.
template <bool COUNT, class STACK, bool PF = false, int TEXM = 0>
__device__ __forceinline__ int inner_step_p(const float4* __restrict__ nodes, int cur, const RaySlabP& s, float best_t,
                                            STACK& stack, int& sp, unsigned& n_nodes, const float4* __restrict__ tris = nullptr,
                                            cudaTextureObject_t ntex = 0)
{
    const float4* np = nodes + 4 * (int64_t)cur;
    float4 A, B, Z;
    float2 L;
    if (TEXM & 2) { A = tex1Dfetch<float4>(ntex, 4 * cur + 0); B = tex1Dfetch<float4>(ntex, 4 * cur + 1); }
    else { A = __ldg(np + 0); B = __ldg(np + 1); }
    if (TEXM & 1) {
        Z = tex1Dfetch<float4>(ntex, 4 * cur + 2);
        const float4 L4 = tex1Dfetch<float4>(ntex, 4 * cur + 3);
        L = make_float2(L4.x, L4.y);
    } else {
        Z = __ldg(np + 2);
        L = __ldg(reinterpret_cast<const float2*>(np + 3));
    }
    if (COUNT) ++n_nodes;
    const float2 naxy = make_float2(-s.axy.x, -s.axy.y), nazz = make_float2(-s.azz.x, -s.azz.y);
    const float2 tc0 = ffma2(make_float2(A.x, A.y), s.ixy, s.nxy);      // child 0: (x, y) slab centres
    const float2 tc1 = ffma2(make_float2(A.z, A.w), s.ixy, s.nxy);      // child 1
    const float2 tcz = ffma2(make_float2(Z.x, Z.y), s.izz, s.nzz);      // z slab centres of child 0 | child 1
    const float2 n0 = ffma2(make_float2(B.x, B.y), naxy, tc0), f0 = ffma2(make_float2(B.x, B.y), s.axy, tc0);
    const float2 n1 = ffma2(make_float2(B.z, B.w), naxy, tc1), f1 = ffma2(make_float2(B.z, B.w), s.axy, tc1);
    const float2 nz = ffma2(make_float2(Z.z, Z.w), nazz, tcz), fz = ffma2(make_float2(Z.z, Z.w), s.azz, tcz);
    const float t0 = fmaxf(fmaxf(n0.x, n0.y), fmaxf(nz.x, 0.f)), e0 = fminf(fminf(f0.x, f0.y), fminf(fz.x, best_t));
    const float t1 = fmaxf(fmaxf(n1.x, n1.y), fmaxf(nz.y, 0.f)), e1 = fminf(fminf(f1.x, f1.y), fminf(fz.y, best_t));
    const bool h0 = t0 <= e0, h1 = t1 <= e1;
    const int l0 = __float_as_int(L.x), l1 = __float_as_int(L.y);
    if (h0 && h1) {
        const bool swp = t1 < t0;
        if (PF) prefetch_link(nodes, tris, swp ? l0 : l1);
        StackOps<STACK>::push(stack, sp, swp ? l0 : l1, swp ? t0 : t1, nullptr, 0);
        return swp ? l1 : l0;
    }
    if (h0) return l0;
    if (h1) return l1;
    return StackOps<STACK>::pop(stack, sp, best_t, nullptr, 0);
}

template <bool COUNT, class STACK, bool PF = false, int TEXM = 0>
__device__ __forceinline__ void trace_loop_p(const float4* __restrict__ nodes, const float4* __restrict__ tris, int root, float ox,
                                             float oy, float oz, float dx, float dy, float dz, float& best_t, uint32_t& best_id,
                                             unsigned& n_nodes, unsigned& n_tris, cudaTextureObject_t ntex = 0)
{
    best_t = LRC_INF;
    best_id = LRC_MISS_ID;
    const RaySlabP s = make_slab_p(ox, oy, oz, dx, dy, dz);
    STACK stack;
    int sp = 0;
    int cur = root;
    while (cur != LRC_SENTINEL) {
        while (cur >= 0) cur = inner_step_p<COUNT, STACK, PF, TEXM>(nodes, cur, s, best_t, stack, sp, n_nodes, tris, ntex);
        if (cur != LRC_SENTINEL) {
            leaf_test(tris, cur, ox, oy, oz, dx, dy, dz, best_t, best_id, n_tris);
            cur = StackOps<STACK>::pop(stack, sp, best_t, nullptr, 0);
        }
    }
}

// ---- K rays per thread ("thread packets", paired node format) -----------------------------------------------------------
// What bounds the one-ray-per-thread kernel is neither HBM nor issue slots but the L1 -> register return path (ncu round 1:
// l1tex data-pipe wavefronts 75-80 % of peak, LSU write-back active 62 %): every lane pulls the 56 bytes of a node record
// into its own registers -- 14 cycles of the SM's 128 B/clk return path per warp and node, however uniform the addresses
// are (which is why neither shared-memory staging nor packed FMA moved the time).  The K rays of one thread are adjacent
// azimuths of one scan line (0.09 deg apart for the 32-line sensor): they walk almost the same nodes, so the thread fetches
// each record ONCE and tests it against all K rays -- 56/K bytes of write-back per ray and node, 1/K of the stack traffic
// and loop control.  A node is entered when ANY ray of the packet hits its box; each ray still culls with its own best
// hit, and the Moller-Trumbore test per ray is the unchanged bit-exact one, so results are identical -- the packet only
// visits the UNION of its rays' nodes.  Ray pairs share one FFMA2 per slab plane (node scalar broadcast to both halves).
template <int K> struct Packet {
    float ox[K], oy[K], oz[K], dx[K], dy[K], dz[K];
    float ix[K], iy[K], iz[K], nx[K], ny[K], nz[K];
    float best_t[K];
    uint32_t best_id[K];
};

template <int K>
__device__ __forceinline__ void packet_set_ray(Packet<K>& pk, int r, float ox, float oy, float oz, float dx, float dy, float dz, bool live)
{
    pk.ox[r] = ox; pk.oy[r] = oy; pk.oz[r] = oz; pk.dx[r] = dx; pk.dy[r] = dy; pk.dz[r] = dz;
    const RaySlab s = make_slab(ox, oy, oz, dx, dy, dz);
    pk.ix[r] = s.ix; pk.iy[r] = s.iy; pk.iz[r] = s.iz; pk.nx[r] = s.nx; pk.ny[r] = s.ny; pk.nz[r] = s.nz;
    pk.best_t[r] = live ? LRC_INF : -1.f;      // a ray that is not cast (past the end, dropped) can never enter a box: tfar < 0 <= tnear
    pk.best_id[r] = LRC_MISS_ID;
}

// slab planes of one axis of one child box for the ray pair (a, b): entry / exit distances
__device__ __forceinline__ void slab_axis2(float c, float h, const float2 I, const float2 N, float2& tn, float2& tf)
{
    const float2 cc = make_float2(c, c), hh = make_float2(h, h);
    const float2 A = make_float2(fabsf(I.x), fabsf(I.y));
    const float2 tc = ffma2(cc, I, N);
    tn = ffma2(hh, make_float2(-A.x, -A.y), tc);
    tf = ffma2(hh, A, tc);
}

template <int K, bool COUNT, class STACK>
__device__ __forceinline__ int inner_step_k(const float4* __restrict__ nodes, int cur, const Packet<K>& pk, STACK& stack, int& sp,
                                            float bmax, unsigned& n_nodes)
{
    const float4* np = nodes + 4 * (int64_t)cur;
    const float4 A = __ldg(np + 0), B = __ldg(np + 1), Z = __ldg(np + 2);
    const float2 L = __ldg(reinterpret_cast<const float2*>(np + 3));
    if (COUNT) ++n_nodes;
    float m0 = LRC_INF, m1 = LRC_INF;             // smallest entry distance among the rays that hit child 0 / child 1
#pragma unroll
    for (int p = 0; p < K; p += 2) {
        const float2 Ix = make_float2(pk.ix[p], pk.ix[p + 1]), Iy = make_float2(pk.iy[p], pk.iy[p + 1]), Iz = make_float2(pk.iz[p], pk.iz[p + 1]);
        const float2 Nx = make_float2(pk.nx[p], pk.nx[p + 1]), Ny = make_float2(pk.ny[p], pk.ny[p + 1]), Nz = make_float2(pk.nz[p], pk.nz[p + 1]);
        float2 nx0, fx0, ny0, fy0, nz0, fz0, nx1, fx1, ny1, fy1, nz1, fz1;
        slab_axis2(A.x, B.x, Ix, Nx, nx0, fx0); slab_axis2(A.y, B.y, Iy, Ny, ny0, fy0); slab_axis2(Z.x, Z.z, Iz, Nz, nz0, fz0);
        slab_axis2(A.z, B.z, Ix, Nx, nx1, fx1); slab_axis2(A.w, B.w, Iy, Ny, ny1, fy1); slab_axis2(Z.y, Z.w, Iz, Nz, nz1, fz1);
        {
            const float t0 = fmaxf(fmaxf(nx0.x, ny0.x), fmaxf(nz0.x, 0.f)), e0 = fminf(fminf(fx0.x, fy0.x), fminf(fz0.x, pk.best_t[p]));
            const float t1 = fmaxf(fmaxf(nx1.x, ny1.x), fmaxf(nz1.x, 0.f)), e1 = fminf(fminf(fx1.x, fy1.x), fminf(fz1.x, pk.best_t[p]));
            m0 = fminf(m0, t0 <= e0 ? t0 : LRC_INF);
            m1 = fminf(m1, t1 <= e1 ? t1 : LRC_INF);
        }
        {
            const float t0 = fmaxf(fmaxf(nx0.y, ny0.y), fmaxf(nz0.y, 0.f)), e0 = fminf(fminf(fx0.y, fy0.y), fminf(fz0.y, pk.best_t[p + 1]));
            const float t1 = fmaxf(fmaxf(nx1.y, ny1.y), fmaxf(nz1.y, 0.f)), e1 = fminf(fminf(fx1.y, fy1.y), fminf(fz1.y, pk.best_t[p + 1]));
            m0 = fminf(m0, t0 <= e0 ? t0 : LRC_INF);
            m1 = fminf(m1, t1 <= e1 ? t1 : LRC_INF);
        }
    }
    const bool h0 = m0 < LRC_INF, h1 = m1 < LRC_INF;
    const int l0 = __float_as_int(L.x), l1 = __float_as_int(L.y);
    if (h0 && h1) {
        const bool swp = m1 < m0;                   // the child some ray enters first goes first, the other one on the stack
        StackOps<STACK>::push(stack, sp, swp ? l0 : l1, swp ? m0 : m1, nullptr, 0);
        return swp ? l1 : l0;
    }
    if (h0) return l0;
    if (h1) return l1;
    return StackOps<STACK>::pop(stack, sp, bmax, nullptr, 0);
}

template <int K>
__device__ __forceinline__ void leaf_test_k(const float4* __restrict__ tris, int link, Packet<K>& pk, unsigned& n_tris)
{
    const unsigned x = ~(unsigned)link;
    unsigned slot = x & 0x0fffffffu;
    const unsigned last = slot + (x >> 28);
#pragma unroll 1
    do {
        const float4* tp = tris + 3 * (int64_t)slot;
        const float4 v0 = __ldg(tp + 0), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
        const uint32_t id = __float_as_uint(v0.w);
        ++n_tris;
#pragma unroll
        for (int r = 0; r < K; ++r) {
            float t;
            if (mt_hit(pk.ox[r], pk.oy[r], pk.oz[r], pk.dx[r], pk.dy[r], pk.dz[r], v0, e1, e2, t)) {
                if (t < pk.best_t[r] || (t == pk.best_t[r] && id < pk.best_id[r])) { pk.best_t[r] = t; pk.best_id[r] = id; }
            }
        }
    } while (slot++ != last);
}

template <int K> __device__ __forceinline__ float packet_bmax(const Packet<K>& pk)
{
    float b = pk.best_t[0];
#pragma unroll
    for (int r = 1; r < K; ++r) b = fmaxf(b, pk.best_t[r]);
    return b;
}

// closest hits of the K rays in pk (best_t = +inf / id = MISS on a miss; rays set up with live = false stay at -1 / MISS)
template <int K, bool COUNT>
__device__ __forceinline__ void trace_packet(const float4* __restrict__ nodes, const float4* __restrict__ tris, int root, Packet<K>& pk,
                                             unsigned& n_nodes, unsigned& n_tris)
{
    StackCull stack;
    int sp = 0;
    int cur = root;
    while (cur != LRC_SENTINEL) {
        const float bmax = packet_bmax(pk);
        while (cur >= 0) cur = inner_step_k<K, COUNT>(nodes, cur, pk, stack, sp, bmax, n_nodes);
        if (cur != LRC_SENTINEL) {
            leaf_test_k<K>(tris, cur, pk, n_tris);
            cur = StackOps<StackCull>::pop(stack, sp, packet_bmax(pk), nullptr, 0);
        }
    }
}

// ---- warp packets: the 32 rays of a warp walk ONE node at a time (paired node format) -----------------------------------
// ncu on the one-ray-per-thread kernel: l1tex data-pipe wavefronts 77 % of peak with ~3.3 wavefronts per load instruction --
// the lanes of a warp sit at ~3 different nodes per step, and every lane keeps its own stack in local memory.  Here `cur`
// and the stack are WARP-uniform: every record is fetched once per warp at a lane-uniform address (one wavefront per load),
// the 32 rays are slab-tested against it, a child is entered when ANY lane hits it (ballot), the nearer child -- by the
// smallest entry distance over the lanes that hit, a REDUX over the float bits, monotone for t >= 0 -- goes first and the
// other one onto one shared-memory stack per warp, which also carries that smallest entry distance so that a popped entry
// is dropped when no lane can still improve (largest best hit over the live lanes).  Every lane still culls with its OWN
// best hit and runs the unchanged Moller-Trumbore test, so the output bits do not change: the warp visits the union of its
// rays' nodes.  Rays of a warp are 32 adjacent beams of one scan line, which is what keeps that union small.
#define LRC_WSTACK 64     // entries per warp (tree height is checked against LRC_STACK_DEPTH = 64 at build time)

__device__ __forceinline__ int wstack_pop(const int* s_link, const unsigned* s_t, int& sp, unsigned bmax)
{
    while (sp > 0) {
        --sp;
        if (s_t[sp] <= bmax) return s_link[sp];
    }
    return LRC_SENTINEL;
}

template <bool COUNT>
__device__ __forceinline__ void trace_warp(const float4* __restrict__ nodes, const float4* __restrict__ tris, int root, int* s_link,
                                           unsigned* s_t, float ox, float oy, float oz, float dx, float dy, float dz, bool live,
                                           float& best_t, uint32_t& best_id, unsigned& n_nodes, unsigned& n_tris)
{
    const unsigned FULL = 0xffffffffu;
    best_t = live ? LRC_INF : -1.f;         // a lane without a ray can never enter a box (tfar < 0 <= tnear) nor accept a hit
    best_id = LRC_MISS_ID;
    const RaySlabP s = make_slab_p(ox, oy, oz, dx, dy, dz);
    const float2 naxy = make_float2(-s.axy.x, -s.axy.y), nazz = make_float2(-s.azz.x, -s.azz.y);
    int sp = 0;
    unsigned bmax = __reduce_max_sync(FULL, live ? 0x7f800000u : 0u);      // bits of the largest best hit among live lanes
    int cur = bmax ? root : LRC_SENTINEL;
    while (cur != LRC_SENTINEL) {
        while (cur >= 0) {
            const float4* np = nodes + 4 * (int64_t)cur;
            const float4 A = __ldg(np + 0), B = __ldg(np + 1), Z = __ldg(np + 2);
            const float2 L = __ldg(reinterpret_cast<const float2*>(np + 3));
            if (COUNT) ++n_nodes;
            const float2 tc0 = ffma2(make_float2(A.x, A.y), s.ixy, s.nxy);
            const float2 tc1 = ffma2(make_float2(A.z, A.w), s.ixy, s.nxy);
            const float2 tcz = ffma2(make_float2(Z.x, Z.y), s.izz, s.nzz);
            const float2 n0 = ffma2(make_float2(B.x, B.y), naxy, tc0), f0 = ffma2(make_float2(B.x, B.y), s.axy, tc0);
            const float2 n1 = ffma2(make_float2(B.z, B.w), naxy, tc1), f1 = ffma2(make_float2(B.z, B.w), s.axy, tc1);
            const float2 nz = ffma2(make_float2(Z.z, Z.w), nazz, tcz), fz = ffma2(make_float2(Z.z, Z.w), s.azz, tcz);
            const float t0 = fmaxf(fmaxf(n0.x, n0.y), fmaxf(nz.x, 0.f)), e0 = fminf(fminf(f0.x, f0.y), fminf(fz.x, best_t));
            const float t1 = fmaxf(fmaxf(n1.x, n1.y), fmaxf(nz.y, 0.f)), e1 = fminf(fminf(f1.x, f1.y), fminf(fz.y, best_t));
            const bool h0 = t0 <= e0, h1 = t1 <= e1;
            const unsigned b0 = __ballot_sync(FULL, h0), b1 = __ballot_sync(FULL, h1);
            const int l0 = __float_as_int(L.x), l1 = __float_as_int(L.y);
            if (b0 && b1) {
                const unsigned m0 = __reduce_min_sync(FULL, h0 ? __float_as_uint(t0) : 0x7f800000u);
                const unsigned m1 = __reduce_min_sync(FULL, h1 ? __float_as_uint(t1) : 0x7f800000u);
                const bool swp = m1 < m0;
                s_link[sp] = swp ? l0 : l1;              // every lane stores the same value to the same word
                s_t[sp] = swp ? m0 : m1;
                ++sp;
                cur = swp ? l1 : l0;
            } else if (b0) cur = l0;
            else if (b1) cur = l1;
            else cur = wstack_pop(s_link, s_t, sp, bmax);
        }
        if (cur != LRC_SENTINEL) {
            leaf_test(tris, cur, ox, oy, oz, dx, dy, dz, best_t, best_id, n_tris);
            bmax = __reduce_max_sync(FULL, live ? __float_as_uint(best_t) : 0u);
            cur = wstack_pop(s_link, s_t, sp, bmax);
        }
    }
}

// Stack-based closest-hit traversal.  VARIANT bit 0: 0 = one node (inner or leaf) per loop trip ("if-if"),
// 1 = "while-while" -- run down inner nodes until a leaf (or the end) is reached, then test the leaf.
// VARIANT bit 1: node records fetched with 256-bit loads.  VARIANT bit 2 (scan.cu): 32-register cap (64 warps per SM).
// VARIANT bit 3: the first top_n nodes (heap order) are read from shared memory.
// VARIANT bit 4: the first stack levels live in shared memory (StackShared).
// VARIANT bit 5: 32-byte quantised node records (trace_loop_q; while-while only).
// VARIANT bit 7: paired node records + packed FMA (trace_loop_p, node format 2; with or without bit 6).
// VARIANT bit 6: stack entries carry their entry distance and are culled against the best hit at pop time (StackCull).
// VARIANT bit 10: the record of a pushed entry is prefetched towards L1 (paired format).  Bits 11 / 12 live in scan.cu:
// per-warp keep counts combined through shared-memory atomics instead of a block-wide barrier; streaming scratch stores.
template <int VARIANT, bool COUNT, class STACK>
__device__ __forceinline__ void trace_loop(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                           const float4* s_top, int top_n, int root, STACK& stack, int* sm, int levels, float ox,
                                           float oy, float oz, float dx, float dy, float dz, float& best_t,
                                           uint32_t& best_id, unsigned& n_nodes, unsigned& n_tris)
{
    best_t = LRC_INF;
    best_id = LRC_MISS_ID;
    const RaySlab s = make_slab(ox, oy, oz, dx, dy, dz);
    int sp = 0;
    constexpr bool WIDE = (VARIANT & 2) != 0;
    constexpr bool TOP = (VARIANT & 8) != 0;
    int cur = (TOP && top_n > 0) ? LRC_TOP_BASE : root;
    if ((VARIANT & 1) == 0) {
        while (cur != LRC_SENTINEL) {
            if (cur >= 0) {
                cur = inner_step<COUNT, WIDE, TOP>(nodes, s_top, top_n, cur, s, best_t, stack, sp, sm, levels, n_nodes);
            } else {
                leaf_test(tris, cur, ox, oy, oz, dx, dy, dz, best_t, best_id, n_tris);
                cur = StackOps<STACK>::pop(stack, sp, best_t, sm, levels);
            }
        }
    } else {
        while (cur != LRC_SENTINEL) {
            while (cur >= 0) cur = inner_step<COUNT, WIDE, TOP>(nodes, s_top, top_n, cur, s, best_t, stack, sp, sm, levels, n_nodes);
            if (cur != LRC_SENTINEL) {
                leaf_test(tris, cur, ox, oy, oz, dx, dy, dz, best_t, best_id, n_tris);
                cur = StackOps<STACK>::pop(stack, sp, best_t, sm, levels);
            }
        }
    }
}

template <bool COUNT>
__device__ __forceinline__ void trace_loop_q(const float4* __restrict__ nodes, const float4* __restrict__ tris, const NodeQ& nq,
                                             int root, float ox, float oy, float oz, float dx, float dy, float dz, float& best_t,
                                             uint32_t& best_id, unsigned& n_nodes, unsigned& n_tris)
{
    best_t = LRC_INF;
    best_id = LRC_MISS_ID;
    const RaySlabQ s = make_slab_q(ox, oy, oz, dx, dy, dz, nq);
    StackLocal stack;
    int sp = 0;
    int cur = root;
    while (cur != LRC_SENTINEL) {
        while (cur >= 0) cur = inner_step_q<COUNT>(nodes, cur, s, best_t, stack, sp, nullptr, 0, n_nodes);
        if (cur != LRC_SENTINEL) {
            leaf_test(tris, cur, ox, oy, oz, dx, dy, dz, best_t, best_id, n_tris);
            cur = StackOps<StackLocal>::pop(stack, sp, best_t, nullptr, 0);
        }
    }
}

// `smem` = the block's dynamic shared memory: the heap-ordered top table (bit 3) or the shared stack levels (bit 4).
template <int VARIANT, bool COUNT>
__device__ __forceinline__ void trace_ray(const float4* __restrict__ nodes, const float4* __restrict__ tris,
                                          const float4* smem, int top_n, int stack_levels, const NodeQ& nq, int root, float ox,
                                          float oy, float oz, float dx, float dy, float dz, float& best_t,
                                          uint32_t& best_id, unsigned& n_nodes, unsigned& n_tris, cudaTextureObject_t ntex = 0)
{
    if (VARIANT & 128) {
        if (VARIANT & (8192 | 16384)) trace_loop_p<COUNT, StackCull, false, ((VARIANT >> 13) & 3)>(nodes, tris, root, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris, ntex);
        else if (VARIANT & 1024) trace_loop_p<COUNT, StackCull, true>(nodes, tris, root, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
        else if (VARIANT & 64) trace_loop_p<COUNT, StackCull>(nodes, tris, root, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
        else trace_loop_p<COUNT, StackLocal>(nodes, tris, root, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
    } else if (VARIANT & 32) {
        trace_loop_q<COUNT>(nodes, tris, nq, root, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
    } else if (VARIANT & 16) {
        StackShared st;
        int* sm = reinterpret_cast<int*>(const_cast<float4*>(smem)) + threadIdx.x;   // this thread's column of the shared table
        trace_loop<VARIANT, COUNT>(nodes, tris, smem, top_n, root, st, sm, stack_levels, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
    } else if (VARIANT & 64) {
        StackCull st;
        trace_loop<VARIANT, COUNT>(nodes, tris, smem, top_n, root, st, nullptr, 0, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
    } else {
        StackLocal st;
        trace_loop<VARIANT, COUNT>(nodes, tris, smem, top_n, root, st, nullptr, 0, ox, oy, oz, dx, dy, dz, best_t, best_id, n_nodes, n_tris);
    }
}
