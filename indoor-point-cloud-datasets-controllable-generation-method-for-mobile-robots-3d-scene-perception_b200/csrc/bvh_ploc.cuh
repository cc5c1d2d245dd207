// bvh_ploc.cuh -- the "build_quality" = 1 builder: parallel locally-ordered clustering (PLOC, Meister & Bittner 2018)
// over the Morton-sorted leaves.  Included by bvh_build.cu only (uses its BuildMeta / write_node helpers).
//
// Why: the Karras radix tree splits where Morton bits change, not where surface area says so -- on the 1M-triangle office
// a LiDAR ray fetches 36 of its node records against 27 for a binned-SAH tree (oracle).  PLOC builds bottom-up: every
// cluster looks at the R clusters before and after it in Morton order, picks the one whose union box has the smallest
// area, and mutually-nearest pairs merge; the survivors are compacted (order kept) and the step repeats until one
// cluster is left.  The tree it yields is of SAH-sweep quality at a small multiple of the LBVH build time.
//
//   k_ploc_init      clusters = leaves (Morton order): id ~slot, padded leaf box, count 1
//   per iteration    k_ploc_search  nearest neighbour inside the window (shared-memory tile with halo), mutual pairs ->
//                                   role (0 stay, 1 merge as owner, 2 absorbed) + per-block counts
//                    k_ploc_scan    exclusive scan of the block counts (one block), totals to pinned host memory
//                    k_ploc_merge   owners create the inner node (children, union box, counts, parents); survivors are
//                                   written to the other cluster buffer at their compacted position
//   k_ploc_tail      once <= 1024 clusters are left: the same loop inside ONE block, clusters in shared memory
//   k_ploc_positions every leaf / inner node walks to the root: first leaf slot of its subtree in depth-first order
//   k_ploc_emit      64 B / 32 B node records, numbered like the radix tree's (left child = last leaf slot of the left
//                    subtree, right child = that + 1, root 0): sibling records share a 128 B line and a subtree's records
//                    and triangles are contiguous; subtrees of <= leaf_size triangles become one leaf link
//   k_tri_records    48 B triangle records in depth-first leaf order
// Merge decisions depend only on the data (ties: area, then a symmetric hash of the pair, then index), node ids come
// from prefix sums, so every rank of a multi-GPU run builds the same tree.
#pragma once

namespace {

constexpr int PL_THREADS = 256;
constexpr int PL_TAIL = 1024;           // clusters handled by the single-block tail kernel
constexpr int PL_MAX_R = 32;

struct PlocTree {            // device arrays, T-1 inner nodes (creation order; the root is the last one)
    int* left; int* right;   // child: >= 0 inner node id, < 0 leaf ~slot (Morton slot)
    int* cl;                 // leaves under the left child
    float4* lo; float4* hi;  // node box; lo.w = leaves under the node (int bits)
    int* parent_node;        // parent of an inner node (-1 root)
    int* parent_leaf;        // parent of a leaf (by Morton slot)
    int* start_node;         // first depth-first leaf slot under an inner node
    int* start_leaf;         // depth-first slot of a leaf
};

__device__ __forceinline__ unsigned ploc_pair_hash(unsigned a, unsigned b)   // symmetric in use: called with (min, max)
{
    unsigned x = a * 0x9E3779B1u + b * 0x85EBCA77u;
    x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
    return x;
}

__device__ __forceinline__ float ploc_area(float lx, float ly, float lz, float hx, float hy, float hz)
{
    const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
    return __fmaf_rn(ex, ey, __fmaf_rn(ey, ez, __fmul_rn(ez, ex)));
}

__global__ void k_ploc_init(int n, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, int* __restrict__ cid,
                            float4* __restrict__ clo, float4* __restrict__ chi)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 l = leaf_lo[i];
    l.w = __int_as_float(1);
    cid[i] = ~i;
    clo[i] = l;
    chi[i] = leaf_hi[i];
}

// nearest neighbour of cluster g among [g-R, g+R] in an SoA tile; w = tile index of g, g0 = global index of tile entry 0
template <int STRIDE>
__device__ __forceinline__ int ploc_nearest(const float (*s_lo)[STRIDE], const float (*s_hi)[STRIDE], int w, int g, int g0, int n, int R)
{
    const float lx = s_lo[0][w], ly = s_lo[1][w], lz = s_lo[2][w], hx = s_hi[0][w], hy = s_hi[1][w], hz = s_hi[2][w];
    int best = -1;
    float best_a = LRC_INF;
    unsigned best_h = 0xffffffffu;
    for (int k = -R; k <= R; ++k) {
        const int j = g + k;
        if (k == 0 || j < 0 || j >= n) continue;
        const int v = j - g0;
        const float a = ploc_area(fminf(lx, s_lo[0][v]), fminf(ly, s_lo[1][v]), fminf(lz, s_lo[2][v]),
                                  fmaxf(hx, s_hi[0][v]), fmaxf(hy, s_hi[1][v]), fmaxf(hz, s_hi[2][v]));
        if (a > best_a) continue;
        const unsigned h = ploc_pair_hash((unsigned)min(g, j), (unsigned)max(g, j));
        if (a < best_a || h < best_h) { best_a = a; best_h = h; best = j; }
    }
    return best;
}

__global__ void __launch_bounds__(PL_THREADS)
k_ploc_search(int n, int R, const float4* __restrict__ lo, const float4* __restrict__ hi, int* __restrict__ nn,
              unsigned char* __restrict__ role, uint2* __restrict__ bsum)
{
    constexpr int WIN = PL_THREADS + 4 * PL_MAX_R;
    constexpr int NNW = PL_THREADS + 2 * PL_MAX_R;
    __shared__ float s_lo[3][WIN], s_hi[3][WIN];
    __shared__ int s_nn[NNW];
    const int b0 = blockIdx.x * PL_THREADS;
    const int g0 = b0 - 2 * R;                       // global index of tile entry 0
    for (int t = threadIdx.x; t < PL_THREADS + 4 * R; t += PL_THREADS) {
        const int g = g0 + t;
        if (g >= 0 && g < n) {
            const float4 a = lo[g], b = hi[g];
            s_lo[0][t] = a.x; s_lo[1][t] = a.y; s_lo[2][t] = a.z;
            s_hi[0][t] = b.x; s_hi[1][t] = b.y; s_hi[2][t] = b.z;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < PL_THREADS + 2 * R; t += PL_THREADS) {     // the block's clusters and a halo of R
        const int g = b0 - R + t;
        s_nn[t] = (g >= 0 && g < n) ? ploc_nearest<WIN>(s_lo, s_hi, t + R, g, g0, n, R) : -1;
    }
    __syncthreads();
    const int g = b0 + threadIdx.x;
    int r = 0;
    if (g < n) {
        const int j = s_nn[R + threadIdx.x];
        if (j >= 0 && s_nn[j - (b0 - R)] == g) r = g < j ? 1 : 2;
        nn[g] = j;
        role[g] = (unsigned char)r;
    }
    const int c1 = __syncthreads_count(r == 1), c2 = __syncthreads_count(r == 2);
    if (threadIdx.x == 0) bsum[blockIdx.x] = make_uint2((unsigned)c1, (unsigned)c2);
}

// exclusive scan of the per-block (owners, absorbed) counts; totals -> tot[0..1] (device) and h_tot (pinned host)
__global__ void __launch_bounds__(1024) k_ploc_scan(uint2* __restrict__ a, int nb, int* __restrict__ h_tot)
{
    __shared__ uint2 warp_sums[32];
    __shared__ uint2 carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = make_uint2(0u, 0u);
    __syncthreads();
    for (int b0 = 0; b0 < nb; b0 += 1024) {
        const int i = b0 + threadIdx.x;
        const uint2 v = i < nb ? a[i] : make_uint2(0u, 0u);
        uint2 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned tx = __shfl_up_sync(0xffffffffu, incl.x, o), ty = __shfl_up_sync(0xffffffffu, incl.y, o);
            if (lane >= o) { incl.x += tx; incl.y += ty; }
        }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0) {
            const uint2 sv = warp_sums[lane];
            uint2 si = sv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned tx = __shfl_up_sync(0xffffffffu, si.x, o), ty = __shfl_up_sync(0xffffffffu, si.y, o);
                if (lane >= o) { si.x += tx; si.y += ty; }
            }
            warp_sums[lane] = make_uint2(si.x - sv.x, si.y - sv.y);
        }
        __syncthreads();
        const uint2 ws = warp_sums[w], c = carry;
        const uint2 excl = make_uint2(incl.x - v.x + ws.x + c.x, incl.y - v.y + ws.y + c.y);
        if (i < nb) a[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = make_uint2(excl.x + v.x, excl.y + v.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) { h_tot[0] = (int)carry.x; h_tot[1] = (int)carry.y; }
}

// exclusive scan of two flags inside a block (blockDim.x a multiple of 32, <= 1024); totals in tot1 / tot2
__device__ __forceinline__ void ploc_block_scan2(bool f1, bool f2, int& ex1, int& ex2, int& tot1, int& tot2, int (*s_w)[2])
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const unsigned b1 = __ballot_sync(0xffffffffu, f1), b2 = __ballot_sync(0xffffffffu, f2);
    const unsigned below = (1u << lane) - 1u;
    __syncthreads();                        // s_w may still be read from a previous call
    if (lane == 0) { s_w[w][0] = __popc(b1); s_w[w][1] = __popc(b2); }
    __syncthreads();
    int o1 = 0, o2 = 0, t1 = 0, t2 = 0;
    for (int k = 0; k < nw; ++k) {
        const int a = s_w[k][0], b = s_w[k][1];
        if (k < w) { o1 += a; o2 += b; }
        t1 += a; t2 += b;
    }
    ex1 = o1 + __popc(b1 & below);
    ex2 = o2 + __popc(b2 & below);
    tot1 = t1; tot2 = t2;
}

__device__ __forceinline__ void ploc_make_node(const PlocTree& tr, int node, int idl, int idr, const float4 llo, const float4 lhi,
                                               const float4 rlo, const float4 rhi, float4& ulo, float4& uhi)
{
    const int cnt_l = __float_as_int(llo.w), cnt_r = __float_as_int(rlo.w);
    ulo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), __int_as_float(cnt_l + cnt_r));
    uhi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.f);
    tr.left[node] = idl; tr.right[node] = idr; tr.cl[node] = cnt_l;
    tr.lo[node] = ulo; tr.hi[node] = uhi;
    if (idl < 0) tr.parent_leaf[~idl] = node; else tr.parent_node[idl] = node;
    if (idr < 0) tr.parent_leaf[~idr] = node; else tr.parent_node[idr] = node;
}

__global__ void __launch_bounds__(PL_THREADS)
k_ploc_merge(int n, int node_base, const int* __restrict__ cid, const float4* __restrict__ lo, const float4* __restrict__ hi,
             const int* __restrict__ nn, const unsigned char* __restrict__ role, const uint2* __restrict__ bscan,
             int* __restrict__ ocid, float4* __restrict__ olo, float4* __restrict__ ohi, PlocTree tr)
{
    __shared__ int s_w[PL_THREADS / 32][2];
    const int g = blockIdx.x * PL_THREADS + threadIdx.x;
    const int r = g < n ? role[g] : 0;
    int ex1, ex2, t1, t2;
    ploc_block_scan2(r == 1, r == 2, ex1, ex2, t1, t2, s_w);
    if (g >= n || r == 2) return;
    const uint2 base = bscan[blockIdx.x];
    const int p = g - (int)base.y - ex2;
    if (r == 1) {
        const int j = nn[g];
        float4 ulo, uhi;
        const int node = node_base + (int)base.x + ex1;
        ploc_make_node(tr, node, cid[g], cid[j], lo[g], hi[g], lo[j], hi[j], ulo, uhi);
        ocid[p] = node; olo[p] = ulo; ohi[p] = uhi;
    } else {
        ocid[p] = cid[g]; olo[p] = lo[g]; ohi[p] = hi[g];
    }
}

// the last <= PL_TAIL clusters: the whole loop in one block, clusters in shared memory
__global__ void __launch_bounds__(PL_TAIL)
k_ploc_tail(int n, int node_base, int R, const int* __restrict__ cid, const float4* __restrict__ lo, const float4* __restrict__ hi,
            PlocTree tr, int* __restrict__ h_iters)
{
    __shared__ float s_lo[3][PL_TAIL], s_hi[3][PL_TAIL];
    __shared__ int s_cnt[PL_TAIL], s_id[PL_TAIL], s_nn[PL_TAIL];
    __shared__ int s_w[PL_TAIL / 32][2];
    const int t = threadIdx.x;
    if (t < n) {
        const float4 a = lo[t], b = hi[t];
        s_lo[0][t] = a.x; s_lo[1][t] = a.y; s_lo[2][t] = a.z; s_cnt[t] = __float_as_int(a.w);
        s_hi[0][t] = b.x; s_hi[1][t] = b.y; s_hi[2][t] = b.z;
        s_id[t] = cid[t];
    }
    __syncthreads();
    int iters = 0;
    while (n > 1) {
        s_nn[t] = t < n ? ploc_nearest<PL_TAIL>(s_lo, s_hi, t, t, 0, n, R) : -1;
        __syncthreads();
        int r = 0, j = -1;
        if (t < n) {
            j = s_nn[t];
            if (j >= 0 && s_nn[j] == t) r = t < j ? 1 : 2;
        }
        int ex1, ex2, t1, t2;
        ploc_block_scan2(r == 1, r == 2, ex1, ex2, t1, t2, s_w);
        // read everything this thread needs before anyone overwrites the tile
        float4 mlo, mhi, plo, phi;
        int mid = 0, pid = 0;
        if (t < n && r != 2) {
            mlo = make_float4(s_lo[0][t], s_lo[1][t], s_lo[2][t], __int_as_float(s_cnt[t]));
            mhi = make_float4(s_hi[0][t], s_hi[1][t], s_hi[2][t], 0.f);
            mid = s_id[t];
            if (r == 1) {
                plo = make_float4(s_lo[0][j], s_lo[1][j], s_lo[2][j], __int_as_float(s_cnt[j]));
                phi = make_float4(s_hi[0][j], s_hi[1][j], s_hi[2][j], 0.f);
                pid = s_id[j];
            }
        }
        __syncthreads();
        if (t < n && r != 2) {
            const int p = t - ex2;
            if (r == 1) {
                float4 ulo, uhi;
                const int node = node_base + ex1;
                ploc_make_node(tr, node, mid, pid, mlo, mhi, plo, phi, ulo, uhi);
                mlo = ulo; mhi = uhi; mid = node;
            }
            s_lo[0][p] = mlo.x; s_lo[1][p] = mlo.y; s_lo[2][p] = mlo.z; s_cnt[p] = __float_as_int(mlo.w);
            s_hi[0][p] = mhi.x; s_hi[1][p] = mhi.y; s_hi[2][p] = mhi.z;
            s_id[p] = mid;
        }
        __syncthreads();
        n -= t2;
        node_base += t1;
        ++iters;
    }
    if (t == 0) h_iters[0] = iters;
}

// first depth-first leaf slot under every leaf and inner node + the tree height (longest leaf-to-root walk)
__global__ void k_ploc_positions(int T, PlocTree tr, BuildMeta* meta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * T - 1) return;
    const bool leaf = i < T;
    int cur = leaf ? ~i : i - T;
    int p = leaf ? tr.parent_leaf[i] : tr.parent_node[cur];
    int start = 0, depth = 0;
    while (p >= 0) {
        if (tr.right[p] == cur) start += tr.cl[p];
        cur = p;
        p = tr.parent_node[p];
        ++depth;
    }
    if (leaf) {
        tr.start_leaf[i] = start;
        int m = depth;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane_id() == 0) atomicMax(&meta->height, m);
    } else {
        tr.start_node[i - T] = start;
    }
}

// link of a child that is (or collapses into) a leaf; `inner` is set when the child keeps its own node record instead
__device__ __forceinline__ int ploc_link(const PlocTree& tr, int child, int leaf_max, const float4* leaf_lo, const float4* leaf_hi,
                                         float4& lo, float4& hi, bool& inner)
{
    inner = false;
    if (child < 0) {
        lo = leaf_lo[~child]; hi = leaf_hi[~child];
        return ~tr.start_leaf[~child];
    }
    lo = tr.lo[child]; hi = tr.hi[child];
    const int cnt = __float_as_int(lo.w);
    if (cnt <= leaf_max) return ~(tr.start_node[child] | ((cnt - 1) << 28));
    inner = true;                                            // the caller knows the record index (a leaf link may be -1 itself)
    return 0;
}

__global__ void k_ploc_emit(int T, PlocTree tr, const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi,
                            float4* __restrict__ nodes_out, int format, NodeQ nq, int leaf_max, BuildMeta* meta,
                            unsigned* __restrict__ live)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= T - 1) return;
    // Record numbering of the radix tree (Karras 2012), which keeps SIBLING records adjacent (one 128 B line): with
    // gamma = last depth-first leaf slot of the left subtree, the left child's record is gamma and the right child's gamma + 1;
    // the root is 0.  (A left inner child ends at gamma, a right inner child starts at gamma + 1, and no two inner nodes
    // share such an end / start slot, so the numbering is a bijection onto 0 .. T-2.)
    float4 l0, h0, l1, h1;
    const int gamma = tr.start_node[a] + tr.cl[a] - 1;
    bool in0, in1;
    int k0 = ploc_link(tr, tr.left[a], leaf_max, leaf_lo, leaf_hi, l0, h0, in0);
    int k1 = ploc_link(tr, tr.right[a], leaf_max, leaf_lo, leaf_hi, l1, h1, in1);
    if (in0) k0 = gamma;
    if (in1) k1 = gamma + 1;
    const int p = tr.parent_node[a];
    const int idx = p < 0 ? 0 : tr.start_node[p] + tr.cl[p] - 1 + (tr.right[p] == a ? 1 : 0);
    emit_node(nodes_out, idx, format, nq, l0, h0, l1, h1, k0, k1);
    live[idx] = (p < 0 || __float_as_int(tr.lo[a].w) > leaf_max) ? 1u : 0u;      // dead: collapsed into a leaf link of an ancestor
    if (p < 0) {
        const float4 lo = tr.lo[a], hi = tr.hi[a];
        meta->root_lo[0] = lo.x; meta->root_lo[1] = lo.y; meta->root_lo[2] = lo.z;
        meta->root_hi[0] = hi.x; meta->root_hi[1] = hi.y; meta->root_hi[2] = hi.z;
        meta->root = 0;
    }
}

// 48 B triangle records at the depth-first slot of each leaf (e1 = v1 - v0, e2 = v2 - v0: one IEEE subtraction each)
__global__ void k_tri_records(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t T,
                              const uint32_t* __restrict__ sorted_ids, const int* __restrict__ dst_slot, float4* __restrict__ tri_rec)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const uint32_t id = sorted_ids[i];
    float a[3], b[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = verts[3 * (int64_t)tris[3 * (int64_t)id + 0] + k];
        b[k] = verts[3 * (int64_t)tris[3 * (int64_t)id + 1] + k];
        c[k] = verts[3 * (int64_t)tris[3 * (int64_t)id + 2] + k];
    }
    const int64_t d = dst_slot[i];
    tri_rec[3 * d + 0] = make_float4(a[0], a[1], a[2], __uint_as_float(id));
    tri_rec[3 * d + 1] = make_float4(__fsub_rn(b[0], a[0]), __fsub_rn(b[1], a[1]), __fsub_rn(b[2], a[2]), 0.f);
    tri_rec[3 * d + 2] = make_float4(__fsub_rn(c[0], a[0]), __fsub_rn(c[1], a[1]), __fsub_rn(c[2], a[2]), 0.f);
}

// Host driver.  Cluster buffers c*[0] hold the leaves on entry.  h_pin: >= 4 ints of page-locked memory.
int ploc_build(lrc_ctx* ctx, int T, int R, int* cid[2], float4* clo[2], float4* chi[2], int* nn, unsigned char* role, uint2* bsum,
               PlocTree tr, int* h_pin, int* iterations, cudaStream_t stream)
{
    int n = T, node_base = 0, cur = 0, iters = 0;
    while (n > PL_TAIL) {
        const int nb = (n + PL_THREADS - 1) / PL_THREADS;
        k_ploc_search<<<nb, PL_THREADS, 0, stream>>>(n, R, clo[cur], chi[cur], nn, role, bsum);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_search");
        k_ploc_scan<<<1, 1024, 0, stream>>>(bsum, nb, h_pin);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_scan");
        k_ploc_merge<<<nb, PL_THREADS, 0, stream>>>(n, node_base, cid[cur], clo[cur], chi[cur], nn, role, bsum, cid[cur ^ 1],
                                                    clo[cur ^ 1], chi[cur ^ 1], tr);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_merge");
        LRC_CUDA(ctx, cudaStreamSynchronize(stream));
        const int merged = h_pin[0];
        if (merged <= 0 || merged != h_pin[1]) return lrc_fail(ctx, LRC_ERR_CUDA, "PLOC build: inconsistent merge counts");
        n -= merged;
        node_base += merged;
        cur ^= 1;
        if (++iters > 100000) return lrc_fail(ctx, LRC_ERR_CAPACITY, "PLOC build: no convergence (degenerate mesh?)");
    }
    if (n > 1) {
        h_pin[2] = 0;
        k_ploc_tail<<<1, PL_TAIL, 0, stream>>>(n, node_base, R, cid[cur], clo[cur], chi[cur], tr, h_pin + 2);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_tail");
        LRC_CUDA(ctx, cudaStreamSynchronize(stream));
        iters += h_pin[2];
    }
    *iterations = iters;
    return LRC_OK;
}

}  // namespace
