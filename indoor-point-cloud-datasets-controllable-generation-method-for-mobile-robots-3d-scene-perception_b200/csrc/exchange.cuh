// exchange.cuh -- the multi-GPU exchange kernels (included by scan.cu inside its anonymous namespace, after the ray
// generator and the compaction): k_push (vector loads / stores), k_push_tma (TMA bulk copies global -> shared -> every peer),
// and the compact wire format (progress words with release / acquire at system scope, rebuild of the other ranks' points).
// The host side that launches them is run_scan in scan.cu; the public entry points are lrc_peer_buffer_*, lrc_set_gather and
// lrc_set_gather_wire (include/lrc.h).
#pragma once

// ---- multi-GPU exchange: push a finished chunk of the compacted cloud into every rank's gather buffer ----------------
// Peer targets (lrc_set_gather): buffers in local HBM or in another GPU's HBM mapped over NVLink (cudaIpcOpenMemHandle).
// After chunk c has been compacted locally, k_push copies its slice [run[c], run[c+1]) of xyz and label -- and the frame
// offsets of the chunk's frames, rebased -- to point_base + position in EVERY target with 16-byte vector loads and
// stores, so each warp store is 512 contiguous bytes on the wire.  (The first version stored 4 bytes at a time from inside
// the compaction kernel: 32 remote stores per point at 8 GPUs, 0.8 ms per chunk; see profiles/.)  The slice bounds are
// read from device memory, so nothing returns to the host; the kernel runs on the auxiliary stream while the next chunk
// is traversed on the caller's stream.
struct GatherTargets {
    int n;
    float* xyz[LRC_MAX_GATHER];
    uint32_t* label[LRC_MAX_GATHER];
    int64_t* frame_offset[LRC_MAX_GATHER];
    int64_t point_base, frame_base, capacity;
    int wire, self;                        // compact wire format: t | label | ray index travel, xyz is rebuilt on arrival
    float* wire_t[LRC_MAX_GATHER];
    uint32_t* wire_ray[LRC_MAX_GATHER];
    int64_t* ready[LRC_MAX_GATHER];
};

struct PushParams {
    const float* xyz;              // local compacted outputs of this call (lrc_out)
    const uint32_t* label;
    const int64_t* frame_offset;
    const long long* run;          // run[0], run[1]: first and one-past-last point of the chunk
    int64_t f0, nf, P;             // frames of the chunk; P = frames of the whole call
    int last;                      // the chunk that owns the closing frame offset
};

// copy n 4-byte words src -> dst; both pointers 4-byte aligned and congruent modulo 16 (else word by word)
__device__ __forceinline__ void push_words(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n, int64_t tid, int64_t nthreads)
{
    if ((((uintptr_t)src) & 15u) != (((uintptr_t)dst) & 15u)) {
        for (int64_t i = tid; i < n; i += nthreads) dst[i] = __ldcg(src + i);
        return;
    }
    int64_t head = (int64_t)(((16u - (((uintptr_t)src) & 15u)) & 15u) >> 2);
    if (head > n) head = n;
    const int64_t nvec = (n - head) >> 2;
    const int64_t tail0 = head + (nvec << 2);
    if (tid < head) dst[tid] = __ldcg(src + tid);
    if (tid < n - tail0) dst[tail0 + tid] = __ldcg(src + tail0 + tid);
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    // the exchange is bound by bytes in flight towards each target (NVLink round trip ~ microseconds), not by SM count:
    // eight 16-byte vectors per thread are loaded before the first store is issued
    int64_t i = tid;
    for (; i + 7 * nthreads < nvec; i += 8 * nthreads) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcg(s4 + i + k * nthreads);
#pragma unroll
        for (int k = 0; k < 8; ++k) __stcs(d4 + i + k * nthreads, v[k]);
    }
    for (; i < nvec; i += nthreads) __stcs(d4 + i, __ldcg(s4 + i));
}

constexpr int PUSH_THREADS = 256;
__global__ void __launch_bounds__(PUSH_THREADS) k_push(PushParams q, const __grid_constant__ GatherTargets gt)
{
    const int k = blockIdx.y;                       // target
    const long long a = q.run[0], b = q.run[1];
    const int64_t tid = (int64_t)blockIdx.x * PUSH_THREADS + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * PUSH_THREADS;
    const long long gp = gt.point_base + a;
    // a scan whose outputs already ARE this rank's region of its own gather buffer (PeerGather.local_out) has nothing to
    // copy to itself: one target less per step, and no second pass over the local cloud
    if (gt.xyz[k] + 3 * gp != q.xyz + 3 * a)
        push_words(reinterpret_cast<const uint32_t*>(q.xyz + 3 * a), reinterpret_cast<uint32_t*>(gt.xyz[k] + 3 * gp), 3 * (b - a), tid, nthreads);
    if (q.label) {
        if (gt.label[k] + gp != q.label + a) push_words(q.label + a, gt.label[k] + gp, b - a, tid, nthreads);
    } else for (int64_t i = tid; i < b - a; i += nthreads) gt.label[k][gp + i] = 0u;
    for (int64_t f = tid; f < q.nf; f += nthreads) gt.frame_offset[k][gt.frame_base + q.f0 + f] = gt.point_base + q.frame_offset[q.f0 + f];
    if (q.last && tid == 0) gt.frame_offset[k][gt.frame_base + q.P] = gt.point_base + b;
}

// ---- the same exchange through the TMA engines (option "push_mode" = 1) ------------------------------------------------------
// k_push moves every byte through registers: loads and stores of 64 x N blocks compete with k_trace for exactly the
// resource that bounds it (the L1 / LSU data path; measured: k_trace 0.39 -> 0.53 ms per chunk at 4 GPUs while the
// exchange runs).  Here ONE thread per block drives 1-D bulk copies: a 16 KB tile of the chunk's compacted stream goes
// global -> shared (cp.async.bulk ... mbarrier::complete_tx) and from shared to EVERY target (cp.async.bulk.global.shared::cta,
// peer memory over NVLink included) -- the tile is read from HBM once instead of once per target, no data touches a
// register, and a handful of blocks keeps megabytes in flight.  Streams: the chunk's xyz bytes and label bytes; their
// 16-byte-aligned middle goes through TMA, the (at most 15-byte) head and tail and the frame offsets through plain stores.
constexpr int TMA_TILE_MAX = 16384;      // bytes per stage (option "push_tile": 2048 ... 16384; shared memory per block = 4 stages)
constexpr int TMA_STAGES = 4;
constexpr int TMA_THREADS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D;\n"
        "bra W;\n"
        "D:\n"
        "}\n" :: "r"(bar), "r"(parity) : "memory");
}

struct TmaStream {
    const char* src;        // first byte of the chunk's slice in the local compacted output
    int64_t dst_off;        // byte offset of that slice inside every target's array
    int64_t bytes;
    int kind;               // 0 = xyz, 1 = label
};

__global__ void __launch_bounds__(TMA_THREADS) k_push_tma(PushParams q, const __grid_constant__ GatherTargets gt, int tile)
{
    extern __shared__ __align__(128) unsigned char tma_smem[];
    __shared__ __align__(8) unsigned long long bars[TMA_STAGES];
    const long long a = q.run[0], b = q.run[1];
    const long long gp = gt.point_base + a;
    // streams of this chunk: xyz | label, or -- compact wire format -- t | label | ray index (xyz is rebuilt by the receiver)
    TmaStream st[3];
    int n_streams = 2;
    if (gt.wire) {
        n_streams = 3;
        st[0].src = reinterpret_cast<const char*>(gt.wire_t[gt.self] + gp); st[0].dst_off = 4 * gp; st[0].bytes = 4 * (b - a); st[0].kind = 2;
        st[2].src = reinterpret_cast<const char*>(gt.wire_ray[gt.self] + gp); st[2].dst_off = 4 * gp; st[2].bytes = 4 * (b - a); st[2].kind = 3;
    } else {
        st[0].src = reinterpret_cast<const char*>(q.xyz + 3 * a); st[0].dst_off = 12 * gp; st[0].bytes = 12 * (b - a); st[0].kind = 0;
        st[2].src = nullptr; st[2].dst_off = 0; st[2].bytes = 0; st[2].kind = 0;
    }
    st[1].src = reinterpret_cast<const char*>(q.label + a);   st[1].dst_off = 4 * gp;  st[1].bytes = q.label ? 4 * (b - a) : 0; st[1].kind = 1;
    auto target_base = [&](int k, int kind) -> char* {
        return kind == 0 ? reinterpret_cast<char*>(gt.xyz[k]) : kind == 1 ? reinterpret_cast<char*>(gt.label[k])
             : kind == 2 ? reinterpret_cast<char*>(gt.wire_t[k]) : reinterpret_cast<char*>(gt.wire_ray[k]);
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < TMA_STAGES; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    int64_t g0 = 0;       // tiles this block has pushed in earlier streams: stage and barrier parity continue from there
    for (int sidx = 0; sidx < n_streams; ++sidx) {
        const TmaStream& S = st[sidx];
        if (S.bytes <= 0) continue;
        // aligned middle [head, head + mid): source and destination are congruent modulo 16 when the rank's point base keeps
        // the alignment (capacity a multiple of 4 points, which PeerGather guarantees); otherwise everything goes the slow way
        int64_t head = (int64_t)((16u - (unsigned)((uintptr_t)S.src & 15u)) & 15u);
        if (head > S.bytes) head = S.bytes;
        bool congruent = true;
        for (int k = 0; k < gt.n; ++k)
            congruent = congruent && ((((uintptr_t)(target_base(k, S.kind) + S.dst_off)) & 15u) == (((uintptr_t)S.src) & 15u));
        const int64_t mid = congruent ? ((S.bytes - head) & ~(int64_t)15) : 0;
        if (!congruent) head = 0;
        const int64_t n_tiles = (mid + tile - 1) / tile;
        // tiles of this block: blockIdx.x, blockIdx.x + gridDim.x, ...
        const int64_t mine = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        if (threadIdx.x == 0 && mine > 0) {
            auto tile_bytes = [&](int64_t i) -> uint32_t {
                const int64_t t = blockIdx.x + i * gridDim.x;
                const int64_t left = mid - t * tile;
                return (uint32_t)(left < tile ? left : tile);
            };
            auto issue_load = [&](int64_t i) {
                const int stage = (int)((g0 + i) % TMA_STAGES);
                const int64_t t = blockIdx.x + i * gridDim.x;
                const uint32_t nb = tile_bytes(i);
                const uint32_t bar = smem_u32(&bars[stage]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(nb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(tma_smem + stage * tile)), "l"(S.src + head + t * tile), "r"(nb), "r"(bar) : "memory");
            };
            static_assert(TMA_STAGES >= 3, "the pipeline keeps TMA_STAGES - 2 loads ahead");
            const int ahead = TMA_STAGES - 2;
            for (int64_t i = 0; i < ahead && i < mine; ++i) issue_load(i);
            for (int64_t i = 0; i < mine; ++i) {
                if (i + ahead < mine) {
                    // the stage of tile i + ahead was last read by the stores of tile i + ahead - STAGES = i - 2: at most the most
                    // recent store group may still be reading shared memory
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    issue_load(i + ahead);
                }
                const int stage = (int)((g0 + i) % TMA_STAGES);
                mbar_wait(smem_u32(&bars[stage]), (unsigned)(((g0 + i) / TMA_STAGES) & 1));
                const int64_t t = blockIdx.x + i * gridDim.x;
                const uint32_t nb = tile_bytes(i);
                for (int k = 0; k < gt.n; ++k) {
                    char* dst = target_base(k, S.kind) + S.dst_off + head + t * tile;
                    if (dst == S.src + head + t * tile) continue;       // this rank's own region already holds the data
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 :: "l"(dst), "r"(smem_u32(tma_smem + stage * tile)), "r"(nb) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        // head, tail (and everything, when not congruent): plain 4-byte stores by block 0
        if (blockIdx.x == 0) {
            const int64_t tail0 = head + mid;
            const int64_t n_words = (head + (S.bytes - tail0)) >> 2;
            for (int64_t i = threadIdx.x; i < n_words * gt.n; i += TMA_THREADS) {
                const int k = (int)(i / n_words);
                const int64_t wd = i - (int64_t)k * n_words;
                const int64_t off = wd < (head >> 2) ? 4 * wd : tail0 + 4 * (wd - (head >> 2));
                char* dst = target_base(k, S.kind) + S.dst_off + off;
                if (dst != S.src + off) *reinterpret_cast<uint32_t*>(dst) = *reinterpret_cast<const uint32_t*>(S.src + off);
            }
        }
        g0 += mine;
        __syncthreads();      // the next stream reuses the stages and their barriers
    }
    if (!q.label && blockIdx.x == 0)
        for (int64_t i = threadIdx.x; i < (b - a) * gt.n; i += TMA_THREADS) gt.label[i / (b - a)][gp + i % (b - a)] = 0u;
    if (blockIdx.x == gridDim.x - 1) {
        for (int64_t i = threadIdx.x; i < q.nf * gt.n; i += TMA_THREADS) {
            const int k = (int)(i / q.nf);
            const int64_t f = i - (int64_t)k * q.nf;
            gt.frame_offset[k][gt.frame_base + q.f0 + f] = gt.point_base + q.frame_offset[q.f0 + f];
        }
        // the offset that closes this chunk (the next chunk writes the same value; the receiver of the compact wire format
        // needs it to know where the chunk's points end)
        if (threadIdx.x < gt.n) gt.frame_offset[threadIdx.x][gt.frame_base + q.f0 + q.nf] = gt.point_base + b;
    }
    // make the bulk stores of this block globally visible before the kernel (and with it the stream-ordered event) completes
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- compact wire format: progress words and the rebuild of the other ranks' points --------------------------------------
// Producer side: after the bulk copies of a chunk have completed (k_push_tma ends with cp.async.bulk.wait_group 0 and the
// kernel boundary orders everything before this launch), one store with release semantics at system scope tells every
// target how many frames of this scan are complete: word = scan number << 32 | frames.
__global__ void k_wire_flag(const __grid_constant__ GatherTargets gt, long long word)
{
    const int k = threadIdx.x;
    if (k < gt.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(gt.ready[k] + gt.self), "l"(word) : "memory");
    }
}

struct WireWait {
    const int64_t* ready;                 // this rank's own progress words, one per rank
    long long need[LRC_MAX_GATHER];       // value to wait for, per rank (0: nothing to wait for)
    int n;
};

// Receiver side, one thread per peer: spin (acquire, system scope) until the peer's progress word has reached `need`.
// ONE resident block, so waiting cannot starve the traversal.  A peer that never arrives (ranks out of step) ends the
// wait after ~4 s and raises the error flag instead of hanging the GPU.
__global__ void k_wire_wait(WireWait w, int* err)
{
    const int p = threadIdx.x;
    if (p >= w.n || w.need[p] == 0) return;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w.ready + p) : "memory");
        if (v >= w.need[p]) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) { atomicExch(err, 1); break; }
        __nanosleep(200);
    }
}

struct RebuildParams {
    float* xyz;                           // this rank's own gather arrays (all ranks' regions)
    const float* wire_t;
    const uint32_t* wire_ray;
    const int64_t* frame_offset;
    int n, self;
    int64_t point_base[LRC_MAX_GATHER], frame_base[LRC_MAX_GATHER];
    int64_t pose0[LRC_MAX_GATHER];        // first pose of every rank in g.poses
    int64_t fa[LRC_MAX_GATHER], fb[LRC_MAX_GATHER];      // frames of every rank rebuilt by this launch
    uint64_t pose_index_base;             // noise: global base (rank r's frame f draws stream base + pose0[r] + f)
};

// p = o + (d / |d|) * t for every point of the peers' frames [fa, fb): the ray is regenerated exactly as the producing rank
// generated it (gen_ray: float64 tables / Philox, one rounding to float32), the point with frame_epilogue's operations.
// blockIdx.y = rank; its points are found through its frame offsets (acquired by k_wire_wait, read around L1).
template <int MODE>
__global__ void __launch_bounds__(256) k_wire_rebuild(RayGen g, RebuildParams q, int sub_blocks)
{
    const int p = blockIdx.y;
    if (p == q.self) return;
    // one frame per group of `sub_blocks` blocks: the pose is loop-invariant, no search for the frame of a point
    const int64_t f = q.fa[p] + (int64_t)(blockIdx.x / sub_blocks);
    if (f >= q.fb[p]) return;
    const int sub = blockIdx.x % sub_blocks;
    const int64_t* off = q.frame_offset + q.frame_base[p];
    const int64_t i0 = __ldcg(off + f), i1 = __ldcg(off + f + 1);       // absolute point slots (carry the rank's base)
    g.pose0 = q.pose0[p];
    g.pose_index_base = q.pose_index_base;
    for (int64_t i = i0 + (int64_t)sub * blockDim.x + threadIdx.x; i < i1; i += (int64_t)sub_blocks * blockDim.x) {
        const int r = (int)__ldcg(q.wire_ray + i);
        const float t = __ldcg(q.wire_t + i);
        const Ray ray = MODE == MODE_SINGLE ? gen_single(g, f, r) : gen_dual(g, f, r);
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(ray.dx, ray.dx), __fmul_rn(ray.dy, ray.dy)), __fmul_rn(ray.dz, ray.dz)));
        __stcs(q.xyz + 3 * i + 0, __fadd_rn(ray.ox, __fmul_rn(__fdiv_rn(ray.dx, nrm), t)));
        __stcs(q.xyz + 3 * i + 1, __fadd_rn(ray.oy, __fmul_rn(__fdiv_rn(ray.dy, nrm), t)));
        __stcs(q.xyz + 3 * i + 2, __fadd_rn(ray.oz, __fmul_rn(__fdiv_rn(ray.dz, nrm), t)));
    }
}

