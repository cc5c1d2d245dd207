// scan_util.cuh -- a small single-block exclusive scan shared by the bin-sort style index builders (plan.cu, nn.cu).
#pragma once
#include "common.cuh"

namespace {

// exclusive scan of n ints by one block; out[n] = total.  `cursor` (optional) receives a copy of the offsets.
__global__ void __launch_bounds__(1024) k_scan_int(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ cursor, int64_t n)
{
    __shared__ int warp_sums[32];
    __shared__ int carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b0 = 0; b0 < n; b0 += 1024) {
        const int64_t i = b0 + threadIdx.x;
        const int v = i < n ? in[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0) {
            int sv = warp_sums[lane], si = sv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, si, o);
                if (lane >= o) si += t;
            }
            warp_sums[lane] = si - sv;
        }
        __syncthreads();
        const int excl = incl - v + warp_sums[w] + carry;
        if (i < n) { out[i] = excl; if (cursor) cursor[i] = excl; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

}  // namespace
