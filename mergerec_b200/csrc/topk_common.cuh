// topk_common.cuh -- ordering keys and the shared-memory sort used by every top-K kernel.
#pragma once
#include "common.cuh"

namespace mr {

typedef unsigned long long u64;

// Monotone map float -> uint32 (larger score = larger key); -0.0 == +0.0; NaN is the greatest (torch.topk).
__device__ __forceinline__ uint32_t score_key(float f) {
    if (f != f) return 0xFFFFFFFFu;
    const uint32_t u = __float_as_uint(__fadd_rn(f, 0.0f));  // -0 -> +0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// (score desc, id asc) as one descending 64-bit order.  0 is never a real key (empty slot / padding).
__device__ __forceinline__ u64 topk_key(float s, uint32_t id) {
    return ((u64)score_key(s) << 32) | (u64)(0xFFFFFFFFu - id);
}
__device__ __forceinline__ float key_score(u64 key) {
    const uint32_t u = (uint32_t)(key >> 32);
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__device__ __forceinline__ uint32_t key_id(u64 key) { return 0xFFFFFFFFu - (uint32_t)key; }

// Bitonic sort of n (power of two) keys in shared memory, descending, by the whole CTA.
// Every thread of the block must call it; ends with a barrier.
__device__ __forceinline__ void block_bitonic_sort_desc(u64* s, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
                const int a = 2 * i - (i & (stride - 1));
                const int b = a + stride;
                const bool desc = ((a & size) == 0);
                const u64 x = s[a], y = s[b];
                if ((x < y) == desc) { s[a] = y; s[b] = x; }
            }
            __syncthreads();
        }
    }
}

}  // namespace mr
