// dare.cu -- DARE merge (drop-and-rescale) given explicit keep masks.
//
// reference: rec_retrieval/merger/algorithms/dare.py:9-31
//     merged = base.clone(); for i: update = weights[i] * (m_i - base); update = dropout(update, p=density); merged += update
// with torch's CPU dropout = `input * (bernoulli(1 - p) / (1 - p))`.  The random keep mask is an INPUT here (the Python
// mirror draws it with torch's generator; tests replay the reference's own masks), so the arithmetic is bit-exact:
//     merged[j] = base[j] (+)_{k in order} fl( fl(w_k * fl(m_k[j] - base[j])) * (keep[k, j] ? scale : 0) )
// One streaming pass, K + 1 fp32 reads + K mask bytes + one write per column; HBM-bound.
#include "common.cuh"

namespace mr {

constexpr int kDareThreads = 256;

template <int K>
__global__ void __launch_bounds__(kDareThreads)
dare_kernel(const float* __restrict__ base, PtrPack<K> m, int64_t d, const float* __restrict__ w,
            const uint8_t* __restrict__ keep, int64_t ld_keep, float scale, float* __restrict__ out) {
    float wk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = w[k];
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < d; j += span) {
        const float b = ldg_stream1(base + j);
        float x[K];
        uint8_t kp[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            x[k] = ldg_stream1(m.p[k] + j);
            kp[k] = keep[(int64_t)k * ld_keep + j];
        }
        float acc = b;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float u = __fmul_rn(wk[k], __fsub_rn(x[k], b));
            acc = __fadd_rn(acc, __fmul_rn(u, kp[k] ? scale : 0.0f));
        }
        out[j] = acc;
    }
}

}  // namespace mr

extern "C" int mr_merge_dare(const float* base, const float* const* models, int K, int64_t d, const float* w,
                             const uint8_t* keep, int64_t ld_keep, float scale, float* out, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_merge_dare: K=%d outside [1,%d]", K, MR_MAX_K);
    MR_REQUIRE(d >= 0, "mr_merge_dare: need d >= 0");
    if (d == 0) return MR_OK;
    MR_REQUIRE(base && models && w && keep && out, "mr_merge_dare: null pointer");
    MR_REQUIRE(ld_keep >= d, "mr_merge_dare: need ld_keep >= d");
    for (int k = 0; k < K; ++k) MR_REQUIRE(models[k] != nullptr, "mr_merge_dare: models[%d] is NULL", k);
    int64_t blocks = (d + kDareThreads - 1) / kDareThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    MR_DISPATCH_K(K, {
        PtrPack<KK> pk;
        for (int k = 0; k < KK; ++k) pk.p[k] = models[k];
        dare_kernel<KK><<<(unsigned)blocks, kDareThreads, 0, (cudaStream_t)stream>>>(base, pk, d, w, keep, ld_keep, scale, out);
    });
    MR_CUDA_LAUNCH_CHECK("mr_merge_dare");
    return MR_OK;
}
