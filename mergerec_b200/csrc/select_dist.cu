// select_dist.cu -- magnitude histograms for the TIES trim over a flat vector that is SHARDED across GPUs
// ("shard the flat dimension d/G ... ncclAllReduce of the (K x bins) histograms per radix pass").
//
// reference semantics: `torch.topk(update.abs(), k)` over the WHOLE flat vector (merger/algorithms/ties.py:14-23).
// With the vector split over ranks the k-th largest magnitude is found by three radix passes over the 31 magnitude
// bits (11 + 10 + 10): every rank histograms its slice (this kernel), the (K x 2048) int64 histograms are summed with
// one all-reduce per pass, and every rank walks the global histogram from the top to the bin holding rank k.  The
// host side (mergerec_b200/merger/sharded.py) turns the global cut into a per-rank cut key for mr_ties_build.
#include "common.cuh"

namespace mr {

constexpr int kMhThreads = 256;
constexpr int kMhBins = 2048;

template <int K>
__global__ void __launch_bounds__(kMhThreads)
mag_hist_kernel(const float* __restrict__ base, PtrPack<K> m, int64_t d, const float* __restrict__ w, int pass,
                const uint32_t* __restrict__ prefix, unsigned long long* __restrict__ hist) {
    extern __shared__ uint32_t s_hist[];  // K * kMhBins
    __shared__ uint32_t s_prefix[K];
    __shared__ float s_w[K];
    for (int i = threadIdx.x; i < K * kMhBins; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < K) {
        s_prefix[threadIdx.x] = pass ? prefix[threadIdx.x] : 0;
        s_w[threadIdx.x] = w ? w[threadIdx.x] : 1.0f;
    }
    __syncthreads();
    const bool weighted = (w != nullptr);
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    const int64_t rounds = (d + span - 1) / span;  // every lane runs every round (warp-wide match below)
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t j = it * span + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        const bool live = j < d;
        const float b = live ? ldg_stream1(base + j) : 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float u = __fsub_rn(live ? ldg_stream1(m.p[k] + j) : 0.f, b);       // ties.py:18
            if (weighted) u = __fmul_rn(u, s_w[k]);                               // ties.py:20 (`update *= w`)
            const uint32_t bits = __float_as_uint(u) & 0x7fffffffu;
            int bin = -1;
            if (live) {
                if (pass == 0) bin = (int)(bits >> 20);
                else if (pass == 1) { if ((bits >> 20) == s_prefix[k]) bin = (int)((bits >> 10) & 1023u); }
                else { if ((bits >> 10) == s_prefix[k]) bin = (int)(bits & 1023u); }
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, bin);
            if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[k * kMhBins + bin], __popc(peers));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * kMhBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

}  // namespace mr

extern "C" int mr_ties_mag_hist(const float* base, const float* const* models, int K, int64_t d, const float* w, int pass,
                                const uint32_t* prefix, int64_t* hist, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_ties_mag_hist: K=%d outside [1,%d]", K, MR_MAX_K);
    MR_REQUIRE(d >= 0 && d < ((int64_t)1 << 32), "mr_ties_mag_hist: need 0 <= d < 2^32");
    MR_REQUIRE(pass >= 0 && pass <= 2, "mr_ties_mag_hist: pass %d outside [0,2]", pass);
    MR_REQUIRE(hist && (pass == 0 || prefix), "mr_ties_mag_hist: null pointer");
    if (d == 0) return MR_OK;
    MR_REQUIRE(base && models, "mr_ties_mag_hist: null pointer");
    for (int k = 0; k < K; ++k) MR_REQUIRE(models[k] != nullptr, "mr_ties_mag_hist: models[%d] is NULL", k);
    int64_t blocks = (d + kMhThreads - 1) / kMhThreads;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    const size_t smem = (size_t)K * kMhBins * sizeof(uint32_t);
    MR_DISPATCH_K(K, {
        PtrPack<KK> pk;
        for (int k = 0; k < KK; ++k) pk.p[k] = models[k];
        cudaFuncSetAttribute(mag_hist_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        mag_hist_kernel<KK><<<(unsigned)blocks, kMhThreads, smem, (cudaStream_t)stream>>>(
            base, pk, d, w, pass, prefix, reinterpret_cast<unsigned long long*>(hist));
    });
    MR_CUDA_LAUNCH_CHECK("mr_ties_mag_hist");
    return MR_OK;
}
