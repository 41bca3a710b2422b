// select_dist.cu -- magnitude histograms for the TIES trim over a flat vector that is SHARDED across GPUs
// ("shard the flat dimension d/G ... ncclAllReduce of the (K x bins) histograms per radix pass").
//
// reference semantics: `torch.topk(update.abs(), k)` over the WHOLE flat vector (merger/algorithms/ties.py:14-23).
// With the vector split over ranks the k-th largest magnitude is found by radix refinement of a WINDOW of magnitude bit
// patterns: every rank histograms its slice (this kernel), the (K x 2048) int64 histograms (+ the counts above the
// window) are summed with one all-reduce per level, and every rank walks the global histogram from the top to the bin
// holding rank k.  The first window comes from per-rank order statistics (mr_ties_select on each slice), so two levels
// usually suffice and almost no element touches the shared-memory histogram.  The host side
// (mergerec_b200/merger/sharded.py) turns the global cut into a per-rank cut key for mr_ties_build.
#include "common.cuh"

namespace mr {

constexpr int kMhThreads = 256;
constexpr int kMhBins = 2048;

// Windowed histogram of the magnitude bit patterns: model k looks at the window [lo_k, lo_k + 2048 << s_k); an element
// inside it adds one to bin (bits - lo_k) >> s_k, an element beyond it to above[k], one below it is ignored.  The host
// chooses the first window from per-rank order statistics (a few per cent wide, so the shared-memory atomics are rare
// and the pass runs at HBM speed) and refines it 11 bits at a time; lo = 0, s = 20 is the always-correct full range.
struct MhCand {           // optional: (bin, local index) of every in-window element of the models whose shift is 0
    uint32_t* list;      // K x cap x 2
    uint32_t* count;     // K (may exceed cap: the caller then falls back)
    uint32_t cap;
};

template <int K, bool COUNT_ABOVE>
__device__ __forceinline__ void mh_one(float b, const float (&x)[K], const float* s_w, bool weighted, const uint32_t* s_lo,
                                       const int* s_sh, uint32_t* s_hist, uint32_t (&above)[K], const MhCand& cand,
                                       int64_t j) {
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float u = __fsub_rn(x[k], b);                 // ties.py:18
        if (weighted) u = __fmul_rn(u, s_w[k]);       // ties.py:20 (`update *= w`)
        const uint32_t bits = __float_as_uint(u) & 0x7fffffffu;
        if (bits >= s_lo[k]) {
            const uint32_t bin = (bits - s_lo[k]) >> s_sh[k];
            if (bin < (uint32_t)kMhBins) {
                atomicAdd(&s_hist[k * kMhBins + bin], 1u);
                if (cand.list && s_sh[k] == 0) {
                    const uint32_t pos = atomicAdd(&cand.count[k], 1u);
                    if (pos < cand.cap) {
                        cand.list[((size_t)k * cand.cap + pos) * 2] = bin;
                        cand.list[((size_t)k * cand.cap + pos) * 2 + 1] = (uint32_t)j;
                    }
                }
            } else if (COUNT_ABOVE) {
                ++above[k];
            }
        }
    }
}

template <int K, bool VEC>
__global__ void __launch_bounds__(kMhThreads)
mag_hist_kernel(const float* __restrict__ base, PtrPack<K> m, int64_t d, const float* __restrict__ w,
                const uint32_t* __restrict__ lo, const int32_t* __restrict__ shift, unsigned long long* __restrict__ hist,
                unsigned long long* __restrict__ above_out, MhCand cand) {
    extern __shared__ uint32_t s_hist[];  // K * kMhBins
    __shared__ uint32_t s_lo[K];
    __shared__ int s_sh[K];
    __shared__ float s_w[K];
    for (int i = threadIdx.x; i < K * kMhBins; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < K) {
        s_lo[threadIdx.x] = lo[threadIdx.x];
        s_sh[threadIdx.x] = shift[threadIdx.x];
        s_w[threadIdx.x] = w ? w[threadIdx.x] : 1.0f;
    }
    __syncthreads();
    const bool weighted = (w != nullptr);
    uint32_t above[K];
#pragma unroll
    for (int k = 0; k < K; ++k) above[k] = 0;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    if (VEC) {
        // hot loop: branch-free classification of the 4 x K elements of a quad (count what lies above the window, note
        // whether anything lies inside); the rare quad with an in-window element is redone by the scalar routine
        uint32_t lo_r[K];
        int sh_r[K];
        float w_r[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { lo_r[k] = s_lo[k]; sh_r[k] = s_sh[k]; w_r[k] = s_w[k]; }
        const int64_t n4 = d >> 2;
        for (int64_t v = gtid; v < n4; v += gsz) {
            const float4 xb = ldg_stream4(base + 4 * v);
            float4 xs[K];
#pragma unroll
            for (int k = 0; k < K; ++k) xs[k] = ldg_stream4(m.p[k] + 4 * v);
            const float bx[4] = {xb.x, xb.y, xb.z, xb.w};
            unsigned long long hit = 0;   // bit 4k + c: element c of model k lies inside its window
#pragma unroll
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float u = __fsub_rn(reinterpret_cast<const float*>(&xs[k])[c], bx[c]);
                    if (weighted) u = __fmul_rn(u, w_r[k]);
                    const uint32_t bits = __float_as_uint(u) & 0x7fffffffu;
                    const bool ge = bits >= lo_r[k];
                    const bool inw = ge && (((bits - lo_r[k]) >> sh_r[k]) < (uint32_t)kMhBins);
                    above[k] += (ge && !inw) ? 1u : 0u;
                    hit |= (unsigned long long)(inw ? 1u : 0u) << (4 * k + c);
                }
            }
            if (hit) {   // work proportional to the number of in-window elements, static register indexing
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint32_t mk = (uint32_t)(hit >> (4 * k)) & 0xFu;
                    if (mk) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if ((mk >> c) & 1u) {
                                float u = __fsub_rn(reinterpret_cast<const float*>(&xs[k])[c], bx[c]);
                                if (weighted) u = __fmul_rn(u, w_r[k]);
                                const uint32_t bin = ((__float_as_uint(u) & 0x7fffffffu) - lo_r[k]) >> sh_r[k];
                                atomicAdd(&s_hist[k * kMhBins + bin], 1u);
                                if (cand.list && sh_r[k] == 0) {
                                    const uint32_t pos = atomicAdd(&cand.count[k], 1u);
                                    if (pos < cand.cap) {
                                        cand.list[((size_t)k * cand.cap + pos) * 2] = bin;
                                        cand.list[((size_t)k * cand.cap + pos) * 2 + 1] = (uint32_t)(4 * v + c);
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    for (int64_t j = (VEC ? (d & ~(int64_t)3) : 0) + gtid; j < d; j += gsz) {
        float x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = m.p[k][j];
        mh_one<K, true>(base[j], x, s_w, weighted, s_lo, s_sh, s_hist, above, cand, j);
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t v = above[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&above_out[k], (unsigned long long)v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * kMhBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], (unsigned long long)s_hist[i]);
}

}  // namespace mr

extern "C" int mr_ties_mag_hist(const float* base, const float* const* models, int K, int64_t d, const float* w,
                                const uint32_t* lo, const int32_t* shift, int64_t* hist, int64_t* above,
                                uint32_t* cand, uint32_t* cand_count, int cand_cap, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_ties_mag_hist: K=%d outside [1,%d]", K, MR_MAX_K);
    MR_REQUIRE(d >= 0 && d < ((int64_t)1 << 32), "mr_ties_mag_hist: need 0 <= d < 2^32");
    MR_REQUIRE(lo && shift && hist && above, "mr_ties_mag_hist: null pointer");
    MR_REQUIRE(!cand || (cand_count && cand_cap > 0), "mr_ties_mag_hist: candidate list needs a counter and a capacity");
    if (d == 0) return MR_OK;
    MR_REQUIRE(base && models, "mr_ties_mag_hist: null pointer");
    bool vec = host_aligned16(base);
    for (int k = 0; k < K; ++k) {
        MR_REQUIRE(models[k] != nullptr, "mr_ties_mag_hist: models[%d] is NULL", k);
        vec = vec && host_aligned16(models[k]);
    }
    const size_t smem = (size_t)K * kMhBins * sizeof(uint32_t);
    // exactly one resident wave: as many CTAs per SM as the shared-memory histograms allow (3 at K = 8), no tail wave
    int per_sm = (int)((size_t)(200 * 1024) / smem);
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    int64_t blocks = ((d + 3) / 4 + kMhThreads - 1) / kMhThreads;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    MR_DISPATCH_K(K, {
        PtrPack<KK> pk;
        for (int k = 0; k < KK; ++k) pk.p[k] = models[k];
        unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
        unsigned long long* a = reinterpret_cast<unsigned long long*>(above);
        const MhCand cd{cand, cand_count, (uint32_t)(cand ? cand_cap : 0)};
        if (vec) {
            cudaFuncSetAttribute(mag_hist_kernel<KK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            mag_hist_kernel<KK, true><<<(unsigned)blocks, kMhThreads, smem, (cudaStream_t)stream>>>(base, pk, d, w, lo, shift, h, a, cd);
        } else {
            cudaFuncSetAttribute(mag_hist_kernel<KK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            mag_hist_kernel<KK, false><<<(unsigned)blocks, kMhThreads, smem, (cudaStream_t)stream>>>(base, pk, d, w, lo, shift, h, a, cd);
        }
    });
    MR_CUDA_LAUNCH_CHECK("mr_ties_mag_hist");
    return MR_OK;
}
