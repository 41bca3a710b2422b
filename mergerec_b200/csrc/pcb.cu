// pcb.cu -- PCB-merging vectors (reference: rec_retrieval/merger/algorithms/pcb.py:9-58).
//
//   tau_k  = m_k - base                                 A_k = clamp(|tau_k|, lo_k, hi_k)   (1% / 99% order statistics of
//   self_k = ((A_k - lo_k) / (hi_k - lo_k))^2                                                |tau_k|, found by the TIES select)
//   task_k = exp(n * self_k) * tanh(tau_k * sum_k tau_k)
//   scale_k = (clamp(task_k, q_k, max_k) - q_k) / (max_k - q_k)     q_k = the int(d (1 - density))-th smallest task_k
//   out_k  = sign(tau_k) A_k scale_k / max(sum_k scale_k, 1e-12) / n
//
// The quantile q_k of the COMPUTED values is an exact order statistic: histogram passes over the order-preserving
// integer image of the float that recompute task_k from base + models each time (dense: 11 + 11 + 10 bits over
// everything; fast: the same on a 1/32 sample, then two windowed passes over everything) -- nothing of size (K, d) is
// materialised besides the output.  Every column sum uses torch.sum(dim=0)'s order (common.cuh).  exp / tanh
// are CUDA's expf / tanhf; torch's CPU kernels use a different libm, so values agree to ~1 ulp, not bit for bit, and
// an element whose task value lies within that distance of q_k may fall on the other side of the clamp (the tests
// confine every difference from the reference to such columns).  This is a baseline merger, not a tuned hot path.
#include <math.h>

#include "common.cuh"

namespace mr {

constexpr int kPcbThreads = 256;
constexpr int kPcbBins = 2048;

struct PcbState {
    uint32_t prefix;  // key bits decided so far
    uint32_t maxkey;  // order-preserving key of the row maximum
    int64_t rank;     // remaining ascending rank inside the chosen bucket
    uint32_t qkey;    // final key of the quantile element
    uint32_t wlo;     // windowed search: first key of the window
    int32_t wshift;   //                  log2(keys per bin)
    int32_t wwidth;   //                  bins of the window that belong to the search
    int32_t miss;     //                  the window did not contain the wanted rank (caller falls back to the dense passes)
    int32_t pad[3];
};

__device__ __forceinline__ uint32_t pcb_key(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float pcb_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// task_k of one column (pcb.py:44-53); also returns tau and A for the build kernel.
template <int K>
__device__ __forceinline__ void pcb_task(float b, const float (&x)[K], const float* __restrict__ lo,
                                         const float* __restrict__ hi, bool tail, float (&tau)[K], float (&A)[K],
                                         float (&task)[K]) {
#pragma unroll
    for (int k = 0; k < K; ++k) tau[k] = __fsub_rn(x[k], b);
    const float total = torch_sum_dim0<K>(tau, tail);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        A[k] = fminf(fmaxf(fabsf(tau[k]), lo[k]), hi[k]);
        const float nrm = __fdiv_rn(__fsub_rn(A[k], lo[k]), __fsub_rn(hi[k], lo[k]));
        const float self_act = expf(__fmul_rn((float)K, __fmul_rn(nrm, nrm)));
        task[k] = __fmul_rn(self_act, tanhf(__fmul_rn(tau[k], total)));
    }
}

// pass 0: bins = key >> 21 (+ row maxima); pass 1: (key >> 10) & 2047 inside prefix; pass 2: key & 1023 inside prefix
template <int K>
__global__ void __launch_bounds__(kPcbThreads)
pcb_hist_kernel(const float* __restrict__ base, PtrPack<K> m, int64_t d, int64_t stride, const float* __restrict__ lo,
                const float* __restrict__ hi, int pass, PcbState* __restrict__ st, uint32_t* __restrict__ hist) {
    extern __shared__ uint32_t s_hist[];  // K * kPcbBins
    __shared__ float s_lo[K], s_hi[K];
    __shared__ uint32_t s_prefix[K];
    for (int i = threadIdx.x; i < K * kPcbBins; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < K) {
        s_lo[threadIdx.x] = lo[threadIdx.x];
        s_hi[threadIdx.x] = hi[threadIdx.x];
        s_prefix[threadIdx.x] = pass ? st[threadIdx.x].prefix : 0;
    }
    __syncthreads();
    uint32_t mx[K];
#pragma unroll
    for (int k = 0; k < K; ++k) mx[k] = 0;
    const int64_t tail0 = d & ~(int64_t)31;
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_vis = (d + stride - 1) / stride;   // stride > 1: a strided sample (columns 0, stride, 2 stride, ...)
    const int64_t rounds = (n_vis + span - 1) / span;  // every lane runs every round (the match below is warp-wide)
    for (int64_t it = 0; it < rounds; ++it) {
        const int64_t j = (it * span + (int64_t)blockIdx.x * blockDim.x + threadIdx.x) * stride;
        const bool live = j < d;
        float x[K], tau[K], A[K], task[K];
        const float b = live ? ldg_stream1(base + j) : 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = live ? ldg_stream1(m.p[k] + j) : 0.f;
        pcb_task<K>(b, x, s_lo, s_hi, j >= tail0, tau, A, task);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t key = pcb_key(task[k]);
            int bin = -1;
            if (live) {
                if (pass == 0) { bin = (int)(key >> 21); mx[k] = key > mx[k] ? key : mx[k]; }
                else if (pass == 1) { if ((key >> 21) == s_prefix[k]) bin = (int)((key >> 10) & 2047u); }
                else { if ((key >> 10) == s_prefix[k]) bin = (int)(key & 1023u); }
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, bin);  // one shared-memory atomic per distinct bin
            if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[k * kPcbBins + bin], __popc(peers));
        }
    }
    if (pass == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            uint32_t v = mx[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const uint32_t o = __shfl_xor_sync(0xffffffffu, v, off);
                v = o > v ? o : v;
            }
            if ((threadIdx.x & 31) == 0 && v) atomicMax(&st[k].maxkey, v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * kPcbBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// one warp per model: walk the histogram to the bucket holding the wanted rank, then clear it for the next pass
__global__ void pcb_pick_kernel(int pass, PcbState* __restrict__ st, uint32_t* __restrict__ hist) {
    const int k = blockIdx.x;
    uint32_t* h = hist + (size_t)k * kPcbBins;
    if (threadIdx.x == 0) {
        const int64_t rank = st[k].rank;
        int64_t cum = 0;
        int chosen = kPcbBins - 1;
        for (int b = 0; b < kPcbBins; ++b) {
            const int64_t c = h[b];
            if (rank < cum + c) { chosen = b; break; }
            cum += c;
        }
        st[k].rank = rank - cum;
        const uint32_t p = (pass == 0) ? (uint32_t)chosen
                                       : ((st[k].prefix << (pass == 1 ? 11 : 10)) | (uint32_t)chosen);
        st[k].prefix = p;
        if (pass == 2) st[k].qkey = p;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kPcbBins; b += blockDim.x) h[b] = 0;
}

// ---- windowed search (fast path) -----------------------------------------------------------------------------------
// The dense passes above send every element through a shared-memory atomic and a handful of bins take most of them.
// The fast path runs them on a 1/32 sample only, centres a window of 2^21 keys on the sample's quantile and then
// counts, over the whole vector, the elements BELOW the window (a register counter) and histograms the few inside it;
// one refinement (1024 keys per bin -> 1) pins the exact key.  A window that does not hold the wanted rank sets `miss`.
template <int K>
__global__ void __launch_bounds__(kPcbThreads)
pcb_window_kernel(const float* __restrict__ base, PtrPack<K> m, int64_t d, const float* __restrict__ lo,
                  const float* __restrict__ hi, int track_max, PcbState* __restrict__ st, uint32_t* __restrict__ hist,
                  unsigned long long* __restrict__ below_out) {
    extern __shared__ uint32_t s_hist[];  // K * kPcbBins
    __shared__ float s_lo[K], s_hi[K];
    __shared__ uint32_t s_wlo[K];
    __shared__ int s_wsh[K], s_ww[K];
    for (int i = threadIdx.x; i < K * kPcbBins; i += blockDim.x) s_hist[i] = 0;
    if (threadIdx.x < K) {
        s_lo[threadIdx.x] = lo[threadIdx.x];
        s_hi[threadIdx.x] = hi[threadIdx.x];
        s_wlo[threadIdx.x] = st[threadIdx.x].wlo;
        s_wsh[threadIdx.x] = st[threadIdx.x].wshift;
        s_ww[threadIdx.x] = st[threadIdx.x].wwidth;
    }
    __syncthreads();
    uint32_t mx[K], below[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { mx[k] = 0; below[k] = 0; }
    const int64_t tail0 = d & ~(int64_t)31;
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < d; j += span) {
        float x[K], tau[K], A[K], task[K];
        const float b = ldg_stream1(base + j);
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = ldg_stream1(m.p[k] + j);
        pcb_task<K>(b, x, s_lo, s_hi, j >= tail0, tau, A, task);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t key = pcb_key(task[k]);
            mx[k] = key > mx[k] ? key : mx[k];
            if (key < s_wlo[k]) {
                ++below[k];
            } else {
                const uint32_t bin = (key - s_wlo[k]) >> s_wsh[k];
                if (bin < (uint32_t)s_ww[k]) atomicAdd(&s_hist[k * kPcbBins + bin], 1u);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t v = below[k], mv = mx[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, off);
            const uint32_t o = __shfl_xor_sync(0xffffffffu, mv, off);
            mv = o > mv ? o : mv;
        }
        if ((threadIdx.x & 31) == 0) {
            if (v) atomicAdd(&below_out[k], (unsigned long long)v);
            if (track_max && mv) atomicMax(&st[k].maxkey, mv);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * kPcbBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// after the sample passes: centre the first window on the sample's quantile key
__global__ void pcb_window_init_kernel(PcbState* st, int K, int64_t q_index) {
    const int k = threadIdx.x;
    if (k < K) {
        const uint32_t q = st[k].qkey;
        st[k].wlo = q > (1u << 20) ? q - (1u << 20) : 0u;
        st[k].wshift = 10;         // 2048 bins x 1024 keys = 2^21 keys (about +-12 % around the sample's quantile)
        st[k].wwidth = kPcbBins;
        st[k].rank = q_index;      // ascending rank wanted in the FULL vector
        st[k].maxkey = 0;          // the sample's maximum is not the row maximum
        st[k].miss = 0;
    }
}

// one warp per model: ascending walk of the window's histogram; first level subtracts what lies below the window
__global__ void pcb_window_pick_kernel(int first, PcbState* __restrict__ st, uint32_t* __restrict__ hist,
                                       unsigned long long* __restrict__ below) {
    const int k = blockIdx.x;
    uint32_t* h = hist + (size_t)k * kPcbBins;
    if (threadIdx.x == 0 && !st[k].miss) {
        int64_t rank = st[k].rank;
        const int width = st[k].wwidth;
        bool ok = true;
        if (first) {
            int64_t total = 0;
            for (int b = 0; b < width; ++b) total += h[b];
            const int64_t bl = (int64_t)below[k];
            if (rank < bl || rank >= bl + total) ok = false;
            rank -= bl;
        }
        if (!ok) {
            st[k].miss = 1;
        } else {
            int64_t cum = 0;
            int chosen = width - 1;
            for (int b = 0; b < width; ++b) {
                const int64_t c = h[b];
                if (rank < cum + c) { chosen = b; break; }
                cum += c;
            }
            const int sh = st[k].wshift;
            const uint32_t nlo = st[k].wlo + ((uint32_t)chosen << sh);
            st[k].rank = rank - cum;
            st[k].wlo = nlo;
            if (sh == 0) {
                st[k].qkey = nlo;
            } else {
                const int nsh = sh > 11 ? sh - 11 : 0;
                st[k].wwidth = (1 << sh) >> nsh;
                st[k].wshift = nsh;
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < kPcbBins; b += blockDim.x) h[b] = 0;
    if (threadIdx.x == 0) below[k] = 0;
}

__global__ void pcb_status_kernel(const PcbState* st, int K, int32_t* status) {
    const int k = threadIdx.x;
    if (k < K) status[k] = st[k].miss ? 0 : 1;
}

__global__ void pcb_init_kernel(PcbState* st, int K, int64_t q_index) {
    const int k = threadIdx.x;
    if (k < K) {
        st[k].prefix = 0; st[k].maxkey = 0; st[k].rank = q_index; st[k].qkey = 0;
        st[k].wlo = 0; st[k].wshift = 0; st[k].wwidth = 0; st[k].miss = 0;
    }
}

template <int K>
__global__ void __launch_bounds__(kPcbThreads)
pcb_build_kernel(const float* __restrict__ base, PtrPack<K> m, int64_t d, const float* __restrict__ lo,
                 const float* __restrict__ hi, const PcbState* __restrict__ st, float* __restrict__ out, int64_t ldo,
                 float* __restrict__ task_out, float* __restrict__ thr_out) {
    __shared__ float s_lo[K], s_hi[K], s_q[K], s_max[K];
    if (threadIdx.x < K) {
        s_lo[threadIdx.x] = lo[threadIdx.x];
        s_hi[threadIdx.x] = hi[threadIdx.x];
        s_q[threadIdx.x] = pcb_unkey(st[threadIdx.x].qkey);
        s_max[threadIdx.x] = pcb_unkey(st[threadIdx.x].maxkey);
        if (thr_out && blockIdx.x == 0) {
            thr_out[2 * threadIdx.x] = s_q[threadIdx.x];
            thr_out[2 * threadIdx.x + 1] = s_max[threadIdx.x];
        }
    }
    __syncthreads();
    const int64_t tail0 = d & ~(int64_t)31;
    const int64_t span = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < d; j += span) {
        float x[K], tau[K], A[K], task[K], scale[K];
        const float b = ldg_stream1(base + j);
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = ldg_stream1(m.p[k] + j);
        const bool tail = j >= tail0;
        pcb_task<K>(b, x, s_lo, s_hi, tail, tau, A, task);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float cl = fminf(fmaxf(task[k], s_q[k]), s_max[k]);
            scale[k] = __fdiv_rn(__fsub_rn(cl, s_q[k]), __fsub_rn(s_max[k], s_q[k]));
        }
        const float denom = fmaxf(torch_sum_dim0<K>(scale, tail), (float)1e-12);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float sgn = (tau[k] > 0.f) ? 1.f : ((tau[k] < 0.f) ? -1.f : 0.f);
            const float v = __fmul_rn(__fmul_rn(sgn, A[k]), scale[k]);
            out[(int64_t)k * ldo + j] = __fdiv_rn(__fdiv_rn(v, denom), (float)K);
            if (task_out) task_out[(int64_t)k * ldo + j] = task[k];
        }
    }
}

static inline size_t pcb_state_bytes() { return (size_t)MR_MAX_K * sizeof(PcbState); }

}  // namespace mr

extern "C" int64_t mr_pcb_workspace_bytes(int K) {
    if (K < 1 || K > MR_MAX_K) return 0;
    return (int64_t)(mr::pcb_state_bytes() + (size_t)K * mr::kPcbBins * sizeof(uint32_t) + (size_t)MR_MAX_K * sizeof(unsigned long long));
}

extern "C" int mr_pcb_vectors(const float* base, const float* const* models, int K, int64_t d, const float* clamp_lo,
                              const float* clamp_hi, int64_t q_index, int dense, int32_t* status, float* out, int64_t ldo,
                              float* task_out, float* thr_out, void* ws, int64_t ws_bytes, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_pcb_vectors: K=%d outside [1,%d]", K, MR_MAX_K);
    MR_REQUIRE(base && models && clamp_lo && clamp_hi && out && ws && status, "mr_pcb_vectors: null pointer");
    MR_REQUIRE(d >= 1 && d < ((int64_t)1 << 32), "mr_pcb_vectors: d=%lld outside [1, 2^32)", (long long)d);
    MR_REQUIRE(q_index >= 0 && q_index < d, "mr_pcb_vectors: quantile index %lld outside [0,d)", (long long)q_index);
    MR_REQUIRE(ldo >= d, "mr_pcb_vectors: need ldo >= d");
    for (int k = 0; k < K; ++k) MR_REQUIRE(models[k] != nullptr, "mr_pcb_vectors: models[%d] is NULL", k);
    if (ws_bytes < mr_pcb_workspace_bytes(K)) {
        set_error("mr_pcb_vectors: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
                  (long long)mr_pcb_workspace_bytes(K));
        return MR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    PcbState* state = reinterpret_cast<PcbState*>(ws);
    uint32_t* hist = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + pcb_state_bytes());
    const size_t hist_bytes = (size_t)K * kPcbBins * sizeof(uint32_t);
    unsigned long long* below = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(hist) + hist_bytes);
    cudaMemsetAsync(hist, 0, hist_bytes + (size_t)MR_MAX_K * sizeof(unsigned long long), st);
    const int64_t stride = 32;                         // sample of the fast path
    const int64_t n_s = (d + stride - 1) / stride;
    const bool fast = !dense && n_s >= 4096;           // small vectors: the dense passes are cheap and exact anyway
    // ascending rank of the quantile inside the sample
    const int64_t q_s = fast ? (int64_t)((long double)q_index * (long double)n_s / (long double)d) : q_index;
    pcb_init_kernel<<<1, 32, 0, st>>>(state, K, fast ? (q_s < n_s ? q_s : n_s - 1) : q_index);
    int64_t blocks = (d + kPcbThreads - 1) / kPcbThreads;
    const int64_t cap = (int64_t)sm_count() * 3;       // one resident wave at K = 8 (64 KB of histograms per CTA)
    if (blocks > cap) blocks = cap;
    int64_t sblocks = (n_s + kPcbThreads - 1) / kPcbThreads;
    if (sblocks > cap) sblocks = cap;
    MR_DISPATCH_K(K, {
        PtrPack<KK> pk;
        for (int k = 0; k < KK; ++k) pk.p[k] = models[k];
        cudaFuncSetAttribute(pcb_hist_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
        cudaFuncSetAttribute(pcb_window_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hist_bytes);
        for (int pass = 0; pass < 3; ++pass) {          // dense radix passes: on the sample (fast) or on everything
            pcb_hist_kernel<KK><<<(unsigned)(fast ? sblocks : blocks), kPcbThreads, hist_bytes, st>>>(
                base, pk, d, fast ? stride : 1, clamp_lo, clamp_hi, pass, state, hist);
            pcb_pick_kernel<<<KK, 256, 0, st>>>(pass, state, hist);
        }
        if (fast) {
            pcb_window_init_kernel<<<1, 32, 0, st>>>(state, KK, q_index);
            for (int level = 0; level < 2; ++level) {
                pcb_window_kernel<KK><<<(unsigned)blocks, kPcbThreads, hist_bytes, st>>>(base, pk, d, clamp_lo, clamp_hi,
                                                                                         level == 0, state, hist, below);
                pcb_window_pick_kernel<<<KK, 256, 0, st>>>(level == 0, state, hist, below);
            }
        }
        pcb_status_kernel<<<1, 32, 0, st>>>(state, KK, status);
        pcb_build_kernel<KK><<<(unsigned)blocks, kPcbThreads, 0, st>>>(base, pk, d, clamp_lo, clamp_hi, state, out, ldo, task_out, thr_out);
    });
    MR_CUDA_LAUNCH_CHECK("mr_pcb_vectors");
    return MR_OK;
}
