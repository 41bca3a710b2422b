// lambda_grad.cu -- the lambda-gradient reduction of collaborative merging (A5).
//
//   out[g, k] = sum_{p : seg_group[p] == g} sum_{j < seg_len[p]} grad_p[j] * T[k, seg_off[p] + j]
//
// The reference gets this from autograd of `base + (w[:,None] * T).sum(0)` plus P slice views
// (rec_retrieval/merger/weight_learning/module/task_wise.py:43-47, layer_wise.py:76-82,
// weight_learning/utils.py:11-15,43-51), which costs P full-size SliceBackward tensors and a (K,d)
// temporary.  Here every gradient tensor is read in place through a pointer table and T is read exactly
// once: algorithmic traffic (K+1)*d*4 bytes, HBM-bound.
//
// Deterministic two-stage reduction (no atomics): stage 1 writes one fp64 partial per (chunk, k),
// stage 2 sums the partials of each group in a fixed order.  fp32 FMAs inside a thread (<= 64 terms per
// accumulator), fp64 from the warp reduction on: relative error vs an fp64 dot product ~1e-7.
#include "common.cuh"

namespace mr {

constexpr int kLgThreads = 256;
constexpr int64_t kLgChunk = 16384;  // elements per work unit

// ---- prologue: cum[p] = number of chunks before segment p (exclusive scan), cum[P] = total --------
__global__ void lg_chunk_table_kernel(const int64_t* __restrict__ seg_len, int P, int64_t* __restrict__ cum) {
    __shared__ int64_t s_part[1024];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int p0 = 0; p0 < P; p0 += blockDim.x) {
        const int p = p0 + threadIdx.x;
        const int64_t mine = (p < P) ? (seg_len[p] + kLgChunk - 1) / kLgChunk : 0;
        s_part[threadIdx.x] = mine;
        __syncthreads();
        for (int off = 1; off < (int)blockDim.x; off <<= 1) {  // Hillis-Steele inclusive scan
            int64_t v = (threadIdx.x >= (unsigned)off) ? s_part[threadIdx.x - off] : 0;
            __syncthreads();
            s_part[threadIdx.x] += v;
            __syncthreads();
        }
        if (p < P) cum[p] = s_carry + s_part[threadIdx.x] - mine;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry += s_part[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) cum[P] = s_carry;
}

template <int V>
struct VecLoad;
template <>
struct VecLoad<4> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
        float4 r = ldg_stream4(p);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
};
template <>
struct VecLoad<2> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v[0]), "=f"(v[1]) : "l"(p));
    }
};
template <>
struct VecLoad<1> {
    static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = ldg_stream1(p); }
};

template <int K, int V>
__device__ __forceinline__ void lg_accumulate(const float* __restrict__ g, const float* __restrict__ t0, int64_t ldT,
                                              int n, float (&acc)[K]) {
    const int nv = n / V;
    for (int i = threadIdx.x; i < nv; i += kLgThreads) {
        float gv[V];
        VecLoad<V>::ld(g + (int64_t)i * V, gv);
        float tv[K][V];
#pragma unroll
        for (int k = 0; k < K; ++k) VecLoad<V>::ld(t0 + (int64_t)k * ldT + (int64_t)i * V, tv[k]);
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int c = 0; c < V; ++c) acc[k] = fmaf(gv[c], tv[k][c], acc[k]);
    }
    for (int i = nv * V + threadIdx.x; i < n; i += kLgThreads) {  // < V leftover elements
        const float gs = g[i];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = fmaf(gs, t0[(int64_t)k * ldT + i], acc[k]);
    }
}

template <int K>
__global__ void __launch_bounds__(kLgThreads)
lg_partial_kernel(const float* const* __restrict__ grad_ptrs, const int64_t* __restrict__ seg_off,
                  const int64_t* __restrict__ seg_len, int P, const float* __restrict__ T, int64_t ldT,
                  const int64_t* __restrict__ cum, double* __restrict__ partial) {
    __shared__ double s_red[kLgThreads / 32][K];
    const int64_t nchunks = cum[P];
    for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
        int lo = 0, hi = P - 1;  // last p with cum[p] <= c
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (cum[mid] <= c) lo = mid; else hi = mid - 1;
        }
        const int p = lo;
        const float* g = grad_ptrs[p];
        float acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = 0.0f;
        if (g != nullptr) {  // a NULL gradient (parameter unused in the forward) contributes zero
            const int64_t start = (c - cum[p]) * kLgChunk;
            const int64_t rem = seg_len[p] - start;
            const int n = (int)(rem < kLgChunk ? rem : kLgChunk);
            const float* gp = g + start;
            const float* tp = T + seg_off[p] + start;
            const uintptr_t both = reinterpret_cast<uintptr_t>(gp) | reinterpret_cast<uintptr_t>(tp) |
                                   (uintptr_t)((ldT & 3) * 4);
            if ((both & 15) == 0) lg_accumulate<K, 4>(gp, tp, ldT, n, acc);
            else if ((both & 7) == 0) lg_accumulate<K, 2>(gp, tp, ldT, n, acc);
            else lg_accumulate<K, 1>(gp, tp, ldT, n, acc);
        }
        double dacc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double v = (double)acc[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            dacc[k] = v;
        }
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) s_red[warp][k] = dacc[k];
        }
        __syncthreads();
        if (threadIdx.x < K) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kLgThreads / 32; ++w) v += s_red[w][threadIdx.x];
            partial[c * K + threadIdx.x] = v;
        }
        __syncthreads();
    }
}

// stage 2: one block per group; fixed summation tree -> deterministic
template <int K>
__global__ void __launch_bounds__(256)
lg_final_kernel(const int32_t* __restrict__ seg_group, int P, const int64_t* __restrict__ cum,
                const double* __restrict__ partial, float* __restrict__ out) {
    __shared__ double s_red[8][K];
    const int g = blockIdx.x;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int p = 0; p < P; ++p) {
        if ((seg_group ? seg_group[p] : 0) != g) continue;
        for (int64_t c = cum[p] + threadIdx.x; c < cum[p + 1]; c += blockDim.x) {
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] += partial[c * K + k];
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        acc[k] = v;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_red[warp][k] = acc[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
        out[g * K + threadIdx.x] = (float)v;
    }
}

static inline int64_t lg_max_chunks(int64_t d, int P) { return d / kLgChunk + P; }
static inline size_t lg_cum_bytes(int P) { return (((size_t)(P + 1) * 8) + 255) & ~(size_t)255; }

}  // namespace mr

extern "C" int64_t mr_lambda_grad_workspace_bytes(int64_t d, int P, int K) {
    if (d < 0 || P < 1 || K < 1) return 0;
    return (int64_t)(mr::lg_cum_bytes(P) + (size_t)mr::lg_max_chunks(d, P) * K * sizeof(double));
}

extern "C" int mr_lambda_grad(const float* const* grad_ptrs, const int64_t* seg_off, const int64_t* seg_len,
                              const int32_t* seg_group, int P, int64_t d, const float* T, int64_t ldT, int K, int G,
                              float* out, void* ws, int64_t ws_bytes, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(P >= 1 && G >= 1 && d >= 0, "mr_lambda_grad: need P >= 1, G >= 1, d >= 0");
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_lambda_grad: K=%d outside [1,%d]", K, MR_MAX_K);
    MR_REQUIRE(grad_ptrs && seg_off && seg_len && T && out && ws, "mr_lambda_grad: null pointer");
    MR_REQUIRE(G == 1 || seg_group, "mr_lambda_grad: G > 1 needs seg_group");
    MR_REQUIRE(ldT >= d, "mr_lambda_grad: need ldT >= d");
    if (ws_bytes < mr_lambda_grad_workspace_bytes(d, P, K)) {
        set_error("mr_lambda_grad: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
                  (long long)mr_lambda_grad_workspace_bytes(d, P, K));
        return MR_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t* cum = reinterpret_cast<int64_t*>(ws);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + lg_cum_bytes(P));
    lg_chunk_table_kernel<<<1, 1024, 0, st>>>(seg_len, P, cum);
    int64_t blocks = lg_max_chunks(d, P);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    MR_DISPATCH_K(K, {
        lg_partial_kernel<KK><<<(unsigned)blocks, kLgThreads, 0, st>>>(grad_ptrs, seg_off, seg_len, P, T, ldT, cum, partial);
        lg_final_kernel<KK><<<G, 256, 0, st>>>(seg_group, P, cum, partial, out);
    });
    MR_CUDA_LAUNCH_CHECK("mr_lambda_grad");
    return MR_OK;
}
