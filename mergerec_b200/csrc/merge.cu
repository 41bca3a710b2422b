// merge.cu -- lambda-weighted task-vector merge (A1, A2, A3, A4, A10).
//
// One streaming pass: read base + K sources with 128-bit loads, K unrolled, write merged once.
// Algorithmic traffic (K+2)*d*4 bytes; HBM-bound.  The arithmetic is *unfused* fp32 in exactly the
// reference's order (the torch CPU kernels it replaces never contract mul+add), so results are
// bit-identical to the reference, not merely within 1e-6.
//
// reference: rec_retrieval/merger/algorithms/task_vector.py:8-34, linear.py:8-27,
//            rec_retrieval/merger/weight_learning/module/task_wise.py:36-48, layer_wise.py:64-83
#include "common.cuh"

namespace mr {

constexpr int kMergeThreads = 256;
constexpr int kMaxSmemSegs = 2048;
constexpr int kMaxMergeSmem = 224 * 1024;   // dynamic shared memory a CTA may opt in to on sm_100a (227 KB) minus slack

struct SegView {
    const int64_t* seg_end;    // dev, P entries (ascending, exclusive)
    const int32_t* seg_group;  // dev, P entries
    int P;
};

// Locate the block holding flat index j (first p with seg_end[p] > j). `hint` caches the last hit.
__device__ __forceinline__ int find_seg(const int64_t* s_end, int P, int64_t j, int hint) {
    if (j < s_end[hint] && (hint == 0 || j >= s_end[hint - 1])) return hint;
    int lo = 0, hi = P - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (s_end[mid] > j) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int K, int ORDER, bool SRC_IS_MODEL>
__device__ __forceinline__ float merge_one(float b, const float (&x)[K], const float* __restrict__ w, bool tail) {
    float prod[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float u = SRC_IS_MODEL ? __fsub_rn(x[k], b) : x[k];
        prod[k] = __fmul_rn(w[k], u);
    }
    if (ORDER == MR_ORDER_BASE_FIRST) {
        float acc = b;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = __fadd_rn(acc, prod[k]);
        return acc;
    } else if (ORDER == MR_ORDER_LINEAR) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc = __fadd_rn(acc, prod[k]);
        return acc;
    } else {
        return __fadd_rn(b, torch_sum_dim0<K>(prod, tail));
    }
}

// SEGMENTED = false: one block [0,d), weights row 0.  true: per-block group weights + per-block tails.
template <int K, int ORDER, bool SRC_IS_MODEL, bool SEGMENTED, bool VEC>
__global__ void __launch_bounds__(kMergeThreads)
merge_kernel(const float* __restrict__ base, PtrPack<K> src, int64_t d, const float* __restrict__ w, int G,
             SegView segs, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* s_w = reinterpret_cast<float*>(smem_raw);                       // G*K
    int64_t* s_end = reinterpret_cast<int64_t*>(smem_raw + ((G * K * 4 + 15) & ~15));  // P (SEGMENTED)
    int32_t* s_grp = reinterpret_cast<int32_t*>(s_end + (SEGMENTED ? segs.P : 0));

    for (int i = threadIdx.x; i < G * K; i += blockDim.x) s_w[i] = w[i];
    if (SEGMENTED) {
        for (int i = threadIdx.x; i < segs.P; i += blockDim.x) {
            s_end[i] = segs.seg_end[i];
            s_grp[i] = segs.seg_group[i];
        }
    }
    __syncthreads();

    float wreg[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wreg[k] = s_w[k];
    const int64_t tail0 = d & ~(int64_t)31;  // un-segmented: interleaved order from here on (K >= 5)

    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    int hint = 0;

    if (VEC) {
        const int64_t n4 = d >> 2;
        for (int64_t v = gtid; v < n4; v += gsz) {
            const int64_t j = v << 2;
            float4 xb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ORDER != MR_ORDER_LINEAR) xb = ldg_stream4(base + j);
            float4 xs[K];
#pragma unroll
            for (int k = 0; k < K; ++k) xs[k] = ldg_stream4(src.p[k] + j);
            float bx[4] = {xb.x, xb.y, xb.z, xb.w};
            float r[4];
            if (!SEGMENTED) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) x[k] = reinterpret_cast<const float*>(&xs[k])[c];
                    r[c] = merge_one<K, ORDER, SRC_IS_MODEL>(bx[c], x, wreg, (j + c) >= tail0);
                }
            } else {
                hint = find_seg(s_end, segs.P, j, hint);
                int p = hint;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    while (j + c >= s_end[p]) ++p;  // a vector may straddle blocks (Recformer offsets)
                    const int64_t beg = p ? s_end[p - 1] : 0;
                    const int64_t n = s_end[p] - beg;
                    const bool tail = (j + c - beg) >= (n & ~(int64_t)31);
                    float x[K];
#pragma unroll
                    for (int k = 0; k < K; ++k) x[k] = reinterpret_cast<const float*>(&xs[k])[c];
                    r[c] = merge_one<K, ORDER, SRC_IS_MODEL>(bx[c], x, s_w + s_grp[p] * K, tail);
                }
            }
            stg_stream4(out + j, make_float4(r[0], r[1], r[2], r[3]));
        }
    }
    // scalar path: the d mod 4 remainder (VEC) or everything (unaligned pointers)
    const int64_t s0 = VEC ? (d & ~(int64_t)3) : 0;
    for (int64_t j = s0 + gtid; j < d; j += gsz) {
        float b = (ORDER != MR_ORDER_LINEAR) ? base[j] : 0.0f;
        float x[K];
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = src.p[k][j];
        if (!SEGMENTED) {
            out[j] = merge_one<K, ORDER, SRC_IS_MODEL>(b, x, wreg, j >= tail0);
        } else {
            hint = find_seg(s_end, segs.P, j, hint);
            const int64_t beg = hint ? s_end[hint - 1] : 0;
            const int64_t n = s_end[hint] - beg;
            out[j] = merge_one<K, ORDER, SRC_IS_MODEL>(b, x, s_w + s_grp[hint] * K, (j - beg) >= (n & ~(int64_t)31));
        }
    }
}

template <int K, int ORDER, bool SRC_IS_MODEL>
static int launch_merge(const float* base, const float* const* src, int64_t d, const float* w, int G,
                        const int64_t* seg_end, const int32_t* seg_group, int P, float* out, cudaStream_t st) {
    PtrPack<K> pack;
    bool vec = host_aligned16(out) && (ORDER == MR_ORDER_LINEAR || host_aligned16(base));
    for (int k = 0; k < K; ++k) {
        pack.p[k] = src[k];
        vec = vec && host_aligned16(src[k]);
    }
    const bool segmented = (P > 1) || (seg_end && seg_group);   // a one-block table may still name a group other than 0
    SegView sv{seg_end, seg_group, segmented ? P : 0};
    size_t smem = ((size_t)(G * K * 4 + 15) & ~(size_t)15) + (segmented ? (size_t)P * 12 : 0) + 16;
    const int64_t work = vec ? (d >> 2) + 3 : d;
    int64_t blocks = (work + kMergeThreads - 1) / kMergeThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    // tables above 48 KB of shared memory (more than ~4,000 blocks) need the opt-in; mr_merge_axpy caps P at what 227 KB hold
#define MR_LAUNCH(SEG, VEC)                                                                            \
    do {                                                                                               \
        auto kern = merge_kernel<K, ORDER, SRC_IS_MODEL, SEG, VEC>;                                    \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        kern<<<(unsigned)blocks, kMergeThreads, smem, st>>>(base, pack, d, w, G, sv, out);             \
    } while (0)
    if (segmented) { if (vec) MR_LAUNCH(true, true); else MR_LAUNCH(true, false); }
    else           { if (vec) MR_LAUNCH(false, true); else MR_LAUNCH(false, false); }
#undef MR_LAUNCH
    MR_CUDA_LAUNCH_CHECK("mr_merge_axpy");
    return MR_OK;
}

// ---- A2: task vectors ----------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(kMergeThreads)
task_vectors_kernel(const float* __restrict__ base, PtrPack<K> models, int64_t d, float* __restrict__ T,
                    int64_t ldT, bool vec) {
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gsz = (int64_t)gridDim.x * blockDim.x;
    int64_t s0 = 0;
    if (vec) {
        const int64_t n4 = d >> 2;
        for (int64_t v = gtid; v < n4; v += gsz) {
            const int64_t j = v << 2;
            const float4 b = ldg_stream4(base + j);
            float4 m[K];
#pragma unroll
            for (int k = 0; k < K; ++k) m[k] = ldg_stream4(models.p[k] + j);
#pragma unroll
            for (int k = 0; k < K; ++k)
                stg_stream4(T + (int64_t)k * ldT + j,
                            make_float4(__fsub_rn(m[k].x, b.x), __fsub_rn(m[k].y, b.y), __fsub_rn(m[k].z, b.z),
                                        __fsub_rn(m[k].w, b.w)));
        }
        s0 = d & ~(int64_t)3;
    }
    for (int64_t j = s0 + gtid; j < d; j += gsz) {
        const float b = base[j];
#pragma unroll
        for (int k = 0; k < K; ++k) T[(int64_t)k * ldT + j] = __fsub_rn(models.p[k][j], b);
    }
}

}  // namespace mr

extern "C" int mr_task_vectors(const float* base, const float* const* models, int K, int64_t d, float* T,
                               int64_t ldT, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(d >= 0 && ldT >= d, "mr_task_vectors: need ldT >= d >= 0");
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_task_vectors: K=%d outside [1,%d]", K, MR_MAX_K);
    if (d == 0) return MR_OK;
    MR_REQUIRE(base && models && T, "mr_task_vectors: null pointer");
    bool vec = host_aligned16(base) && host_aligned16(T) && (ldT % 4 == 0);
    for (int k = 0; k < K && k < MR_MAX_K; ++k) vec = vec && host_aligned16(models[k]);
    const int64_t work = vec ? (d >> 2) + 3 : d;
    int64_t blocks = (work + kMergeThreads - 1) / kMergeThreads;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    MR_DISPATCH_K(K, {
        PtrPack<KK> pack;
        for (int k = 0; k < KK; ++k) pack.p[k] = models[k];
        task_vectors_kernel<KK><<<(unsigned)blocks, kMergeThreads, 0, (cudaStream_t)stream>>>(base, pack, d, T, ldT, vec);
    });
    MR_CUDA_LAUNCH_CHECK("mr_task_vectors");
    return MR_OK;
}

extern "C" int mr_merge_axpy(const float* base, const float* const* src, int K, int64_t d, const float* w,
                             int G, const int64_t* seg_end, const int32_t* seg_group, int P, int order,
                             int src_is_model, float* out, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(d >= 0 && G >= 1 && P >= 1, "mr_merge_axpy: need d >= 0, G >= 1, P >= 1");
    MR_REQUIRE(K >= 1 && K <= MR_MAX_K, "mr_merge_axpy: K=%d outside [1,%d]", K, MR_MAX_K);
    if (d == 0) return MR_OK;
    MR_REQUIRE(src && w && out, "mr_merge_axpy: null pointer");
    MR_REQUIRE(order == MR_ORDER_LINEAR || base, "mr_merge_axpy: base required unless order == LINEAR");
    MR_REQUIRE(P == 1 || (seg_end && seg_group), "mr_merge_axpy: P > 1 needs seg_end and seg_group");
    MR_REQUIRE((size_t)G * K * 4 + (size_t)P * 12 + 64 <= (size_t)kMaxMergeSmem,
               "mr_merge_axpy: the block table (P=%d blocks, G=%d groups) does not fit %d bytes of shared memory", P, G, kMaxMergeSmem);
    MR_REQUIRE(order >= 0 && order <= 2, "mr_merge_axpy: bad order %d", order);
    MR_REQUIRE(!(order == MR_ORDER_LINEAR && src_is_model), "mr_merge_axpy: LINEAR takes sources as they are");
    cudaStream_t st = (cudaStream_t)stream;
    MR_DISPATCH_K(K, {
        if (order == MR_ORDER_BASE_FIRST) {
            return src_is_model ? launch_merge<KK, MR_ORDER_BASE_FIRST, true>(base, src, d, w, G, seg_end, seg_group, P, out, st)
                                : launch_merge<KK, MR_ORDER_BASE_FIRST, false>(base, src, d, w, G, seg_end, seg_group, P, out, st);
        } else if (order == MR_ORDER_SUM_FIRST) {
            return src_is_model ? launch_merge<KK, MR_ORDER_SUM_FIRST, true>(base, src, d, w, G, seg_end, seg_group, P, out, st)
                                : launch_merge<KK, MR_ORDER_SUM_FIRST, false>(base, src, d, w, G, seg_end, seg_group, P, out, st);
        } else {
            return launch_merge<KK, MR_ORDER_LINEAR, false>(base, src, d, w, G, seg_end, seg_group, P, out, st);
        }
    });
    return MR_OK;
}
