// distill.cu -- the distillation step that sits between the merger and the evaluator (SURVEY.md section 8(f),
// rank 1): per-sample catalogue logits of the merged model, the distillation loss against the single-domain
// ("teacher") logits, and the gradient with respect to the merged model's representation.
//
// reference: rec_retrieval/module/distiller/sequence/module.py:59-76 (`_forward_distill`: for every sample i a GEMV
//            `rep[i] @ item_embeddings[dataset_index].T`, a teacher row `score_embeddings[dataset_index][sequence_id]`,
//            `loss_fn(merged.unsqueeze(0), single.unsqueeze(0))`, then `torch.stack(losses).mean()`),
//            rec_retrieval/module/recommender/loss_fn.py:36-231 (the loss classes),
//            merge_train.py:116-126 (teacher logits = normalised sequence embeddings @ normalised item table).
//
// The reference launches B GEMVs, copies B teacher rows from host memory and runs ~10 small kernels per sample.
// Here one step is three streaming kernels:
//   ds_logits_kernel  every domain's item table is read ONCE for all samples of that domain (HBM-bound:
//                     rows * E * 4 bytes per table) -> logits (B, ld);
//   ds_loss_kernel    one CTA per sample over its two logit rows (L2-resident) -> loss[b] and dloss/dlogits;
//   ds_grad_kernel    second pass over the item tables: grad_rep[b] = sum_n gz[b, n] * items[n] (+ a finish kernel
//                     that sums the per-CTA partials in a fixed order and applies the upstream gradient).
// All reductions are deterministic (no atomics).  Arithmetic is fp32 with fp64 block sums in the loss kernel; the
// contract is a floating-point tolerance (tests: 2e-5 relative against an fp64 oracle), not bit equality.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace mr {

constexpr int kDsThreads = 256;
constexpr int kDsWarps = kDsThreads / 32;
constexpr int kDsNB = 4;  // samples of one domain handled by one pass over its item table
constexpr int kDsLossThreads = 256;
constexpr int kDsLossCluster = 8;  // CTAs (one thread-block cluster) sharing one sample
namespace cg = cooperative_groups;

struct DsGroup {
    const float* items;  // (rows, E) row-major
    int32_t rows;
    int32_t nb;
    int32_t sample[kDsNB];
    int32_t cta_begin;
    int32_t cta_count;
};
struct DsPlan {
    int32_t ngroups;
    int32_t total_ctas;
    DsGroup g[MR_DISTILL_MAX_GROUPS];
};
struct DsLossArgs {
    const float* teacher[MR_DISTILL_MAX_B];
    int32_t n[MR_DISTILL_MAX_B];
};

// Host: group the samples by domain (ascending domain, ascending sample index, <= kDsNB per group) and give
// every group a share of the grid proportional to its rows.  Deterministic, so the logits, gradient and finish
// launches of one step agree on the layout of the partials.
static int ds_build_plan(const float* const* item_ptrs, const int64_t* item_rows, int nD, const int32_t* sample_domain,
                         int B, DsPlan* plan) {
    MR_REQUIRE(item_ptrs && item_rows && sample_domain, "mr_distill: null host table");
    MR_REQUIRE(B >= 1 && B <= MR_DISTILL_MAX_B, "mr_distill: B=%d outside [1,%d]", B, MR_DISTILL_MAX_B);
    MR_REQUIRE(nD >= 1, "mr_distill: need at least one item table");
    int ng = 0;
    int64_t total_rows = 0;
    for (int d = 0; d < nD; ++d) {
        int in_group = kDsNB;  // forces a new group at the first sample of the domain
        for (int b = 0; b < B; ++b) {
            const int32_t sd = sample_domain[b];
            MR_REQUIRE(sd >= 0 && sd < nD, "mr_distill: sample %d has dataset index %d outside [0,%d)", b, sd, nD);
            if (sd != d) continue;
            MR_REQUIRE(item_ptrs[d] != nullptr, "mr_distill: item table %d is NULL", d);
            MR_REQUIRE(host_aligned16(item_ptrs[d]), "mr_distill: item table %d is not 16-byte aligned", d);
            MR_REQUIRE(item_rows[d] >= 0 && item_rows[d] <= INT32_MAX, "mr_distill: item table %d has %lld rows", d,
                       (long long)item_rows[d]);
            if (in_group == kDsNB) {
                MR_REQUIRE(ng < MR_DISTILL_MAX_GROUPS, "mr_distill: more than %d (domain, 4-sample) groups",
                           MR_DISTILL_MAX_GROUPS);
                DsGroup& G = plan->g[ng++];
                G.items = item_ptrs[d];
                G.rows = (int32_t)item_rows[d];
                G.nb = 0;
                for (int s = 0; s < kDsNB; ++s) G.sample[s] = 0;
                total_rows += item_rows[d];
                in_group = 0;
            }
            DsGroup& G = plan->g[ng - 1];
            G.sample[G.nb++] = b;
            ++in_group;
        }
    }
    // One persistent CTA per SM (the stage ring fills shared memory): hand out exactly sm_count CTAs in proportion to
    // the rows of each group (largest remainders first), at least one and at most one per 16 rows.
    const int64_t target = sm_count() > ng ? sm_count() : ng;
    int64_t given = 0;
    int64_t share[MR_DISTILL_MAX_GROUPS], rem[MR_DISTILL_MAX_GROUPS];
    for (int i = 0; i < ng; ++i) {
        const int64_t rows = plan->g[i].rows;
        const int64_t cap = rows > 0 ? (rows + 15) / 16 : 1;
        int64_t c = total_rows > 0 ? target * rows / total_rows : 1;
        rem[i] = total_rows > 0 ? (target * rows) % total_rows : 0;
        if (c < 1) { c = 1; rem[i] = 0; }
        if (c >= cap) { c = cap; rem[i] = -1; }
        share[i] = c;
        given += c;
    }
    while (given < target) {
        int best = -1;
        for (int i = 0; i < ng; ++i)
            if (rem[i] >= 0 && (best < 0 || rem[i] > rem[best])) best = i;
        if (best < 0) break;
        const int64_t cap = (plan->g[best].rows + 15) / 16;
        ++share[best];
        ++given;
        rem[best] = share[best] >= cap ? -1 : 0;
    }
    int next = 0;
    for (int i = 0; i < ng; ++i) {
        plan->g[i].cta_begin = next;
        plan->g[i].cta_count = (int32_t)share[i];
        next += (int)share[i];
    }
    plan->ngroups = ng;
    plan->total_ctas = next;
    return MR_OK;
}

static inline int64_t ds_max_ctas() { return (int64_t)sm_count() + 2 * MR_DISTILL_MAX_GROUPS; }

__device__ __forceinline__ int ds_find_group(const DsPlan& plan, int cta) {
    int gi = 0;
    while (gi + 1 < plan.ngroups && cta >= plan.g[gi + 1].cta_begin) ++gi;
    return gi;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    return fmaf(a.w, b.w, acc);
}

// Sum V per-lane values over the warp in V-1 + log2(32/V) shuffles instead of 5 V: each halving step sends one half
// of the values to the partner lane.  Afterwards the lanes with (lane % (32/V)) == 0 hold the total of value
// index lane / (32/V) (every lane of that sub-group does).
template <int V>
__device__ __forceinline__ float warp_multi_sum(float (&v)[V], int lane) {
    int off = 16;
#pragma unroll
    for (int n = V; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < n / 2; ++j) {
            const float send = up ? v[j] : v[j + n / 2];
            const float keep = up ? v[j + n / 2] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = v[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// ---- item-table streaming: 1-D TMA bulk copies into a ring of shared-memory stages ------------------------------------
// A CTA owns rows [r0, r1) of one table and walks them in tiles of R rows (R * E * 4 bytes, contiguous in HBM).
// Thread 0 keeps S tiles in flight (`cp.async.bulk` completing on one mbarrier per stage); the eight warps wait for a
// stage, consume two rows each per 16-row slab straight from shared memory, and a CTA barrier hands the stage back.
// Bytes in flight per SM = (S-1) tiles (~144 KB) independent of register pressure.
__device__ __forceinline__ uint32_t ds_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ds_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ds_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool ds_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ds_mbar_wait(uint32_t bar, uint32_t parity) {
    while (!ds_mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void ds_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct DsTiling {
    int R, S;
    size_t smem;
};
static inline DsTiling ds_tiling(int E) {
    DsTiling t;
    t.R = E > 384 ? 16 : 32;
    const size_t stage = (size_t)t.R * E * 4;
    const size_t fit = (size_t)(200 * 1024) / stage;
    t.S = (int)(fit < 4 ? fit : 4);
    t.smem = 128 + t.S * stage;
    return t;
}

struct DsStream {
    uint32_t bar0;
    float* stage0;
    const float* items;
    int64_t r0, r1;
    int R, S, E, ntiles;
    __device__ __forceinline__ void init(unsigned char* smem, const DsGroup& G, int local, int R_, int S_, int E_) {
        bar0 = ds_smem_u32(smem);
        stage0 = reinterpret_cast<float*>(smem + 128);
        items = G.items;
        r0 = (int64_t)G.rows * local / G.cta_count;
        r1 = (int64_t)G.rows * (local + 1) / G.cta_count;
        R = R_; S = S_; E = E_;
        ntiles = (int)((r1 - r0 + R - 1) / R);
        if (threadIdx.x == 0) {
            for (int s = 0; s < S; ++s) ds_mbar_init(bar0 + 8 * s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int t = 0; t < S && t < ntiles; ++t) issue(t);
    }
    __device__ __forceinline__ int rows_in(int tile) const {
        const int64_t left = r1 - (r0 + (int64_t)tile * R);
        return (int)(left < R ? left : R);
    }
    __device__ __forceinline__ void issue(int tile) const {
        const int s = tile % S;
        const uint32_t bytes = (uint32_t)rows_in(tile) * (uint32_t)E * 4u;
        ds_mbar_expect_tx(bar0 + 8 * s, bytes);
        ds_bulk_load(ds_smem_u32(stage0 + (size_t)s * R * E), items + (r0 + (int64_t)tile * R) * E, bytes, bar0 + 8 * s);
    }
    __device__ __forceinline__ const float4* wait(int tile) const {
        const int s = tile % S;
        ds_mbar_wait(bar0 + 8 * s, (uint32_t)((tile / S) & 1));
        return reinterpret_cast<const float4*>(stage0 + (size_t)s * R * E);
    }
    __device__ __forceinline__ void release(int tile) const {  // all threads; the stage is refilled with tile + S
        __syncthreads();
        if (threadIdx.x == 0 && tile + S < ntiles) issue(tile + S);
    }
};

// ---- logits[b, n] = <rep[b], items_dom(b)[n]> ---------------------------------------------------------------------
// The group's <= NB representation vectors stay in registers; a warp takes two rows of the stage at a time.
template <int EV, int NB>
__global__ void __launch_bounds__(kDsThreads, 1)
ds_logits_kernel(const __grid_constant__ DsPlan plan, const float* __restrict__ rep, int E, int R, int S,
                 float* __restrict__ logits, int64_t ld) {
    extern __shared__ __align__(128) unsigned char ds_smem[];
    const DsGroup& G = plan.g[ds_find_group(plan, (int)blockIdx.x)];
    DsStream st;
    st.init(ds_smem, G, (int)blockIdx.x - G.cta_begin, R, S, E);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int EQ = E >> 2;
    float4 u[NB][EV];
#pragma unroll
    for (int s = 0; s < NB; ++s)
#pragma unroll
        for (int i = 0; i < EV; ++i) {
            const int c = lane + 32 * i;
            u[s][i] = (s < G.nb && c < EQ) ? *reinterpret_cast<const float4*>(rep + (int64_t)G.sample[s] * E + 4 * c)
                                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    for (int tile = 0; tile < st.ntiles; ++tile) {
        const float4* sm = st.wait(tile);
        const int nrows = st.rows_in(tile);
        const int64_t row0 = st.r0 + (int64_t)tile * R;
        for (int rr = warp * 2; rr < nrows; rr += kDsWarps * 2) {
            const bool two = (rr + 1 < nrows);
            const float4* pa = sm + (size_t)rr * EQ;
            const float4* pb = two ? pa + EQ : pa;
            float4 a[EV], b[EV];
#pragma unroll
            for (int i = 0; i < EV; ++i) {
                const int c = lane + 32 * i;
                a[i] = (c < EQ) ? pa[c] : make_float4(0.f, 0.f, 0.f, 0.f);
                b[i] = (c < EQ) ? pb[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float acc[NB * 2];
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                float xa = 0.f, xb = 0.f;
#pragma unroll
                for (int i = 0; i < EV; ++i) {
                    xa = dot4(u[s][i], a[i], xa);
                    xb = dot4(u[s][i], b[i], xb);
                }
                acc[2 * s] = xa;
                acc[2 * s + 1] = xb;
            }
            const float tot = warp_multi_sum<NB * 2>(acc, lane);
            constexpr int kSub = 32 / (NB * 2);
            if ((lane % kSub) == 0) {
                const int j = lane / kSub, s = j >> 1, h = j & 1;
                if (s < G.nb && (h == 0 || two)) logits[(int64_t)G.sample[s] * ld + row0 + rr + h] = tot;
            }
        }
        st.release(tile);
    }
}

// ---- per-CTA partial of grad_rep[b, :] = sum_n gz[b, n] * items[n, :] -----------------------------------------------
// Same stream; the tile's <= 32 gradient values of every sample are fetched by the lanes before the wait and
// broadcast with shuffles.
template <int EV, int NB>
__global__ void __launch_bounds__(kDsThreads, 1)
ds_grad_kernel(const __grid_constant__ DsPlan plan, const float* __restrict__ gz, int64_t ldg, int E, int R, int S,
               float* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char ds_smem[];
    const DsGroup& G = plan.g[ds_find_group(plan, (int)blockIdx.x)];
    DsStream st;
    st.init(ds_smem, G, (int)blockIdx.x - G.cta_begin, R, S, E);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int EQ = E >> 2;
    float4 acc[NB][EV];
#pragma unroll
    for (int s = 0; s < NB; ++s)
#pragma unroll
        for (int i = 0; i < EV; ++i) acc[s][i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* gzrow[NB];
#pragma unroll
    for (int s = 0; s < NB; ++s) gzrow[s] = gz + (int64_t)G.sample[s < G.nb ? s : 0] * ldg;
    for (int tile = 0; tile < st.ntiles; ++tile) {
        const int nrows = st.rows_in(tile);
        const int64_t row0 = st.r0 + (int64_t)tile * R;
        float gv[NB];
#pragma unroll
        for (int s = 0; s < NB; ++s) gv[s] = (s < G.nb && lane < nrows) ? __ldg(gzrow[s] + row0 + lane) : 0.f;
        const float4* sm = st.wait(tile);
        for (int rr = warp * 2; rr < nrows; rr += kDsWarps * 2) {
            const bool two = (rr + 1 < nrows);
            const float4* pa = sm + (size_t)rr * EQ;
            const float4* pb = two ? pa + EQ : pa;
            float4 a[EV], b[EV];
#pragma unroll
            for (int i = 0; i < EV; ++i) {
                const int c = lane + 32 * i;
                a[i] = (c < EQ) ? pa[c] : make_float4(0.f, 0.f, 0.f, 0.f);
                b[i] = (c < EQ) ? pb[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                const float ga = __shfl_sync(0xffffffffu, gv[s], rr);
                const float gb0 = __shfl_sync(0xffffffffu, gv[s], (rr + 1) & 31);
                const float gb = two ? gb0 : 0.f;
#pragma unroll
                for (int i = 0; i < EV; ++i) {
                    acc[s][i].x = fmaf(ga, a[i].x, fmaf(gb, b[i].x, acc[s][i].x));
                    acc[s][i].y = fmaf(ga, a[i].y, fmaf(gb, b[i].y, acc[s][i].y));
                    acc[s][i].z = fmaf(ga, a[i].z, fmaf(gb, b[i].z, acc[s][i].z));
                    acc[s][i].w = fmaf(ga, a[i].w, fmaf(gb, b[i].w, acc[s][i].w));
                }
            }
        }
        st.release(tile);
    }
    // cross-warp sum in a fixed order, one sample slot at a time; the (idle) stage ring is the scratch
    float4* s_red = reinterpret_cast<float4*>(ds_smem + 128);  // [kDsWarps][EQ]
#pragma unroll
    for (int s = 0; s < NB; ++s) {
        if (s < G.nb) {  // uniform over the CTA
#pragma unroll
            for (int i = 0; i < EV; ++i) {
                const int c = lane + 32 * i;
                if (c < EQ) s_red[warp * EQ + c] = acc[s][i];
            }
            __syncthreads();
            for (int c = threadIdx.x; c < EQ; c += kDsThreads) {
                float4 t = s_red[c];
#pragma unroll
                for (int w = 1; w < kDsWarps; ++w) {
                    const float4 o = s_red[w * EQ + c];
                    t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
                }
                *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.x * kDsNB + s) * E + 4 * c) = t;
            }
            __syncthreads();
        }
    }
}

// grad_rep[b, :] = grad_out[b] * sum over the CTAs of b's group (ascending) of their partials
__global__ void __launch_bounds__(256)
ds_grad_finish_kernel(const __grid_constant__ DsPlan plan, const float* __restrict__ partial,
                      const float* __restrict__ grad_out, int E, float* __restrict__ grad_rep) {
    const int b = blockIdx.x;
    int gi = -1, slot = 0;
    for (int i = 0; i < plan.ngroups && gi < 0; ++i)
        for (int s = 0; s < plan.g[i].nb; ++s)
            if (plan.g[i].sample[s] == b) { gi = i; slot = s; break; }
    if (gi < 0) return;
    const DsGroup& G = plan.g[gi];
    const float scale = grad_out ? grad_out[b] : 1.0f;
    const int EQ = E >> 2;
    for (int c = threadIdx.x; c < EQ; c += blockDim.x) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < G.cta_count; ++k) {
            const float4 o = *reinterpret_cast<const float4*>(partial + ((int64_t)(G.cta_begin + k) * kDsNB + slot) * E + 4 * c);
            t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
        }
        t.x *= scale; t.y *= scale; t.z *= scale; t.w *= scale;
        *reinterpret_cast<float4*>(grad_rep + (int64_t)b * E + 4 * c) = t;
    }
}

// ---- block reductions of the loss kernel ---------------------------------------------------------------------------
struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {  // larger value, then lower index (torch.argmax: first)
    return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ ArgMax block_argmax(ArgMax x, ArgMax* s_tmp) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ArgMax o;
        o.v = __shfl_xor_sync(0xffffffffu, x.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, x.i, off);
        x = better(x, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_tmp[warp] = x;
    __syncthreads();
    ArgMax r = s_tmp[0];
    for (int w = 1; w < kDsLossThreads / 32; ++w) r = better(r, s_tmp[w]);
    return r;
}
template <int NV>
__device__ void block_sum(double (&v)[NV], double* s_tmp /* [warps][NV] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) s_tmp[warp * NV + j] = v[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double t = 0.0;
        for (int w = 0; w < kDsLossThreads / 32; ++w) t += s_tmp[w * NV + j];
        v[j] = t;
    }
}

// cluster-wide versions: block result -> one slot per CTA in shared memory -> every CTA combines the eight slots over
// distributed shared memory in rank order (deterministic)
__device__ ArgMax cluster_argmax(ArgMax x, ArgMax* s_tmp, ArgMax* s_slot, cg::cluster_group& cl) {
    x = block_argmax(x, s_tmp);
    if (threadIdx.x == 0) *s_slot = x;
    cl.sync();
    ArgMax r = *cl.map_shared_rank(s_slot, 0);
    for (int c = 1; c < kDsLossCluster; ++c) r = better(r, *cl.map_shared_rank(s_slot, c));
    cl.sync();
    return r;
}
template <int NV>
__device__ void cluster_sum(double (&v)[NV], double* s_tmp, double* s_slot /* [NV] */, cg::cluster_group& cl) {
    block_sum<NV>(v, s_tmp);
    if (threadIdx.x == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) s_slot[j] = v[j];
    cl.sync();
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double t = 0.0;
        for (int c = 0; c < kDsLossCluster; ++c) t += cl.map_shared_rank(s_slot, c)[j];
        v[j] = t;
    }
    cl.sync();
}

struct DsLossMix {  // loss = cA*CE(z, target) + cK*KD_T + cE*entropy(z) + cM*MSE + cL*ListNet_T  (or the pairwise hinge)
    float cA, cK, cE, cM, cL;
    int target_from_merged;  // CE target = argmax z (pseudo label of the merged model) instead of argmax t
    int pairwise;
    int needs_teacher;
};

static int ds_loss_mix(int type, float coef, DsLossMix* m) {
    *m = DsLossMix{0.f, 0.f, 0.f, 0.f, 0.f, 0, 0, 1};
    switch (type) {
        case MR_LOSS_CE: case MR_LOSS_SINGLE_PSEUDO_LABEL: m->cA = 1.f; break;   // loss_fn.py:36-43, 139-153
        case MR_LOSS_KD: m->cK = 1.f; break;                                      // loss_fn.py:46-59
        case MR_LOSS_MSE: m->cM = 1.f; break;                                     // loss_fn.py:180-187
        case MR_LOSS_ADAMERGING: m->cE = 1.f; m->needs_teacher = 0; break;        // loss_fn.py:62-68
        case MR_LOSS_ADAMERGING_KD: m->cE = 1.f; m->cK = coef; break;             // loss_fn.py:71-87
        case MR_LOSS_MERGED_PSEUDO_LABEL: m->cA = 1.f; m->target_from_merged = 1; m->needs_teacher = 0; break;  // :90-105
        case MR_LOSS_MERGED_PSEUDO_LABEL_KD: m->cA = 1.f; m->target_from_merged = 1; m->cK = coef; break;       // :108-129
        case MR_LOSS_SINGLE_PSEUDO_LABEL_KD: m->cA = 1.f; m->cK = coef; break;    // loss_fn.py:156-177
        case MR_LOSS_PAIRWISE: m->pairwise = 1; break;                            // loss_fn.py:190-211
        case MR_LOSS_LISTNET: m->cL = 1.f; break;                                 // loss_fn.py:214-229
        default: set_error("mr_distill_loss: unknown loss type %d", type); return MR_ERR_INVALID_ARG;
    }
    return MR_OK;
}

// One thread-block CLUSTER of eight CTAs per sample (a single CTA kept one SM busy for ~50 us on a 25,000-item row; the
// expf / logf / fp64 work is what costs, not the bytes).  z = merged-model logits, t = teacher logits (n each).
// Streaming passes over the two rows (just written / L2-resident): maxima, sums, optional entropy pass, gradient
// write; between passes the CTAs exchange their partials through distributed shared memory.
__global__ void __cluster_dims__(kDsLossCluster, 1, 1) __launch_bounds__(kDsLossThreads)
ds_loss_kernel(const __grid_constant__ DsLossArgs args, const float* __restrict__ logits, int64_t ld, DsLossMix mix,
               float T, float margin, float* __restrict__ loss, float* __restrict__ gz, int64_t ldg) {
    __shared__ ArgMax s_am[kDsLossThreads / 32];
    __shared__ double s_sum[(kDsLossThreads / 32) * 5];
    __shared__ ArgMax s_am_slot;
    __shared__ double s_sum_slot[5];
    cg::cluster_group cl = cg::this_cluster();
    const int b = blockIdx.y;
    const int gt = (int)blockIdx.x * kDsLossThreads + (int)threadIdx.x;   // thread index inside the cluster
    constexpr int GS = kDsLossCluster * kDsLossThreads;
    const bool writer = (blockIdx.x == 0 && threadIdx.x == 0);
    const int n = args.n[b];
    const float* __restrict__ z = logits + (int64_t)b * ld;
    const float* __restrict__ t = args.teacher[b];
    float* __restrict__ g = gz ? gz + (int64_t)b * ldg : nullptr;
    const bool has_t = (t != nullptr);
    if (n <= 0) { if (writer) loss[b] = 0.f; return; }

    ArgMax az{-INFINITY, 0x7fffffff}, at{-INFINITY, 0x7fffffff};
#pragma unroll 4
    for (int i = gt; i < n; i += GS) {
        az = better(az, ArgMax{z[i], i});
        if (has_t) at = better(at, ArgMax{t[i], i});
    }
    az = cluster_argmax(az, s_am, &s_am_slot, cl);
    if (has_t) at = cluster_argmax(at, s_am, &s_am_slot, cl);

    if (mix.pairwise) {  // hinge on (best, second-best) teacher items
        ArgMax an{-INFINITY, 0x7fffffff};
    #pragma unroll 4
    for (int i = gt; i < n; i += GS)
            if (i != at.i) an = better(an, ArgMax{t[i], i});
        an = cluster_argmax(an, s_am, &s_am_slot, cl);
        const int neg = (an.i == 0x7fffffff) ? 0 : an.i;  // n == 1: argmax of an all -inf row is index 0
        const float h = margin - (z[at.i] - z[neg]);
        const bool active = h > 0.f;
        if (g) {
        #pragma unroll 4
    for (int i = gt; i < n; i += GS) {
                float v = 0.f;
                if (active) v = (i == at.i ? -1.f : 0.f) + (i == neg ? 1.f : 0.f);
                g[i] = v;
            }
        }
        if (writer) loss[b] = active ? h : 0.f;
        return;
    }

    const float invT = 1.0f / T;
    const bool useT = (mix.cK != 0.f || mix.cL != 0.f);
    // sums: [0] S1 = sum exp(z-mz); [1] ST = sum exp((z-mz)/T); [2] PT = sum e_t; [3] sum e_t ((t-mt)-(z-mz)); [4] sum (z-t)^2
    // (for ListNet [3] holds sum e_t (z-mz) instead)
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int i = gt; i < n; i += GS) {
        const float zc = z[i] - az.v;
        acc[0] += (double)expf(zc);
        if (useT) {
            acc[1] += (double)expf(zc * invT);
            const float tc = t[i] - at.v;
            const float et = expf(tc * invT);
            acc[2] += (double)et;
            acc[3] += (double)et * (double)(mix.cL != 0.f ? zc : tc - zc);
        }
        if (mix.cM != 0.f) {
            const float dlt = z[i] - t[i];
            acc[4] += (double)dlt * (double)dlt;
        }
    }
    cluster_sum<5>(acc, s_sum, s_sum_slot, cl);
    const double S1 = acc[0], ST = acc[1], PT = acc[2];
    double total = 0.0;
    const int target = mix.target_from_merged ? az.i : at.i;
    if (mix.cA != 0.f) total += (double)mix.cA * (log(S1) - (double)(z[target] - az.v));
    if (mix.cK != 0.f) total += (double)mix.cK * (double)T * (double)T * (acc[3] / ((double)T * PT) - log(PT) + log(ST));
    if (mix.cL != 0.f) total += (double)mix.cL * (log(ST) - acc[3] / ((double)T * PT));
    if (mix.cM != 0.f) total += (double)mix.cM * acc[4] / (double)n;

    const float inv_S1 = (float)(1.0 / S1), inv_ST = useT ? (float)(1.0 / ST) : 0.f, inv_PT = useT ? (float)(1.0 / PT) : 0.f;
    double hbar = 0.0;
    if (mix.cE != 0.f) {  // entropy of softmax(z) with the reference's +1e-8 inside the log (loss_fn.py:65-66)
        double e2[2] = {0.0, 0.0};
    #pragma unroll 4
    for (int i = gt; i < n; i += GS) {
            const float p = expf(z[i] - az.v) * inv_S1;
            const float lp = logf(p + 1e-8f);
            e2[0] += (double)p * (double)lp;
            e2[1] += (double)p * (double)(lp + p / (p + 1e-8f));
        }
        cluster_sum<2>(e2, s_sum, s_sum_slot, cl);
        total += (double)mix.cE * (-e2[0]);
        hbar = -e2[1];
    }
    if (writer) loss[b] = (float)total;

    if (g) {
        const float kd_scale = mix.cK * T + mix.cL * invT;  // d/dz of T^2 KL(P || Q_T) is T (Q - P); ListNet: (Q - P)/T
        const float mse_scale = mix.cM * 2.0f / (float)n;
    #pragma unroll 4
    for (int i = gt; i < n; i += GS) {
            const float zc = z[i] - az.v;
            float v = 0.f;
            if (mix.cA != 0.f || mix.cE != 0.f) {
                const float p = expf(zc) * inv_S1;
                if (mix.cA != 0.f) v += mix.cA * (p - (i == target ? 1.f : 0.f));
                if (mix.cE != 0.f) {
                    const float h = -(logf(p + 1e-8f) + p / (p + 1e-8f));
                    v += mix.cE * p * (h - (float)hbar);
                }
            }
            if (useT) {
                const float q = expf(zc * invT) * inv_ST;
                const float pt = expf((t[i] - at.v) * invT) * inv_PT;
                v += kd_scale * (q - pt);
            }
            if (mix.cM != 0.f) v += mse_scale * (z[i] - t[i]);
            g[i] = v;
        }
    }
}

// x[r, :] /= ||x[r, :]||_2   (merge_train.py:122-123: `e / e.norm(dim=-1, keepdim=True)`; a zero row gives NaN there
// and here).  One warp per row, fp32 sum of squares in lane order then a shuffle tree.
__global__ void __launch_bounds__(256)
ds_normalize_rows_kernel(const float* __restrict__ x, int64_t rows, int E, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float* p = x + r * E;
        float ss = 0.f;
        for (int e = lane; e < E; e += 32) ss = fmaf(p[e], p[e], ss);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        const float nrm = sqrtf(ss);
        for (int e = lane; e < E; e += 32) out[r * E + e] = p[e] / nrm;
    }
}

#define MR_DISPATCH_EV_NB(E, NBV, ...)                                                             \
    do {                                                                                           \
        const int ev__ = ((E) + 127) / 128;                                                        \
        const int nb__ = (NBV) <= 1 ? 1 : ((NBV) <= 2 ? 2 : 4);                                    \
        if (ev__ <= 1) { constexpr int EVV = 1; MR_DISPATCH_NB_(nb__, __VA_ARGS__); }              \
        else if (ev__ <= 2) { constexpr int EVV = 2; MR_DISPATCH_NB_(nb__, __VA_ARGS__); }         \
        else if (ev__ <= 4) { constexpr int EVV = 4; MR_DISPATCH_NB_(nb__, __VA_ARGS__); }         \
        else if (ev__ <= 6) { constexpr int EVV = 6; MR_DISPATCH_NB_(nb__, __VA_ARGS__); }         \
        else { constexpr int EVV = 8; MR_DISPATCH_NB_(nb__, __VA_ARGS__); }                        \
    } while (0)
#define MR_DISPATCH_NB_(nb, ...)                                          \
    do {                                                                  \
        if ((nb) == 1) { constexpr int NBB = 1; __VA_ARGS__; }            \
        else if ((nb) == 2) { constexpr int NBB = 2; __VA_ARGS__; }       \
        else { constexpr int NBB = 4; __VA_ARGS__; }                      \
    } while (0)

static int ds_max_nb(const DsPlan& plan) {
    int m = 1;
    for (int i = 0; i < plan.ngroups; ++i) m = plan.g[i].nb > m ? plan.g[i].nb : m;
    return m;
}

}  // namespace mr

extern "C" int mr_distill_logits(const float* rep, int B, int E, const float* const* item_ptrs, const int64_t* item_rows,
                                 int nD, const int32_t* sample_domain, float* logits, int64_t ld, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(rep && logits, "mr_distill_logits: null pointer");
    MR_REQUIRE(E >= 4 && E <= MR_DISTILL_MAX_E && (E & 3) == 0, "mr_distill_logits: E=%d must be a multiple of 4 in [4,%d]",
               E, MR_DISTILL_MAX_E);
    MR_REQUIRE(host_aligned16(rep), "mr_distill_logits: rep is not 16-byte aligned");
    DsPlan plan;
    const int rc = ds_build_plan(item_ptrs, item_rows, nD, sample_domain, B, &plan);
    if (rc != MR_OK) return rc;
    for (int i = 0; i < plan.ngroups; ++i)
        MR_REQUIRE(plan.g[i].rows <= ld, "mr_distill_logits: ld=%lld is smaller than a table's %d rows", (long long)ld,
                   plan.g[i].rows);
    if (plan.total_ctas == 0) return MR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const DsTiling tl = ds_tiling(E);
    MR_DISPATCH_EV_NB(E, ds_max_nb(plan), {
        cudaFuncSetAttribute(ds_logits_kernel<EVV, NBB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem);
        ds_logits_kernel<EVV, NBB><<<plan.total_ctas, kDsThreads, tl.smem, st>>>(plan, rep, E, tl.R, tl.S, logits, ld);
    });
    MR_CUDA_LAUNCH_CHECK("mr_distill_logits");
    return MR_OK;
}

extern "C" int mr_distill_loss(const float* logits, int64_t ld, const float* const* teacher_rows, const int64_t* n_per_sample,
                               int B, int loss_type, float temperature, float coefficient, float margin, float* loss,
                               float* grad_logits, int64_t ldg, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(logits && n_per_sample && loss, "mr_distill_loss: null pointer");
    MR_REQUIRE(B >= 1 && B <= MR_DISTILL_MAX_B, "mr_distill_loss: B=%d outside [1,%d]", B, MR_DISTILL_MAX_B);
    DsLossMix mix;
    const int rc = ds_loss_mix(loss_type, coefficient, &mix);
    if (rc != MR_OK) return rc;
    MR_REQUIRE(!mix.needs_teacher || teacher_rows, "mr_distill_loss: loss type %d needs teacher rows", loss_type);
    const bool usesT = (mix.cK != 0.f || mix.cL != 0.f);
    MR_REQUIRE(!usesT || temperature > 0.f, "mr_distill_loss: temperature must be > 0");
    DsLossArgs args;
    for (int b = 0; b < B; ++b) {
        MR_REQUIRE(n_per_sample[b] >= 0 && n_per_sample[b] <= ld && n_per_sample[b] <= INT32_MAX,
                   "mr_distill_loss: sample %d has %lld logits (ld=%lld)", b, (long long)n_per_sample[b], (long long)ld);
        MR_REQUIRE(!grad_logits || n_per_sample[b] <= ldg, "mr_distill_loss: ldg too small");
        args.n[b] = (int32_t)n_per_sample[b];
        args.teacher[b] = teacher_rows ? teacher_rows[b] : nullptr;
        MR_REQUIRE(!mix.needs_teacher || args.teacher[b], "mr_distill_loss: teacher row %d is NULL", b);
    }
    ds_loss_kernel<<<dim3(kDsLossCluster, B), kDsLossThreads, 0, (cudaStream_t)stream>>>(args, logits, ld, mix, usesT ? temperature : 1.0f, margin,
                                                                  loss, grad_logits, ldg);
    MR_CUDA_LAUNCH_CHECK("mr_distill_loss");
    return MR_OK;
}

extern "C" int64_t mr_distill_grad_workspace_bytes(int E) {
    if (E < 1) return 0;
    return mr::ds_max_ctas() * mr::kDsNB * (int64_t)E * (int64_t)sizeof(float);
}

extern "C" int mr_distill_grad(const float* grad_logits, int64_t ldg, const float* grad_out, int B, int E,
                               const float* const* item_ptrs, const int64_t* item_rows, int nD, const int32_t* sample_domain,
                               float* grad_rep, void* ws, int64_t ws_bytes, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(grad_logits && grad_rep && ws, "mr_distill_grad: null pointer");
    MR_REQUIRE(E >= 4 && E <= MR_DISTILL_MAX_E && (E & 3) == 0, "mr_distill_grad: E=%d must be a multiple of 4 in [4,%d]", E,
               MR_DISTILL_MAX_E);
    MR_REQUIRE(host_aligned16(grad_rep) && host_aligned16(ws), "mr_distill_grad: grad_rep / ws not 16-byte aligned");
    if (ws_bytes < mr_distill_grad_workspace_bytes(E)) {
        set_error("mr_distill_grad: workspace too small (%lld < %lld bytes)", (long long)ws_bytes,
                  (long long)mr_distill_grad_workspace_bytes(E));
        return MR_ERR_WORKSPACE;
    }
    DsPlan plan;
    const int rc = ds_build_plan(item_ptrs, item_rows, nD, sample_domain, B, &plan);
    if (rc != MR_OK) return rc;
    for (int i = 0; i < plan.ngroups; ++i)
        MR_REQUIRE(plan.g[i].rows <= ldg, "mr_distill_grad: ldg=%lld is smaller than a table's %d rows", (long long)ldg,
                   plan.g[i].rows);
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(ws);
    const DsTiling tl = ds_tiling(E);
    MR_DISPATCH_EV_NB(E, ds_max_nb(plan), {
        cudaFuncSetAttribute(ds_grad_kernel<EVV, NBB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.smem);
        ds_grad_kernel<EVV, NBB><<<plan.total_ctas, kDsThreads, tl.smem, st>>>(plan, grad_logits, ldg, E, tl.R, tl.S, partial);
    });
    ds_grad_finish_kernel<<<B, 256, 0, st>>>(plan, partial, grad_out, E, grad_rep);
    MR_CUDA_LAUNCH_CHECK("mr_distill_grad");
    return MR_OK;
}

extern "C" int mr_normalize_rows(const float* x, int64_t rows, int E, float* out, mr_stream_t stream) {
    using namespace mr;
    MR_REQUIRE(rows >= 0 && E >= 1, "mr_normalize_rows: need rows >= 0, E >= 1");
    if (rows == 0) return MR_OK;
    MR_REQUIRE(x && out, "mr_normalize_rows: null pointer");
    int64_t blocks = (rows + 7) / 8;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    ds_normalize_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, rows, E, out);
    MR_CUDA_LAUNCH_CHECK("mr_normalize_rows");
    return MR_OK;
}
