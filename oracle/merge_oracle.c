/*
 * oracle/merge_oracle.c -- CPU restatement of MergeRec's merger + evaluator arithmetic.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (mergerec_b200/) never links, imports or calls anything in this directory.
 *
 * Every function restates, in plain C with unfused IEEE fp32 arithmetic
 * (build with -ffp-contract=off, no -ffast-math), what the reference computes through
 * PyTorch CPU kernels.  File:line citations are relative to /root/reference.
 *
 * Parity pinning: the reference has no tests / golden vectors of its own,
 * so this oracle is pinned against outputs of the *unmodified reference modules run in the
 * build container* (tests/golden/make_golden.py -> the .npz files under tests/golden/) and, when
 * /root/reference is importable, against the live reference (tests/test_oracle_vs_reference.py).
 *
 * Third-party arithmetic restated here (the reference delegates to torch CPU kernels,
 * requirements.txt:1 pins torch~=2.6, installed 2.11.0):
 *   - torch.sum(dim=0) over a contiguous (K, n) fp32 block, K <= 16: sequential in k for
 *     columns j < 32*floor(n/32); 4-way interleaved partial sums for the trailing n mod 32
 *     columns (identical to sequential when K <= 4).  Blocks with n < 8 take a different
 *     TensorIterator path that is NOT restated (no tensor in the named architectures is
 *     that small); K >= 17 switches to cascade summation and is out of contract.
 *   - torch.topk tie order is implementation-defined; this oracle uses the canonical rule
 *     "larger value first, then lower index first" (SURVEY.md section 0.1 D3).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;

#define ORC_MAX_K 16

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline u32 f2u(float f) { u32 u; memcpy(&u, &f, 4); return u; }

/* ---- torch.sum(dim=0) summation order (see header) -------------------------------- */
static inline float sum_seq(const float* p, int K) {
    float s = 0.0f;
    for (int k = 0; k < K; ++k) s = s + p[k];
    return s;
}
static inline float sum_inter4(const float* p, int K) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    int full = K / 4;
    for (int i = 0; i < full; ++i) {
        a0 = a0 + p[4 * i + 0];
        a1 = a1 + p[4 * i + 1];
        a2 = a2 + p[4 * i + 2];
        a3 = a3 + p[4 * i + 3];
    }
    for (int r = 4 * full; r < K; ++r) a0 = a0 + p[r];
    return ((a0 + a1) + a2) + a3;
}
/* column j (0-based inside its block) of a contiguous block with n columns */
static inline float torch_sum_dim0(const float* p, int K, i64 j, i64 n) {
    if (K >= 5 && j >= (n & ~(i64)31)) return sum_inter4(p, K);
    return sum_seq(p, K);
}

/* ---- A2: get_task_vectors  (merger/algorithms/task_vector.py:8-10) ----------------- */
void orc_task_vectors(const float* base, const float* const* models, int K, i64 d, float* T) {
    for (int k = 0; k < K; ++k) {
        const float* m = models[k];
        float* t = T + (i64)k * d;
#pragma omp parallel for schedule(static)
        for (i64 j = 0; j < d; ++j) t[j] = m[j] - base[j];
    }
}

/* ---- A1: merge_task_vector  (merger/algorithms/task_vector.py:13-34) ---------------
 * merged = base.clone(); for k: merged += w_k * (m_k - base)   (w_k rounded to fp32) */
void orc_merge_task_vector(const float* base, const float* const* models, int K, i64 d,
                           const float* w, float* out) {
#pragma omp parallel for schedule(static)
    for (i64 j = 0; j < d; ++j) {
        float b = base[j];
        float acc = b;
        for (int k = 0; k < K; ++k) {
            float u = models[k][j] - b;
            float p = w[k] * u;
            acc = acc + p;
        }
        out[j] = acc;
    }
}

/* ---- A10: merge_linear  (merger/algorithms/linear.py:8-27) ------------------------- */
void orc_merge_linear(const float* const* models, int K, i64 d, const float* w, float* out) {
#pragma omp parallel for schedule(static)
    for (i64 j = 0; j < d; ++j) {
        float acc = 0.0f;
        for (int k = 0; k < K; ++k) {
            float p = w[k] * models[k][j];
            acc = acc + p;
        }
        out[j] = acc;
    }
}

/* ---- A3 / A4: lambda-merge from stored task vectors ---------------------------------
 * task-wise  (weight_learning/module/task_wise.py:36-48): one block [0,d), one group;
 * layer-wise (weight_learning/module/layer_wise.py:64-83): one block per state_dict tensor,
 *            group = encoder layer index or "others" (layer_wise.py:13-33).
 * merged[j] = base[j] + sum_dim0_k( w[g,k] * T[k,j] ), summed in torch.sum(dim=0) order with
 * the block being [seg_begin[p], seg_end[p]).  w is (G,K) row-major. */
void orc_lambda_merge(const float* base, const float* T, i64 ldT, int K, i64 d, const float* w,
                      const i64* seg_begin, const i64* seg_end, const int32_t* seg_group, int P,
                      float* out) {
    (void)d;
    for (int p = 0; p < P; ++p) {
        i64 s = seg_begin[p], e = seg_end[p], n = e - s;
        const float* wg = w + (i64)seg_group[p] * K;
#pragma omp parallel for schedule(static) if (n > 65536)
        for (i64 j = s; j < e; ++j) {
            float prod[ORC_MAX_K];
            for (int k = 0; k < K; ++k) prod[k] = wg[k] * T[(i64)k * ldT + j];
            float sum = torch_sum_dim0(prod, K, j - s, n);
            out[j] = base[j] + sum;
        }
    }
}

/* ---- A5: lambda-gradient  (autograd of A3/A4; weight_learning/utils.py:11-15,43-51) --
 * out[g,k] = sum_{p in g} sum_j grad[j] * T[k,j], accumulated in fp64 (reference truth for the
 * tolerance test; torch's own fp32 cascade sum matches this to ~1e-7). out is (G,K) doubles. */
void orc_lambda_grad(const float* grad, const float* T, i64 ldT, int K, const i64* seg_begin,
                     const i64* seg_end, const int32_t* seg_group, int P, int G, double* out) {
    for (i64 i = 0; i < (i64)G * K; ++i) out[i] = 0.0;
    for (int p = 0; p < P; ++p) {
        i64 s = seg_begin[p], e = seg_end[p];
        int g = seg_group[p];
        for (int k = 0; k < K; ++k) {
            const float* t = T + (i64)k * ldT;
            double acc = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : acc) if (e - s > 65536)
            for (i64 j = s; j < e; ++j) acc += (double)grad[j] * (double)t[j];
            out[(i64)g * K + k] += acc;
        }
    }
}

/* ---- radix select: k-th largest (1-based) of n distinct-or-not u64 keys; destroys a --- */
static u64 select_kth_largest(u64* a, i64 n, i64 k) {
    u64 prefix = 0;
    i64* hist = (i64*)malloc(65536 * sizeof(i64));
    for (int shift = 48; shift >= 0; shift -= 16) {
        memset(hist, 0, 65536 * sizeof(i64));
#pragma omp parallel
        {
            i64* loc = (i64*)calloc(65536, sizeof(i64));
#pragma omp for schedule(static) nowait
            for (i64 i = 0; i < n; ++i) loc[(a[i] >> shift) & 0xFFFF]++;
#pragma omp critical
            for (int b = 0; b < 65536; ++b) hist[b] += loc[b];
            free(loc);
        }
        i64 above = 0;
        int dg = 65535;
        for (; dg >= 0; --dg) {
            if (above + hist[dg] >= k) break;
            above += hist[dg];
        }
        k -= above;
        prefix |= ((u64)dg) << shift;
        i64 m = 0;
        for (i64 i = 0; i < n; ++i)
            if (((a[i] >> shift) & 0xFFFF) == (u64)dg) a[m++] = a[i];
        n = m;
    }
    free(hist);
    return prefix;
}

/* composite selection key: larger |u| first, then LOWER flat index first. d must be <= 2^32. */
static inline u64 ties_key(float u, i64 j) {
    return ((u64)(f2u(u) & 0x7FFFFFFFu) << 32) | (u64)(0xFFFFFFFFu - (u32)j);
}

/* ---- A6 (selection part): _compute_sparse_updates  (merger/algorithms/ties.py:8-28) ---
 * u = m_k - base (times w_k when w != NULL, ties.py:20-21); keep the k_cnt = int(density*d)
 * largest |u| over the WHOLE flat vector (ties.py:14-15,23).  Canonical tie rule: lowest index.
 * Returns per model the 64-bit cut: element j is kept iff ties_key(u_j, j) >= cut[k]. */
void orc_ties_select(const float* base, const float* const* models, int K, i64 d, const float* w,
                     i64 k_cnt, u64* cut) {
    if (k_cnt <= 0) {
        for (int k = 0; k < K; ++k) cut[k] = ~(u64)0;
        return;
    }
    if (k_cnt >= d) {
        for (int k = 0; k < K; ++k) cut[k] = 0;
        return;
    }
    u64* keys = (u64*)malloc((size_t)d * sizeof(u64));
    for (int k = 0; k < K; ++k) {
        const float* m = models[k];
#pragma omp parallel for schedule(static)
        for (i64 j = 0; j < d; ++j) {
            float u = m[j] - base[j];
            if (w) u = u * w[k];
            keys[j] = ties_key(u, j);
        }
        cut[k] = select_kth_largest(keys, d, k_cnt);
    }
    free(keys);
}

static inline float trimmed_update(const float* base, const float* const* models, const float* w,
                                   const u64* cut, int k, i64 j) {
    float u = models[k][j] - base[j];
    if (w) u = u * w[k];
    return ties_key(u, j) >= cut[k] ? u : 0.0f;
}

/* ---- A6-A8: get_ties_vectors  (merger/algorithms/ties.py:55-72; sign rule :31-52) ------
 * That is (K,d) row-major with leading dimension d.  trim_mask / elect_mask (optional, may be
 * NULL) receive one byte per (k,j): trim_mask = survived the magnitude trim; elect_mask =
 * survived trim AND sign election (i.e. That[k,j] != 0). */
void orc_ties_vectors(const float* base, const float* const* models, int K, i64 d, const u64* cut,
                      float* That, uint8_t* trim_mask, uint8_t* elect_mask) {
#pragma omp parallel for schedule(static)
    for (i64 j = 0; j < d; ++j) {
        float s[ORC_MAX_K], pp[ORC_MAX_K], nn[ORC_MAX_K];
        for (int k = 0; k < K; ++k) {
            s[k] = trimmed_update(base, models, NULL, cut, k, j);
            pp[k] = s[k] > 0.0f ? s[k] : 0.0f;
            nn[k] = s[k] < 0.0f ? s[k] : 0.0f;
        }
        float pos = torch_sum_dim0(pp, K, j, d); /* ties.py:35 */
        float neg = torch_sum_dim0(nn, K, j, d); /* ties.py:36 */
        float sign;
        if (pos != 0.0f && neg != 0.0f) { /* ties.py:38-45 */
            sign = fabsf(pos) >= fabsf(neg) ? 1.0f : -1.0f;
        } else { /* ties.py:47-48 */
            float t = pos + neg;
            sign = t > 0.0f ? 1.0f : (t < 0.0f ? -1.0f : 0.0f);
        }
        if (sign == 0.0f) sign = 1.0f; /* ties.py:50 */
        float sel[ORC_MAX_K];
        int cnt = 0;
        for (int k = 0; k < K; ++k) { /* ties.py:61-65 */
            sel[k] = sign > 0.0f ? pp[k] : nn[k];
            if (sel[k] != 0.0f) cnt++;
        }
        for (int k = 0; k < K; ++k) { /* ties.py:68-70: x / float(cnt); 0/0 -> NaN -> 0 */
            float v = cnt ? sel[k] / (float)cnt : 0.0f;
            That[(i64)k * d + j] = v;
            if (trim_mask) trim_mask[(i64)k * d + j] = (ties_key(models[k][j] - base[j], j) >= cut[k]);
            if (elect_mask) elect_mask[(i64)k * d + j] = (sel[k] != 0.0f);
        }
    }
}

/* ---- A9: merge_ties  (merger/algorithms/ties.py:75-83) ---------------------------------
 * base + sum_dim0_k trim_k( w_k * (m_k - base) ); no sign election, no mean. cut must come from
 * orc_ties_select called with the same w. */
void orc_merge_ties(const float* base, const float* const* models, int K, i64 d, const float* w,
                    const u64* cut, float* out) {
#pragma omp parallel for schedule(static)
    for (i64 j = 0; j < d; ++j) {
        float s[ORC_MAX_K];
        for (int k = 0; k < K; ++k) s[k] = trimmed_update(base, models, w, cut, k, j);
        float delta = torch_sum_dim0(s, K, j, d); /* ties.py:81 */
        out[j] = base[j] + delta;                 /* ties.py:83 */
    }
}

/* ---- B2: top-K per row  (evaluator/evaluator.py:43) ------------------------------------
 * canonical order: score descending, then item id ascending; -0.0 == +0.0; NaN sorts first
 * (torch.topk treats NaN as the greatest value). ids are int32 (id_base + column). */
static inline u32 score_key(float f) {
    if (f != f) return 0xFFFFFFFFu;
    f = f + 0.0f; /* -0 -> +0 */
    u32 u = f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline u64 topk_key(float score, u32 id) {
    return ((u64)score_key(score) << 32) | (u64)(0xFFFFFFFFu - id);
}
static void heap_sift_down(u64* h, int n, int i) { /* min-heap */
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && h[l] < h[m]) m = l;
        if (r < n && h[r] < h[m]) m = r;
        if (m == i) return;
        u64 t = h[i]; h[i] = h[m]; h[m] = t;
        i = m;
    }
}
static int cmp_u64_desc(const void* a, const void* b) {
    u64 x = *(const u64*)a, y = *(const u64*)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}
/* scores (Q, N) with row stride ld; out_val/out_id (Q, K). Requires K <= N. */
void orc_topk_rows(const float* scores, i64 Q, i64 N, i64 ld, int K, int32_t id_base,
                   float* out_val, int32_t* out_id) {
#pragma omp parallel
    {
        u64* h = (u64*)malloc((size_t)K * sizeof(u64));
#pragma omp for schedule(static)
        for (i64 q = 0; q < Q; ++q) {
            const float* row = scores + q * ld;
            int n = 0;
            for (i64 c = 0; c < N; ++c) {
                u64 key = topk_key(row[c], (u32)(id_base + (int32_t)c));
                if (n < K) {
                    h[n++] = key;
                    if (n == K)
                        for (int i = K / 2 - 1; i >= 0; --i) heap_sift_down(h, K, i);
                } else if (key > h[0]) {
                    h[0] = key;
                    heap_sift_down(h, K, 0);
                }
            }
            qsort(h, (size_t)n, sizeof(u64), cmp_u64_desc);
            for (int i = 0; i < n; ++i) {
                u32 id = 0xFFFFFFFFu - (u32)(h[i] & 0xFFFFFFFFu);
                out_id[q * K + i] = (int32_t)id;
                out_val[q * K + i] = row[(i64)((int32_t)id - id_base)];
            }
        }
        free(h);
    }
}

/* merge G per-shard top-K lists (G, Q, K) into one (Q, K) under the same order (the multi-GPU
 * exchange step; no reference counterpart -- the reference is single-GPU, README.md:51-53). */
typedef struct { u64 key; float val; } orc_kv;
static int cmp_kv_desc(const void* a, const void* b) {
    u64 x = ((const orc_kv*)a)->key, y = ((const orc_kv*)b)->key;
    return x < y ? 1 : (x > y ? -1 : 0);
}
void orc_topk_merge(const float* vals, const int32_t* ids, int G, i64 Q, int K, float* out_val,
                    int32_t* out_id) {
#pragma omp parallel
    {
        orc_kv* kv = (orc_kv*)malloc((size_t)G * K * sizeof(orc_kv));
#pragma omp for schedule(static)
        for (i64 q = 0; q < Q; ++q) {
            for (int g = 0; g < G; ++g)
                for (int i = 0; i < K; ++i) {
                    i64 src = ((i64)g * Q + q) * K + i;
                    kv[g * K + i].key = topk_key(vals[src], (u32)ids[src]);
                    kv[g * K + i].val = vals[src];
                }
            qsort(kv, (size_t)G * K, sizeof(orc_kv), cmp_kv_desc);
            for (int i = 0; i < K; ++i) {
                out_id[q * K + i] = (int32_t)(0xFFFFFFFFu - (u32)(kv[i].key & 0xFFFFFFFFu));
                out_val[q * K + i] = kv[i].val;
            }
        }
        free(kv);
    }
}

/* ---- B3/B4 helper: rank of the label inside each row's top-K list, -1 when absent -------
 * (evaluator/metrics.py:51-59 `true in pred`, :79-84 `pred.index(true)`) */
void orc_label_rank(const int32_t* ids, i64 Q, int K, const i64* labels, int32_t* rank) {
    for (i64 q = 0; q < Q; ++q) {
        int32_t r = -1;
        for (int i = 0; i < K; ++i)
            if ((i64)ids[q * K + i] == labels[q]) { r = i; break; }
        rank[q] = r;
    }
}

/* ---- B1: scores = U @ I.T in fp32 (module/recommender/module.py:137) --------------------
 * plain k-sequential fp32 accumulation; the reference's MKL sgemm bits are not reproducible,
 * so parity tests use exact-grid inputs (every order gives the same bits) or fp64 near-tie checks. */
void orc_scores(const float* U, const float* I, i64 Q, i64 N, int E, float* out) {
#pragma omp parallel for schedule(static)
    for (i64 q = 0; q < Q; ++q)
        for (i64 n = 0; n < N; ++n) {
            const float* u = U + q * E;
            const float* it = I + n * E;
            float acc = 0.0f;
            for (int e = 0; e < E; ++e) acc = acc + u[e] * it[e];
            out[q * N + n] = acc;
        }
}
