/*
 * mergerec_b200.h -- C ABI of libmergerec_b200.so: the B200 (sm_100a) replacement for MergeRec's
 * two data-parallel hot paths (merger, evaluator).  Plain pointers and sizes only; no torch types.
 *
 * Conventions:
 *   - every pointer marked "dev" is a CUDA device pointer owned by the caller (PyTorch allocates all
 *     inputs, outputs and workspaces; the library owns no device memory);
 *   - every pointer marked "host" is ordinary host memory read during the call;
 *   - all entry points are stream-ordered and asynchronous: they enqueue work on `stream` and return;
 *     nothing synchronises the host;
 *   - functions never throw: the return value is 0 on success, <0 for an argument error
 *     (mr_status), >0 for a cudaError_t.  mr_last_error() gives a thread-local message.
 *   - "flat" vectors are the reference's flattened state_dicts: fp32, length d, tensors concatenated
 *     in dict order (reference: rec_retrieval/merger/utils/model_operations.py:47-63).
 *
 * Citations "ref:" are file:line under the reference checkout (DIALLab-SKKU/MergeRec).
 */
#ifndef MERGEREC_B200_H_
#define MERGEREC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* mr_stream_t; /* == cudaStream_t */

enum mr_status {
    MR_OK = 0,
    MR_ERR_INVALID_ARG = -1,
    MR_ERR_UNSUPPORTED = -2,
    MR_ERR_WORKSPACE = -3,
};

#define MR_MAX_K 16 /* bit-exact torch.sum(dim=0) order is only defined for K <= 16 */

int mr_version(void);
const char* mr_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Merger: elementwise merges over flat vectors (HBM-bound streaming kernels).
 * ---------------------------------------------------------------------------------------------- */

/* T[k, j] = models[k][j] - base[j]                       ref: merger/algorithms/task_vector.py:8-10
 * models: host array of K dev pointers (d floats each). T: dev, K rows with leading dimension ldT. */
int mr_task_vectors(const float* base, const float* const* models, int K, int64_t d, float* T,
                    int64_t ldT, mr_stream_t stream);

enum mr_merge_order {
    /* acc = base; acc = acc + fl(w_k * u_k)  k = 0..K-1    ref: merger/algorithms/task_vector.py:30-32 */
    MR_ORDER_BASE_FIRST = 0,
    /* out = base + sum_dim0_k fl(w_k * u_k) in torch.sum(dim=0) order (sequential in k; 4-way
     * interleaved on the trailing n mod 32 columns of each block when K >= 5)
     *                                                       ref: weight_learning/module/task_wise.py:43-47,
     *                                                            weight_learning/module/layer_wise.py:76-82 */
    MR_ORDER_SUM_FIRST = 1,
    /* acc = 0; acc = acc + fl(w_k * src_k); base unused      ref: merger/algorithms/linear.py:23-25 */
    MR_ORDER_LINEAR = 2,
};

/* The lambda-weighted merge (A1, A3, A4, A10 of SURVEY.md section 8(a)).
 *   src           host array of K dev pointers: fine-tuned flat models (src_is_model = 1, u_k = src_k - base
 *                 computed on the fly) or task-vector rows (src_is_model = 0, u_k = src_k).
 *   w             dev, (G, K) row-major fp32 lambdas; row g applies to blocks with seg_group == g.
 *   seg_end       dev, P ascending exclusive block ends (seg_end[P-1] == d), or NULL when P == 1
 *                 (one block [0, d), group 0).  Blocks are the reference's sum(dim=0) operands:
 *                 the whole vector for task-wise, one state_dict tensor each for layer-wise.
 *   seg_group     dev, P group ids in [0, G), or NULL when P == 1.
 *   out           dev, d floats.  May alias base. */
int mr_merge_axpy(const float* base, const float* const* src, int K, int64_t d, const float* w, int G,
                  const int64_t* seg_end, const int32_t* seg_group, int P, int order, int src_is_model,
                  float* out, mr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MERGEREC_B200_H_ */
