"""numpy/ctypes front-end of the CPU oracle (``merge_oracle.c``) -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path never routes through it.

Each function restates one reference function on numpy arrays; citations are relative to
``/root/reference``.  Parity pinning: the reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified reference run in
the build container (``tests/golden/make_golden.py``) and against the live reference when
``/root/reference`` is importable.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmergerec_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile merge_oracle.c with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "merge_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> None:
    lib().orc_set_threads(C.c_int(n))


# ----------------------------------------------------------------------------- helpers
def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrs: Sequence[np.ndarray]):
    arr = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        arr[i] = a.ctypes.data
    return arr


# ----------------------------------------------------------------------------- A0: flat layout
def flatten_model(model: "OrderedDict[str, np.ndarray]") -> Tuple[np.ndarray, "OrderedDict[str, tuple]"]:
    """merger/utils/model_operations.py:47-63 -- concat of reshape(-1) in dict order, ints -> fp32."""
    shape_dict = OrderedDict((k, tuple(v.shape)) for k, v in model.items())
    flat = np.concatenate([np.asarray(v).reshape(-1).astype(np.float32) for v in model.values()])
    return flat, shape_dict


def unflatten_model(flat: np.ndarray, shape_dict) -> "OrderedDict[str, np.ndarray]":
    """merger/utils/model_operations.py:66-90."""
    out, start = OrderedDict(), 0
    for name, shape in shape_dict.items():
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        out[name] = flat[start:start + n].reshape(shape)
        start += n
    assert start == flat.size, "Flattened tensor size does not match the expected size."
    return out


def segment_table(shape_dict, layer_wise: bool):
    """Blocks and lambda groups.

    task-wise (weight_learning/module/task_wise.py:36-48): one block [0, d), one group "all".
    layer-wise (weight_learning/module/layer_wise.py:13-33): one block per tensor, group key =
    name.split(".")[3] when "encoder.layer." is in the name, else "others"; group ids follow
    first-appearance order (defaultdict insertion order).
    Returns (seg_begin int64[P], seg_end int64[P], seg_group int32[P], group_keys list[str]).
    """
    sizes = [int(np.prod(s, dtype=np.int64)) if len(s) else 1 for s in shape_dict.values()]
    d = int(sum(sizes))
    if not layer_wise:
        return (np.zeros(1, np.int64), np.full(1, d, np.int64), np.zeros(1, np.int32), ["all"])
    keys: List[str] = []
    begins, ends, groups = [], [], []
    off = 0
    for name, n in zip(shape_dict.keys(), sizes):
        key = name.split(".")[3] if "encoder.layer." in name else "others"
        if key not in keys:
            keys.append(key)
        begins.append(off)
        ends.append(off + n)
        groups.append(keys.index(key))
        off += n
    return (np.asarray(begins, np.int64), np.asarray(ends, np.int64), np.asarray(groups, np.int32), keys)


# ----------------------------------------------------------------------------- merger
def task_vectors(base: np.ndarray, models: Sequence[np.ndarray]) -> np.ndarray:
    """A2 get_task_vectors, merger/algorithms/task_vector.py:8-10."""
    base = _f32(base)
    models = [_f32(m) for m in models]
    K, d = len(models), base.size
    T = np.empty((K, d), np.float32)
    lib().orc_task_vectors(_ptr(base), _ptr_array(models), C.c_int(K), C.c_int64(d), _ptr(T))
    return T


def merge_task_vector(base, models, weights: Sequence[float]) -> np.ndarray:
    """A1 merge_task_vector, merger/algorithms/task_vector.py:13-34 (weights: python floats -> fp32)."""
    assert len(models) == len(weights), "Number of models and weights should match."
    base = _f32(base)
    models = [_f32(m) for m in models]
    w = np.asarray(weights, dtype=np.float64).astype(np.float32)
    out = np.empty_like(base)
    lib().orc_merge_task_vector(_ptr(base), _ptr_array(models), C.c_int(len(models)),
                                C.c_int64(base.size), _ptr(w), _ptr(out))
    return out


def merge_linear(models, weights: Sequence[float]) -> np.ndarray:
    """A10 merge_linear, merger/algorithms/linear.py:8-27."""
    assert len(models) == len(weights), "Number of models and weights should match."
    models = [_f32(m) for m in models]
    w = np.asarray(weights, dtype=np.float64).astype(np.float32)
    out = np.empty_like(models[0])
    lib().orc_merge_linear(_ptr_array(models), C.c_int(len(models)), C.c_int64(out.size), _ptr(w), _ptr(out))
    return out


def lambda_merge(base, T, w, seg_begin=None, seg_end=None, seg_group=None) -> np.ndarray:
    """A3/A4 _merge_task_vectors (task_wise.py:36-48, layer_wise.py:64-83). w is (G,K) fp32."""
    base, T = _f32(base), _f32(T)
    K, d = T.shape
    w = _f32(w).reshape(-1, K)
    if seg_begin is None:
        seg_begin, seg_end, seg_group = np.zeros(1, np.int64), np.full(1, d, np.int64), np.zeros(1, np.int32)
    seg_begin = np.ascontiguousarray(seg_begin, np.int64)
    seg_end = np.ascontiguousarray(seg_end, np.int64)
    seg_group = np.ascontiguousarray(seg_group, np.int32)
    out = np.zeros(d, np.float32)  # layer_wise.py:65 zero-initialises
    lib().orc_lambda_merge(_ptr(base), _ptr(T), C.c_int64(d), C.c_int(K), C.c_int64(d), _ptr(w),
                           _ptr(seg_begin), _ptr(seg_end), _ptr(seg_group), C.c_int(len(seg_begin)), _ptr(out))
    return out


def combine_lambda(gw, pw, gb, use_softmax: bool) -> np.ndarray:
    """w = gw * (softmax(pw) if use_softmax else pw) + gb  (task_wise.py:37-42, layer_wise.py:67-74).

    fp32 throughout; softmax follows torch.softmax's max-subtracted form.  Bits of exp() are
    libm-dependent, so parity tests feed the kernels the same fp32 w instead of comparing this.
    """
    gw, pw, gb = _f32(gw), _f32(pw), _f32(gb)
    if use_softmax:
        m = pw.max(axis=-1, keepdims=True)
        e = np.exp(pw - m).astype(np.float32)
        pw = e / e.sum(axis=-1, keepdims=True, dtype=np.float32)
    return (gw * pw + gb).astype(np.float32)


def lambda_grad(grad, T, G: int = 1, seg_begin=None, seg_end=None, seg_group=None) -> np.ndarray:
    """A5: dL/dw[g,k] = sum_{p in g} sum_j grad[j] T[k,j]; fp64 truth, returns (G,K) float64."""
    grad, T = _f32(grad), _f32(T)
    K, d = T.shape
    if seg_begin is None:
        seg_begin, seg_end, seg_group = np.zeros(1, np.int64), np.full(1, d, np.int64), np.zeros(1, np.int32)
    seg_begin = np.ascontiguousarray(seg_begin, np.int64)
    seg_end = np.ascontiguousarray(seg_end, np.int64)
    seg_group = np.ascontiguousarray(seg_group, np.int32)
    out = np.zeros((G, K), np.float64)
    lib().orc_lambda_grad(_ptr(grad), _ptr(T), C.c_int64(d), C.c_int(K), _ptr(seg_begin), _ptr(seg_end),
                          _ptr(seg_group), C.c_int(len(seg_begin)), C.c_int(G), _ptr(out))
    return out


def ties_topk_count(density: float, d: int) -> int:
    """ties.py:14-15: int(density * numel) in Python double arithmetic."""
    return int(density * d)


def ties_select(base, models, density: float, weights: Optional[Sequence[float]] = None) -> np.ndarray:
    """A6 selection (ties.py:8-23) under the canonical lowest-index tie rule. Returns uint64 cut[K]:
    element j of model k survives iff ((bits(|u|) << 32) | (0xFFFFFFFF - j)) >= cut[k]."""
    base = _f32(base)
    models = [_f32(m) for m in models]
    K, d = len(models), base.size
    assert d <= 2 ** 32
    cut = np.zeros(K, np.uint64)
    w = None if weights is None else np.asarray(weights, np.float64).astype(np.float32)
    lib().orc_ties_select(_ptr(base), _ptr_array(models), C.c_int(K), C.c_int64(d),
                          None if w is None else _ptr(w), C.c_int64(ties_topk_count(density, d)), _ptr(cut))
    return cut


def ties_vectors(base, models, density: float, return_masks: bool = False):
    """A6-A8 get_ties_vectors, merger/algorithms/ties.py:55-72."""
    base = _f32(base)
    models = [_f32(m) for m in models]
    K, d = len(models), base.size
    cut = ties_select(base, models, density)
    That = np.empty((K, d), np.float32)
    trim = np.empty((K, d), np.uint8) if return_masks else None
    elect = np.empty((K, d), np.uint8) if return_masks else None
    lib().orc_ties_vectors(_ptr(base), _ptr_array(models), C.c_int(K), C.c_int64(d), _ptr(cut), _ptr(That),
                           None if trim is None else _ptr(trim), None if elect is None else _ptr(elect))
    if return_masks:
        return That, trim.astype(bool), elect.astype(bool), cut
    return That


def merge_ties(base, models, weights: Sequence[float], density: float) -> np.ndarray:
    """A9 merge_ties, merger/algorithms/ties.py:75-83 (trim + weighted sum only)."""
    assert len(models) == len(weights), "Number of models and weights should match."
    base = _f32(base)
    models = [_f32(m) for m in models]
    cut = ties_select(base, models, density, weights)
    w = np.asarray(weights, np.float64).astype(np.float32)
    out = np.empty_like(base)
    lib().orc_merge_ties(_ptr(base), _ptr_array(models), C.c_int(len(models)), C.c_int64(base.size),
                         _ptr(w), _ptr(cut), _ptr(out))
    return out


def lns_vectors(base, models, density: float = 0.05) -> np.ndarray:
    """get_localize_and_stitch_vectors, merger/algorithms/localize_and_stitch.py:8-49, under the canonical
    lowest-index tie rule of the top-k (the same select as TIES)."""
    base = _f32(base)
    models = [_f32(m) for m in models]
    K, d = len(models), base.size
    u = np.stack([m - base for m in models]).astype(np.float32)                       # :26
    if int(density * d) <= 0:                                                          # :30-33
        return np.zeros_like(u)
    cut = ties_select(base, models, density)
    key = (u.view(np.uint32).astype(np.uint64) & np.uint64(0x7FFFFFFF)) << np.uint64(32)
    key |= (np.uint64(0xFFFFFFFF) - np.arange(d, dtype=np.uint64))[None, :]
    masks = (key >= cut[:, None]).astype(np.float32)                                   # :36-39
    denom = np.maximum(masks.sum(axis=0, dtype=np.float32), np.float32(1.0))           # :42-43
    return ((masks / denom).astype(np.float32) * u).astype(np.float32)                 # :44-48


def merge_localize_and_stitch(base, models, weights: Sequence[float], density: float = 0.05) -> np.ndarray:
    """merge_localize_and_stitch, localize_and_stitch.py:52-82: base + sum_dim0(vectors * w)."""
    assert len(models) == len(weights), "Number of models and weights should match."
    w = np.asarray(weights, np.float64).astype(np.float32).reshape(1, -1)
    return lambda_merge(base, lns_vectors(base, models, density), w)


# ----------------------------------------------------------------------------- evaluator
def scores_f32(U, I) -> np.ndarray:
    """B1: scores = U @ I.T (module/recommender/module.py:137), fp32 BLAS."""
    return _f32(U) @ _f32(I).T


def scores_f64(U, I) -> np.ndarray:
    return np.asarray(U, np.float64) @ np.asarray(I, np.float64).T


def to_bf16(x) -> np.ndarray:
    """fp32 -> bf16 -> fp32, round to nearest even (torch's `.to(torch.bfloat16)`)."""
    u = _f32(x).view(np.uint32).astype(np.uint64)
    r = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    nan_inf = (u & np.uint64(0x7F800000)) == np.uint64(0x7F800000)
    r = np.where(nan_inf, u & np.uint64(0xFFFF0000), r)
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def scores_bf16(U, I) -> np.ndarray:
    """The reference's default pipeline (Lightning precision="bf16-mixed", configs/base.py:41): bf16 operands, fp32
    accumulation, bf16 result.  Exact (order-independent) whenever the fp32 sum of the bf16 products is exact, e.g.
    on the grid catalogs; returned as fp32 holding bf16 values."""
    return to_bf16(to_bf16(U).astype(np.float32) @ to_bf16(I).astype(np.float32).T)


def normalize(x) -> np.ndarray:
    """_maybe_normalize for cosine similarity (module.py:74-77): x / max(||x||_2, 1e-12)."""
    x = _f32(x)
    nrm = np.sqrt((x * x).sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)
    return (x / np.maximum(nrm, np.float32(1e-12))).astype(np.float32)


def topk_rows(scores, k: int, id_base: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """B2 (evaluator/evaluator.py:43) under the canonical order (score desc, id asc)."""
    scores = _f32(scores)
    Q, N = scores.shape
    assert k <= N
    vals = np.empty((Q, k), np.float32)
    ids = np.empty((Q, k), np.int32)
    lib().orc_topk_rows(_ptr(scores), C.c_int64(Q), C.c_int64(N), C.c_int64(N), C.c_int(k),
                        C.c_int32(id_base), _ptr(vals), _ptr(ids))
    return vals, ids


def topk_merge(vals, ids) -> Tuple[np.ndarray, np.ndarray]:
    """Merge (G,Q,K) per-shard lists into (Q,K) under the canonical order."""
    vals = _f32(vals)
    ids = np.ascontiguousarray(ids, np.int32)
    G, Q, K = vals.shape
    ov = np.empty((Q, K), np.float32)
    oi = np.empty((Q, K), np.int32)
    lib().orc_topk_merge(_ptr(vals), _ptr(ids), C.c_int(G), C.c_int64(Q), C.c_int(K), _ptr(ov), _ptr(oi))
    return ov, oi


def label_rank(ids, labels) -> np.ndarray:
    ids = np.ascontiguousarray(ids, np.int32)
    labels = np.ascontiguousarray(labels, np.int64)
    Q, K = ids.shape
    rank = np.empty(Q, np.int32)
    lib().orc_label_rank(_ptr(ids), C.c_int64(Q), C.c_int(K), _ptr(labels), _ptr(rank))
    return rank


def _log2_f32(x: int) -> float:
    """float(torch.log2(torch.tensor(x)))  (metrics.py:84): int64 -> fp32 log2 -> python float.
    torch's result equals the correctly rounded fp32 log2 for every x in 2..1025 (pinned by
    tests/golden/evaluator.npz:ndcg_gain_table); numpy's own fp32 log2 is off by one ulp at 8 of them."""
    return float(np.float32(math.log2(x)))


def evaluate_ids(pred_ids, labels, metrics: Sequence[str], ks: Sequence[int], prefix: str = "") -> Dict[str, float]:
    """B3/B4: Recall (metrics.py:35-59) and NDCG (metrics.py:62-88) on predicted id lists, with
    the reference's row-order python ``sum()/len`` and its key order (evaluator.py:11-15,45-47)."""
    pred = np.asarray(pred_ids).tolist()
    true = np.asarray(labels).tolist()
    out: Dict[str, float] = {}
    for metric in metrics:
        for k in ks:
            vals = []
            for p, t in zip(pred, true):
                p = p[:k]
                if metric == "RECALL":
                    vals.append(1.0 if t in p else 0.0)
                elif metric == "NDCG":
                    vals.append(1 / _log2_f32(p.index(t) + 2) if t in p else 0.0)
                else:
                    raise KeyError(metric)
            name = {"RECALL": "Recall", "NDCG": "NDCG"}[metric]
            out[f"{prefix}{name}@{k}"] = sum(vals) / len(vals) if vals else 0.0
    return out


def evaluate(scores, labels, metrics: Sequence[str], ks: Sequence[int], prefix: str = "") -> Dict[str, float]:
    """Evaluator.__call__ (evaluator/evaluator.py:31-49) with canonical top-K order."""
    _, ids = topk_rows(scores, max(ks))
    return evaluate_ids(ids, labels, metrics, ks, prefix)


# ---- distillation step -------------------------------------------------------------
# fp64 restatement of rec_retrieval/module/recommender/loss_fn.py on ONE sample (the reference calls every loss with
# `(1, num_items)` tensors, sequence/module.py:69) plus the analytic gradient w.r.t. the merged model's logits.
def _softmax64(x):
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - x.max())
    return e / e.sum()


def _lse64(x):
    x = np.asarray(x, dtype=np.float64)
    m = x.max()
    return m + math.log(np.exp(x - m).sum())


def distill_loss(z, t, loss_type: str, temperature: float = 1.0, coefficient: float = 0.0, margin: float = 0.0):
    """(loss, d loss / d z) for merged logits z (n,) and teacher logits t (n,).  loss_type: the reference's LossType
    names plus "PAIRWISE" / "LISTNET" (classes of loss_fn.py that the factory does not list).  argmax = first maximum."""
    z = np.asarray(z, dtype=np.float64)
    t = None if t is None else np.asarray(t, dtype=np.float64)
    n = z.shape[0]
    T = float(temperature)

    def ce(target):                                     # F.cross_entropy, loss_fn.py:41-42
        p = _softmax64(z)
        g = p.copy()
        g[target] -= 1.0
        return _lse64(z) - z[target], g

    def kd():                                           # loss_fn.py:52-58
        P, Q = _softmax64(t / T), _softmax64(z / T)
        logP = t / T - _lse64(t / T)
        logQ = z / T - _lse64(z / T)
        return float((P * (logP - logQ)).sum()) * T * T, T * (Q - P)

    def entropy():                                      # loss_fn.py:64-67 (fp32 1e-8 inside the log)
        p = _softmax64(z)
        eps = float(np.float32(1e-8))
        h = -(np.log(p + eps) + p / (p + eps))
        return float(-(p * np.log(p + eps)).sum()), p * (h - float((p * h).sum()))

    if loss_type in ("CE", "SINGLE_PSEUDO_LABEL"):
        return ce(int(np.argmax(t)))
    if loss_type == "KD":
        return kd()
    if loss_type == "MSE":                              # loss_fn.py:185
        return float(((z - t) ** 2).mean()), 2.0 * (z - t) / n
    if loss_type == "ADAMERGING":
        return entropy()
    if loss_type == "ADAMERGING_KD":                    # loss_fn.py:81-86
        (a, ga), (b, gb) = entropy(), kd()
        return a + coefficient * b, ga + coefficient * gb
    if loss_type == "MERGED_PSEUDO_LABEL":              # loss_fn.py:96-104
        return ce(int(np.argmax(z)))
    if loss_type == "MERGED_PSEUDO_LABEL_KD":
        (a, ga), (b, gb) = ce(int(np.argmax(z))), kd()
        return a + coefficient * b, ga + coefficient * gb
    if loss_type == "SINGLE_PSEUDO_LABEL_KD":
        (a, ga), (b, gb) = ce(int(np.argmax(t))), kd()
        return a + coefficient * b, ga + coefficient * gb
    if loss_type == "PAIRWISE":                         # loss_fn.py:195-210
        pos = int(np.argmax(t))
        masked = t.copy()
        masked[pos] = -np.inf
        neg = int(np.argmax(masked))
        h = margin - (z[pos] - z[neg])
        g = np.zeros(n)
        if h > 0:
            g[pos] -= 1.0
            g[neg] += 1.0
        return max(h, 0.0), g
    if loss_type == "LISTNET":                          # loss_fn.py:221-228
        P, Q = _softmax64(t / T), _softmax64(z / T)
        logQ = z / T - _lse64(z / T)
        return float(-(P * logQ).sum()), (Q - P) / T
    raise ValueError(f"unknown loss type {loss_type}")


def teacher_scores(item_embedding, sequence_embedding) -> np.ndarray:
    """merge_train.py:116-126: normalised sequence embeddings @ normalised item embeddings.T (fp64 here)."""
    I = np.asarray(item_embedding, dtype=np.float64)
    S = np.asarray(sequence_embedding, dtype=np.float64)
    I = I / np.linalg.norm(I, axis=-1, keepdims=True)
    S = S / np.linalg.norm(S, axis=-1, keepdims=True)
    return S @ I.T


def distill_step(rep, item_tables, dataset_indexes, teacher_rows, loss_type: str, temperature: float = 1.0,
                 coefficient: float = 0.0, margin: float = 0.0):
    """`_forward_distill` (sequence/module.py:59-76) in fp64: per-sample losses, their mean, and the gradient of the
    mean w.r.t. the representations (item tables carry no gradient, callbacks.py:48-50)."""
    rep = np.asarray(rep, dtype=np.float64)
    B = rep.shape[0]
    losses = np.zeros(B)
    grad = np.zeros_like(rep)
    for i, d in enumerate(dataset_indexes):
        items = np.asarray(item_tables[d], dtype=np.float64)
        z = items @ rep[i]
        lv, gz = distill_loss(z, None if teacher_rows is None else teacher_rows[i], loss_type, temperature, coefficient, margin)
        losses[i] = lv
        grad[i] = (gz @ items) / B
    return losses, float(losses.mean()), grad


# ---- PCB merging --------------------------------------------------------------------
def _sum_dim0(X: np.ndarray) -> np.ndarray:
    """torch.sum(X, dim=0) for a (K, d) fp32 matrix in ATen's order (0 + 1*x is exact, so the lambda merge does it)."""
    X = _f32(X)
    return lambda_merge(np.zeros(X.shape[1], np.float32), X, np.ones((1, X.shape[0]), np.float32))


def pcb_vectors(base, models, density: float = 0.2, return_task: bool = False):
    """get_pcb_vectors, merger/algorithms/pcb.py:37-58, in numpy fp32 (numpy's exp / tanh stand in for torch's: both
    are within an ulp of the true value, so the restatement is pinned to the golden vectors at ~1e-6, not bitwise)."""
    base = _f32(base)
    tv = np.stack([_f32(m) - base for m in models]).astype(np.float32)                 # :38
    n, d = tv.shape

    def clamp(x, min_ratio, max_ratio):                                                # :16-29
        s = np.sort(x, axis=1)
        lo = s[:, int(d * min_ratio)][:, None]
        hi = s[:, int(d * (1 - max_ratio) - 1)][:, None]
        return np.minimum(np.maximum(x, lo), hi), lo, hi

    def normalize(x):                                                                  # :9-13
        mn, mx = x.min(axis=1, keepdims=True), x.max(axis=1, keepdims=True)
        return ((x - mn) / (mx - mn)).astype(np.float32)

    A, lo, hi = clamp(np.abs(tv), 0.01, 0.01)                                          # :42
    clamped = (np.sign(tv) * A).astype(np.float32)                                     # :43
    self_pcb = normalize(A)
    self_pcb = (self_pcb * self_pcb).astype(np.float32)                                # :45
    self_act = np.exp(np.float32(n) * self_pcb).astype(np.float32)                     # :46
    cross = (tv * _sum_dim0(tv)[None, :]).astype(np.float32)                           # :48
    task = (self_act * np.tanh(cross).astype(np.float32)).astype(np.float32)           # :49-51
    cl, q, mx = clamp(task, 1 - density, 0)
    scale = normalize(cl)                                                              # :53
    pcb = (clamped * scale).astype(np.float32)                                         # :54
    pcb = (pcb / np.maximum(_sum_dim0(scale), np.float32(1e-12))[None, :]).astype(np.float32)   # :55
    pcb = (pcb / np.float32(n)).astype(np.float32)                                     # :56
    if return_task:
        return pcb, task, q[:, 0], mx[:, 0], lo[:, 0], hi[:, 0]
    return pcb


def merge_pcb(base, models, weights: Sequence[float], density: float = 0.2) -> np.ndarray:
    """merge_pcb, pcb.py:61-73: base-first accumulation of weights[i] * pcb_vector_i."""
    assert len(models) == len(weights), "Number of models and weights should match."
    base = _f32(base)
    vec = pcb_vectors(base, models, density)
    acc = base.copy()
    for i in range(len(models)):
        acc = (acc + (np.float32(weights[i]) * vec[i]).astype(np.float32)).astype(np.float32)
    return acc


# ---- DARE (merger/algorithms/dare.py:9-31) with explicit keep masks --------------------------------------------------
def merge_dare(base, models, weights: Sequence[float], density: float, masks) -> np.ndarray:
    """merged = base; merged += fl(fl(w_i * (m_i - base)) * noise_i), noise = keep / (1 - p) in fp32 (torch's CPU dropout:
    `input * bernoulli_(1 - p).div_(1 - p)`; p == 0 returns the input, p == 1 multiplies by zeros)."""
    assert len(models) == len(weights), "Number of models and weights should match."
    base = _f32(base)
    p = float(density)
    masks = np.asarray(masks).astype(bool)
    acc = base.copy()
    for i, m in enumerate(models):
        update = (np.float32(weights[i]) * (_f32(m) - base).astype(np.float32)).astype(np.float32)
        if p == 1.0:
            noise = np.zeros_like(update)
        elif p == 0.0:
            noise = np.ones_like(update)
        else:
            noise = (masks[i].astype(np.float32) / np.float32(1.0 - p)).astype(np.float32)
        acc = (acc + (update * noise).astype(np.float32)).astype(np.float32)
    return acc
