"""``Evaluator`` -- full-catalog top-K evaluation (reference: rec_retrieval/evaluator/evaluator.py:6-49).

``Evaluator(metrics, ks)(scores, labels, metric_prefix)`` keeps the reference contract (a materialised (Q, N)
score matrix in, ``{prefix}{Recall|NDCG}@{k}`` python floats out, same key order).  Because that contract forces
the caller to build the score matrix (module/recommender/module.py:137, :344-352), two additions take the
embeddings instead and never materialise it: ``topk_embeddings`` and ``evaluate_embeddings`` run the fused
tensor-core scoring + per-row top-K kernel (``mr_score_topk``), shard-aware through ``ShardedItemTable``.

Top-K order is (score desc, item id asc): ``torch.topk``'s order among equal scores is unspecified (SURVEY.md
0.1-D3), so ids can differ from the raw reference only inside groups of exactly equal scores.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import torch

from .. import _lib
from .enums import MetricType
from .metrics import label_rank
from .sharded import (MAX_FUSED_TOPK, MR_SCORE_BF16, MR_SCORE_TF32X3, ShardedItemTable, exchange_packed,
                      new_packed_list, split_tf32, to_bf16)

MAX_TOPK = 1024


def topk_rows(scores: torch.Tensor, k: int, id_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """(values (Q, k) fp32, ids (Q, k) int32) of a (Q, N) score matrix: `mr_topk_rows`, the drop-in for
    ``torch.topk(scores, k, dim=1)`` (evaluator.py:43).  CPU inputs are copied to the GPU in row chunks."""
    dev = _lib.require_cuda()
    lib = _lib.load()
    if scores.dim() != 2:
        raise ValueError("scores must be (Q, N)")
    Q, N = scores.shape
    if k > N:
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")  # torch.topk's error class
    out_v = torch.empty((Q, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((Q, k), dtype=torch.int32, device=dev)
    rows_per_chunk = Q if scores.is_cuda else max(1, min(Q, (1 << 30) // max(4 * N, 1)))
    for q0 in range(0, Q, max(rows_per_chunk, 1)):
        q1 = min(Q, q0 + rows_per_chunk)
        chunk = scores[q0:q1].to(device=dev, dtype=torch.float32, non_blocking=True)
        if chunk.stride(1) != 1:
            chunk = chunk.contiguous()
        _lib.check(lib.mr_topk_rows(_lib.dptr(chunk), q1 - q0, N, chunk.stride(0), k, id_base,
                                    _lib.dptr(out_v[q0:q1]), _lib.dptr(out_i[q0:q1]), _lib.stream_handle()),
                   "mr_topk_rows")
    return out_v, out_i


def score_topk(user_hi: torch.Tensor, user_lo: Optional[torch.Tensor], table: ShardedItemTable, k: int,
               mode: int = MR_SCORE_TF32X3, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Local fused scoring + top-k of pre-split queries against this rank's shard (`mr_score_topk`).  `out` =
    (values (Q, k) fp32, ids (Q, k) int32) to write into (e.g. the planes of the exchange buffer)."""
    dev = _lib.require_cuda()
    lib = _lib.load()
    Q, E = user_hi.shape
    if E != table.dim:
        raise ValueError(f"embedding dims differ: queries {E}, items {table.dim}")
    if out is None:
        out_v = torch.empty((Q, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((Q, k), dtype=torch.int32, device=dev)
    else:
        out_v, out_i = out
    ws_bytes = int(lib.mr_score_topk_workspace_bytes(Q, table.n_local, E, k))
    if ws_bytes < 0:
        _lib.check(ws_bytes, "mr_score_topk_workspace_bytes")
    ws = table.workspace(ws_bytes)
    _lib.check(lib.mr_score_topk(_lib.dptr(user_hi), _lib.dptr(user_lo), Q, _lib.dptr(table.hi), _lib.dptr(table.lo),
                                 table.n_local, E, k, table.id_base, mode, _lib.dptr(out_v), _lib.dptr(out_i),
                                 _lib.dptr(ws), ws_bytes, _lib.stream_handle()), "mr_score_topk")
    return out_v, out_i


class PreparedQueries:
    """Query embeddings in the operand format of the scoring kernel ((hi, lo) TF32 split, or bf16), prepared once and
    reusable against any number of item tables / evaluations (`Evaluator.prepare_queries`)."""

    def __init__(self, user_emb: torch.Tensor, normalize: bool = False, bf16: bool = False):
        dev = _lib.require_cuda()
        users = user_emb.to(device=dev, dtype=torch.float32)
        if users.dim() != 2:
            raise ValueError("query embeddings must be (Q, E)")
        if normalize:
            users = torch.nn.functional.normalize(users, p=2, dim=-1)   # module/recommender/module.py:74-77
        self.bf16 = bool(bf16)
        self.hi, self.lo = (to_bf16(users), None) if self.bf16 else split_tf32(users)
        self.shape = tuple(users.shape)


class Evaluator:
    def __init__(self, metrics: List[str], ks: List[int]):
        self.metric_names = metrics
        self.ks = ks
        self._max_k = max(ks)
        self._metrics = []
        for metric in metrics:
            for k in ks:
                self._metrics.append(MetricType[metric].metric_cls(k))

    # ------------------------------------------------------------------ reference contract
    def evaluate(self, scores: torch.Tensor, labels: torch.Tensor, metric_prefix: str = "") -> Dict[str, float]:
        return self(scores, labels, metric_prefix)

    def __call__(self, scores: torch.Tensor, labels: torch.Tensor, metric_prefix: str = "") -> Dict[str, float]:
        _, ids = topk_rows(scores, self._max_k)
        return self.metrics_from_ids(ids, labels, metric_prefix)

    # ------------------------------------------------------------------ fused additions
    def topk_embeddings(self, user_emb: Union[torch.Tensor, PreparedQueries], item_emb: Union[torch.Tensor, ShardedItemTable],
                        k: Optional[int] = None, normalize: bool = False,
                        mode: int = MR_SCORE_TF32X3) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top-k (values fp32, global ids int32), both (Q, k), of ``user_emb @ item_emb.T`` without the matrix.

        ``item_emb`` is an (N, E) tensor (single GPU) or a ``ShardedItemTable`` (reusable, possibly one shard of
        a multi-GPU catalog -- then every rank passes the same queries and receives the same merged lists).
        ``normalize`` applies the reference's cosine normalisation to the queries (and to a raw item tensor).
        ``mode``: ``MR_SCORE_TF32X3`` (default, fp32-faithful), ``MR_SCORE_TF32X1``, or ``MR_SCORE_BF16`` -- the
        bf16-compat mode that mirrors the reference's default ``precision="bf16-mixed"`` runs (bf16 operands, fp32
        accumulation, scores rounded to bf16; configs/base.py:41) with the canonical tie rule on equal scores."""
        k = self._max_k if k is None else k
        if k > MAX_FUSED_TOPK:
            raise ValueError(f"fused top-k supports k <= {MAX_FUSED_TOPK}")
        _lib.require_cuda()
        table = item_emb if isinstance(item_emb, ShardedItemTable) else ShardedItemTable(
            item_emb, normalize=normalize, bf16=(mode == MR_SCORE_BF16))
        if table.bf16 != (mode == MR_SCORE_BF16):
            raise ValueError("the item table was prepared for a different scoring mode (bf16 table <-> MR_SCORE_BF16)")
        if k > table.n_total:
            raise RuntimeError(f"selected index k out of range (k={k}, N={table.n_total})")
        queries = user_emb if isinstance(user_emb, PreparedQueries) else PreparedQueries(user_emb, normalize, table.bf16)
        if queries.bf16 != table.bf16:
            raise ValueError("the queries were prepared for a different scoring mode than the item table")
        if table.group is None or table.world == 1:
            return score_topk(queries.hi, queries.lo, table, k, mode)
        # sharded: the kernel writes this rank's list into the buffer that the single all-gather sends
        local, loc_v, loc_i = new_packed_list(queries.shape[0], k, queries.hi.device)
        score_topk(queries.hi, queries.lo, table, k, mode, out=(loc_v, loc_i))
        return exchange_packed(local, k, table.group, gathered=table.gather_buffer(queries.shape[0], k))

    def topk_embeddings_streamed(self, user_emb: Union[torch.Tensor, "PreparedQueries"], host_items: torch.Tensor,
                                 k: Optional[int] = None, id_base: int = 0, n_total: Optional[int] = None, group=None,
                                 normalize: bool = False, mode: int = MR_SCORE_TF32X3,
                                 first_rows: int = 32768, max_chunks: int = 12) -> Tuple[torch.Tensor, torch.Tensor]:
        """`topk_embeddings` for an item table that still lives in HOST memory (what `load_item_embeddings` /
        `ItemEncoderMixin.encode_items(...).cpu()` hand over): the rows are copied to the GPU in chunks on a copy stream
        while the previous chunk is being scored, so the transfer hides behind the tensor-core work instead of preceding
        it.  Chunk sizes ramp up geometrically from `first_rows` (only the first, small copy is exposed); every chunk is
        split into its TF32 operands, scored into its own sorted list, and the lists are merged (`mr_topk_merge_packed`).
        Because a (query, item) score does not depend on how the table is cut, the result is bit-identical to the
        one-table call.  `host_items` = this rank's rows (pinned memory makes the copies asynchronous); `id_base`,
        `n_total`, `group` as for `ShardedItemTable`."""
        from .sharded import exchange_packed, topk_merge_packed
        k = self._max_k if k is None else k
        if k > MAX_FUSED_TOPK:
            raise ValueError(f"fused top-k supports k <= {MAX_FUSED_TOPK}")
        dev = _lib.require_cuda()
        if host_items.dim() != 2:
            raise ValueError("item table must be (N, E)")
        n_local = int(host_items.shape[0])
        n_total = n_local if n_total is None else int(n_total)
        if k > n_total:
            raise RuntimeError(f"selected index k out of range (k={k}, N={n_total})")
        bf16 = mode == MR_SCORE_BF16
        queries = user_emb if isinstance(user_emb, PreparedQueries) else PreparedQueries(user_emb, normalize, bf16)
        Q = queries.shape[0]
        # chunk plan: geometric ramp, at most max_chunks lists (k * chunks candidates per row must stay mergeable)
        max_chunks = max(1, min(max_chunks, 8192 // max(k, 1)))
        bounds, lo, rows = [], 0, max(int(first_rows), 1)
        while lo < n_local:
            if len(bounds) == max_chunks - 1:
                rows = n_local - lo
            hi = min(n_local, lo + rows)
            bounds.append((lo, hi))
            lo, rows = hi, rows * 2
        C = max(len(bounds), 1)
        partial = torch.empty((C, 2, Q, k), dtype=torch.int32, device=dev)
        if not bounds:        # an empty shard: every list is empty
            partial[:, 0].view(torch.float32).fill_(float("-inf"))
            partial[:, 1].fill_(-1)
        main = torch.cuda.current_stream(dev)
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
        copy = self._copy_stream
        copy.wait_stream(main)
        ws = None
        for c, (a, b) in enumerate(bounds):
            with torch.cuda.stream(copy):
                chunk = host_items[a:b].to(device=dev, dtype=torch.float32, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy)
            main.wait_event(ready)
            chunk.record_stream(main)
            table = ShardedItemTable(chunk, id_base=id_base + a, n_total=n_total, normalize=normalize, bf16=bf16)
            table._ws = ws                                            # one scratch for all chunks (grown if needed)
            if b - a >= k:
                score_topk(queries.hi, queries.lo, table, k, mode, out=(partial[c, 0].view(torch.float32), partial[c, 1]))
            else:
                self._score_small_chunk(queries, table, k, mode, partial[c])
            ws = table._ws
        vals, ids = topk_merge_packed(partial, k) if C > 1 else (partial[0, 0].view(torch.float32), partial[0, 1])
        if group is not None:
            import torch.distributed as dist
            if dist.get_world_size(group) > 1:
                local = torch.stack([vals.contiguous().view(torch.int32), ids.contiguous()])
                return exchange_packed(local, k, group)
        return vals, ids

    @staticmethod
    def _score_small_chunk(queries: "PreparedQueries", table: ShardedItemTable, k: int, mode: int, out: torch.Tensor) -> None:
        """A chunk with fewer than k rows: its list has only n < k entries; the rest of the (Q, k) plane is marked empty."""
        n = table.n_local
        out[0].view(torch.float32).fill_(float("-inf"))
        out[1].fill_(-1)
        if n:
            v, i = score_topk(queries.hi, queries.lo, table, n, mode)
            out[0].view(torch.float32)[:, :n] = v
            out[1][:, :n] = i

    def evaluate_embeddings_streamed(self, user_emb, host_items: torch.Tensor, labels: torch.Tensor, metric_prefix: str = "",
                                     **kw) -> Dict[str, float]:
        """`evaluate_embeddings` with a host-resident item table (see `topk_embeddings_streamed`)."""
        _, ids = self.topk_embeddings_streamed(user_emb, host_items, self._max_k, **kw)
        return self.metrics_from_ids(ids, labels, metric_prefix)

    @staticmethod
    def prepare_queries(user_emb: torch.Tensor, normalize: bool = False, mode: int = MR_SCORE_TF32X3) -> PreparedQueries:
        """Split / convert the query embeddings once; pass the result as `user_emb` to `topk_embeddings` /
        `evaluate_embeddings` as often as needed (several catalogs, several ks, repeated evaluation)."""
        return PreparedQueries(user_emb, normalize, mode == MR_SCORE_BF16)

    def evaluate_embeddings(self, user_emb: Union[torch.Tensor, PreparedQueries], item_emb: Union[torch.Tensor, ShardedItemTable],
                            labels: torch.Tensor, metric_prefix: str = "", normalize: bool = False,
                            mode: int = MR_SCORE_TF32X3) -> Dict[str, float]:
        _, ids = self.topk_embeddings(user_emb, item_emb, self._max_k, normalize, mode)
        return self.metrics_from_ids(ids, labels, metric_prefix)

    # ------------------------------------------------------------------ shared tail
    def metrics_from_ids(self, ids: torch.Tensor, labels: torch.Tensor, metric_prefix: str = "") -> Dict[str, float]:
        """One rank lookup on the GPU (Q int32 back to the host), then every (metric, k) from the same ranks."""
        ranks = label_rank(ids, labels).cpu().numpy()
        found = ranks.compress(ranks >= 0)     # compressed once; row order (= the reference's summation order) is kept
        n = int(ranks.shape[0])
        return {metric_prefix + m.name: m.from_found(found, n) for m in self._metrics}
