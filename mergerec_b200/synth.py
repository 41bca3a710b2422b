"""Synthetic inputs of the named shapes (there is no network for checkpoints or datasets).

Shape tables restate the state_dict layouts the reference merges:

* BLaIR-base = HF ``RobertaModel`` (RoBERTa-base + pooler) wrapped as ``BaseEncoderModel.model``
  (module/models/encoder/blair.py:11-16) -> keys ``model.embeddings.*``, ``model.encoder.layer.N.*``,
  ``model.pooler.dense.*``; P = 199 tensors, d = 124,645,632.
* Recformer = ``RecformerEmbeddings`` (module/models/encoder/recformer/models.py:85-96, incl. the
  int64 ``position_ids`` buffer that flatten promotes to fp32) + HF ``LongformerEncoder``;
  Recformer-large: P = 535, d = 433,610,754.

Generators are numpy (PCG64) so that fixtures made in the build container and tests run on the
GPU box see identical bits.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Sequence, Tuple

import numpy as np

ShapeTable = "OrderedDict[str, Tuple[int, ...]]"


def _encoder_layer(prefix: str, h: int, ffn: int, longformer: bool) -> List[Tuple[str, Tuple[int, ...]]]:
    rows: List[Tuple[str, Tuple[int, ...]]] = []
    proj = ["query", "key", "value"] + (["query_global", "key_global", "value_global"] if longformer else [])
    for p in proj:
        rows += [(f"{prefix}.attention.self.{p}.weight", (h, h)), (f"{prefix}.attention.self.{p}.bias", (h,))]
    rows += [
        (f"{prefix}.attention.output.dense.weight", (h, h)),
        (f"{prefix}.attention.output.dense.bias", (h,)),
        (f"{prefix}.attention.output.LayerNorm.weight", (h,)),
        (f"{prefix}.attention.output.LayerNorm.bias", (h,)),
        (f"{prefix}.intermediate.dense.weight", (ffn, h)),
        (f"{prefix}.intermediate.dense.bias", (ffn,)),
        (f"{prefix}.output.dense.weight", (h, ffn)),
        (f"{prefix}.output.dense.bias", (h,)),
        (f"{prefix}.output.LayerNorm.weight", (h,)),
        (f"{prefix}.output.LayerNorm.bias", (h,)),
    ]
    return rows


def roberta_shapes(layers: int = 12, hidden: int = 768, ffn: int = 3072, vocab: int = 50265,
                   max_pos: int = 514, type_vocab: int = 1, pooler: bool = True) -> ShapeTable:
    """BLaIR-base / RoBERTa-base layout (defaults: P = 199, d = 124,645,632)."""
    rows: List[Tuple[str, Tuple[int, ...]]] = [
        ("model.embeddings.word_embeddings.weight", (vocab, hidden)),
        ("model.embeddings.position_embeddings.weight", (max_pos, hidden)),
        ("model.embeddings.token_type_embeddings.weight", (type_vocab, hidden)),
        ("model.embeddings.LayerNorm.weight", (hidden,)),
        ("model.embeddings.LayerNorm.bias", (hidden,)),
    ]
    for i in range(layers):
        rows += _encoder_layer(f"model.encoder.layer.{i}", hidden, ffn, longformer=False)
    if pooler:
        rows += [("model.pooler.dense.weight", (hidden, hidden)), ("model.pooler.dense.bias", (hidden,))]
    return OrderedDict(rows)


def recformer_shapes(layers: int = 24, hidden: int = 1024, ffn: int = 4096, vocab: int = 50265,
                     max_pos: int = 4098, token_types: int = 4, max_items: int = 51) -> ShapeTable:
    """Recformer layout (defaults = Recformer-large: P = 535, d = 433,610,754).

    ``model.embeddings.position_ids`` (1, max_pos) sits between the embedding parameters and the
    encoder, so every later tensor starts at a flat offset = 2 (mod 4)."""
    rows: List[Tuple[str, Tuple[int, ...]]] = [
        ("model.embeddings.word_embeddings.weight", (vocab, hidden)),
        ("model.embeddings.position_embeddings.weight", (max_pos, hidden)),
        ("model.embeddings.token_type_embeddings.weight", (token_types, hidden)),
        ("model.embeddings.item_position_embeddings.weight", (max_items, hidden)),
        ("model.embeddings.LayerNorm.weight", (hidden,)),
        ("model.embeddings.LayerNorm.bias", (hidden,)),
        ("model.embeddings.position_ids", (1, max_pos)),
    ]
    for i in range(layers):
        rows += _encoder_layer(f"model.encoder.layer.{i}", hidden, ffn, longformer=True)
    return OrderedDict(rows)


def tiny_shapes(layers: int = 2, hidden: int = 24, ffn: int = 40, vocab: int = 53, max_pos: int = 10,
                recformer: bool = False) -> ShapeTable:
    """HF-named toy layout with ragged, non-multiple-of-32 (and, for recformer, non-multiple-of-4
    offset) tensors for the parity tests."""
    if recformer:
        return recformer_shapes(layers, hidden, ffn, vocab, max_pos, token_types=4, max_items=5)
    return roberta_shapes(layers, hidden, ffn, vocab, max_pos, type_vocab=1, pooler=True)


def numel(shape: Sequence[int]) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def total_numel(shapes: ShapeTable) -> int:
    return sum(numel(s) for s in shapes.values())


# ------------------------------------------------------------------------------ state dicts
def make_state_dicts(shapes: ShapeTable, K: int, seed: int = 0, sigma: float = 1e-3,
                     init_std: float = 0.02) -> Tuple["OrderedDict[str, np.ndarray]", List["OrderedDict[str, np.ndarray]"]]:
    """base ~ N(0, init_std) (LayerNorm weights 1, biases 0, position_ids = arange as int64);
    domain model k = base + sigma * N(0,1) with its own stream."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("position_ids"):
            base[name] = np.arange(numel(shape), dtype=np.int64).reshape(shape)
        elif "LayerNorm.weight" in name:
            base[name] = np.ones(shape, np.float32)
        elif name.endswith(".bias"):
            base[name] = np.zeros(shape, np.float32)
        else:
            base[name] = (rng.standard_normal(shape, dtype=np.float32) * np.float32(init_std)).astype(np.float32)
    models = []
    for k in range(K):
        rk = np.random.Generator(np.random.PCG64(seed + 1 + k))
        m: "OrderedDict[str, np.ndarray]" = OrderedDict()
        for name, shape in shapes.items():
            if name.endswith("position_ids"):
                m[name] = base[name].copy()
            else:
                m[name] = (base[name] + np.float32(sigma) * rk.standard_normal(shape, dtype=np.float32)).astype(np.float32)
        models.append(m)
    return base, models


def make_flat(d: int, K: int, seed: int = 0, sigma: float = 1e-3, tie_free: bool = False,
              quantize: float = 0.0) -> Tuple[np.ndarray, List[np.ndarray]]:
    """Flat base (d,) and K flat models.

    tie_free: |m_k - base| are all distinct by construction (a permutation of distinct fp32 values,
              exactly representable differences) -> raw torch.topk parity holds bit-for-bit.
    quantize: round task vectors to a multiple of ``quantize`` (many exact magnitude ties)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if tie_free:
        # base on a coarse dyadic grid, updates = +-(i+1) * 2^-20 : m - base is exact and distinct
        base = (rng.integers(-64, 64, size=d).astype(np.float32) * np.float32(2.0 ** -6)).astype(np.float32)
        models = []
        for k in range(K):
            rk = np.random.Generator(np.random.PCG64(seed + 1 + k))
            mag = (rk.permutation(d).astype(np.float32) + np.float32(1.0)) * np.float32(2.0 ** -20)
            sgn = np.where(rk.random(d) < 0.5, np.float32(-1.0), np.float32(1.0)).astype(np.float32)
            models.append((base + sgn * mag).astype(np.float32))
        return base, models
    base = (rng.standard_normal(d, dtype=np.float32) * np.float32(0.02)).astype(np.float32)
    models = []
    for k in range(K):
        rk = np.random.Generator(np.random.PCG64(seed + 1 + k))
        tv = np.float32(sigma) * rk.standard_normal(d, dtype=np.float32)
        if quantize:
            tv = (np.round(tv / np.float32(quantize)) * np.float32(quantize)).astype(np.float32)
        models.append((base + tv).astype(np.float32))
    return base, models


# ------------------------------------------------------------------------------ catalogs
def make_catalog(Q: int, N: int, E: int, kind: str = "grid", seed: int = 2,
                 planted: float = 0.5) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """User embeddings (Q,E), item table (N,E), labels (Q,) int64  (SURVEY.md section 8(d) config 1).

    kind="grid":  entries uniform in {-8..8}/8 -> every dot product is exact in fp32 and in TF32 and
                  independent of summation order; ties are plentiful (exercises the tie rule).
    kind="gauss": N(0,1) rows, L2-normalised (cosine similarity, configs/base.py:35).
    For a ``planted`` fraction of the queries u_q = item[label] + noise so metrics are non-trivial."""
    rng = np.random.Generator(np.random.PCG64(seed))
    labels = rng.integers(0, N, size=Q).astype(np.int64)
    plant = rng.random(Q) < planted
    if kind == "grid":
        items = (rng.integers(-8, 9, size=(N, E)).astype(np.float32) / np.float32(8.0)).astype(np.float32)
        users = (rng.integers(-8, 9, size=(Q, E)).astype(np.float32) / np.float32(8.0)).astype(np.float32)
        noise = rng.integers(-2, 3, size=(Q, E)).astype(np.float32) / np.float32(8.0)
        planted_u = np.clip(items[labels] + noise, -1.0, 1.0).astype(np.float32)
        users[plant] = planted_u[plant]
    elif kind == "gauss":
        items = rng.standard_normal((N, E), dtype=np.float32)
        users = rng.standard_normal((Q, E), dtype=np.float32)
        noise = rng.standard_normal((Q, E), dtype=np.float32)
        users[plant] = (items[labels] + np.float32(0.5) * noise)[plant]
        items = items / np.linalg.norm(items, axis=1, keepdims=True).astype(np.float32)
        users = users / np.linalg.norm(users, axis=1, keepdims=True).astype(np.float32)
        items, users = items.astype(np.float32), users.astype(np.float32)
    else:
        raise ValueError(kind)
    return users, items, labels


def make_distill_case(B: int, E: int, rows: Sequence[int], n_seq: int = 6, seed: int = 0, planted: float = 1.0):
    """Inputs of one distillation step: per-domain merged-model item tables, teacher
    item / sequence embeddings (un-normalised), the batch's (dataset index, sequence id) pairs and the merged model's
    representations.  Student tables are the teacher's plus noise so that losses are non-trivial."""
    rng = np.random.Generator(np.random.PCG64(seed))
    D = len(rows)
    t_items = [rng.standard_normal((n, E), dtype=np.float32) for n in rows]
    t_seqs = [rng.standard_normal((n_seq, E), dtype=np.float32) for _ in rows]
    tables = [(ti / np.linalg.norm(ti, axis=-1, keepdims=True) + np.float32(0.05) * rng.standard_normal(ti.shape, dtype=np.float32)).astype(np.float32)
              for ti in t_items]
    dataset_indexes = [int(x) for x in rng.integers(0, D, size=B)]
    sequence_ids = [int(x) for x in rng.integers(0, n_seq, size=B)]
    rep = np.stack([planted * t_seqs[d][s] / np.linalg.norm(t_seqs[d][s]) for d, s in zip(dataset_indexes, sequence_ids)]).astype(np.float32)
    rep = (rep + np.float32(0.3) * rng.standard_normal(rep.shape, dtype=np.float32)).astype(np.float32)
    return dict(tables=tables, teacher_items=t_items, teacher_seqs=t_seqs, dataset_indexes=dataset_indexes,
                sequence_ids=sequence_ids, rep=rep)
