"""Workloads of bench.py: each is one BASELINE.json configuration, built from synthetic inputs of the named
shapes.  Algorithmic byte counts follow SURVEY.md section 8(d) / DESIGN.md "Measurement"."""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import numpy as np
import torch

from mergerec_b200 import synth

GB = 1e9


class Workload:
    name = ""
    metric = ""
    unit = ""
    dtype = "f32"
    scaling = "weak"
    launches_per_step = 0
    e2e_steps_cap = 5
    h2d_bytes = 0
    d2h_bytes = 0

    def __init__(self, rank: int, world: int, device: Optional[torch.device]):
        self.rank, self.world, self.device = rank, world, device

    def extra(self) -> Dict:
        return {}


def ncu_traffic(kernel: str):
    """(bytes per launch, source file) of the committed ncu --set full capture of `kernel`, or (None, None)."""
    import json
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get(kernel)
        return (e["bytes"], e["source"]) if e else (None, None)
    except (OSError, ValueError, KeyError):
        return None, None


def _with_traffic(roof: Dict, kernel: str) -> Dict:
    roof["traffic"], src = ncu_traffic(kernel)
    if src:
        roof["traffic_source"] = src
    return roof


def _pinned(t: torch.Tensor) -> torch.Tensor:
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t)
    return out


class LambdaMergeK8(Workload):
    """Per-layer lambda merge of 8 BLaIR-base task vectors (the merge half of BASELINE config 2; A4).
    Algorithmic bytes per step: (K+2) * d * 4 (read base + K task-vector rows, write merged)."""

    name = "lambda_merge_k8"
    metric = "merge GB/s (algorithmic bytes / time)"
    unit = "GB/s"
    K = 8
    launches_per_step = 1

    def __init__(self, rank, world, device):
        super().__init__(rank, world, device)
        self.shapes = synth.roberta_shapes()
        self.d = synth.total_numel(self.shapes)
        self.bytes_per_step = (self.K + 2) * self.d * 4

    def config(self):
        return {"workload": "BLaIR-base (RoBERTa-base, d=124,645,632, P=199) K=8 per-layer lambda merge (G=13) "
                            "from stored task vectors; merge half of BASELINE config 2",
                "K": self.K, "d": self.d, "G": 13, "l2": "inputs (4.5 GB) exceed L2, no flush needed",
                "parallelism": f"replicas x{self.world}" if self.world > 1 else "1 GPU"}

    # -- device-resident arm
    def setup(self):
        from mergerec_b200.merger.algorithms import get_task_vectors
        from mergerec_b200.merger.layout import FlatLayout
        g = torch.Generator(device=self.device).manual_seed(1234 + self.rank)
        d, K = self.d, self.K
        self.base = torch.randn(d, generator=g, device=self.device) * 0.02
        self.models = [self.base + 1e-3 * torch.randn(d, generator=g, device=self.device) for _ in range(K)]
        self.T = get_task_vectors(self.base, self.models)
        self.rows = list(self.T.unbind(0))
        self.layout = FlatLayout.from_shape_dict(self.shapes)
        self.seg_end, self.seg_group, keys = self.layout.device_blocks(True, self.device)
        rng = np.random.Generator(np.random.PCG64(5))
        self.w = torch.from_numpy(rng.uniform(0.1, 0.5, size=(len(keys), K)).astype(np.float32)).to(self.device)
        self.out = torch.empty(d, dtype=torch.float32, device=self.device)

    def step(self):
        from mergerec_b200 import _lib
        from mergerec_b200.merger.algorithms._common import merge_axpy
        merge_axpy(self.base, self.rows, self.w, _lib.MR_ORDER_SUM_FIRST, False, self.seg_end, self.seg_group, out=self.out)

    def units_per_step_all_ranks(self):
        return self.bytes_per_step * self.world / GB

    # -- end-to-end arm: pinned host flat vectors in, merged flat vector out
    def setup_e2e(self):
        self.h_base = _pinned(self.base.cpu())
        self.h_T = _pinned(self.T.cpu())
        self.h_out = torch.empty(self.d, dtype=torch.float32, pin_memory=True)
        self.h2d_bytes = (self.K + 1) * self.d * 4
        self.d2h_bytes = self.d * 4

    def step_e2e(self):
        self.base.copy_(self.h_base, non_blocking=True)
        self.T.copy_(self.h_T, non_blocking=True)
        self.step()
        self.h_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # -- roofline of the dominant kernel (the only one here)
    def roofline(self, peaks):
        from bench import event_time_ms
        ms = event_time_ms(self.step, 20)
        achieved = self.bytes_per_step / GB / (ms * 1e-3)
        return _with_traffic({"bound": "hbm", "kernel": "mr::merge_kernel<8, SUM_FIRST, segmented, vec4>", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "ms_per_launch": ms,
                "algorithmic_bytes_per_launch": self.bytes_per_step}, "merge_kernel")

    # -- CPU arms (oracle port of the reference algorithm)
    def _cpu_inputs(self, d):
        rng = np.random.Generator(np.random.PCG64(7))
        base = rng.standard_normal(d, dtype=np.float32) * np.float32(0.02)
        T = rng.standard_normal((self.K, d), dtype=np.float32) * np.float32(1e-3)
        from oracle import oracle as orc
        sb, se, sg, keys = orc.segment_table(self.shapes, layer_wise=True)
        w = rng.uniform(0.1, 0.5, size=(len(keys), self.K)).astype(np.float32)
        return base, T, w, sb, se, sg

    def _cpu_time(self, reps):
        from oracle import oracle as orc
        base, T, w, sb, se, sg = self._cpu_inputs(self.d)
        orc.lambda_merge(base, T, w, sb, se, sg)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            orc.lambda_merge(base, T, w, sb, se, sg)
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)), orc.max_threads()

    def cpu_baseline(self):
        t, cores = self._cpu_time(3)
        return {"value": self.bytes_per_step / GB / t, "unit": self.unit, "cores": cores, "kind": "port",
                "sample": f"full workload (d={self.d}, K={self.K}), median of 3 after 1 warm-up, OpenMP C oracle",
                "seconds_per_step": t}

    def reference_arm(self, steps, warmup):
        t, cores = self._cpu_time(max(1, min(steps, 5)))
        return {"value": self.bytes_per_step / GB / t, "ms_per_step": t * 1e3, "cores": cores,
                "sample": f"full workload (d={self.d}, K={self.K}) per step, OpenMP C oracle port"}


class TiesCfg2(LambdaMergeK8):
    """BASELINE config 2: TIES merge (trim 20 %, sign election, disjoint mean) of 8 BLaIR-base domain models with
    per-layer lambda.  One step = the reference-shaped pipeline `get_ties_vectors(density=0.2)` (A6-A8) followed
    by the layer-wise lambda merge (A4).  Algorithmic bytes per step: (2K+1)*d*4 + (K+2)*d*4."""

    name = "ties_cfg2"
    launches_per_step = 12  # own kernels: init, 3 x pass, 3 x pick, cand_hist, compact, final, build, merge

    def __init__(self, rank, world, device):
        super().__init__(rank, world, device)
        self.bytes_build = (2 * self.K + 1) * self.d * 4
        self.bytes_merge = (self.K + 2) * self.d * 4
        self.bytes_per_step = self.bytes_build + self.bytes_merge

    def config(self):
        return {"workload": "BASELINE config 2: TIES merge (density 0.2, global trim, sign election, disjoint mean) "
                            "of K=8 BLaIR-base (RoBERTa-base, d=124,645,632, P=199) domain models + per-layer "
                            "lambda merge (G=13)",
                "K": self.K, "d": self.d, "G": 13, "density": 0.2,
                "l2": "inputs (4.5 GB) exceed L2, no flush needed", "cuda_graph": getattr(self, "graph", None) is not None,
                "select_status": "checked on the host after the timed loop (stream-ordered step)",
                "parallelism": f"replicas x{self.world}" if self.world > 1 else "1 GPU"}

    def setup(self):
        super().setup()
        from mergerec_b200.merger.layout import alloc_rows
        self.That = alloc_rows(self.K, self.d, self.device)
        self.Trows = list(self.That.unbind(0))
        del self.T, self.rows
        self._status = None
        self._try_capture()

    def _eager_step(self, defer=True):
        """select -> build -> per-layer merge, stream-ordered: the select's status word is checked by the host after
        the timed loop (`finish`), not between the kernels (`defer=False` is the public API's immediate check)."""
        from mergerec_b200 import _lib
        from mergerec_b200.merger.algorithms import ties as T
        from mergerec_b200.merger.algorithms._common import merge_axpy
        if defer:
            cut, self._status = T.select_kth_largest(self.base, self.models, int(0.2 * self.d), None, defer_status=True)
        else:
            cut = T.ties_select(self.base, self.models, 0.2)
        T._build(self.base, self.models, cut, _lib.MR_TIES_VECTORS, out=self.That, ldo=self.That.stride(0))
        merge_axpy(self.base, self.Trows, self.w, _lib.MR_ORDER_SUM_FIRST, False, self.seg_end, self.seg_group, out=self.out)

    def _try_capture(self):
        """One CUDA graph for the whole step (12 launches of this package, all stream-ordered)."""
        self.graph = None
        if os.environ.get("MR_BENCH_NO_GRAPH"):
            return
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self._eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._eager_step()
            g.replay()
            torch.cuda.synchronize()
            self.finish()
            self.graph = g
        except Exception as e:  # noqa: BLE001 -- report, keep the eager path
            print(f"[ties_cfg2] CUDA graph capture unavailable ({type(e).__name__}: {e}); eager launches", file=__import__("sys").stderr)
            self.graph = None
            torch.cuda.synchronize()

    def step(self):
        if getattr(self, "graph", None) is not None:
            self.graph.replay()
        else:
            self._eager_step()

    def finish(self):
        """After a timed loop: the deferred status of the last select (every step ran the same inputs)."""
        from mergerec_b200.merger.algorithms import ties as T
        if getattr(self, "_status", None) is not None:
            T.verify_select_status(self._status)

    def step_fused(self):
        from mergerec_b200.merger.algorithms import ties as T
        T.merge_ties_lambda(self.base, self.models, 0.2, self.w, self.seg_end, self.seg_group, out=self.out)

    def setup_e2e(self):
        self.h_base = _pinned(self.base.cpu())
        self.h_models = [_pinned(m.cpu()) for m in self.models]
        self.h_out = torch.empty(self.d, dtype=torch.float32, pin_memory=True)
        self.h2d_bytes = (self.K + 1) * self.d * 4
        self.d2h_bytes = self.d * 4

    def step_e2e(self):
        self.base.copy_(self.h_base, non_blocking=True)
        for m, h in zip(self.models, self.h_models):
            m.copy_(h, non_blocking=True)
        self._e2e_kernels()
        self.h_out.copy_(self.out, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def _e2e_kernels(self):
        self._eager_step(defer=False)      # the public get_ties_vectors path: status checked on the host right away

    def roofline(self, peaks):
        from bench import event_time_ms
        from mergerec_b200 import _lib
        from mergerec_b200.merger.algorithms import ties as T
        cut = T.ties_select(self.base, self.models, 0.2)
        lib = _lib.load()
        K, d = self.K, self.d
        ws_bytes = int(lib.mr_ties_workspace_bytes(d, K))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
        cut2 = torch.empty(K, dtype=torch.int64, device=self.device)
        status = torch.zeros(K, dtype=torch.int32, device=self.device)
        parr = _lib.ptr_array(self.models)

        def select():
            _lib.check(lib.mr_ties_select(_lib.dptr(self.base), parr, K, d, None, int(0.2 * d), _lib.dptr(cut2),
                                          _lib.dptr(status), _lib.dptr(ws), ws_bytes, _lib.stream_handle()), "select")

        def build():
            T._build(self.base, self.models, cut, _lib.MR_TIES_VECTORS, out=self.That, ldo=self.That.stride(0))

        # lambda-gradient reduction (A5; the backward kernel of BASELINE config 3) against the same K x d rows
        from mergerec_b200.merger.weight_learning.module._base import _lambda_grad
        grad = torch.randn(self.d, device=self.device)
        seg_off, seg_len = self.layout.device_segments(self.device)
        grads = [grad[o:o + n] for o, n in zip(seg_off.tolist(), seg_len.tolist())]
        lgrad = lambda: _lambda_grad(grads, self.layout, self.That, self.seg_group, int(self.w.shape[0]))  # noqa: E731
        lgrad()
        ms_lgrad = event_time_ms(lgrad, 5)
        lgrad_bytes = (K + 1) * d * 4
        del grad, grads

        ms_build = event_time_ms(build, 10)
        ms_select = event_time_ms(select, 10)
        ms_merge = event_time_ms(self._merge_only, 10)
        ms_fused = event_time_ms(self.step_fused, 5)
        ach = self.bytes_build / GB / (ms_build * 1e-3)
        sel_bytes = (K + 1) * d * 4
        return _with_traffic({"bound": "hbm", "kernel": "mr::ties_build_kernel<8, VECTORS, vec4> (get_ties_vectors build pass)",
                "achieved": ach, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "ms_per_launch": ms_build,
                "algorithmic_bytes_per_launch": self.bytes_build,
                "other_kernels": {
                    "ties_select (sample + 1 full pass + finish)": {"ms": ms_select, "GB/s": sel_bytes / GB / (ms_select * 1e-3), "bytes": sel_bytes},
                    "lambda-gradient reduction (lambda_grad_kernel, incl. host pointer-table upload)": {"ms": ms_lgrad, "GB/s": lgrad_bytes / GB / (ms_lgrad * 1e-3), "bytes": lgrad_bytes},
                    "lambda merge (merge_kernel)": {"ms": ms_merge, "GB/s": self.bytes_merge / GB / (ms_merge * 1e-3), "bytes": self.bytes_merge},
                    "fused select + build + merge without materialising That (merge_ties_lambda)": {"ms": ms_fused, "GB/s": self.bytes_merge / GB / (ms_fused * 1e-3), "bytes": self.bytes_merge},
                }}, "ties_build_kernel")

    def extra(self):
        """The second hot path, reported beside the merger line: a short run of the evaluator workload
        (`--workload eval_cfg5` is the full bench of it).  Single GPU only; failures are reported, not raised."""
        if self.world != 1 or self.rank != 0 or os.environ.get("MR_BENCH_SKIP_EVAL_EXTRA"):
            return {}
        try:
            from bench import event_time_ms, measured_peaks
            for name in ("base", "models", "That", "Trows", "out"):   # free the merger's 9 GB first
                if hasattr(self, name):
                    delattr(self, name)
            torch.cuda.empty_cache()
            ev = EvalCatalog(0, 1, self.device)   # BASELINE config 5 at full size (about 0.4 s per step)
            ev.setup()
            ev.step()
            torch.cuda.synchronize()
            ms = event_time_ms(ev.step, 3)
            roof = ev.roofline(measured_peaks())
            return {"evaluator": {"metric": ev.metric, "value": ev.Q * ev.N / (ms * 1e-3), "unit": ev.unit,
                                  "eval_seqs_per_s": ev.Q / (ms * 1e-3), "ms_per_step": ms, "config": ev.config(),
                                  "metrics": ev.last, "roofline": roof}}
        except Exception as e:  # noqa: BLE001
            return {"evaluator": {"error": repr(e)}}

    def _merge_only(self):
        from mergerec_b200 import _lib
        from mergerec_b200.merger.algorithms._common import merge_axpy
        merge_axpy(self.base, self.Trows, self.w, _lib.MR_ORDER_SUM_FIRST, False, self.seg_end, self.seg_group, out=self.out)

    def _cpu_time(self, reps):
        """Oracle port on a bounded sample: the full K = 8 but a prefix of the flat vector (the select is global
        over whatever vector it is given, so the per-element work is the same)."""
        from oracle import oracle as orc
        frac = 8
        d = self.d // frac
        rng = np.random.Generator(np.random.PCG64(7))
        base = rng.standard_normal(d, dtype=np.float32) * np.float32(0.02)
        models = [base + np.float32(1e-3) * rng.standard_normal(d, dtype=np.float32) for _ in range(self.K)]
        w = rng.uniform(0.1, 0.5, size=(1, self.K)).astype(np.float32)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            That = orc.ties_vectors(base, models, 0.2)
            orc.lambda_merge(base, That, w)
            ts.append(time.perf_counter() - t0)
        self._cpu_sample = f"d/{frac} = {d} flat elements of the K={self.K} workload (select is global over the sample), OpenMP C oracle port"
        return float(np.median(ts)) * frac, orc.max_threads()

    def cpu_baseline(self):
        t, cores = self._cpu_time(2)
        return {"value": self.bytes_per_step / GB / t, "unit": self.unit, "cores": cores, "kind": "port",
                "sample": self._cpu_sample + "; time scaled x8 to the full vector", "seconds_per_step": t}

    def reference_arm(self, steps, warmup):
        t, cores = self._cpu_time(max(1, min(steps, 3)))
        return {"value": self.bytes_per_step / GB / t, "ms_per_step": t * 1e3, "cores": cores,
                "sample": self._cpu_sample + "; time scaled x8 to the full vector"}


class EvalCatalog(Workload):
    """BASELINE config 5 (evaluator): full-catalog scoring of Q query embeddings against an N-item table (E = 768),
    fused per-row top-100, label rank and Recall/NDCG.  The item table is sharded over the ranks (N / world rows
    each, strong scaling); every rank scores all queries against its shard, the per-rank top-K lists are
    allgathered over NCCL and merged.  One step = one pass of all Q queries over the whole catalog.

    Default sizes are config 5 itself: Q = 65,536 queries, N = 1,000,000 items, E = 768, top-100 (1.0e14 multiply-adds
    worth of logical FLOPs, x3 tensor passes); MR_BENCH_EVAL_Q / _N / _K / _E override them."""

    name = "eval_cfg5"
    metric = "catalog scores/sec (Q*N / time, fused scoring + top-K + Recall/NDCG)"
    unit = "scores/s"
    dtype = "tf32x3 (fp32-faithful split, fp32 accumulate)"
    scaling = "strong"
    e2e_steps_cap = 3

    def __init__(self, rank, world, device):
        super().__init__(rank, world, device)
        self.Q = int(os.environ.get("MR_BENCH_EVAL_Q", 65536))
        self.N = int(os.environ.get("MR_BENCH_EVAL_N", 1_000_000))
        self.E = int(os.environ.get("MR_BENCH_EVAL_E", 768))
        self.K = int(os.environ.get("MR_BENCH_EVAL_K", 100))
        self.mode = int(os.environ.get("MR_BENCH_EVAL_MODE", 0))
        self.flops = 2.0 * self.Q * self.N * self.E
        self.launches_per_step = 4 if world == 1 else 5  # split_tf32(users), score_topk, topk_merge, [shard merge], label_rank

    def config(self):
        return {"workload": f"BASELINE config 5: {self.Q} query seqs x {self.N}-item catalog, E={self.E}, "
                            f"top-{self.K}, Recall/NDCG@{{10,{self.K}}}; item table sharded over {self.world} GPU(s), "
                            "NCCL allgather of per-GPU top-K + merge",
                "Q": self.Q, "N": self.N, "E": self.E, "K": self.K, "mode": "tf32x3" if self.mode == 0 else "tf32x1",
                "l2": "item table (3.1 GB x2 hi/lo at N=1M) exceeds L2, no flush needed",
                "parallelism": f"item-sharded x{self.world}" if self.world > 1 else "1 GPU"}

    def setup(self):
        import torch.distributed as dist
        from mergerec_b200.evaluator import Evaluator, ShardedItemTable, shard_bounds
        # every rank generates the same full table chunk by chunk (seeded), keeps its own rows and the rows of the
        # labels: half of the queries are planted near their label item so Recall / NDCG are non-trivial
        g = torch.Generator(device=self.device).manual_seed(99)       # same queries and labels on every rank
        users = torch.randn(self.Q, self.E, generator=g, device=self.device)
        self.labels = torch.randint(0, self.N, (self.Q,), generator=g, device=self.device)
        planted = torch.rand(self.Q, generator=g, device=self.device) < 0.5
        lo, hi = shard_bounds(self.N, self.world, self.rank)
        self.items = torch.empty(hi - lo, self.E, device=self.device)
        label_rows = torch.empty(self.Q, self.E, device=self.device)
        chunk = 65536
        for c0 in range(0, self.N, chunk):
            c1 = min(self.N, c0 + chunk)
            gi = torch.Generator(device=self.device).manual_seed(1000 + c0 // chunk)
            rows = torch.nn.functional.normalize(torch.randn(c1 - c0, self.E, generator=gi, device=self.device), dim=-1)
            a, b = max(c0, lo), min(c1, hi)
            if a < b:
                self.items[a - lo:b - lo] = rows[a - c0:b - c0]
            sel = (self.labels >= c0) & (self.labels < c1)
            if bool(sel.any()):
                label_rows[sel] = rows[self.labels[sel] - c0]
        users = torch.where(planted[:, None], label_rows + 0.5 * users / self.E ** 0.5, users)
        self.users = torch.nn.functional.normalize(users, dim=-1)
        del label_rows, users
        self.lo = lo
        self.group = dist.group.WORLD if self.world > 1 else None
        self.table = ShardedItemTable(self.items, id_base=lo, n_total=self.N, group=self.group)
        self.ev = Evaluator(["RECALL", "NDCG"], [10, self.K])
        self.last = None

    def step(self):
        self.last = self.ev.evaluate_embeddings(self.users, self.table, self.labels, mode=self.mode)

    def units_per_step_all_ranks(self):
        return float(self.Q) * float(self.N)

    def setup_e2e(self):
        self.h_users = _pinned(self.users.cpu())
        self.h_items = _pinned(self.items.cpu())
        self.h_labels = _pinned(self.labels.cpu())
        self.h2d_bytes = self.h_users.numel() * 4 + self.h_items.numel() * 4 + self.h_labels.numel() * 8
        self.d2h_bytes = self.Q * 4          # the rank of each label; the metric floats are finished on the host

    def step_e2e(self):
        from mergerec_b200.evaluator import ShardedItemTable
        users = self.h_users.to(self.device, non_blocking=True)
        items = self.h_items.to(self.device, non_blocking=True)
        labels = self.h_labels.to(self.device, non_blocking=True)
        table = ShardedItemTable(items, id_base=self.lo, n_total=self.N, group=self.group)
        self.last = self.ev.evaluate_embeddings(users, table, labels, mode=self.mode)

    def roofline(self, peaks):
        from bench import event_time_ms
        from mergerec_b200.evaluator.evaluator import score_topk
        from mergerec_b200.evaluator.sharded import split_tf32
        u_hi, u_lo = split_tf32(self.users)
        ms = event_time_ms(lambda: score_topk(u_hi, u_lo, self.table, self.K, self.mode), 2)
        passes = 3 if self.mode == 0 else 1
        local_flops = 2.0 * self.Q * self.table.n_local * self.E
        ach = passes * local_flops / (ms * 1e-3) / 1e12
        # a launch of a few ms runs at boost clocks (burst peak); one of hundreds of ms sits at the 1 kW power cap
        # like the driver's sustained cuBLAS loop (sustained peak).  TF32 runs at half the bf16 rate.
        sustained = ms > 100.0
        peak = (peaks["bf16_tflops_sustained"] if sustained else peaks["bf16_tflops"]) / 2
        return {"bound": "tensor", "kernel": "mr::st::score_topk_kernel<2, 32> (tcgen05.mma kind::tf32, cta_group::2) + list merge",
                "achieved": ach, "peak": peak,
                "peak_source": peaks["source"] + (": sustained" if sustained else ": burst") + " bf16 cuBLAS / 2 (tf32 runs at half the bf16 rate)",
                "unit": "TFLOP/s", "frac": ach / peak, "traffic": None, "ms_per_launch": ms,
                "tensor_passes": passes, "logical_tflops": local_flops / (ms * 1e-3) / 1e12,
                "algorithmic_flops_per_launch": passes * local_flops,
                "frac_of_burst_peak": ach / (peaks["bf16_tflops"] / 2),
                "frac_of_sustained_peak": ach / (peaks["bf16_tflops_sustained"] / 2)}

    def extra(self):
        return {"metrics": self.last}

    # -- CPU arms: fp32 sgemm + canonical top-K + metrics (oracle port of module.py:137 + evaluator.py:31-49)
    def _cpu_time(self, reps):
        from oracle import oracle as orc
        Qs, Ns = min(self.Q, 256), min(self.N, 100_000)
        rng = np.random.Generator(np.random.PCG64(7))
        users = rng.standard_normal((Qs, self.E), dtype=np.float32)
        items = rng.standard_normal((Ns, self.E), dtype=np.float32)
        labels = rng.integers(0, Ns, size=Qs)
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            scores = orc.scores_f32(users, items)
            orc.evaluate(scores, labels, ["RECALL", "NDCG"], [10, min(self.K, Ns)])
            ts.append(time.perf_counter() - t0)
        t = float(np.median(ts[1:]))
        self._cpu_sample = (f"{Qs} queries x {Ns} items (E={self.E}) of the workload: numpy/BLAS sgemm + OpenMP C top-K + "
                            "python metric loops (oracle port); rate scaled linearly")
        return t, Qs * Ns, max(orc.max_threads(), os.cpu_count() or 1)

    def cpu_baseline(self):
        t, n, cores = self._cpu_time(2)
        return {"value": n / t, "unit": self.unit, "cores": cores, "kind": "port", "sample": self._cpu_sample,
                "seconds_per_sample": t}

    def reference_arm(self, steps, warmup):
        t, n, cores = self._cpu_time(max(1, min(steps, 3)))
        return {"value": n / t, "ms_per_step": t * 1e3, "cores": cores, "sample": self._cpu_sample}


class CollabStepCfg3(Workload):
    """BASELINE config 3: one MergeRec collaborative-merging optimisation step at K = 8 BLaIR-base (RoBERTa-base)
    domains -- layer-wise lambda merge (A4) into one flat buffer whose slices become the encoder's parameters,
    encoder forward + backward in PyTorch on a pseudo-user batch (16 sequences x 512 tokens), lambda-gradient
    reduction (A5), Adam on lambda (stack B of SURVEY.md section 3; `load_merging_module` + `DistillSequenceModule`'s
    optimiser set-up, sequence/module.py:94-100).  The two kernels of this package are timed inside the step and
    alone; the encoder is stock HF RoBERTa under bf16 autocast like the reference's Trainer (configs/base.py:41)."""

    name = "collab_cfg3"
    metric = "collaborative-merging steps/sec (lambda merge + encoder fwd/bwd + lambda-gradient + Adam)"
    unit = "steps/s"
    dtype = "f32 merge / lambda-gradient, bf16-autocast encoder"
    K = 8
    launches_per_step = 3   # merge_kernel, lg_partial_kernel, lg_finish kernel (the encoder's kernels are PyTorch's)
    e2e_steps_cap = 3

    def __init__(self, rank, world, device):
        super().__init__(rank, world, device)
        self.B, self.L = 16, 512
        self.shapes = synth.roberta_shapes()
        self.d = synth.total_numel(self.shapes)

    def config(self):
        return {"workload": "BASELINE config 3: collaborative merging step, K=8 BLaIR-base (RoBERTa-base, d=124,645,632), "
                            "layer-wise lambda (G=13), batch 16 x 512 tokens, Adam(lr 1e-3) on lambda",
                "K": self.K, "d": self.d, "batch": self.B, "seq_len": self.L,
                "l2": "4.5 GB of task vectors per merge / gradient pass exceed L2",
                "parallelism": f"data-parallel replicas x{self.world} (lambda-gradient all-reduce of G x K floats)"
                if self.world > 1 else "1 GPU"}

    def setup(self):
        import torch.distributed as dist
        from transformers import RobertaConfig, RobertaModel
        from mergerec_b200.merger.enums import LearnType, MergeType
        from mergerec_b200.merger.weight_learning.module import load_merging_module

        class Wrapped(torch.nn.Module):      # the reference's BaseEncoderModel keeps the HF model under `.model`
            def __init__(self, hf):
                super().__init__()
                self.model = hf

            def forward(self, ids):
                return self.model(input_ids=ids).last_hidden_state[:, 0]   # CLS pooling (encoder/_base.py:32-39)

        torch.manual_seed(0)
        cfg = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1,
                            layer_norm_eps=1e-5, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        model = Wrapped(RobertaModel(cfg, add_pooling_layer=True)).to(self.device)
        pre = {k: v.detach().clone() for k, v in model.state_dict().items()}
        g = torch.Generator(device=self.device).manual_seed(100 + self.rank)
        fts = [{k: (v + 1e-3 * torch.randn(v.shape, generator=g, device=self.device)) if v.is_floating_point() else v.clone()
                for k, v in pre.items()} for _ in range(self.K)]
        self.module = load_merging_module(MergeType.TASK_VECTOR, LearnType.LAYER_WISE, model, pre, fts, ignore_keys=set(),
                                          initial_per_weight=0.2, disable_softmax=True)
        del fts, pre
        self.opt = torch.optim.Adam(self.module.trainable_parameters(True, True, False), lr=1e-3)
        self.ids = torch.randint(3, 50000, (self.B, self.L), generator=g, device=self.device)
        self.group = dist.group.WORLD if self.world > 1 else None
        self.loss = None

    def step(self):
        import torch.distributed as dist
        self.opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            rep = self.module(self.ids)
        loss = rep.float().square().mean()
        loss.backward()
        if self.group is not None:
            for p in self.module.per_weights.parameters():
                dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=self.group)
        self.opt.step()
        self.loss = loss.detach()

    def units_per_step_all_ranks(self):
        return float(self.world)

    def setup_e2e(self):
        self.h_ids = _pinned(self.ids.cpu())
        self.h2d_bytes = self.h_ids.numel() * 8
        self.d2h_bytes = 4

    def step_e2e(self):
        self.ids.copy_(self.h_ids, non_blocking=True)
        self.step()
        float(self.loss)      # device -> host read of the step's loss

    def roofline(self, peaks):
        from bench import event_time_ms
        from mergerec_b200.merger.weight_learning.module._base import _lambda_grad
        m = self.module
        w = m._effective_weights().detach()
        ms_merge = event_time_ms(lambda: m._merge_flat(w), 10)
        _, seg_group, keys = m._blocks()
        grad = torch.randn(self.d, device=self.device)
        grads = [grad[o:o + n] for o, n in zip(m._layout.offsets, m._layout.sizes)]
        ms_grad = event_time_ms(lambda: _lambda_grad(grads, m._layout, m._task_rows(), seg_group, len(keys)), 10)
        # rank 0 only from here: with several ranks `step` holds an all-reduce, so time it only on one GPU
        ms_step = event_time_ms(self.step, 3) if self.world == 1 else None
        bytes_merge, bytes_grad = (self.K + 2) * self.d * 4, (self.K + 1) * self.d * 4
        ach = bytes_grad / GB / (ms_grad * 1e-3)
        return _with_traffic({"bound": "hbm", "kernel": "mr::lg_partial_kernel<8> (+ finish): lambda-gradient reduction (A5)",
                "achieved": ach, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "ms_per_launch": ms_grad,
                "algorithmic_bytes_per_launch": bytes_grad,
                "other_kernels": {"lambda merge forward (merge_kernel, A4)": {"ms": ms_merge, "GB/s": bytes_merge / GB / (ms_merge * 1e-3), "bytes": bytes_merge}},
                "step_ms": ms_step, "merge_plus_grad_share_of_step": None if ms_step is None else (ms_merge + ms_grad) / ms_step}, "lg_partial_kernel")

    def extra(self):
        return {"loss": None if self.loss is None else float(self.loss),
                "per_weights": {k: [round(x, 6) for x in v.tolist()] for k, v in list(self.module.per_weights.items())[:2]}}

    # -- CPU arm: the oracle port of the two merger kernels of the step (the encoder is PyTorch on both sides)
    def _cpu_time(self, reps):
        from oracle import oracle as orc
        rng = np.random.Generator(np.random.PCG64(7))
        d = self.d
        base = rng.standard_normal(d, dtype=np.float32) * np.float32(0.02)
        T = rng.standard_normal((self.K, d), dtype=np.float32) * np.float32(1e-3)
        sb, se, sg, keys = orc.segment_table(self.shapes, layer_wise=True)
        w = rng.uniform(0.1, 0.5, size=(len(keys), self.K)).astype(np.float32)
        grad = rng.standard_normal(d, dtype=np.float32)
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            orc.lambda_merge(base, T, w, sb, se, sg)
            orc.lambda_grad(grad, T, len(keys), sb, se, sg)
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts[1:])), orc.max_threads()

    def cpu_baseline(self):
        t, cores = self._cpu_time(2)
        return {"value": 1.0 / t, "unit": self.unit, "cores": cores, "kind": "port",
                "sample": "the step's two merger passes only (lambda merge + lambda-gradient, full d, K=8), OpenMP C oracle "
                          "port; the encoder forward/backward is not included on the CPU side",
                "seconds_per_step": t}

    def reference_arm(self, steps, warmup):
        t, cores = self._cpu_time(max(1, min(steps, 3)))
        return {"value": 1.0 / t, "ms_per_step": t * 1e3, "cores": cores,
                "sample": "the step's two merger passes only (lambda merge + lambda-gradient, full d, K=8), OpenMP C oracle port"}


class TiesSharded(TiesCfg2):
    """SURVEY.md section 8(e), merger row: the TIES merge of BASELINE config 2 with the flat vector SHARDED over the
    ranks (d/G columns per GPU, 32-column aligned).  One step = global trim threshold (per-slice order statistic, then two
    windowed radix levels of `mr_ties_mag_hist` with one NCCL all-reduce of K x 2049 int64 each + two all-gathers of K values), TIES build of
    the local slice, task-wise lambda merge of the local slice.  Strong scaling: the whole job's algorithmic bytes
    ((2K+1) + (K+2)) * d * 4 divided by the slowest rank's time.  Results are bit-identical to the single-GPU merge
    (tests/test_sharded_merger_gpu.py)."""

    name = "ties_sharded"
    scaling = "strong"
    launches_per_step = 14  # per-slice select (10), 2 x mag_hist, build, merge (plus torch's tiny histogram-walk ops and the collectives)

    def config(self):
        return {"workload": "TIES (density 0.2) of K=8 BLaIR-base models, flat vector sharded d/G per GPU; global trim via "
                            "all-reduced radix histograms; local build + task-wise lambda merge",
                "K": self.K, "d": self.d, "density": 0.2, "l2": "per-rank inputs (4.5 GB / G) exceed L2 for G <= 8",
                "parallelism": f"flat dimension sharded x{self.world}, 2 all-reduces of 131 KB + 2 small all-gathers per step"
                if self.world > 1 else "1 GPU (same code path, no collective)"}

    def setup(self):
        import torch.distributed as dist
        from mergerec_b200.merger.layout import alloc_rows
        from mergerec_b200.merger.sharded import flat_shard_bounds
        g = torch.Generator(device=self.device).manual_seed(1234)    # same vectors on every rank, then keep the slice
        d, K = self.d, self.K
        lo, hi = flat_shard_bounds(d, self.world, self.rank)
        self.lo, self.hi = lo, hi
        base = torch.randn(d, generator=g, device=self.device) * 0.02
        self.models = []
        for _ in range(K):
            self.models.append((base + 1e-3 * torch.randn(d, generator=g, device=self.device))[lo:hi].clone())
        self.base = base[lo:hi].clone()
        del base
        self.group = dist.group.WORLD if self.world > 1 else None
        self.That = alloc_rows(K, hi - lo, self.device)
        self.Trows = list(self.That.unbind(0))
        self.w = torch.full((1, K), 0.3, dtype=torch.float32, device=self.device)
        self.out = torch.empty(hi - lo, dtype=torch.float32, device=self.device)

    def _select(self):
        from mergerec_b200.merger.sharded import sharded_select
        return sharded_select(self.base, self.models, int(0.2 * self.d), self.d, None, self.group)

    def _build(self, cut):
        from mergerec_b200 import _lib
        from mergerec_b200.merger.algorithms import ties as T
        T._build(self.base, self.models, cut, _lib.MR_TIES_VECTORS, out=self.That, ldo=max(self.That.stride(0), self.hi - self.lo))

    def _merge_only(self):
        from mergerec_b200 import _lib
        from mergerec_b200.merger.algorithms._common import merge_axpy
        merge_axpy(self.base, self.Trows, self.w, _lib.MR_ORDER_SUM_FIRST, False, out=self.out)

    def step(self):
        self._cut = self._select()
        self._build(self._cut)
        self._merge_only()

    def _e2e_kernels(self):
        self.step()

    def units_per_step_all_ranks(self):
        return self.bytes_per_step / GB      # strong scaling: one job, whatever the number of ranks

    def setup_e2e(self):
        # (every rank passes here once, after the timed loop) exact integer checksum of the merged vector's bit
        # patterns summed over the ranks: independent of the sharding because the sharded merge is bit-identical to
        # the single-GPU one -- compare the lines of different --gpus
        chk = self.out.view(torch.int32).to(torch.int64).sum().reshape(1)
        if self.group is not None:
            import torch.distributed as dist
            dist.all_reduce(chk, group=self.group)
        self._checksum = int(chk.item())
        n = self.hi - self.lo
        self.h_base = _pinned(self.base.cpu())
        self.h_models = [_pinned(m.cpu()) for m in self.models]
        self.h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
        self.h2d_bytes = (self.K + 1) * n * 4
        self.d2h_bytes = n * 4

    def roofline(self, peaks):
        from bench import event_time_ms
        from mergerec_b200.merger.sharded import CudaKernels
        n, K = self.hi - self.lo, self.K
        cut = self._cut                      # rank 0 only from here on: kernels, no collectives
        hist = torch.zeros((K, 2048), dtype=torch.int64, device=self.device)
        above = torch.zeros(K, dtype=torch.int64, device=self.device)
        lo = ((cut >> 32) - 64).clamp(min=0).to(torch.int32)                  # the first window of this (iid) workload
        sh = torch.zeros(K, dtype=torch.int32, device=self.device)
        ms_hist = event_time_ms(lambda: CudaKernels.mag_hist(self.base, self.models, None, lo, sh, hist, above), 10)
        lo_w = ((cut >> 32) - (1 << 17)).clamp(min=0).to(torch.int32)         # a wide window (shards with different statistics)
        sh_w = torch.full((K,), 7, dtype=torch.int32, device=self.device)
        ms_hist_wide = event_time_ms(lambda: CudaKernels.mag_hist(self.base, self.models, None, lo_w, sh_w, hist, above), 10)
        ms_select = event_time_ms(self._select, 5) if self.world == 1 else None
        ms_build = event_time_ms(lambda: self._build(cut), 10)
        ms_merge = event_time_ms(self._merge_only, 10)
        b_build, b_hist, b_merge = (2 * K + 1) * n * 4, (K + 1) * n * 4, (K + 2) * n * 4
        ach = b_build / GB / (ms_build * 1e-3)
        return {"bound": "hbm", "kernel": "mr::ties_build_kernel<8, VECTORS, vec4> on this rank's slice",
                "achieved": ach, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "ms_per_launch": ms_build,
                "algorithmic_bytes_per_launch": b_build,
                "other_kernels": {
                    "mag_hist_kernel (one windowed radix level)": {"ms": ms_hist, "GB/s": b_hist / GB / (ms_hist * 1e-3), "bytes": b_hist},
                    "mag_hist_kernel, window 2^18 bit patterns wide (1.3 % of the elements inside)": {"ms": ms_hist_wide, "GB/s": b_hist / GB / (ms_hist_wide * 1e-3)},
                    "sharded select (per-slice order statistic + windowed level(s) + all-reduce + tie scan)": {"ms": ms_select},
                    "lambda merge of the slice (merge_kernel)": {"ms": ms_merge, "GB/s": b_merge / GB / (ms_merge * 1e-3), "bytes": b_merge}}}

    def extra(self):
        return {"slice": [self.lo, self.hi], "merged_bits_checksum": self._checksum}


class DistillStep(Workload):
    """The distillation half of a collaborative-merging step (reference:
    module/distiller/sequence/module.py:59-76 + loss_fn.py KD loss): catalogue logits of 16 pseudo-user representations
    against their domains' item tables (8 domains x 25,000 items, E = 768 -- the per-domain share of BASELINE config 4's
    200K-item union catalogue), KD loss against device-resident teacher logits, gradient w.r.t. the representations.
    Algorithmic bytes per step: every used item table once in the forward and once in the backward pass."""

    name = "distill_step"
    metric = "distillation samples/sec (catalogue logits + KD loss + representation gradient)"
    unit = "samples/s"
    dtype = "f32"
    launches_per_step = 4     # ds_logits, ds_loss, ds_grad, ds_grad_finish
    D, N, E, B, NSEQ = 8, 25000, 768, 16, 512

    def config(self):
        return {"workload": "distillation step: 16 samples over 8 domains x 25,000 items, E=768, KD loss (T=2), "
                            "teacher logits device-resident (512 sequences per domain)",
                "B": self.B, "domains": self.D, "items_per_domain": self.N, "E": self.E,
                "l2": "item tables (614 MB) exceed L2", "cuda_graph": getattr(self, "graph", None) is not None,
                "parallelism": f"data-parallel replicas x{self.world}" if self.world > 1 else "1 GPU"}

    def setup(self):
        from mergerec_b200.module.distiller import TeacherScores
        from mergerec_b200.module.recommender.loss_fn import DistillKDLoss
        g = torch.Generator(device=self.device).manual_seed(77 + self.rank)
        mk = lambda *shape: torch.randn(*shape, generator=g, device=self.device)
        t_items = [mk(self.N, self.E) for _ in range(self.D)]
        t_seqs = [mk(self.NSEQ, self.E) for _ in range(self.D)]
        self.teacher = TeacherScores(t_items, t_seqs)
        self.tables = [torch.nn.functional.normalize(t, dim=-1) + 0.02 * mk(self.N, self.E) for t in t_items]
        self.dom = [b % self.D for b in range(self.B)]
        self.seq_ids = [int(x) for x in torch.randint(0, self.NSEQ, (self.B,), generator=g, device=self.device).tolist()]
        self.rep = torch.nn.functional.normalize(torch.stack([t_seqs[d][s] for d, s in zip(self.dom, self.seq_ids)]) + 0.3 * mk(self.B, self.E), dim=-1)
        self.rep.requires_grad_(True)
        self.spec = DistillKDLoss(2.0).spec
        _, self.ptrs = self.teacher.rows(self.dom, self.seq_ids)
        self.bytes_tables = self.D * self.N * self.E * 4
        self.loss = None
        self._try_capture()

    def _eager_step(self):
        from mergerec_b200.module.distiller.sequence.module import fused_distill_losses
        self.rep.grad = None
        loss = fused_distill_losses(self.rep, self.tables, self.dom, self.ptrs, self.spec).mean()
        loss.backward()
        self.loss = loss.detach()

    def _try_capture(self):
        """The step is launch-bound (about ten launches for ~0.23 ms of GPU work): capture forward + backward in one CUDA
        graph and replay it.  Every entry point of the C ABI is stream-ordered and takes its small tables by value, so
        the capture needs nothing special; if it fails for any reason the eager step stays."""
        self.graph = None
        if os.environ.get("MR_BENCH_NO_GRAPH"):
            return
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._eager_step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            self.rep.grad = None
            with torch.cuda.graph(g):
                self._eager_step()
            g.replay()
            torch.cuda.synchronize()
            self.graph = g
        except Exception as e:  # noqa: BLE001 -- report, keep the eager path
            print(f"[distill_step] CUDA graph capture unavailable ({type(e).__name__}: {e}); eager launches", file=__import__("sys").stderr)
            self.graph = None
            torch.cuda.synchronize()

    def step(self):
        if getattr(self, "graph", None) is not None:
            self.graph.replay()
        else:
            self._eager_step()

    def units_per_step_all_ranks(self):
        return float(self.B * self.world)

    def setup_e2e(self):
        self.h_rep = _pinned(self.rep.detach().cpu())
        self.h_grad = torch.empty((self.B, self.E), dtype=torch.float32, pin_memory=True)
        self.h2d_bytes = self.B * self.E * 4
        self.d2h_bytes = self.B * self.E * 4 + 4

    def step_e2e(self):
        with torch.no_grad():
            self.rep.copy_(self.h_rep, non_blocking=True)
        if getattr(self, "graph", None) is None:
            _, self.ptrs = self.teacher.rows(self.dom, self.seq_ids)
        self.step()
        self.h_grad.copy_(self.rep.grad, non_blocking=True)
        float(self.loss)

    def roofline(self, peaks):
        from bench import event_time_ms
        from mergerec_b200.module.distiller.sequence.module import distill_logits
        rep = self.rep.detach()
        out = distill_logits(rep, self.tables, self.dom)
        ms = event_time_ms(lambda: distill_logits(rep, self.tables, self.dom, out=out), 20)
        ms_step = event_time_ms(self.step, 10)
        ach = self.bytes_tables / GB / (ms * 1e-3)
        return _with_traffic({"bound": "hbm", "kernel": "mr::ds_logits_kernel<6, 2> (catalogue logits, every item table read once)",
                "achieved": ach, "peak": peaks["hbm_gbs"], "peak_source": peaks["source"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "ms_per_launch": ms,
                "algorithmic_bytes_per_launch": self.bytes_tables, "step_ms": ms_step,
                "step_GB/s (tables read twice)": 2 * self.bytes_tables / GB / (ms_step * 1e-3)}, "ds_logits_kernel")

    def extra(self):
        return {"loss": None if self.loss is None else float(self.loss)}

    def _cpu_time(self, reps):
        from oracle import oracle as orc
        rng = np.random.Generator(np.random.PCG64(7))
        tables = [rng.standard_normal((self.N, self.E), dtype=np.float32) * np.float32(0.05) for _ in range(self.D)]
        rep = rng.standard_normal((self.B, self.E), dtype=np.float32)
        trows = [rng.standard_normal(self.N, dtype=np.float32) for _ in range(self.B)]
        ts = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            orc.distill_step(rep, tables, self.dom if hasattr(self, "dom") else [b % self.D for b in range(self.B)], trows, "KD", temperature=2.0)
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts[1:])), os.cpu_count() or 1

    def cpu_baseline(self):
        t, cores = self._cpu_time(2)
        return {"value": self.B / t, "unit": self.unit, "cores": cores, "kind": "port",
                "sample": "the full step (16 samples, 8 x 25,000 x 768 tables), numpy fp64 oracle port (BLAS threads)", "seconds_per_step": t}

    def reference_arm(self, steps, warmup):
        t, cores = self._cpu_time(max(1, min(steps, 3)))
        return {"value": self.B / t, "ms_per_step": t * 1e3, "cores": cores,
                "sample": "the full step (16 samples, 8 x 25,000 x 768 tables), numpy fp64 oracle port"}


WORKLOADS = {LambdaMergeK8.name: LambdaMergeK8, TiesCfg2.name: TiesCfg2, EvalCatalog.name: EvalCatalog,
             CollabStepCfg3.name: CollabStepCfg3, DistillStep.name: DistillStep, TiesSharded.name: TiesSharded}
DEFAULT_WORKLOAD = TiesCfg2.name
