"""On-disk formats of MergeRec: host-side glue so the merger / evaluator can consume the
reference's checkpoints and lambda logs unchanged.  No arithmetic happens here.

reference: scripts/2_ft_postprocess/extract.py:7-20   Lightning ckpt -> state_dict.pt + item_embedding.pt
           utils.py:17-29                             remove_duplicate_prefix (strip ONE leading "model.")
           merge_test.py:20-25                        torch.load, pop "item_embeddings", strip the prefix
           merge_test.py:47-69                        weight file modes: "average", "uniform", a line of a jsonl log
           rec_retrieval/module/callbacks.py:139-174  SaveWeightsCallback: one python-dict repr per line
"""
from __future__ import annotations

import ast
import re
from pathlib import Path
from typing import Any, Dict, Iterable, List, Union

import torch

from .merger.types import PathStr, StateDict

WeightsDict = Dict[str, Dict[str, List[float]]]


def extract_checkpoint(model_checkpoint: PathStr, output_dir: PathStr) -> None:
    """Lightning checkpoint -> ``state_dict.pt`` (still holding ``item_embeddings``) and ``item_embedding.pt``."""
    model_checkpoint, output_dir = Path(model_checkpoint), Path(output_dir)
    if not model_checkpoint.exists():
        raise FileNotFoundError(f"Model checkpoint not found: {model_checkpoint}")
    output_dir.mkdir(parents=True, exist_ok=True)
    state_dict = torch.load(model_checkpoint, map_location="cpu")["state_dict"]
    torch.save(state_dict["item_embeddings"], output_dir / "item_embedding.pt")
    torch.save(state_dict, output_dir / "state_dict.pt")


def remove_duplicate_prefix(state_dict: StateDict) -> StateDict:
    """Strip one leading ``"model."`` from every key that has it; other keys are kept as they are."""
    return {(k.replace("model.", "", 1) if k.startswith("model.") else k): v for k, v in state_dict.items()}


def load_finetuned_state_dict(path: PathStr, drop_item_embeddings: bool = True) -> StateDict:
    """One fine-tuned domain model as ``merge_test.py`` prepares it for ``load_merging_module``.

    ``merge_train.py`` does not pop ``item_embeddings`` (the key intersection in ``load_merging_module`` drops it,
    _factory.py:55); pass ``drop_item_embeddings=False`` to mirror that."""
    state_dict = torch.load(Path(path), map_location="cpu")
    if drop_item_embeddings:
        state_dict.pop("item_embeddings")          # KeyError when absent, like the reference
    return remove_duplicate_prefix(state_dict)


def load_item_embeddings(path: PathStr) -> torch.Tensor:
    """``item_embedding.pt`` written by ``extract_checkpoint`` (the evaluator's item table / the teacher's inputs)."""
    return torch.load(Path(path), map_location="cpu")


def read_weight_log(path: PathStr) -> List[Dict[str, Any]]:
    """Lines of a ``SaveWeightsCallback`` log: python reprs of ``{"epoch", "step", "weights"}``.  The reference reads
    them back with ``eval`` (merge_test.py:67); ``ast.literal_eval`` accepts exactly the literals that callback writes
    (dicts, lists, strings, ints, floats incl. ``inf``/``nan`` spelled by repr) without executing anything."""
    text = Path(path).read_text().strip()
    return [_literal(line) for line in text.splitlines()] if text else []


_NAN_TOKEN = "__mergerec_nan__"
# a bare inf / nan NAME (what repr() writes for non-finite floats), i.e. not part of a quoted key or another word
_BARE_NONFINITE = re.compile(r"(?<![\w'\"])(-?)(inf|nan)(?![\w'\"])")


def _restore_nan(obj):
    if isinstance(obj, str) and obj == _NAN_TOKEN:
        return float("nan")
    if isinstance(obj, dict):
        return {k: _restore_nan(v) for k, v in obj.items()}
    if isinstance(obj, list):
        return [_restore_nan(v) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_restore_nan(v) for v in obj)
    return obj


def _literal(line: str):
    try:
        return ast.literal_eval(line)
    except ValueError:
        # repr() writes non-finite floats as the bare names inf / nan, which literal_eval refuses (the reference's eval
        # would fail on them too unless the names are defined).  Only stand-alone tokens are rewritten, so group keys
        # that contain "inf" / "nan" stay intact; nan comes back as float("nan"), inf as float("inf").
        def sub(m):
            return f"{m.group(1)}1e999" if m.group(2) == "inf" else f"'{_NAN_TOKEN}'"
        return _restore_nan(ast.literal_eval(_BARE_NONFINITE.sub(sub, line)))


def resolve_weights(weight_file: PathStr, weight_file_line: Union[int, float], num_models: int) -> WeightsDict:
    """The lambda set ``merge_test.py`` loads into the merging module (merge_test.py:47-69):
    file NAME ``"average"`` -> 1/K each, ``"uniform"`` -> ``weight_file_line`` each (it is a float there),
    anything else -> the ``weights`` entry of line ``weight_file_line`` of that jsonl log."""
    weight_file = Path(weight_file)
    if weight_file.name == "average":
        per = [1.0 / num_models] * num_models
    elif weight_file.name == "uniform":
        per = [weight_file_line] * num_models
    else:
        return read_weight_log(weight_file)[int(weight_file_line)]["weights"]
    return {"global_weights": {"all": [1.0]}, "global_biases": {"all": [0.0]}, "per_weights": {"all": per}}


class WeightLog:
    """Writer of the reference's lambda log (SaveWeightsCallback, callbacks.py:139-174): ``{version}.jsonl`` in
    ``save_dir``, one ``repr`` of ``{"epoch", "step", "weights": module.serialize_weights()}`` per logged step."""

    def __init__(self, version: str, save_dir: PathStr = "weights", log_every_steps: int = 5):
        self.save_dir = Path(save_dir)
        self.save_dir.mkdir(parents=True, exist_ok=True)
        self.save_file = self.save_dir / f"{version}.jsonl"
        self.log_every_steps = log_every_steps
        self._fh = open(self.save_file, "w", encoding="utf-8")

    def on_train_batch_end(self, epoch: int, global_step: int, batch_idx: int, merged_model) -> None:
        if batch_idx % self.log_every_steps == 0:
            line = {"epoch": epoch, "step": global_step, "weights": merged_model.serialize_weights()}
            self._fh.write(f"{line}\n")

    def flush(self) -> None:
        self._fh.flush()

    def close(self) -> None:
        if self._fh:
            self._fh.close()
            self._fh = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def load_finetuned_state_dicts(paths: Iterable[PathStr], drop_item_embeddings: bool = True) -> List[StateDict]:
    return [load_finetuned_state_dict(p, drop_item_embeddings) for p in paths]
