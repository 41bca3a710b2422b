"""CPU: the reference's on-disk formats -- checkpoint extraction, prefix stripping,
weight-file modes and the lambda log -- round-tripped through synthetic files."""
import math

import pytest
import torch

from mergerec_b200 import io as mio


def fake_lightning_ckpt(path, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {"model.model.embeddings.word_embeddings.weight": torch.randn(7, 4, generator=g),
          "model.model.encoder.layer.0.attention.weight": torch.randn(4, 4, generator=g),
          "model.pooler.dense.bias": torch.randn(4, generator=g),
          "temperature": torch.tensor(0.05),
          "item_embeddings": torch.randn(11, 4, generator=g)}
    torch.save({"state_dict": sd, "epoch": 3, "global_step": 42}, path)
    return sd


def test_extract_and_load_roundtrip(tmp_path):
    ckpt = tmp_path / "last.ckpt"
    sd = fake_lightning_ckpt(ckpt)
    out = tmp_path / "extracted" / "beauty"
    mio.extract_checkpoint(ckpt, out)                       # creates the directory like extract.py:11-13
    assert sorted(p.name for p in out.iterdir()) == ["item_embedding.pt", "state_dict.pt"]
    assert torch.equal(mio.load_item_embeddings(out / "item_embedding.pt"), sd["item_embeddings"])
    loaded = mio.load_finetuned_state_dict(out / "state_dict.pt")
    # item_embeddings popped, ONE "model." stripped, keys without the prefix kept (utils.py:17-29)
    assert list(loaded.keys()) == ["model.embeddings.word_embeddings.weight", "model.encoder.layer.0.attention.weight",
                                   "pooler.dense.bias", "temperature"]
    assert torch.equal(loaded["model.embeddings.word_embeddings.weight"], sd["model.model.embeddings.word_embeddings.weight"])
    kept = mio.load_finetuned_state_dict(out / "state_dict.pt", drop_item_embeddings=False)
    assert "item_embeddings" in kept
    with pytest.raises(FileNotFoundError):
        mio.extract_checkpoint(tmp_path / "missing.ckpt", out)
    torch.save({k: v for k, v in sd.items() if k != "item_embeddings"}, tmp_path / "no_items.pt")
    with pytest.raises(KeyError):
        mio.load_finetuned_state_dict(tmp_path / "no_items.pt")


class _FakeModule:
    def __init__(self, per):
        self.per = per

    def serialize_weights(self):
        return {"global_weights": {"all": [1.0]}, "global_biases": {"all": [0.0]}, "per_weights": {"all": list(self.per)}}


def test_weight_file_modes(tmp_path):
    assert mio.resolve_weights("whatever/average", 0, 4) == {
        "global_weights": {"all": [1.0]}, "global_biases": {"all": [0.0]}, "per_weights": {"all": [0.25] * 4}}
    assert mio.resolve_weights("uniform", 0.3, 3)["per_weights"]["all"] == [0.3, 0.3, 0.3]
    with mio.WeightLog("run1", tmp_path / "weights", log_every_steps=2) as log:
        for step in range(5):
            log.on_train_batch_end(epoch=0, global_step=step, batch_idx=step, merged_model=_FakeModule([0.1 * step, 1.0 / 3.0]))
        log.flush()
    lines = (tmp_path / "weights" / "run1.jsonl").read_text().splitlines()
    assert len(lines) == 3                                   # batch_idx 0, 2, 4
    assert lines[1] == repr({"epoch": 0, "step": 2, "weights": _FakeModule([0.2, 1.0 / 3.0]).serialize_weights()})
    assert [eval(l) for l in lines] == mio.read_weight_log(tmp_path / "weights" / "run1.jsonl")   # same as the reference's eval
    w = mio.resolve_weights(tmp_path / "weights" / "run1.jsonl", 2, 2)
    assert w["per_weights"]["all"] == [0.4, 1.0 / 3.0]       # exact float round trip through repr
    assert mio.read_weight_log(tmp_path / "weights" / "run1.jsonl")[-1]["step"] == 4


def test_weight_log_non_finite(tmp_path):
    p = tmp_path / "w.jsonl"
    p.write_text(repr({"epoch": 0, "step": 0, "weights": {"per_weights": {"all": [float("inf"), 1.0]}}}) + "\n")
    assert math.isinf(mio.read_weight_log(p)[0]["weights"]["per_weights"]["all"][0])


def test_weight_log_nan_and_keys_containing_inf(tmp_path):
    """A NaN lambda comes back as float('nan') (not None); group keys that merely contain "inf" / "nan" are untouched."""
    p = tmp_path / "w.jsonl"
    weights = {"per_weights": {"info_layer": [float("nan"), float("-inf"), 0.25], "nanny": [1.0]},
               "global_weights": {"info_layer": [float("inf")]}}
    p.write_text(repr({"epoch": 1, "step": 7, "weights": weights}) + "\n")
    got = mio.read_weight_log(p)[0]
    pw = got["weights"]["per_weights"]
    assert set(pw.keys()) == {"info_layer", "nanny"} and pw["nanny"] == [1.0]
    assert math.isnan(pw["info_layer"][0]) and pw["info_layer"][1] == float("-inf") and pw["info_layer"][2] == 0.25
    assert got["weights"]["global_weights"]["info_layer"] == [float("inf")] and got["step"] == 7


def test_weight_checkpoint_callback_keeps_the_best_lambda_set():
    """callbacks.py:177-205: mean of the metrics matching the monitor regex, lower is better, restore at the end."""
    import torch
    from mergerec_b200.module.callbacks import WeightCheckpointCallback

    class _Merged:
        def __init__(self):
            self.w = [0.0]

        def serialize_weights(self):
            return {"per_weights": {"all": list(self.w)}}

        def load_weights_from_dict(self, d):
            self.w = list(d["per_weights"]["all"])

    class _PL:
        merged_model = _Merged()

    class _Trainer:
        callback_metrics = {}

    cb, pl, tr = WeightCheckpointCallback(monitor=r"val/loss.*"), _PL(), _Trainer()
    for step, (a, b) in enumerate([(3.0, 1.0), (1.0, 0.5), (2.0, 2.0)]):
        pl.merged_model.w = [float(step)]
        tr.callback_metrics = {"val/loss/dataloader_idx_0": torch.tensor(a), "val/loss/dataloader_idx_1": torch.tensor(b), "train/loss": torch.tensor(9.0)}
        cb.on_validation_epoch_end(tr, pl)
    assert cb.best_score == 0.75 and cb.best_weights == {"per_weights": {"all": [1.0]}}
    cb.load_weights(pl)
    assert pl.merged_model.w == [1.0]
    tr.callback_metrics = {"train/loss": torch.tensor(1.0)}
    import pytest
    with pytest.raises(RuntimeError):
        cb.on_validation_epoch_end(tr, pl)
