"""DARE merge (reference: rec_retrieval/merger/algorithms/dare.py).  The reference's dropout masks come from torch's CPU
generator; tests/golden/dare.npz stores the masks (replayed from the same seed) and the merged vectors, so both the
oracle (CPU) and the CUDA kernel (GPU) are checked bit for bit against the unmodified reference."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import assert_bit_equal, golden
from mergerec_b200 import synth
from oracle import oracle as orc


def _case(case):
    g = golden("dare")
    base, models = synth.make_flat(case["d"], case["K"], seed=case["seed"])
    masks = np.unpackbits(g[f"{case['name']}/masks"], axis=1)[:, :case["d"]].astype(bool)
    return base, models, masks, g[f"{case['name']}/merged"]


@pytest.mark.parametrize("case", gc.DARE_CASES, ids=lambda c: c["name"])
def test_oracle_matches_reference(case):
    base, models, masks, ref = _case(case)
    assert_bit_equal(orc.merge_dare(base, models, case["weights"], case["density"], masks), ref, case["name"])
    if 0.0 < case["density"] < 1.0:
        assert abs(masks.mean() - (1.0 - case["density"])) < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize("case", gc.DARE_CASES, ids=lambda c: c["name"])
def test_cuda_matches_reference_given_the_masks(case):
    from mergerec_b200.merger.algorithms import merge_dare
    base, models, masks, ref = _case(case)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    got = merge_dare(dev(base), [dev(m) for m in models], case["weights"], case["density"], masks=dev(masks))
    assert_bit_equal(got.cpu().numpy(), ref, case["name"])


@pytest.mark.gpu
def test_model_merger_dare_draws_reproducible_masks():
    from mergerec_b200.merger import ModelMerger
    shapes = synth.tiny_shapes(recformer=False)
    base, models = synth.make_state_dicts(shapes, 3, seed=21, sigma=1e-2)
    to_t = lambda sd: {k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}  # noqa: E731
    merger = ModelMerger([to_t(m) for m in models], to_t(base))
    g1 = torch.Generator(device="cuda").manual_seed(5)
    g2 = torch.Generator(device="cuda").manual_seed(5)
    a = merger.merge("dare", 0.5, density=0.3, generator=g1)
    b = merger.merge("dare", 0.5, density=0.3, generator=g2)
    assert all(torch.equal(a[k], b[k]) for k in a)
    flat = torch.cat([v.reshape(-1).float() for v in a.values()])
    bflat = merger.base_model
    changed = (flat != bflat).float().mean().item()
    tau_nonzero = torch.stack([(m != bflat) for m in merger.models]).any(0).float().mean().item()
    assert 0.5 * tau_nonzero < changed <= tau_nonzero          # about 1 - 0.3^3 of the columns with a non-zero update move
    with pytest.raises(ValueError):
        merger.merge("dare", 0.5, density=1.5)
    with pytest.raises(ValueError):
        merger.merge("dare", 0.5, density=0.3, masks=torch.ones(2, 5, dtype=torch.bool))
    p0 = merger.merge("dare", [0.3, 0.3, 0.3], density=0.0)      # p = 0: plain task arithmetic
    tv = merger.merge("task_vector", [0.3, 0.3, 0.3])
    assert all(torch.equal(p0[k], tv[k]) for k in tv)
