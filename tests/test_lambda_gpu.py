"""GPU parity for the collaborative-merging module (A3-A5, A11): lambda-merge forward bit-exact, lambda-gradient
within 1e-5 relative of an fp64 dot product, and 3 Adam steps through load_merging_module against the reference's
trajectory."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from helpers import assert_bit_equal, flatten_np, golden, shape_dict_of, state_dict_case
from mergerec_b200 import synth
from mergerec_b200.merger.algorithms import get_task_vectors
from mergerec_b200.merger.enums import LearnType, MergeType
from mergerec_b200.merger.layout import FlatLayout
from mergerec_b200.merger.weight_learning import (TaskVectorMergingModuleLayerWise, TaskVectorMergingModuleTaskWise,
                                                  load_merging_module)
from mergerec_b200.merger.weight_learning.module._base import _lambda_grad
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

LAMBDA_GRAD_RTOL = 1e-5  # north star: lambda-gradient within 1e-5 relative of the fp64 dot product


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


class _NoModel(torch.nn.Module):
    def forward(self, x):
        return x


@pytest.mark.parametrize("case", gc.LAMBDA_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("learn", ["task", "layer"])
@pytest.mark.parametrize("softmax", [0, 1])
def test_module_merge_and_grad_vs_golden(case, learn, softmax):
    g = golden("lambda_merge")
    _, base, models = state_dict_case(case)
    shape_dict = {k: torch.Size(v) for k, v in shape_dict_of(base).items()}
    tb = dev(flatten_np(base))
    T = get_task_vectors(tb, [dev(flatten_np(m)) for m in models])
    cls = TaskVectorMergingModuleTaskWise if learn == "task" else TaskVectorMergingModuleLayerWise
    mod = cls(tb, T, _NoModel(), shape_dict, disable_softmax=not softmax).cuda()
    tag = f"{case['name']}/{learn}/softmax{softmax}"
    keys = list(mod.per_weights.keys())
    assert keys == list(g[f"{tag}/keys"])
    with torch.no_grad():
        for i, key in enumerate(keys):
            mod.per_weights[key].copy_(dev(g[f"{tag}/per_weights"][i]))
            mod.global_weights[key].copy_(dev(g[f"{tag}/global_weights"][i]))
            mod.global_biases[key].copy_(dev(g[f"{tag}/global_biases"][i]))
    merged = mod._merge_task_vectors()
    if not softmax:
        # w = gw * pw + gb is two correctly rounded fp32 ops on either device -> identical lambdas -> identical bits
        assert_bit_equal(host(merged), g[f"{tag}/merged"], "A3/A4 merged (module)")
    else:
        # softmax runs through the device's expf: lambdas differ from the CPU's by an ulp -> ~|T| * 6e-8 absolute
        np.testing.assert_allclose(host(merged), g[f"{tag}/merged"], rtol=1e-6, atol=2e-8)
    grad_out = dev(g[f"{tag}/grad_out"])
    (merged * grad_out).sum().backward()
    for name, store in (("grad_per_weights", mod.per_weights), ("grad_global_weights", mod.global_weights),
                        ("grad_global_biases", mod.global_biases)):
        got = np.stack([host(store[k].grad) for k in keys])
        np.testing.assert_allclose(got, g[f"{tag}/{name}"], rtol=2e-4, atol=2e-6, err_msg=name)


@pytest.mark.parametrize("recformer", [False, True])
@pytest.mark.parametrize("K", [1, 3, 8, 16])
def test_lambda_grad_kernel_vs_fp64(recformer, K):
    """Pointer-table gradients (separate allocations, some missing) against the fp64 oracle."""
    shapes = synth.tiny_shapes(recformer=recformer, hidden=40, ffn=72, vocab=211)
    layout = FlatLayout.from_shape_dict(shapes)
    rng = np.random.Generator(np.random.PCG64(K))
    T = rng.standard_normal((K, layout.d), dtype=np.float32)
    grads_h = [rng.standard_normal(n, dtype=np.float32) for n in layout.sizes]
    missing = {1, len(grads_h) - 2}
    flat = np.concatenate([np.zeros_like(gh) if i in missing else gh for i, gh in enumerate(grads_h)])
    grads = [None if i in missing else dev(gh) for i, gh in enumerate(grads_h)]
    for layer_wise in (False, True):
        sb, se, sg, keys = orc.segment_table(shapes, layer_wise=layer_wise)
        _, seg_group, dkeys = layout.device_blocks(layer_wise, "cuda")
        assert dkeys == keys
        # segments are always the tensors; task-wise maps them all to group 0
        grp = seg_group if layer_wise else None
        Tt = dev(T)
        got = host(_lambda_grad(grads, layout, Tt, grp, len(keys))).astype(np.float64)
        tsb, tse, tsg, _ = orc.segment_table(shapes, layer_wise=True)
        want = orc.lambda_grad(flat, T, len(keys), tsb, tse, tsg if layer_wise else np.zeros_like(tsg))
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= LAMBDA_GRAD_RTOL * scale, (np.abs(got - want).max(), scale)


def test_lambda_grad_is_deterministic():
    shapes = synth.tiny_shapes(hidden=64, ffn=128, vocab=1000)
    layout = FlatLayout.from_shape_dict(shapes)
    g = torch.Generator(device="cuda").manual_seed(0)
    T = torch.randn(5, layout.d, generator=g, device="cuda")
    grad = torch.randn(layout.d, generator=g, device="cuda")
    grads = [grad[o:o + n] for o, n in zip(layout.offsets, layout.sizes)]
    _, seg_group, keys = layout.device_blocks(True, "cuda")
    a = _lambda_grad(grads, layout, T, seg_group, len(keys))
    b = _lambda_grad(grads, layout, T, seg_group, len(keys))
    assert torch.equal(a, b)


@pytest.mark.parametrize("case", gc.MODULE_CASES, ids=lambda c: c["name"])
def test_load_merging_module_adam_trajectory(case):
    """Stack B of SURVEY.md section 3 on a toy encoder: merged0 bit-exact (no softmax) and the lambda trajectory of
    3 Adam steps equal to the reference's autograd run within fp32 noise."""
    from toy_model import ToyEncoder, make_toy_state_dicts

    g = golden("module_e2e")
    pre, fts = make_toy_state_dicts(case["K"], seed=case["seed"])
    torch.manual_seed(1234)
    model = ToyEncoder()
    mod = load_merging_module(MergeType[case["merge_type"]], LearnType[case["learn_type"]], model, pre, fts,
                              ignore_keys=set(), ties_density=case.get("density"), initial_per_weight=0.3,
                              disable_softmax=case["disable_softmax"])
    assert next(mod.per_weights.parameters()).is_cuda
    keys = list(mod.per_weights.keys())
    assert keys == list(g[f"{case['name']}/keys"])
    merged0 = host(mod._merge_task_vectors())
    if case["disable_softmax"]:
        assert_bit_equal(merged0, g[f"{case['name']}/merged0"], "merged0")
    else:
        np.testing.assert_allclose(merged0, g[f"{case['name']}/merged0"], rtol=1e-6, atol=2e-8)
    rng = np.random.Generator(np.random.PCG64(case["seed"] + 7))
    ids = dev(rng.integers(0, 37, size=(3, 6, 5)).astype(np.int64))
    tgt = dev(rng.standard_normal((3, 6, 24), dtype=np.float32))
    opt = torch.optim.Adam(mod.trainable_parameters(True, True, False), lr=1e-2)
    traj, losses = [], []
    for step in range(3):
        opt.zero_grad()
        rep = mod(ids[step])
        loss = ((rep - tgt[step]) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
        traj.append(np.stack([host(mod.per_weights[k]) for k in keys]))
    np.testing.assert_allclose(np.asarray(losses), g[f"{case['name']}/losses"], rtol=1e-4)
    # Adam normalises the gradient, so tiny gradient noise moves lambda by at most ~lr * relative noise
    np.testing.assert_allclose(np.stack(traj), g[f"{case['name']}/per_weights_traj"], rtol=0, atol=2e-4)
    sd = mod.get_state_dict()
    assert list(sd.keys()) == list(g[f"{case['name']}/sd_keys"])
    flat = torch.cat([v.reshape(-1) for v in sd.values()])
    np.testing.assert_allclose(host(flat), g[f"{case['name']}/final_flat"], rtol=0, atol=2e-5)
    ser = mod.serialize_weights()
    assert set(ser) == {"global_weights", "global_biases", "per_weights"} and list(ser["per_weights"]) == keys
    mod.load_weights_from_dict(ser)


def test_lambda_grad_full_size_blair_base():
    """K = 8, d = 124,645,632 (BASELINE config 3 size): kernel vs an fp64 reduction done on the GPU in chunks."""
    shapes = synth.roberta_shapes()
    layout = FlatLayout.from_shape_dict(shapes)
    K, d = 8, layout.d
    g = torch.Generator(device="cuda").manual_seed(3)
    T = torch.randn(K, d, generator=g, device="cuda") * 1e-3
    grad = torch.randn(d, generator=g, device="cuda")
    grads = [grad[o:o + n] for o, n in zip(layout.offsets, layout.sizes)]
    _, seg_group, keys = layout.device_blocks(True, "cuda")
    got = _lambda_grad(grads, layout, T, seg_group, len(keys)).double()
    want = torch.zeros(len(keys), K, dtype=torch.float64, device="cuda")
    grp = seg_group.tolist()
    for p, (o, n) in enumerate(zip(layout.offsets, layout.sizes)):
        want[grp[p]] += T[:, o:o + n].double() @ grad[o:o + n].double()
    scale = want.abs().max()
    assert (got - want).abs().max() <= LAMBDA_GRAD_RTOL * scale


def test_merge_test_flow_from_files(tmp_path):
    """Stack A of SURVEY.md section 3 from the reference's on-disk formats: extracted state_dict.pt files (one
    "model." prefix too many, item_embeddings inside) -> load_merging_module -> weight file ("average" and a jsonl
    log line) -> get_state_dict, against the oracle's task-arithmetic merge of the same tensors."""
    from toy_model import ToyEncoder, make_toy_state_dicts
    from mergerec_b200 import io as mio
    from mergerec_b200.merger.enums import LearnType, MergeType
    from mergerec_b200.merger.weight_learning.module import load_merging_module
    K = 3
    pre, fts = make_toy_state_dicts(K, seed=91)
    paths = []
    for k, ft in enumerate(fts):
        sd = {"model." + n: v for n, v in ft.items()}
        sd["item_embeddings"] = torch.randn(5, 24)
        p = tmp_path / f"domain{k}" / "state_dict.pt"
        p.parent.mkdir()
        torch.save(sd, p)
        paths.append(p)
    loaded = mio.load_finetuned_state_dicts(paths)
    assert all(list(sd.keys()) == list(pre.keys()) for sd in loaded)
    torch.manual_seed(5)
    model = ToyEncoder()
    model.load_state_dict(pre)
    mod = load_merging_module(MergeType.TASK_VECTOR, LearnType.TASK_WISE, model, model.state_dict(), loaded,
                              ignore_keys=set(), disable_softmax=True)
    keys = list(pre.keys())
    fb = np.concatenate([pre[k].numpy().reshape(-1).astype(np.float32) for k in keys])
    fm = [np.concatenate([ft[k].numpy().reshape(-1).astype(np.float32) for k in keys]) for ft in fts]
    T = orc.task_vectors(fb, fm)

    def merged_flat():
        sd = mod.get_state_dict()
        return np.concatenate([sd[k].detach().cpu().numpy().reshape(-1) for k in keys])

    mod.load_weights_from_dict(mio.resolve_weights("average", 0, K))
    w = np.full((1, K), np.float32(1.0) * np.float32(1.0 / K) + np.float32(0.0), np.float32)
    assert_bit_equal(merged_flat(), orc.lambda_merge(fb, T, w), "average weights")
    with mio.WeightLog("v1", tmp_path / "weights", log_every_steps=1) as log:
        mod.load_weights_from_dict({"global_weights": {"all": [1.0]}, "global_biases": {"all": [0.0]},
                                    "per_weights": {"all": [0.7, 0.2, 0.4]}})
        log.on_train_batch_end(0, 0, 0, mod)
        mod.load_weights_from_dict(mio.resolve_weights("uniform", 0.3, K))
        log.on_train_batch_end(0, 1, 1, mod)
    mod.load_weights_from_dict(mio.resolve_weights(tmp_path / "weights" / "v1.jsonl", 0, K))
    w = np.asarray([[0.7, 0.2, 0.4]], np.float32)
    assert_bit_equal(merged_flat(), orc.lambda_merge(fb, T, w), "line 0 of the lambda log")


def test_lambda_merge_and_gradient_capture_in_a_cuda_graph():
    """The merge forward and its lambda-gradient backward are allocation-free and stream-ordered once warm (cached pointer
    table and workspace): both capture into ONE CUDA graph, and a replay with new lambdas / new upstream gradients gives the
    eager result bit for bit."""
    shapes = synth.tiny_shapes(hidden=64, ffn=128, vocab=1000)
    layout = FlatLayout.from_shape_dict(shapes)
    K = 5
    g = torch.Generator(device="cuda").manual_seed(1)
    T = torch.randn(K, layout.d, generator=g, device="cuda")
    base = torch.randn(layout.d, generator=g, device="cuda")
    seg_end, seg_group, keys = layout.device_blocks(True, "cuda")
    G = len(keys)
    from mergerec_b200 import _lib
    from mergerec_b200.merger.algorithms._common import merge_axpy
    w = torch.rand((G, K), generator=g, device="cuda")
    grad = torch.randn(layout.d, generator=g, device="cuda")
    grads = [grad[o:o + n] for o, n in zip(layout.offsets, layout.sizes)]       # fixed addresses: views of one buffer
    merged = torch.empty(layout.d, device="cuda")
    rows = list(T.unbind(0))

    def step():
        merge_axpy(base, rows, w, _lib.MR_ORDER_SUM_FIRST, False, seg_end, seg_group, out=merged)
        return _lambda_grad(grads, layout, T, seg_group, G)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = step()
    for trial in range(3):
        w.copy_(torch.rand((G, K), generator=g, device="cuda"))
        grad.copy_(torch.randn(layout.d, generator=g, device="cuda"))
        graph.replay()
        got_merged, got_gw = merged.clone(), out.clone()
        want_gw = step()
        assert torch.equal(got_merged, merged) and torch.equal(got_gw, want_gw), trial
