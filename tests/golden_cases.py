"""Case definitions shared by tests/golden/make_golden.py (reference side, build container only)
and the parity tests (oracle / CUDA side).  Inputs are regenerated from these seeds."""

# d values: 4099 = odd (no 16-byte alignment of anything), 4165 = 32*130 + 5 (a 5-column tail for
# the interleaved torch.sum order when K >= 5), 8192 = aligned.
MERGE_FLAT_CASES = [
    dict(name="k3_d4099", K=3, d=4099, seed=11, weights=[0.3, 0.7, -0.123456789]),
    dict(name="k8_d4165", K=8, d=4165, seed=12, weights=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]),
    dict(name="k1_d8192", K=1, d=8192, seed=13, weights=[1.0]),
    dict(name="k16_d1000", K=16, d=1000, seed=14, weights=[0.05 * (i + 1) for i in range(16)]),
]

MODEL_MERGER_CASES = [
    dict(name="tiny_roberta_k3", recformer=False, K=3, seed=21,
         merges=[("task_vector", 0.3, {}), ("linear", [0.2, 0.3, 0.5], {}), ("task_vector", [0.5, -0.25, 1.0], {})]),
    dict(name="tiny_recformer_k5", recformer=True, K=5, seed=22,
         merges=[("task_vector", 0.2, {}), ("linear", 0.2, {})]),
]

LAMBDA_CASES = [
    dict(name="roberta_k3", recformer=False, K=3, seed=31),
    dict(name="roberta_k8", recformer=False, K=8, seed=32),
    dict(name="recformer_k8", recformer=True, K=8, seed=33),
    dict(name="recformer_k5", recformer=True, K=5, seed=34),
]

TIES_CASES = [
    dict(name="tiefree_k3", K=3, d=4165, seed=41, tie_free=True, density=0.2, weights=[0.3, 0.5, 0.7]),
    dict(name="tiefree_k8", K=8, d=4165, seed=42, tie_free=True, density=0.2,
         weights=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]),
    dict(name="tiefree_k5_dense", K=5, d=2051, seed=43, tie_free=True, density=0.7, weights=[1.0, 0.5, 0.25, 2.0, 1.5]),
    dict(name="gauss_k4", K=4, d=10007, seed=44, tie_free=False, density=0.2, weights=[0.3, 0.3, 0.3, 0.3]),
    dict(name="quant_k6", K=6, d=4165, seed=45, tie_free=False, quantize=2.5e-4, density=0.2,
         weights=[0.5, 0.5, 0.5, 0.5, 0.5, 0.5]),
]

LNS_CASES = [
    dict(name="tiefree_k3", K=3, d=4165, seed=71, tie_free=True, density=0.05, weights=[1.0, 1.0, 1.0]),
    dict(name="tiefree_k8", K=8, d=4165, seed=72, tie_free=True, density=0.2,
         weights=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]),
    dict(name="gauss_k5", K=5, d=10007, seed=73, tie_free=False, density=0.05, weights=[0.3, 0.5, 0.7, 0.9, 1.1]),
    dict(name="tiny_density", K=2, d=64, seed=74, tie_free=True, density=0.01, weights=[1.0, 1.0]),   # int(0.64) = 0
]

EVAL_CASES = [
    dict(name="gauss_small", kind="gauss", Q=64, N=500, E=32, seed=51, metrics=["NDCG", "RECALL"], ks=[1, 5, 10, 50], prefix=""),
    dict(name="grid_small", kind="grid", Q=64, N=500, E=16, seed=52, metrics=["NDCG", "RECALL"], ks=[1, 5, 10, 50], prefix="test/"),
    dict(name="gauss_cfg1", kind="gauss", Q=256, N=20000, E=768, seed=2, metrics=["NDCG", "RECALL"], ks=[10], prefix=""),
    dict(name="grid_cfg1", kind="grid", Q=256, N=20000, E=768, seed=2, metrics=["NDCG", "RECALL"], ks=[10], prefix=""),
    dict(name="grid_top100", kind="grid", Q=96, N=3000, E=64, seed=53, metrics=["RECALL", "NDCG"], ks=[1, 10, 100], prefix="val_"),
]

# BASELINE config 4's evaluator shape (200,000-item union catalog, E = 1024, top-50) on a slice of the queries.
EVAL_CFG4_CASES = [
    dict(name="grid_cfg4_q64", kind="grid", Q=64, N=200_000, E=1024, seed=54, metrics=["NDCG", "RECALL"], ks=[10, 50], prefix=""),
]

MODULE_CASES = [
    dict(name="tv_taskwise", merge_type="TASK_VECTOR", learn_type="TASK_WISE", K=3, seed=61, disable_softmax=True),
    dict(name="tv_layerwise", merge_type="TASK_VECTOR", learn_type="LAYER_WISE", K=3, seed=62, disable_softmax=True),
    dict(name="ties_layerwise_softmax", merge_type="TIES", learn_type="LAYER_WISE", K=5, seed=63, density=0.2,
         disable_softmax=False),
]

# Distillation step.  rows: items per domain (ragged, odd sizes, one domain with a
# single sample and one with more than four so that a table is walked by two sample groups).
DISTILL_LOSSES = [
    ("CE", {}), ("KD", dict(temperature=2.0)), ("MSE", {}), ("ADAMERGING", {}),
    ("ADAMERGING_KD", dict(temperature=0.5, coefficient=0.3)), ("MERGED_PSEUDO_LABEL", {}),
    ("MERGED_PSEUDO_LABEL_KD", dict(temperature=2.0, coefficient=0.7)), ("SINGLE_PSEUDO_LABEL", {}),
    ("SINGLE_PSEUDO_LABEL_KD", dict(temperature=1.0, coefficient=0.5)), ("PAIRWISE", dict(margin=0.2)),
    ("LISTNET", dict(temperature=0.1)),
]
DISTILL_CASES = [
    dict(name="b6_e64", B=6, E=64, rows=[37, 130, 257], n_seq=5, seed=81, scale=8.0),
    dict(name="b16_e768", B=16, E=768, rows=[301, 64, 999], n_seq=7, seed=82, scale=20.0),
    dict(name="b9_e1024", B=9, E=1024, rows=[513, 1], n_seq=4, seed=83, scale=5.0),
]

# PCB merging: d with a 5-column interleaved tail, an odd d, a larger d.
PCB_CASES = [
    dict(name="k3_d4165", K=3, d=4165, seed=91, density=0.2, weights=[0.3, 0.5, 0.7]),
    dict(name="k8_d4165", K=8, d=4165, seed=92, density=0.2, weights=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]),
    dict(name="k5_d10007", K=5, d=10007, seed=93, density=0.1, weights=[1.0, 0.5, 0.25, 2.0, 1.5]),
    dict(name="k2_d40000", K=2, d=40000, seed=94, density=0.5, weights=[0.6, 0.4]),
    dict(name="k3_d77", K=3, d=77, seed=95, density=0.2, weights=[0.3, 0.5, 0.7]),   # d < 100: int(0.01 d) = 0 -> clamp_min = row minimum
]

# DARE: the reference draws its dropout masks from torch's CPU generator; the golden file stores those masks (replayed
# from the same seed) next to the merged vector.
DARE_CASES = [
    dict(name="k3_d4099", K=3, d=4099, seed=101, density=0.2, weights=[0.3, 0.7, -0.123456789], torch_seed=7),
    dict(name="k8_d4165", K=8, d=4165, seed=102, density=0.9, weights=[0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], torch_seed=8),
    dict(name="k2_p0", K=2, d=1000, seed=103, density=0.0, weights=[0.5, 0.5], torch_seed=9),
    dict(name="k2_p1", K=2, d=1000, seed=104, density=1.0, weights=[0.5, 0.5], torch_seed=10),
]

# Stack B end to end: load_merging_module -> encoder -> distillation loss -> lambda-gradient -> Adam (3 steps).
COLLAB_DISTILL_CASES = [
    dict(name="layerwise_kd", K=3, seed=55, learn_type="LAYER_WISE", loss="KD", kw=dict(temperature=2.0), rows=[17, 40, 33]),
    dict(name="taskwise_ce", K=4, seed=56, learn_type="TASK_WISE", loss="CE", kw={}, rows=[25, 64]),
]
