"""Callers either side of the two hot paths: the distillation losses and the
collaborative-merging step's logits -> loss -> gradient, as CUDA kernels behind the reference's class names
(rec_retrieval/module/recommender/loss_fn.py, rec_retrieval/module/distiller/sequence/module.py)."""
