"""mergerec_b200 -- B200-native (sm_100a) implementation of MergeRec's merger and evaluator hot paths.

Drop-in for ``rec_retrieval.merger`` / ``rec_retrieval.evaluator`` of DIALLab-SKKU/MergeRec: same class and
function names, argument meaning and error behaviour, with the arithmetic running in hand-written CUDA
kernels reached through the C ABI in ``include/mergerec_b200.h``.  No CPU fallback.
"""
__version__ = "0.1.0"
