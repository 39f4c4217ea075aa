"""jacobiforcing_b200 — B200-native greedy Jacobi decoding for Qwen2 models.

Host side (Python) of the hot path: a mirror of the reference's `inference_engine` operator surface
(LLM / SamplingParams / ModelRunner.run) over a thin C-ABI CUDA library (csrc/ -> libjacobi_b200.so, see
include/jacobi_b200.h).  There is no CPU fallback: importing the runtime without the built library raises.
"""
__version__ = "0.1.0"
