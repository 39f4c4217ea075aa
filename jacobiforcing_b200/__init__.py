"""jacobiforcing_b200 — B200-native greedy Jacobi decoding for Qwen2 models.

Host side (Python) of the hot path: a mirror of the reference's `inference_engine` operator surface
(LLM / SamplingParams / ModelRunner.run) over a thin C-ABI CUDA library (csrc/ -> libjacobi_b200.so, see
include/jacobi_b200.h).  There is no CPU fallback: importing the runtime without the built library raises.
"""
__version__ = "0.1.0"

from .sampling_params import SamplingParams   # noqa: E402,F401


def __getattr__(name):                       # lazy: importing the package must not need the built library
    if name == "LLM":
        from .llm import LLM
        return LLM
    if name == "Config":
        from .config import Config
        return Config
    raise AttributeError(name)
