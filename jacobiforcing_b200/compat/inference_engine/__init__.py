"""Import alias for scripts written against the reference engine.

Put `<repo>/jacobiforcing_b200/compat` on `sys.path` (or `PYTHONPATH`) and

    from inference_engine import LLM, SamplingParams

resolves to this engine's `LLM` / `SamplingParams` - the surface the reference's own script tests use
(`inference_engine/tests/test_jacobi_decoding_greedy.py:22`: `llm.generate(prompts, sampling_params, use_tqdm=False)`,
`llm.scheduler.waiting / running`, `llm.model_runner.jacobi_decoder.stats`).  Nothing is re-implemented here: every name is
looked up in `jacobiforcing_b200` when it is first used, so importing the alias needs neither the built library nor a GPU;
constructing an `LLM` does (there is no CPU fallback)."""
import importlib
import sys

import jacobiforcing_b200 as _jf

SamplingParams = _jf.SamplingParams

# pure-Python submodules a reference script may import by their reference path
for _ref, _ours in (("sampling_params", "jacobiforcing_b200.sampling_params"), ("config", "jacobiforcing_b200.config")):
    sys.modules[f"{__name__}.{_ref}"] = importlib.import_module(_ours)


def __getattr__(name):
    if name == "llm":                                   # `import inference_engine.llm` / `from inference_engine.llm import LLM`
        mod = importlib.import_module("jacobiforcing_b200.llm")
        sys.modules[f"{__name__}.llm"] = mod
        return mod
    return getattr(_jf, name)                           # LLM, Config (lazy in jacobiforcing_b200 too)
