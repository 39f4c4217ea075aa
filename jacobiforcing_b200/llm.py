"""LLM — the user-facing class (inference_engine/llm.py).  `LLM(model, **Config fields).generate(prompts,
SamplingParams(...))` returns [{"text", "token_ids"}].  Jacobi knobs travel in SamplingParams: the reference's
`generate(..., jacobi_block_len=...)` keyword route is broken upstream (it replaces dataclass fields that do not
exist, inference_engine/llm.py:126-144 -> TypeError), so only the SamplingParams route is kept."""
from .engine.llm_engine import LLMEngine


class LLM(LLMEngine):
    pass
