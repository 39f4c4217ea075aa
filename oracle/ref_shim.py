"""CPU ORACLE tooling (not product code): run the UNMODIFIED reference Jacobi functions in this container.

The reference pins transformers 4.53.0 (requirements.txt:97); this image has 5.5.0.  Five monkey-patches, applied to
`transformers` (never to the reference files, which are imported as they lie under /root/reference), make
modeling/cllm2_qwen2_modeling_kv_terminate_on_eos_improved_multiblock_lookahead_unified.py and
modeling/cllm2_qwen2_modeling_kv_terminate_on_eos_improved.py run on CPU:

  1. Qwen2DecoderLayer.attention_type        attribute read at unified.py:207,455; removed in transformers 5
  2. Qwen2DecoderLayer.forward               accept legacy `past_key_value=` / `cache_position=`, return a 1-tuple
                                             when the legacy kwarg was used (reference indexes [0], :213,:461)
  3. DynamicCache.key_cache / value_cache    list-like views over cache.layers[i].keys/.values (:38-57, :100-126, :500-502)
  4. create_causal_mask (reference namespace) all-ones short mask -> None (what flash_attention_2 effectively does),
                                             `input_embeds` -> `inputs_embeds`
  5. DynamicCache.delete_false_key_value     re-bound to the multiblock module's guarded version (import-order hazard)

/root/reference does not exist on the GPU box: only oracle/make_golden.py (run here, output committed under
tests/golden/) and the `-m "not gpu"` pinning test that skips when the reference is absent use this module.
"""
from __future__ import annotations

import os
import sys

REFERENCE_ROOT = os.environ.get("JF_REFERENCE_ROOT", "/root/reference")
MB_MODULE = "modeling.cllm2_qwen2_modeling_kv_terminate_on_eos_improved_multiblock_lookahead_unified"
SB_MODULE = "modeling.cllm2_qwen2_modeling_kv_terminate_on_eos_improved"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, *MB_MODULE.split(".")) + ".py")


_installed = {}


def install():
    """Returns (multiblock_module, single_block_module) with the shim applied.  Idempotent."""
    if _installed:
        return _installed["mb"], _installed["sb"]
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    import importlib
    from transformers.cache_utils import DynamicCache
    from transformers.models.qwen2 import modeling_qwen2 as mq
    from transformers.masking_utils import create_causal_mask as _ccm

    class _KVList:                                              # (3)
        def __init__(self, cache, attr):
            self.c, self.a = cache, attr

        def __len__(self):
            return sum(1 for l in self.c.layers if getattr(l, "is_initialized", False))

        def __bool__(self):
            return len(self) > 0

        def __getitem__(self, i):
            return getattr(self.c.layers[i], self.a)

        def __setitem__(self, i, v):
            setattr(self.c.layers[i], self.a, v)

    DynamicCache.key_cache = property(lambda self: _KVList(self, "keys"))
    DynamicCache.value_cache = property(lambda self: _KVList(self, "values"))

    if not getattr(mq.Qwen2DecoderLayer, "_jf_shimmed", False):  # (2)
        _orig = mq.Qwen2DecoderLayer.forward

        def _fwd(self, hidden_states, *a, past_key_value=None, cache_position=None, **kw):
            legacy = past_key_value is not None
            if legacy and kw.get("past_key_values") is None:
                kw["past_key_values"] = past_key_value
            out = _orig(self, hidden_states, *a, **kw)
            return (out,) if legacy else out

        mq.Qwen2DecoderLayer.forward = _fwd
        mq.Qwen2DecoderLayer._jf_shimmed = True
    mq.Qwen2DecoderLayer.attention_type = "full_attention"      # (1)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sb = importlib.import_module(SB_MODULE)
    mb = importlib.import_module(MB_MODULE)                      # imported LAST (finding 7)

    def _mask(**kw):                                             # (4)
        kw["attention_mask"] = None
        if "input_embeds" in kw:
            kw["inputs_embeds"] = kw.pop("input_embeds")
        return _ccm(**kw)

    mb.create_causal_mask = _mask
    sb.create_causal_mask = _mask
    DynamicCache.delete_false_key_value = mb._delete_false_key_value   # (5)
    _installed["mb"], _installed["sb"] = mb, sb
    return mb, sb


def bind(model):
    """Attach the reference functions to an HF Qwen2ForCausalLM instance the way the reference drivers do
    (JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:24-27: `Qwen2ForCausalLM.jacobi_forward_greedy_multiblock = ...`)."""
    import types
    mb, sb = install()
    model.jacobi_forward_greedy_multiblock = types.MethodType(mb.jacobi_forward_greedy_multiblock, model)
    model.jacobi_forward_greedy = types.MethodType(sb.jacobi_forward_greedy, model)
    return model
