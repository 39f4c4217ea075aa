"""Synthetic checkpoint directory (there are no model files offline): what `LLM("<dir>")` needs from disk - config.json,
*.safetensors with the HF tensor names the reference's loader walks (inference_engine/utils/loader.py:12-47) and the
files AutoTokenizer reads - written the way SURVEY.md section 8(c) probed: Qwen2ForCausalLM.save_pretrained(safe
serialization) + a PreTrainedTokenizerFast over a WordLevel vocabulary."""
from __future__ import annotations

import os


def make_checkpoint_dir(path: str, name: str = "tiny", seed: int = 0, shards: int = 2, dtype=None):
    """Returns (hf_model_fp32_with_bf16_rounded_weights, tokenizer).  The weights on disk are bf16, split into `shards`
    safetensors files (the loader must walk all of them)."""
    import torch
    from safetensors.torch import save_file
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    from oracle.hf_model import build_model, state_dict_bf16
    os.makedirs(path, exist_ok=True)
    model = build_model(name, seed=seed, dtype=torch.float32)
    sd = state_dict_bf16(model)
    with torch.no_grad():
        for k, v in model.state_dict().items():
            v.copy_(v.to(torch.bfloat16).to(torch.float32))
    model.config.save_pretrained(path)
    names = sorted(sd)
    per = (len(names) + shards - 1) // shards
    for i in range(shards):
        part = {k: sd[k].contiguous() for k in names[i * per:(i + 1) * per]}
        if part:
            save_file(part, os.path.join(path, f"model-{i + 1:05d}-of-{shards:05d}.safetensors"))
    V = model.config.vocab_size
    vocab = {f"w{i}": i for i in range(V - 2)}
    vocab["<eos>"] = V - 2
    vocab["<pad>"] = V - 1
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="w0"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, eos_token="<eos>", pad_token="<pad>", unk_token="w0")
    fast.save_pretrained(path)
    return model, fast
