"""CPU side of the checkpoint-directory path: the synthetic checkpoint is a directory AutoConfig / AutoTokenizer /
safetensors read offline, Config resolves a training output directory to its latest checkpoint-N (reference:
inference_engine/config.py:31-52) and the tensor names on disk are the ones jf_load_tensor accepts."""
import glob
import os


def test_synthetic_checkpoint_dir_is_loadable_offline(tmp_path):
    from safetensors import safe_open
    from transformers import AutoConfig, AutoTokenizer
    from jacobiforcing_b200.config import Config
    from tests.ckpt_helpers import make_checkpoint_dir
    make_checkpoint_dir(str(tmp_path / "checkpoint-3"), seed=1)
    model, tok = make_checkpoint_dir(str(tmp_path / "checkpoint-12"), seed=0, shards=2)
    cfg = Config(str(tmp_path), max_model_len=1024, max_num_batched_tokens=2048)
    assert cfg.model.endswith("checkpoint-12")
    assert cfg.hf_config.hidden_size == model.config.hidden_size
    assert AutoConfig.from_pretrained(cfg.model).num_hidden_layers == model.config.num_hidden_layers
    t2 = AutoTokenizer.from_pretrained(cfg.model, use_fast=True)
    assert t2.eos_token_id == tok.eos_token_id == model.config.vocab_size - 2
    assert t2.encode("w5 w17 w3") == [5, 17, 3]
    names = set()
    for f in sorted(glob.glob(os.path.join(cfg.model, "*.safetensors"))):
        with safe_open(f, "pt", "cpu") as sf:
            names |= set(sf.keys())
    L = model.config.num_hidden_layers
    for l in range(L):
        for leaf in ("self_attn.q_proj.weight", "self_attn.k_proj.bias", "self_attn.o_proj.weight", "mlp.gate_proj.weight",
                     "mlp.down_proj.weight", "input_layernorm.weight", "post_attention_layernorm.weight"):
            assert f"model.layers.{l}.{leaf}" in names
    assert {"model.embed_tokens.weight", "model.norm.weight", "lm_head.weight"} <= names
