"""Tensor-parallel sharding of HF Qwen2 tensors — the narrowing the reference does in its weight loaders
(inference_engine/layers/linear.py:76-153 column/row-parallel `param_data.narrow(...)`, layers/embed_head.py:20-32 vocab
shards), as one pure function of the HF parameter name so it can be tested on CPU.  Embeddings and norms stay whole
(the embedding is replicated: the hidden state is needed in full on every rank)."""
from __future__ import annotations

from typing import Tuple

COLUMN_PARALLEL = ("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj", "mlp.gate_proj", "mlp.up_proj")
ROW_PARALLEL = ("self_attn.o_proj", "mlp.down_proj")
SUPPORTED_TP = (1, 2, 4, 8)


def shard_dim(name: str, shard_attention: bool = True) -> Tuple[int, bool]:
    """(dimension to narrow or -1, applies_to_bias).  shard_attention=False: MLP-only tensor parallelism - the attention
    block (q/k/v/o) is replicated on every rank, only gate/up/down and the LM head are sharded (SURVEY 8(e) option A:
    one exchange per transformer block, no head-divisibility constraint)."""
    if name == "lm_head.weight":
        return 0, False
    if not shard_attention and "self_attn." in name:
        return -1, False
    for key in COLUMN_PARALLEL:
        if key in name:
            return 0, True                      # output features: rows of the weight, and the bias
    for key in ROW_PARALLEL:
        if key in name:
            return (1, False) if name.endswith(".weight") else (-1, False)   # input features: columns; bias whole
    return -1, False


def shard_tensor(name: str, t, tp_rank: int, tp_size: int, shard_attention: bool = True):
    """This rank's slice of an HF tensor (a view; the caller makes it contiguous)."""
    if tp_size == 1:
        return t
    dim, _ = shard_dim(name, shard_attention)
    if dim < 0 or (t.dim() == 1 and dim == 1):
        return t
    n = t.shape[dim]
    if n % tp_size:
        raise ValueError(f"{name}: dimension {dim} of size {n} is not divisible by tensor_parallel_size={tp_size}")
    per = n // tp_size
    return t.narrow(dim, tp_rank * per, per)


def check_divisibility(cfg, tp_size: int, shard_attention: bool = True) -> None:
    """The reference asserts head divisibility (inference_engine/models/qwen3.py:31,34); with the attention block
    replicated (shard_attention=False) only the MLP width and the vocabulary have to divide."""
    g = (lambda k: cfg[k] if isinstance(cfg, dict) else getattr(cfg, k))
    keys = ("intermediate_size", "vocab_size")
    if shard_attention:
        keys = ("num_attention_heads", "num_key_value_heads") + keys
    for key in keys:
        if int(g(key)) % tp_size:
            raise ValueError(f"{key}={g(key)} is not divisible by tensor_parallel_size={tp_size}")
    if tp_size not in SUPPORTED_TP:
        raise NotImplementedError("tensor_parallel_size must be 1, 2, 4 or 8 on this build")
    if tp_size > 1 and (int(g("intermediate_size")) // tp_size) % 64:
        raise ValueError(f"intermediate_size / tensor_parallel_size = {int(g('intermediate_size')) // tp_size} must be a "
                         "multiple of 64")


class LocalTPGroup:
    """Rendezvous of tensor-parallel ranks that live in ONE process, one host thread per rank (ranks on several devices
    with peer access, or sharing one device whose SMs they split): what torch.distributed provides between processes."""

    def __init__(self, size: int):
        import threading
        self.size = size
        self.handles = [None] * size
        self._barrier = threading.Barrier(size)

    def barrier(self) -> None:
        self._barrier.wait()
