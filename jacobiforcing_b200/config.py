"""Config — the reference's engine configuration (inference_engine/config.py:7-52), plus two ways to run without a
checkpoint directory (there are no model files offline): `hf_config` may be passed in, and `state_dict` may carry
HF-named tensors for the loader."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Optional


@dataclass
class Config:
    model: Optional[str]
    max_num_batched_tokens: int = 16384
    max_num_seqs: int = 512
    max_model_len: int = 8192
    gpu_memory_utilization: float = 0.9
    tensor_parallel_size: int = 1
    enforce_eager: bool = False
    hf_config: Any = None
    eos: int = -1
    pad: int = -1
    kvcache_block_size: int = 256
    num_kvcache_blocks: int = -1

    jacobi_enabled: bool = True
    jacobi_block_len: int = 64
    jacobi_max_blocks: int = 2
    jacobi_spawn_ratio: float = 0.8
    jacobi_lookahead_start_ratio: float = 0.0
    jacobi_n_gram_pool_size: int = 4
    jacobi_max_iterations: int = 128

    # extensions of this build
    state_dict: Any = field(default=None, repr=False)     # HF-named tensors instead of a checkpoint directory
    jacobi_multiblock_enabled: bool = False                # serve "jacobi_multiblock_rejection_recycling" on device
    draft_seed: int = 0                                    # seeds every random draft token (reference: unseeded RNGs)
    device: int = 0
    tp_shard_attention: Optional[bool] = None              # True: heads sharded like the reference (heads must divide tp);
                                                           # False: MLP-only tensor parallelism (attention replicated on
                                                           # every rank: one exchange per block, any tp in {2,4,8});
                                                           # None: shard when the heads divide, else replicate

    tp_devices: Any = None                                 # CUDA device of every rank when the ranks are threads of this
                                                           # process (default: rank r -> device r mod device count)

    def __post_init__(self):
        self.model = _resolve_checkpoint_dir(self.model)
        if self.kvcache_block_size % 256:
            raise AssertionError("kvcache_block_size must be a multiple of 256")
        if not 1 <= self.tensor_parallel_size <= 8:
            raise AssertionError("tensor_parallel_size must lie in [1, 8]")
        if self.hf_config is None:
            if self.model is None or not os.path.isdir(self.model):
                raise AssertionError(f"Model path does not exist: {self.model}")
            from transformers import AutoConfig
            self.hf_config = AutoConfig.from_pretrained(self.model)
        self.max_model_len = min(self.max_model_len, self.hf_config.max_position_embeddings)
        if self.max_num_batched_tokens < self.max_model_len:
            raise AssertionError("max_num_batched_tokens must cover max_model_len")


def _resolve_checkpoint_dir(path: Optional[str]) -> Optional[str]:
    """A training output directory is replaced by its latest `checkpoint-N` sub-directory, as the reference does
    (inference_engine/config.py:31-41)."""
    if path is None or not os.path.isdir(path):
        return path
    steps = []
    for name in os.listdir(path):
        head, _, tail = name.partition("-")
        if head == "checkpoint" and tail.isdigit() and os.path.isdir(os.path.join(path, name)):
            steps.append(int(tail))
    return os.path.join(path, f"checkpoint-{max(steps)}") if steps else path
