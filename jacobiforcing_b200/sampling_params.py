"""SamplingParams — same fields, defaults and validation as the reference (inference_engine/sampling_params.py:5-38)."""
from dataclasses import dataclass


@dataclass
class SamplingParams:
    temperature: float = 1.0
    max_tokens: int = 64
    ignore_eos: bool = False

    # "autoregressive", "jacobi", "jacobi_multiblock_rejection_recycling"
    decode_strategy: str = "autoregressive"

    jacobi_block_len: int = 64          # n_token_seq_len
    jacobi_max_iterations: int = 128

    jacobi_max_blocks: int = 2           # K
    jacobi_spawn_ratio: float = 0.85     # r
    jacobi_lookahead_start_ratio: float = 0.0
    jacobi_n_gram_pool_size: int = 4

    jacobi_on_policy: bool = False

    def __post_init__(self):
        assert self.temperature >= 0.0, "temperature must be non-negative"
        if self.jacobi_on_policy and self.temperature == 0.0:
            raise ValueError(
                "jacobi_on_policy=True requires temperature > 0 (non-greedy decoding). "
                "On-policy learning is only supported with non-greedy Jacobi decoding."
            )

    @property
    def use_jacobi(self) -> bool:
        return self.jacobi_block_len is not None
