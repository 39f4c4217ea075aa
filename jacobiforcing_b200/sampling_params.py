"""Per-request sampling / decoding parameters.

Field names, order, defaults and validation follow the reference's dataclass (inference_engine/sampling_params.py:5-38)
because requests written against it must construct unchanged; only `SamplingParams(...)` carries Jacobi settings (the
reference's `generate(..., jacobi_block_len=...)` keyword route replaces fields that do not exist, SURVEY finding 9)."""
from dataclasses import dataclass

DECODE_STRATEGIES = ("autoregressive", "jacobi", "jacobi_multiblock_rejection_recycling")


@dataclass
class SamplingParams:
    # sampling
    temperature: float = 1.0
    max_tokens: int = 64
    ignore_eos: bool = False
    # one of DECODE_STRATEGIES
    decode_strategy: str = "autoregressive"
    # Jacobi block: n draft positions verified per forward, iteration cap per decode call
    jacobi_block_len: int = 64
    jacobi_max_iterations: int = 128
    # multiblock + rejection recycling: K blocks in flight, spawn ratio r, lookahead gate, n-gram pool
    jacobi_max_blocks: int = 2
    jacobi_spawn_ratio: float = 0.85
    jacobi_lookahead_start_ratio: float = 0.0
    jacobi_n_gram_pool_size: int = 4
    # on-policy rollout records (non-greedy only)
    jacobi_on_policy: bool = False

    def __post_init__(self):
        assert self.temperature >= 0.0, "temperature must be non-negative"
        greedy = self.temperature == 0.0
        if self.jacobi_on_policy and greedy:
            raise ValueError("jacobi_on_policy=True requires temperature > 0 (non-greedy decoding). "
                             "On-policy learning is only supported with non-greedy Jacobi decoding.")

    @property
    def use_jacobi(self) -> bool:
        return self.jacobi_block_len is not None
