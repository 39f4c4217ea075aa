"""Request state seen by the scheduler, the block manager and the model runner.

Mirrors the attribute surface of the reference's `Sequence` (inference_engine/engine/sequence.py:14-156: token ids,
prompt / completion split, block table, the per-request Jacobi knobs copied from `SamplingParams`, draft fields) so the
reference's engine code and tests can use it unchanged; the implementation is this repository's own."""
from __future__ import annotations

import itertools
from enum import Enum, auto
from typing import Iterable, List, Optional

from ..sampling_params import SamplingParams

_JACOBI_KNOBS = ("jacobi_block_len", "jacobi_max_blocks", "jacobi_spawn_ratio", "jacobi_lookahead_start_ratio",
                 "jacobi_n_gram_pool_size", "jacobi_max_iterations", "jacobi_on_policy")


class SequenceStatus(Enum):
    WAITING = auto()
    RUNNING = auto()
    FINISHED = auto()


def _as_int(tok) -> int:
    return int(tok.item()) if hasattr(tok, "item") else int(tok)


class Sequence:
    block_size = 256                      # KV page size in tokens
    counter = itertools.count()           # process-wide request ids

    def __init__(self, token_ids: List[int], sampling_params: Optional[SamplingParams] = None):
        sp = sampling_params if sampling_params is not None else SamplingParams()
        self.seq_id = next(Sequence.counter)
        self.status = SequenceStatus.WAITING
        self.token_ids = list(token_ids)
        self.num_prompt_tokens = self.num_tokens = len(self.token_ids)
        self.last_token = self.token_ids[-1]
        self.num_cached_tokens = 0
        self.block_table: List[int] = []
        # per-request sampling / decoding knobs
        self.sampling_params = sp
        self.temperature, self.max_tokens, self.ignore_eos = sp.temperature, sp.max_tokens, sp.ignore_eos
        self.decode_strategy = sp.decode_strategy
        for knob in _JACOBI_KNOBS:
            setattr(self, knob, getattr(sp, knob))
        # Jacobi draft state
        self.draft_tokens = None
        self._prefill_draft = None
        self._first_correct = None        # HF-convention strategies: first_correct_token of the prefill

    # ---- container protocol ------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return self.num_tokens

    def __getitem__(self, key):
        return self.token_ids[key]

    # ---- derived quantities ------------------------------------------------------------------------------------
    @property
    def is_finished(self) -> bool:
        return self.status is SequenceStatus.FINISHED

    @property
    def num_completion_tokens(self) -> int:
        return self.num_tokens - self.num_prompt_tokens

    @property
    def prompt_token_ids(self) -> List[int]:
        return self.token_ids[:self.num_prompt_tokens]

    @property
    def completion_token_ids(self) -> List[int]:
        return self.token_ids[self.num_prompt_tokens:]

    @property
    def num_blocks(self) -> int:
        return -(-self.num_tokens // self.block_size)

    @property
    def num_cached_blocks(self) -> int:
        return self.num_cached_tokens // self.block_size

    def block(self, i: int) -> List[int]:
        """Token ids of page i (the last one may be partial)."""
        assert 0 <= i < self.num_blocks
        return self.token_ids[i * self.block_size:(i + 1) * self.block_size]

    # ---- growth ------------------------------------------------------------------------------------------------
    def append_token(self, token_id) -> None:
        if isinstance(token_id, (list, tuple)):
            raise ValueError(f"append_token expects a single integer, got {type(token_id)} with {len(token_id)} elements")
        self.extend_tokens((token_id,))

    def extend_tokens(self, tokens: Iterable) -> None:
        new = [_as_int(t) for t in tokens] if tokens else []
        if new:
            self.token_ids += new
            self.last_token = new[-1]
            self.num_tokens = len(self.token_ids)

    # ---- drafts ------------------------------------------------------------------------------------------------
    def has_draft(self) -> bool:
        return bool(self.draft_tokens is not None and len(self.draft_tokens))

    def clear_draft(self) -> None:
        self.draft_tokens = None
