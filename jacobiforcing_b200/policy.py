"""Host-side preparation of the device policy inputs (csrc/jf_policy.h).

The accept / shift / re-draft state machine itself runs inside the persistent kernel; the host only turns the
reference's floating-point knobs into the integers the device compares against, and pre-draws the random draft
choices the reference's driver makes between calls (JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:176-180,
210-215 use the unseeded global `random`; here the RNG is an explicit argument - SURVEY.md finding 10).
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass
from typing import List, Optional, Sequence

MODE_HF_SINGLE, MODE_HF_MULTI, MODE_ENGINE, MODE_ENGINE_NG = 0, 1, 2, 3
MODES = {"single": MODE_HF_SINGLE, "multiblock": MODE_HF_MULTI, "engine": MODE_ENGINE, "engine_nongreedy": MODE_ENGINE_NG}
STOP_REASONS = {0: "", 1: "eos", 2: "max_new_tokens", 3: "max_calls", 4: "max_iterations", 5: "error",
                }
POLICY_ERRORS = {0: "", 1: "reference would fail to broadcast draft rows against greedy rows",
                 2: "pad_token_id must be provided when spawning pseudo-active blocks.",
                 3: "block list capacity exceeded (K >= 3 bookkeeping drift is not supported on device)",
                 4: "draft size mismatch at promotion", 5: "policy capacity exceeded", 6: "draft source exhausted"}
N_MAX, ROWS_MAX, POOL_MAX = 128, 16, 16


@dataclass
class PolicyKnobs:
    mode: int
    n: int
    K: int = 2
    r: float = 0.85
    lookahead_start_ratio: float = 0.0
    n_gram_pool_size: int = 4
    eos: Optional[int] = None
    pad: Optional[int] = None
    max_iteration_count: int = 128
    max_new_tokens: int = 512
    max_calls: int = 1024
    vocab: int = 0
    # non-greedy engine mode: sampling temperature, 64-bit key of the counter-based generator, counter offset
    temperature: float = 0.0
    seed: int = 0
    rng_offset: int = 0

    def validate(self) -> None:
        if not (1 <= self.n <= N_MAX):
            raise ValueError(f"block length n={self.n} outside [1, {N_MAX}]")
        if self.mode in (MODE_ENGINE, MODE_ENGINE_NG) and self.n < 2:
            raise ValueError("Jacobi draft length L must be >= 2")          # jacobi_decoding.py:316-317
        if self.mode == MODE_ENGINE_NG and not self.temperature > 0.0:
            raise ValueError("non-greedy Jacobi decoding needs temperature > 0")
        if self.mode == MODE_HF_MULTI:
            if not (1 <= self.K <= 2):
                raise NotImplementedError("K >= 3 relies on the reference's block-list drift (SURVEY.md finding 5); "
                                          "the device policy supports K in {1, 2}")
            if not (0 <= self.n_gram_pool_size <= POOL_MAX):
                raise ValueError(f"n_gram_pool_size outside [0, {POOL_MAX}]")

    def as_ints(self) -> List[int]:
        """Field order of struct jf::PolicyKnobs."""
        n = self.n
        spawn = int(math.ceil(self.r * n))                                     # unified.py:262
        lsr_min = next((a for a in range(0, 2 * N_MAX + 1) if a / n >= self.lookahead_start_ratio), 1 << 30)
        return [self.mode, n, self.K, spawn, self.n_gram_pool_size,
                -1 if self.eos is None else int(self.eos), -1 if self.pad is None else int(self.pad),
                self.max_iteration_count, self.max_new_tokens, self.max_calls, lsr_min, self.vocab,
                struct.unpack("<i", struct.pack("<f", float(self.temperature)))[0],
                int(self.seed) & 0xFFFFFFFF, (int(self.seed) >> 32) & 0xFFFFFFFF, int(self.rng_offset)]


def mt_state_words(rng) -> List[int]:
    """The 625 words (624 of MT19937 state + position) of a random.Random: handed to the device, which continues the
    generator exactly where the host left it (csrc/jf_policy.h mt_randbelow == random.Random._randbelow)."""
    version, internal, _gauss = rng.getstate()
    if version != 3 or len(internal) != 625:
        raise ValueError("unsupported random.Random state layout")
    return [int(x) for x in internal]


def draw_prefill_draft(rng, prompt: Sequence[int], n: int) -> List[int]:
    """n x random.choice(prompt ids) (MR_humaneval.py:176-180)."""
    p = list(prompt)
    return [rng.choice(p) for _ in range(n)]
