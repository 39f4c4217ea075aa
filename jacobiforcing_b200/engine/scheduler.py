"""Prefill-first continuous batching over the paged KV cache.

Behaviourally equivalent to the reference's scheduler (inference_engine/engine/scheduler.py:28-97): a step either admits
waiting requests for prefill (bounded by `max_num_seqs`, `max_num_batched_tokens` and free pages) or, when none can be
admitted, schedules the running ones for a decode call, preempting from the tail when pages run out.  Jacobi requests
get their pages from the runner (the draft length is only known there), and `postprocess_jacobi` only decides whether
a request is finished - its tokens were appended by the decode call itself (:80-97)."""
from __future__ import annotations

from collections import deque
from typing import Deque, List, Sequence as Seq, Tuple

from ..config import Config
from .block_manager import BlockManager
from .sequence import Sequence, SequenceStatus


class Scheduler:
    def __init__(self, config: Config):
        self.max_num_seqs = config.max_num_seqs
        self.max_num_batched_tokens = config.max_num_batched_tokens
        self.eos = config.eos
        self.block_manager = BlockManager(config.num_kvcache_blocks, config.kvcache_block_size)
        self.waiting: Deque[Sequence] = deque()
        self.running: Deque[Sequence] = deque()

    # ---- queue -------------------------------------------------------------------------------------------------
    def set_kv_cache(self, kv_cache) -> None:
        self.block_manager.kv_cache = kv_cache

    def add(self, seq: Sequence) -> None:
        self.waiting.append(seq)

    def is_finished(self) -> bool:
        return not (self.waiting or self.running)

    # ---- one scheduling decision -------------------------------------------------------------------------------
    def schedule(self) -> Tuple[List[Sequence], bool]:
        """(sequences, is_prefill)."""
        admitted = self._admit_waiting()
        if admitted:
            return admitted, True
        chosen = self._pick_running()
        if not chosen:
            # the reference asserts here too (scheduler.py:58); say why
            raise AssertionError("nothing can be scheduled: the KV cache / token budget cannot hold a single request "
                                 "(raise num_kvcache_blocks / max_num_batched_tokens or lower jacobi_block_len)")
        self.running.extendleft(reversed(chosen))        # keep them at the head, in order
        return chosen, False

    def _admit_waiting(self) -> List[Sequence]:
        admitted: List[Sequence] = []
        budget = self.max_num_batched_tokens
        while self.waiting and len(admitted) < self.max_num_seqs:
            seq = self.waiting[0]
            if len(seq) > budget or not self.block_manager.can_allocate(seq):
                break
            self.block_manager.allocate(seq)
            budget -= len(seq) - seq.num_cached_tokens
            seq.status = SequenceStatus.RUNNING
            self.running.append(self.waiting.popleft())
            admitted.append(seq)
        return admitted

    @staticmethod
    def jacobi_min_tokens(seq: Sequence) -> int:
        """Tokens the next Jacobi decode call must be able to hold to make progress: one block of accepted tokens plus
        the in-flight margin the device loop writes past them (draft + pseudo block + next token)."""
        L = int(seq.jacobi_block_len)
        remaining = max(1, int(seq.max_tokens) - seq.num_completion_tokens)
        return len(seq) + min(remaining, L) + 3 * L + 1

    def _pick_running(self) -> List[Sequence]:
        chosen: List[Sequence] = []
        while self.running and len(chosen) < self.max_num_seqs:
            seq = self.running.popleft()
            jacobi = str(getattr(seq, "decode_strategy", "")).startswith("jacobi")
            need = self.jacobi_min_tokens(seq) if jacobi else None
            evicted_self = False
            while not self.block_manager.can_append(seq, need):
                if self.running:
                    self.preempt(self.running.pop())     # youngest first
                else:
                    self.preempt(seq)
                    evicted_self = True
                    break
            if evicted_self:
                continue
            if jacobi:
                # reserve the minimum now (several requests may pass can_append against the same free pages); the
                # runner grows the reservation towards the whole remaining budget when pages allow, else caps the call
                self.block_manager.ensure(seq, need)
            else:
                self.block_manager.may_append(seq)
            chosen.append(seq)
        return chosen

    def preempt(self, seq: Sequence) -> None:
        self.block_manager.deallocate(seq)
        seq.status = SequenceStatus.WAITING
        self.waiting.appendleft(seq)

    # ---- after a step ------------------------------------------------------------------------------------------
    def _retire(self, seq: Sequence) -> None:
        seq.status = SequenceStatus.FINISHED
        self.block_manager.deallocate(seq)
        self.running.remove(seq)

    def postprocess(self, seqs: Seq[Sequence], token_ids: Seq[int]) -> None:
        for seq, tok in zip(seqs, token_ids):
            seq.append_token(tok)
            hit_eos = tok == self.eos and not seq.ignore_eos
            if hit_eos or seq.num_completion_tokens == seq.max_tokens:
                self._retire(seq)

    def postprocess_jacobi(self, seqs: Seq[Sequence], token_ids_batch: Seq[Seq[int]]) -> None:
        for seq, new_tokens in zip(seqs, token_ids_batch):
            hit_eos = (not seq.ignore_eos) and self.eos in new_tokens
            if hit_eos or seq.num_completion_tokens >= seq.max_tokens:
                self._retire(seq)
