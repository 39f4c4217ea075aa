"""BlockManager — page bookkeeping for the paged KV cache (inference_engine/engine/block_manager.py).  The hot path of
this build never clears or hashes pages (the reference's Jacobi path also allocates without clearing and rolls back
logically, :114-121, :534-564); the xxhash prefix cache is a prefill feature outside the hot path."""
from __future__ import annotations

from collections import deque

from .sequence import Sequence


class BlockManager:
    def __init__(self, num_blocks: int, block_size: int = 256):
        assert num_blocks > 0
        self.block_size = block_size
        self.num_blocks = num_blocks
        self.free_block_ids = deque(range(num_blocks))
        self.used_block_ids = set()
        self.kv_cache = None

    def _take(self) -> int:
        if not self.free_block_ids:
            raise RuntimeError("out of KV cache blocks")
        b = self.free_block_ids.popleft()
        self.used_block_ids.add(b)
        return b

    def can_allocate(self, seq: Sequence) -> bool:
        return len(self.free_block_ids) >= seq.num_blocks + 1

    def allocate(self, seq: Sequence) -> None:
        assert not seq.block_table
        for _ in range(seq.num_blocks):
            seq.block_table.append(self._take())

    def deallocate(self, seq: Sequence) -> None:
        for b in reversed(seq.block_table):
            self.used_block_ids.discard(b)
            self.free_block_ids.append(b)
        seq.num_cached_tokens = 0
        seq.block_table.clear()

    def pages_missing(self, seq: Sequence, n_tokens: int) -> int:
        """Pages the table lacks to cover n_tokens."""
        return max(0, (n_tokens + self.block_size - 1) // self.block_size - len(seq.block_table))

    def can_append(self, seq: Sequence, need_tokens: int = None) -> bool:
        """Reference predicate (:192-193): a page is needed only when the next token opens one.  A Jacobi decode call
        appends many tokens; the scheduler passes the token count the call must be able to hold."""
        if need_tokens is not None:
            return len(self.free_block_ids) >= self.pages_missing(seq, need_tokens)
        return len(self.free_block_ids) >= (len(seq) % self.block_size == 1)

    def may_append(self, seq: Sequence) -> None:
        self.ensure(seq, len(seq))

    def may_append_batch(self, seq: Sequence, num_tokens: int) -> None:
        self.ensure(seq, len(seq))

    def ensure(self, seq: Sequence, n_tokens: int) -> None:
        """_allocate_block_no_clear (:114-121) until the table covers n_tokens."""
        need = (n_tokens + self.block_size - 1) // self.block_size
        while len(seq.block_table) < need:
            seq.block_table.append(self._take())

    def trim_kv_only_fast(self, seq: Sequence, num_tokens: int) -> None:
        """Logical roll-back (:534-564): nothing is cleared; attention trusts the committed length."""
        seq.num_cached_tokens = max(0, seq.num_cached_tokens - num_tokens)
