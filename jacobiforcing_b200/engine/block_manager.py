"""BlockManager — page bookkeeping for the paged KV cache (inference_engine/engine/block_manager.py).

Two jobs.  (1) Jacobi path (:114-121, :267-276, :534-564): pages are allocated WITHOUT clearing and rolled back logically
- the attention kernel trusts the committed length, never the page contents.  (2) Prefill prefix cache (:135-168): every
FULL page of a sequence is identified by a hash chained over all tokens up to and including it; a new request whose
leading pages carry known hashes (and the same token ids - hashes are verified, never trusted) shares those pages
(reference counted) and starts with `num_cached_tokens` > 0, so the prefill forward skips them.  A freed page keeps its
identity until it is handed out again; the free list is FIFO, so recently used prefixes survive longest.
A page becomes shareable when its tokens are known to be committed: the prompt's full pages at allocation (the prefill
that fills them runs before any later request's), everything else when the sequence is released."""
from __future__ import annotations

import struct
from collections import deque
from typing import Dict, List

from .sequence import Sequence

try:                                     # the reference hashes with xxhash64 (:71-97); any 64-bit hash serves
    import xxhash

    def _hash64(data: bytes) -> int:
        return xxhash.xxh64(data).intdigest()
except Exception:                        # pragma: no cover - xxhash is present in this image
    import hashlib

    def _hash64(data: bytes) -> int:
        return int.from_bytes(hashlib.blake2b(data, digest_size=8).digest(), "little")


class Block:
    __slots__ = ("block_id", "ref_count", "hash", "token_ids")

    def __init__(self, block_id: int):
        self.block_id = block_id
        self.ref_count = 0
        self.hash = -1
        self.token_ids: List[int] = []


class BlockManager:
    def __init__(self, num_blocks: int, block_size: int = 256, enable_prefix_cache: bool = True):
        assert num_blocks > 0
        self.block_size = block_size
        self.num_blocks = num_blocks
        self.blocks = [Block(i) for i in range(num_blocks)]
        self.hash_to_block_id: Dict[int, int] = {}
        self.free_block_ids = deque(range(num_blocks))
        self.used_block_ids = set()
        self.kv_cache = None
        self.enable_prefix_cache = enable_prefix_cache
        self.stats = {"prefix_hits": 0, "prefix_hit_tokens": 0}

    # ---- hashing -------------------------------------------------------------------------------------------------
    @classmethod
    def compute_hash(cls, token_ids, prefix: int = -1) -> int:
        """64-bit hash of one full page chained to the hash of the page before it (:71-97)."""
        head = b"" if prefix == -1 else struct.pack("<Q", prefix & 0xFFFFFFFFFFFFFFFF)
        return _hash64(head + struct.pack(f"<{len(token_ids)}q", *[int(t) for t in token_ids]))

    def _register(self, block: Block, h: int, token_ids) -> None:
        block.hash = h
        block.token_ids = list(token_ids)
        self.hash_to_block_id[h] = block.block_id

    def _lookup(self, h: int, token_ids) -> int:
        """Page id holding exactly these tokens under this chained hash, or -1."""
        b = self.hash_to_block_id.get(h, -1)
        if b == -1:
            return -1
        blk = self.blocks[b]
        if blk.hash != h or blk.token_ids != list(token_ids):
            return -1
        return b

    # ---- allocation ----------------------------------------------------------------------------------------------
    def _take(self, block_id: int = -1) -> int:
        """A free page (a specific one when asked); its previous identity is forgotten unless it is being revived."""
        if not self.free_block_ids:
            raise RuntimeError("out of KV cache blocks")
        if block_id == -1:
            block_id = self.free_block_ids.popleft()
            blk = self.blocks[block_id]
            if blk.hash != -1 and self.hash_to_block_id.get(blk.hash) == block_id:
                del self.hash_to_block_id[blk.hash]
            blk.hash, blk.token_ids = -1, []
        else:
            self.free_block_ids.remove(block_id)
        blk = self.blocks[block_id]
        assert blk.ref_count == 0
        blk.ref_count = 1
        self.used_block_ids.add(block_id)
        return block_id

    def can_allocate(self, seq: Sequence) -> bool:
        return len(self.free_block_ids) >= seq.num_blocks + 1

    def allocate(self, seq: Sequence) -> None:
        """Pages for the whole sequence; leading full pages found in the prefix cache are shared (:135-168)."""
        assert not seq.block_table
        h = -1
        miss = not self.enable_prefix_cache
        for i in range(seq.num_blocks):
            token_ids = seq.block(i)
            full = len(token_ids) == self.block_size
            h = self.compute_hash(token_ids, h) if (full and self.enable_prefix_cache) else -1
            block_id = self._lookup(h, token_ids) if (h != -1 and not miss) else -1
            if block_id == -1:
                miss = True                                  # a chain cannot resume after a miss
                block_id = self._take()
            else:
                seq.num_cached_tokens += self.block_size
                self.stats["prefix_hits"] += 1
                self.stats["prefix_hit_tokens"] += self.block_size
                if block_id in self.used_block_ids:
                    self.blocks[block_id].ref_count += 1
                else:
                    self._take(block_id)                     # revive a freed page that still holds the prefix
            if h != -1:
                self._register(self.blocks[block_id], h, token_ids)
            seq.block_table.append(block_id)

    def register_committed(self, seq: Sequence, n_tokens: int) -> None:
        """Make the full pages among the first n_tokens (all committed to the KV cache) shareable."""
        if not self.enable_prefix_cache:
            return
        h = -1
        for i in range(min(n_tokens // self.block_size, len(seq.block_table))):
            token_ids = seq.token_ids[i * self.block_size:(i + 1) * self.block_size]
            h = self.compute_hash(token_ids, h)
            blk = self.blocks[seq.block_table[i]]
            if blk.hash == -1 and self._lookup(h, token_ids) == -1:
                self._register(blk, h, token_ids)

    def deallocate(self, seq: Sequence) -> None:
        self.register_committed(seq, min(seq.num_cached_tokens, len(seq) - 1))
        for b in reversed(seq.block_table):
            blk = self.blocks[b]
            blk.ref_count -= 1
            if blk.ref_count == 0:
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
