"""Host side of the engine surface (no GPU, no CUDA library): SamplingParams / Sequence / BlockManager / Scheduler follow
the reference's contracts (inference_engine/sampling_params.py:5-38, engine/sequence.py:14-156,
engine/block_manager.py:114-121, 534-564, engine/scheduler.py:28-97)."""
import types

import pytest

from jacobiforcing_b200.engine.block_manager import BlockManager
from jacobiforcing_b200.engine.scheduler import Scheduler
from jacobiforcing_b200.engine.sequence import Sequence, SequenceStatus
from jacobiforcing_b200.sampling_params import SamplingParams


def _cfg(**kw):
    d = dict(max_num_seqs=4, max_num_batched_tokens=4096, eos=7, num_kvcache_blocks=8, kvcache_block_size=256)
    d.update(kw)
    return types.SimpleNamespace(**d)


def test_sampling_params_defaults_and_validation():
    sp = SamplingParams()
    assert (sp.temperature, sp.max_tokens, sp.ignore_eos, sp.decode_strategy) == (1.0, 64, False, "autoregressive")
    assert (sp.jacobi_block_len, sp.jacobi_max_iterations, sp.jacobi_max_blocks) == (64, 128, 2)
    assert sp.jacobi_spawn_ratio == 0.85 and sp.jacobi_n_gram_pool_size == 4 and sp.use_jacobi
    with pytest.raises(AssertionError):
        SamplingParams(temperature=-1.0)
    with pytest.raises(ValueError):                      # on-policy needs non-greedy decoding (sampling_params.py:27-32)
        SamplingParams(temperature=0.0, jacobi_on_policy=True)


def test_sequence_bookkeeping():
    s = Sequence([1, 2, 3], SamplingParams(max_tokens=5, decode_strategy="jacobi", jacobi_block_len=8))
    assert len(s) == 3 and s.num_prompt_tokens == 3 and s.num_completion_tokens == 0 and s.num_blocks == 1
    s.append_token(9)
    s.extend_tokens([4, 5])
    assert s.token_ids == [1, 2, 3, 9, 4, 5] and s.last_token == 5 and s.completion_token_ids == [9, 4, 5]
    with pytest.raises(ValueError):                      # the reference's guard against passing a list (sequence.py:104-110)
        s.append_token([1, 2])
    s.extend_tokens([])                                  # no-op
    assert len(s) == 6 and not s.has_draft()
    s.draft_tokens = [1]
    assert s.has_draft()
    s.clear_draft()
    assert not s.has_draft()
    big = Sequence(list(range(257)))
    assert big.num_blocks == 2                           # page = 256 tokens


def test_block_manager_allocate_ensure_trim_deallocate():
    bm = BlockManager(4, 256)
    s = Sequence(list(range(300)))
    assert bm.can_allocate(s)                            # needs num_blocks + 1 free (block_manager.py can_allocate)
    bm.allocate(s)
    assert len(s.block_table) == 2 and len(bm.free_block_ids) == 2
    bm.ensure(s, 600)                                    # _allocate_block_no_clear until the table covers 600 tokens
    assert len(s.block_table) == 3
    s.num_cached_tokens = 300
    bm.trim_kv_only_fast(s, 40)                          # logical roll-back only
    assert s.num_cached_tokens == 260 and len(s.block_table) == 3
    bm.trim_kv_only_fast(s, 1000)
    assert s.num_cached_tokens == 0
    with pytest.raises(RuntimeError):
        bm.ensure(s, 256 * 5)                            # out of blocks
    bm.deallocate(s)
    assert not s.block_table and len(bm.free_block_ids) == 4 and not bm.used_block_ids
    t = Sequence(list(range(256 * 4)))
    assert not bm.can_allocate(t)                        # 4 blocks + 1 spare do not fit in 4


def test_scheduler_prefill_first_then_decode_and_finish():
    sch = Scheduler(_cfg())
    a = Sequence([1, 2, 3], SamplingParams(max_tokens=2, decode_strategy="jacobi"))
    b = Sequence([4, 5], SamplingParams(max_tokens=4, decode_strategy="jacobi", ignore_eos=True))
    sch.add(a)
    sch.add(b)
    seqs, is_prefill = sch.schedule()
    assert is_prefill and seqs == [a, b] and all(s.status == SequenceStatus.RUNNING for s in seqs)
    assert a.block_table and b.block_table
    seqs, is_prefill = sch.schedule()                    # nothing waiting: decode step over the running sequences
    assert not is_prefill and seqs == [a, b]
    # the Jacobi decoder has already appended the tokens; the scheduler only checks EOS / max_tokens (:80-97)
    a.extend_tokens([9, 9])
    b.extend_tokens([7, 1])                              # contains EOS (7) but ignore_eos is set
    sch.postprocess_jacobi([a, b], [[9, 9], [7, 1]])
    assert a.is_finished and not a.block_table           # max_tokens reached -> pages returned
    assert not b.is_finished
    b.extend_tokens([1, 1])
    sch.postprocess_jacobi([b], [[1, 1]])
    assert b.is_finished and sch.is_finished()


def test_scheduler_eos_finishes_and_autoregressive_postprocess():
    sch = Scheduler(_cfg())
    a = Sequence([1, 2], SamplingParams(max_tokens=8))
    sch.add(a)
    sch.schedule()
    sch.schedule()
    sch.postprocess([a], [5])
    assert not a.is_finished and a.token_ids[-1] == 5
    sch.postprocess([a], [7])                            # EOS
    assert a.is_finished and sch.is_finished()


def test_scheduler_respects_limits_and_preempts():
    sch = Scheduler(_cfg(max_num_seqs=1, num_kvcache_blocks=3))
    a = Sequence(list(range(10)), SamplingParams(decode_strategy="jacobi"))
    b = Sequence(list(range(10)), SamplingParams(decode_strategy="jacobi"))
    sch.add(a)
    sch.add(b)
    seqs, is_prefill = sch.schedule()
    assert is_prefill and seqs == [a]                    # max_num_seqs = 1
    seqs, is_prefill = sch.schedule()                    # b is still waiting and fits: prefill first
    assert is_prefill and seqs == [b]
    # exhaust the pool: the decode scheduling must preempt the youngest running sequence (scheduler.py:47-56)
    sch.block_manager.free_block_ids.clear()
    seqs, is_prefill = sch.schedule()
    assert not is_prefill and seqs == [a]
    assert b.status == SequenceStatus.WAITING and not b.block_table and sch.waiting[0] is b
    big = Scheduler(_cfg(max_num_batched_tokens=8))
    big.add(Sequence(list(range(9))))
    with pytest.raises(AssertionError):                  # nothing can be scheduled: the reference asserts too (:58)
        big.schedule()


def test_jacobi_requests_under_kv_memory_pressure_finish_without_crashing():
    """8 Jacobi requests with max_tokens=2048 on a 40-page cache.  The scheduler reserves one block's worth
    per admitted request and preempts the youngest when pages run out; the runner shortens a decode call whose whole
    remaining budget does not fit instead of raising 'out of KV cache blocks'."""
    from types import SimpleNamespace
    from jacobiforcing_b200.engine.model_runner import ModelRunner
    cfg = _cfg()
    cfg.num_kvcache_blocks = 40
    sch = Scheduler(cfg)
    bm = sch.block_manager
    runner = SimpleNamespace(block_manager=bm)
    L = 32
    sp = SamplingParams(temperature=0.0, max_tokens=2048, decode_strategy="jacobi", jacobi_block_len=L, ignore_eos=True)
    reqs = [Sequence(list(range(1, 300 + 7 * i)), sp) for i in range(8)]
    for r in reqs:
        sch.add(r)
    steps = preempted = 0
    while not sch.is_finished():
        steps += 1
        assert steps < 2000, "scheduler does not make progress"
        before = len(sch.waiting)
        seqs, is_prefill = sch.schedule()
        if not is_prefill:
            preempted += max(0, len(sch.waiting) - before)
        if is_prefill:
            for s in seqs:
                s.num_cached_tokens = len(s)
            continue
        remaining = max(1, sp.max_tokens - seqs[0].num_completion_tokens)
        groups = {}
        for s in seqs:
            groups.setdefault(sp.max_tokens - s.num_completion_tokens, []).append(s)
        outs = {}
        for rem, grp in groups.items():
            new = ModelRunner._fit_new_tokens(runner, grp, max(1, rem), L)
            assert 1 <= new <= max(1, rem)
            for s in grp:
                bm.ensure(s, len(s) + new + 3 * L + 1)        # what _run_jacobi_group reserves; must not raise
                toks = [7] * new
                s.extend_tokens(toks)
                outs[s.seq_id] = toks
        sch.postprocess_jacobi(seqs, [outs[s.seq_id] for s in seqs])
    assert all(r.num_completion_tokens >= 2048 for r in reqs)
    assert len(bm.free_block_ids) == 40 and not bm.used_block_ids
    assert preempted > 0                                       # the pressure was real


def test_can_append_predicates():
    bm = BlockManager(1, 256)
    s = Sequence(list(range(256)))
    bm.allocate(s)
    assert not bm.free_block_ids
    assert bm.can_append(s)                                    # reference predicate: only a page-opening token needs one
    s.extend_tokens([1])
    assert not bm.can_append(s)                                # len % 256 == 1 and nothing free
    assert bm.can_append(s, need_tokens=256) and not bm.can_append(s, need_tokens=258)


# ---- prefix cache (reference: inference_engine/engine/block_manager.py:71-97 hash chain, :135-168 allocate) ------------
def _seq(tokens, **kw):
    from jacobiforcing_b200.engine.sequence import Sequence
    from jacobiforcing_b200.sampling_params import SamplingParams
    return Sequence(list(tokens), SamplingParams(temperature=0.0, **kw))


def test_prefix_cache_shares_full_pages_and_verifies_tokens():
    from jacobiforcing_b200.engine.block_manager import BlockManager
    bm = BlockManager(16, 256)
    base = list(range(600))                       # 2 full pages + 88 tokens
    a = _seq(base)
    bm.allocate(a)
    assert a.num_cached_tokens == 0 and len(a.block_table) == 3
    b = _seq(base[:512] + [7] * 50)               # same first two pages, different tail
    bm.allocate(b)
    assert b.num_cached_tokens == 512 and b.block_table[:2] == a.block_table[:2] and b.block_table[2] != a.block_table[2]
    assert bm.blocks[a.block_table[0]].ref_count == 2
    c = _seq(base[:256] + [9] * 300)              # only the first page matches; the chain stops at the first miss
    bm.allocate(c)
    assert c.num_cached_tokens == 256 and c.block_table[0] == a.block_table[0] and c.block_table[1] != a.block_table[1]
    d = _seq([5] + base[1:600])                   # differs in the first token: nothing shared
    bm.allocate(d)
    assert d.num_cached_tokens == 0 and not set(d.block_table) & set(a.block_table)
    # a shared page is freed only with its last user, and keeps its identity while it sits on the free list
    shared = a.block_table[0]
    for s in (a, b, c):
        s.num_cached_tokens = len(s)
    bm.deallocate(a); bm.deallocate(b)
    assert shared in bm.used_block_ids
    bm.deallocate(c)
    assert shared not in bm.used_block_ids
    e = _seq(base)
    bm.allocate(e)
    assert e.num_cached_tokens == 512 and e.block_table[0] == shared      # revived from the free list
    assert bm.stats["prefix_hit_tokens"] == 512 + 256 + 512


def test_prefix_cache_forgets_a_page_that_is_handed_out_again():
    from jacobiforcing_b200.engine.block_manager import BlockManager
    bm = BlockManager(3, 256)
    a = _seq(list(range(300)))
    bm.allocate(a)
    a.num_cached_tokens = 300
    bm.deallocate(a)
    b = _seq([1] * 700)                           # needs all three pages: the cached one is recycled
    bm.allocate(b)
    b.num_cached_tokens = 0
    bm.deallocate(b)
    c = _seq(list(range(300)))
    bm.allocate(c)
    assert c.num_cached_tokens == 0               # the hash entry died with the page's reuse


def test_generated_pages_become_shareable_when_the_sequence_is_released():
    from jacobiforcing_b200.engine.block_manager import BlockManager
    bm = BlockManager(8, 256)
    a = _seq(list(range(100)))
    bm.allocate(a)
    a.extend_tokens(list(range(1000, 1000 + 500)))        # decode grew it to 600 tokens, all committed but the last
    bm.ensure(a, 600)
    a.num_cached_tokens = 600
    toks = list(a.token_ids)
    bm.deallocate(a)
    b = _seq(toks[:520])
    bm.allocate(b)
    assert b.num_cached_tokens == 512


def test_disabled_prefix_cache_never_shares():
    from jacobiforcing_b200.engine.block_manager import BlockManager
    bm = BlockManager(8, 256, enable_prefix_cache=False)
    a, b = _seq(list(range(600))), _seq(list(range(600)))
    bm.allocate(a); bm.allocate(b)
    assert b.num_cached_tokens == 0 and not set(a.block_table) & set(b.block_table)
