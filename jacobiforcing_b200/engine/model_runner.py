"""ModelRunner — the per-rank operator the engine calls (`run(seqs, is_prefill)`), mirroring
inference_engine/engine/model_runner.py:155-291 (init), :777-963 (Jacobi prefill with drafting) and :1464-1550 (run
dispatch), with the Python JacobiDecoder loop + per-iteration forwards replaced by ONE device-resident loop per call
(jf_jacobi_run) behind the C ABI.  `jacobi_decoder.stats` keeps the six keys the reference's tests read
(engine/jacobi_decoding.py:101-108).
"""
from __future__ import annotations

import glob
import os
import random
from typing import Dict, List, Optional

from ..config import Config
from ..policy import MODES, PolicyKnobs, mt_state_words
from ..runtime import PAGE, JacobiRuntime, SeqHandle
from .sequence import Sequence


class _DecoderStats:
    """Stand-in for JacobiDecoder: only its `stats` dictionary is part of the surface the tests read."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.stats = {"num_chunk_calls": 0, "num_jacobi_iterations": 0, "tokens_accepted": 0, "tokens_per_call": [],
                      "tokens_per_iteration": [], "iterations_per_call": []}


def load_model(runner: "ModelRunner", path: str) -> None:
    """Same walk as inference_engine/utils/loader.py:12-47: every tensor of every *.safetensors file is handed to the
    weight loader under its HF name; q/k/v and gate/up packing happens behind jf_load_tensor."""
    from safetensors import safe_open
    files = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    for file in files:
        with safe_open(file, "pt", "cpu") as f:
            for name in f.keys():
                if name.endswith("rotary_emb.inv_freq"):
                    continue
                runner.rt.load_tensor(name, f.get_tensor(name))
    runner.rt.finalize()


class ModelRunner:
    def __init__(self, config: Config, rank: int = 0, event=None):
        self.config = config
        self.rank = rank
        self.world_size = config.tensor_parallel_size
        tp_group = None
        threaded = False
        if config.tensor_parallel_size > 1:
            # Two ways to run the ranks (reference: spawned workers + SHM RPC, engine/llm_engine.py:31-44,
            # model_runner.py:423-458).  Under a launcher (torchrun, one process per GPU, torch.distributed initialised)
            # every process builds the same engine and makes the same calls (SPMD).  From a plain script the ranks are
            # host threads of THIS process, one per GPU (runtime.ThreadedTensorParallel).  Either way the ranks only meet
            # inside the kernel, over NVLink peer memory.
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                if dist.get_world_size() != config.tensor_parallel_size:
                    raise RuntimeError("torch.distributed world size != tensor_parallel_size")
                self.rank = dist.get_rank()
                config.device = self.rank
                tp_group = dist.group.WORLD
            else:
                threaded = True
        hf = config.hf_config
        self.block_size = config.kvcache_block_size
        assert self.block_size == PAGE, "the device KV cache uses 256-token pages"
        max_len = config.max_model_len + 1024              # +1024-token Jacobi buffer (model_runner.py:656-658)
        self.max_pages_per_seq = (max_len + PAGE - 1) // PAGE
        if config.num_kvcache_blocks <= 0:
            config.num_kvcache_blocks = self._size_kv_cache(hf, config)
        max_seqs = min(config.max_num_seqs, 64)
        rows = max(min(config.max_num_batched_tokens, 2048), 64)
        kw = dict(max_rows=rows, max_seqs=max_seqs, num_kv_pages=config.num_kvcache_blocks,
                  max_pages_per_seq=self.max_pages_per_seq, gen_capacity=config.max_model_len + 1024 + 64, max_calls=1024,
                  trace_capacity=1 << 20, tp_shard_attention=config.tp_shard_attention)
        if threaded:
            from ..runtime import ThreadedTensorParallel
            self.rt = ThreadedTensorParallel(hf, self.world_size, devices=config.tp_devices, **kw)
        else:
            self.rt = JacobiRuntime(hf, device=config.device, tp_rank=(self.rank if self.world_size > 1 else 0),
                                    tp_size=self.world_size, tp_group=tp_group, **kw)
        self.rt.free_pages = []                            # pages are owned by the scheduler's BlockManager
        if config.state_dict is not None:
            self.rt.load_state_dict(config.state_dict)
        else:
            load_model(self, config.model)
        self.kv_cache = ("device", config.num_kvcache_blocks)   # opaque: the tensor lives behind the C ABI
        self.block_manager = None                          # set by LLMEngine (llm_engine.py:54)
        self.jacobi_decoder = _DecoderStats()
        self._rng = random.Random(config.draft_seed)
        self.max_seqs = max_seqs

    @staticmethod
    def _size_kv_cache(hf, config: Config) -> int:
        """gpu_memory_utilization of the free memory after weights (model_runner.py:523-622), capped at 40 000 blocks."""
        import torch
        free, total = torch.cuda.mem_get_info(config.device)
        heads = hf.num_attention_heads
        hd = getattr(hf, "head_dim", None) or hf.hidden_size // heads
        per_page = hf.num_hidden_layers * 2 * hf.num_key_value_heads * PAGE * hd * 2
        nparams = 2 * hf.vocab_size * hf.hidden_size + hf.num_hidden_layers * (
            hf.hidden_size * (heads + 2 * hf.num_key_value_heads) * hd + heads * hd * hf.hidden_size +
            3 * hf.hidden_size * hf.intermediate_size)
        budget = free * config.gpu_memory_utilization - 2.2 * nparams - (4 << 30)
        return int(max(16, min(40000, budget // per_page)))

    # ---- RPC-compatible entry (llm_engine.py:88: model_runner.call("run", seqs, is_prefill)) -----------------------
    def call(self, method_name, *args):
        return getattr(self, method_name)(*args)

    def exit(self):
        self.rt.close()

    # ---- helpers --------------------------------------------------------------------------------------------------
    def _handle(self, seq: Sequence, ctx: int) -> SeqHandle:
        return SeqHandle(seq.seq_id, pages=seq.block_table, ctx=ctx, tokens=list(seq.token_ids))

    def _reserve(self, seq: Sequence, n_tokens: int) -> None:
        if n_tokens > self.max_pages_per_seq * PAGE:
            raise RuntimeError(f"sequence of {n_tokens} tokens exceeds max_model_len + Jacobi buffer")
        self.block_manager.ensure(seq, n_tokens)

    # ---- prefill with drafting (model_runner.py:777-963) -------------------------------------------------------------
    def _jacobi_prefill_with_drafting(self, seqs: List[Sequence]) -> List[list]:
        for seq in seqs:
            L = int(seq.jacobi_block_len)
            prompt = list(seq.token_ids)
            draft = [self._rng.choice(prompt) for _ in range(L)]          # :797 random prompt tokens
            self._reserve(seq, len(prompt) + L)
            h = self._handle(seq, seq.num_cached_tokens)                  # > 0: leading pages came from the prefix cache
            first, drafted, _ = self.rt.prefill(h, prompt, draft)         # argmax(logits[P-1 : P+L-1]) (:906-918)
            seq._prefill_draft = list(drafted)
            seq._first_correct = first
            seq.num_cached_tokens = len(prompt)                           # :929-941 tokens truncated, KV trimmed
        return [[] for _ in seqs]

    # ---- decode -------------------------------------------------------------------------------------------------------
    def _fit_new_tokens(self, seqs: List[Sequence], remaining: int, L: int) -> int:
        """Largest per-call token budget <= remaining whose KV reservation (len + budget + 3L + 1 per request) fits the
        free pages.  The reference grows pages per iteration and preempts (scheduler.py:47-60); one device-resident call
        cannot, so a call that does not fit entirely is SHORTENED and the request returns to the scheduler, which
        re-schedules or preempts.  The scheduler has already reserved one block's worth per request."""
        bm = self.block_manager
        free = len(bm.free_block_ids)
        need = lambda new: sum(bm.pages_missing(s, len(s) + new + 3 * L + 1) for s in seqs)
        new = remaining
        while new > L and need(new) > free:
            new = max(L, new - PAGE // 2)
        new = min(new, remaining)
        if need(new) > free:
            raise RuntimeError("out of KV cache blocks: not even one Jacobi block fits "
                               f"({need(new)} pages needed, {free} free)")
        return new

    def _run_jacobi_group(self, seqs: List[Sequence], mode: str) -> List[List[int]]:
        s0 = seqs[0]
        L = int(s0.jacobi_block_len)
        remaining = max(1, int(s0.max_tokens) - s0.num_completion_tokens)
        remaining = self._fit_new_tokens(seqs, remaining, L)
        eos = self.config.eos if self.config.eos is not None and self.config.eos >= 0 else None
        pad = self.config.pad if self.config.pad is not None and self.config.pad >= 0 else None
        V = self.rt.vocab
        handles, fcs, fcts, chs, pads = [], [], [], [], []
        if mode == "engine":
            knobs = PolicyKnobs(mode=MODES["engine"], n=L, eos=eos, pad=pad,
                                max_iteration_count=int(s0.jacobi_max_iterations), max_new_tokens=remaining, vocab=V)
            n_pad = knobs.max_iteration_count * (L - 1) + L
        elif mode == "engine_nongreedy":
            # rejection-sampling verify (jacobi_decoding_nongreedy.py:300-355): every random draw of the call comes from
            # a counter-based generator keyed by a fresh 64-bit seed out of this runner's seeded RNG
            knobs = PolicyKnobs(mode=MODES["engine_nongreedy"], n=L, eos=eos, pad=pad,
                                max_iteration_count=int(s0.jacobi_max_iterations), max_new_tokens=remaining, vocab=V,
                                temperature=float(s0.temperature), seed=self._rng.getrandbits(64))
            n_pad = min(knobs.max_iteration_count, remaining + 1) * (L - 1) + L
        else:
            n_calls = (remaining + L - 1) // L + 3
            knobs = PolicyKnobs(mode=MODES["multiblock"], n=L, K=int(s0.jacobi_max_blocks),
                                r=float(s0.jacobi_spawn_ratio),
                                lookahead_start_ratio=float(s0.jacobi_lookahead_start_ratio),
                                n_gram_pool_size=int(s0.jacobi_n_gram_pool_size), eos=eos, pad=pad,
                                max_iteration_count=int(s0.jacobi_max_iterations), max_new_tokens=remaining,
                                max_calls=n_calls, vocab=V)
        for seq in seqs:
            S = len(seq)
            self._reserve(seq, S + remaining + 3 * L + 1)
            draft = getattr(seq, "_prefill_draft", None)
            if mode == "engine_nongreedy":
                handles.append(self._handle(seq, S - 1))
                fcs.append(0)                                                 # the non-greedy decoder never reads the
                pads.append([self._rng.randrange(V) for _ in range(n_pad)])  # prefill draft: seed + random (:216-236)
            elif mode == "engine":
                handles.append(self._handle(seq, S - 1))
                fcs.append(0)
                fcts.append((list(draft) + [0] * L)[:L] if draft else
                            [self._rng.randrange(V) for _ in range(L)])      # _init_draft (:142-171)
                pads.append([self._rng.randrange(V) for _ in range(n_pad)])
            else:
                handles.append(self._handle(seq, S))
                if draft:                                                     # first call = prefill_drafted_n_gram
                    fcs.append(int(seq._first_correct))
                    fcts.append(list(draft)[:L])
                else:                                                         # resumed: [first_correct, random...]
                    fcs.append(int(seq._first_correct))
                    fcts.append([int(seq._first_correct)] + [self._rng.choice(seq.token_ids) for _ in range(L - 1)])
                chs.append(mt_state_words(random.Random(self._rng.getrandbits(64))))
            seq._prefill_draft = None
        results, run = self.rt.jacobi_run(handles, knobs, fcs, (fcts or None), rng_states=(chs or None),
                                          pad_stream=(pads or None), want_trace=True)
        out = []
        st = self.jacobi_decoder.stats
        per_seq_ctx: Dict[int, List[int]] = {}
        for rec in run.get("trace", []):
            per_seq_ctx.setdefault(rec["seq"], []).append(rec["kv_before"])
        for i, (seq, res) in enumerate(zip(seqs, results)):
            seq.extend_tokens(res.tokens)
            seq.num_cached_tokens = len(seq)                                  # invariant at jacobi_decoding.py:404-407
            seq._kv_ctx = res.ctx
            seq._first_correct = res.next_token
            out.append(list(res.tokens))
            st["num_chunk_calls"] += 1
            st["num_jacobi_iterations"] += res.total_iterations
            st["tokens_accepted"] += len(res.tokens)
            st["tokens_per_call"].append(len(res.tokens))
            st["iterations_per_call"].append(res.total_iterations)
            ctxs = per_seq_ctx.get(i, []) + [res.ctx]
            st["tokens_per_iteration"].extend(max(0, b - a) for a, b in zip(ctxs, ctxs[1:]))
        return out

    def _run_jacobi(self, seqs: List[Sequence], mode: str) -> List[List[int]]:
        groups: Dict[tuple, List[int]] = {}
        for i, s in enumerate(seqs):                                          # group-by-L (jacobi_decoding.py:500-515)
            key = (int(s.jacobi_block_len), int(s.jacobi_max_iterations), int(s.max_tokens) - s.num_completion_tokens,
                   int(s.jacobi_max_blocks), float(s.jacobi_spawn_ratio), int(s.jacobi_n_gram_pool_size),
                   float(s.jacobi_lookahead_start_ratio), float(s.temperature))
            groups.setdefault(key, []).append(i)
        out: List[Optional[List[int]]] = [None] * len(seqs)
        for idxs in groups.values():
            for j in range(0, len(idxs), self.max_seqs):
                chunk = idxs[j:j + self.max_seqs]
                res = self._run_jacobi_group([seqs[i] for i in chunk], mode)
                for i, r in zip(chunk, res):
                    out[i] = r
        return out

    def _run_autoregressive(self, seqs: List[Sequence], is_prefill: bool) -> List[int]:
        if any(s.temperature > 0 for s in seqs):
            raise NotImplementedError("autoregressive sampling (temperature > 0): use decode_strategy='jacobi', whose "
                                      "non-greedy verify samples from the same distribution")
        tokens = []
        if is_prefill:
            for seq in seqs:
                prompt = list(seq.token_ids)
                self._reserve(seq, len(prompt) + 1)
                first, _, _ = self.rt.prefill(self._handle(seq, seq.num_cached_tokens), prompt, [])
                seq.num_cached_tokens = len(prompt)
                tokens.append(int(first))
            return tokens
        toks, groups, handles = [], [], []
        for i, seq in enumerate(seqs):
            self._reserve(seq, len(seq) + 1)
            toks.append(seq.token_ids[-1])
            groups.append((i, 1, i, len(seq) - 1))
            handles.append(self._handle(seq, len(seq) - 1))
        for j in range(0, len(seqs), self.max_seqs):
            hs = handles[j:j + self.max_seqs]
            n = len(hs)
            g = [(k, 1, k, groups[j + k][3]) for k in range(n)]
            commits = [(k, k, groups[j + k][3]) for k in range(n)]
            res = self.rt.forward(toks[j:j + n], g, hs, 0, n, commits)
            tokens.extend(int(t) for t in res["greedy"][:n])
        for seq in seqs:
            seq.num_cached_tokens = len(seq)
        return tokens

    # ---- run dispatch (model_runner.py:1464-1550) ---------------------------------------------------------------------
    def run(self, seqs: List[Sequence], is_prefill: bool):
        mb = [s for s in seqs if s.decode_strategy == "jacobi_multiblock_rejection_recycling"]
        if mb and not self.config.jacobi_multiblock_enabled:
            raise NotImplementedError(
                "Decode strategy 'jacobi_multiblock_rejection_recycling' is not implemented. "
                "Supported strategies: 'autoregressive', 'jacobi'")
        jac = [s for s in seqs if s.decode_strategy == "jacobi"]
        ar = [s for s in seqs if s.decode_strategy == "autoregressive"]
        if len({s.decode_strategy for s in seqs}) > 1:
            raise NotImplementedError(
                f"Mixed decode strategies in same batch not supported. Got {len(jac)} Jacobi, {len(mb)} multiblock "
                f"and {len(ar)} autoregressive sequences.")
        if jac or mb:
            if any(getattr(s, "jacobi_on_policy", False) for s in seqs):
                raise NotImplementedError("on-policy rollout records (jacobi_decoding_nongreedy_on_policy.py) are not "
                                          "part of this build")
            temps = [float(s.temperature) for s in seqs]
            nongreedy = all(t > 0.0 for t in temps)
            if not nongreedy and any(t > 0.0 for t in temps):                 # model_runner.py:331-373
                raise NotImplementedError("Mixed greedy / non-greedy temperatures in one batch are not supported")
            if nongreedy and mb:
                raise NotImplementedError("the multiblock strategy is greedy only")
            if is_prefill:
                return self._jacobi_prefill_with_drafting(seqs)
            return self._run_jacobi(seqs, "engine_nongreedy" if nongreedy else ("engine" if jac else "multiblock"))
        return self._run_autoregressive(seqs, is_prefill)
