"""JacobiRuntime — Python host of the C-ABI library: the per-rank pieces the reference's ModelRunner owns
(inference_engine/engine/model_runner.py): weights (HF tensor names, loader.py:27-47), the paged KV cache with its
block manager (engine/block_manager.py, pages of 256 tokens), prefill-with-drafting (:777-963; HF: unified.py:175-225)
and the Jacobi decode operator, here ONE device-resident loop per call instead of a Python loop of forwards.

numpy carries host buffers across the ABI; torch tensors are accepted for weights (host or device pointers).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _cabi
from .policy import MODES, POLICY_ERRORS, STOP_REASONS, PolicyKnobs, draw_prefill_draft, mt_state_words
from .tp import LocalTPGroup, check_divisibility, shard_tensor

PAGE = 256


def _i32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class SeqHandle:
    """Host bookkeeping of one sequence: its pages and how many tokens are committed (Sequence.block_table /
    num_cached_tokens in the reference, engine/sequence.py:44-56)."""
    seq_id: int
    pages: List[int] = field(default_factory=list)
    ctx: int = 0
    tokens: List[int] = field(default_factory=list)


@dataclass
class JacobiResult:
    tokens: List[int]
    calls: int
    stop_reason: str
    total_new_tokens: int
    total_iterations: int
    iters: List[int]
    ctx: int
    next_token: int = -1          # first_correct_token a following call starts from (HF convention)


class JacobiRuntime:
    def __init__(self, hf_config, *, device: int = 0, max_rows: int = 1024, max_seqs: int = 8,
                 num_kv_pages: int = 64, max_pages_per_seq: int = 8, attn_chunk: int = 64, attn_slots: int = 0,
                 gen_capacity: int = 4096, max_calls: int = 256, trace_capacity: int = 0,
                 tp_rank: int = 0, tp_size: int = 1, tp_group=None, tp_shard_attention: Optional[bool] = None,
                 num_ctas: int = 0):
        """tp_group: a torch.distributed group (one process per GPU) or a tp.LocalTPGroup (one thread per rank in this
        process).  tp_shard_attention: True = heads sharded too (the reference's layout; q and kv heads must divide tp);
        False = MLP-only tensor parallelism (attention replicated, ONE exchange per block, any tp in {2,4,8});
        None = shard when the heads divide (measured faster at every batch size), else replicate."""
        self.lib = _cabi.load()
        g = (lambda k, d=None: hf_config[k] if isinstance(hf_config, dict) else getattr(hf_config, k, d))
        heads = int(g("num_attention_heads"))
        hidden = int(g("hidden_size"))
        head_dim = g("head_dim", None) or hidden // heads
        theta = g("rope_theta", None)
        if theta is None:
            rp = g("rope_parameters", None) or {}
            theta = rp.get("rope_theta", 1e6)
        self.model_cfg = _cabi.ModelCfg(
            hidden, int(g("num_hidden_layers")), heads, int(g("num_key_value_heads")), int(head_dim),
            int(g("intermediate_size")), int(g("vocab_size")), int(bool(g("tie_word_embeddings", False))),
            int(g("max_position_embeddings")), float(g("rms_norm_eps", 1e-6)), float(theta))
        if tp_shard_attention is None:
            tp_shard_attention = heads % tp_size == 0 and int(g("num_key_value_heads")) % tp_size == 0
        self.tp_shard_attention = bool(tp_shard_attention) and tp_size > 1
        check_divisibility(hf_config, tp_size, self.tp_shard_attention)
        self.tp_rank, self.tp_size = tp_rank, tp_size
        self.runtime_cfg = _cabi.RuntimeCfg(device, max_rows, max_seqs, num_kv_pages, max_pages_per_seq, attn_chunk,
                                            attn_slots, gen_capacity, max_calls, trace_capacity, tp_rank, tp_size,
                                            tp_size if self.tp_shard_attention else 1, num_ctas)
        self.vocab = int(g("vocab_size"))
        self.vocab_local = self.vocab // tp_size
        self.hidden = hidden
        self.max_rows = (max(max_rows, 16) + 15) & ~15
        self.max_seqs = max_seqs
        self.max_pages_per_seq = max_pages_per_seq
        self.max_calls = max(max_calls, 4)
        self.trace_capacity = trace_capacity
        h = C.c_void_p()
        _cabi.check(self.lib.jf_create(C.byref(self.model_cfg), C.byref(self.runtime_cfg), C.byref(h)), None)
        self.h = h
        self.free_pages = list(range(num_kv_pages - 1, -1, -1))
        self._next_seq = 0
        self.stats = {"forwards": 0, "device_ms": 0.0, "rows": 0}
        self._tp_group = tp_group
        if tp_size > 1:
            self._connect_peers(tp_group)

    def _tp_sync(self) -> None:
        """Host-side rendezvous before every tensor-parallel launch: the ranks are independent processes and only meet
        inside the kernel, whose cross-GPU waits are bounded (a rank that is seconds late - GC, tokenizer, page faults -
        would otherwise trip its peers' watchdog)."""
        if self.tp_size > 1:
            if isinstance(self._tp_group, LocalTPGroup):
                self._tp_group.barrier()
                return
            import torch.distributed as dist
            if dist.is_initialized():
                dist.barrier(group=self._tp_group)

    def _connect_peers(self, group) -> None:
        """Exchange the CUDA IPC handles of the activation arenas (torch.distributed is only the transport) so the
        kernel can write partial sums / finished rows into the peers' memory and read their argmax keys over NVLink."""
        if isinstance(group, LocalTPGroup):
            group.handles[self.tp_rank] = self.h.value
            group.barrier()
            arr = (C.c_void_p * self.tp_size)(*group.handles)
            _cabi.check(self.lib.jf_tp_attach_local(self.h, arr), self.h)
            group.barrier()
            return
        import torch.distributed as dist
        mine = (C.c_ubyte * 64)()
        _cabi.check(self.lib.jf_tp_export(self.h, mine), self.h)
        gathered = [None] * self.tp_size
        dist.all_gather_object(gathered, bytes(mine), group=group)
        blob = (C.c_ubyte * (64 * self.tp_size))(*b"".join(gathered))
        _cabi.check(self.lib.jf_tp_import(self.h, blob), self.h)
        dist.barrier(group=group)

    # ---- lifecycle -----------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.jf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.jf_launch_count(self.h))

    def device_info(self) -> dict:
        sm = C.c_int32()
        bw, bkv, bact = C.c_int64(), C.c_int64(), C.c_int64()
        _cabi.check(self.lib.jf_device_info(self.h, C.byref(sm), C.byref(bw), C.byref(bkv), C.byref(bact)), self.h)
        return {"sm_count": sm.value, "bytes_weights": bw.value, "bytes_kv": bkv.value, "bytes_act": bact.value}

    # ---- stage profile (ProfileTimer equivalent, model_runner.py:26-152) -------------------------------------------
    def profile_enable(self, forward_index: int) -> None:
        _cabi.check(self.lib.jf_profile_enable(self.h, forward_index), self.h)

    def profile_read(self) -> dict:
        """Per grid barrier of the profiled forward: microseconds CTA 0 spent working in the stage that ends at the
        barrier and waiting at it; stage names follow the kernel's static schedule."""
        stamps = np.zeros(4096, np.int64)
        n, khz = C.c_int32(0), C.c_int32(0)
        _cabi.check(self.lib.jf_profile_read(self.h, stamps.ctypes.data_as(C.POINTER(C.c_int64)), 4096, C.byref(n),
                                             C.byref(khz)), self.h)
        nrec = n.value // 3
        # the opening and the closing record carry -(wall clock ns) in their id slot: measured SM clock of the profiled
        # forward (the attribute clock rate is the maximum; under load the SMs run slower)
        khz_eff = float(khz.value)
        if nrec >= 2 and stamps[0] < 0 and stamps[3 * (nrec - 1)] < 0:
            dt_ns = float(-stamps[3 * (nrec - 1)] - (-stamps[0]))
            dclk = float(stamps[3 * (nrec - 1) + 1] - stamps[1])
            if dt_ns > 0 and dclk > 0:
                khz_eff = dclk / dt_ns * 1e6
            nrec -= 1                                       # the closing record is not a stage
        st = stamps[:n.value].astype(np.float64) / (khz_eff / 1e3)       # microseconds
        kinds = {1: "embed+norm", 2: "qkv_gemm", 3: "qkv_rope", 4: "attention", 5: "attn_combine", 6: "o_gemm",
                 7: "o_reduce_norm", 8: "gate_up_gemm", 9: "down_gemm", 10: "down_reduce_norm", 11: "lm_head_argmax",
                 12: "policy", 13: "o_gemm+norm", 14: "down_gemm+norm", 15: "qkv_gemm+rope", 16: "gate_up+down_gemm+norm"}
        rows = []
        layer, seen_qkv = -1, False
        for i in range(1, nrec):                            # triples (stage id, arrival of CTA 0, release)
            sid = int(stamps[3 * i])
            arrive, release = st[3 * i + 1], st[3 * i + 2]
            prev_release = st[3 * i - 1]
            if sid in (2, 15):
                layer += 1
            name = kinds.get(sid, f"#{sid}")
            if 2 <= sid <= 10 or sid >= 13:
                name = f"L{layer}.{name}"
            rows.append({"stage": name, "work_us": arrive - prev_release, "wait_us": release - arrive})
        return {"sm_clock_khz": khz.value, "sm_clock_khz_measured": khz_eff, "stages": rows}

    # ---- weights (loader.py:27-47 loop stays with the caller; this is the per-tensor weight_loader) ---------------
    def load_tensor(self, name: str, t) -> None:
        import torch
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        t = shard_tensor(name, t.detach(), self.tp_rank, self.tp_size, self.tp_shard_attention).contiguous()
        if t.is_cuda:
            # the library copies on its own stream: the producer kernels on torch's stream must have finished
            torch.cuda.current_stream(t.device).synchronize()
        if t.dtype == torch.bfloat16:
            dt = _cabi.DTYPE_BF16
        elif t.dtype == torch.float16:
            dt = _cabi.DTYPE_F16
        else:
            t = t.float().contiguous()
            dt = _cabi.DTYPE_F32
        shape = (C.c_int64 * t.dim())(*t.shape)
        _cabi.check(self.lib.jf_load_tensor(self.h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim(), dt), self.h)

    def load_state_dict(self, sd: Dict[str, "object"]) -> None:
        for name, t in sd.items():
            if name.endswith("rotary_emb.inv_freq"):
                continue
            self.load_tensor(name, t)
        _cabi.check(self.lib.jf_finalize_weights(self.h), self.h)

    def finalize(self) -> None:
        _cabi.check(self.lib.jf_finalize_weights(self.h), self.h)

    def set_rope_tables(self, cos, sin) -> None:
        cos = np.ascontiguousarray(np.asarray(cos, dtype=np.float32))
        sin = np.ascontiguousarray(np.asarray(sin, dtype=np.float32))
        _cabi.check(self.lib.jf_set_rope_tables(self.h, _ptr(cos), _ptr(sin), cos.shape[0]), self.h)

    # ---- block manager (engine/block_manager.py: allocate / may_append / deallocate; no prefix cache on this path) --
    def new_sequence(self) -> SeqHandle:
        s = SeqHandle(self._next_seq)
        self._next_seq += 1
        return s

    def ensure_capacity(self, seq: SeqHandle, n_tokens: int) -> None:
        need = (n_tokens + PAGE - 1) // PAGE
        if need > self.max_pages_per_seq:
            raise RuntimeError(f"sequence needs {need} KV pages, max_pages_per_seq={self.max_pages_per_seq}")
        while len(seq.pages) < need:
            if not self.free_pages:
                raise RuntimeError("out of KV pages")
            seq.pages.append(self.free_pages.pop())

    def release(self, seq: SeqHandle) -> None:
        self.free_pages.extend(reversed(seq.pages))
        seq.pages = []
        seq.ctx = 0

    def _block_tables(self, seqs: Sequence[SeqHandle]) -> np.ndarray:
        bt = np.zeros((len(seqs), self.max_pages_per_seq), np.int32)
        for i, s in enumerate(seqs):
            bt[i, :len(s.pages)] = s.pages
        return bt

    # ---- one planned forward (the forward_step_batch operator) ------------------------------------------------------
    def forward(self, tokens, groups, seqs: Sequence[SeqHandle], lm_row0: int, lm_rows: int, commits=(),
                want_logits: bool = False, want_hidden: bool = False) -> dict:
        tokens = _i32(tokens)
        T = int(tokens.shape[0])
        groups = _i32(groups).reshape(-1, 4)
        commits = _i32(commits).reshape(-1, 3)
        bt = self._block_tables(seqs)
        greedy = np.zeros(T, np.int32)
        logits = np.zeros((lm_rows, self.vocab_local), np.float32) if want_logits else None   # this rank's vocabulary slice
        hidden = np.zeros((T, self.hidden), np.uint16) if want_hidden else None
        ms = C.c_float()
        self._tp_sync()
        _cabi.check(self.lib.jf_forward(self.h, T, _ptr(tokens), groups.shape[0], _ptr(groups), len(seqs), _ptr(bt),
                                        lm_row0, lm_rows, commits.shape[0], _ptr(commits), _ptr(greedy), _ptr(logits),
                                        _ptr(hidden), C.byref(ms)), self.h)
        self.stats["forwards"] += 1
        self.stats["device_ms"] += ms.value
        self.stats["rows"] += T
        return {"greedy": greedy, "logits": logits, "hidden": hidden, "ms": ms.value}

    def commit_rows(self, seqs: Sequence[SeqHandle], commits) -> None:
        """Append in-flight K/V rows of the last forward to the cache: commits = [(src_row, seq_index, pos)]."""
        commits = _i32(commits).reshape(-1, 3)
        if commits.shape[0] == 0:
            return
        bt = self._block_tables(seqs)
        self._tp_sync()
        _cabi.check(self.lib.jf_commit_rows(self.h, len(seqs), _ptr(bt), commits.shape[0], _ptr(commits)), self.h)

    # ---- prefill-with-drafting (unified.py:175-225 / model_runner.py:777-963) ---------------------------------------
    def prefill(self, seq: SeqHandle, prompt: Sequence[int], draft: Sequence[int]):
        """Forward prompt ++ draft, commit the prompt's KV, return (first_correct_token, prefill_drafted_n_gram) =
        argmax(logits[-n-1:-1]).  Long prompts are processed in chunks of max_rows rows; logits are computed for the
        last n+1 rows only.  seq.ctx > 0: the first seq.ctx prompt tokens are already in the sequence's pages (prefix
        cache hit, engine/block_manager.py) and are skipped - all but the last prompt token, whose logits are needed."""
        prompt = [int(t) for t in prompt]
        draft = [int(t) for t in draft]
        P, n = len(prompt), len(draft)
        if P < 1:
            raise ValueError("empty prompt")
        if seq.ctx < 0 or seq.ctx > P:
            raise ValueError("cached prefix longer than the prompt")
        row = prompt + draft
        total = len(row)
        self.ensure_capacity(seq, total)
        want_from = P - 1                      # first row whose argmax is needed
        out: List[int] = []
        start = min(seq.ctx, P - 1)
        ms = 0.0
        while start < total:
            end = min(total, start + self.max_rows)
            if end < total and total - end < n + 1:   # keep the n+1 logit rows inside one chunk
                end = max(start + 1, total - (n + 1))
            toks = row[start:end]
            T = end - start
            commits = [(i, 0, start + i) for i in range(T) if start + i < P]
            lm0 = max(0, want_from - start)
            lm_row0, lm_rows = 0, 0
            if end > want_from:
                lm_row0 = (lm0 // 8) * 8
                lm_rows = T - lm_row0
            res = self.forward(toks, [(0, T, 0, start)], [seq], lm_row0, lm_rows, commits)
            ms += res["ms"]
            self.stats["prefill_rows"] = self.stats.get("prefill_rows", 0) + T
            if end > want_from:
                out.extend(int(t) for t in res["greedy"][lm0:T])
            start = end
        seq.ctx = P
        seq.tokens = list(prompt)
        drafted = out[:n] if n > 0 else out[:1]
        return drafted[0], drafted, ms

    # ---- the device-resident Jacobi loop -----------------------------------------------------------------------------
    def jacobi_run(self, seqs: Sequence[SeqHandle], knobs: PolicyKnobs, first_correct: Sequence[int],
                   first_call_tokens: Optional[Sequence[Sequence[int]]], rng_states=None, pad_stream=None,
                   max_forwards: int = 0, want_trace: bool = False):
        knobs.validate()
        B = len(seqs)
        n = knobs.n
        kc = _cabi.PolicyKnobsC(*knobs.as_ints())
        prompts = _i32(np.concatenate([_i32(s.tokens) for s in seqs]))
        plens = _i32([len(s.tokens) for s in seqs])
        ctxs = _i32([s.ctx for s in seqs])
        reserve = knobs.max_new_tokens + 3 * n + 1
        for s in seqs:
            self.ensure_capacity(s, len(s.tokens) + reserve)
        bt = self._block_tables(seqs)
        fc = _i32(first_correct) if first_correct is not None else None
        fct = _i32(first_call_tokens).reshape(B, n) if first_call_tokens is not None else None
        ch, ch_rows = None, 0
        if rng_states is not None:                       # [B][625] MT19937 words (mt_state_words)
            ch = np.ascontiguousarray(np.asarray(rng_states, dtype=np.uint32)).reshape(B, 625)
            ch_rows = 625
        pad, pad_len = None, 0
        if pad_stream is not None:
            pad = _i32(pad_stream).reshape(B, -1)
            pad_len = pad.shape[1]
        out_cap = knobs.max_new_tokens + 3 * n + 1
        out = np.zeros((B, out_cap), np.int32)
        summ = np.zeros((B, 8), np.int32)
        iters = np.zeros((B, self.max_calls), np.int32)
        stats = np.zeros(4, np.float32)
        trace = np.zeros(max(self.trace_capacity, 1), np.int32) if want_trace else None
        tl = C.c_int32(0)
        self._tp_sync()
        rc = self.lib.jf_jacobi_run(self.h, B, C.byref(kc), _ptr(prompts), _ptr(plens), _ptr(ctxs), _ptr(fc), _ptr(fct),
                                    _ptr(ch), ch_rows, _ptr(pad), pad_len, _ptr(bt), max_forwards, _ptr(out), out_cap,
                                    _ptr(summ), _ptr(iters), _ptr(stats), _ptr(trace), C.byref(tl))
        if rc == _cabi.JF_ERR_POLICY:
            code = int(summ[:, 4].max())
            msg = POLICY_ERRORS.get(code, f"policy error {code}")
            raise (ValueError if code == 2 else RuntimeError)(msg)
        _cabi.check(rc, self.h)
        results = []
        for i, s in enumerate(seqs):
            ngen, calls, stop, total_new, _err, total_iters, ctx, nxt = [int(x) for x in summ[i]]
            toks = out[i, :ngen].tolist()
            s.tokens = s.tokens + toks
            s.ctx = ctx
            results.append(JacobiResult(toks, calls, STOP_REASONS.get(stop, str(stop)), total_new, total_iters,
                                        iters[i, :calls].tolist(), ctx, nxt))
        self.stats["forwards"] += int(stats[0])
        self.stats["device_ms"] += float(stats[1])
        self.stats["rows"] += int(stats[2])
        run = {"forwards": int(stats[0]), "device_ms": float(stats[1]), "rows": int(stats[2])}
        if want_trace:
            run["trace"] = parse_trace(trace[:tl.value])
        return results, run


def parse_trace(words: np.ndarray) -> List[dict]:
    out, i = [], 0
    w = words.tolist()
    while i + 5 <= len(w):
        it, s, ctx, B, L = w[i:i + 5]
        rows = [w[i + 5 + r * L: i + 5 + (r + 1) * L] for r in range(B)]
        out.append({"iter": it, "seq": s, "kv_before": ctx, "rows": rows})
        i += 5 + B * L
    return out


def hf_generate(rt: JacobiRuntime, prompts: Sequence[Sequence[int]], *, n: int, max_new_tokens: int, mode: str = "multiblock",
                K: int = 2, r: float = 0.85, lookahead_start_ratio: float = 0.0, n_gram_pool_size: int = 4,
                eos: Optional[int] = None, pad: Optional[int] = None, max_iteration_count: int = 128,
                max_calls: int = 1024, rngs=None, want_trace: bool = False):
    """The reference driver protocol (JacobiForcing/jacobi_forcing_inference_MR_humaneval.py:152-240) for a batch of
    independent sequences: prefill with a random draft, then ONE device-resident run that chains the per-block calls.
    `rngs[i]` is the random.Random the i-th sequence draws its draft tokens from."""
    import random
    B = len(prompts)
    rngs = rngs or [random.Random(2 + i) for i in range(B)]
    seqs, fcs, fcts, chs = [], [], [], []
    prefill_ms = 0.0
    for p, rng in zip(prompts, rngs):
        s = rt.new_sequence()
        draft = draw_prefill_draft(rng, p, n)
        fc, drafted, ms = rt.prefill(s, p, draft)
        prefill_ms += ms
        seqs.append(s); fcs.append(fc); fcts.append(drafted)
        chs.append(mt_state_words(rng))      # the device continues this generator for the later calls' drafts
    knobs = PolicyKnobs(mode=MODES[mode], n=n, K=K, r=r, lookahead_start_ratio=lookahead_start_ratio,
                        n_gram_pool_size=n_gram_pool_size, eos=eos, pad=pad, max_iteration_count=max_iteration_count,
                        max_new_tokens=max_new_tokens, max_calls=min(max_calls, rt.max_calls), vocab=rt.vocab)
    results, run = rt.jacobi_run(seqs, knobs, fcs, fcts, rng_states=chs, want_trace=want_trace)
    run["prefill_ms"] = prefill_ms
    return seqs, results, run


class ThreadedTensorParallel:
    """tensor_parallel_size ranks inside ONE process: one JacobiRuntime per rank, each driven by its own host thread (the
    launches must overlap: the ranks meet inside the kernels), wired with tp.LocalTPGroup / jf_tp_attach_local.  What the
    reference gets by spawning worker processes behind `LLM(..., tensor_parallel_size=N)` (engine/llm_engine.py:31-44,
    model_runner.py:423-458) - a plain script, no torchrun.  `devices[r]` is rank r's CUDA device; ranks may share a device
    (its SMs are then divided between them), which is how the single-GPU tests cover this class.

    The object has the JacobiRuntime surface the engine uses; every call is collective (fanned out to all ranks) and
    returns rank 0's result.  Host bookkeeping objects (SeqHandle) passed in are mutated by rank 0 only: the other ranks
    work on copies."""

    def __init__(self, hf_config, tp_size: int, devices: Optional[Sequence[int]] = None, **kw):
        import copy
        from concurrent.futures import ThreadPoolExecutor
        import torch
        self._copy = copy
        self.tp_size = tp_size
        _cabi.load()                                             # once, before the rank threads race for it
        ndev = torch.cuda.device_count()
        devices = list(devices) if devices is not None else [r % max(1, ndev) for r in range(tp_size)]
        sharing = max(devices.count(d) for d in set(devices))
        if sharing > 1 and not kw.get("num_ctas"):
            sms = torch.cuda.get_device_properties(devices[0]).multi_processor_count
            kw["num_ctas"] = (sms - 4) // sharing                 # co-resident persistent kernels
        self.group = LocalTPGroup(tp_size)
        self._pool = [ThreadPoolExecutor(max_workers=1, thread_name_prefix=f"jf-tp{r}") for r in range(tp_size)]
        futs = [self._pool[r].submit(JacobiRuntime, hf_config, device=devices[r], tp_rank=r, tp_size=tp_size,
                                     tp_group=self.group, **kw) for r in range(tp_size)]
        self.ranks: List[JacobiRuntime] = [f.result() for f in futs]
        r0 = self.ranks[0]
        self.vocab, self.hidden, self.max_rows, self.max_seqs = r0.vocab, r0.hidden, r0.max_rows, r0.max_seqs
        self.max_pages_per_seq, self.max_calls, self.trace_capacity = r0.max_pages_per_seq, r0.max_calls, r0.trace_capacity
        self.tp_shard_attention = r0.tp_shard_attention

    def _all(self, method: str, *args, **kwargs):
        def call(r):
            a, k = (args, kwargs) if r == 0 else self._copy.deepcopy((args, kwargs))
            return getattr(self.ranks[r], method)(*a, **k)
        futs = [self._pool[r].submit(call, r) for r in range(self.tp_size)]
        errs, out = [], None
        for r, f in enumerate(futs):
            try:
                res = f.result()
                if r == 0:
                    out = res
            except BaseException as e:       # noqa: BLE001 - collect, keep joining: a dead rank must not strand its peers
                errs.append(e)
                try:
                    self.group._barrier.abort()
                except Exception:
                    pass
        if errs:
            raise errs[0]
        return out

    # ---- the JacobiRuntime surface -------------------------------------------------------------------------------
    def load_tensor(self, name, t):
        def call(r):
            return self.ranks[r].load_tensor(name, t)            # every rank narrows the same tensor to its shard
        for f in [self._pool[r].submit(call, r) for r in range(self.tp_size)]:
            f.result()

    def load_state_dict(self, sd):
        for name, t in sd.items():
            if not name.endswith("rotary_emb.inv_freq"):
                self.load_tensor(name, t)
        self.finalize()

    def finalize(self):
        return self._all("finalize")

    def set_rope_tables(self, cos, sin):
        return self._all("set_rope_tables", cos, sin)

    def prefill(self, seq, prompt, draft):
        return self._all("prefill", seq, prompt, draft)

    def forward(self, *a, **k):
        return self._all("forward", *a, **k)

    def commit_rows(self, *a, **k):
        return self._all("commit_rows", *a, **k)

    def jacobi_run(self, *a, **k):
        return self._all("jacobi_run", *a, **k)

    def new_sequence(self):
        return self.ranks[0].new_sequence()

    def ensure_capacity(self, seq, n_tokens):
        return self.ranks[0].ensure_capacity(seq, n_tokens)

    def release(self, seq):
        return self.ranks[0].release(seq)

    @property
    def free_pages(self):
        return self.ranks[0].free_pages

    @free_pages.setter
    def free_pages(self, v):
        for rt in self.ranks:
            rt.free_pages = list(v)

    @property
    def stats(self):
        return self.ranks[0].stats

    @property
    def launches(self) -> int:
        return sum(rt.launches for rt in self.ranks)

    def device_info(self):
        return self.ranks[0].device_info()

    def close(self):
        for r, rt in enumerate(self.ranks):
            self._pool[r].submit(rt.close).result()
        for p in self._pool:
            p.shutdown(wait=True)
