"""LLMEngine — request queue + step loop (inference_engine/engine/llm_engine.py:17-202).  Rank 0 lives in the caller's
process; there is no worker RPC because a launch of the persistent kernel already covers a whole decode call."""
from __future__ import annotations

import atexit
from dataclasses import fields
from time import perf_counter

from ..config import Config
from ..sampling_params import SamplingParams
from .model_runner import ModelRunner
from .scheduler import Scheduler
from .sequence import Sequence


class LLMEngine:
    def __init__(self, model, tokenizer_path=None, **kwargs):
        config_fields = {f.name for f in fields(Config)}
        config = Config(model, **{k: v for k, v in kwargs.items() if k in config_fields})
        self.config = config
        self._exited = False
        self.model_runner = ModelRunner(config, 0, [])
        self.tokenizer = None
        from ..config import _resolve_checkpoint_dir
        # a training output directory is resolved to its latest checkpoint-N, like config.model
        tok_path = _resolve_checkpoint_dir(tokenizer_path) if tokenizer_path else config.model
        if tok_path is not None:
            try:
                from transformers import AutoTokenizer
                self.tokenizer = AutoTokenizer.from_pretrained(tok_path, use_fast=True)
            except Exception:
                if tokenizer_path is not None:
                    raise
        if self.tokenizer is not None:
            config.eos = self.tokenizer.eos_token_id if self.tokenizer.eos_token_id is not None else -1
            config.pad = (self.tokenizer.pad_token_id if self.tokenizer.pad_token_id is not None
                          else config.eos)
        self.scheduler = Scheduler(config)
        self.model_runner.block_manager = self.scheduler.block_manager
        self.scheduler.set_kv_cache(self.model_runner.kv_cache)
        atexit.register(self.exit)

    def exit(self):
        if getattr(self, "_exited", False):
            return
        self._exited = True
        self.model_runner.call("exit")

    def add_request(self, prompt, sampling_params: SamplingParams):
        if isinstance(prompt, str):
            if self.tokenizer is None:
                raise ValueError("no tokenizer loaded: pass prompts as lists of token ids")
            prompt = self.tokenizer.encode(prompt)
        self.scheduler.add(Sequence(prompt, sampling_params))

    def step(self):
        seqs, is_prefill = self.scheduler.schedule()
        result = self.model_runner.call("run", seqs, is_prefill)
        if result is None or (is_prefill and len(result) > 0 and result[0] == []):
            return [], sum(len(seq) for seq in seqs)
        token_ids = result
        is_jacobi = len(token_ids) > 0 and isinstance(token_ids[0], (list, tuple))
        if is_jacobi:
            self.scheduler.postprocess_jacobi(seqs, token_ids)
            num_new = sum(len(t) for t in token_ids)
        else:
            self.scheduler.postprocess(seqs, token_ids)
            num_new = len(seqs)
        outputs = [(seq.seq_id, seq.completion_token_ids) for seq in seqs if seq.is_finished]
        return outputs, (sum(len(seq) for seq in seqs) if is_prefill else -num_new)

    def is_finished(self):
        return self.scheduler.is_finished()

    def generate(self, prompts, sampling_params, use_tqdm: bool = True):
        pbar = None
        if use_tqdm:
            from tqdm.auto import tqdm
            pbar = tqdm(total=len(prompts), desc="Generating", dynamic_ncols=True)
        if not isinstance(sampling_params, list):
            sampling_params = [sampling_params] * len(prompts)
        for prompt, sp in zip(prompts, sampling_params):
            self.add_request(prompt, sp)
        outputs = {}
        prefill_tp = decode_tp = 0.0
        while not self.is_finished():
            t = perf_counter()
            output, num_tokens = self.step()
            if pbar is not None:
                dt = max(perf_counter() - t, 1e-9)
                if num_tokens > 0:
                    prefill_tp = num_tokens / dt
                else:
                    decode_tp = -num_tokens / dt
                pbar.set_postfix({"Prefill": f"{int(prefill_tp)}tok/s", "Decode": f"{int(decode_tp)}tok/s"})
            for seq_id, toks in output:
                outputs[seq_id] = toks
                if pbar is not None:
                    pbar.update(1)
        if pbar is not None:
            pbar.close()
        outs = [outputs[k] for k in sorted(outputs)]
        dec = (lambda ids: self.tokenizer.decode(ids)) if self.tokenizer is not None else (lambda ids: "")
        return [{"text": dec(list(t)), "token_ids": [int(x) for x in t]} for t in outs]
