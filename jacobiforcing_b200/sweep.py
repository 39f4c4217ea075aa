"""Hyper-parameter sweep CLI over (block size n, K, r, n-gram pool size, lookahead start ratio) — the reference's
JacobiForcing/jacobi_forcing_inference_MR_humaneval_config_grid_search.py:53-77 (one grid point per process) together
with its launcher scripts/inference/scanning_hyperparameter_jacobi_decoding_mr.sh:17-45 (the grid), in ONE process:
the model is loaded once and every grid point runs the device-resident loop on the same prompts.

    python -m jacobiforcing_b200.sweep DF_FILE MODEL_DIR TOKENIZER_DIR CSV_PATH MAX_CALLS MAX_NEW_TOKENS \
        --n_token_seq_len 64 32 16 --K 1 2 --r 0.5 0.85 --n_gram_pool_size 1 4 --lookahead_start_ratio 0.0

Same positionals, same knob names (each now takes one OR several values), same run id (`ntok{n}_K{K}_r{r:.2f}_ng{pool}`),
same per-run CSV (`<run_id>_<csv basename>`, schema of drivers.CSV_COLUMNS = MR_humaneval.py:259-273) plus one
`sweep_summary.csv` with the aggregate row of every grid point.  DF_FILE is a parquet / jsonl file with a `prompt` column
(wrapped in the reference's HumanEval instruction + chat template when the tokenizer has one) or `synthetic:N:P` for N
random prompts of P tokens (there are no datasets offline).  K >= 3 relies on a block-list drift of the reference
(SURVEY finding 5) and is reported as unsupported, not run.
"""
from __future__ import annotations

import argparse
import csv
import itertools
import json
import os
import random
import sys
from typing import Dict, Iterable, List, Optional, Sequence

from .drivers import CSV_COLUMNS, profile_rows

PROMPT_TEMPLATE = ("Please continue to complete the function. You are not allowed to modify the given code and do the "
                   "completion only. Please return all completed function in a codeblock. Here is the given code to do "
                   "completion:\n```\n{}\n```")


def build_argparser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("df_file", nargs="?", default="synthetic:4:64")
    p.add_argument("model_name", nargs="?", default=None)
    p.add_argument("tokenizer_name", nargs="?", default=None)
    p.add_argument("csv_path", nargs="?", default="profiling_results/diffusion_profile_humaneval.csv")
    p.add_argument("max_calls", type=int, nargs="?", default=1024)
    p.add_argument("max_new_tokens", type=int, nargs="?", default=1024)
    p.add_argument("--n_token_seq_len", type=int, nargs="+", default=[64])
    p.add_argument("--K", type=int, nargs="+", default=[2])
    p.add_argument("--r", type=float, nargs="+", default=[0.8])
    p.add_argument("--n_gram_pool_size", type=int, nargs="+", default=[4])
    p.add_argument("--lookahead_start_ratio", type=float, nargs="+", default=[0.0])
    p.add_argument("--out_prefix", default=None)
    p.add_argument("--limit", type=int, default=0, help="use only the first N prompts")
    p.add_argument("--seed", type=int, default=0, help="seeds the random draft tokens (the reference's are unseeded)")
    p.add_argument("--device", type=int, default=0)
    return p


def run_id(n: int, K: int, r: float, pool: int, prefix: Optional[str] = None) -> str:
    """grid_search.py:86"""
    return prefix or f"ntok{n}_K{K}_r{r:.2f}_ng{pool}"


def grid(args) -> List[Dict]:
    """Every combination, in the launcher's loop order (block size, K, r, pool; scanning_...mr.sh:17-25)."""
    return [dict(n=n, K=K, r=r, pool=pool, lsr=lsr)
            for n, K, r, pool, lsr in itertools.product(args.n_token_seq_len, args.K, args.r, args.n_gram_pool_size,
                                                        args.lookahead_start_ratio)]


def csv_path_for(csv_path: str, rid: str) -> str:
    """grid_search.py:317-321: a generic file name gets the run id prepended; with several grid points always."""
    base_dir, base = os.path.dirname(csv_path), os.path.basename(csv_path)
    return os.path.join(base_dir, f"{rid}_{base}")


def load_prompts(df_file: str, tokenizer, vocab: int, limit: int = 0):
    """(prompt token ids, task ids)."""
    if df_file.startswith("synthetic:"):
        _, n, plen = df_file.split(":")
        rng = random.Random(1)
        hi = vocab - 2 if vocab > 4 else vocab
        prompts = [[rng.randrange(hi) for _ in range(int(plen))] for _ in range(int(n))]
        tasks = [f"synthetic_{i}" for i in range(int(n))]
    else:
        if tokenizer is None:
            raise ValueError("a dataset of text prompts needs a tokenizer directory")
        if df_file.endswith(".parquet"):
            import pandas as pd
            records = pd.read_parquet(df_file).to_dict(orient="records")
        else:
            with open(df_file) as f:
                records = [json.loads(line) for line in f if line.strip()]
        prompts, tasks = [], []
        for idx, row in enumerate(records):
            text = PROMPT_TEMPLATE.format(str(row["prompt"]).strip())
            if getattr(tokenizer, "chat_template", None):
                text = tokenizer.apply_chat_template([{"role": "user", "content": text}], tokenize=False,
                                                     add_generation_prompt=True)
            prompts.append([int(t) for t in tokenizer.encode(text)])
            tasks.append(str(row.get("task_id", f"idx_{idx}")))
    if limit > 0:
        prompts, tasks = prompts[:limit], tasks[:limit]
    return prompts, tasks


def summarise(rows: Sequence[Dict]) -> Dict:
    """The aggregate the reference prints per run (grid_search.py:323-353): totals over all prompts and over the prompts
    that ended on EOS."""
    def agg(sel: Iterable[Dict]) -> Dict:
        sel = list(sel)
        toks = sum(r["new_tokens"] for r in sel)
        its = sum(r["total_iterations"] for r in sel)
        gen_s = sum(r["new_tokens"] / r["toks_per_sec"] for r in sel if r["toks_per_sec"] > 0)
        return {"prompts": len(sel), "new_tokens": toks, "total_iterations": its,
                "avg_iter_per_token": its / toks if toks else float("nan"),
                "tokens_per_iteration": toks / its if its else float("nan"),
                "toks_per_sec": toks / gen_s if gen_s > 0 else float("nan")}
    out = agg(rows)
    eos = agg(r for r in rows if r["stop_reason"] == "eos")
    out.update({f"eos_{k}": v for k, v in eos.items()})
    return out


def write_csv(path: str, rows: Sequence[Dict], columns: Sequence[str]) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(columns), extrasaction="ignore")
        w.writeheader()
        for r in rows:
            w.writerow(r)


def sweep(rt, prompts, tasks, points: Sequence[Dict], *, csv_path: str, max_calls: int, max_new_tokens: int,
          eos: Optional[int], pad: Optional[int], seed: int = 0, out_prefix: Optional[str] = None, log=print) -> List[Dict]:
    """Runs every grid point on `rt` (a JacobiRuntime); returns the summary rows (also written next to the CSVs)."""
    summary = []
    for pt in points:
        rid = run_id(pt["n"], pt["K"], pt["r"], pt["pool"], out_prefix if len(points) == 1 else None)
        row = {"run_id": rid, "n_token_seq_len": pt["n"], "K": pt["K"], "r": pt["r"], "n_gram_pool_size": pt["pool"],
               "lookahead_start_ratio": pt["lsr"], "status": "ok"}
        try:
            if pt["K"] >= 3:
                raise NotImplementedError("device multiblock policy supports K in {1, 2}")
            rows = profile_rows(rt, prompts, n=pt["n"], max_new_tokens=max_new_tokens, K=pt["K"], r=pt["r"],
                                n_gram_pool_size=pt["pool"], lookahead_start_ratio=pt["lsr"], eos=eos, pad=pad,
                                mode="multiblock", max_calls=max_calls, task_ids=tasks, seed=seed)
        except NotImplementedError as e:
            row["status"] = f"unsupported: {e}"
            summary.append(row)
            log(f"[RUN] {rid}: {row['status']}")
            continue
        path = csv_path_for(csv_path, rid)
        write_csv(path, rows, CSV_COLUMNS)
        row.update(summarise(rows))
        row["csv"] = path
        summary.append(row)
        log(f"[RUN] {rid}: {row['new_tokens']} tokens, {row['tokens_per_iteration']:.3f} tokens/iteration, "
            f"{row['toks_per_sec']:.1f} tokens/s -> {path}")
    cols = ["run_id", "n_token_seq_len", "K", "r", "n_gram_pool_size", "lookahead_start_ratio", "status", "prompts",
            "new_tokens", "total_iterations", "avg_iter_per_token", "tokens_per_iteration", "toks_per_sec", "eos_prompts",
            "eos_new_tokens", "eos_total_iterations", "eos_avg_iter_per_token", "eos_tokens_per_iteration",
            "eos_toks_per_sec", "csv"]
    write_csv(os.path.join(os.path.dirname(csv_path) or ".", "sweep_summary.csv"), summary, cols)
    return summary


def main(argv: Optional[Sequence[str]] = None) -> int:
    args = build_argparser().parse_args(argv)
    if args.model_name is None:
        print("MODEL_DIR is required (a checkpoint directory with config.json + *.safetensors)", file=sys.stderr)
        return 2
    from transformers import AutoConfig
    from .config import _resolve_checkpoint_dir
    from .engine.model_runner import load_model
    from .runtime import JacobiRuntime
    model_dir = _resolve_checkpoint_dir(args.model_name)
    hf = AutoConfig.from_pretrained(model_dir)
    tokenizer = None
    tok_dir = _resolve_checkpoint_dir(args.tokenizer_name) if args.tokenizer_name else model_dir
    try:
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(tok_dir, use_fast=True)
    except Exception:
        if args.tokenizer_name:
            raise
    eos = tokenizer.eos_token_id if tokenizer is not None else None
    pad = tokenizer.pad_token_id if tokenizer is not None and tokenizer.pad_token_id is not None else (eos if eos is not None else 0)
    prompts, tasks = load_prompts(args.df_file, tokenizer, hf.vocab_size, args.limit)
    points = grid(args)
    n_max = max(p["n"] for p in points)
    longest = max(len(p) for p in prompts)
    pages = (longest + args.max_new_tokens + 4 * n_max + 255) // 256 + 1
    rows_cap = min(4096, max(longest + n_max, 2 * n_max * 16, 64))
    rt = JacobiRuntime(hf, device=args.device, max_rows=rows_cap, max_seqs=1, num_kv_pages=pages, max_pages_per_seq=pages,
                       gen_capacity=longest + args.max_new_tokens + 4 * n_max + 8,
                       max_calls=min(args.max_calls, args.max_new_tokens + 8) + 8)

    class _Runner:          # load_model walks the safetensors files into runner.rt (engine/model_runner.py)
        pass
    runner = _Runner()
    runner.rt = rt
    load_model(runner, model_dir)
    try:
        sweep(rt, prompts, tasks, points, csv_path=args.csv_path, max_calls=args.max_calls,
              max_new_tokens=args.max_new_tokens, eos=eos, pad=pad, seed=args.seed, out_prefix=args.out_prefix)
    finally:
        rt.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
