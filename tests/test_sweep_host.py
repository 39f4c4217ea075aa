"""CPU tests of the sweep CLI (jacobiforcing_b200/sweep.py): grid enumeration, run ids and CSV naming follow the
reference (JacobiForcing/jacobi_forcing_inference_MR_humaneval_config_grid_search.py:53-86, :317-321,
scripts/inference/scanning_hyperparameter_jacobi_decoding_mr.sh:17-25); the runner is stubbed (no GPU here)."""
import csv
import os

from jacobiforcing_b200 import sweep as sw
from jacobiforcing_b200.drivers import CSV_COLUMNS


def test_positionals_and_knobs_parse_like_the_reference():
    a = sw.build_argparser().parse_args(["data.parquet", "/m", "/t", "out/diffusion_profile_humaneval.csv", "64", "512",
                                         "--n_token_seq_len", "64", "32", "--K", "1", "2", "--r", "0.5", "0.85",
                                         "--n_gram_pool_size", "4"])
    assert (a.df_file, a.model_name, a.tokenizer_name, a.max_calls, a.max_new_tokens) == ("data.parquet", "/m", "/t", 64, 512)
    pts = sw.grid(a)
    assert len(pts) == 2 * 2 * 2 * 1 * 1
    assert pts[0] == dict(n=64, K=1, r=0.5, pool=4, lsr=0.0) and pts[-1] == dict(n=32, K=2, r=0.85, pool=4, lsr=0.0)
    assert sw.run_id(32, 2, 0.85, 4) == "ntok32_K2_r0.85_ng4"
    assert sw.csv_path_for("out/diffusion_profile_humaneval.csv", "ntok32_K2_r0.85_ng4") == \
        "out/ntok32_K2_r0.85_ng4_diffusion_profile_humaneval.csv"
    d = sw.build_argparser().parse_args([])                     # the reference's defaults
    assert (d.n_token_seq_len, d.K, d.r, d.n_gram_pool_size, d.lookahead_start_ratio) == ([64], [2], [0.8], [4], [0.0])
    assert (d.max_calls, d.max_new_tokens) == (1024, 1024)


def test_synthetic_prompts_and_sweep_outputs(tmp_path, monkeypatch):
    prompts, tasks = sw.load_prompts("synthetic:3:10", None, 1000)
    assert len(prompts) == 3 and all(len(p) == 10 for p in prompts) and tasks[2] == "synthetic_2"
    calls = []

    def fake_profile_rows(rt, prompts, **kw):
        calls.append(kw)
        return [{"index": i, "task_id": kw["task_ids"][i], "prompt_tokens": len(p), "new_tokens": 32, "calls": 2,
                 "total_iterations": 16 + kw["n"] // 8, "avg_iter_per_call": 8.0, "avg_iter_per_token": 0.5,
                 "time_sec": 0.1, "toks_per_sec": 320.0, "stop_reason": "eos" if i == 0 else "max_new_tokens",
                 "token_ids": [1, 2]} for i, p in enumerate(prompts)]
    monkeypatch.setattr(sw, "profile_rows", fake_profile_rows)
    a = sw.build_argparser().parse_args(["synthetic:3:10", "/m", "/t", str(tmp_path / "diffusion_profile_humaneval.csv"),
                                         "8", "32", "--n_token_seq_len", "16", "8", "--K", "2", "3"])
    summary = sw.sweep(None, prompts, tasks, sw.grid(a), csv_path=a.csv_path, max_calls=a.max_calls,
                       max_new_tokens=a.max_new_tokens, eos=5, pad=6)
    assert [s["status"][:11] for s in summary] == ["ok", "unsupported", "ok", "unsupported"]
    assert len(calls) == 2 and calls[0]["n"] == 16 and calls[0]["mode"] == "multiblock" and calls[0]["eos"] == 5
    ok = summary[0]
    assert ok["prompts"] == 3 and ok["new_tokens"] == 96 and ok["eos_prompts"] == 1
    assert abs(ok["toks_per_sec"] - 320.0) < 1e-6
    with open(ok["csv"]) as f:
        rows = list(csv.DictReader(f))
    assert list(rows[0].keys()) == CSV_COLUMNS and len(rows) == 3
    assert os.path.exists(tmp_path / "sweep_summary.csv")
