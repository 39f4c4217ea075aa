"""LLM("<checkpoint dir>") end to end (-m gpu): config.json -> Config / AutoConfig, *.safetensors -> load_model ->
jf_load_tensor per HF tensor name (reference: inference_engine/utils/loader.py:12-47, config.py:31-52), tokenizer files ->
AutoTokenizer -> eos / pad ids, string prompts in, text out.  The checkpoint is synthetic (tests/ckpt_helpers.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ckpt(tmp_path_factory):
    from tests.ckpt_helpers import make_checkpoint_dir
    root = tmp_path_factory.mktemp("ckpt")
    # a training output directory: the engine must pick the latest checkpoint-N (reference config.py:31-41)
    make_checkpoint_dir(str(root / "checkpoint-10"), seed=5)
    model, tok = make_checkpoint_dir(str(root / "checkpoint-200"), seed=0, shards=3)
    return str(root), model, tok


def test_llm_from_checkpoint_dir_matches_hf_generate(ckpt):
    from jacobiforcing_b200 import LLM, SamplingParams
    root, model, tok = ckpt
    llm = LLM(root, max_model_len=2048, num_kvcache_blocks=32, max_num_seqs=4, max_num_batched_tokens=2048, draft_seed=3)
    try:
        assert llm.config.model.endswith("checkpoint-200")
        assert llm.config.eos == tok.eos_token_id and llm.config.pad == tok.pad_token_id
        texts = [" ".join(f"w{(7 * i + 3 * j) % 900}" for j in range(12 + 4 * i)) for i in range(3)]
        sp = dict(temperature=0.0, max_tokens=40, ignore_eos=True)
        ar = llm.generate(texts, SamplingParams(decode_strategy="autoregressive", **sp), use_tqdm=False)
        jc = llm.generate(texts, SamplingParams(decode_strategy="jacobi", jacobi_block_len=8, **sp), use_tqdm=False)
        assert [a["token_ids"][:40] for a in ar] == [j["token_ids"][:40] for j in jc]
        for text, a in zip(texts, ar):
            ids = tok.encode(text)
            assert len(a["token_ids"]) >= 40 and a["text"] == tok.decode(a["token_ids"])
            with torch.no_grad():
                ref = model.generate(torch.tensor([ids]), max_new_tokens=40, do_sample=False,
                                     eos_token_id=None, pad_token_id=tok.pad_token_id)[0, len(ids):].tolist()
            same = next((i for i, (x, y) in enumerate(zip(a["token_ids"], ref)) if x != y), 40)
            if same < 40:          # only a near-tie of the oracle may differ
                with torch.no_grad():
                    lg = model(torch.tensor([ids + ref[:same]])).logits[0, -1]
                t2 = lg.topk(2).values
                assert float(t2[0] - t2[1]) < 3e-2, f"diverged at {same} with margin {float(t2[0] - t2[1])}"
    finally:
        llm.exit()


def test_missing_shard_is_an_error(ckpt, tmp_path):
    """A checkpoint directory that lacks one safetensors shard must not load (per-layer presence check)."""
    import shutil
    from jacobiforcing_b200 import LLM
    root, _, _ = ckpt
    broken = tmp_path / "broken"
    shutil.copytree(os.path.join(root, "checkpoint-200"), broken)
    os.remove(broken / "model-00002-of-00003.safetensors")
    with pytest.raises(ValueError, match="never loaded"):
        LLM(str(broken), max_model_len=2048, num_kvcache_blocks=16, max_num_seqs=2, max_num_batched_tokens=2048)


def test_sweep_cli_on_checkpoint_dir(ckpt, tmp_path):
    """python -m jacobiforcing_b200.sweep over a 2 x 2 grid on the synthetic checkpoint: one CSV per grid point with the
    reference's schema, a summary, and the same ids for every grid point (greedy Jacobi is exact: the knobs only change
    the iteration count)."""
    import csv
    from jacobiforcing_b200 import sweep as sw
    from jacobiforcing_b200.drivers import CSV_COLUMNS
    root, _, _ = ckpt
    out = tmp_path / "prof" / "diffusion_profile_humaneval.csv"
    rc = sw.main(["synthetic:2:24", root, root, str(out), "64", "48", "--n_token_seq_len", "16", "8", "--K", "1", "2",
                  "--r", "0.85", "--n_gram_pool_size", "4"])
    assert rc == 0
    with open(tmp_path / "prof" / "sweep_summary.csv") as f:
        summ = list(csv.DictReader(f))
    assert [s["run_id"] for s in summ] == ["ntok16_K1_r0.85_ng4", "ntok16_K2_r0.85_ng4", "ntok8_K1_r0.85_ng4",
                                           "ntok8_K2_r0.85_ng4"]
    for s in summ:
        assert s["status"] == "ok" and int(s["prompts"]) == 2 and int(s["new_tokens"]) >= 2 * 47
        with open(s["csv"]) as f:
            rows = list(csv.DictReader(f))
        assert list(rows[0].keys()) == CSV_COLUMNS and len(rows) == 2
        assert all(float(r["avg_iter_per_token"]) <= 1.0 + 1e-9 for r in rows)      # TPF >= 1
