"""The `inference_engine` import alias (jacobiforcing_b200/compat): a script written against the reference
(`from inference_engine import LLM, SamplingParams`, inference_engine/tests/test_jacobi_decoding_greedy.py:22) gets this
engine's classes.  Runs in a subprocess so the alias does not leak into the other tests' module table."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
from inference_engine import LLM, SamplingParams
import jacobiforcing_b200 as jf
from inference_engine.sampling_params import SamplingParams as SP2
from inference_engine.config import Config
assert SamplingParams is jf.SamplingParams and SP2 is SamplingParams
assert LLM is jf.LLM and Config is jf.Config
sp = SamplingParams(temperature=0.0, max_tokens=64, decode_strategy="jacobi", jacobi_block_len=32, jacobi_max_iterations=128)
assert sp.decode_strategy == "jacobi" and sp.jacobi_block_len == 32
import torch
if not torch.cuda.is_available():
    try:
        LLM("/nonexistent-model-dir")
    except Exception as e:                      # no CPU fallback: construction fails loudly without a GPU / model
        print("raised", type(e).__name__)
    else:
        raise SystemExit("LLM() must not succeed without a GPU")
print("ok")
"""


def test_inference_engine_alias_resolves_to_this_engine():
    out = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, os.path.join(ROOT, "jacobiforcing_b200", "compat")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("ok")
