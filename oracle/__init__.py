"""CPU ORACLE - test infrastructure only (never imported by jacobiforcing_b200/).

  jacobi_policy.py   pure-Python restatement of the reference's greedy Jacobi state machines
  hf_model.py        seeded random-init Qwen2 checkpoints + an HF-transformers forward with the fwd interface the
                     policy restatement drives (the arithmetic of the path lives in `transformers`, which the
                     reference pins at 4.53.0 and this image ships at 5.5.0)
  ref_shim.py        compat shim that lets the UNMODIFIED reference functions run here (needs /root/reference)
  make_golden.py     runs the unmodified reference under the shim and writes tests/golden/*.json

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this package.
"""
