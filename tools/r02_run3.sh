set -x
timeout 900 python tools/diag_7b_rows.py 2>&1 | grep -v Warning | tail -40 > gpurun_out/r02_diag7b.txt; cat gpurun_out/r02_diag7b.txt
timeout 900 python -m pytest tests/test_gpu_tp_virtual.py -x -q -s 2>&1 | tail -40 > gpurun_out/r02_tpv_b.txt; tail -30 gpurun_out/r02_tpv_b.txt
timeout 600 python -m pytest tests/test_gpu_checkpoint_dir.py -x -q 2>&1 | tail -60 > gpurun_out/r02_ckpt_b.txt; tail -30 gpurun_out/r02_ckpt_b.txt
