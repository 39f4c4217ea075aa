set -x
timeout 900 python -m pytest tests/test_gpu_tp_virtual.py -x -q -s 2>&1 | tail -60 > gpurun_out/r02_tpv_a.txt; tail -40 gpurun_out/r02_tpv_a.txt
timeout 600 python -m pytest tests/test_gpu_checkpoint_dir.py -x -q 2>&1 | tail -60 > gpurun_out/r02_ckpt_a.txt; tail -50 gpurun_out/r02_ckpt_a.txt
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tp_virtual.py --deselect tests/test_gpu_checkpoint_dir.py 2>&1 | tail -15 > gpurun_out/r02_gputests_b.txt; tail -8 gpurun_out/r02_gputests_b.txt
