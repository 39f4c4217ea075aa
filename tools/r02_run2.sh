# round-2: virtual-rank TP tests + the existing GPU suite (single GPU)
set -x
timeout 900 python -m pytest tests/test_gpu_tp_virtual.py -x -q -s 2>&1 | tail -40 > gpurun_out/r02_tpv_a.txt; tail -30 gpurun_out/r02_tpv_a.txt
timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_tp_virtual.py 2>&1 | tail -15 > gpurun_out/r02_gputests_b.txt; tail -8 gpurun_out/r02_gputests_b.txt
