set -x
timeout 900 python -m pytest tests/test_gpu_nongreedy.py -x -q -s 2>&1 | tail -40 > gpurun_out/r02_ng_a.txt; tail -40 gpurun_out/r02_ng_a.txt
