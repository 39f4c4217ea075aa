set -x
timeout 900 python -m pytest tests/test_gpu_tp_virtual.py tests/test_gpu_nongreedy.py -x -q 2>&1 | tail -12
