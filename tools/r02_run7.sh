set -x
timeout 900 python -m pytest tests/test_gpu_tp_virtual.py -x -q -k "plain_process" 2>&1 | tail -15
