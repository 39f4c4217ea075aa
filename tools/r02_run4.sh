set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_gputests_c.txt; tail -25 gpurun_out/r02_gputests_c.txt
