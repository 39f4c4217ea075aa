set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity_7b.py -x -q 2>&1 | tail -3 > gpurun_out/s3_tests19.txt; tail -3 gpurun_out/s3_tests19.txt
for i in 1 2; do
JF_LIB_PATH=build/libjacobi_b200_prev19.so python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('oneround', d['value'], d['roofline']['ms_per_forward'])"
done
