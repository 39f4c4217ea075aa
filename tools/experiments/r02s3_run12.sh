set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_engine_api.py tests/test_gpu_nongreedy.py -x -q 2>&1 | tail -3 > gpurun_out/s3_tests12.txt; tail -3 gpurun_out/s3_tests12.txt
for i in 1 2; do
JF_EARLY_FENCE=0 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('late', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('early', d['value'], d['roofline']['ms_per_forward'])"
done
python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage12_n1.txt 2>&1; tail -12 gpurun_out/s3_stage12_n1.txt | head -9
