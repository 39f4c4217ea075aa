# session 3, run 2: worker loop without a stack-resident context (WorkerView) - parity tests, A/B against the previous library
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_engine_api.py -x -q 2>&1 | tail -4 > gpurun_out/s3_tests2.txt; tail -4 gpurun_out/s3_tests2.txt
for i in 1 2; do
JF_LIB_PATH=build/libjacobi_b200_sess3start.so python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new', d['value'], d['roofline']['ms_per_forward'])"
done
python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage2_n1.txt 2>&1; tail -12 gpurun_out/s3_stage2_n1.txt
