# session 3, run 3: SwiGLU epilogue through shared memory + bulk stores - parity tests, A/B, marks
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_engine_api.py tests/test_gpu_parity_7b.py -x -q 2>&1 | tail -4 > gpurun_out/s3_tests3.txt; tail -4 gpurun_out/s3_tests3.txt
for i in 1 2; do
JF_SWIGLU_BULK=0 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('direct', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bulk', d['value'], d['roofline']['ms_per_forward'])"
done
JF_PROF_MARKS=gpurun_out/s3_marks3_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage3_n1.txt 2>&1; tail -12 gpurun_out/s3_stage3_n1.txt
python tools/prof_marks2.py gpurun_out/s3_marks3_n1.txt 1860 | head -8
