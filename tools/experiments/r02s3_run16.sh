set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 > gpurun_out/s3_tests16.txt; tail -3 gpurun_out/s3_tests16.txt
for i in 1 2; do
JF_LIB_PATH=build/libjacobi_b200_prev16.so python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step8', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('step16', d['value'], d['roofline']['ms_per_forward'])"
done
JF_PROF_MARKS=gpurun_out/s3_marks16_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage16_n1.txt 2>&1
python tools/prof_marks2.py gpurun_out/s3_marks16_n1.txt 1860 > gpurun_out/s3_marks16_tl.txt; head -6 gpurun_out/s3_marks16_tl.txt
