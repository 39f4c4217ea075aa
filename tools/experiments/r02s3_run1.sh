# session 3, run 1: fused attention + combine (tagged partial results) - parity tests, A/B against the previous library and
# against JF_FUSE_ATTN=0, stage profile with marks
set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_engine_api.py tests/test_gpu_parity_7b.py -x -q 2>&1 | tail -8 > gpurun_out/s3_tests1.txt; tail -8 gpurun_out/s3_tests1.txt
for i in 1 2; do
JF_LIB_PATH=build/libjacobi_b200_sess3start.so python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prev', d['value'], d['roofline']['ms_per_forward'])"
JF_FUSE_ATTN=0 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new-nofuse', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new', d['value'], d['roofline']['ms_per_forward'])"
done
python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage_n1.txt 2>&1; tail -12 gpurun_out/s3_stage_n1.txt
JF_PROF_MARKS=gpurun_out/s3_marks_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage_n1_marks.txt 2>&1
python tools/prof_marks2.py gpurun_out/s3_marks_n1.txt 1860 | head -12
