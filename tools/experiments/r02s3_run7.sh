set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3 > gpurun_out/s3_tests7.txt; tail -3 gpurun_out/s3_tests7.txt
for i in 1 2; do
JF_L2_AHEAD=0 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('noahead', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ahead8', d['value'], d['roofline']['ms_per_forward'])"
done
JF_PROF_MARKS=gpurun_out/s3_marks7_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage7_n1.txt 2>&1; tail -12 gpurun_out/s3_stage7_n1.txt
python tools/prof_marks2.py gpurun_out/s3_marks7_n1.txt 1860 | head -8
