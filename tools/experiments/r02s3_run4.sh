set -x
JF_PROF_MARKS=gpurun_out/s3_marks4_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage4_n1.txt 2>&1; tail -12 gpurun_out/s3_stage4_n1.txt | head -4
python tools/prof_marks2.py gpurun_out/s3_marks4_n1.txt 1860 | grep "605"
JF_DEBUG_SKIP=32768 JF_PROF_MARKS=gpurun_out/s3_marks4b_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage4b_n1.txt 2>&1; tail -12 gpurun_out/s3_stage4b_n1.txt | head -4
python tools/prof_marks2.py gpurun_out/s3_marks4b_n1.txt 1860 | grep "605"
for i in 1 2; do
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fence', d['value'], d['roofline']['ms_per_forward'])"
JF_DEBUG_SKIP=32768 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nofence', d['value'], d['roofline']['ms_per_forward'])"
done
