set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/s3_tests14.txt; tail -4 gpurun_out/s3_tests14.txt
for i in 1 2; do
JF_LIB_PATH=build/libjacobi_b200_sess3start.so python bench.py --no-cpu-baseline --mode multiblock --batch 8 --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b8 prev', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --mode multiblock --batch 8 --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b8 new', d['value'], d['roofline']['ms_per_forward'])"
done
python tools/stage_profile.py --gen 64 --batch 8 --mode multiblock > gpurun_out/s3_stage14_b8.txt 2>&1; tail -12 gpurun_out/s3_stage14_b8.txt
