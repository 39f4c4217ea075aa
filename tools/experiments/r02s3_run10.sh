set -x
for d in 0 12 24; do
JF_DEBUG_SKIP=$((d<<20)) python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('depth $d', d['value'], d['roofline']['ms_per_forward'])"
done
JF_DEBUG_SKIP=$((16<<20)) python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage10_n1.txt 2>&1; tail -12 gpurun_out/s3_stage10_n1.txt
