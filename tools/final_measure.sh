# round-end measurement batch (GPU box, 1 GPU): tests, bench (ours + reference arm), ncu launch list + full capture,
# stage profile; outputs under gpurun_out/
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/r01_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tail -1 > gpurun_out/r01_bench_reference_arm.json
python tools/stage_profile.py --gen 64 > gpurun_out/r01_stage_profile.txt 2>&1
JF_PROF_MARKS=gpurun_out/marks_final.txt python tools/stage_profile.py --gen 64 > /dev/null 2>&1; python tools/prof_marks.py gpurun_out/marks_final.txt > gpurun_out/r01_stage_marks.txt 2>&1
python bench.py --no-cpu-baseline --mode multiblock --batch 8 --steps 2 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_bench_7b_multiblock_b8.json
python bench.py --no-cpu-baseline --model qwen2-0.5b --n 8 2>/dev/null | tail -1 > gpurun_out/r01_bench_05b_n8.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_ncu_launches.csv python bench.py --steps 1 --warmup 3 --gen 64 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jf_persistent -s 7 -c 1 -o gpurun_out/prof_r01_final2 python bench.py --steps 1 --warmup 3 --gen 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
head -c 600 gpurun_out/r01_bench_n1.json
