# round-2 end-of-round measurement batch (GPU box, 1 GPU): tests, smoke, bench (ours + reference arm), stage profile + marks,
# ncu launch list + full capture, batch-8 and 0.5B configurations; outputs under gpurun_out/
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gputests_final.txt; cat gpurun_out/r02_gputests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/r02_smoke_final.txt; cat gpurun_out/r02_smoke_final.txt
python bench.py 2>gpurun_out/bench_n1.err | tail -1 > gpurun_out/r02_bench_n1.json; head -c 400 gpurun_out/r02_bench_n1.json; echo
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tail -1 > gpurun_out/r02_bench_reference_arm.json; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_profile_n1.txt 2>&1; grep -A10 "^forward 3" gpurun_out/r02_stage_profile_n1.txt
python bench.py --no-cpu-baseline --mode multiblock --batch 8 --steps 2 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_bench_7b_multiblock_b8.json; head -c 250 gpurun_out/r02_bench_7b_multiblock_b8.json; echo
python bench.py --no-cpu-baseline --model qwen2-0.5b --n 8 2>/dev/null | tail -1 > gpurun_out/r02_bench_05b_n8.json; head -c 250 gpurun_out/r02_bench_05b_n8.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:jf_persistent -c 40 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 1 --warmup 3 --gen 64 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jf_persistent -s 7 -c 1 -o gpurun_out/prof_r02_final python bench.py --steps 1 --warmup 3 --gen 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
