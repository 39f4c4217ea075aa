# round-2 first GPU batch: all GPU tests, smoke, the headline bench and a stage profile of the unchanged kernel
set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputests_a.txt; tail -5 gpurun_out/r02_gputests_a.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 > gpurun_out/r02_smoke_a.txt; cat gpurun_out/r02_smoke_a.txt
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_a.err | tail -1 > gpurun_out/r02_bench_n1_a.json
head -c 400 gpurun_out/r02_bench_n1_a.json
python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_profile_a.txt 2>&1; tail -12 gpurun_out/r02_stage_profile_a.txt
