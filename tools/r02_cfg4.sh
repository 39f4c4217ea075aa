# BASELINE config 4: Qwen2.5-Math-7B shape, n=32, TP=4, batch 16, 512/512, multiblock + rejection recycling
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline --model qwen2.5-math-7b --mode multiblock --batch 16 2>gpurun_out/bench_cfg4.err | tail -1 > gpurun_out/r02_bench_cfg4_tp4_b16.json; head -c 300 gpurun_out/r02_bench_cfg4_tp4_b16.json; echo; tail -3 gpurun_out/bench_cfg4.err
timeout 900 $TR bench.py --gpus 4 --steps 2 --warmup 3 --no-cpu-baseline --model qwen2.5-math-7b --mode multiblock --batch 16 --tp-attn shard 2>gpurun_out/bench_cfg4sa.err | tail -1 > gpurun_out/r02_bench_cfg4_tp4_b16_shardattn.json; head -c 300 gpurun_out/r02_bench_cfg4_tp4_b16_shardattn.json; echo; tail -3 gpurun_out/bench_cfg4sa.err
timeout 600 $TR tools/stage_profile.py --gen 64 --batch 16 --mode multiblock --model qwen2.5-math-7b --out gpurun_out/stage_profile_cfg4.json > gpurun_out/r02_stage_profile_cfg4.txt 2>&1; tail -14 gpurun_out/r02_stage_profile_cfg4.txt
