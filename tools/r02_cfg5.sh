# BASELINE config 5: Qwen2.5-Coder-7B shape, n=64, TP=8, batch 32 (single-block Jacobi; prompts of 256 tokens), plus the
# batch-1 headline on 8 GPUs with the end-of-round build
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q 2>&1 | tail -3
timeout 900 $TR bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline --block-size 64 --batch 32 --prompt 256 --gen 256 --mode single 2>gpurun_out/bench_cfg5.err | tail -1 > gpurun_out/r02_bench_cfg5_tp8_b32_n64.json; head -c 260 gpurun_out/r02_bench_cfg5_tp8_b32_n64.json; echo; tail -3 gpurun_out/bench_cfg5.err
