# tensor-parallel batch on N GPUs: usage  bash tools/r02_tp.sh N
N=$1
set -x
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -s 2>&1 | tail -15 > gpurun_out/r02_tp_tests_n$N.txt; tail -12 gpurun_out/r02_tp_tests_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tp$N.err | tail -1 > gpurun_out/r02_bench_n${N}_tp$N.json; head -c 700 gpurun_out/r02_bench_n${N}_tp$N.json; echo
timeout 600 $TR tools/stage_profile.py --gen 64 --out gpurun_out/stage_profile_tp$N.json > gpurun_out/r02_stage_profile_tp$N.txt 2>&1; tail -14 gpurun_out/r02_stage_profile_tp$N.txt
if [ "$N" -le 4 ]; then
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --tp-attn shard 2>gpurun_out/bench_tp${N}sa.err | tail -1 > gpurun_out/r02_bench_n${N}_tp${N}_shardattn.json; head -c 400 gpurun_out/r02_bench_n${N}_tp${N}_shardattn.json; echo
fi
