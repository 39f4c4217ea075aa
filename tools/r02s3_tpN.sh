# session 3: tensor-parallel re-measurement on N GPUs: usage  bash tools/r02s3_tpN.sh N [tests]
N=$1
set -x
if [ "$2" = "tests" ]; then timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -s 2>&1 | tail -4 > gpurun_out/r02_tp_tests_n$N.txt; tail -4 gpurun_out/r02_tp_tests_n$N.txt; fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tp$N.err | tail -1 > gpurun_out/r02_bench_n${N}_tp$N.json; head -c 500 gpurun_out/r02_bench_n${N}_tp$N.json; echo
