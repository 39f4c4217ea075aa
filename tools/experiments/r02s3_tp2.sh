# session 3: the latency changes under real tensor parallelism (2 GPUs): TP tests, TP bench, B=8 A/B on one GPU
set -x
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -s 2>&1 | tail -6 > gpurun_out/s3_tp_tests_n2.txt; tail -6 gpurun_out/s3_tp_tests_n2.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_tp2.err | tail -1 > gpurun_out/s3_bench_n2_tp2.json; python -c "import json; d=json.load(open('gpurun_out/s3_bench_n2_tp2.json')); print('tp2', d['value'], d.get('tp_check'), d['config'].get('parallelism'))"
JF_LIB_PATH=build/libjacobi_b200_sess3start.so timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tp2 prev', d['value'])"
CUDA_VISIBLE_DEVICES=0 JF_LIB_PATH=build/libjacobi_b200_sess3start.so python bench.py --no-cpu-baseline --mode multiblock --batch 8 --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b8 prev', d['value'], d['roofline']['ms_per_forward'])" &
CUDA_VISIBLE_DEVICES=1 python bench.py --no-cpu-baseline --mode multiblock --batch 8 --steps 2 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('b8 new', d['value'], d['roofline']['ms_per_forward'])"
wait
