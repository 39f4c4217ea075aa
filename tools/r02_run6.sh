set -x
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_d.err | tail -1 > gpurun_out/r02_bench_n1_d.json; python -c "import json; d=json.load(open('gpurun_out/r02_bench_n1_d.json')); print(d['value'], d['roofline']['frac'], d['roofline']['ms_per_forward'])"
JF_ATTN_HSPLIT=2 python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_d2.err | tail -1 > gpurun_out/r02_bench_n1_hs2.json; python -c "import json; d=json.load(open('gpurun_out/r02_bench_n1_hs2.json')); print(d['value'], d['roofline']['frac'], d['roofline']['ms_per_forward'])"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
