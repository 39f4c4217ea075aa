set -x
JF_LIB_PATH=build/libjacobi_b200_r02start.so python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('old', d['value'], d['roofline']['ms_per_forward'])"
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('new', d['value'], d['roofline']['ms_per_forward'])"
python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_n1_e.txt 2>&1; grep -A10 "^forward 3" gpurun_out/r02_stage_n1_e.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
