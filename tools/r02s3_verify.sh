# session 3: verification of the final build: all GPU tests, smoke, headline bench line
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gputests_final.txt; cat gpurun_out/r02_gputests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r02_smoke_final.txt; cat gpurun_out/r02_smoke_final.txt
python bench.py --no-cpu-baseline --steps 3 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('final', d['value'], d['roofline']['ms_per_forward'], d['roofline']['frac'])"
