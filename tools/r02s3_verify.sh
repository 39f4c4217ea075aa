# session 3: verification of the final build: all GPU tests and the smoke check
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_gputests_final.txt; cat gpurun_out/r02_gputests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 > gpurun_out/r02_smoke_final.txt; cat gpurun_out/r02_smoke_final.txt
