set -x
python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_n1_pred.txt 2>&1; grep -A12 "^forward 3" gpurun_out/r02_stage_n1_pred.txt
JF_PROF_MARKS=gpurun_out/marks_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_n1_marks.txt 2>&1
python tools/prof_marks2.py gpurun_out/marks_n1.txt 1870 | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
