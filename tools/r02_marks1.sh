set -x
JF_PROF_MARKS=gpurun_out/marks_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_n1_marks.txt 2>&1
grep "SM clock" gpurun_out/r02_stage_n1_marks.txt
python tools/prof_marks2.py gpurun_out/marks_n1.txt 1870 | head -20
