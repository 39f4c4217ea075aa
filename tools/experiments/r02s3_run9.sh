set -x
JF_DEBUG_SKIP=262144 JF_PROF_MARKS=gpurun_out/s3_marks9_n1.txt timeout 300 python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage9_n1.txt 2>&1; tail -12 gpurun_out/s3_stage9_n1.txt | head -4
python tools/prof_marks2.py gpurun_out/s3_marks9_n1.txt 1860 | head -8
