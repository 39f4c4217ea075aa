set -x
JF_DEBUG_SKIP=524288 JF_PROF_MARKS=gpurun_out/s3_marks13_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage13_n1.txt 2>&1; tail -12 gpurun_out/s3_stage13_n1.txt | head -9
python tools/prof_marks2.py gpurun_out/s3_marks13_n1.txt 1860 > gpurun_out/s3_marks13_tl.txt; head -6 gpurun_out/s3_marks13_tl.txt
