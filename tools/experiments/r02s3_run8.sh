set -x
JF_PROF_MARKS=gpurun_out/s3_marks8_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage8_n1.txt 2>&1; tail -12 gpurun_out/s3_stage8_n1.txt | head -8
python tools/prof_marks2.py gpurun_out/s3_marks8_n1.txt 1860 | head -8
JF_ATTN_PREFETCH=0 JF_PROF_MARKS=gpurun_out/s3_marks8b_n1.txt python tools/stage_profile.py --gen 64 > gpurun_out/s3_stage8b_n1.txt 2>&1; tail -12 gpurun_out/s3_stage8b_n1.txt | head -8
python tools/prof_marks2.py gpurun_out/s3_marks8b_n1.txt 1860 | head -8
