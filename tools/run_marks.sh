# developer aid (GPU box): per-stage profile + fine-grained marks of one forward, optional env passthrough
OUT=${1:-n32}
JF_PROF_MARKS=gpurun_out/marks_$OUT.txt python tools/stage_profile.py --gen 64 --n 32 > gpurun_out/stage_$OUT.txt 2>&1
python tools/prof_marks.py gpurun_out/marks_$OUT.txt > gpurun_out/marks_${OUT}_agg.txt 2>&1
tail -14 gpurun_out/stage_$OUT.txt
grep "^(" gpurun_out/marks_${OUT}_agg.txt
