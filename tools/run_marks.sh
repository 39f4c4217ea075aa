# developer aid (GPU box): per-stage profile of one forward (no marks), then the fine-grained marks of the same forward
# usage: run_marks.sh NAME [extra stage_profile.py args]
OUT=${1:-n32}; shift
python tools/stage_profile.py --gen 64 --n 32 "$@" > gpurun_out/stage_$OUT.txt 2>&1
tail -12 gpurun_out/stage_$OUT.txt
JF_PROF_MARKS=gpurun_out/marks_$OUT.txt python tools/stage_profile.py --gen 64 --n 32 "$@" > gpurun_out/stage_${OUT}_marks.txt 2>&1
python tools/prof_marks.py gpurun_out/marks_$OUT.txt > gpurun_out/marks_${OUT}_agg.txt 2>&1
grep "^(" gpurun_out/marks_${OUT}_agg.txt
