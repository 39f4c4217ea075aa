N=$1
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
JF_PROF_MARKS=gpurun_out/marks_tp$N.txt timeout 600 $TR tools/stage_profile.py --gen 64 --out gpurun_out/stage_profile_tp${N}m.json > gpurun_out/r02_stage_profile_tp${N}_marks.txt 2>&1
python tools/prof_marks.py gpurun_out/marks_tp$N.txt > gpurun_out/marks_tp${N}_agg.txt 2>&1; grep "^(" gpurun_out/marks_tp${N}_agg.txt | grep "50\|90"
JF_DEBUG_SKIP=4096 JF_PROF_MARKS=gpurun_out/marks_tp${N}g.txt timeout 600 $TR tools/stage_profile.py --gen 64 --out gpurun_out/stage_profile_tp${N}g.json > gpurun_out/r02_stage_profile_tp${N}_gpuscope.txt 2>&1
tail -12 gpurun_out/r02_stage_profile_tp${N}_gpuscope.txt
python tools/prof_marks.py gpurun_out/marks_tp${N}g.txt > gpurun_out/marks_tp${N}g_agg.txt 2>&1; grep "^(" gpurun_out/marks_tp${N}g_agg.txt | grep "50\|90"
