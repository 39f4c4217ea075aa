set -x
python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_profile_n1.txt 2>&1; grep -A12 "^forward 3" gpurun_out/r02_stage_profile_n1.txt
JF_DEBUG_SKIP=32768 python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_profile_n1_xspin.txt 2>&1; grep -A12 "^forward 3" gpurun_out/r02_stage_profile_n1_xspin.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 1 --warmup 3 --gen 64 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jf_persistent -s 7 -c 1 -o gpurun_out/prof_r02 python bench.py --steps 1 --warmup 3 --gen 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_full.log
