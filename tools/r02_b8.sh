set -x
run() { name=$1; shift; env "$@" python tools/stage_profile.py --gen 64 --batch 8 --mode multiblock $EXTRA > gpurun_out/r02_stage_b8_$name.txt 2>&1; grep -A12 "^forward 3" gpurun_out/r02_stage_b8_$name.txt; }
EXTRA="" run base A=1
EXTRA="" run nofusenorm JF_FUSE_NORM=0
EXTRA="--attn-chunk 128" run chunk128 A=1
EXTRA="--attn-chunk 256" run chunk256 A=1
