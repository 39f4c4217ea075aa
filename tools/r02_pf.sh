set -x
for cfg in "0 0" "24 8192" "24 16384" "24 24576" "40 24576"; do
set -- $cfg
JF_L2_LOOKAHEAD=$1 JF_DEBUG_SKIP=$2 python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_pf_$1_$2.txt 2>&1; grep "forward 3\|gate_up\|o_gemm\|attn_combine" gpurun_out/r02_stage_pf_$1_$2.txt
done
