set -x
for tc in 0 1; do
JF_ATTN_TC=$tc timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_parity_7b.py tests/test_gpu_engine_api.py tests/test_gpu_nongreedy.py -x -q 2>&1 | tail -3
JF_ATTN_TC=$tc python tools/stage_profile.py --gen 64 > gpurun_out/r02_stage_n1_tc$tc.txt 2>&1; grep "^forward 3\|attention\|attn_combine" gpurun_out/r02_stage_n1_tc$tc.txt
JF_ATTN_TC=$tc python tools/stage_profile.py --gen 64 --batch 8 --mode multiblock > gpurun_out/r02_stage_b8_tc$tc.txt 2>&1; grep "^forward 3\|attention\|attn_combine" gpurun_out/r02_stage_b8_tc$tc.txt
done
