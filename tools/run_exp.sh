# developer aid (GPU box): stage profile under a list of JF_DEBUG_SKIP values (results invalid, timing only)
for v in "$@"; do   # bits: 1 rope, 2 attention, 4 combine, 8 o-norm stage bodies; 16 no W run-ahead; 2048 launch marks
  echo "=== JF_DEBUG_SKIP=$v"
  JF_DEBUG_SKIP=$v python tools/stage_profile.py --gen 64 --n 32 2>&1 | grep -E "forward|attention|attn_combine|qkv_rope"
done
