python tools/train_diag.py bf16 32 attn 2>&1 | grep -vE "^\s+(down|up|bot|sa|inc|outc)" | tail -12
echo ---- simt
SPDM_ENC_W1_SIMT=1 python tools/train_diag.py bf16 32 attn 2>&1 | grep -E "vision_encoder|global" | head -12
