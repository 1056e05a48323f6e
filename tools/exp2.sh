SPDM_CL_TIMING=1 python - <<'PY' > gpurun_out/exp_chain_timing.txt 2>&1
import torch, sys
sys.path.insert(0, '.')
import state_policy_diffusionmodel_b200 as spdm
from bench import synth_batch
torch.manual_seed(0)
B=256
m = spdm.Diffusion_DDIM(noise_steps=1000, obs_horizon=10, pred_horizon=30, observation_dim=135, prediction_dim=5, model="UNet_Film", inpaint_horizon=1).cuda().eval()
m.configure(precision="bf16", graph_steps=0, batch_max=B)
m.use_ddim(3)
devb = {k: v.cuda() for k, v in synth_batch(B, 10, 1).items()}
out = m.sample(devb, batched=True)
torch.cuda.synchronize()
PY
