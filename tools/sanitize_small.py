"""Small end-to-end exercise of the round-2 kernels for compute-sanitizer (memcheck): chains + pair fold + fused apply at batch 64,
TF32 plan, simple U-Net, ResNet18 encoder, uint8 frames, clip_sample."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import state_policy_diffusionmodel_b200 as spdm
from oracle import fixtures

B = 64
sd = fixtures.make_unet_weights(attention=True, seed=0)
esd = fixtures.make_encoder_weights()
batch = fixtures.make_batch(B, seed=1)
x_T = fixtures.make_xT(B)
sch = spdm.DDIMScheduler(num_train_timesteps=4, beta_schedule="linear", clip_sample=True, prediction_type="epsilon")
sch.set_timesteps(4)
for precision in ("bf16", "tf32"):
    plan = spdm.DenoisePlan(attention=True, precision=precision, batch_max=B, inpaint_rows=1, graph_steps=2)
    plan.load_unet_state_dict(sd)
    plan.load_encoder_state_dict(esd)
    plan.set_schedule("ddim", sch.coef_table(), sch.timesteps)
    u8 = (batch["image"] * 255).to(torch.uint8).permute(0, 1, 3, 4, 2).contiguous()
    plan.encode_cond(u8, batch["position"], batch["action"], batch["velocity"])
    inp = torch.cat([batch["position"][:, -1:], batch["action"][:, -1:]], dim=-1).reshape(B, -1)
    out = plan.sample(x_T, inpaint=inp)
    torch.cuda.synchronize()
    print(precision, float(out.abs().mean()))
    plan.close()
net = spdm.UNet(1, 1, 1000, global_cond_dim=1350).cuda().eval()
net.load_state_dict(fixtures.make_simple_unet_weights(seed=7))
with torch.no_grad():
    o = net(torch.rand(3, 1, 31, 5).cuda(), torch.tensor([5, 6, 7]).cuda(), torch.randn(3, 1, 10, 135).cuda())
print("simple", float(o.abs().mean()))
rp = spdm.DenoisePlan(attention=False, precision="bf16", batch_max=4, cond_dim=519, encoder="resnet18", graph_steps=0)
rp.load_encoder_state_dict(fixtures.make_resnet_weights())
f = rp.encode_images(torch.rand(7, 3, 96, 96))
torch.cuda.synchronize()
print("resnet", float(f.abs().mean()))
