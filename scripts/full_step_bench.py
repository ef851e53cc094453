"""One COMPLETE ViewNeTI train step on one GPU through the drop-in API: batched conditioning path (CUDA mappers + CLIP
encoder) -> frozen SD-2.1 UNet forward -> fp32 MSE -> dgrad backward through the 32 context tensors ->
CLIP encoder backward -> mapper gradients -> AdamW.  Seeded random weights at the real shapes (no checkpoints offline)."""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch

from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.training.coach import Coach
from view_neti_b200.training.synthetic import build_conditioning, synthetic_prompt
from view_neti_b200.unet import UNet2DConditionModel

steps = int(os.environ.get("STEPS", 20))
dev = "cuda"
g = torch.Generator().manual_seed(0)
cond = build_conditioning(dev)
unet = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, dev)
coach = Coach(cfg=None, unet=unet, conditioning=cond, optimizer=torch.optim.AdamW(cond.parameters(), lr=1e-3),
              generator=torch.Generator(device=dev).manual_seed(1))
batch = synthetic_prompt(1, dev)
latents = torch.randn(1, 4, 64, 64, generator=g).to(dev)
if os.environ.get("VAE"):       # start from the image: vae.encode(pixel_values) every step (reference coach.py:165-169)
    from view_neti_b200.models.vae import SD21_VAE, AutoencoderKL, init_state_dict as vae_init
    coach.vae = AutoencoderKL(vae_init(SD21_VAE, 0), SD21_VAE, dev)
    batch["pixel_values"] = torch.rand(1, 3, 512, 512, generator=g).to(dev) * 2 - 1
    latents = None
losses = []
for _ in range(4):
    losses.append(float(coach.train_step(latents, batch)))
torch.cuda.synchronize()
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = coach.train_step(latents, batch)
e1.record()
torch.cuda.synchronize()
wall = (time.time() - t0) / steps * 1e3
print(json.dumps({"full_step_ms": round(e0.elapsed_time(e1) / steps, 3), "wall_ms": round(wall, 3),
                  "images_per_s": round(1e3 / (e0.elapsed_time(e1) / steps), 2), "first_losses": [round(l, 4) for l in losses],
                  "last_loss": round(float(loss), 4), "trainable_params": sum(p.numel() for p in cond.parameters())}))
