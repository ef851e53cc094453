"""BASELINE config 5 on one GPU: novel-view inference for one prompt = PromptManager.embed_prompt over the 50 DDIM
timesteps (batched conditioning path, frozen mappers) + sd_pipeline_call (50 steps x batched CFG UNet forward + fused
guidance / DDIM update) at 512x512 (64x64 latents), latents out.  Seeded random weights at the SD-2.1 shapes."""
import json
import os
import sys
import time

sys.path.insert(0, os.getcwd())
import torch

from view_neti_b200.prompt_manager import PromptManager
from view_neti_b200.schedulers import DDIMScheduler
from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.sd_pipeline_call import ViewNeTIPipeline, sd_pipeline_call
from view_neti_b200.training.synthetic import OBJECT_TOKEN_ID, VIEW_TOKEN_IDS, build_conditioning, synthetic_prompt
from view_neti_b200.unet import UNet2DConditionModel

dev = "cuda"
steps = int(os.environ.get("STEPS", 50))
hw = int(os.environ.get("SIZE", 512))
cond = build_conditioning(dev)
unet = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, dev)
sched = DDIMScheduler("v_prediction")
sched.set_timesteps(steps)
timesteps = [int(t) for t in sched.timesteps]
pm = PromptManager(tokenizer=None, text_encoder=cond, timesteps=timesteps, placeholder_view_token_ids=VIEW_TOKEN_IDS,
                   placeholder_object_token_ids=[OBJECT_TOKEN_ID], chunk=10)
ids = synthetic_prompt(1, dev)["input_ids"]
neg = torch.randn(1, 77, 1024, generator=torch.Generator().manual_seed(3)).to(dev)
pipe = ViewNeTIPipeline(unet, sched, negative_prompt_embeds=neg)
res = {}
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    embeds = pm.embed_prompt(ids)
    torch.cuda.synchronize(); t1 = time.time()
    out = sd_pipeline_call(pipe, embeds, height=hw, width=hw, num_inference_steps=steps, guidance_scale=7.5,
                           generator=torch.Generator(device=dev).manual_seed(rep), output_type="latent")
    torch.cuda.synchronize(); t2 = time.time()
    res = {"embed_prompt_ms": round((t1 - t0) * 1e3, 2), "denoise_ms": round((t2 - t1) * 1e3, 2),
           "ms_per_denoise_step": round((t2 - t1) * 1e3 / steps, 3), "images_per_s": round(1.0 / (t2 - t0), 3)}
print(json.dumps({"config": f"1 prompt, {steps} DDIM steps, CFG 7.5, {hw}x{hw}, latents out", **res,
                  "finite": bool(torch.isfinite(out.images).all())}))
