"""Host-side time of the phases of Coach.train_step (no synchronisation inside the step: what the Python thread spends before
it can enqueue the next phase).  If the sum exceeds the device time of a step, the step is host-bound."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.training.coach import Coach
from view_neti_b200.training.synthetic import build_conditioning, synthetic_prompt
from view_neti_b200.unet import UNet2DConditionModel

dev = "cuda"
cond = build_conditioning(dev)
unet = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, dev)
opt = torch.optim.AdamW(cond.parameters(), lr=1e-3)
coach = Coach(cfg=None, unet=unet, conditioning=cond, optimizer=opt, generator=torch.Generator(device=dev).manual_seed(1))
batch = synthetic_prompt(1, dev)
latents = torch.randn(1, 4, 64, 64, device=dev)
for _ in range(6):
    coach.train_step(latents, batch)
torch.cuda.synchronize()
acc = {}
N = 20
t_all = time.perf_counter()
for _ in range(N):
    t0 = time.perf_counter()
    noise = torch.randn(latents.shape, generator=coach.generator, device=dev)
    ts = torch.randint(0, 1000, (1,), generator=coach.generator, device=dev).long()
    noisy = coach.noise_scheduler.add_noise(latents, noise, ts)
    t1 = time.perf_counter()
    hs = coach.get_text_conditioning(input_ids=batch["input_ids"], timesteps=ts,
                                     input_ids_placeholder_object=batch["input_ids_placeholder_object"],
                                     input_ids_placeholder_view=batch["input_ids_placeholder_view"], device=dev)
    t2 = time.perf_counter()
    pred = unet(noisy, ts, hs).sample
    t3 = time.perf_counter()
    target = coach.noise_scheduler.get_velocity(latents, noise, ts)
    loss = F.mse_loss(pred.float(), target.float())
    t4 = time.perf_counter()
    loss.backward()
    t5 = time.perf_counter()
    opt.step(); opt.zero_grad(set_to_none=True)
    t6 = time.perf_counter()
    for k, v in (("noise+add_noise", t1 - t0), ("conditioning (mappers + CLIP fwd enqueue)", t2 - t1), ("unet forward enqueue", t3 - t2),
                 ("target + mse", t4 - t3), ("backward (UNet bwd + CLIP bwd + mappers)", t5 - t4), ("AdamW", t6 - t5)):
        acc[k] = acc.get(k, 0.0) + v
torch.cuda.synchronize()
wall = (time.perf_counter() - t_all) / N * 1e3
print(f"wall per step {wall:.3f} ms (host + device, synchronised only at the end)")
for k, v in acc.items():
    print(f"  host {v / N * 1e3:7.3f} ms  {k}")
print(f"  host total {sum(acc.values()) / N * 1e3:.3f} ms")
