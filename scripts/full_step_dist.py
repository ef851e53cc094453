"""Batch-parallel COMPLETE train steps on N GPUs of one box (BASELINE configs 2/3: per-GPU batch 1, frozen weights replicated,
ONE all-reduce of the flat mapper-gradient buffer per step - SURVEY.md 8e):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/full_step_dist.py

Every rank draws its own noise / timesteps / latents (per-rank generator), gradients of M_v + M_o (283 392 floats) are
averaged by FlatGradAllReducer inside Coach.train_step, and after the run the ranks compare their mapper parameters:
they must be bit-identical (same initial weights + same averaged gradients)."""
import json
import os
import sys

sys.path.insert(0, os.getcwd())
import torch
import torch.distributed as dist

from view_neti_b200.sd21 import SD21, init_state_dict
from view_neti_b200.training.coach import Coach
from view_neti_b200.training.synthetic import build_conditioning, synthetic_prompt
from view_neti_b200.unet import UNet2DConditionModel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
steps = int(os.environ.get("STEPS", 20))
cond = build_conditioning(dev, seed=0)                       # identical initial mappers on every rank
unet = UNet2DConditionModel(init_state_dict(SD21, 0), SD21, dev)
coach = Coach(cfg=None, unet=unet, conditioning=cond, optimizer=torch.optim.AdamW(cond.parameters(), lr=1e-3),
              generator=torch.Generator(device=dev).manual_seed(100 + rank))
batch = synthetic_prompt(1, dev)
latents = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(200 + rank)).to(dev)
for _ in range(4):
    coach.train_step(latents, batch)
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = coach.train_step(latents, batch)
e1.record()
torch.cuda.synchronize(); dist.barrier()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
flat = torch.cat([p.detach().reshape(-1) for p in cond.parameters()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
losses = [torch.zeros(1, device=dev) for _ in range(world)]
dist.all_gather(losses, loss.reshape(1).float())
if rank == 0:
    print(json.dumps({"n_gpus": world, "full_step_ms": round(float(ms), 3), "images_per_s": round(world * 1e3 / float(ms), 2),
                      "mapper_params_identical_across_ranks": same, "per_rank_loss": [round(float(l), 4) for l in losses],
                      "allreduce_floats": int(flat.numel())}))
dist.destroy_process_group()
