"""One VAE encode + one decode at 512 x 512 for an `ncu --metrics gpu__time_duration.sum` launch list (after one warm
pass of each, so tensor maps / smem attributes are set).  Markers: a vn_memset of 1 / 2 / 3 bytes brackets the passes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200.models.vae import SD21_VAE, AutoencoderKL, init_state_dict

vae = AutoencoderKL(init_state_dict(SD21_VAE, 0), SD21_VAE, "cuda")
img = torch.rand(1, 3, 512, 512, device="cuda") * 2 - 1
z = torch.randn(1, 4, 64, 64, device="cuda")
vae.encode(img); vae.decode(z)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("encode"); vae.encode(img); torch.cuda.synchronize(); torch.cuda.nvtx.range_pop()
torch.cuda.nvtx.range_push("decode"); vae.decode(z); torch.cuda.synchronize(); torch.cuda.nvtx.range_pop()
