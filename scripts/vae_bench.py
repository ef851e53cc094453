"""VAE encode (every train step, reference training/coach.py:165-169) and decode (per image, reference
sd_pipeline_call.py:115) on one GPU: ms per call by CUDA events, algorithmic TFLOP/s (2*MAC of every conv / linear /
attention product), launches per call.  Inputs are resident; the 84 M-parameter weights + ~1 GB of activations per call
stream through L2 (126 MB), which flushes it between iterations."""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from view_neti_b200 import ops
from view_neti_b200.models.vae import SD21_VAE, AutoencoderKL, init_state_dict, param_table


def flops(cfg, H, W):
    """(encode, decode) algorithmic FLOPs for an H x W image."""
    shapes = {n: s for n, s, _ in param_table(cfg)}
    enc = dec = 0.0
    nlev = len(cfg.block_out_channels)
    for name, s in shapes.items():
        if not name.endswith(".weight") or len(s) == 1:
            continue
        side = name.split(".")[0]
        if name.startswith(("quant_conv", "post_quant_conv")):
            lev, side = nlev - 1, "encoder" if name.startswith("quant") else "decoder"
        elif ".down_blocks." in name:
            lev = int(name.split(".")[2]) + (1 if "downsamplers" in name else 0)
        elif ".up_blocks." in name:
            i = int(name.split(".")[2])
            lev = nlev - 1 - i - (1 if "upsamplers" in name else 0)
        elif "mid_block" in name or name.endswith(("encoder.conv_out.weight", "decoder.conv_in.weight")):
            lev = nlev - 1
        else:
            lev = 0
        px = (H >> lev) * (W >> lev)
        f = 2.0 * px * math.prod(s)
        if side == "encoder":
            enc += f
        else:
            dec += f
    hw = (H >> (nlev - 1)) * (W >> (nlev - 1))
    att = 2 * 2.0 * hw * hw * cfg.block_out_channels[-1]
    return enc + att, dec + att


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    cfg = SD21_VAE
    vae = AutoencoderKL(init_state_dict(cfg, 0), cfg, "cuda")
    out = {"weights_MB": round(vae.engine.weight_bytes() / 2 ** 20, 1), "gn_fused_elems": vae.engine.FUSED_GN_ELEMS,
           "thin_on_gemm": vae.engine.THIN_ON_GEMM}
    shapes = [(1, 512, 512)] if os.environ.get("VAE_BENCH_QUICK") else [(1, 512, 512), (1, 512, 384), (2, 512, 512)]
    for (nb, H, W) in shapes:
        img = torch.rand(nb, 3, H, W, device="cuda") * 2 - 1
        z = torch.randn(nb, 4, H // 8, W // 8, device="cuda")
        fe, fd = flops(cfg, H, W)
        ops.launch_count_reset()
        vae.encode(img)
        le = ops.launch_count()
        ops.launch_count_reset()
        vae.decode(z)
        ld = ops.launch_count()
        te = timed(lambda: vae.encode(img))
        td = timed(lambda: vae.decode(z))
        out[f"{nb}x{H}x{W}"] = {
            "encode_ms": round(te, 3), "encode_gflop": round(nb * fe / 1e9, 1), "encode_tflops": round(nb * fe / te / 1e9, 1),
            "encode_launches": le, "images_per_s_encode": round(nb / te * 1e3, 1),
            "decode_ms": round(td, 3), "decode_gflop": round(nb * fd / 1e9, 1), "decode_tflops": round(nb * fd / td / 1e9, 1),
            "decode_launches": ld, "scratch_MB": round(vae.engine.scratch_bytes() / 2 ** 20, 1)}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.environ.get("VAE_BENCH_OUT", "gpurun_out/vae_bench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
