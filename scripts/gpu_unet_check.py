"""Print the end-to-end parity figures of the CUDA UNet against the CPU oracle (with a per-block trace)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests.unet_parity import run_parity
from view_neti_b200.sd21 import SD21, TINY

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
cases = {"tiny": (TINY, 2, 16, 16), "tiny1": (TINY, 1, 8, 8), "sd32": (SD21, 1, 32, 32), "sd64": (SD21, 1, 64, 64)}
cfg, nb, h, w = cases[which]
torch.set_num_threads(os.cpu_count())
t0 = time.time()
r = run_parity(cfg, nb, h, w, trace=True)
print(f"[{which}] {time.time() - t0:.1f}s  eps_mse {r['eps_mse']:.3e} eps_rel {r['eps_rel']:.3e} loss_rel {r['loss_rel']:.3e} "
      f"grad_flat_rel {r['grad_flat_rel']:.3e} grad_worst_rel {r['grad_worst_rel']:.3e}")
for k, v in r["trace_f"].items():
    print(f"  fwd {k}: {v:.3e}   bwd(dout): {r['trace_b'].get(k, float('nan')):.3e}")
for k, v in r["per_grad"].items():
    print(f"  grad {k}: {v:.3e}")
