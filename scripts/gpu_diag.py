"""Run every per-op parity check and print all error figures (does not stop at the first failure)."""
import os
import sys
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import opchecks

only = sys.argv[1] if len(sys.argv) > 1 else ""
bad = 0
for fn, kw in opchecks.all_checks():
    if only and only not in fn.__name__:
        continue
    t0 = time.time()
    try:
        res = fn(**kw)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        bad += 1
        print(f"EXC  {fn.__name__} {kw}: {e!r}", flush=True)
        traceback.print_exc()
        try:
            torch.cuda.synchronize()
        except Exception as e2:  # noqa: BLE001
            print("device is in an error state, stopping:", e2, flush=True)
            break
        continue
    for label, err, tol in res:
        ok = err <= tol
        bad += 0 if ok else 1
        print(f"{'ok  ' if ok else 'FAIL'} {label}: {err:.3e} (tol {tol:.1e})", flush=True)
print(f"done, {bad} failing", flush=True)
