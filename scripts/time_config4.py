"""Time BASELINE config 4 (16,384 loops x 401 steps, n = 20, m = p = 4) through the fused FP64 tensor-core kernel for each
CTA size (DDMPC_DMMA_WARPS = 1 | 2 | 4; unset = the launcher's own choice).  Uses bench.secondary_config4."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

for w in (None, "4", "2", "1") if "--all" in sys.argv else (None,):
    if w is None:
        os.environ.pop("DDMPC_DMMA_WARPS", None)
    else:
        os.environ["DDMPC_DMMA_WARPS"] = w
    B = int(os.environ.get("CONFIG4_LOOPS", "16384"))
    r = bench.secondary_config4(torch.device("cuda", 0), B)
    print("loops", B, "warps/CTA", w or "auto", json.dumps({k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in r.items() if k.startswith("n_mpc")}), flush=True)
