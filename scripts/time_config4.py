"""Time BASELINE config 4 (16,384 loops x 401 steps, n = 20, m = p = 4) through the fused FP64 tensor-core kernel for each
CTA size (set option "dmma_warps" = 1 | 2 | 4; default 1).  Uses bench.secondary_config4."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from direct_data_driven_mpc_b200 import _lib

peak = _lib.probe_fp64_tflops(True)
for w in (None, 4, 2, 1) if "--all" in sys.argv else (None,):
    B = int(os.environ.get("CONFIG4_LOOPS", "16384"))
    r = bench.secondary_config4(torch.device("cuda", 0), B, fp64_peak=peak, dmma_warps=w)
    print("loops", B, "warps/CTA", w or "default", "fp64 peak", round(peak, 2),
          json.dumps({k: {kk: round(vv, 4) for kk, vv in v.items()} for k, v in r.items() if k.startswith("n_mpc")}), flush=True)
