"""NCCL leg of the sharding tests (the CPU suite covers the same logic on gloo, tests/test_sharding_gloo.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/check_nccl_sharding.py

Every rank runs its contiguous shard of a config-3 job (equal and unequal shard sizes, NONE and CONVEX slack bound) through
run_sharded_closed_loops - global scenario ids as Philox streams, all_gather of the trajectories over NCCL - and rank 0
compares the gathered result BIT for bit with the whole job run on its own GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from direct_data_driven_mpc_b200 import ControllerSet, scenarios as S
from direct_data_driven_mpc_b200.sharding import run_sharded_closed_loops

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
dev = torch.device("cuda", torch.cuda.current_device())
ok = True
for total, slack, c in ((16384, 0, 1.0), (16384 + 37, 0, 1.0), (8192, 1, 1.0), (4096 + 5, 1, 0.3)):
    sc = S.config3_batch(total, seed=0)
    prm, plant = sc["params"], sc["plant"]
    cs = ControllerSet(prm["n"], 2, 2, sc["u_d"], sc["y_d"], prm["L"], prm["Q"], prm["R"], prm["eps_max"], prm["lamb_alpha"],
                       prm["lamb_sigma"], c, slack, 1, 4, True, device=dev)
    batch = {k: sc[k] for k in ("x0", "u_past0", "y_past0", "u_s", "y_s")}
    res = run_sharded_closed_loops(cs, plant, batch, 401, noise_seed=0, noise_eps=0.002, gather="all")
    u, y = res[0], res[1]
    if rank == 0:
        uf, yf, st, it = cs.closed_loop(plant, sc["x0"], sc["u_past0"], sc["y_past0"], sc["u_s"], sc["y_s"], 401, noise_seed=0,
                                        scenario_id0=0, noise_eps=0.002)
        # (shards of another size may take another kernel: the same numbers to 1e-9, bit-identical when the kernel is the same)
        du, dy = float((u - uf).abs().max()), float((y - yf).abs().max())
        same = du <= 1e-9 and dy <= 1e-9 and int(st.max()) == 0
        ok = ok and same
        print(f"{world} ranks, {total} loops, slack {slack}, c {c}: gathered {tuple(u.shape)}, max |du| {du:.2e} |dy| {dy:.2e} "
              f"bit-identical {bool(torch.equal(u, uf) and torch.equal(y, yf))} -> {'ok' if same else 'MISMATCH'}", flush=True)
    dist.barrier()
if rank == 0:
    print("RESULT:", "PASS" if ok else "FAIL")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
