"""torchrun --nproc-per-node N tools/strips_nccl_check.py : one membrane on N GPUs (strips, NCCL exchange) == the same membrane on
one GPU, bit for bit. Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import kmc_b200
from kmc_b200.strips import DistStrips, halo_for
from common import apply_regime

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
na, nb, regime, seed, steps, every = 60000, 20000, "hot", 5, int(os.environ.get("STEPS", "96")), 4
na, nb = (90000, 30000) if int(os.environ.get("WORLD_SIZE", "1")) > 2 else (na, nb)
L = 52000.0 if na == 60000 else 78000.0
box = (L, L, 400.0)
mk = lambda cap_a, cap_b: apply_regime(kmc_b200.default_params(box=box, n_receptor=cap_a, n_ligand=cap_b, seed=seed, device=local), regime)
# every rank builds the same global start state (deterministic initialiser) and, for the check, the single-GPU answer
k = kmc_b200.Kmc(mk(na, nb)); k.init_random(seed=17, sort_cells=True)
start = k.get_packed()
t0 = time.perf_counter(); k.step(steps); k.sync(); t_single = time.perf_counter() - t0
end = k.get_packed(); bonds = k.series()["bond_num"]; k.close()
ds = DistStrips(mk(int(na * 0.9), int(nb * 0.9)) if world > 1 else mk(na, nb), every, halo_width=halo_for(every), dist=dist)
ds.load_global(*start)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter(); ds.step(steps); ds.k.sync(); dist.barrier(); t_strips = time.perf_counter() - t0
rec, lig, rl, rs, rc = ds.gather(na, nb)
ok = bool(np.array_equal(rec, end[0]) and np.array_equal(lig, end[1]) and np.array_equal(rl, end[2]) and np.array_equal(rs, end[3]) and np.array_equal(rc, end[4]))
if rank == 0:
    print(json.dumps({"strips_equal_single_gpu": ok, "world": world, "molecules": na + nb, "steps": steps, "refresh_every": every, "bonds": bonds,
                      "single_gpu_s": t_single, "strips_s": t_strips, "bytes_sent_per_rank": ds.bytes_sent, "backend": dist.get_backend()}))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
