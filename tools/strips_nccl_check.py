"""torchrun --nproc-per-node N tools/strips_nccl_check.py : one membrane on N GPUs (strips, the library's own NCCL refresh) == the same
membrane on one GPU, bit for bit, and the all-reduced bond.dat row equal to the single-GPU row. Prints one JSON line on rank 0.
(bench.py --gpus N runs the same check and reports it under "strips_check".)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
import torch
import torch.distributed as dist

from kmc_b200.strips import nccl_check

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
res = nccl_check(dist, local, molecules=int(os.environ.get("MOLECULES", "100000")), steps=int(os.environ.get("STEPS", "96")), every=int(os.environ.get("EVERY", "8")))
if dist.get_rank() == 0:
    print(json.dumps(res))
dist.destroy_process_group()
sys.exit(0 if res["equal_single_gpu"] and res["series_equal"] else 1)
