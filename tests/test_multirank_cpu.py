"""CPU, world_size 2, gloo: the host-side logic of the N>1 path (replica partition, seed derivation, max-over-ranks
timing, gather of the per-replica series). The data path has no collective (independent replicas / patches)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, os.path.join(%r, "kmc-with-a-diffusion-reaction-algorithm_b200"))
    import torch.distributed as dist
    from kmc_b200.sharding import replica_range, rank_seed, max_over_ranks, gather_series
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = replica_range(rank, world, 7)
    rows = [(r, {"seed": rank_seed(50, rank, world, 7) + (r - lo), "bond_num": 10 * r}) for r in range(lo, hi)]
    allrows = gather_series(rows, dist)
    ms = max_over_ranks(10.0 + 5.0 * rank, dist)
    dist.barrier()
    if rank == 0:
        print(json.dumps({"rows": allrows, "ms": ms, "world": world}))
    dist.destroy_process_group()
''') % ROOT


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["ms"] == 15.0
    assert [r[0] for r in out["rows"]] == list(range(7))
    assert [r[1]["seed"] for r in out["rows"]] == [50 + r for r in range(7)]   # replica r keeps seed base + r for any GPU count
