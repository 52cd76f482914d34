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


RING = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, os.path.join(%r, "kmc-with-a-diffusion-reaction-algorithm_b200"))
    import torch, torch.distributed as dist
    from kmc_b200.strips import ring_exchange
    dist.init_process_group("gloo")
    r, n = dist.get_rank(), dist.get_world_size()
    ok = True
    for trial in range(3):
        to_low = bytes([r, 0, trial]) * (5 + r + trial); to_high = bytes([r, 1, trial]) * (2 + 3 * r)
        if trial == 2: to_low = b""
        from_low, from_high = ring_exchange(torch, dist, "cpu", to_low, to_high)
        lo, hi = (r - 1) %% n, (r + 1) %% n
        exp_low = bytes([lo, 1, trial]) * (2 + 3 * lo)                       # the lower neighbour's to_high
        exp_high = b"" if trial == 2 else bytes([hi, 0, trial]) * (5 + hi + trial)   # the upper neighbour's to_low
        ok = ok and from_low == exp_low and from_high == exp_high
    res = [None] * n
    dist.all_gather_object(res, ok)
    if r == 0: print(json.dumps({"ok": all(res), "world": n}))
    dist.destroy_process_group()
''') % ROOT


def _run_ring(tmp_path, nproc, port):
    script = tmp_path / ("ring%d.py" % nproc)
    script.write_text(RING)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc, "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    import json
    return json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])


def test_strip_ring_exchange_two_ranks(tmp_path):
    """halo messages between strips: with two ranks both neighbours are the same peer (through the middle boundary and through the
    periodic seam) -- the pairing of untagged point-to-point operations must still be right"""
    assert _run_ring(tmp_path, 2, 29741) == {"ok": True, "world": 2}


def test_strip_ring_exchange_three_ranks(tmp_path):
    assert _run_ring(tmp_path, 3, 29743) == {"ok": True, "world": 3}
