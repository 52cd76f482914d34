"""Host-side logic of the multi-GPU path (one process per GPU): which replicas / which patch a rank owns, how seeds are
derived, and how per-rank results are combined. No data-path collective exists (replicas and patches are independent,
DESIGN.md section 7); torch.distributed is used only for the barrier, the max-over-ranks of the device time and the final gather
of the time series. Works with the gloo backend on CPU (tests/test_multirank_cpu.py) and nccl on the GPU box."""


def replica_range(rank, world, n_replicas):
    """contiguous block partition of replica ids 0..n_replicas-1 over ranks (sizes differ by at most one)"""
    base, extra = divmod(n_replicas, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_seed(base_seed, rank, world, n_replicas=None):
    """Philox key of a rank: replica r of the whole ensemble always gets base_seed + r no matter how many GPUs share the
    work (so an ensemble result is independent of the GPU count); independent patches use base_seed + 1000*rank."""
    if n_replicas is None:
        return base_seed + 1000 * rank
    return base_seed + replica_range(rank, world, n_replicas)[0]


def max_over_ranks(value, dist=None, device=None):
    """max of a python float over ranks (the timing rule: device time, max over ranks)"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_series(local_rows, dist=None):
    """local_rows: list of (replica_id, dict) owned by this rank -> full list ordered by replica id on every rank"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(local_rows, key=lambda r: r[0])
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local_rows)
    return sorted([r for part in out for r in part], key=lambda r: r[0])
