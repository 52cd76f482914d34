"""Strip decomposition: thin callers of the C ABI (one kmc handle per rank, csrc/kmc_strips.cu).

The refresh itself -- classify, pack, NCCL send/recv with the two x-neighbours, merge -- lives in the library
(kmc_strip_refresh, on the handle's stream, no host synchronisation; kmc_step calls it every refresh_every steps once
kmc_strip_comm_init has run). What is left here: DistStrips hands the NCCL unique id from rank 0 to the other ranks through
torch.distributed (any transport would do) and LocalStrips drives K logical ranks inside one process (one GPU: how the tests
prove equality with the single-GPU run). The host-path refresh (strip_begin_refresh / strip_message / strip_rebuild and the
ring_exchange of byte strings below) is the slow reference implementation the device path is tested against."""
import numpy as np

from . import Kmc

REC_DT = np.dtype([("ref", "<i4"), ("ligRef", "<i4"), ("site", "<i4"), ("cisRef", "<i4"), ("pose", "<f8", (6,))])
LIG_DT = np.dtype([("ref", "<i4"), ("recRef", "<i4", (3,)), ("pose", "<f8", (24,))])
assert REC_DT.itemsize == 64 and LIG_DT.itemsize == 208

# per-step reach of information: largest overlap reach (ligand-ligand centres, 2 rB + 2*2rB/sqrt3 = 129.3 A at default radii) plus
# twice the largest displacement of a molecule in one step (complex members swing up to ~31 A); rounded up generously
D1 = 200.0


def halo_for(refresh_every, complex_extent=400.0, params=None):
    """halo width that keeps the owned strip exact for `refresh_every` steps between refreshes (kmc_strip_halo_width when the
    parameters are given; the constant D1 covers the default radii)"""
    if params is not None:
        from . import strip_halo_width
        return strip_halo_width(params, refresh_every, complex_extent)
    return refresh_every * D1 + complex_extent


def msg_bytes(n_rec, n_lig):
    return n_rec * 64 + n_lig * 208


def parse_message(buf):
    if len(buf) < 16:
        return np.zeros(0, REC_DT), np.zeros(0, LIG_DT)
    n = np.frombuffer(buf, dtype="<i8", count=2)
    rec = np.frombuffer(buf, dtype=REC_DT, count=int(n[0]), offset=16)
    lig = np.frombuffer(buf, dtype=LIG_DT, count=int(n[1]), offset=16 + int(n[0]) * 64)
    return rec, lig


def assemble_global(messages, n_rec, n_lig):
    """owned sets of all ranks -> global packed state (rec[n_rec,6], lig[n_lig,24], rec_lig, rec_site, rec_cis as 0-based
    global indices / site 2..4, the kmc_get_packed conventions). Every molecule must be owned exactly once."""
    rec = np.full((n_rec, 6), np.nan); lig = np.full((n_lig, 24), np.nan)
    rl = np.full(n_rec, -1, np.int32); rs = np.zeros(n_rec, np.int32); rc = np.full(n_rec, -1, np.int32)
    seen_r = np.zeros(n_rec, np.int32); seen_l = np.zeros(n_lig, np.int32)
    for buf in messages:
        r, l = parse_message(buf)
        a = r["ref"] - 1
        rec[a] = r["pose"]; seen_r[a] += 1
        rl[a] = np.where(r["ligRef"] > 0, r["ligRef"] - n_rec - 1, -1); rs[a] = np.where(r["ligRef"] > 0, r["site"] + 2, 0)
        rc[a] = np.where(r["cisRef"] > 0, r["cisRef"] - 1, -1)
        b = l["ref"] - n_rec - 1
        lig[b] = l["pose"]; seen_l[b] += 1
    assert (seen_r == 1).all() and (seen_l == 1).all(), "ownership is not a partition: %d/%d receptors, %d/%d ligands owned once" % (
        (seen_r == 1).sum(), n_rec, (seen_l == 1).sum(), n_lig)
    return rec, lig, rl, rs, rc


def ring_exchange(torch, dist, device, to_low, to_high):
    """Every rank sends `to_low` to rank-1 and `to_high` to rank+1 (periodic) and returns (what rank-1 sent upwards, what rank+1
    sent downwards). Point-to-point operations (NCCL has no tags: operations between one pair of ranks are matched in posting
    order, so sends are posted [to_low, to_high] and receives [from the upper neighbour, from the lower neighbour]; with two
    ranks, where both neighbours are the same peer, that makes the peer's first receive meet my first send)."""
    rank, n = dist.get_rank(), dist.get_world_size()
    lo, hi = (rank - 1) % n, (rank + 1) % n

    def as_t(b):
        return torch.frombuffer(bytearray(b) if b else bytearray(1), dtype=torch.uint8).to(device)

    def run(ops):
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    sizes = torch.tensor([len(to_low), len(to_high)], dtype=torch.int64, device=device)
    sz_hi = torch.zeros(2, dtype=torch.int64, device=device); sz_lo = torch.zeros(2, dtype=torch.int64, device=device)
    run([dist.P2POp(dist.isend, sizes, lo), dist.P2POp(dist.isend, sizes.clone(), hi), dist.P2POp(dist.irecv, sz_hi, hi), dist.P2POp(dist.irecv, sz_lo, lo)])
    n_from_hi, n_from_lo = int(sz_hi[0].item()), int(sz_lo[1].item())      # the upper neighbour's to_low, the lower neighbour's to_high
    r_hi = torch.empty(max(n_from_hi, 1), dtype=torch.uint8, device=device); r_lo = torch.empty(max(n_from_lo, 1), dtype=torch.uint8, device=device)
    run([dist.P2POp(dist.isend, as_t(to_low), lo), dist.P2POp(dist.isend, as_t(to_high), hi), dist.P2POp(dist.irecv, r_hi, hi), dist.P2POp(dist.irecv, r_lo, lo)])
    return r_lo[:n_from_lo].cpu().numpy().tobytes(), r_hi[:n_from_hi].cpu().numpy().tobytes()


class StripRank:
    def __init__(self, params, rank, nranks, halo_width):
        self.k = Kmc(params)
        self.rank, self.nranks = rank, nranks
        self.k.strip_configure(rank, nranks, halo_width)

    def load_global(self, rec, lig, rl=None, rs=None, rc=None, step_done=0):
        self.k.strip_load_global(rec, lig, rl, rs, rc, step_done)


class LocalStrips:
    """K logical ranks in one process (all on one GPU): the in-process stand-in for the NCCL exchange.
    device_refresh=True keeps the refresh on the GPU (device-to-device copies instead of NCCL)."""

    def __init__(self, make_params, nranks, refresh_every, halo_width=None, device_refresh=False, guard=False):
        self.n, self.every, self.device_refresh, self.guard = nranks, refresh_every, device_refresh, guard
        self.halo = halo_width if halo_width is not None else halo_for(refresh_every)
        self.ranks = [StripRank(make_params(r), r, nranks, self.halo) for r in range(nranks)]
        self.since = 0

    def load_global(self, *a, **kw):
        for r in self.ranks:
            r.load_global(*a, **kw)

    def refresh(self):
        if self.device_refresh:
            return self._refresh_dev()
        for r in self.ranks:
            r.k.strip_begin_refresh()
        low = [r.k.strip_message(0) for r in self.ranks]; high = [r.k.strip_message(1) for r in self.ranks]
        for i, r in enumerate(self.ranks):
            if self.n == 1:
                r.k.strip_rebuild(b"", b"")
            else:   # what my lower neighbour sent upwards arrives as from_low; what my upper neighbour sent downwards as from_high
                r.k.strip_rebuild(high[(i - 1) % self.n], low[(i + 1) % self.n])
        self.since = 0

    def _refresh_dev(self):
        from . import strip_refresh_local
        strip_refresh_local([r.k for r in self.ranks], self.every if self.guard else 0)
        self.since = 0

    def series(self):
        """bond.dat row of the whole membrane: the ranks' owned-only parts added up (values of the last refresh)"""
        parts = [r.k.strip_series(reduce=False) for r in self.ranks]
        out = dict(parts[0])
        for k in ("bond_num_rl", "bond_num_mono_cis", "bond_num_cis", "bond_num", "n_complexes", "n_in_complexes"):
            out[k] = sum(p[k] for p in parts)
        out["max_complex"] = max(p["max_complex"] for p in parts)
        out["cluster_size"] = out["n_in_complexes"] / out["n_complexes"] if out["n_complexes"] else 0.0
        return out

    def oligomer_hist(self, nbins=64):
        return sum(r.k.strip_oligomer_hist(reduce=False, nbins=nbins) for r in self.ranks)

    def step(self, n):
        while n > 0:
            m = min(n, self.every - self.since)
            for r in self.ranks:
                r.k.step(m)
            self.since += m; n -= m
            if self.since == self.every:
                self.refresh()

    def gather(self, n_rec, n_lig):
        for r in self.ranks:
            r.k.strip_begin_refresh()
        return assemble_global([r.k.strip_message(2) for r in self.ranks], n_rec, n_lig)


class DistStrips:
    """one rank per process. Backend nccl: the library's own refresh (C++ + NCCL on the handle's stream, kmc_step refreshes by
    itself). Any other backend (gloo, CPU plumbing tests): the host-path refresh with the byte strings sent through torch."""

    def __init__(self, params, refresh_every, halo_width=None, dist=None, device=None, device_refresh=None):
        import torch
        self.device_refresh = (dist.get_backend() == "nccl") if device_refresh is None else device_refresh
        self.torch, self.dist = torch, dist
        self.rank, self.n = dist.get_rank(), dist.get_world_size()
        self.every = refresh_every
        self.halo = halo_width if halo_width is not None else halo_for(refresh_every)
        self.device = device if device is not None else ("cuda:%d" % params.device if dist.get_backend() == "nccl" else "cpu")
        self.sr = StripRank(params, self.rank, self.n, self.halo)
        self.k = self.sr.k
        self.since = 0
        self.bytes_sent = 0
        if self.device_refresh:
            from . import strip_unique_id
            box = [strip_unique_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            self.k.strip_comm_init(box[0], refresh_every)

    def load_global(self, *a, **kw):
        self.sr.load_global(*a, **kw)

    def _exchange(self, to_low, to_high):
        out = ring_exchange(self.torch, self.dist, self.device, to_low, to_high)
        self.bytes_sent += len(to_low) + len(to_high)
        return out

    def refresh(self):
        if self.device_refresh:
            self.k.strip_refresh(); self.since = 0
            return
        self.k.strip_begin_refresh()
        if self.n == 1:
            self.k.strip_rebuild(b"", b"")
        else:
            from_low, from_high = self._exchange(self.k.strip_message(0), self.k.strip_message(1))
            self.k.strip_rebuild(from_low, from_high)
        self.since = 0

    def step(self, n):
        if self.device_refresh:             # the library refreshes every `every` steps on its own
            self.k.step(n)
            return
        while n > 0:
            m = min(n, self.every - self.since)
            self.k.step(m)
            self.since += m; n -= m
            if self.since == self.every:
                self.refresh()

    def series(self):
        return self.k.strip_series(reduce=True)

    def oligomer_hist(self, nbins=64):
        return self.k.strip_oligomer_hist(reduce=True, nbins=nbins)

    def gather(self, n_rec, n_lig):
        """global state on every rank (all_gather of the owned sets; host path, for checks)"""
        self.k.strip_begin_refresh()
        mine = self.k.strip_message(2)
        out = [None] * self.n
        self.dist.all_gather_object(out, mine)
        return assemble_global(out, n_rec, n_lig)


def nccl_check(dist, local, molecules=100000, steps=64, every=8, regime_scale=20.0, seed=5, pre=1200):
    """ONE membrane on the live NCCL ranks == the same membrane on one GPU, bit for bit (collective; returns a dict on every rank).
    Hot, dense regime (on-rates x20, fast dissociation, 6.6x the default density) so that bonds form and break, complexes
    straddle boundaries and units migrate within the window. Every rank runs the single-GPU trajectory itself and compares the
    records of the units it owns after the strip run with it; the bond.dat row of the whole membrane (kmc_strip_get_series,
    all-reduced in the library) must equal the single-GPU row, the oligomer histogram likewise."""
    import torch
    from . import Kmc, default_params, strip_halo_width
    na, nb = (3 * molecules) // 4, molecules - (3 * molecules) // 4
    L = 26000.0 * (molecules / 20000.0) ** 0.5
    box = (L, L, 400.0)

    def mk():
        p = default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed, device=local)
        p.cis_on *= regime_scale; p.mono_cis_on *= regime_scale
        p.off, p.cis_off, p.mono_cis_off = 2e-5, 2e-5, 1e-4
        return p
    k = Kmc(mk()); k.init_random(seed=17, sort_cells=True)
    k.step(pre)                                         # complexes first (on one GPU), so that the strip window has them everywhere
    start = k.get_packed(); mx0 = k.series()["max_complex"]
    k.step(steps)
    end = k.get_packed(); want = k.series(); want_hist = k.oligomer_hist(); k.close()
    halo = strip_halo_width(mk(), every, 400.0)
    ds = DistStrips(mk(), every, halo_width=halo, dist=dist)
    ds.load_global(*start, step_done=pre)
    ds.step(steps)
    buf = np.zeros(64 * na + 208 * nb, dtype=np.uint8)
    nr, nl = ds.k.strip_get_records(2, buf)
    rec = np.frombuffer(buf, dtype=REC_DT, count=nr); lig = np.frombuffer(buf, dtype=LIG_DT, count=nl, offset=64 * nr)
    a, b = rec["ref"] - 1, lig["ref"] - na - 1
    ok = bool(np.array_equal(rec["pose"], end[0][a]) and np.array_equal(lig["pose"], end[1][b]) and
              np.array_equal(np.where(rec["ligRef"] > 0, rec["ligRef"] - na - 1, -1), end[2][a]) and
              np.array_equal(np.where(rec["ligRef"] > 0, rec["site"] + 2, 0), end[3][a]) and
              np.array_equal(np.where(rec["cisRef"] > 0, rec["cisRef"] - 1, -1), end[4][a]))
    got = ds.series(); got_hist = ds.oligomer_hist()
    # (the running-max complex of the strip run starts at the load: compare it only if the window's own maximum reaches the earlier one)
    keys = ("bond_num", "bond_num_rl", "bond_num_cis", "bond_num_mono_cis", "n_complexes", "n_in_complexes", "cluster_size")
    series_ok = all(got[q] == want[q] for q in keys) and bool(np.array_equal(got_hist, want_hist)) and got["max_complex"] <= want["max_complex"] and \
        (got["max_complex"] == want["max_complex"] or want["max_complex"] == mx0)
    ds.k.sync()
    t = torch.tensor([1 if ok else 0, 1 if series_ok else 0, nr + nl, -(nr + nl)], dtype=torch.int64, device="cuda:%d" % local)
    tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    refreshes = steps // every
    ds.k.close()
    return {"equal_single_gpu": bool(tmin[0].item() == 1 and tsum[2].item() == na + nb), "series_equal": bool(tmin[1].item() == 1),
            "ranks": dist.get_world_size(), "molecules": molecules, "steps": steps, "refreshes": refreshes, "refresh_every": every, "halo": halo,
            "bonds": want["bond_num"], "complexes": want["n_complexes"], "max_complex": want["max_complex"],
            "owned_min": int(-tmin[3].item()), "owned_total": int(tsum[2].item()), "backend": dist.get_backend(),
            "exchange": "kmc_strip_refresh: C++ ncclSend/ncclRecv on the handle's stream"}
