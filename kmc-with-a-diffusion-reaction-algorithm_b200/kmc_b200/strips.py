"""Strip decomposition driver: one membrane across several GPUs (one kmc handle per rank, csrc/kmc_strips.cu).

StripRank wraps one rank's handle; the exchange of the boundary bands is done here, either between handles living in one
process (LocalStrips: K logical ranks on one GPU, used to prove equality with the single-GPU run) or between processes with
torch.distributed point-to-point operations (DistStrips: NCCL send/recv over NVLink when the backend is nccl; gloo on CPU for
the message plumbing tests)."""
import numpy as np

from . import Kmc

REC_DT = np.dtype([("ref", "<i4"), ("ligRef", "<i4"), ("site", "<i4"), ("cisRef", "<i4"), ("pose", "<f8", (6,))])
LIG_DT = np.dtype([("ref", "<i4"), ("recRef", "<i4", (3,)), ("pose", "<f8", (24,))])
assert REC_DT.itemsize == 64 and LIG_DT.itemsize == 208

# per-step reach of information: largest overlap reach (ligand-ligand centres, 2 rB + 2*2rB/sqrt3 = 129.3 A at default radii) plus
# twice the largest displacement of a molecule in one step (complex members swing up to ~31 A); rounded up generously
D1 = 200.0


def halo_for(refresh_every, complex_extent=400.0):
    """halo width that keeps the owned strip exact for `refresh_every` steps between refreshes"""
    return refresh_every * D1 + complex_extent


class _DevBuf:
    """a raw device allocation of the library seen as a torch uint8 tensor (zero copy, __cuda_array_interface__)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (max(int(nbytes), 1),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def dev_tensor(torch, ptr, nbytes, device):
    return torch.as_tensor(_DevBuf(ptr, nbytes), device=device)[:nbytes]


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

    def __init__(self, make_params, nranks, refresh_every, halo_width=None, device_refresh=False):
        self.n, self.every, self.device_refresh = nranks, refresh_every, device_refresh
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
        import torch
        dev = "cuda:%d" % self.ranks[0].k.p.device
        for r in self.ranks:
            r.k.strip_begin_refresh_dev()
        low = [r.k.strip_message_dev(0) for r in self.ranks]; high = [r.k.strip_message_dev(1) for r in self.ranks]
        for i, r in enumerate(self.ranks):
            if self.n == 1:
                r.k.strip_rebuild_dev(0, 0, 0, 0); continue
            src = (high[(i - 1) % self.n], low[(i + 1) % self.n])          # (from lower-x neighbour, from higher-x neighbour)
            for side, (ptr, nr, nl) in enumerate(src):
                dst = r.k.strip_recv_dev(side, nr, nl)
                nb = msg_bytes(nr, nl)
                if nb:
                    dev_tensor(torch, dst, nb, dev).copy_(dev_tensor(torch, ptr, nb, dev))
            torch.cuda.synchronize()
            r.k.strip_rebuild_dev(src[0][1], src[0][2], src[1][1], src[1][2])
        self.since = 0

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
    """one rank per process; boundary bands travel with torch.distributed send/recv (NCCL over NVLink with backend nccl)"""

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

    def load_global(self, *a, **kw):
        self.sr.load_global(*a, **kw)

    def _exchange(self, to_low, to_high):
        out = ring_exchange(self.torch, self.dist, self.device, to_low, to_high)
        self.bytes_sent += len(to_low) + len(to_high)
        return out

    def _refresh_dev(self):
        """device messages, NCCL point-to-point GPU to GPU; same pairing rules as ring_exchange"""
        torch, dist, dev = self.torch, self.dist, self.device
        k = self.k
        import os, time
        timing = os.environ.get("KMC_STRIP_TIMING")
        if timing:
            k.sync(); t0 = time.perf_counter()
        k.strip_begin_refresh_dev()
        if timing:
            t1 = time.perf_counter()
        if self.n == 1:
            k.strip_rebuild_dev(0, 0, 0, 0); self.since = 0; return
        lo, hi = (self.rank - 1) % self.n, (self.rank + 1) % self.n
        (pl, rl_, ll_), (ph, rh_, lh_) = k.strip_message_dev(0), k.strip_message_dev(1)

        def run(ops):
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        counts = torch.tensor([rl_, ll_, rh_, lh_], dtype=torch.int64, device=dev)
        c_hi = torch.zeros(4, dtype=torch.int64, device=dev); c_lo = torch.zeros(4, dtype=torch.int64, device=dev)
        run([dist.P2POp(dist.isend, counts, lo), dist.P2POp(dist.isend, counts.clone(), hi), dist.P2POp(dist.irecv, c_hi, hi), dist.P2POp(dist.irecv, c_lo, lo)])
        c_hi, c_lo = c_hi.tolist(), c_lo.tolist()
        fl = (c_lo[2], c_lo[3])          # from the lower neighbour: its message towards higher x
        fh = (c_hi[0], c_hi[1])          # from the higher neighbour: its message towards lower x
        t_send_lo = dev_tensor(torch, pl, max(msg_bytes(rl_, ll_), 1), dev); t_send_hi = dev_tensor(torch, ph, max(msg_bytes(rh_, lh_), 1), dev)
        t_recv_hi = dev_tensor(torch, k.strip_recv_dev(1, *fh), max(msg_bytes(*fh), 1), dev); t_recv_lo = dev_tensor(torch, k.strip_recv_dev(0, *fl), max(msg_bytes(*fl), 1), dev)
        run([dist.P2POp(dist.isend, t_send_lo, lo), dist.P2POp(dist.isend, t_send_hi, hi), dist.P2POp(dist.irecv, t_recv_hi, hi), dist.P2POp(dist.irecv, t_recv_lo, lo)])
        torch.cuda.synchronize()
        if timing:
            t2 = time.perf_counter()
        self.bytes_sent += msg_bytes(rl_, ll_) + msg_bytes(rh_, lh_)
        k.strip_rebuild_dev(fl[0], fl[1], fh[0], fh[1])
        if timing:
            t3 = time.perf_counter()
            self.timing = getattr(self, "timing", [0.0, 0.0, 0.0, 0])
            self.timing[0] += t1 - t0; self.timing[1] += t2 - t1; self.timing[2] += t3 - t2; self.timing[3] += 1
        self.since = 0

    def refresh(self):
        if self.device_refresh:
            return self._refresh_dev()
        self.k.strip_begin_refresh()
        if self.n == 1:
            self.k.strip_rebuild(b"", b"")
        else:
            from_low, from_high = self._exchange(self.k.strip_message(0), self.k.strip_message(1))
            self.k.strip_rebuild(from_low, from_high)
        self.since = 0

    def step(self, n):
        while n > 0:
            m = min(n, self.every - self.since)
            self.k.step(m)
            self.since += m; n -= m
            if self.since == self.every:
                self.refresh()

    def gather(self, n_rec, n_lig):
        """global state on every rank (all_gather of the owned sets)"""
        self.k.strip_begin_refresh()
        mine = self.k.strip_message(2)
        out = [None] * self.n
        self.dist.all_gather_object(out, mine)
        return assemble_global(out, n_rec, n_lig)
