"""GPU: strip decomposition of one membrane. K logical ranks on ONE GPU (in-process exchange instead of NCCL) must reproduce
the single-handle run bit for bit: same positions, same bond table -- periodic seam, migrating units, complexes that straddle
boundaries and refreshes every few steps included."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import kmc_b200
from kmc_b200.strips import LocalStrips, halo_for
from common import apply_regime


def single_run(box, na, nb, regime, seed, steps):
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), regime))
    k.init_random(seed=17, sort_cells=True)
    start = k.get_packed()
    k.step(steps)
    return start, k.get_packed(), k.series(), k.oligomer_hist()


@pytest.mark.parametrize("nranks,every,dev", [(2, 3, False), (4, 2, False), (3, 5, False), (2, 3, True), (4, 2, True), (3, 5, True)])
def test_strips_equal_single_gpu(nranks, every, dev):
    na, nb, box, regime, seed, steps = 15000, 5000, (26000.0, 26000.0, 400.0), "hot", 5, 60
    start, end, series, hist = single_run(box, na, nb, regime, seed, steps)
    cap = lambda r: apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), regime)   # capacity = whole system
    ls = LocalStrips(cap, nranks, every, halo_width=halo_for(every), device_refresh=dev)
    ls.load_global(*start)
    ls.step(steps)
    rec, lig, rl, rs, rc = ls.gather(na, nb)
    assert np.array_equal(rl, end[2]) and np.array_equal(rs, end[3]) and np.array_equal(rc, end[4]), "bond tables differ"
    assert np.array_equal(rec, end[0]) and np.array_equal(lig, end[1]), "positions differ (must be bit-identical: same device arithmetic)"
    assert series["bond_num"] >= 5          # bonds formed during the run (complexes are the subject of the next test)
    if dev:
        # the bond.dat row of the WHOLE membrane from the ranks' owned-only parts (main.cpp:2247-2253): equal to the single-GPU row
        got = ls.series()
        for key in ("bond_num", "bond_num_rl", "bond_num_cis", "bond_num_mono_cis", "max_complex", "n_complexes", "n_in_complexes", "cluster_size"):
            assert got[key] == series[key], (key, got[key], series[key])
        assert np.array_equal(ls.oligomer_hist(), hist)
    # the decomposition is real: no rank owns everything
    owned = [len(kmc_b200.strips.parse_message(r.k.strip_message(2))[0]) for r in ls.ranks]
    assert sum(owned) == na and max(owned) < 0.8 * na


@pytest.mark.parametrize("dev", [False, True])
def test_strips_with_complexes_across_boundaries(dev):
    """start from a state that already has complexes everywhere (evolved on one GPU), then continue on 4 strips"""
    na, nb, box, regime, seed = 6000, 2000, (12000.0, 12000.0, 400.0), "hot", 9
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), regime))
    k.init_random(seed=3, sort_cells=True)
    k.step(1500)
    mid = k.get_packed()
    assert k.series()["bond_num"] > 80
    k.step(40)
    end = k.get_packed()
    ls = LocalStrips(lambda r: apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), regime), 4, 2, halo_width=1000.0, device_refresh=dev)
    ls.load_global(*mid, step_done=1500)
    ls.step(40)
    rec, lig, rl, rs, rc = ls.gather(na, nb)
    assert np.array_equal(rl, end[2]) and np.array_equal(rc, end[4])
    assert np.array_equal(rec, end[0]) and np.array_equal(lig, end[1])
    if dev:
        got, want = ls.series(), k.series()
        for key in ("bond_num", "bond_num_rl", "bond_num_cis", "bond_num_mono_cis", "n_complexes", "n_in_complexes", "cluster_size"):
            assert got[key] == want[key], (key, got[key], want[key])
        assert np.array_equal(ls.oligomer_hist(), k.oligomer_hist())


def oligomerised_state(na=6000, nb=2000, box=(12000.0, 12000.0, 400.0), seed=9, steps=1500):
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), "hot"))
    k.init_random(seed=3, sort_cells=True)
    k.step(steps)
    return k


def test_two_ranks_units_inside_both_bands():
    """nranks = 2: both neighbours of a rank are the same peer, so a unit that lies inside BOTH bands of its owner must still arrive
    once (k_strip_classify lists it in one message only). Halo nearly half a strip wide: most units are in both bands."""
    na, nb, box, seed = 6000, 2000, (12000.0, 12000.0, 400.0), 9
    k = oligomerised_state(na, nb, box, seed)
    mid = k.get_packed()
    k.step(24)
    end = k.get_packed()
    mk = lambda r: apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), "hot")
    for dev in (False, True):
        ls = LocalStrips(mk, 2, 3, halo_width=2950.0, device_refresh=dev)
        ls.load_global(*mid, step_done=1500)
        ls.step(24)
        rec, lig, rl, rs, rc = ls.gather(na, nb)
        assert np.array_equal(rl, end[2]) and np.array_equal(rc, end[4]), "bond tables differ (device refresh %s)" % dev
        assert np.array_equal(rec, end[0]) and np.array_equal(lig, end[1]), "positions differ (device refresh %s)" % dev


def test_complex_wider_than_the_halo_budget_is_refused():
    """Exactness of a strip run needs halo >= refresh_every * reach + the x-extent of the widest unit. The classify kernel
    measures every owned unit; an oligomer wider than the budget makes the run fail with KMC_ERR_CAPACITY instead of silently
    diverging from the single-GPU trajectory."""
    na, nb, box, seed = 6000, 2000, (12000.0, 12000.0, 400.0), 9
    k = oligomerised_state(na, nb, box, seed)
    mid = k.get_packed()
    assert k.series()["max_complex"] >= 4
    mk = lambda r: apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), "hot")
    every = 2
    d1 = halo_for(1, 0.0, params=mk(0))
    ok = LocalStrips(mk, 4, every, halo_width=every * d1 + 1000.0, device_refresh=True, guard=True)      # generous budget: runs
    ok.load_global(*mid, step_done=1500); ok.step(4)
    for r in ok.ranks:
        r.k.sync()
    tight = LocalStrips(mk, 4, every, halo_width=every * d1 + 40.0, device_refresh=True, guard=True)   # 40 A: narrower than any bound pair
    tight.load_global(*mid, step_done=1500)
    with pytest.raises(kmc_b200.KmcError, match="wider than the halo budget"):
        tight.step(4)
        for r in tight.ranks:
            r.k.sync()


def test_records_round_trip_restores_a_rank():
    """kmc_strip_get_records(3) -> kmc_strip_load_records: the per-rank slab (owned + halo copies) through HOST buffers; the restored
    ranks continue bit for bit like the ones that never left the device (the e2e path of bench.py --gpus N)."""
    na, nb, box, seed = 6000, 2000, (12000.0, 12000.0, 400.0), 9
    k = oligomerised_state(na, nb, box, seed, steps=600)
    mid = k.get_packed()
    mk = lambda r: apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), "hot")
    a = LocalStrips(mk, 3, 4, halo_width=halo_for(4), device_refresh=True)
    a.load_global(*mid, step_done=600)
    a.step(8)
    b = LocalStrips(mk, 3, 4, halo_width=halo_for(4), device_refresh=True)
    for ra, rb in zip(a.ranks, b.ranks):
        buf = np.zeros(64 * na + 208 * nb, dtype=np.uint8)
        nr, nl = ra.k.strip_get_records(3, buf)
        assert 0 < nr < na and 0 < nl < nb
        rb.k.strip_load_records(buf, nr, nl, step_done=608)
    a.step(12); b.step(12)
    ga, gb = a.gather(na, nb), b.gather(na, nb)
    for x, y in zip(ga, gb):
        assert np.array_equal(x, y)
    # owned sets through the same call: a partition of the membrane
    tot = 0
    for r in b.ranks:
        buf = np.zeros(64 * na + 208 * nb, dtype=np.uint8)
        nr, nl = r.k.strip_get_records(2, buf)
        tot += nr + nl
    assert tot == na + nb
