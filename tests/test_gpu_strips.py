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
    return start, k.get_packed(), k.series()


@pytest.mark.parametrize("nranks,every,dev", [(2, 3, False), (4, 2, False), (3, 5, False), (2, 3, True), (4, 2, True), (3, 5, True)])
def test_strips_equal_single_gpu(nranks, every, dev):
    na, nb, box, regime, seed, steps = 15000, 5000, (26000.0, 26000.0, 400.0), "hot", 5, 60
    start, end, series = single_run(box, na, nb, regime, seed, steps)
    cap = lambda r: apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed), regime)   # capacity = whole system
    ls = LocalStrips(cap, nranks, every, halo_width=halo_for(every), device_refresh=dev)
    ls.load_global(*start)
    ls.step(steps)
    rec, lig, rl, rs, rc = ls.gather(na, nb)
    assert np.array_equal(rl, end[2]) and np.array_equal(rs, end[3]) and np.array_equal(rc, end[4]), "bond tables differ"
    assert np.array_equal(rec, end[0]) and np.array_equal(lig, end[1]), "positions differ (must be bit-identical: same device arithmetic)"
    assert series["bond_num"] >= 5          # bonds formed during the run (complexes are the subject of the next test)
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
