"""The fused small-system step (csrc/kmc_small.cu: one CTA per replica, the whole time step main.cpp:461-2202 in one kernel,
many steps per launch; BASELINE configs[2]) against the general multi-kernel path of the same library and against the oracle.
The fused kernel runs the same device functions through a per-replica view of the state, so the two paths must agree BIT FOR BIT
(poses included: same arithmetic, same libdevice), whatever the chunking of the steps into launches."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import kmc_b200
import pyoracle
from common import apply_regime, compare_states, load_golden_state


def _pair_of_paths(monkeypatch, params_fn):
    monkeypatch.delenv("KMC_FUSED", raising=False)
    fused = kmc_b200.Kmc(params_fn())
    monkeypatch.setenv("KMC_FUSED", "0")
    general = kmc_b200.Kmc(params_fn())
    monkeypatch.delenv("KMC_FUSED", raising=False)
    assert fused.path() == "fused" and general.path() == "general"
    return fused, general


def _same(a, b, replicas):
    for x, y in zip(a.get_packed(), b.get_packed()):
        assert np.array_equal(x, y)
    for r in range(replicas):
        assert a.series(r) == b.series(r)
        assert a.complexes(r) == b.complexes(r)
        assert np.array_equal(a.accepted(r), b.accepted(r))
    ea, eb = a.events(), b.events()
    for key in ("rl_on", "mono_cis_on", "cis_on", "rl_off", "mono_cis_off", "cis_off", "reverted"):
        assert ea[key] == eb[key], key


@pytest.mark.parametrize("mode", [kmc_b200.MODE_REPLAY, kmc_b200.MODE_PRODUCTION])
def test_fused_equals_general_path_bit_for_bit(golden_dir, monkeypatch, mode):
    """4 replicas of a reference-evolved hot state (multi-ligand complexes, association, dissociation), odd and even chunks"""
    g = load_golden_state(os.path.join(golden_dir, "hot200_step40000.npz"))
    R_ = 4
    fused, general = _pair_of_paths(monkeypatch, lambda: apply_regime(kmc_b200.default_params(box=tuple(g["params"]["box"]), seed=17, n_replicas=R_, mode=mode), "hot"))
    for k in (fused, general):
        for r in range(R_):
            k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"], replica=r)
    for chunk in (1, 1, 7, 64, 333, 1000, 1):
        fused.step(chunk); general.step(chunk)
        _same(fused, general, R_)
    ev = fused.events()
    assert ev["rl_on"] > 0 and ev["rl_off"] > 0 and ev["cis_on"] + ev["mono_cis_on"] > 0 and ev["reverted"] > 0
    # replicas with different seeds diverge: they are independent systems
    assert not np.array_equal(fused.get_state(0)[0], fused.get_state(1)[0])


def test_fused_from_random_start_and_long_launch(monkeypatch):
    """bond-free random start (the ensemble benchmark's state), one launch of 5000 steps against 5000 graph launches"""
    fused, general = _pair_of_paths(monkeypatch, lambda: kmc_b200.default_params(seed=5, n_replicas=3))
    fused.init_random(seed=9)
    general.set_packed(*fused.get_packed())
    fused.step(5000); general.step(5000)
    _same(fused, general, 3)
    assert fused.series(0)["step"] == 5000


def test_fused_ticket_queue_for_large_ensembles(monkeypatch):
    """More replicas than CTAs fit on the device: a persistent grid deals the replicas in 64-step chunks through a ticket queue.
    Forced here with a grid of 3 CTAs for 7 replicas; odd launch lengths, launches shorter and longer than a chunk."""
    monkeypatch.setenv("KMC_SMALL_GRID", "3")
    fused, general = _pair_of_paths(monkeypatch, lambda: apply_regime(kmc_b200.default_params(box=(2500.0, 2500.0, 400.0), seed=23, n_replicas=7), "hot"))
    monkeypatch.delenv("KMC_SMALL_GRID")
    fused.init_random(seed=4)
    general.set_packed(*fused.get_packed())
    for chunk in (5, 64, 333, 1, 1000, 129):
        fused.step(chunk); general.step(chunk)
        _same(fused, general, 7)
    assert fused.series(3)["bond_num"] > 0 and fused.series(0)["step"] == 5 + 64 + 333 + 1 + 1000 + 129


@pytest.mark.parametrize("grid", [None, "1"])
def test_fused_four_replicas_per_cta_in_lockstep(monkeypatch, grid):
    """Ensembles that fill the device advance four replicas per CTA in lockstep (k_small_step<4>). Forced here on 7 replicas: two
    groups, the second one with a spare slot; with a grid of one CTA the groups go through the ticket queue as well."""
    monkeypatch.setenv("KMC_SMALL_SLOTS", "4")
    if grid: monkeypatch.setenv("KMC_SMALL_GRID", grid)
    fused, general = _pair_of_paths(monkeypatch, lambda: apply_regime(kmc_b200.default_params(box=(2500.0, 2500.0, 400.0), seed=29, n_replicas=7), "hot"))
    monkeypatch.delenv("KMC_SMALL_SLOTS"); monkeypatch.delenv("KMC_SMALL_GRID", raising=False)
    fused.init_random(seed=6)
    general.set_packed(*fused.get_packed())
    for chunk in (3, 64, 500, 1, 1200):
        fused.step(chunk); general.step(chunk)
        _same(fused, general, 7)
    assert fused.series(6)["bond_num"] > 0 and fused.events()["reverted"] > 0


def test_four_slot_kernel_against_oracle_with_spare_slots(golden_dir, monkeypatch):
    """the lockstep kernel on ONE system (three spare slots that only keep the barriers company), per-step against the oracle from the
    reference-evolved state that sits before a lay-down and `goto lable4` back edges (main.cpp:1141-1189, 1628 -> 1438)"""
    monkeypatch.setenv("KMC_SMALL_SLOTS", "4")
    g = load_golden_state(os.path.join(golden_dir, "hot200_step180900_goto.npz"))
    o = pyoracle.Oracle(apply_regime(pyoracle.default_params(box=tuple(g["params"]["box"]), use_grid=1, stream_mode=1, seed=g["params"]["keyed_seed"]), "hot"))
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=tuple(g["params"]["box"]), seed=g["params"]["keyed_seed"]), "hot"))
    assert k.path() == "fused"
    for x in (o, k):
        x.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    ev0 = o.events().copy()
    for step in range(300):
        o.step(1); k.step(1)
        assert np.array_equal(o.accepted()[1:], k.accepted()[1:]), step
    compare_states(o.get_state(), k.get_state(), "four-slot goto window")
    assert o.results() == k.complexes()
    assert o.events()[9] - ev0[9] >= 1, "goto lable4 not taken inside the window"


def test_fused_small_odd_sizes_against_oracle():
    """odd molecule count (the all-pairs schedule differs for odd and even N), more molecules than threads, crowded box"""
    for na, nb, box in ((31, 10, (900.0, 900.0, 300.0)), (170, 71, (3000.0, 3000.0, 400.0))):
        po = apply_regime(pyoracle.default_params(box=box, n_receptor=na, n_ligand=nb, use_grid=1, stream_mode=1, seed=3), "hot")
        pg = apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=3), "hot")
        o, k = pyoracle.Oracle(po), kmc_b200.Kmc(pg)
        assert k.path() == "fused"
        k.init_random(seed=2)
        o.set_state(*k.get_state())
        for _ in range(6):
            o.step(250); k.step(250)
            compare_states(o.get_state(), k.get_state(), "odd %d+%d" % (na, nb))
            assert o.results() == k.complexes()
            assert np.array_equal(o.accepted()[1:], k.accepted()[1:])
        assert k.series()["bond_num"] > 0
        k.close()


def test_fused_reports_capacity_overflow(monkeypatch):
    """a work list too small for the step must come back as KMC_ERR_CAPACITY from the fused path too"""
    monkeypatch.setenv("KMC_TEST_PENDCAP", "1")
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=(1200.0, 1200.0, 300.0), seed=1), "hot"))
    assert k.path() == "fused"
    k.init_random(seed=1)
    with pytest.raises(kmc_b200.KmcError, match="overflow"):
        for _ in range(40):
            k.step(50); k.sync()
