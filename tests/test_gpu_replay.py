"""GPU (B200) parity tests proper: the CUDA sweep, called through the C ABI (libkmc_b200.so), against the
CPU oracle (oracle/libkmc_oracle.so, itself pinned bit-for-bit to the unmodified reference) on the same inputs
and the same keyed Philox stream. Bond formation/breakage, accept/reject decisions and complex member lists
must be identical; positions agree to 1e-12 relative (libm sin/cos/atan2/acos of CUDA vs glibc)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import kmc_b200
import pyoracle
from common import apply_regime, compare_states, load_golden_state, log_errors


def make_pair(na, nb, box, regime, seed, grid=1, **kw):
    po = apply_regime(pyoracle.default_params(box=box, n_receptor=na, n_ligand=nb, use_grid=grid, stream_mode=1, seed=seed), regime)
    pg = apply_regime(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=seed, mode=kmc_b200.MODE_REPLAY, **kw), regime)
    return pyoracle.Oracle(po), kmc_b200.Kmc(pg)


def lockstep(o, k, nsteps, check_every, label, per_step_accept=False):
    done, worst = 0, 0.0
    while done < nsteps:
        m = min(check_every, nsteps - done)
        if per_step_accept:
            for _ in range(m):
                o.step(1); k.step(1)
                assert np.array_equal(o.accepted()[1:], k.accepted()[1:]), "%s: accept/reject differs at step %d" % (label, done + 1)
                done += 1
        else:
            o.step(m); k.step(m); done += m
        worst = max(worst, compare_states(o.get_state(), k.get_state(), "%s step %d" % (label, done)))
        c, s = o.counts(), k.series()
        for key in ("bond_num", "bond_num_rl", "bond_num_cis", "bond_num_mono_cis", "max_complex"):
            assert c[key] == s[key], (label, done, key, c[key], s[key])
        assert c["tot_cluster_num"] == s["n_complexes"] and c["tot_proteins_in_cluster"] == s["n_in_complexes"]
        assert c["cluster_size"] == s["cluster_size"]
        assert o.results() == k.complexes(), "%s: complex member lists differ at step %d" % (label, done)
    # strict figures (absolute in Angstrom, relative to |x| itself) next to the floored one the assertion uses; absolute error bounded too
    a, r = log_errors(label, worst, o.get_state(), k.get_state())
    assert a <= 1e-9, "%s: absolute position error %.3e A" % (label, a)
    return worst


def test_default_system_free_diffusion():
    """configs[0] geometry: 150 receptors + 50 ligands, paper parameters; own scalable initial state."""
    o, k = make_pair(150, 50, (5773, 5773, 1000), "default", seed=11)
    k.init_random(seed=5)
    R, st, rn = k.get_state()
    o.set_state(R, st, rn)
    lockstep(o, k, 300, 50, "default", per_step_accept=True)


@pytest.mark.parametrize("name,regime", [("dense_step30000.npz", "dense"), ("hot200_step40000.npz", "hot")])
def test_from_reference_state_with_complexes(golden_dir, name, regime):
    """start from a state the REFERENCE evolved (63 / 46 bonds, multi-ligand complexes) and replay 1500 steps:
    rigid complex moves, lay-down, alignment passes, shuffles, association, dissociation."""
    g = load_golden_state(os.path.join(golden_dir, name))
    o, k = make_pair(150, 50, tuple(g["params"]["box"]), regime, seed=2024)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    lockstep(o, k, 200, 1, name, per_step_accept=True)
    lockstep(o, k, 1300, 100, name)
    ev_o, ev_k = o.events(), k.events()
    assert (ev_k["rl_on"], ev_k["mono_cis_on"], ev_k["cis_on"], ev_k["rl_off"], ev_k["mono_cis_off"], ev_k["cis_off"]) == tuple(int(x) for x in ev_o[:6])


def test_laydown_and_goto_back_edge_fire_inside_the_window(golden_dir):
    """S2e lay-down (main.cpp:1141-1189) and the `goto lable4` back edge of the multi-ligand alignment (main.cpp:1628 -> 1438) are
    rare events; this start state (tests/golden/make_goto_state.py) sits right before both. The oracle's own counters must advance
    inside the per-step lockstep window, so the GPU's S2e/S2f code is known to have executed them, decisions and member order equal."""
    g = load_golden_state(os.path.join(golden_dir, "hot200_step180900_goto.npz"))
    o, k = make_pair(150, 50, tuple(g["params"]["box"]), "hot", seed=g["params"]["keyed_seed"])
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    ev0 = o.events().copy()
    lockstep(o, k, 600, 1, "goto-window", per_step_accept=True)
    ev1 = o.events()
    assert ev1[8] - ev0[8] >= 1, "no lay-down inside the window"
    assert ev1[9] - ev0[9] >= 2, "goto lable4 not taken inside the window"


def test_hot_from_scratch_long():
    """dense hot system from a bond-free start, 20 000 steps: complexes form, grow, break."""
    o, k = make_pair(150, 50, (2500, 2500, 400), "hot", seed=7)
    k.init_random(seed=3)
    R, st, rn = k.get_state()
    o.set_state(R, st, rn)
    lockstep(o, k, 20000, 1000, "hot-long")
    assert k.series()["bond_num"] > 10


def test_small_system_hot40(golden_dir):
    g = load_golden_state(os.path.join(golden_dir, "hot40_step200000.npz"))
    o, k = make_pair(30, 10, tuple(g["params"]["box"]), "hot", seed=99)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    lockstep(o, k, 5000, 250, "hot40")


def test_replicas_are_independent_systems():
    """ensemble batching (configs[2]): replica r must equal a single system run with seed + r."""
    R_ = 5
    pg = apply_regime(kmc_b200.default_params(box=(2500, 2500, 400), seed=100, n_replicas=R_), "hot")
    k = kmc_b200.Kmc(pg)
    k.init_random(seed=40)
    oracles = []
    for r in range(R_):
        po = apply_regime(pyoracle.default_params(box=(2500, 2500, 400), use_grid=1, stream_mode=1, seed=100 + r), "hot")
        o = pyoracle.Oracle(po)
        o.set_state(*k.get_state(r))
        oracles.append(o)
    k.step(1500)
    for r, o in enumerate(oracles):
        o.step(1500)
        compare_states(o.get_state(), k.get_state(r), "replica %d" % r)
        assert o.results() == k.complexes(r)


@pytest.mark.parametrize("regime,steps,every", [("default", 1000, 250), ("hot", 200, 50)])
def test_config2_1e5_molecules_replay(regime, steps, every):
    """configs[1]: 1e5-molecule membrane (75 000 receptors + 25 000 ligands, default density), replay check vs the oracle:
    10^3 steps with the paper's parameters, 200 steps of the hot variant (association, dissociation, complexes fire)."""
    na, nb = 75000, 25000
    box = kmc_b200.scaled_box(na + nb)
    o, k = make_pair(na, nb, box, regime, seed=1)
    k.init_random(seed=1, sort_cells=True)
    o.set_state(*k.get_state())
    lockstep(o, k, steps, every, "1e5-" + regime)
    assert np.array_equal(o.accepted()[1:], k.accepted()[1:])
    ev_o, ev_k = o.events(), k.events()
    assert (ev_k["rl_on"], ev_k["mono_cis_on"], ev_k["cis_on"], ev_k["rl_off"], ev_k["mono_cis_off"], ev_k["cis_off"]) == tuple(int(x) for x in ev_o[:6])
    assert ev_k["rl_on"] > 0 and ev_k["reverted"] == int(ev_o[6])
    k.close()


def test_crowded_tiles_generic_path():
    """2000 molecules at 6x the default density: more entries per tile window than the shared-memory staging holds, so the tile
    kernel's overflow path (entries read from global memory, in-place evaluation) is what is being checked against the oracle"""
    na, nb = 1500, 500
    o, k = make_pair(na, nb, (7500.0, 7500.0, 400.0), "hot", seed=21)
    k.init_random(seed=8)
    o.set_state(*k.get_state())
    lockstep(o, k, 60, 1, "crowded", per_step_accept=True)
    lockstep(o, k, 540, 60, "crowded")
    assert k.series()["bond_num"] >= 1


def test_gpu_against_keyed_reference_directly(golden_dir):
    """No oracle in between: the bond table after 1000/2000/3000 steps must hash to what the UNMODIFIED reference produced when driven
    by the same keyed Philox stream (tests/golden/ref_kat.json: keyed_hot200, written by oracle/ref_harness.cpp --keyed)."""
    import json
    import refio
    kk = json.load(open(os.path.join(golden_dir, "ref_kat.json")))["keyed_hot200"]
    g = load_golden_state(os.path.join(golden_dir, kk["start"]))
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=tuple(g["params"]["box"]), seed=kk["seed"], mode=kmc_b200.MODE_REPLAY), "hot"))
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    done = g["step"]
    for fr in kk["frames"]:
        k.step(fr["step"] - done); done = fr["step"]
        R, st, rn = k.get_state()
        assert "%016x" % refio.fnv1a64(rn, st) == fr["hash_bonds"], "bond table differs from the keyed reference at step %d" % done
        s = k.series()
        assert (s["bond_num"], s["bond_num_rl"], s["bond_num_cis"], s["bond_num_mono_cis"]) == (fr["bond_num"], fr["bond_num_rl"], fr["bond_num_cis"], fr["bond_num_mono_cis"])


def test_asynchronous_snapshot_is_the_state_at_the_time_of_the_call():
    """kmc_get_packed_async: a device-side snapshot taken on the handle's stream crosses PCIe on a copy stream while the next steps
    already run. It must hold the state after the steps enqueued BEFORE the call -- not what the following steps make of it --, and
    a second snapshot queued behind the first must not disturb it. Membrane (general path) and ensemble (fused step)."""
    import torch
    for kw in (dict(box=kmc_b200.scaled_box(40000), n_receptor=30000, n_ligand=10000), dict(n_replicas=24)):
        a = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(seed=6, **kw), "hot"))
        b = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(seed=6, **kw), "hot"))
        a.init_random(seed=3, sort_cells=True)
        b.set_packed(*a.get_packed())
        pinned = lambda arrs: [torch.from_numpy(x.copy()).pin_memory().numpy() for x in arrs]
        out1, out2 = pinned(a.get_packed()), pinned(a.get_packed())
        a.step(60); a.get_packed_async(out1); a.step(45); a.get_packed_async(out2); a.step(10)
        a.snapshot_wait()
        b.step(60)
        for x, y in zip(out1, b.get_packed()):
            assert np.array_equal(x, y)
        b.step(45)
        for x, y in zip(out2, b.get_packed()):
            assert np.array_equal(x, y)
        b.step(10)
        for x, y in zip(a.get_packed(), b.get_packed()):
            assert np.array_equal(x, y)
        a.close(); b.close()


def test_errors_are_reported_not_swallowed():
    """bad input must come back as an error code with a message (the reference checks nothing)"""
    k = kmc_b200.Kmc(kmc_b200.default_params())
    k.init_random(seed=1)
    rec, lig, rl, rs, rc = k.get_packed()
    bad_rl, bad_rs = rl.copy(), rs.copy()
    bad_rl[0], bad_rs[0] = 3, 2
    bad_rl[1], bad_rs[1] = 3, 2                      # two receptors on the same ligand site
    with pytest.raises(kmc_b200.KmcError, match="R-L bond"):
        k.set_packed(rec, lig, bad_rl, bad_rs, rc)
    bad_rc = rc.copy(); bad_rc[5] = 7                # cis partner that does not point back
    with pytest.raises(kmc_b200.KmcError, match="cis bond"):
        k.set_packed(rec, lig, rl, rs, bad_rc)
    R, st, rn = k.get_state()
    R2 = R.copy(); R2[3, 2, 1, 0] += 1.0             # receptor beads no longer stacked
    with pytest.raises(kmc_b200.KmcError, match="stack"):
        k.set_state(R2, st, rn)
    rn2 = rn.copy(); rn2[1, 2] = 151                 # half a bond
    with pytest.raises(kmc_b200.KmcError, match="bond"):
        k.set_state(R, st, rn2)
    with pytest.raises(kmc_b200.KmcError):
        kmc_b200.Kmc(kmc_b200.default_params(n_replicas=0))
    with pytest.raises(kmc_b200.KmcError):
        kmc_b200.Kmc(kmc_b200.default_params(mode=7))
    k.set_packed(rec, lig, rl, rs, rc)               # and the handle is still usable afterwards
    k.step(3)
    assert k.series()["step"] == 3


@pytest.mark.parametrize("reuse,skin,drift", [("1", None, None), ("4", None, None), ("6", None, None), ("7", "3", "2.5"), ("3", "30", "0.5"), ("5", "1.5", "40")])
def test_list_reuse_settings_are_all_exact(golden_dir, monkeypatch, reuse, skin, drift):
    """The sparse path rebuilds the neighbour grid / pair list every KMC_REUSE-th step and reuses them in between; far movers
    (beyond KMC_SKIN) and molecules that drifted more than KMC_DRIFT from their grid entry are handled as special entries. Every
    setting must reproduce the oracle: tiny skins / drifts make most molecules special, large ones stretch the stale list."""
    monkeypatch.setenv("KMC_RESOLVE", "cells")
    monkeypatch.setenv("KMC_ADAPT", "0")           # stay on the requested setting however many special entries it makes
    monkeypatch.setenv("KMC_REUSE", reuse)
    if skin: monkeypatch.setenv("KMC_SKIN", skin)
    if drift: monkeypatch.setenv("KMC_DRIFT", drift)
    g = load_golden_state(os.path.join(golden_dir, "hot200_step40000.npz"))
    o, k = make_pair(150, 50, tuple(g["params"]["box"]), "hot", seed=77)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    lockstep(o, k, 150, 1, "reuse-%s" % reuse, per_step_accept=True)
    lockstep(o, k, 850, 50, "reuse-%s" % reuse)
    # a state pushed in from outside in the middle of a reuse window must force a rebuild
    R, st, rn = o.get_state()
    k.set_state(R, st, rn, step_done=o.counts()["step"], max_complex=o.counts()["max_complex"])
    lockstep(o, k, 60, 1, "reuse-%s after set_state" % reuse, per_step_accept=True)
    k.close()


def test_binary_checkpoint_continues_bit_for_bit(golden_dir, tmp_path):
    """kmc_write_checkpoint_bin / kmc_read_checkpoint_bin: a run restored from the binary checkpoint continues exactly
    like the uninterrupted one -- poses bit for bit, bonds, running-max complex, complex tables -- whereas the reference's own
    position.cpt keeps three decimals. Two replicas, hot regime (complexes form and break in the window)."""
    g = load_golden_state(os.path.join(golden_dir, "hot200_step40000.npz"))
    def make():
        k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=tuple(g["params"]["box"]), seed=31, n_replicas=2), "hot"))
        for r in range(2):
            k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"], replica=r)
        return k
    a = make()
    a.step(333)
    path = str(tmp_path / "state.kmcb")
    a.write_checkpoint_bin(path)
    a.step(400)
    b = make()
    b.read_checkpoint_bin(path)
    assert b.series(0)["step"] == g["step"] + 333
    b.step(400)
    for x, y in zip(a.get_packed(), b.get_packed()):
        assert np.array_equal(x, y)
    for r in range(2):
        assert a.series(r) == b.series(r)
        assert a.complexes(r) == b.complexes(r)
    other = kmc_b200.Kmc(kmc_b200.default_params(n_receptor=10, n_ligand=5))
    with pytest.raises(kmc_b200.KmcError, match="molecules"):
        other.read_checkpoint_bin(path)
    with pytest.raises(kmc_b200.KmcError):
        other.read_checkpoint_bin(str(tmp_path / "missing.kmcb"))


def test_list_reuse_backs_off_when_molecules_outrun_their_entries(golden_dir, monkeypatch):
    """With a 1.5 A skin most molecules are 'special entries' of every reuse step; at the next host synchronisation the library
    falls back to rebuilding the grid every step (with its default skin). The switch happens in the middle of a run and must not
    change a bit."""
    monkeypatch.setenv("KMC_RESOLVE", "cells"); monkeypatch.setenv("KMC_REUSE", "6"); monkeypatch.setenv("KMC_SKIN", "1.5")
    g = load_golden_state(os.path.join(golden_dir, "hot200_step40000.npz"))
    o, k = make_pair(150, 50, tuple(g["params"]["box"]), "hot", seed=12)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    o.step(9); k.step(9)
    assert k.events()["special_entries"] > 10
    k.sync()                                       # the back-off is decided here
    lockstep(o, k, 300, 1, "adapt", per_step_accept=True)
    assert k.events()["special_entries"] == 0      # every step rebuilds now: no special entries any more
    k.close()


def test_list_reuse_first_back_off_level_is_exact(golden_dir, monkeypatch):
    """Without tuning knobs in the environment the first back-off is a WIDER list (30 A skin, 60 A drift allowance, rebuilt every 4th
    step on coarser cells) rather than a rebuild every step. Met here on the 200-molecule hot state with the trigger lowered; the
    switch re-lays the neighbour grid in the middle of a run and must not change a bit."""
    monkeypatch.setenv("KMC_FUSED", "0"); monkeypatch.setenv("KMC_RESOLVE", "cells"); monkeypatch.setenv("KMC_ADAPT_MIN", "2")
    g = load_golden_state(os.path.join(golden_dir, "hot200_step40000.npz"))
    o, k = make_pair(150, 50, (4000.0, 4000.0, 400.0), "hot", seed=41)          # (the reference-evolved state in a wider box: the wide list is only taken where it stays short)
    assert k.path() == "general"
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    o.step(15); k.step(15)                         # (a reuse step: builds are steps 1, 7, 13)
    before = k.events()
    assert before["special_entries"] > 2
    k.sync()                                       # the back-off is decided here
    lockstep(o, k, 400, 1, "adapt-level-1", per_step_accept=True)
    after = k.events()
    assert after["list_pairs"] > 1.5 * before["list_pairs"]      # the wide list, not the every-step rebuild (whose steps have no special entries at all)
    lockstep(o, k, 1600, 100, "adapt-level-1")
    k.close()


def test_full_size_membrane_properties(monkeypatch):
    """BASELINE configs[3] size on one GPU (1e6 molecules, default density), where the oracle is too slow to follow: the
    size-independent properties the domain offers. (a) Exactness of the list reuse at full size: a handle that rebuilds its
    neighbour grid every step and one that reuses it for 6 steps end bit-identical. (b) What the sweep must conserve:
    rigid bodies keep their edge lengths, no committed overlap (receptor centres >= 2 rA, receptor-ligand beads >= rA + rB,
    ligand-ligand beads >= 2 rB; plain Euclidean distances, the reference has no minimum image), a symmetric bond table whose
    counters add up, ligand centres inside the slab up to one reflection."""
    from scipy.spatial import cKDTree
    na, nb = 750000, 250000
    box = kmc_b200.scaled_box(na + nb)
    runs = []
    for reuse in ("1", "6"):
        monkeypatch.setenv("KMC_REUSE", reuse)
        k = kmc_b200.Kmc(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=9))
        k.init_random(seed=4, sort_cells=True)
        k.step(26)
        k.sync()
        runs.append((k.get_packed(), k.series(), k.events()))
        k.close()
    (pa, sa, ea), (pb, sb, eb) = runs
    for x, y in zip(pa, pb):
        assert np.array_equal(x, y)
    assert sa == sb and ea["reverted"] == eb["reverted"] and ea["reverted"] > 0
    rec, lig, rl, rs, rc = pb
    p = kmc_b200.default_params()
    rA, rB = p.rA, p.rB
    c, s2, s3 = rec[:, 0:2], rec[:, 2:4], rec[:, 4:6]
    assert np.allclose(np.linalg.norm(s2 - c, axis=1), rA, rtol=0, atol=1e-9) and np.allclose(np.linalg.norm(s3 - c, axis=1), rA, rtol=0, atol=1e-9)
    L = lig.reshape(nb, 8, 3)
    rs_ = 2 * rB / np.sqrt(3.0)
    for q in (1, 2, 3):
        assert np.allclose(np.linalg.norm(L[:, q] - L[:, 0], axis=1), rs_, rtol=0, atol=1e-9)
        assert np.allclose(np.linalg.norm(L[:, 4 + q] - L[:, 0], axis=1), rs_ + rB, rtol=0, atol=1e-9)
    assert np.allclose(np.linalg.norm(L[:, 4] - L[:, 0], axis=1), rB, rtol=0, atol=1e-9)
    assert L[:, 0, 2].min() > -8.0 and L[:, 0, 2].max() < box[2] + 8.0
    # no committed overlap
    assert len(cKDTree(c).query_pairs(2 * rA - 1e-9)) == 0
    beads = L[:, 1:4].reshape(-1, 3)
    owner = np.repeat(np.arange(nb), 3)
    pairs = cKDTree(beads).query_pairs(2 * rB - 1e-9, output_type="ndarray")
    assert not len(pairs) or (owner[pairs[:, 0]] == owner[pairs[:, 1]]).all()            # only beads of one ligand are that close
    rec_beads = np.concatenate([np.column_stack([c, np.full(na, z)]) for z in (0.0, 2 * rA, 4 * rA, 6 * rA)])
    low = beads[beads[:, 2] < 6 * rA + rA + rB]                      # only beads this close to the membrane can touch a receptor
    assert len(low) > 1000
    assert cKDTree(rec_beads).query_ball_point(low, rA + rB - 1e-9, return_length=True).sum() == 0
    # bond table: symmetric, counters add up
    bound = np.nonzero(rl >= 0)[0]
    assert sb["bond_num_rl"] == len(bound) and sb["bond_num"] == sb["bond_num_rl"] + sb["bond_num_cis"] + sb["bond_num_mono_cis"]
    cis = np.nonzero(rc >= 0)[0]
    assert (rc[rc[cis]] == cis).all() and len(cis) == 2 * (sb["bond_num_cis"] + sb["bond_num_mono_cis"])
