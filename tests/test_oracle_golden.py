"""CPU: the oracle (oracle/kmc_oracle.cpp) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py -> ref_kat.json), and, when oracle/_ref is present, against the reference run live."""
import json
import os

import numpy as np
import pytest

import pyoracle
import refio
from common import apply_regime

KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_kat.json")))


def hashes(o):
    R, st, rn = o.get_state()
    return "%016x" % refio.fnv1a64(R[..., 0], R[..., 1], R[..., 2]), "%016x" % refio.fnv1a64(rn, st)


def check_against_frames(o, frames, upto=None):
    done = 0
    for fr in frames:
        if upto is not None and fr["step"] > upto:
            break
        o.step(fr["step"] - done)
        done = fr["step"]
        c = o.counts()
        hr, hb = hashes(o)
        assert hb == fr["hash_bonds"], "bond table differs from the reference at step %d" % done
        assert hr == fr["hash_R"], "coordinates differ from the reference at step %d" % done
        for k in ("bond_num", "bond_num_rl", "bond_num_cis", "bond_num_mono_cis"):
            assert c[k] == fr[k]
    return done


@pytest.mark.parametrize("grid", [0, 1])
def test_default_1000_steps(grid):
    """SURVEY 8c second known-answer vector: default parameters, 750 755 draws, R_x[1][1][1] = -151.64284743143637."""
    o = pyoracle.Oracle(pyoracle.default_params(use_grid=grid))
    o.init_reference()
    check_against_frames(o, KAT["default_1000"]["frames"])
    R, _, _ = o.get_state()
    assert R[1, 1, 1, 0] == -151.64284743143637 and R[200, 4, 2, 0] == 1961.758352843294
    assert o.counts()["rand2_draws"] == KAT["default_1000"]["summary"]["rand2_draws"] == 750755


def test_dense_10000_steps():
    """dense oligomerising regime (SURVEY 8c first vector), first two golden frames."""
    p = apply_regime(pyoracle.default_params(box=(2500, 2500, 400), use_grid=1), "dense")
    o = pyoracle.Oracle(p)
    o.init_reference()
    assert check_against_frames(o, KAT["dense_30000"]["frames"], upto=10000) == 10000


@pytest.mark.slow
def test_dense_30000_steps_kat():
    """the full first known-answer vector: 20 087 886 rand2 draws, 1 012 220 rand draws, 63 bonds,
    hashes 4c33ca174ef7e3d9 / 55c51a0c3e2c74d7."""
    p = apply_regime(pyoracle.default_params(box=(2500, 2500, 400), use_grid=1), "dense")
    o = pyoracle.Oracle(p)
    o.init_reference()
    check_against_frames(o, KAT["dense_30000"]["frames"])
    c = o.counts()
    assert (c["rand2_draws"], c["rand_draws"]) == (20087886, 1012220)
    assert hashes(o) == ("4c33ca174ef7e3d9", "55c51a0c3e2c74d7")


def test_hot40_200000_steps():
    """N=40, hot off-rates: dissociation (all three kinds), lay-downs and the `goto lable4` back edge all fire."""
    p = apply_regime(pyoracle.default_params(box=(1000, 1000, 300), n_receptor=30, n_ligand=10, use_grid=1), "hot")
    o = pyoracle.Oracle(p)
    o.init_reference()
    check_against_frames(o, KAT["hot40_200000"]["frames"])
    ev = o.events()
    assert ev[3] > 0 and ev[4] > 0 and ev[5] > 0 and ev[8] > 0 and ev[9] > 0
    c = o.counts()
    s = KAT["hot40_200000"]["summary"]
    assert (c["rand2_draws"], c["rand_draws"]) == (s["rand2_draws"], s["rand_draws"])


def test_grid_equals_all_pairs():
    """the O(N) neighbour grid of the oracle must not change a single bit relative to the reference-style loops."""
    outs = []
    for grid in (0, 1):
        p = apply_regime(pyoracle.default_params(box=(2500, 2500, 400), use_grid=grid), "hot")
        o = pyoracle.Oracle(p)
        o.init_reference()
        o.step(3000)
        outs.append((hashes(o), o.counts()["rand2_draws"]))
    assert outs[0] == outs[1]


def test_restart_from_golden_state(golden_dir):
    """state injection (bonds included) + continuation equals the reference's own continuation."""
    from common import load_golden_state
    g = load_golden_state(os.path.join(golden_dir, "hot200_step40000.npz"))
    fr = KAT["hot200_40000"]["frames"][-1]
    assert "%016x" % refio.fnv1a64(g["res_nei"], g["status"]) == fr["hash_bonds"]
    p = apply_regime(pyoracle.default_params(box=tuple(g["params"]["box"]), use_grid=1), "hot")
    o = pyoracle.Oracle(p)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    c = o.counts()
    assert (c["bond_num"], c["bond_num_rl"], c["bond_num_cis"], c["bond_num_mono_cis"]) == (fr["bond_num"], fr["bond_num_rl"], fr["bond_num_cis"], fr["bond_num_mono_cis"])
    if not refio.ref_available("n200"):
        pytest.skip("oracle/_ref not built here")
    # continue 500 steps in the reference (restart through its own position.cpt reader) and in the oracle
    sets = dict(cell_range_x=2500, cell_range_y=2500, cell_range_z=400, Diss_Rate=2e-5, cis_Diss_Rate=2e-5, mono_cis_Diss_Rate=1e-4)
    frame = dict(step=g["step"], bond_num=c["bond_num"], bond_num_rl=c["bond_num_rl"], bond_num_cis=c["bond_num_cis"],
                 bond_num_mono_cis=c["bond_num_mono_cis"], max_complex=g["max_complex"], R=g["R"], status=g["status"], res_nei=g["res_nei"])
    s, frames = refio.run_ref("n200", 200, 500, sets=sets, scales=dict(cis_Ass_Rate=20, mono_cis_Ass_Rate=20), in_frame=frame)
    o.step(500)
    R, st, rn = o.get_state()
    assert np.array_equal(rn, frames[-1]["res_nei"]) and np.array_equal(st, frames[-1]["status"])
    assert np.array_equal(R, frames[-1]["R"])


@pytest.mark.skipif(not refio.ref_available("n200"), reason="oracle/_ref not built (no /root/reference on this box)")
def test_live_reference_other_seed():
    """a stream the golden file has never seen: reference and oracle bit-equal, incl. draw counts."""
    st2, str_ = 0x1234567887654321, 0x0F0F0F0F12345678
    p = apply_regime(pyoracle.default_params(box=(2500, 2500, 400), use_grid=1, rand2_state=st2, rand_state=str_), "hot")
    o = pyoracle.Oracle(p)
    o.init_reference()
    o.step(4000)
    sets = dict(cell_range_x=2500, cell_range_y=2500, cell_range_z=400, Diss_Rate=2e-5, cis_Diss_Rate=2e-5, mono_cis_Diss_Rate=1e-4)
    s, frames = refio.run_ref("n200", 200, 4000, sets=sets, scales=dict(cis_Ass_Rate=20, mono_cis_Ass_Rate=20), rand2_state=st2, rand_state=str_)
    R, st, rn = o.get_state()
    assert np.array_equal(R, frames[-1]["R"]) and np.array_equal(rn, frames[-1]["res_nei"]) and np.array_equal(st, frames[-1]["status"])
    c = o.counts()
    assert (c["rand2_draws"], c["rand_draws"]) == (s["rand2_draws"], s["rand_draws"])


def test_compact_pose_invariants(golden_dir):
    """What the CUDA layout relies on: in reference states every receptor is a stack of four beads with
    identical xy per site, site 4 above the centre, z exactly the template -- checked on evolved reference states."""
    from common import load_golden_state
    for name in ("dense_step30000.npz", "hot200_step40000.npz", "hot40_step200000.npz"):
        g = load_golden_state(os.path.join(golden_dir, name))
        R = g["R"]
        na = R.shape[0] - 1 - (R.shape[0] - 1) // 4
        A = R[1:na + 1]
        for j in range(1, 5):
            assert np.array_equal(A[:, j, 1:4, :2], A[:, 1, 1:4, :2])
            assert np.array_equal(A[:, j, 4, :2], A[:, 1, 1, :2])
            assert np.all(A[:, j, 1:4, 2] == (2 * j - 2) * 20.0) and np.all(A[:, j, 4, 2] == (2 * j - 1) * 20.0)


def test_keyed_oracle_against_keyed_reference_golden(golden_dir):
    """The keyed (Philox) stream: golden frames written by the UNMODIFIED reference when ref_harness.cpp keys every draw by its call
    site (return address -> main.cpp line -> slot), molecule, partner and step. The oracle in stream_mode=1 must reproduce them bit for
    bit -- this is the stream the GPU replay tests use."""
    from common import load_golden_state
    kk = KAT["keyed_hot200"]
    g = load_golden_state(os.path.join(golden_dir, kk["start"]))
    p = apply_regime(pyoracle.default_params(box=tuple(g["params"]["box"]), use_grid=1, stream_mode=1, seed=kk["seed"]), "hot")
    o = pyoracle.Oracle(p)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    done = g["step"]
    for fr in kk["frames"]:
        o.step(fr["step"] - done); done = fr["step"]
        hr, hb = hashes(o)
        assert (hb, hr) == (fr["hash_bonds"], fr["hash_R"]), "keyed oracle differs from the keyed reference at step %d" % done
    assert o.counts()["rand_draws"] == kk["summary"]["rand_draws"]


@pytest.mark.skipif(not refio.ref_available("n40"), reason="oracle/_ref not built (no /root/reference on this box)")
def test_keyed_reference_live_n40(golden_dir):
    from common import load_golden_state
    g = load_golden_state(os.path.join(golden_dir, "hot40_step200000.npz"))
    sets = dict(cell_range_x=1000, cell_range_y=1000, cell_range_z=300, Diss_Rate=2e-5, cis_Diss_Rate=2e-5, mono_cis_Diss_Rate=1e-4)
    fr = dict(step=g["step"], bond_num=0, bond_num_rl=0, bond_num_cis=0, bond_num_mono_cis=0, max_complex=g["max_complex"], R=g["R"], status=g["status"], res_nei=g["res_nei"])
    s, frames = refio.run_ref("n40", 40, 20000, sets=sets, scales=dict(cis_Ass_Rate=20, mono_cis_Ass_Rate=20), in_frame=fr, keyed_seed=77)
    p = apply_regime(pyoracle.default_params(box=(1000, 1000, 300), n_receptor=30, n_ligand=10, use_grid=1, stream_mode=1, seed=77), "hot")
    o = pyoracle.Oracle(p)
    o.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    o.step(20000)
    R, st, rn = o.get_state()
    assert np.array_equal(R, frames[-1]["R"]) and np.array_equal(rn, frames[-1]["res_nei"]) and np.array_equal(st, frames[-1]["status"])
