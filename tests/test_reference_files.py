"""CPU: the reference's remaining text files -- position.cpt writer + reader, test.gro, parameter.log --
produced by the library's host functions must be byte-identical to the files the UNMODIFIED reference wrote for the same state
(tests/golden/ref_kat.json: ref_files40 + hot40_step5000.npz, made by tests/golden/make_golden.py from oracle/_ref/kmcref_n40)."""
import json
import os

import numpy as np

import kmc_b200
import pyoracle
from common import apply_regime, load_golden_state

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLDEN, "ref_kat.json")))["ref_files40"]
G = load_golden_state(os.path.join(GOLDEN, "hot40_step5000.npz"))
NA, NB = 30, 10


def test_position_cpt_writer_is_byte_identical(tmp_path):
    p = G["params"]
    out = tmp_path / "position.cpt"
    kmc_b200.checkpoint_write(str(out), NA, NB, G["R"], G["status"], G["res_nei"],
                              (p["bond_num"], p["bond_num_rl"], p["bond_num_cis"], p["bond_num_mono_cis"], G["max_complex"], G["step"]))
    assert out.read_text() == KAT["position_cpt"]


def test_position_cpt_reader_parses_like_the_reference(tmp_path):
    """the reader is a token stream (main.cpp:231-266): the 3-decimal file gives the rounded state; a full-precision file round-trips exactly"""
    f = tmp_path / "position.cpt"
    f.write_text(KAT["position_cpt"])
    R, st, rn, c = kmc_b200.checkpoint_read(str(f), NA, NB)
    assert np.array_equal(st, G["status"]) and np.array_equal(rn, G["res_nei"])
    used = np.zeros(R.shape[:3], bool); used[1:NA + 1, 1:5, 1:5] = True; used[NA + 1:, 1:5, 1:3] = True
    assert np.abs(R - G["R"])[used].max() <= 5.0001e-4          # lossy by design
    assert np.array_equal(R[used], np.round(G["R"], 3)[used]) or np.abs(R[used] - np.round(G["R"][used], 3)).max() < 1e-9
    assert c[5] == G["step"] and c[4] == G["max_complex"]
    # restart equivalence with the reference's own reader: feed the oracle the parsed state and the reference the same file
    import refio
    if refio.ref_available("n40"):
        fr = dict(step=int(c[5]), bond_num=int(c[0]), bond_num_rl=int(c[1]), bond_num_cis=int(c[2]), bond_num_mono_cis=int(c[3]), max_complex=int(c[4]),
                  R=R, status=st, res_nei=rn)
        s, frames = refio.run_ref("n40", 40, 300, sets=KAT["sets"], scales=KAT["scales"], in_frame=fr)
        po = apply_regime(pyoracle.default_params(box=(1000, 1000, 300), n_receptor=NA, n_ligand=NB, use_grid=1), "hot")
        o = pyoracle.Oracle(po)
        o.set_state(R, st, rn, step_done=int(c[5]), max_complex=int(c[4]))
        o.step(300)
        assert np.array_equal(o.get_state()[0], frames[-1]["R"]) and np.array_equal(o.get_state()[2], frames[-1]["res_nei"])


def test_gro_frame_is_byte_identical(tmp_path):
    out = tmp_path / "test.gro"
    kmc_b200.gro_append(str(out), NA, NB, G["R"], 10.0, G["step"], (1000, 1000, 300))
    assert out.read_text() == KAT["test_gro"]


def test_parameter_log_is_byte_identical(tmp_path):
    p = apply_regime(kmc_b200.default_params(box=(1000, 1000, 300), n_receptor=NA, n_ligand=NB), "hot")
    out = tmp_path / "parameter.log"
    kmc_b200.parameter_log_write(p, str(out))
    assert out.read_text() == KAT["parameter_log"]
