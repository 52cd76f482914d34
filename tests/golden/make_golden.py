#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerates the golden vectors from the UNMODIFIED reference
(oracle/_ref/kmcref_*, built by oracle/build_ref.py from /root/reference/main.cpp). Run in the build
container only (the reference source does not exist on the GPU box); the outputs are committed.

  ref_kat.json          hashes / counters of reference runs under the sequential xorshift streams of
                        oracle/ref_harness.cpp (FNV-1a-64 over the raw bytes of R_x,R_y,R_z and of
                        res_nei,protein_status), incl. the two known-answer vectors of SURVEY.md 8c
  dense_step30000.npz   full reference state of the dense oligomerising system after 30000 steps
                        (63 bonds, complexes up to 10 members): start state for GPU replay tests
  hot40_step200000.npz  N=40 hot system (dissociation + goto lable4 exercised) after 200000 steps
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refio  # noqa: E402

DENSE = dict(sets=dict(cell_range_x=2500, cell_range_y=2500, cell_range_z=400), scales=dict(cis_Ass_Rate=20, mono_cis_Ass_Rate=20))
HOT = dict(Diss_Rate=2e-5, cis_Diss_Rate=2e-5, mono_cis_Diss_Rate=1e-4)


def summarize(fr):
    R = fr["R"]
    return dict(step=fr["step"], bond_num=fr["bond_num"], bond_num_rl=fr["bond_num_rl"], bond_num_cis=fr["bond_num_cis"],
                bond_num_mono_cis=fr["bond_num_mono_cis"], max_complex=fr["max_complex"],
                hash_R="%016x" % refio.fnv1a64(R[..., 0], R[..., 1], R[..., 2]),
                hash_bonds="%016x" % refio.fnv1a64(fr["res_nei"], fr["status"]))


def save_state(name, fr, params):
    np.savez_compressed(os.path.join(HERE, name), R=fr["R"], status=fr["status"], res_nei=fr["res_nei"], step=fr["step"],
                        max_complex=fr["max_complex"], params=json.dumps(params))


def keyed_entry(kat):
    """the UNMODIFIED reference driven by the keyed Philox stream (ref_harness.cpp --keyed: draw = f(seed; call site -> slot, molecule,
    partner, step)), continued from the hot 40000-step state: pins the keyed mode of the oracle and, through the bond table, the GPU"""
    g40 = np.load(os.path.join(HERE, "hot200_step40000.npz"))
    last = kat["hot200_40000"]["frames"][-1]
    fr_in = dict(step=int(g40["step"]), bond_num=last["bond_num"], bond_num_rl=last["bond_num_rl"], bond_num_cis=last["bond_num_cis"],
                 bond_num_mono_cis=last["bond_num_mono_cis"], max_complex=int(g40["max_complex"]), R=g40["R"], status=g40["status"], res_nei=g40["res_nei"])
    sets_k = dict(DENSE["sets"]); sets_k.update(HOT)
    s, fr = refio.run_ref("n200", 200, 3000, frames_every=1000, sets=sets_k, scales=DENSE["scales"], in_frame=fr_in, keyed_seed=4242)
    kat["keyed_hot200"] = dict(summary=s, frames=[summarize(f) for f in fr], seed=4242, start="hot200_step40000.npz", sets=sets_k, scales=DENSE["scales"])


def main():
    if "--only-keyed" in sys.argv:          # the key layout of the Philox stream changed: only this entry depends on it
        path = os.path.join(HERE, "ref_kat.json")
        kat = json.load(open(path))
        keyed_entry(kat)
        json.dump(kat, open(path, "w"), indent=1)
        print(json.dumps(kat["keyed_hot200"]["frames"][-1]))
        return
    kat = {}
    s, fr = refio.run_ref("n200", 200, 1000)
    kat["default_1000"] = dict(summary=s, frames=[summarize(f) for f in fr])
    s, fr = refio.run_ref("n200", 200, 30000, frames_every=5000, **DENSE)
    kat["dense_30000"] = dict(summary=s, frames=[summarize(f) for f in fr], sets=DENSE["sets"], scales=DENSE["scales"])
    save_state("dense_step30000.npz", fr[-1], dict(box=[2500, 2500, 400], cis_on_scale=20, mono_cis_on_scale=20))
    sets = dict(cell_range_x=1000, cell_range_y=1000, cell_range_z=300); sets.update(HOT)
    s, fr = refio.run_ref("n40", 40, 200000, frames_every=50000, sets=sets, scales=DENSE["scales"])
    kat["hot40_200000"] = dict(summary=s, frames=[summarize(f) for f in fr], sets=sets, scales=DENSE["scales"])
    save_state("hot40_step200000.npz", fr[-1], dict(box=[1000, 1000, 300], cis_on_scale=20, mono_cis_on_scale=20, **HOT))
    sets = dict(DENSE["sets"]); sets.update(HOT)
    s, fr = refio.run_ref("n200", 200, 40000, frames_every=10000, sets=sets, scales=DENSE["scales"])
    kat["hot200_40000"] = dict(summary=s, frames=[summarize(f) for f in fr], sets=sets, scales=DENSE["scales"])
    save_state("hot200_step40000.npz", fr[-1], dict(box=[2500, 2500, 400], cis_on_scale=20, mono_cis_on_scale=20, **HOT))
    # the reference's own output records (bond.dat, cluster.log; main.cpp:2247-2253, 2291-2305) of the dense run
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        wd = os.path.join(td, "wd")
        cmd = [os.path.join(refio.REF_DIR, "kmcref_n200"), "--steps", "10000", "--workdir", wd]
        for k2, v2 in DENSE["sets"].items():
            cmd += ["--set", "%s=%r" % (k2, float(v2))]
        for k2, v2 in DENSE["scales"].items():
            cmd += ["--scale", "%s=%r" % (k2, float(v2))]
        subprocess.run(cmd, cwd=td, check=True, capture_output=True)
        kat["ref_records"] = dict(bond_dat=open(os.path.join(wd, "bond.dat")).read(), cluster_log=open(os.path.join(wd, "cluster.log")).read(),
                                  note="dense system, steps 5000 and 10000, written by the unmodified reference")
    keyed_entry(kat)
    # the reference's own position.cpt / test.gro / parameter.log of a 5000-step N=40 hot run, with the state they describe
    with tempfile.TemporaryDirectory() as td:
        wd = os.path.join(td, "wd"); out = os.path.join(td, "f.bin")
        sets40 = dict(cell_range_x=1000, cell_range_y=1000, cell_range_z=300); sets40.update(HOT)
        cmd = [os.path.join(refio.REF_DIR, "kmcref_n40"), "--steps", "5000", "--workdir", wd, "--out", out]
        for k2, v2 in sets40.items():
            cmd += ["--set", "%s=%r" % (k2, float(v2))]
        for k2, v2 in DENSE["scales"].items():
            cmd += ["--scale", "%s=%r" % (k2, float(v2))]
        subprocess.run(cmd, cwd=td, check=True, capture_output=True)
        fr = refio.read_frames(out, 40)[-1]
        kat["ref_files40"] = dict(position_cpt=open(os.path.join(wd, "position.cpt")).read(), test_gro=open(os.path.join(wd, "test.gro")).read(),
                                  parameter_log=open(os.path.join(wd, "parameter.log")).read(), sets=sets40, scales=DENSE["scales"])
        save_state("hot40_step5000.npz", fr, dict(box=[1000, 1000, 300], cis_on_scale=20, mono_cis_on_scale=20, bond_num=fr["bond_num"], bond_num_rl=fr["bond_num_rl"],
                                                  bond_num_cis=fr["bond_num_cis"], bond_num_mono_cis=fr["bond_num_mono_cis"], **HOT))
    json.dump(kat, open(os.path.join(HERE, "ref_kat.json"), "w"), indent=1)
    print(json.dumps({k: v["frames"][-1] for k, v in kat.items()}, indent=1))


if __name__ == "__main__":
    main()
