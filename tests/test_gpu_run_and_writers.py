"""GPU: the handle-level writers and the reference's main loop on top of the library (kmc_run, host/kmc_main):
bond.dat / cluster.log / test.gro / position.cpt written every output period (main.cpp:2206-2305) must be byte-identical to the
same records formatted from the ORACLE's numbers for the same trajectory."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import kmc_b200
import pyoracle
from common import apply_regime, load_golden_state

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KMC_MAIN = os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200", "host", "kmc_main")


def test_kmc_run_writes_the_reference_records(golden_dir, tmp_path):
    """two output periods of kmc_run from a reference-evolved state with complexes; every file compared byte for byte with the
    formatter fed by the oracle's state at the same steps"""
    g = load_golden_state(os.path.join(golden_dir, "dense_step30000.npz"))
    box = tuple(g["params"]["box"])
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=box, seed=8), "dense"))
    o = pyoracle.Oracle(apply_regime(pyoracle.default_params(box=box, use_grid=1, stream_mode=1, seed=8), "dense"))
    for x in (k, o):
        x.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    every, dt, na, nb = 50, k.p.dt, 150, 50
    out = tmp_path / "gpu"; out.mkdir()
    k.run(2 * every, every, str(out))
    want = tmp_path / "oracle"; want.mkdir()
    bond, cluster = "", ""
    for period in range(2):
        o.step(every)
        c = o.counts()
        R, st, rn = o.get_state()
        bond += kmc_b200.format_bond_dat(dt, c["step"], c["bond_num_rl"], c["bond_num_mono_cis"], c["bond_num_cis"], c["bond_num"], c["cluster_size"], c["max_complex"])
        cluster += kmc_b200.format_cluster_log(dt, c["step"], o.results())
        kmc_b200.gro_append(str(want / "test.gro"), na, nb, R, dt, c["step"], box)
        kmc_b200.checkpoint_write(str(want / "position.cpt"), na, nb, R, st, rn,
                                  (c["bond_num"], c["bond_num_rl"], c["bond_num_cis"], c["bond_num_mono_cis"], c["max_complex"], c["step"]))
    assert (out / "bond.dat").read_text() == bond
    assert (out / "cluster.log").read_text() == cluster
    assert (out / "test.gro").read_text() == (want / "test.gro").read_text()
    assert (out / "position.cpt").read_text() == (want / "position.cpt").read_text()
    assert bond.count("\n") == 2 and cluster.count("Hello Cluster!") == 2


def test_kmc_run_reports_capacity_overflow(golden_dir, tmp_path, monkeypatch):
    """a device work list that overflows during a long kmc_run must come back as KMC_ERR_CAPACITY (never silently dropped):
    the pending-findings list is shrunk to one entry, the crowded system needs more"""
    monkeypatch.setenv("KMC_TEST_PENDCAP", "1")
    na, nb = 1500, 500
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=(7500.0, 7500.0, 400.0), n_receptor=na, n_ligand=nb, seed=21), "hot"))
    k.init_random(seed=8)
    with pytest.raises(kmc_b200.KmcError, match="overflow"):
        k.run(400, 200, str(tmp_path))


def test_kmc_main_binary_start_and_restart(tmp_path):
    """host/kmc_main (the reference's program shape over the C ABI): a fresh start writes parameter.log and the four output
    files exactly like kmc_run driven through the binding; a second invocation finds position.cpt and continues (main.cpp:226-268)"""
    if not os.path.exists(KMC_MAIN):
        pytest.skip("host/kmc_main not built")
    args = ["--output-every", "40", "--box", "2500", "2500", "400", "--seed", "5", "--set", "cis_on=%r" % (0.00096 * 20), "--set", "mono_cis_on=%r" % (0.000047 * 20)]
    wd = tmp_path / "bin"; wd.mkdir()
    p = subprocess.run([KMC_MAIN, "--steps", "80"] + args, cwd=wd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "CPT file not exist" in p.stdout
    lib = tmp_path / "lib"; lib.mkdir()
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=(2500.0, 2500.0, 400.0), seed=5), "dense"))
    k.init_random(seed=5, sort_cells=False)
    k.run(80, 40, str(lib))
    for name in ("bond.dat", "cluster.log", "test.gro", "position.cpt"):
        assert (wd / name).read_text() == (lib / name).read_text(), name
    assert (wd / "parameter.log").read_text().startswith("          box size: x y z")
    # restart: the same command with a larger step count continues from position.cpt and appends
    p = subprocess.run([KMC_MAIN, "--steps", "120"] + args, cwd=wd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert "CPT file exist" in p.stdout
    rows = (wd / "bond.dat").read_text().splitlines()
    assert len(rows) == 3 and rows[2].split()[0] == "1200.000"
    assert (wd / "cluster.log").read_text().count("Hello Cluster!") == 3
    assert (wd / "test.gro").read_text().count("Hello Gro!") == 3
