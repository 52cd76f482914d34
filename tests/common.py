"""Shared helpers of the test-suite: parameter sets used by both the oracle (checker) and the CUDA library
(the thing checked), state comparison with the tolerances the north star states."""
import json
import os

import numpy as np

POS_RTOL = 1e-12      # "positions matching to within 1e-12 relative in fp64" (BASELINE.json north_star)
POS_FLOOR = 100.0     # Angstrom. A position is a vector: its error is measured relative to max(|r|, 100 A), r = the point's
                      # position vector. (Every rotation mixes x and y, so each coordinate carries rounding noise of the size
                      # of ulp(|r|): a point at (50, 60000) has x-noise of ulp(60000) per step, whatever x is. A point that
                      # passes near the origin has no meaningful self-relative error at all; 100 A is the rigid-body lever arm
                      # -- ligand site radius 64.6 A, complex radius ~100 A -- that turns an orientation rounding error into a
                      # coordinate error.) The strict per-coordinate figures are logged next to it (log_errors).

HOT = dict(off=2e-5, cis_off=2e-5, mono_cis_off=1e-4)


def apply_regime(p, regime, n_total=None):
    """regime: 'default' | 'dense' | 'hot' (dense + fast dissociation). Mutates a Params struct of either binding."""
    if regime in ("dense", "hot"):
        p.cis_on *= 20
        p.mono_cis_on *= 20
    if regime == "hot":
        for k, v in HOT.items():
            setattr(p, k, v)
    return p


def load_golden_state(path):
    z = np.load(path)
    return dict(R=z["R"], status=z["status"], res_nei=z["res_nei"], step=int(z["step"]), max_complex=int(z["max_complex"]),
                params=json.loads(str(z["params"])))


def strict_errors(ref, got):
    """(max |x_ref - x| in Angstrom, max |x_ref - x| / |x_ref| over coordinates with |x_ref| > 1e-6 A): the strict figures next to the
    floored one that compare_states asserts"""
    Rr, Rg = ref[0], got[0]
    d = np.abs(Rr - Rg)
    big = np.abs(Rr) > 1e-6
    return float(d.max()), float((d[big] / np.abs(Rr[big])).max()) if big.any() else 0.0


def log_errors(label, floored, ref, got):
    """one line per comparison into gpurun_out/position_errors.jsonl (brought back from the GPU box; summarised under profiles/)"""
    a, r = strict_errors(ref, got)
    line = json.dumps({"test": label, "rel_floored_100A": floored, "abs_A": a, "rel_strict": r})
    print("position error", line)
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "position_errors.jsonl"), "a") as f:
            f.write(line + "\n")
    return a, r


def compare_states(ref, got, label=""):
    """ref/got = (R, status, res_nei). Bond table must be identical; positions within POS_RTOL relative
    (relative to max(|r_ref|, POS_FLOOR), r = position vector of the point). Returns max relative position error."""
    Rr, sr, nr = ref
    Rg, sg, ng = got
    assert np.array_equal(sr, sg), label + ": protein_status differs"
    assert np.array_equal(nr, ng), label + ": res_nei differs"
    scale = np.maximum(np.sqrt((Rr * Rr).sum(axis=-1, keepdims=True)), POS_FLOOR)
    err = np.abs(Rr - Rg) / scale
    worst = float(err.max())
    assert worst <= POS_RTOL, "%s: position error %.3e exceeds %.1e at %s" % (label, worst, POS_RTOL, np.unravel_index(err.argmax(), err.shape))
    return worst
