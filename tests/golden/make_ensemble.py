#!/usr/bin/env python3
"""tests/golden/make_ensemble.py -- 64-seed ensemble of the UNMODIFIED reference (oracle/_ref/kmcref_n200) in the dense
oligomerising regime (box 2500x2500x400 A, cis on-rates x20), 30 000 steps each, for the statistical (production-mode) test:
per seed the bond-count time series every 5000 steps and the final oligomer sizes. Output: ref_ensemble_dense.json.
Run in the build container only (needs the reference binary); ~30 s per seed per core."""
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refio  # noqa: E402

NA, NB, STEPS, EVERY = 150, 50, 30000, 5000
SETS = dict(cell_range_x=2500, cell_range_y=2500, cell_range_z=400)
SCALES = dict(cis_Ass_Rate=20, mono_cis_Ass_Rate=20)


def complex_sizes(res_nei):
    """sizes of ligand-rooted complexes (main.cpp:525-562) from a bond table"""
    n = NA + NB
    seen, sizes = set(), []
    for root in range(NA + 1, n + 1):
        if root in seen:
            continue
        comp, stack = {root}, [root]
        while stack:
            m = stack.pop()
            nb = [res_nei[m][2], res_nei[m][3]] if m <= NA else [res_nei[m][2], res_nei[m][3], res_nei[m][4]]
            for v in nb:
                if v > 0 and v not in comp:
                    comp.add(v); stack.append(v)
        seen |= comp
        sizes.append(len(comp))
    return sizes


def one(seed):
    s2 = (0x9E3779B97F4A7C15 * (seed + 1)) & 0xFFFFFFFFFFFFFFFF or 1
    sr = (0xD1B54A32D192ED03 * (seed + 1)) & 0xFFFFFFFFFFFFFFFF or 1
    summ, frames = refio.run_ref("n200", NA + NB, STEPS, sets=SETS, scales=SCALES, rand2_state=s2, rand_state=sr, frames_every=EVERY)
    series = [dict(step=f["step"], rl=f["bond_num_rl"], mono=f["bond_num_mono_cis"], cis=f["bond_num_cis"], bonds=f["bond_num"]) for f in frames]
    last = frames[-1]
    return dict(seed=seed, series=series, max_complex=last["max_complex"], sizes=complex_sizes(last["res_nei"].tolist()))


if __name__ == "__main__":
    nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        rows = list(ex.map(one, range(nseeds)))
    json.dump(dict(regime="dense", box=[2500, 2500, 400], scales=SCALES, steps=STEPS, every=EVERY, seeds=rows),
              open(os.path.join(HERE, "ref_ensemble_dense.json"), "w"))
    print("final R-L bonds: mean %.1f sd %.1f" % (np.mean([r["series"][-1]["rl"] for r in rows]), np.std([r["series"][-1]["rl"] for r in rows])))
