#!/usr/bin/env python3
"""tools/affinity_sweep.py -- BASELINE.json configs[4]: affinity sweep (k_on/k_off of the receptor-ligand bond, angle cutoffs) on
a large membrane, output = oligomer-size histograms and the bond.dat columns per grid point (one JSON line each).

  python tools/affinity_sweep.py --molecules 1000000 --steps 2000 --on 0.004,0.04,0.4 --off 3.48e-13,3.48e-7 --thetaot 45,90

With torchrun the grid points are dealt round-robin to the ranks (independent runs, no communication); each rank prints its own
lines, rank 0 a final summary line."""
import argparse
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--molecules", type=int, default=1000000)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--on", default="0.04", help="Ass_Rate values (per ns), comma separated")
    ap.add_argument("--off", default="3.48e-13", help="Diss_Rate values")
    ap.add_argument("--thetaot", default="90", help="bond_thetaot_cutoff values (degrees)")
    ap.add_argument("--thetapd", default="45", help="bond_thetapd_cutoff values (degrees)")
    ap.add_argument("--density-scale", type=float, default=1.0, help=">1 shrinks the box (faster encounters)")
    ap.add_argument("--mode", default="production", choices=["production", "replay"])
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--nbins", type=int, default=64)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    grid = list(itertools.product([float(x) for x in a.on.split(",")], [float(x) for x in a.off.split(",")],
                                  [float(x) for x in a.thetaot.split(",")], [float(x) for x in a.thetapd.split(",")]))
    M = a.molecules
    L = 5773.0 * (M / 200.0 / a.density_scale) ** 0.5
    done = 0
    for gi, (on, off, tot, tpd) in enumerate(grid):
        if gi % world != rank:
            continue
        p = kmc_b200.default_params(box=(L, L, 1000.0), n_receptor=3 * M // 4, n_ligand=M - 3 * M // 4, seed=a.seed + gi, device=local,
                                    mode=kmc_b200.MODE_PRODUCTION if a.mode == "production" else kmc_b200.MODE_REPLAY)
        p.on, p.off, p.thetaot_cut, p.thetapd_cut = on, off, tot, tpd
        k = kmc_b200.Kmc(p)
        k.init_random(seed=a.seed, sort_cells=True)
        t0 = time.perf_counter()
        k.step(a.steps); k.sync()
        dt = time.perf_counter() - t0
        hist = k.oligomer_hist(nbins=a.nbins)
        s = k.series()
        print(json.dumps({"grid_point": gi, "rank": rank, "on": on, "off": off, "thetaot_cut": tot, "thetapd_cut": tpd, "molecules": M, "steps": a.steps,
                          "series": s, "oligomer_hist": {str(i): int(c) for i, c in enumerate(hist) if c}, "moves_per_s": M * a.steps / dt}), flush=True)
        k.close(); done += 1
    if rank == 0:
        print(json.dumps({"summary": "affinity sweep", "grid_points": len(grid), "world": world, "mode": a.mode}), flush=True)


if __name__ == "__main__":
    main()
