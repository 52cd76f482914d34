"""the oligomerised regime of bench.py's extras on its own: 1.25e6 molecules, pre-assembled complexes, per-kernel times"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200
from kmc_b200.synth import oligomerised_state
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1250000
na, nb = 3 * M // 4, M // 4
p = kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=na, n_ligand=nb, seed=1)
st = oligomerised_state(p, seed=1, bound_fraction=0.6)
k = kmc_b200.Kmc(p)
k.set_packed(*st)
k.step(120); k.sync()
ms = k.step_timed(300) / 300
s = k.series()
k.profile(True); k.step(60); k.sync(); prof = k.profile_get(); k.profile(False)
print(json.dumps({"ms_per_mc_step": ms, "moves_per_s": M / (ms * 1e-3), "bonds": s["bond_num"], "complexes": s["n_complexes"], "max_complex": s["max_complex"],
                  "kernels_us": {n: round(1e3 * v[0] / 60, 1) for n, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[1]}}))
