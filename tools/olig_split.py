import os, sys, json
sys.path.insert(0, "/root/repo/kmc-with-a-diffusion-reaction-algorithm_b200")
import kmc_b200
from kmc_b200.synth import oligomerised_state
M = 1250000
na, nb = 3 * M // 4, M // 4
for share in (0.0, 1.0):
    p = kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=na, n_ligand=nb, seed=1)
    st = oligomerised_state(p, seed=1, bound_fraction=0.6, two_ligand_share=share)
    k = kmc_b200.Kmc(p); k.set_packed(*st); k.step(60); k.sync()
    ms = k.step_timed(100) / 100
    s = k.series()
    k.profile(True); k.step(30); k.sync(); prof = k.profile_get(); k.profile(False)
    print(json.dumps({"two_ligand_share": share, "ms": ms, "complexes": s["n_complexes"], "bonds": s["bond_num"], "kernels_us": {n: round(1e3 * v[0] / 30, 1) for n, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:5] if v[1]}}), flush=True)
    k.close()
