import sys, ctypes as C, numpy as np
sys.path.insert(0, '/root/repo/kmc-with-a-diffusion-reaction-algorithm_b200')
import kmc_b200
M = 1250000
p = kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=3*M//4, n_ligand=M-3*M//4, seed=1)
k = kmc_b200.Kmc(p); k.init_random(seed=1, sort_cells=True)
k.step(20); k.sync()
e0 = np.zeros(16, dtype=np.int64); kmc_b200.lib().kmc_get_events(k.h, e0.ctypes.data)
ms = k.step_timed(20)
e1 = np.zeros(16, dtype=np.int64); kmc_b200.lib().kmc_get_events(k.h, e1.ctypes.data)
d = e1 - e0
n = d[15]
print("ms/step", ms/20, "CTAs", n/20)
for i, name in enumerate(["cs window load", "row prefix", "staging gather", "phase1 cut scan", "phase2 survivors"]):
    print("%-18s %8.0f cycles/CTA" % (name, d[10+i]/n))
