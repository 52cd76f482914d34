"""two-ligand complexes only (tools/olig_split.py, share 1.0): a few steps, for an ncu capture of the complex kernels"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200
from kmc_b200.synth import oligomerised_state
M = 1250000
p = kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=3 * M // 4, n_ligand=M // 4, seed=1)
k = kmc_b200.Kmc(p); k.set_packed(*oligomerised_state(p, seed=1, bound_fraction=0.6, two_ligand_share=float(sys.argv[1]) if len(sys.argv) > 1 else 1.0))
k.profile(True); k.step(12); k.sync()
