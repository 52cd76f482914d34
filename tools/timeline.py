"""KMC_TIMELINE=1 python tools/timeline.py : where the kernels of one step graph run on the device clock (1.25e6-molecule bench workload)"""
import os, sys
os.environ["KMC_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200
M = 1250000
k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=3 * M // 4, n_ligand=M // 4, seed=1))
k.init_random(seed=1, sort_cells=True)
k.step(304); k.sync()          # last step: a reuse step
kmc_b200.lib().kmc_timeline_print(k.h)
k.step(3); k.sync()            # 307 = 6*51+1: last step is a build step
print("---- build step", file=sys.stderr)
kmc_b200.lib().kmc_timeline_print(k.h)
print("ms/step with stamps", k.step_timed(300) / 300)
