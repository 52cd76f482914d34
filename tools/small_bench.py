"""small systems: the default system, 64 and 1024 replicas of it (fused step: one CTA per replica), the 1e5-molecule membrane.
KMC_LIB selects a build variant; argv[1:] = replica counts"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200
reps = [int(x) for x in sys.argv[1:]] or [1, 64, 1024]
for R in reps:
    k = kmc_b200.Kmc(kmc_b200.default_params(n_replicas=R, seed=3)); k.init_random(seed=2)
    k.step(200); k.sync()
    ms = k.step_timed(2000)
    print("%s replicas %4d (%s): %.2f us/step, %.3e moves/s" % (os.environ.get("KMC_LIB", "default").split("/")[-1], R, k.path(), ms * 1e3 / 2000, 200 * R * 2000 / (ms * 1e-3)), flush=True)
    k.close()
if len(sys.argv) == 1:
    k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(100000), n_receptor=75000, n_ligand=25000, seed=1)); k.init_random(seed=1, sort_cells=True)
    k.step(200); k.sync(); ms = k.step_timed(1000)
    print("1e5 molecules: %.1f us/step, %.3e moves/s" % (ms, 1e5 * 1000 / (ms * 1e-3)), flush=True)
