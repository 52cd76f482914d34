"""step time in a complex-rich state (most molecules bound): the fresh-start bench.py workload has almost no complexes"""
import sys, time
sys.path.insert(0, '/root/repo/kmc-with-a-diffusion-reaction-algorithm_b200'); sys.path.insert(0, '/root/repo/tests')
import kmc_b200
from common import apply_regime
M = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = 5773.0 * (M / 200.0 / 5.3) ** 0.5          # the dense regime of SURVEY 8 1/2 (2500 A box for 200 molecules)
p = apply_regime(kmc_b200.default_params(box=(L, L, 400.0), n_receptor=3 * M // 4, n_ligand=M - 3 * M // 4, seed=3), "dense")
k = kmc_b200.Kmc(p); k.init_random(seed=2, sort_cells=True)
for target in (0, 2000, 10000, 30000):
    k.step(target - k.series()["step"]); k.sync()
    ms = k.step_timed(200) / 200
    s = k.series()
    k.profile(True); k.step(50); k.sync(); prof = k.profile_get(); k.profile(False)
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:4]
    print("step %6d bonds %7d complexes %6d max %3d : %.3f ms/step  %.3e moves/s  top: %s" % (s["step"], s["bond_num"], s["n_complexes"], s["max_complex"], ms, M / (ms * 1e-3),
          ", ".join("%s %.3f" % (n, v[0] / 50) for n, v in top)), flush=True)
