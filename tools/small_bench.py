import sys, time
sys.path.insert(0, '/root/repo/kmc-with-a-diffusion-reaction-algorithm_b200')
import kmc_b200
for R in (1, 64, 1024):
    p = kmc_b200.default_params(n_replicas=R, seed=3)
    k = kmc_b200.Kmc(p); k.init_random(seed=2)
    k.step(200); k.sync()
    ms = k.step_timed(2000)
    print("replicas %4d: %.1f us/step, %.3e moves/s" % (R, ms*1e3/2000, 200*R*2000/(ms*1e-3)), k.series(0), flush=True)
    k.close()
