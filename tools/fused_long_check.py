"""fused small-system step against the general path, bit for bit, over a long hot run (associations, dissociations, complexes):
python tools/fused_long_check.py [replicas] [steps]. 32 x 30001 (one replica per CTA) and 1024 x 3001 (four replicas per CTA in
lockstep + ticket queue) were run on the B200 in both sweep orders: identical."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import kmc_b200
from common import apply_regime
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30001


def mk(fused, mode):
    os.environ["KMC_FUSED"] = fused
    k = kmc_b200.Kmc(apply_regime(kmc_b200.default_params(box=(2500.0, 2500.0, 400.0), seed=77, n_replicas=R, mode=mode), "hot"))
    del os.environ["KMC_FUSED"]
    return k


for mode in (kmc_b200.MODE_REPLAY, kmc_b200.MODE_PRODUCTION):
    a, b = mk("1", mode), mk("0", mode)
    a.init_random(seed=5); b.set_packed(*a.get_packed())
    a.step(N); b.step(N)
    same = all(np.array_equal(x, y) for x, y in zip(a.get_packed(), b.get_packed())) and all(a.series(r) == b.series(r) for r in range(0, R, max(1, R // 16)))
    ea, eb = a.events(), b.events()
    print("mode %d, %d replicas x %d steps, %s vs %s: bit-identical %s; events %d on / %d off (general: %d / %d)" % (mode, R, N, a.path(), b.path(), same, ea["rl_on"], ea["rl_off"], eb["rl_on"], eb["rl_off"]), flush=True)
    assert same and ea["rl_on"] == eb["rl_on"] and ea["rl_off"] == eb["rl_off"]
    a.close(); b.close()
