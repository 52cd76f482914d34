"""two logical strip ranks of 1.25e6 molecules each on ONE GPU: a few refreshes, for an ncu launch list of the refresh kernels
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/refresh_launches.csv python tools/refresh_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200
from kmc_b200.strips import StripRank
M, world, every = 1250000, 2, 24
G = M * world
gna, gnb = 3 * G // 4, G // 4
gbox = kmc_b200.scaled_box(G)
pg = kmc_b200.default_params(box=gbox, n_receptor=gna, n_ligand=gnb)
halo = kmc_b200.strip_halo_width(pg, every, 400.0)
frac = (gbox[0] / world + 2 * halo) / (gbox[0] / world)
mk = lambda: kmc_b200.default_params(box=gbox, n_receptor=int(gna / world * frac * 1.04) + 2000, n_ligand=int(gnb / world * frac * 1.04) + 2000)
ranks = [StripRank(mk(), r, world, halo) for r in range(world)]
for r in ranks:
    r.k.strip_init_random(gna, gnb, seed=1)
for it in range(3):
    for r in ranks:
        r.k.step(every)
    kmc_b200.strip_refresh_local([r.k for r in ranks], every)
for r in ranks:
    r.k.sync()
print("ok", [r.k.live_counts() for r in ranks])
