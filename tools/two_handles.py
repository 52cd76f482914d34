"""experiment: do the steps of TWO handles on one GPU overlap (the latency-bound tail of one under the proposals of the other)?
One 1.25e6-molecule membrane against two independent 625 000-molecule membranes stepped side by side on their own streams."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200
import torch

def make(M, seed):
    k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=3 * M // 4, n_ligand=M // 4, seed=seed))
    k.init_random(seed=seed, sort_cells=True); k.step(60); k.sync(); return k

def timed(ks, n):
    for k in ks: k.sync()
    t0 = time.perf_counter()
    for k in ks: k.step(n)
    for k in ks: k.sync()
    return (time.perf_counter() - t0) / n

one = make(1250000, 1)
print("one handle, 1.25e6 molecules: %.1f us per step" % (1e6 * timed([one], 600)), flush=True)
one.close()
for parts in (2, 4):
    ks = [make(1250000 // parts, 10 + i) for i in range(parts)]
    print("%d handles x %d molecules side by side: %.1f us per step of all; one of them alone: %.1f us" % (parts, 1250000 // parts, 1e6 * timed(ks, 600), 1e6 * timed(ks[:1], 600)), flush=True)
    for k in ks: k.close()
