"""what the asynchronous snapshot (kmc_get_packed_async) overlaps with: uploads, steps, the small series record"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "kmc-with-a-diffusion-reaction-algorithm_b200"))
import kmc_b200, torch, numpy as np
M = 1250000
k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=3 * M // 4, n_ligand=M // 4, seed=1))
k.init_random(seed=1, sort_cells=True); k.step(60); k.sync()
st = k.get_packed()
pin = [torch.from_numpy(a).pin_memory().numpy() for a in st]
pout = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in st]
def T(f, n=5):
    k.sync(); k.snapshot_wait(); t0 = time.perf_counter()
    for _ in range(n): f()
    k.sync(); k.snapshot_wait(); return 1e3 * (time.perf_counter() - t0) / n
print("set_packed            %.2f ms" % T(lambda: k.set_packed(*pin)))
print("step(100)             %.2f ms" % T(lambda: k.step(100)))
print("get_packed (sync)     %.2f ms" % T(lambda: k.get_packed(out=pout)))
print("get_packed_async+wait %.2f ms" % T(lambda: (k.get_packed_async(pout), k.snapshot_wait())))
def call_only():
    t0 = time.perf_counter(); k.get_packed_async(pout); return time.perf_counter() - t0
k.sync(); print("async call returns in %.3f ms" % (1e3 * call_only())); k.snapshot_wait()
print("step(100) + async     %.2f ms (overlap => ~ step alone)" % T(lambda: (k.get_packed_async(pout), k.step(100))))
print("set+step+async        %.2f ms" % T(lambda: (k.set_packed(*pin), k.step(100), k.get_packed_async(pout))))
print("set+step+async+series %.2f ms" % T(lambda: (k.set_packed(*pin), k.step(100), k.get_packed_async(pout), k.series())))
print("set+step+sync get     %.2f ms" % T(lambda: (k.set_packed(*pin), k.step(100), k.get_packed(out=pout))))
print("set+step+series+async %.2f ms" % T(lambda: (k.set_packed(*pin), k.step(100), k.series(), k.get_packed_async(pout))))
