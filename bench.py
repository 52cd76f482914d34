#!/usr/bin/env python3
"""bench.py -- throughput of the per-timestep KMC sweep (main.cpp:461-2202) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--molecules M] [--mc-steps S] [--workload membrane|ensemble1024] [--impl reference]

metric   molecule-moves/s = molecules x MC steps / seconds, whole job (all GPUs)
"step"   one batch of S (default 100) MC steps of the sweep over the membrane resident on the GPUs: state in, S steps, state and
         records out (the reference's own batch is 5000 steps between two output records)
workload per GPU a membrane patch of M molecules (3:1 receptors:ligands, default densities and parameters,
         fresh non-overlapping random start), default M = 1.25e6 = the per-GPU share of the 1e7-molecule
         membrane of BASELINE.json configs[4] on 8 GPUs; working set (> 400 MB incl. neighbour grid) exceeds
         the 126 MB L2, so no L2 flush is needed between iterations.
         N > 1: ONE membrane of N x M molecules cut into N strips along x; the halo refresh (classify, pack, NCCL send/recv, merge) is
         enqueued by the library on its own stream every --refresh-every steps, inside the timed region.
value    device time (CUDA events on the library's own stream, N = 1 and N > 1 alike), max over ranks
e2e      the same batch through the C ABI with HOST buffers (pinned): N = 1 kmc_set_packed + kmc_step + kmc_get_packed_async + kmc_get_series
         per batch, pipelined (the download of batch i overlaps batch i + 1; all transfers inside the timed region);
         N > 1 every rank moves its own slab: kmc_strip_load_records + kmc_step + kmc_strip_get_records + kmc_strip_get_series
         (the bond.dat row of the whole membrane, all-reduced); host wall clock, max over ranks
roofline dominant kernel: its algorithmic bytes per molecule x molecules / its CUDA-event time
cpu_baseline the UNMODIFIED reference (oracle/_ref/kmcref_n200_shipped = main.cpp, only `main` renamed) on one
         host core, default 150+50 system, bounded sample
extras   (N = 1, default workload) short measurements of the other named regimes: production (checkerboard) order, an oligomerised
         state of the same 1.25e6 molecules, the 1e5-molecule config, the 1024-replica ensemble
strips_check (N > 1) a 1e5-molecule hot membrane on the live NCCL ranks compared bit for bit with the single-GPU run
--impl reference: the unmodified reference on ALL host cores (one independent copy per core, the only
         parallelism the serial program allows), same metric.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "kmc-with-a-diffusion-reaction-algorithm_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# algorithmic bytes per molecule and launch of each kernel (DESIGN.md "Kernels"), 3:1 receptor:ligand mix
B_ALG_STEP = 232.0      # whole step, fp64 pose read+written once (2*96) + bond words 24 + label 8 + cell key/index 8
# algorithmic bytes per launch of a kernel, per molecule of the system (3:1 receptors:ligands) unless the kernel works on a list
B_ALG_KERNEL = {
    "k_propose_rec": 0.75 * (48 + 48 + 4 + 4),        # receptors: old pose in, proposed pose out, unit word, cis word
    "k_propose_lig": 0.25 * (192 + 192 + 4 + 4),      # ligands: old pose in, proposed pose out, unit word, complex-size word
    "k_propose": 0.75 * (48 + 48 + 4 + 4) + 0.25 * (192 + 192 + 4 + 4),      # the fused proposal kernel: both of the above
    "k_resolve_tiles": 81,                            # per entry: id 4, unit 4, far 1, centres old+new 32, bond words 8-12 (+72 B beads for a ligand probe)
                                                      # = 49 B receptor / 121 B ligand -> 67 B in the 3:1 mix, + cellStart window 2.56 cells x 1.34 x 4 B
    "k_cells_cut": 12 + 6 * 4 * 0.5 + 3.6 * 12,       # per entry: own id/centre/cell, row extents (shared between neighbours), ~3.6 candidates x 12 B
    "k_grid_scatter": 18 + 4 + 4 + 4,
    "k_finish": 4 + 4 + 0.75 * 8,                     # unit word + unit result, bond words of the receptors
}
B_ALG_PER_LIST_PAIR = 8 + 2 * 40                      # k_pairs_eval: the pair + two records (centres old/new 32, unit key 4, flags 4)


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the top kernels on the default workload (ncu --set full), profiles/"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {"k_propose_lig": 65.0e6 + 28.5e6, "k_propose_rec": 60.0e6 + 42.5e6, "k_pairs_eval": 49.3e6 + 0.1e6, "k_resolve_tiles": 154.8e6 + 5.7e6}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0])); self.max_mhz = float(out[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def time_reference_once(mc_steps, shipped=True, tag="n200", n=200):
    """runs the unmodified reference for mc_steps steps on one core; returns (moves/s, seconds of the sweep)"""
    exe = os.path.join(ROOT, "oracle", "_ref", "kmcref_%s%s" % (tag, "_shipped" if shipped else ""))
    with tempfile.TemporaryDirectory() as td:
        p = subprocess.run([exe, "--steps", str(mc_steps), "--workdir", os.path.join(td, "wd")], cwd=td, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(p.stderr)
        s = json.loads(p.stdout.strip().splitlines()[-1])
    return n * mc_steps / s["seconds"], s["seconds"]


def reference_arm(args):
    """the reference's own CPU implementation on all host cores: one independent copy of the serial program per core"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "kmcref_n200_shipped")
    cores = os.cpu_count() or 1
    s_ref = args.ref_mc_steps
    times = []
    for it in range(args.warmup + args.steps):
        with tempfile.TemporaryDirectory() as td:
            t0 = time.perf_counter()
            procs = [subprocess.Popen([exe, "--steps", str(s_ref), "--workdir", os.path.join(td, "wd%d" % c)], cwd=td,
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for c in range(cores)]
            for p in procs:
                p.wait()
            dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = cores * 200 * s_ref * len(times) / total
    line = {"impl": "reference", "metric": "molecule-moves/s", "value": value, "unit": "molecule-moves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "unmodified main.cpp default system (150 receptors + 50 ligands, paper parameters), one independent copy per host core; "
                                   "the reference cannot run the 1.25e6-molecule patch (O(N^2) time, (N+1)^2-int results matrix)",
                       "mc_steps_per_step": s_ref, "molecules_per_copy": 200, "copies": cores},
            "cpu_baseline": {"value": value, "unit": "molecule-moves/s", "cores": cores, "kind": "reference",
                             "sample": "%d copies x %d MC steps x %d timed batches of the default 200-molecule system, rand2 as shipped" % (cores, s_ref, len(times))},
            "e2e": {"value": value, "unit": "molecule-moves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def top_kernels(k, steps, n=5):
    """per-kernel CUDA-event time (us per MC step) of `steps` steps"""
    k.profile(True); k.step(steps); k.sync(); prof = k.profile_get(); k.profile(False)
    return {name: round(1e3 * v[0] / steps, 2) for name, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:n] if v[1]}


def extras_single_gpu(kmc_b200, args, local, M, na, nb):
    """the other named regimes, measured briefly on this GPU (each: warm-up, CUDA-event timing, top kernels)"""
    import numpy as np
    out = {}
    box = kmc_b200.scaled_box(M)
    # production mode (checkerboard sweep order, KMC_MODE_PRODUCTION): same membrane, same kernels
    try:
        k = kmc_b200.Kmc(kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=args.seed, mode=kmc_b200.MODE_PRODUCTION, device=local))
        k.init_random(seed=args.seed, sort_cells=True)
        k.step(120); k.sync()
        ms = k.step_timed(300) / 300
        ev = k.events()
        out["production"] = {"value": M / (ms * 1e-3), "unit": "molecule-moves/s", "ms_per_mc_step": ms, "molecules": M,
                             "order": "checkerboard (2x2 colours of the neighbour-grid cells), index order inside a colour",
                             "pending_findings_last_step": ev["pending_findings"], "list_pairs": ev["list_pairs"], "special_entries": ev["special_entries"]}
        k.close()
    except Exception as ex:
        out["production"] = {"error": str(ex)[:200]}
    # oligomerised state of the same membrane: pre-assembled complexes at the default density, then evolved
    try:
        from kmc_b200.synth import oligomerised_state
        p = kmc_b200.default_params(box=box, n_receptor=na, n_ligand=nb, seed=args.seed, device=local)
        st = oligomerised_state(p, seed=args.seed, bound_fraction=0.6)
        k = kmc_b200.Kmc(p)
        k.set_packed(*st)
        k.step(120); k.sync()
        ms = k.step_timed(300) / 300
        s, ev = k.series(), k.events()
        hist = k.oligomer_hist(nbins=16)
        out["oligomerised"] = {"value": M / (ms * 1e-3), "unit": "molecule-moves/s", "ms_per_mc_step": ms, "molecules": M,
                               "state": "pre-assembled aligned complexes (1-2 ligands, 1-6 receptors, cis partners) at the default density, evolved 420 steps",
                               "ligands_bound_fraction": float(1.0 - hist[1] / nb), "bonds": s["bond_num"], "complexes": s["n_complexes"], "max_complex": s["max_complex"],
                               "oligomer_hist": {str(i): int(c) for i, c in enumerate(hist) if c}, "adapted_rebuild_every_step": bool(ev["special_entries"] == 0),
                               "top_kernels_us": top_kernels(k, 60)}
        k.close()
    except Exception as ex:
        out["oligomerised"] = {"error": str(ex)[:200]}
    # configs[1]: 1e5 molecules on one GPU (L2 resident, launch-latency bound)
    try:
        k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(100000), n_receptor=75000, n_ligand=25000, seed=args.seed, device=local))
        k.init_random(seed=args.seed, sort_cells=True)
        k.step(200); k.sync()
        ms = k.step_timed(1000) / 1000
        out["config_1e5"] = {"value": 1e5 / (ms * 1e-3), "unit": "molecule-moves/s", "ms_per_mc_step": ms, "molecules": 100000}
        k.close()
    except Exception as ex:
        out["config_1e5"] = {"error": str(ex)[:200]}
    # the whole 1e7-molecule membrane of configs[3] / configs[4] on ONE GPU (it fits: 1.5 GB of state): the streaming regime, where the
    # latency-bound tail of the step (nine dependent list kernels) is amortised over 8x more molecules than in the headline workload
    try:
        M7 = 10000000
        k = kmc_b200.Kmc(kmc_b200.default_params(box=kmc_b200.scaled_box(M7), n_receptor=3 * M7 // 4, n_ligand=M7 // 4, seed=args.seed, device=local))
        k.init_random(seed=args.seed, sort_cells=True)
        k.step(24); k.sync()
        ms = k.step_timed(120) / 120
        peak, _ = read_peaks()
        out["membrane_1e7_one_gpu"] = {"value": M7 / (ms * 1e-3), "unit": "molecule-moves/s", "ms_per_mc_step": ms, "molecules": M7,
                                       "step_frac_of_hbm_roofline": M7 / (ms * 1e-3) * B_ALG_STEP / 1e9 / peak}
        k.close()
    except Exception as ex:
        out["membrane_1e7_one_gpu"] = {"error": str(ex)[:200]}
    # configs[2]: 1024 replicas of the default system in one handle
    try:
        out["ensemble1024"] = ensemble_measure(kmc_b200, 1024, local, args.seed, steps=2000)
    except Exception as ex:
        out["ensemble1024"] = {"error": str(ex)[:200]}
    return out


def ensemble_measure(kmc_b200, replicas, local, seed, steps=2000):
    k = kmc_b200.Kmc(kmc_b200.default_params(n_replicas=replicas, seed=seed, device=local))
    k.init_random(seed=seed + 1)
    k.step(200); k.sync()
    ms = k.step_timed(steps) / steps
    out = {"value": 200 * replicas / (ms * 1e-3), "unit": "molecule-moves/s", "us_per_mc_step": 1e3 * ms, "replicas": replicas, "molecules_per_replica": 200,
           "step_path": k.path() + (" (one CTA per replica, the whole step in one kernel, %d steps per launch)" % steps if k.path() == "fused" else "")}
    k.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--molecules", type=int, default=1250000, help="molecules per GPU (3:1 receptors:ligands)")
    ap.add_argument("--mc-steps", type=int, default=100, help="MC steps per bench step (the reference writes its records every 5000 steps: main.cpp:2206)")
    ap.add_argument("--ref-mc-steps", type=int, default=100)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cell-edge", type=float, default=0.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--workload", default="membrane", choices=["membrane", "ensemble1024"],
                    help="ensemble1024: BASELINE configs[2], 1024 replicas of the default system dealt to the GPUs (no communication)")
    ap.add_argument("--decomp", default="auto", choices=["auto", "strips", "patches"],
                    help="N>1: 'strips' = ONE membrane of N*M molecules cut into N strips along x, halos refreshed over NCCL; "
                         "'patches' = N independent membranes of M molecules (no communication); auto = strips, patches if that fails")
    ap.add_argument("--refresh-every", type=int, default=24, help="strips: MC steps between halo refreshes")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import kmc_b200

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = "cuda:%d" % local
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.workload == "ensemble1024":
        return ensemble_workload(args, kmc_b200, torch, dist, rank, world, local, allmax)

    M = args.molecules
    na, nb = (3 * M) // 4, M - (3 * M) // 4
    S = args.mc_steps
    decomp = "patches" if world == 1 else args.decomp
    strips_note, ds, check = None, None, None
    if decomp in ("auto", "strips"):
        try:
            from kmc_b200.strips import DistStrips, nccl_check
            check = nccl_check(dist, local)                                # correctness of the live NCCL path, attached to the throughput
            every = args.refresh_every
            S = ((S + every - 1) // every) * every                      # whole refresh intervals per bench step
            G = M * world
            gna, gnb = (3 * G) // 4, G - (3 * G) // 4
            gbox = kmc_b200.scaled_box(G)
            pg = kmc_b200.default_params(box=gbox, n_receptor=gna, n_ligand=gnb, seed=args.seed)
            halo = kmc_b200.strip_halo_width(pg, every, 400.0)
            frac = (gbox[0] / world + 2 * halo) / (gbox[0] / world)
            p = kmc_b200.default_params(box=gbox, n_receptor=int(gna / world * frac * 1.04) + 2000, n_ligand=int(gnb / world * frac * 1.04) + 2000,
                                        seed=args.seed, mode=kmc_b200.MODE_REPLAY, device=local, cell_edge=args.cell_edge)
            ds = DistStrips(p, every, halo_width=halo, dist=dist)
            t_init = time.perf_counter()
            if hasattr(ds.k, "strip_init_random"):
                ds.k.strip_init_random(gna, gnb, seed=args.seed)           # generated on the GPU; every rank keeps its slab
            else:
                grec, glig = kmc_b200.generate_packed(pg, seed=args.seed, sort_cells=True)      # every rank: the same global start state
                ds.load_global(grec, glig)
                del grec, glig
            ds.k.sync()
            t_init = time.perf_counter() - t_init
            k = ds.k
            decomp = "strips"
            strips_note = "ONE %d-molecule membrane (L=%.0f A) cut into %d strips along x; halo %.0f A refreshed every %d steps by the library (C++ ncclSend/ncclRecv on its own stream, no host synchronisation)" % (
                G, gbox[0], world, halo, every)
        except Exception as ex:            # keep the run alive: fall back to independent patches and say why
            if args.decomp == "strips":
                raise
            strips_note = "strips failed (%s): fell back to independent patches" % str(ex)[:200]
            decomp, ds = "patches", None
        # all ranks must take the same path (a rank falling back alone would leave the others waiting in NCCL)
        okt = torch.tensor([1 if ds is not None else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0 and ds is not None:
            ds.k.close(); ds = None; decomp = "patches"
            strips_note = "strips failed on another rank: fell back to independent patches"
    if decomp == "patches":
        p = kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=na, n_ligand=nb, seed=args.seed + 1000 * rank,
                                    mode=kmc_b200.MODE_REPLAY, device=local, cell_edge=args.cell_edge)
        k = kmc_b200.Kmc(p)
        t_init = time.perf_counter()
        k.init_random(seed=args.seed + 7919 * rank, sort_cells=True)
        k.sync()
        t_init = time.perf_counter() - t_init

    def barrier():
        if dist is not None:
            dist.barrier()
        k.sync()

    for _ in range(args.warmup):
        k.step(S)
    # ---- timed region: EXACTLY K steps, barrier + sync on both sides; device time from CUDA events on the library's stream
    # (with strips the refreshes -- kernels and NCCL operations -- are enqueued on that stream by kmc_step itself) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0 = k.events()
    barrier()
    ms = 0.0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        ms += k.step_timed(S)
    barrier()
    t_wall = time.perf_counter() - t_wall
    ev1 = k.events()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    ms_max = allmax(ms)
    wall_max = allmax(t_wall * 1e3)
    value = world * M * S * args.steps / (ms_max * 1e-3)

    # ---- per-kernel CUDA-event timing (separate, untimed pass: events around every launch) ----
    k.profile(True)
    for _ in range(3):
        k.step(S)
    k.sync()
    prof = k.profile_get()
    k.profile(False)
    tot_prof = sum(v[0] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][0])
    peak, peak_src = read_peaks()
    top_name, (top_ms, top_n) = top
    evp = k.events()
    n_local = M if ds is None else (ev_live(k) or M)
    balg = B_ALG_KERNEL.get(top_name)
    if top_name == "k_pairs_eval":
        balg = B_ALG_PER_LIST_PAIR * evp["list_pairs"] / n_local
    achieved = balg * n_local / (top_ms / top_n * 1e-3) / 1e9 if balg else None

    # ---- e2e through the C ABI with host buffers ----
    e2e_t = []
    n_e2e = 2 + max(3, args.steps // 2)
    if ds is None:
        rec, lig, rl, rs, rc = k.get_packed()
        pin = [torch.from_numpy(a).pin_memory().numpy() for a in (rec, lig, rl, rs, rc)]          # inputs, pinned host memory
        pout = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (rec, lig, rl, rs, rc)]  # results, pinned host memory
        h2d = sum(a.nbytes for a in pin); d2h = h2d + 64
        # pipelined, as an application that writes records every batch would run it: the result of batch i crosses PCIe on the
        # library's copy stream (kmc_get_packed_async: device-side snapshot first) while batch i + 1 is uploaded and computed; the
        # timed region holds every upload, every step and every download of its batches (the last download is waited for inside it)
        for it in range(2):
            k.set_packed(*pin, step_done=1000 + it * S); k.step(S); k.series(); k.get_packed_async(pout)
        k.snapshot_wait()
        barrier()
        t0 = time.perf_counter()
        for it in range(2, n_e2e):
            k.set_packed(*pin, step_done=1000 + it * S)
            k.step(S)
            k.series()                       # (the 64-byte record first: a small download queued behind the 116 MB one would wait for it)
            k.get_packed_async(pout)
        k.snapshot_wait()
        e2e_t.append((time.perf_counter() - t0) / (n_e2e - 2))
    else:       # strips: every rank moves its own slab (owned units + halo copies in, owned units out) and gets the global bond.dat row
        cap = 64 * p.n_receptor + 208 * p.n_ligand
        bin_ = torch.zeros(cap, dtype=torch.uint8).pin_memory().numpy(); bout = torch.zeros(cap, dtype=torch.uint8).pin_memory().numpy()
        nr, nl = k.strip_get_records(3, bin_)
        h2d = 64 * nr + 208 * nl; d2h = 0
        for it in range(n_e2e):
            barrier()
            t0 = time.perf_counter()
            k.strip_load_records(bin_, nr, nl, step_done=1000 + it * S)
            k.step(S)
            onr, onl = k.strip_get_records(2, bout)
            row = ds.series()
            dt = time.perf_counter() - t0
            d2h = 64 * onr + 208 * onl + 64
            if it >= 2:
                e2e_t.append(dt)
    e2e_val = world * M * S / allmax(sum(e2e_t) / len(e2e_t))
    h2d_all = allmax(float(h2d)); d2h_all = allmax(float(d2h))

    extras = None
    if rank == 0 and world == 1 and M == 1250000 and not args.no_extras:
        k.close()
        extras = extras_single_gpu(kmc_b200, args, local, M, na, nb)

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                v, sec = time_reference_once(600)
                cpu = {"value": v, "unit": "molecule-moves/s", "cores": 1, "kind": "reference",
                       "sample": "unmodified main.cpp (oracle/_ref/kmcref_n200_shipped: only `main` renamed, rand2 as shipped), default 150+50 system, 600 MC steps in %.1f s" % sec}
                v2, _ = time_reference_once(3000, shipped=False)
                cpu["extra_rng_fixed"] = {"value": v2, "note": "same object with rand2 interposed by a cheap xorshift stream (labelled extra, not the baseline)"}
            except Exception as ex:  # the prebuilt reference did not travel
                cpu = {"value": None, "unit": "molecule-moves/s", "cores": 1, "kind": "reference", "sample": "unavailable: %s" % ex}
        traffic = ncu_traffic()
        line = {"metric": "molecule-moves/s", "value": value, "unit": "molecule-moves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "%d-molecule membrane patch per GPU (%d receptors + %d ligands, default density L=%.0f A, paper parameters, "
                                       "fresh random start); %s" % (M, na, nb, kmc_b200.scaled_box(M)[0],
                                                                  "per-GPU share of the 1e7-molecule membrane (configs[4])" if M == 1250000 else "custom size"),
                           "mc_steps_per_step": S, "mode": "replay (index-order sweep, keyed Philox)",
                           "neighbour_list": "grid + pair list rebuilt every %s steps, reused in between (exact: far movers / drifted molecules are special entries)" % os.environ.get("KMC_REUSE", "6"), "l2": "working set > L2 (no flush needed)",
                           "decomposition": strips_note if decomp == "strips" else ("independent patches per GPU, no data-path collective" + ("; " + strips_note if strips_note else "")),
                           "timing": "CUDA events on the library stream (N = 1 and N > 1 alike); host wall clock between the barriers: %.3f ms per step" % (wall_max / args.steps),
                           "ms_per_mc_step": ms_max / args.steps / S, "init_s": round(t_init, 3),
                           "reference_arm_note": "the reference cannot run this workload (O(N^2) time and memory): its arm runs the default 200-molecule system, so the ratio compares different N"},
                "roofline": {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": achieved / peak if achieved else None,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, ncu --set full (profiles/)
                             "traffic": traffic.get(top_name) if M == 1250000 else None,
                             "alg_bytes_per_molecule": balg, "kernel_ms": top_ms / top_n, "kernel_share_of_step": top_ms / tot_prof,
                             "step_frac": value / world * B_ALG_STEP / 1e9 / peak,
                             "kernels_ms_per_mc_step": {n: round(v[0] / (3 * S), 4) for n, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[1]}},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "molecule-moves/s", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                        "note": "bytes per rank (max over ranks); every rank moves its own slab" if ds is not None else "whole state in and out every batch; pipelined: the download of batch i (kmc_get_packed_async) overlaps batch i + 1, all transfers inside the timed region"},
                "gpu_launches": ev1["launches"] - ev0["launches"],
                "work_lists_last_step": {q: evp[q] for q in ("list_pairs", "special_entries", "pending_findings", "reaction_pairs")},
                "events_in_timed_region": {q: ev1[q] - ev0[q] for q in ("rl_on", "mono_cis_on", "cis_on", "rl_off", "reverted", "rebuilds")},
                "clocks": sampler.result() if sampler else None}
        if check is not None:
            line["strips_check"] = check
        if ds is not None:
            line["series_whole_membrane"] = {q: row[q] for q in ("step", "bond_num_rl", "bond_num_mono_cis", "bond_num_cis", "bond_num", "n_complexes", "max_complex")}
        if extras is not None:
            line["extras"] = extras
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ev_live(k):
    try:
        return k.live_counts()[0] + k.live_counts()[1]
    except Exception:
        return None


def ensemble_workload(args, kmc_b200, torch, dist, rank, world, local, allmax):
    """BASELINE configs[2]: 1024 replicas of the default system, contiguous blocks of replica ids per GPU, no communication"""
    from kmc_b200.sharding import replica_range
    total = 1024
    lo, hi = replica_range(rank, world, total)
    S = max(args.mc_steps, 2000)
    k = kmc_b200.Kmc(kmc_b200.default_params(n_replicas=hi - lo, seed=args.seed + lo, device=local))
    k.init_random(seed=args.seed + 1 + lo)
    for _ in range(args.warmup):
        k.step(S)
    if dist is not None:
        dist.barrier()
    k.sync()
    ev0 = k.events()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = sum(k.step_timed(S) for _ in range(args.steps))
    k.sync()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    if dist is not None:
        dist.barrier()
    ev1 = k.events()
    ms_max = allmax(ms)
    value = 200.0 * total * S * args.steps / (ms_max * 1e-3)
    # e2e: reference-shaped state of every replica in, S steps, series of every replica out
    states = [k.get_state(r) for r in range(min(hi - lo, 8))]
    t0 = time.perf_counter()
    for r, st in enumerate(states):
        k.set_state(*st, replica=r)
    k.step(S)
    rows = [k.series(r) for r in range(hi - lo)]
    e2e_s = allmax(time.perf_counter() - t0)
    if rank == 0:
        peak, peak_src = read_peaks()
        line = {"metric": "molecule-moves/s", "value": value, "unit": "molecule-moves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "ensemble1024: 1024 independent replicas of the reference's default system (150 receptors + 50 ligands, paper parameters), "
                                       "%d per GPU, no inter-GPU communication (BASELINE configs[2])" % (hi - lo), "mc_steps_per_step": S,
                           "us_per_mc_step": 1e3 * ms_max / args.steps / S, "step_path": k.path(),
                           "l2": "state resident in shared memory for the length of a launch (fused step: one CTA per replica, the whole time step in one kernel, %d steps per launch); not an HBM-bound workload" % S,
                           "timing": "CUDA events on the library stream"},
                "roofline": {"bound": "hbm", "kernel": "whole step", "achieved": value * B_ALG_STEP / 1e9, "peak": peak * world, "peak_source": peak_src, "unit": "GB/s",
                             "frac": value * B_ALG_STEP / 1e9 / (peak * world), "traffic": None},
                "cpu_baseline": None,
                "e2e": {"value": 200.0 * total * S / e2e_s, "unit": "molecule-moves/s", "h2d_bytes_per_step": int(len(states) * 3 * 25 * 201 * 8), "d2h_bytes_per_step": int(64 * (hi - lo)),
                        "note": "kmc_set_state of 8 replicas + S steps + kmc_get_series of every replica"},
                "gpu_launches": ev1["launches"] - ev0["launches"], "clocks": sampler.result() if sampler else None,
                "series_replica0": rows[0]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
