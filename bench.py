#!/usr/bin/env python3
"""bench.py -- throughput of the per-timestep KMC sweep (main.cpp:461-2202) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--molecules M] [--mc-steps S] [--impl reference]

metric   molecule-moves/s = molecules x MC steps / seconds, whole job (all GPUs)
"step"   one batch of S (default 100) MC steps of the sweep over the membrane resident on the GPUs: state in, S steps, state and
         records out (the reference's own batch is 5000 steps between two output records)
workload per GPU a membrane patch of M molecules (3:1 receptors:ligands, default densities and parameters,
         fresh non-overlapping random start), default M = 1.25e6 = the per-GPU share of the 1e7-molecule
         membrane of BASELINE.json configs[4] on 8 GPUs; working set (> 400 MB incl. neighbour grid) exceeds
         the 126 MB L2, so no L2 flush is needed between iterations
value    device time (CUDA events on the library's own stream), max over ranks
e2e      the same batch through the C ABI with HOST buffers: kmc_set_packed (H2D) + kmc_step + kmc_get_packed
         (D2H) + kmc_get_series, host wall clock
roofline dominant kernel: its algorithmic bytes per molecule x molecules / its CUDA-event time
cpu_baseline the UNMODIFIED reference (oracle/_ref/kmcref_n200_shipped = main.cpp, only `main` renamed) on one
         host core, default 150+50 system, bounded sample
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
    "k_resolve_tiles": 81,                            # per entry: id 4, unit 4, far 1, centres old+new 32, bond words 8-12 (+72 B beads for a ligand probe)
                                                      # = 49 B receptor / 121 B ligand -> 67 B in the 3:1 mix, + cellStart window 2.56 cells x 1.34 x 4 B
    "k_cells_cut": 12 + 6 * 4 * 0.5 + 3.6 * 12,       # per entry: own id/centre/cell, row extents (shared between neighbours), ~3.6 candidates x 12 B
    "k_grid_scatter": 18 + 4 + 4 + 4,
    "k_finish": 4 + 4 + 0.75 * 8,                     # unit word + unit result, bond words of the receptors
}
NCU_TRAFFIC = {"k_propose_lig": 65.0e6 + 28.5e6, "k_propose_rec": 60.0e6 + 42.5e6, "k_pairs_eval": 49.3e6 + 0.1e6, "k_resolve_tiles": 154.8e6 + 5.7e6}
B_ALG_PER_LIST_PAIR = 8 + 2 * 40                      # k_pairs_eval: the pair + two records (centres old/new 32, unit key 4, flags 4)


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
    ap.add_argument("--decomp", default="auto", choices=["auto", "strips", "patches"],
                    help="N>1: 'strips' = ONE membrane of N*M molecules cut into N strips along x, halos refreshed over NCCL; "
                         "'patches' = N independent membranes of M molecules (no communication); auto = strips, patches if that fails")
    ap.add_argument("--refresh-every", type=int, default=32, help="strips: MC steps between halo refreshes")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    import kmc_b200

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    M = args.molecules
    na, nb = (3 * M) // 4, M - (3 * M) // 4
    S = args.mc_steps
    decomp = "patches" if world == 1 else args.decomp
    strips_note, ds = None, None
    if decomp in ("auto", "strips"):
        try:
            from kmc_b200.strips import DistStrips, halo_for
            every = args.refresh_every
            S = ((S + every - 1) // every) * every                      # whole refresh intervals per bench step
            G = M * world
            gna, gnb = (3 * G) // 4, G - (3 * G) // 4
            gbox = kmc_b200.scaled_box(G)
            pg = kmc_b200.default_params(box=gbox, n_receptor=gna, n_ligand=gnb, seed=args.seed)
            grec, glig = kmc_b200.generate_packed(pg, seed=args.seed, sort_cells=True)      # every rank: the same global start state
            halo = halo_for(every)
            frac = (gbox[0] / world + 2 * halo) / (gbox[0] / world)
            p = kmc_b200.default_params(box=gbox, n_receptor=int(gna / world * frac * 1.08) + 2000, n_ligand=int(gnb / world * frac * 1.08) + 2000,
                                        seed=args.seed, mode=kmc_b200.MODE_REPLAY, device=local, cell_edge=args.cell_edge)
            ds = DistStrips(p, every, halo_width=halo, dist=dist)
            ds.load_global(grec, glig)
            k = ds.k
            decomp = "strips"
            strips_note = "ONE %d-molecule membrane (L=%.0f A) cut into %d strips along x; halo %.0f A refreshed every %d steps, boundary bands over NCCL send/recv" % (
                G, gbox[0], world, halo, every)
        except Exception as ex:            # keep the run alive: fall back to independent patches and say why
            if args.decomp == "strips":
                raise
            strips_note = "strips failed (%s): fell back to independent patches" % str(ex)[:200]
            decomp, ds = "patches", None
        # all ranks must take the same path (a rank falling back alone would leave the others waiting in NCCL)
        okt = torch.tensor([1 if ds is not None else 0], dtype=torch.int64, device="cuda:%d" % local)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) == 0 and ds is not None:
            ds.k.close(); ds = None; decomp = "patches"
            strips_note = "strips failed on another rank: fell back to independent patches"
    if decomp == "patches":
        p = kmc_b200.default_params(box=kmc_b200.scaled_box(M), n_receptor=na, n_ligand=nb, seed=args.seed + 1000 * rank,
                                    mode=kmc_b200.MODE_REPLAY, device=local, cell_edge=args.cell_edge)
        k = kmc_b200.Kmc(p)
        k.init_random(seed=args.seed + 7919 * rank, sort_cells=True)
    stepper = ds if ds is not None else k

    def barrier():
        if dist is not None:
            dist.barrier()
        k.sync()

    for _ in range(args.warmup):
        stepper.step(S)
    # ---- timed region: EXACTLY K steps, barrier + sync on both sides. Patches: device time from CUDA events on the library's
    # stream. Strips: host wall clock between the barriers (the halo refresh is host-orchestrated: NCCL + rebuild kernels) ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ev0 = k.events()
    barrier()
    ms = 0.0
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        if ds is None:
            ms += k.step_timed(S)
        else:
            ds.step(S)
    barrier()
    t_wall = time.perf_counter() - t_wall
    if ds is not None:
        ms = t_wall * 1e3
    ev1 = k.events()
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda:%d" % local)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * M * S * args.steps / (ms_max * 1e-3)

    # ---- per-kernel CUDA-event timing (separate, untimed pass: events around every launch) ----
    k.profile(True)
    for _ in range(3):
        stepper.step(S)
    k.sync()
    prof = k.profile_get()
    k.profile(False)
    tot_prof = sum(v[0] for v in prof.values())
    top = max(prof.items(), key=lambda kv: kv[1][0])
    peak, peak_src = read_peaks()
    top_name, (top_ms, top_n) = top
    evp = k.events()
    balg = B_ALG_KERNEL.get(top_name)
    if top_name == "k_pairs_eval":
        balg = B_ALG_PER_LIST_PAIR * evp["list_pairs"] / M
    achieved = balg * M / (top_ms / top_n * 1e-3) / 1e9 if balg else None

    # ---- e2e through the C ABI with host buffers ----
    e2e_t = []
    if ds is None:
        rec, lig, rl, rs, rc = k.get_packed()
        pin = [torch.from_numpy(a).pin_memory().numpy() for a in (rec, lig, rl, rs, rc)]          # inputs, pinned host memory
        pout = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in (rec, lig, rl, rs, rc)]  # results, pinned host memory
        h2d = sum(a.nbytes for a in pin); d2h = h2d + 64
        for it in range(2 + max(3, args.steps // 2)):
            barrier()
            t0 = time.perf_counter()
            k.set_packed(*pin, step_done=1000 + it * S)
            k.step(S)
            k.get_packed(out=pout)
            k.series()
            dt = time.perf_counter() - t0
            if it >= 2:
                e2e_t.append(dt)
    else:       # strips: host arrays of the global state in, each rank's owned molecules out
        h2d = d2h = 0
        for it in range(2 + max(3, args.steps // 2)):
            barrier()
            t0 = time.perf_counter()
            ds.load_global(grec, glig, step_done=1000 + it * S)
            ds.step(S)
            k.strip_begin_refresh()
            owned = k.strip_message(2)
            dt = time.perf_counter() - t0
            h2d = d2h = len(owned)
            if it >= 2:
                e2e_t.append(dt)
    te = torch.tensor([sum(e2e_t) / len(e2e_t)], dtype=torch.float64, device="cuda:%d" % local)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * M * S / float(te.item())

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
        line = {"metric": "molecule-moves/s", "value": value, "unit": "molecule-moves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "%d-molecule membrane patch per GPU (%d receptors + %d ligands, default density L=%.0f A, paper parameters, "
                                       "fresh random start); %s" % (M, na, nb, p.box[0],
                                                                  "per-GPU share of the 1e7-molecule membrane (configs[4])" if M == 1250000 else "custom size"),
                           "mc_steps_per_step": S, "mode": "replay (index-order sweep, keyed Philox)",
                           "neighbour_list": "grid + pair list rebuilt every %s steps, reused in between (exact: far movers / drifted molecules are special entries)" % os.environ.get("KMC_REUSE", "6"), "l2": "working set > L2 (no flush needed)",
                           "decomposition": strips_note if decomp == "strips" else ("independent patches per GPU, no data-path collective" + ("; " + strips_note if strips_note else "")),
                           "timing": "host wall clock between barriers (halo refresh is host-orchestrated)" if decomp == "strips" else "CUDA events on the library stream",
                           "ms_per_mc_step": ms_max / args.steps / S},
                "roofline": {"bound": "hbm", "kernel": top_name, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                             "frac": achieved / peak if achieved else None,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this workload, ncu --set full
                             # (profiles/r01e_ncu_full_raw.csv; k_resolve_tiles: profiles/r01b_ncu_full_raw.csv)
                             "traffic": NCU_TRAFFIC.get(top_name) if M == 1250000 else None,
                             "alg_bytes_per_molecule": balg, "kernel_ms": top_ms / top_n, "kernel_share_of_step": top_ms / tot_prof,
                             "step_frac": value / world * B_ALG_STEP / 1e9 / peak,
                             "kernels_ms_per_mc_step": {n: round(v[0] / (3 * S), 4) for n, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if v[1]}},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_val, "unit": "molecule-moves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": ev1["launches"] - ev0["launches"],
                "work_lists_last_step": {q: evp[q] for q in ("list_pairs", "special_entries", "pending_findings", "reaction_pairs")},
                "events_in_timed_region": {q: ev1[q] - ev0[q] for q in ("rl_on", "mono_cis_on", "cis_on", "rl_off", "reverted", "rebuilds")},
                "clocks": sampler.result() if sampler else None}
        print(json.dumps(line))
    if ds is not None and getattr(ds, "timing", None) and rank == 0:
        tt = ds.timing
        print("strip refresh timing (ms per refresh): classify+pack %.3f, exchange %.3f, merge %.3f over %d refreshes" % (
            1e3 * tt[0] / tt[3], 1e3 * tt[1] / tt[3], 1e3 * tt[2] / tt[3], tt[3]), file=sys.stderr)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
