// csrc/kmc_small.cu -- the whole time step (main.cpp:461-2202) of a SMALL system as one kernel, many steps per launch
// (BASELINE configs[2]: ensembles of the reference's default 200-molecule system).
//
// The general path (kmc_kernels.cu) spends one CUDA-graph node per stage: 13 dependent nodes per step, ~30 us for a system
// whose arithmetic takes a microsecond. Here the stages of a step are separated by __syncthreads() instead of kernel
// boundaries, a replica never leaves its SM, and nothing crosses between replicas, so a launch covers any number of steps.
//
// It is the SAME step: every stage calls the device functions of kmc_kernels.cu (propose_one_rec, lig_move_*,
// complex_move_thread, pair_eval / publish, pend_sweep, react_pair, react_resolve_block, finish_body) through a per-replica VIEW
// of the device state -- a copy of Args in shared memory whose work lists point at this replica's slices of the global lists and
// whose scalars (list lengths, step counter) live in shared memory. For the length of a launch the replica's poses (both
// buffers) and its bond table live in shared memory too: the view's pose / bond pointers are SHIFTED so that indexing them with
// a molecule's global index lands in the shared arrays -- the device functions neither know nor care.
//
// Only the neighbour search differs: no grid, no far movers, no ghosts. Per molecule an fp32 search record in shared memory (old
// centre, share of the reach + this step's displacement); a pair list built with a drift allowance per molecule and rebuilt when
// too many molecules have left theirs; a molecule that may be outside its allowance this step (periodic wrap, alignment snap) is
// cut against every other molecule by the whole CTA. Pairs within reach are queued and classified by the same pair_eval as
// everywhere else, one directed pair per thread.
//
// Results are bit-identical to the general path (tests/test_gpu_small.py, tools/fused_long_check.py) and to the oracle (every
// small-system test of tests/test_gpu_replay.py runs through this kernel). DESIGN.md section 3b has the measurements.
#pragma once
#ifdef SMALL_TIMING
#include <cstdio>
#endif

namespace kmc {

#ifndef SMALL_T
#define SMALL_T 128           // threads per replica
#endif
#ifndef SMALL_MINB
#define SMALL_MINB 4          // CTAs per SM the register allocation aims at
#endif
#define SMALL_MAXNB 128       // ligands per replica (complex work list)
#define SMALL_LIST 768        // pairs the list of a replica holds
#define SMALL_ITEMS 512       // directed pairs queued for exact classification per step (more: classified in place)
#ifndef SMALL_DMAX
#define SMALL_DMAX 64.0f      // drift (Angstrom) the pair list allows a molecule before it has to be searched on its own (24 / 48 / 64 / 100: 29.8 / 28.0 / 27.9 / 28.2 us per step at 1 024 replicas); halved by a replica whose list overflows
#endif
#ifndef SMALL_SPEC_MAX
#define SMALL_SPEC_MAX 6      // more molecules than this outside their allowance: the list is rebuilt
#endif

struct SmallShared {
    Args view[2];                       // per-replica views of the device state (see above); [1] = the pose buffers swapped (odd steps)
    SmallSearch search;
    int scal[S_COUNT];
    unsigned long long step64;
    union { int touch[TOUCH_CAP]; unsigned long long fastPairs[TOUCH_CAP / 2]; };      // S3 candidates pairs (written by the classification, read by react_pairs_body), then the molecules whose bonds changed (written from react_resolve_block on: only counted here)
    int cxList[SMALL_MAXNB];            // root ligands of the complexes with more than one member | bit 30: several ligands
    unsigned list[SMALL_LIST];          // the pair list: (local index a << 16) | local index b
    unsigned items[SMALL_ITEMS];        // this step's classification work: (probe << 16) | neighbour, both directions of every pair within reach
    int ncx, nlist, listValid, nitems, building;
};

// dynamic shared memory of k_small_step for replicas of NA receptors + NB ligands
__host__ __device__ inline size_t small_dyn_bytes(int NA, int NB) { return ((size_t)NA * 96 + (size_t)NB * 384 + (size_t)(2 * ((NA + 11) & ~3) + NA + 3 * NB) * 4 + 31) & ~(size_t)15; }
// (a CTA that holds `slots` replicas takes slots x (sizeof(SmallShared) + small_dyn_bytes): lists and views are dynamic as well)

KD int small_gid(const Consts &K, int rep, int m) { return m < K.NA ? rep * K.NA + m : K.NAt + rep * K.NB + (m - K.NA); }

// S1 for one replica (main.cpp:514-562): union-find over its bond graph, unit heads, sizes, breadth-first member rows; the member
// rows of replica r live in members[r*N, (r+1)*N)
// (todo = false: a slot of a multi-replica CTA that has nothing to re-derive only keeps the CTA's barriers company)
template <int T> KD void small_rebuild(SmallShared &sm, const Args &V, int rep, int tid, bool todo) {
    const Dev &D = V.D; const Consts &K = V.K;
    const int N = todo ? K.NA + K.NB : 0, NA_ = todo ? K.NA : 0, NB_ = todo ? K.NB : 0;
    for (int m = tid; m < N; m += T) {
        const int gid = small_gid(K, rep, m);
        const int uid = gid < K.NAt ? K.NBt + gid : gid - K.NAt;
        D.ufParent[uid] = uid; D.bfsMark[gid] = 0;
        if (gid >= K.NAt) { D.cxSize[gid - K.NAt] = 0; D.cxOff[gid - K.NAt] = -1; }
    }
    __syncthreads();                    // (every thread has read the touch counters that sent it here)
    if (tid == 0 && todo) { sm.scal[S_MEMBER_CURSOR] = rep * N; sm.ncx = 0; sm.scal[S_NTOUCH] = 0; sm.scal[S_TOPO_DIRTY] = 0; }
    __syncthreads();
    for (int m = tid; m < NA_; m += T) {
        const int a = rep * K.NA + m, ua = K.NBt + a;
        const int l = D.recLig[a];
        if (l >= 0) uf_union(D.ufParent, ua, l);
        const int c = D.recCis[a];
        if (c > a) uf_union(D.ufParent, ua, K.NBt + c);
    }
    __syncthreads();
    for (int m = tid; m < N; m += T) {
        const int gid = small_gid(K, rep, m);
        const int uid = gid < K.NAt ? K.NBt + gid : gid - K.NAt;
        const int r = uf_find(D.ufParent, uid);
        const int head = r < K.NBt ? K.NAt + r : r - K.NBt;
        D.unitOf[gid] = head;
        sm.search.meta[m].x = head;          // (the proposals read the head from here; from the next step on it is the unit key mark_far leaves: same head)
        if (r < K.NBt) atomicAdd(&D.cxSize[r], 1);
    }
    __syncthreads();
    for (int b = tid; b < NB_; b += T) {
        const int h = rep * K.NB + b;
        sm.search.ligFree[b] = D.unitOf[K.NAt + h] == K.NAt + h && D.cxSize[h] <= 1;
        if (D.unitOf[K.NAt + h] != K.NAt + h) continue;
        const int size = D.cxSize[h];
        note_max_complex(K, D, h, size);
        if (size <= 1) continue;
        const int off = atomicAdd(&sm.scal[S_MEMBER_CURSOR], size);
        D.cxOff[h] = off;
        int nlig; build_member_row(K, D, h, size, D.members + off, 1, nlig);
        for (int i = 0; i < size; i++) D.rowWork[off + i] = D.members[off + i];
        sm.cxList[atomicAdd(&sm.ncx, 1)] = h | (nlig == 1 ? 0 : 0x40000000);
    }
    __syncthreads();
}

// S2c for one free ligand, straight between the committed and the new pose buffers, one point at a time
KD void small_propose_lig(const Args &A, uint64_t step, int gid, uint32_t me) {          // me = reference id (1-based number inside the replica)
    KARGS
    const Consts &K = cK;
    const int h = gid - K.NAt;
    if (!D.small->ligFree[h - K.smallRep * K.NB]) return;          // member of a complex: moved by complex_move_thread
    const double *__restrict__ src = D.lig + (size_t)h * 24;
    double *__restrict__ dst = D.lign + (size_t)h * 24;
    const double ox = src[0], oy = src[1], oz = src[2];
    const uint64_t seed = seed_of(K, replica_of_gid(K, gid));
    double u0, u1, u2, u3, u4, u5;
    keyed_uniform2<true>(seed, me, 0, step, 0, u0, u1); keyed_uniform2<true>(seed, me, 0, step, 2, u2, u3); keyed_uniform2<true>(seed, me, 0, step, 4, u4, u5);
    LigMove M; lig_move_setup(K, M, ox, oy, oz, u0, u1, u2, u3, u4, u5);
#pragma unroll 1
    for (int q = 1; q < 8; q++) {
        double o[3]; lig_move_point(M, src[3 * q], src[3 * q + 1], src[3 * q + 2], o);
        dst[3 * q] = o[0]; dst[3 * q + 1] = o[1]; dst[3 * q + 2] = o[2];
    }
    dst[0] = M.c[0]; dst[1] = M.c[1]; dst[2] = M.c[2];
    mark_far(K, D, gid, ox, oy, M.c[0], M.c[1], oz, M.c[2], unit_key(K, gid, ox, oy), F_FREE_RL, 0u, make_float2(0.f, 0.f));
    if (K.mode) D.ukey[gid] = unit_key(K, gid, ox, oy);
    D.unitRes[gid] = 0; D.pendCnt[gid] = 0;
}

// S2g for one directed pair (main.cpp:640-664 / 804-849 / 1759-1826 restated as in kmc_kernels.cu): probe p against neighbour o,
// local indices; the same pair_eval / publish as everywhere else. One copy of the code for every caller (instruction cache).
__device__ __noinline__ void small_classify(const Args &V, int rep, int p, int o) {
    const Dev &D = V.D; const Consts &K = V.K;
    const PairSink none = {nullptr, nullptr, 0, nullptr, 0};
    const ProbeCtx pc = make_probe(K, fetch_rec_small(K, D, small_gid(K, rep, p)));
    int cf = -1;
    const int rr = pair_eval(K, D, pc, fetch_rec_small(K, D, small_gid(K, rep, o)), &cf, none);
    publish(D, pc.u, rr, cf);
}
// a pair within each other's reach this step: both directions are queued for the dense classification pass
KD void small_queue_pair(const Args &V, SmallShared &sm, int rep, int a, int b) {
    const int i = atomicAdd(&sm.nitems, 2);
    if (i + 1 < SMALL_ITEMS) { sm.items[i] = ((unsigned)a << 16) | (unsigned)b; sm.items[i + 1] = ((unsigned)b << 16) | (unsigned)a; }
    else { small_classify(V, rep, a, b); small_classify(V, rep, b, a); }
}
KD bool small_in_reach(const SmallSearch &S, int a, int b) {
    const float4 ca = S.cen[a], cb = S.cen[b];
    const float ex = cb.x - ca.x, ey = cb.y - ca.y, r = ca.z + cb.z;
    return ex * ex + ey * ey <= r * r;
}

// SLOTS replicas per CTA, T threads each (slot = threadIdx.x / T). SLOTS = 1: one replica per CTA, up to four CTAs per SM, each
// in its own stage of the step (T = 128); ensembles of at most one replica per SM take T = 256 -- a molecule per thread, no
// register cap: 7.3 instead of 8.8 us per step for one system, 9.9 instead of 12.3 us for 128 replicas. SLOTS = 4: the four replicas an SM can hold advance in LOCKSTEP -- the barriers between
// the stages are CTA wide --, so all warps of the SM run the same stage and share its instruction lines (the kernel is bound by
// its instruction-cache footprint); used for ensembles that fill the device.
// queue == nullptr: CTA b advances replicas [b SLOTS, (b+1) SLOTS) by nsteps. Otherwise (more replicas than the device holds at
// once) the grid is persistent and the work is dealt in tickets: ticket t = chunk t / G of replica group t % G (chunk = an even
// number of steps); a CTA that takes a ticket waits until the group's previous chunk has been published (queue[1 + group] counts
// its finished chunks; that chunk's ticket was drawn earlier by a CTA that is running, so the wait always ends), loads the
// replicas, advances them and writes them back. Every SM stays busy to the end whatever G modulo the resident CTAs is.
template <int SLOTS, int T>
__global__ void __launch_bounds__(T * SLOTS, (SLOTS == 1 && T == 128) ? SMALL_MINB : 1) k_small_step(const __grid_constant__ Args A, unsigned long long step0, int nsteps, int chunk, int *queue) {
    __shared__ int s_ticket;
    extern __shared__ __align__(16) unsigned char small_dyn[];
    // (the slots' thread numbering is rotated by one warp per slot: stages with a handful of work items run on "warp 0" of every
    // slot, and in lockstep those would otherwise be the CTA's warps 0, 4, 8, 12 -- all on the same warp scheduler)
    const int slot = threadIdx.x / T, tid = (threadIdx.x + 32 * slot) & (T - 1);
    const int NA = A.K.NA, NB = A.K.NB, N = NA + NB, R = A.K.R, G = (R + SLOTS - 1) / SLOTS;
    SmallShared *const SM = reinterpret_cast<SmallShared *>(small_dyn);          // [SLOTS], then the slots' pose / bond regions
    SmallShared &sm = SM[slot];
    unsigned char *const myDyn = small_dyn + (size_t)SLOTS * sizeof(SmallShared) + (size_t)slot * small_dyn_bytes(NA, NB);
  for (;;) {
    int group = blockIdx.x, s0 = 0, s1 = nsteps, myChunk = 0;
    if (queue) {
        __syncthreads();                 // (the last ticket's shared state is no longer in use)
        if (threadIdx.x == 0) s_ticket = atomicAdd(&queue[0], 1);
        __syncthreads();
        const int ticket = s_ticket, nchunks = (nsteps + chunk - 1) / chunk;
        if (ticket >= G * nchunks) break;
        group = ticket % G; myChunk = ticket / G;
        s0 = myChunk * chunk; s1 = min(nsteps, s0 + chunk);
        if (threadIdx.x == 0) while (((volatile int *)queue)[1 + group] < myChunk) __nanosleep(100);
        __syncthreads();
        __threadfence();
    }
    const int rep = group * SLOTS + slot;
    const bool active = rep < R;         // (the last group of an ensemble may be short: its spare slots only keep the barriers company)
    const int NAa = active ? NA : 0, NBa = active ? NB : 0;
    // resident state of the replica: poses (committed + new buffer) and bond table
    double2 *const sRecC = reinterpret_cast<double2 *>(myDyn), *const sRecS2 = sRecC + 2 * NA, *const sRecS3 = sRecS2 + 2 * NA;
    double *const sLig = reinterpret_cast<double *>(sRecS3 + 2 * NA);
    int *const sBond = reinterpret_cast<int *>(sLig + 2 * (size_t)NB * 24);
    // (finish_body reads the bond words of four consecutive receptors as one 16-byte load at GLOBAL indices that are multiples
    // of four: the shared copies start at the same phase, with four spare words on either side)
    const int ph = (rep * NA) & 3, NA8 = (NA + 11) & ~3;
    int *const sRecLig = sBond + 4 + ph, *const sRecCis = sBond + NA8 + 4 + ph, *const sRecSite = sBond + 2 * NA8, *const sLigRec = sRecSite + NA;
    for (int i = tid; i < 2 * NA8; i += T) sBond[i] = -1;
    __syncthreads();
    for (int i = tid; i < NAa; i += T) {
        const int a = rep * NA + i;
        // (__ldcg: past the L1 -- the replica's previous chunk may have been written by another SM)
        sRecC[i] = __ldcg(&A.D.recC[a]); sRecS2[i] = __ldcg(&A.D.recS2[a]); sRecS3[i] = __ldcg(&A.D.recS3[a]);
        sRecLig[i] = __ldcg(&A.D.recLig[a]); sRecCis[i] = __ldcg(&A.D.recCis[a]); sRecSite[i] = __ldcg(&A.D.recSite[a]);
    }
    for (int i = tid; i < NBa * 12; i += T) reinterpret_cast<double2 *>(sLig)[i] = __ldcg(reinterpret_cast<const double2 *>(A.D.lig + (size_t)rep * NB * 24) + i);
    for (int i = tid; i < NBa * 3; i += T) sLigRec[i] = __ldcg(&A.D.ligRec[(size_t)rep * NB * 3 + i]);
    if (tid == 0) {
        // the views: this replica's slices of the work lists; scalars, search records, poses and bonds in shared memory
        const int repv = active ? rep : 0;          // (a spare slot gets a well-formed view of replica 0 that it never uses)
        sm.view[0] = A;
        Dev &V = sm.view[0].D;
        V.scal = sm.scal; V.step64 = &sm.step64; V.touchList = sm.touch; V.reactList = nullptr; V.small = &sm.search; V.pairsFast = sm.fastPairs; V.pairFastCap = TOUCH_CAP / 2;
        const int pendPer = A.D.pendCap / R, pairPer = A.D.pairCap / R, candPer = A.D.candCap / R;
        V.pendList = A.D.pendList + (size_t)repv * pendPer; V.pendCap = pendPer;
        V.pairs = A.D.pairs + (size_t)repv * pairPer; V.pairCap = pairPer;
        V.candRL = A.D.candRL + (size_t)repv * 2 * candPer; V.candCis = A.D.candCis + (size_t)repv * 2 * candPer; V.candCap = candPer;
        V.rejList = A.D.rejList + (size_t)repv * N; V.rejPartner = A.D.rejPartner + (size_t)repv * N;
        const ptrdiff_t a0 = (ptrdiff_t)repv * NA, b0 = (ptrdiff_t)repv * NB;
        V.recC = sRecC - a0; V.recCn = sRecC + NA - a0; V.recS2 = sRecS2 - a0; V.recS2n = sRecS2 + NA - a0; V.recS3 = sRecS3 - a0; V.recS3n = sRecS3 + NA - a0;
        V.lig = sLig - b0 * 24; V.lign = sLig + (ptrdiff_t)NB * 24 - b0 * 24;
        V.recLig = sRecLig - a0; V.recCis = sRecCis - a0; V.recSite = sRecSite - a0; V.ligRec = sLigRec - b0 * 3;
        sm.view[0].K.smallRep = repv;
        sm.view[1] = sm.view[0];
        Dev &W = sm.view[1].D;           // S4 (main.cpp:2164-2202) is a pointer swap: odd steps of the launch see the buffers exchanged
        W.recC = V.recCn; W.recCn = V.recC; W.recS2 = V.recS2n; W.recS2n = V.recS2; W.recS3 = V.recS3n; W.recS3n = V.recS3; W.lig = V.lign; W.lign = V.lig;
        for (int i = 0; i < S_COUNT; i++) sm.scal[i] = 0;
        sm.scal[S_NA_LIVE] = A.K.NAt; sm.scal[S_NB_LIVE] = A.K.NBt;
        sm.scal[S_TOPO_DIRTY] = active;  // the complex tables are re-derived at the start of every launch
        sm.search.recBase = repv * NA; sm.search.ligBase = A.K.NAt + repv * NB - NA; sm.search.N = N; sm.search.dmax = SMALL_DMAX; sm.search.nspec = 0;
        sm.search.share[0] = search_share(A.K, false); sm.search.share[1] = search_share(A.K, true);
        sm.listValid = active ? 0 : 1; sm.nlist = 0; sm.ncx = 0; sm.nitems = 0;
    }
    for (int m = tid; m < N; m += T) sm.search.ref[m] = make_float2(0.f, 0.f);
    __syncthreads();
    SmallSearch &S = sm.search;
    // a condition of one slot that guards a stage with barriers inside becomes a condition of the CTA (the barrier of
    // __syncthreads_or is spared when the CTA holds one replica)
    auto any_slot = [&](bool mine) -> bool { return SLOTS == 1 ? mine : (__syncthreads_or(mine) != 0); };
#ifdef SMALL_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#define SMALL_TICK(i) do { asm volatile("" ::: "memory"); if (threadIdx.x == 0) { const long long t_ = clock64(); tacc[i] += t_ - tprev; tprev = t_; } asm volatile("" ::: "memory"); } while (0)
#else
#define SMALL_TICK(i) do {} while (0)
#endif
    for (int s = s0; s < s1; s++) {
        const Args &V = sm.view[s & 1];          // (chunks are even: every chunk starts on buffer set 0)
        const Dev &D = V.D; const Consts &K = V.K;
        const uint64_t step = step0 + (uint64_t)s + 1;
        // ---- S1: only when the last step's reactions touched the bond table (or at the start of a launch) ----
        {
            const bool dirty = active && (sm.scal[S_NTOUCH] > 0 || sm.scal[S_TOPO_DIRTY]);
            if (any_slot(dirty)) small_rebuild<T>(sm, V, rep, tid, dirty);
        }
        if (tid == 0) {                  // (nothing below reads these before the next barrier; nobody still reads the last step's values: barrier at its end)
            sm.step64 = step;
            sm.scal[S_NPEND] = 0; sm.scal[S_NPAIR] = 0; sm.scal[S_NCAND_RL] = 0; sm.scal[S_NCAND_CIS] = 0; sm.scal[S_NREJ] = 0;
            sm.nitems = 0;
        }
        SMALL_TICK(0);
        // ---- S2 proposals: every molecule is proposed by exactly one thread (a unit head, or the thread of its complex) ----
        // (work index: receptors, ligands from the next multiple of 32 on -- a warp runs ONE of the code paths --, then the complexes)
        if (active) {
            const int NAp = (NA + 31) & ~31, NBp = (NB + 31) & ~31, nw = NAp + NBp + sm.ncx;
            for (int w = tid; w < nw; w += T) {
                if (w < NAp) {
                    if (w >= NA) continue;
                    const int gid = rep * NA + w;
                    propose_one_rec<true>(V, step, 0u, K.NAt, gid, S.meta[w].x & UNIT_MASK, D.recCis[gid], load_rec(D.recC, D.recS2, D.recS3, gid), make_float2(0.f, 0.f), (uint32_t)(w + 1));
                } else if (w < NAp + NBp) {
                    if (w - NAp < NB) small_propose_lig(V, step, K.NAt + rep * NB + (w - NAp), (uint32_t)(NA + w - NAp + 1));
                } else { const int e = sm.cxList[w - NAp - NBp]; complex_move_thread(V, e & 0x3fffffff, !(e & 0x40000000), step, 0u); }
            }
        }
        __syncthreads();
        SMALL_TICK(1);
        // ---- pair list: rebuilt when too many molecules have left their allowance (or a launch starts). A molecule stays covered
        // by the list while all its poses lie within dmax of its centre at build time (ref): listed are all pairs with
        // |ref_a - ref_b| <= share(a) + share(b) + 2 dmax ----
        bool useList = true;
        {
            const bool build = active && (!sm.listValid || S.nspec > SMALL_SPEC_MAX);
            if (SLOTS > 1 && tid == 0) sm.building = build;
            if (any_slot(build)) {
                if (SLOTS == 1) __syncthreads();             // (everyone has read nspec; any_slot was that barrier otherwise)
                if (tid == 0 && build) { S.nspec = 0; sm.nlist = 0; sm.listValid = 1; }
#ifdef SMALL_TIMING
#endif
                __syncthreads();
                // the scans of the replicas that rebuild (usually one of the CTA's) are spread over ALL threads of the CTA: a task
                // is a quarter of one molecule's cyclic scan range (at most 32 neighbours); hits are collected in a bit mask so
                // that the scan itself is a loop without side effects (the loads of several iterations are in flight together)
                for (int k = 0; k < SLOTS; k++) {
                    if (SLOTS > 1 ? !SM[k].building : !build) continue;
                    SmallShared &B = SM[k]; SmallSearch &Sk = B.search;
                    const int half = N / 2, len = (half + 3) / 4;
                    for (int task = threadIdx.x; task < 4 * N; task += T * SLOTS) {
                        const int m = task >> 2, q = task & 3;
                        const float4 me = Sk.cen[m];
                        if (q == 0) {
                            Sk.ref[m] = make_float2(me.x, me.y);
                            const bool special = me.w > Sk.dmax;          // (this step's own move is already too long: periodic wrap, alignment snap)
                            Sk.isSpec[m] = special;
                            if (special) { const int i = atomicAdd(&Sk.nspec, 1); if (i < SMALL_SPEC) Sk.spec[i] = m; }
                        }
                        const float rm = me.z - me.w + 2.f * Sk.dmax + 0.0625f;          // the share alone (+ margins), plus both allowances
                        const int dmax_ = (2 * half == N && m >= half) ? half - 1 : half;      // even N: the pair (m, m + N/2) is listed by its lower member
                        const int dlo = 1 + q * len, cnt = min(dmax_, dlo + len - 1) - dlo + 1;          // neighbours m + dlo ... (cyclic), cnt <= 32
                        if (cnt <= 0) continue;
                        int a0 = m + dlo; if (a0 >= N) a0 -= N;
                        const int n1 = min(cnt, N - a0);
                        unsigned mask = 0;
#pragma unroll 4
                        for (int i = 0; i < n1; i++) {
                            const float4 o = Sk.cen[a0 + i];
                            const float ex = o.x - me.x, ey = o.y - me.y, r = rm + (o.z - o.w);
                            mask |= (ex * ex + ey * ey <= r * r ? 1u : 0u) << i;
                        }
#pragma unroll 4
                        for (int i = n1; i < cnt; i++) {
                            const float4 o = Sk.cen[i - n1];
                            const float ex = o.x - me.x, ey = o.y - me.y, r = rm + (o.z - o.w);
                            mask |= (ex * ex + ey * ey <= r * r ? 1u : 0u) << i;
                        }
                        while (mask) {
                            const int i = __ffs(mask) - 1; mask &= mask - 1;
                            int j = a0 + i; if (j >= N) j -= N;
                            const int slot_ = atomicAdd(&B.nlist, 1);
                            if (slot_ < SMALL_LIST) B.list[slot_] = ((unsigned)m << 16) | (unsigned)j;
                        }
                    }
                }
                __syncthreads();
                if (build) {
                    useList = sm.nlist <= SMALL_LIST && S.nspec <= SMALL_SPEC;      // else too crowded for the list: every pair, in place (and the next step tries again)
                    if (tid == 0) { sm.listValid = useList; if (!useList) S.dmax = fmaxf(8.f, 0.5f * S.dmax); }      // (read again only after the next barrier; a denser replica settles on a shorter allowance)
                }
            }
        }
        SMALL_TICK(2);
        // ---- S2g: list pairs (both members inside their allowance) + every pair of a special molecule: those within reach this
        // step are queued, then classified exactly by a dense pass (one directed pair per thread) ----
        if (!active) {}
        else if (!useList) {                  // (a replica too crowded for the list: every pair)
            for (int m = tid; m < N; m += T)
                for (int j = m + 1; j < N; j++) if (small_in_reach(S, m, j)) small_queue_pair(V, sm, rep, m, j);
        } else {
            const int nl = sm.nlist;
            for (int q = tid; q < nl; q += T) {
                const unsigned w = sm.list[q];
                const int a = (int)(w >> 16), b = (int)(w & 0xffffu);
                if (!(S.isSpec[a] | S.isSpec[b]) && small_in_reach(S, a, b)) small_queue_pair(V, sm, rep, a, b);
            }
            const int nsp = S.nspec;
            for (int k = 0; k < nsp; k++) {
                const int f = S.spec[k];
                for (int t = tid; t < N; t += T)
                    if (t != f && !(S.isSpec[t] && t < f) && small_in_reach(S, f, t)) small_queue_pair(V, sm, rep, f, t);      // two special molecules: once, by the lower one
            }
        }
        __syncthreads();
        SMALL_TICK(7);
        {
            const int ni = min(sm.nitems, SMALL_ITEMS);
            for (int q = tid; q < ni; q += T) small_classify(V, rep, (int)(sm.items[q] >> 16), (int)(sm.items[q] & 0xffffu));
        }
        __syncthreads();
        SMALL_TICK(3);
        // ---- the order dependence of the sweep, from the pending findings (the whole CTA settles one replica after the other) ----
        for (int k = 0; k < SLOTS; k++) {
            const int np = SM[k].scal[S_NPEND];
            if (np > 0) { const Dev &Dk = SM[k].view[s & 1].D; pend_resolve_block(Dk, min(np, Dk.pendCap)); }
        }
        // ---- S3 ----
        if (active) react_pairs_body<true>(K, D, tid, T);
        __syncthreads();
        SMALL_TICK(4);
        {
            bool any = false;
            for (int k = 0; k < SLOTS; k++)
                if (SM[k].scal[S_NCAND_RL] | SM[k].scal[S_NCAND_CIS]) { react_resolve_block(SM[k].view[s & 1].D); any = true; }
            if (any) __syncthreads();
        }
        SMALL_TICK(5);
        if (tid == 0) S.nspec = 0;          // (the special molecules of the NEXT step register during its proposals)
        if (active) finish_body<true>(K, D, tid, T, rep * NA, (rep + 1) * NA);
        __syncthreads();
        SMALL_TICK(6);
    }
#ifdef SMALL_TIMING
    if (threadIdx.x == 0 && rep == 0 && nsteps >= 1000)          // (diagnostic build only) cycles per step: S1 | proposals | list | classify | pending + S3 pairs | S3 resolve | finish
        printf("k_small_step cycles/step: S1 %lld | proposals %lld | list %lld | queue %lld | classify %lld | pending + S3 pairs %lld | S3 resolve %lld | finish %lld; last list %d pairs\n", tacc[0] / nsteps, tacc[1] / nsteps, tacc[2] / nsteps, tacc[7] / nsteps, tacc[3] / nsteps, tacc[4] / nsteps, tacc[5] / nsteps, tacc[6] / nsteps, sm.nlist);
#endif
    // the committed state goes back to the global arrays: into the buffers the host regards as committed after this launch
    // (it swaps its pointers when the step count is odd)
    {
        const int cb = (s1 - s0) & 1;          // (odd only in the last chunk of an odd launch)
        double2 *gC = cb ? A.D.recCn : A.D.recC, *gS2 = cb ? A.D.recS2n : A.D.recS2, *gS3 = cb ? A.D.recS3n : A.D.recS3;
        double *gL = cb ? A.D.lign : A.D.lig;
        for (int i = tid; i < NAa; i += T) {
            const int a = rep * NA + i;
            gC[a] = sRecC[cb * NA + i]; gS2[a] = sRecS2[cb * NA + i]; gS3[a] = sRecS3[cb * NA + i];
            A.D.recLig[a] = sRecLig[i]; A.D.recCis[a] = sRecCis[i]; A.D.recSite[a] = sRecSite[i];
        }
        for (int i = tid; i < NBa * 12; i += T) reinterpret_cast<double2 *>(gL + (size_t)rep * NB * 24)[i] = reinterpret_cast<const double2 *>(sLig + (size_t)cb * NB * 24)[i];
        for (int i = tid; i < NBa * 3; i += T) A.D.ligRec[(size_t)rep * NB * 3 + i] = sLigRec[i];
    }
    if (tid == 0) {
        if (sm.scal[S_OVERFLOW]) atomicOr(&A.D.scal[S_OVERFLOW], sm.scal[S_OVERFLOW]);
        if (rep == 0 && s1 == nsteps) A.D.step64[0] = step0 + (unsigned long long)nsteps;
    }
    if (!queue) break;
    __threadfence();                     // this chunk's state is out before the replica's next chunk may start
    __syncthreads();
    if (threadIdx.x == 0) atomicExch(&queue[1 + group], myChunk + 1);
  }
}

}  // namespace kmc
