// csrc/kmc_kernels.cu -- hand-written sm_100a kernels of the per-timestep KMC sweep
// (replaces /root/reference/main.cpp:461-2202).
//
//  step s:   [complexes]  k_uf_init (+ step begin) -> k_uf_hook -> k_uf_flatten -> k_cx_build   (S1, 514-562; bodies run only
//                                                                                           when the bond table changed)
//            [propose]    k_propose_rec, k_propose_lig, k_propose_complex                  (S2a-S2f, 577-1732)
//            [grid]       k_scan_* -> k_grid_scatter -> k_cells_cut                        (cell list + pair list; sparse path:
//                                                                                           every 6th step, reused in between)
//            [resolve]    k_pairs_eval (+ k_special_pairs) | k_resolve_tiles -> k_pend_resolve   (S2g, 1759-1860 + ordering)
//            [reactions]  k_react_pairs -> k_react_resolve -> k_finish (copy-back of rejected units, dissociation)   (S3, 1876-2141)
//            pointer swap                                                                  (S4, 2164-2202)
//
// Sequential semantics in parallel: the reference sweeps molecules in index order and every overlap test
// sees the already-updated positions of earlier molecules (Gauss-Seidel, main.cpp:577/642). A unit's
// PROPOSAL depends only on its own old pose and its keyed draws, so all proposals are computed at once;
// its accept/reject decision depends on earlier units only through their (rare) rejections, and is
// resolved by a monotone fixed point over the recorded findings: a unit is decided as soon as every earlier unit it touches is.
#include "kmc_device.cuh"
#include <cooperative_groups.h>
#include <cuda_pipeline.h>
#define REC_TILE 256
#ifndef CX_SMALL
#define CX_SMALL 12        // complexes up to this size: one thread each; larger: one warp each
#endif

namespace kmc {

// kernel arguments: device pointers + constants travel by value in the kernel parameter bank
struct Args { Dev D; Consts K; };
#define KARGS const Dev &D = A.D; const Consts &cK = A.K; (void)D; (void)cK;

// ------------------------------------------------------------------------------------------------
// S1: complexes. Union-find over the bond graph; ligand uids are the low numbers so the root of a
// ligand-containing component is its lowest-index ligand = the reference's BFS root (main.cpp:525).
// ------------------------------------------------------------------------------------------------
KD int uf_find(int *parent, int x) {
    int p = parent[x];
    while (p != x) {
        int gp = parent[p];
        if (gp != p) parent[x] = gp;   // path halving (benign race: only ever replaces by an ancestor)
        x = p; p = gp;
    }
    return x;
}
KD void uf_union(int *parent, int a, int b) {
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a > b) { int t = a; a = b; b = t; }          // hook the larger root under the smaller
        int old = atomicCAS(&parent[b], b, a);
        if (old == b) return;
        b = old;
    }
}

// ---- work lists of the complex proposal kernels (cxRoots), by kind so that the threads of a warp run the same code path:
// list 0: small complexes (<= CX_SMALL members, one THREAD each) with ONE ligand (S2e), front of cxRoots;
// list 1: small complexes with several ligands (S2f: shuffles, passes), second half of cxRoots;
// list 2: large complexes (one warp each), from the back of the first half. rootSlot[h] remembers where root h is listed.
KD int *root_list_entry(const Consts &K, const Dev &D, int list, int pos) {
    return list == 0 ? &D.cxRoots[pos] : (list == 1 ? &D.cxRoots[K.NBt + pos] : &D.cxRoots[K.NBt - 1 - pos]);
}
KD int root_list_counter(int list) { return list == 0 ? S_NCX : (list == 1 ? S_NCX_MULTI : S_NCX_BIG); }
KD void root_list_add(const Consts &K, const Dev &D, int h, int size, int nlig) {        // (atomic: the parallel rebuild appends concurrently)
    const int list = size > CX_SMALL ? 2 : (nlig == 1 ? 0 : 1);
    const int pos = atomicAdd(&D.scal[root_list_counter(list)], 1);
    *root_list_entry(K, D, list, pos) = h;
    D.rootSlot[h] = (list << 28) | pos;
}
KD void root_list_remove(const Consts &K, const Dev &D, int h) {                         // single-threaded (incremental update)
    const int slot = D.rootSlot[h];
    if (slot < 0) return;
    const int list = slot >> 28, pos = slot & 0x0fffffff;
    const int last = --D.scal[root_list_counter(list)];
    if (pos != last) { const int h2 = *root_list_entry(K, D, list, last); *root_list_entry(K, D, list, pos) = h2; D.rootSlot[h2] = (list << 28) | pos; }
    D.rootSlot[h] = -1;
}
// neighbours in the bond graph in the reference's order (main.cpp:543-551): receptor -> its ligand, its cis partner; ligand -> sites 2,3,4
KD int bond_neighbours(const Consts &K, const Dev &D, int m, int cand[3]) {
    int nc = 0;
    if (m < K.NAt) {
        if (D.recLig[m] >= 0) cand[nc++] = K.NAt + D.recLig[m];
        if (D.recCis[m] >= 0) cand[nc++] = D.recCis[m];
    } else {
        const int hh = m - K.NAt;
        for (int s = 0; s < 3; s++) if (D.ligRec[hh * 3 + s] >= 0) cand[nc++] = D.ligRec[hh * 3 + s];
    }
    return nc;
}
// breadth-first member row of the complex rooted at ligand h (main.cpp:528-560), `mark` = value that flags a visited molecule
KD void build_member_row(const Consts &K, const Dev &D, int h, int size, int *row, int mark, int &nlig) {
    int head = 0, tail = 0;
    row[tail++] = K.NAt + h; D.bfsMark[K.NAt + h] = mark; D.rowPos[K.NAt + h] = 0;
    nlig = 0;
    while (head < tail) {
        const int m = row[head++];
        nlig += m >= K.NAt;
        int cand[3]; const int nc = bond_neighbours(K, D, m, cand);
        for (int c = 0; c < nc; c++)
            if (D.bfsMark[cand[c]] != mark) { D.bfsMark[cand[c]] = mark; if (tail < size) { D.rowPos[cand[c]] = tail; row[tail++] = cand[c]; } }
    }
}
KD void note_max_complex(const Consts &K, const Dev &D, int h, int size) {
    // main.cpp:896-898 (read first: one hot address). With strips only the rank that owns the root counts it: a halo copy in the
    // outer (possibly stale) part of the halo may show a complex the true trajectory never had
    if (size > D.maxComplex[replica_of_gid(K, K.NAt + h)] && (K.strips <= 1 || d_strip_owner(K, D.lig[(size_t)h * 24]) == K.stripRank)) atomicMax(&D.maxComplex[replica_of_gid(K, K.NAt + h)], size);
}

// First kernel of a step (one warp; begin = 0: only the complex tables, for the strip refresh). Lane 0
//  (1) advances the device-side step counter and resets the per-step scalars;
//  (2) S1, incrementally: the bond table changed only at the few molecules S3 of the last step touched (touchList). Their
//      connected components are re-derived here -- unit heads, sizes, breadth-first member rows (main.cpp:514-562), work lists --
//      and everything else stands. Only when the table changed wholesale (state loaded, strip refresh, more than TOUCH_CAP changes
//      in one step, member storage exhausted) S_TOPO_DIRTY is set and the parallel rebuild kernels below run their bodies.
__global__ void k_step_begin(const __grid_constant__ Args A, int begin) {
    KARGS
    const Consts &K = cK;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (begin) {
        D.step64[0] += 1; D.scal[S_EPOCH] += 1;
        if (D.scal[S_NSPEC] > D.scal[S_NSPEC_MAX]) D.scal[S_NSPEC_MAX] = D.scal[S_NSPEC];
        D.scal[S_NFAR] = 0; D.scal[S_NPEND] = 0; D.scal[S_NPAIR] = 0; D.scal[S_NSPEC] = 0; D.scal[S_NCAND_RL] = 0; D.scal[S_NCAND_CIS] = 0; D.scal[S_NREJ] = 0; D.scal[S_NREACT] = 0; D.scal[S_REC_TICKET] = 0;
        if (cK.phase == 0) D.scal[S_NSURV] = 0;          // (a reuse step keeps the pair list of the last build step)
    }
    const int nt = D.scal[S_NTOUCH];
    D.scal[S_NTOUCH] = 0;
    if (!D.scal[S_TOPO_DIRTY] && nt > 0) {
        const int stamp = 2 + (D.scal[S_EPOCH] & 0x3fffffff) * 2 + (begin ? 0 : 1);          // never 0 / 1 (the parallel rebuild's marks), unique per pass
        int *Q = D.bfsQueue;
        for (int ti = 0; ti < min(nt, TOUCH_CAP) && !D.scal[S_TOPO_DIRTY]; ti++) {
            const int t = D.touchList[ti];
            if (D.cxStamp[t] == stamp) continue;
            // the connected component of t (any order) and its lowest-index ligand
            int qn = 0, minLig = 0x7fffffff, minRec = 0x7fffffff;
            Q[qn++] = t; D.cxStamp[t] = stamp;
            for (int qi = 0; qi < qn; qi++) {
                const int m = Q[qi];
                if (m >= K.NAt) minLig = min(minLig, m); else minRec = min(minRec, m);
                int cand[3]; const int nc = bond_neighbours(K, D, m, cand);
                for (int c = 0; c < nc; c++) if (D.cxStamp[cand[c]] != stamp) { D.cxStamp[cand[c]] = stamp; Q[qn++] = cand[c]; }
            }
            // the ligands of the component lose their old table entries and listings. (The root a member belonged to before is a
            // ligand of this component or of a sibling piece of a split; every piece holds an end of a changed bond, so it is
            // re-derived in this same pass -- and must not be unlisted again once a sibling has listed it afresh.)
            for (int qi = 0; qi < qn; qi++) {
                const int m = Q[qi];
                if (m >= K.NAt) { root_list_remove(K, D, m - K.NAt); D.cxSize[m - K.NAt] = 0; D.cxOff[m - K.NAt] = -1; }
            }
            if (minLig == 0x7fffffff) { for (int qi = 0; qi < qn; qi++) D.unitOf[Q[qi]] = minRec; continue; }      // ligand-free unit: receptor / cis pair
            const int h = minLig - K.NAt, size = qn;
            for (int qi = 0; qi < qn; qi++) D.unitOf[Q[qi]] = minLig;
            D.cxSize[h] = size;
            note_max_complex(K, D, h, size);
            if (size <= 1) continue;
            const int off = D.scal[S_MEMBER_CURSOR];
            if (off + size > K.NT) { D.scal[S_TOPO_DIRTY] = 1; break; }          // member storage exhausted (old rows are not reclaimed): compact by a full rebuild
            D.scal[S_MEMBER_CURSOR] = off + size;
            D.cxOff[h] = off;
            int nlig; build_member_row(K, D, h, size, D.members + off, stamp, nlig);
            for (int i = 0; i < size; i++) D.rowWork[off + i] = D.members[off + i];
            root_list_add(K, D, h, size, nlig);
        }
        if (!D.scal[S_TOPO_DIRTY]) D.events[EV_REBUILDS] += 1;
    }
    if (D.scal[S_TOPO_DIRTY]) { D.scal[S_MEMBER_CURSOR] = 0; D.scal[S_NCX] = 0; D.scal[S_NCX_BIG] = 0; D.scal[S_NCX_MULTI] = 0; D.events[EV_REBUILDS] += 1; }
}
// The parallel rebuild of the whole table: ONE cooperative kernel (grid-wide barriers between its four phases), a single graph
// node that returns at once unless S_TOPO_DIRTY is set (a state was loaded, a strip refresh renumbered the molecules, more than
// TOUCH_CAP bonds changed in one step, the member storage needs compacting).
//   init: every molecule its own set; hook: one thread per receptor, its ligand edge and (once per pair) its cis edge; flatten:
//   unit heads and component sizes; build: one thread per root ligand, breadth-first member order exactly as main.cpp:528-560
__global__ void __launch_bounds__(256) k_cx_rebuild(const __grid_constant__ Args A) {
    KARGS
    if (!D.scal[S_TOPO_DIRTY]) return;          // (uniform over the grid: nobody reaches a barrier)
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    const int nA = nA_live(D), nB = nB_live(D);
    for (int i = tid; i < cK.NT; i += nth) {
        D.ufParent[i] = i; D.bfsMark[i] = 0;          // (dead slots of a strip are harmless singletons)
        if (i < cK.NBt) { D.cxSize[i] = 0; D.cxOff[i] = -1; D.rootSlot[i] = -1; }
    }
    grid.sync();
    for (int a = tid; a < nA; a += nth) {
        const int ua = cK.NBt + a;
        const int l = D.recLig[a];
        if (l >= 0) uf_union(D.ufParent, ua, l);
        const int c = D.recCis[a];
        if (c > a) uf_union(D.ufParent, ua, cK.NBt + c);
    }
    grid.sync();
    for (int i = tid; i < cK.NT; i += nth) {          // uid
        if (!(i < cK.NBt ? i < nB : i - cK.NBt < nA)) continue;
        const int r = uf_find(D.ufParent, i);
        const int gid = i < cK.NBt ? cK.NAt + i : i - cK.NBt;
        const int head = r < cK.NBt ? cK.NAt + r : r - cK.NBt;
        D.unitOf[gid] = head;
        if (r < cK.NBt) atomicAdd(&D.cxSize[r], 1);
    }
    grid.sync();
    for (int h = tid; h < nB; h += nth) {
        if (D.unitOf[cK.NAt + h] != cK.NAt + h) continue;
        const int size = D.cxSize[h];
        note_max_complex(cK, D, h, size);
        if (size <= 1) continue;
        const int off = atomicAdd(&D.scal[S_MEMBER_CURSOR], size);
        D.cxOff[h] = off;
        int nlig; build_member_row(cK, D, h, size, D.members + off, 1, nlig);
        for (int i = 0; i < size; i++) D.rowWork[off + i] = D.members[off + i];       // (single-ligand complexes never permute their row)
        root_list_add(cK, D, h, size, nlig);
    }
}

// Order of the sweep. Every molecule carries the key of its unit: (colour << 30) | head gid. In replay mode the colour is 0,
// so units are swept in index order like the reference (main.cpp:577); in production mode the colour is the 2x2
// checkerboard colour of the head's cell, so the sweep runs colour by colour and by index inside a colour.
#define UNIT_MASK 0x3fffffff
KD bool unit_before(int v, int u) { return (unsigned)v < (unsigned)u; }
KD int unit_key(const Consts &K, int head, double hx, double hy) {
    if (K.mode == 0) return head;
    const int cx = (int)floor((hash_x(K, hx) - K.keyX0) * K.keyInv), cy = (int)floor((hy - K.keyY0) * K.keyInv);
    return (int)((unsigned)head | ((unsigned)((cx & 1) | ((cy & 1) << 1)) << 30));
}

// ------------------------------------------------------------------------------------------------
// S2 proposals
// ------------------------------------------------------------------------------------------------
KD uint64_t seed_of(const Consts &cK, int replica) { return cK.seed + (uint64_t)replica; }

#define F_FAR 1
#define F_GHOST 2
#define F_FREE_RL 4      // receptor: no ligand bound / ligand: at least one free site
#define F_FREE_CIS 8
#define F_DISP 16        // list-reuse step: the molecule has drifted away from its grid entry (or jumped since the grid was built)
// a special entry of a list-reuse step: appended to specList and chained to the cell of the centre it stands for
KD void register_special(const Consts &cK, const Dev &D, int entry, int cell, unsigned stamp) {
    const int idx = atomicAdd(&D.scal[S_NSPEC], 1);
    D.specList[idx] = entry;
    D.specNext[idx] = atomicExch(&D.cellHead[cell], ((unsigned long long)stamp << 32) | (unsigned)(idx + 1));
}
// called once per molecule per step by whoever computed its proposal: the 48-byte neighbour record of the molecule (centre
// old/new, unit key, far/free flags: what the resolve kernels read) and its place in the neighbour search.
//  build step (phase 0): counting-sort histogram of the grid -- entry in the cell of the OLD centre; a far mover gets a second,
//    ghost entry in the cell of its proposal -- and the fp32 centre of the entry (bcen), which later steps measure drift from;
//  reuse step (phase 1): the grid and the pair list of the last build step stand; a far mover, or a molecule whose old centre
//    has drifted more than K.drift from its entry, is SPECIAL: the stale structures do not cover it, k_special_pairs does.
// oz / nz: height of a ligand's centre (0 for a receptor), kept as fp32 in the record: the resolve kernels use it to discard
// pairs that are close in the membrane plane but far apart in z before any bead is fetched. bc = D.bcen[gid], loaded by the caller.
// fp32 centre of the molecule's grid entry: only list-reuse steps (phase 1) measure drift from it
KD float2 entry_centre(const Consts &K, const Dev &D, int gid) { return K.phase == 1 ? D.bcen[gid] : make_float2(0.f, 0.f); }
// half of the largest centre-centre distance at which two molecules can still matter to each other (overlap or S3), split per
// molecule so that the pair radius is a sum: receptor-receptor max(2 rA, cis reach), ligand-ligand reachLL, receptor-ligand
// max(reachRL, reachOn) all fit under w(a) + w(b) (the receptor's share is raised until the mixed pair is covered)
KD float search_share(const Consts &K, bool rec) {
    const float wl = 0.5f * (float)K.reachLL;
    return rec ? fmaxf(0.5f * fmaxf((float)K.ovAA, (float)K.reachCis), fmaxf((float)K.reachRL, (float)K.reachOn) - wl) : wl;
}
KD void mark_far(const Consts &cK, const Dev &D, int gid, double ox, double oy, double nx, double ny, double oz, double nz, int ukey, int freeFlags,
                 unsigned stamp, float2 bc) {
    const double dx = nx - ox, dy = ny - oy;
    if (cK.phase == 2) {
        // fused small-system step: no grid, no far movers (a record with both poses is complete whatever the displacement; the CTA
        // of the replica searches pairs on these records). Search radius of the molecule this step = its share of the reach + its
        // displacement, so that any pose combination of a pair within reach implies |O_a - O_b| <= r(a) + r(b). The pair list of
        // the replica was built around `ref` with dmax of slack per molecule: a molecule whose poses may lie further out is special
        SmallSearch &S = *D.small;
        const bool rec = gid < cK.NAt;
        const int li = rec ? gid - S.recBase : gid - S.ligBase;
        const float disp = sqrtf((float)(dx * dx + dy * dy)) * 1.0001f + 0.0625f;
        const float4 c = make_float4((float)ox, (float)oy, S.share[rec] + disp + 0.0625f, disp);
        S.cen[li] = c;
        const float2 rf = S.ref[li];
        const float ex = c.x - rf.x, ey = c.y - rf.y;
        const bool special = sqrtf(ex * ex + ey * ey) + disp > S.dmax;
        S.isSpec[li] = special;
        if (special) { const int i = atomicAdd(&S.nspec, 1); if (i < SMALL_SPEC) S.spec[i] = li; }
        S.meta[li] = make_int2(ukey, freeFlags);          // (old / new centre and height: read from the resident poses, fetch_rec)
        return;
    }
    const bool far = dx * dx + dy * dy > cK.skin * cK.skin;
    const int rep = replica_of_gid(cK, gid);
    int flags = freeFlags | (far ? F_FAR : 0);
    if (cK.phase == 0) {
        if (D.bcen) D.bcen[gid] = make_float2((float)ox, (float)oy);
        D.molSlot[gid] = atomicAdd(&D.cellCount[cell_of(cK, rep, ox, oy)], 1);
        if (far) {
            const int c2 = cell_of(cK, rep, nx, ny);
            const int slot = atomicAdd(&D.cellCount[c2], 1);
            D.farList[atomicAdd(&D.scal[S_NFAR], 1)] = make_int4(gid, c2, slot, 0);
        }
    } else {
        const double ex = ox - (double)bc.x, ey = oy - (double)bc.y;
        const bool disp = ex * ex + ey * ey > cK.drift * cK.drift;
        if (disp) flags |= F_DISP;
        if (far || disp) {
            register_special(cK, D, gid, cell_of(cK, rep, ox, oy), stamp);
            if (far) register_special(cK, D, gid | GHOST_BIT, cell_of(cK, rep, nx, ny), stamp);
        }
    }
    double2 *nr = reinterpret_cast<double2 *>(D.nrec) + (size_t)gid * 3;
    nr[0] = make_double2(ox, oy); nr[1] = make_double2(nx, ny);
    reinterpret_cast<int4 *>(nr)[2] = make_int4(__float_as_int((float)oz), ukey, flags, __float_as_int((float)nz));
}

// free receptor (main.cpp:584-636), ligand-free cis dimer (682-799): one thread per receptor. (Receptors and ligands are
// separate kernels: the receptor path needs far fewer registers, and the two run side by side on forked branches of the step graph.)
// Persistent CTAs, software pipeline: while a thread transforms receptor gid of tile t out of shared memory, the 64 bytes it needs
// for tile t + gridDim (centre, two sites, unit word, cis word, entry centre) are already in flight (cp.async, 16 B per request,
// each thread copies and later reads only its own slot: no barrier, just the pipeline wait). The load latency that every warp
// used to sit out at its start now overlaps the previous tile's arithmetic.
struct RecStage { double2 c[REC_TILE], s2[REC_TILE], s3[REC_TILE]; float2 bc[REC_TILE]; int head[REC_TILE], cis[REC_TILE]; unsigned ref[REC_TILE]; };
KD void rec_prefetch(const Consts &K, const Dev &D, RecStage &S, int tile, int nLive) {
    const int gid = tile * REC_TILE + threadIdx.x, t = threadIdx.x;
    if (gid < nLive) {
        __pipeline_memcpy_async(&S.c[t], &D.recC[gid], 16); __pipeline_memcpy_async(&S.s2[t], &D.recS2[gid], 16);
        __pipeline_memcpy_async(&S.s3[t], &D.recS3[gid], 16);
        __pipeline_memcpy_async(&S.head[t], &D.unitOf[gid], 4); __pipeline_memcpy_async(&S.cis[t], &D.recCis[gid], 4);
        if (K.phase == 1) __pipeline_memcpy_async(&S.bc[t], &D.bcen[gid], 8);
        if (D.refA) __pipeline_memcpy_async(&S.ref[t], &D.refA[gid], 4);          // strips: the reference id keys the random stream
    }
}
template <bool SMALL = false>          // SMALL: called by the fused small-system kernel (compact Philox, see philox4x32_10)
KD void propose_one_rec(const Args &A, uint64_t step, unsigned stamp, int nLive, int gid, int head, int p, Rec ra, float2 bc, uint32_t me);
KD void propose_rec_body(const Args &A) {
    KARGS
    const Consts &K = cK;
    __shared__ RecStage ST[2];
    const uint64_t step = D.step64[0];
    const unsigned stamp = (unsigned)D.scal[S_EPOCH];
    const int nLive = nA_live(D);
    const int ntiles = (nLive + REC_TILE - 1) / REC_TILE;          // (live receptors only: a strip's capacity padding costs nothing)
    // RECDYN = 1 hands the tiles out dynamically (the first one is the CTA's own index, the following ones come from a ticket
    // counter that k_step_begin resets), so that a CTA that becomes resident late -- the complex kernels of the side branches share
    // the SMs -- takes fewer tiles. Measured on the membrane bench: no gain in the step (0.1537 vs 0.1550 ms), +1.2 us in the kernel
    // (two barriers per tile), so the static stride stays the default.
#ifndef RECDYN
#define RECDYN 0
#endif
    __shared__ int s_next;
    int tile = blockIdx.x;
    if (threadIdx.x == 0) s_next = RECDYN ? gridDim.x + atomicAdd(&D.scal[S_REC_TICKET], 1) : tile + gridDim.x;
    if (tile < ntiles) rec_prefetch(K, D, ST[0], tile, nLive);
    __pipeline_commit();
    __syncthreads();
    int next = s_next;
    for (int it = 0; tile < ntiles; it++) {
        if (RECDYN) __syncthreads();                          // (everyone has read s_next)
        if (RECDYN && threadIdx.x == 0) s_next = gridDim.x + atomicAdd(&D.scal[S_REC_TICKET], 1);      // the tile after the next one: its latency hides behind this tile's arithmetic
        if (next < ntiles) rec_prefetch(K, D, ST[(it + 1) & 1], next, nLive);
        __pipeline_commit();
        __pipeline_wait_prior(1);                 // everything but the newest group has landed: this tile's slot is ready
        const RecStage &S = ST[it & 1];
        const int gid = tile * REC_TILE + threadIdx.x, t = threadIdx.x;
        if (gid < nLive) {
            Rec ra; ra.cx = S.c[t].x; ra.cy = S.c[t].y; ra.s2x = S.s2[t].x; ra.s2y = S.s2[t].y; ra.s3x = S.s3[t].x; ra.s3y = S.s3[t].y;
            propose_one_rec<false>(A, step, stamp, nLive, gid, S.head[t], S.cis[t], ra, K.phase == 1 ? S.bc[t] : make_float2(0.f, 0.f), D.refA ? S.ref[t] : ref_id(K, D, gid));
        }
        if (RECDYN) { __syncthreads(); tile = next; next = s_next; } else { tile = next; next = tile + gridDim.x; }
    }
    __pipeline_wait_prior(0);
}
template <bool SMALL>
KD void propose_one_rec(const Args &A, uint64_t step, unsigned stamp, int nLive, int gid, int head, int p, Rec ra, float2 bc, uint32_t me) {
    KARGS
    const Consts &K = cK;
    if (gid >= nLive || head != gid) return;      // not the head of a unit
    const int rep = replica_of_gid(K, gid);
    const uint64_t seed = seed_of(cK, rep);
    {
        const int a = gid;
        double u0, u1; keyed_uniform2<SMALL>(seed, me, 0, step, 0, u0, u1);
        const double u2 = keyed_uniform<SMALL>(seed, me, 0, step, 2);
        const double phai = mul(mul(u1, 2.0), K.pai);
        double sp, cp; KMC_SINCOS(K, phai, &sp, &cp);
        if (p < 0) {
            // ---- S2a ----
            const double amp = mul(K.ampA, u0);
            const double shx = mul(amp, cp), shy = mul(amp, sp);
            Rec t = {add(ra.cx, shx), add(ra.cy, shy), add(ra.s2x, shx), add(ra.s2y, shy), add(ra.s3x, shx), add(ra.s3y, shy)};
            const double PBx = wrap_offset(t.cx, K.Lx), PBy = wrap_offset(t.cy, K.Ly);
            t.cx = sub(t.cx, PBx); t.s2x = sub(t.s2x, PBx); t.s3x = sub(t.s3x, PBx);
            t.cy = sub(t.cy, PBy); t.s2y = sub(t.s2y, PBy); t.s3y = sub(t.s3y, PBy);
            const double psai = mul(sub(mul(2.0, u2), 1.0), K.rotA);
            double ss, cs; KMC_SINCOS(K, psai, &ss, &cs);
            Rec n; n.cx = t.cx; n.cy = t.cy;
            rotz(cs, ss, t.s2x, t.s2y, t.cx, t.cy, n.s2x, n.s2y);
            rotz(cs, ss, t.s3x, t.s3y, t.cx, t.cy, n.s3x, n.s3y);
            store_rec(D.recCn, D.recS2n, D.recS3n, a, n);
            mark_far(cK, D, a, ra.cx, ra.cy, n.cx, n.cy, 0.0, 0.0, unit_key(K, a, ra.cx, ra.cy), F_FREE_RL | F_FREE_CIS, stamp, bc);
            if (K.mode) D.ukey[a] = unit_key(K, a, ra.cx, ra.cy);
        } else {
            // ---- S2b: this receptor is the lower index of a ligand-free cis pair ----
            Rec rb = load_rec(D.recC, D.recS2, D.recS3, p);
            const double amp = mul(K.ampCis, u0);
            const double shx = mul(amp, cp), shy = mul(amp, sp);
            Rec ta = {add(ra.cx, shx), add(ra.cy, shy), add(ra.s2x, shx), add(ra.s2y, shy), add(ra.s3x, shx), add(ra.s3y, shy)};
            Rec tb = {add(rb.cx, shx), add(rb.cy, shy), add(rb.s2x, shx), add(rb.s2y, shy), add(rb.s3x, shx), add(rb.s3y, shy)};
            const double PBx = mul(K.Lx, round(dvd(dvd(add(ta.cx, tb.cx), 2.0), K.Lx)));
            const double PBy = mul(K.Ly, round(dvd(dvd(add(ta.cy, tb.cy), 2.0), K.Ly)));
            ta.cx = sub(ta.cx, PBx); ta.s2x = sub(ta.s2x, PBx); ta.s3x = sub(ta.s3x, PBx);
            ta.cy = sub(ta.cy, PBy); ta.s2y = sub(ta.s2y, PBy); ta.s3y = sub(ta.s3y, PBy);
            tb.cx = sub(tb.cx, PBx); tb.s2x = sub(tb.s2x, PBx); tb.s3x = sub(tb.s3x, PBx);
            tb.cy = sub(tb.cy, PBy); tb.s2y = sub(tb.s2y, PBy); tb.s3y = sub(tb.s3y, PBy);
            const double psai = mul(sub(mul(2.0, u2), 1.0), K.rotCis);
            double ss, cs; KMC_SINCOS(K, psai, &ss, &cs);
            // rotation centre from the PRE-translation bead centres, summed bead by bead (main.cpp:744-753)
            double cmx = 0, cmy = 0;
            for (int j = 0; j < 4; j++) { cmx = add(add(cmx, ra.cx), rb.cx); cmy = add(add(cmy, ra.cy), rb.cy); }
            cmx = dvd(cmx, 8.0); cmy = dvd(cmy, 8.0);
            Rec na, nb;
            rotz(cs, ss, ta.cx, ta.cy, cmx, cmy, na.cx, na.cy);
            rotz(cs, ss, ta.s2x, ta.s2y, cmx, cmy, na.s2x, na.s2y);
            rotz(cs, ss, ta.s3x, ta.s3y, cmx, cmy, na.s3x, na.s3y);
            rotz(cs, ss, tb.cx, tb.cy, cmx, cmy, nb.cx, nb.cy);
            rotz(cs, ss, tb.s2x, tb.s2y, cmx, cmy, nb.s2x, nb.s2y);
            rotz(cs, ss, tb.s3x, tb.s3y, cmx, cmy, nb.s3x, nb.s3y);
            if (cis_misaligned(K, na, nb)) snap_cis(K, nb, na);          // "relax", main.cpp:770-799
            store_rec(D.recCn, D.recS2n, D.recS3n, a, na);
            store_rec(D.recCn, D.recS2n, D.recS3n, p, nb);
            mark_far(cK, D, a, ra.cx, ra.cy, na.cx, na.cy, 0.0, 0.0, unit_key(K, a, ra.cx, ra.cy), F_FREE_RL, stamp, bc);
            mark_far(cK, D, p, rb.cx, rb.cy, nb.cx, nb.cy, 0.0, 0.0, unit_key(K, a, ra.cx, ra.cy), F_FREE_RL, stamp, K.phase == 1 ? D.bcen[p] : bc);
            if (K.mode) { const int key = unit_key(K, a, ra.cx, ra.cy); D.ukey[a] = key; D.ukey[p] = key; }
        }
        D.unitRes[gid] = 0; D.pendCnt[gid] = 0;
    }
}
#ifndef RECMINB
#define RECMINB 4
#endif
#ifndef LIGMINB
#define LIGMINB 8
#endif
__global__ void __launch_bounds__(256, RECMINB) k_propose_rec(const __grid_constant__ Args A) { propose_rec_body(A); }

// ---- free ligands (S2c, main.cpp:905-969): one thread per ligand, poses staged through shared memory by the bulk-copy engine ----
// A ligand pose is 192 contiguous bytes (AoS). Every thread asks the copy engine for ITS row (cp.async.bulk, global -> shared, one
// 192-byte bulk copy per ligand, completion counted in bytes on one mbarrier per CTA) into a padded tile (row stride 13 x 16 B),
// so no load instruction and no register is spent on staging. The pose then STAYS in shared memory: the thread keeps only the
// rotation matrix, the shift and the centre in registers and carries one point at a time through translation, wrap, reflection
// and rotation (each point's arithmetic is exactly the reference's, the points are independent), which halves the register
// footprint of holding all eight points and doubles the warps an SM can keep in flight. Moved rows go back with bulk stores
// (shared -> global); rows of ligands that do not move here (members of complexes: the complex kernels write those) are not stored.
// S2c for one free ligand (main.cpp:905-969), split so that a caller can carry the eight points through it one at a time:
// the centre decides the wrap and the reflection of the whole body and is the pivot of the rotation; every other point is
// shifted, reflected, wrapped and rotated about the new centre (the points are independent; each one's arithmetic is the reference's)
struct LigMove { double shx, shy, shz, PBx, PBy, twoPBz, c[3]; Rot3 R3; bool reflect; };
KD void lig_move_setup(const Consts &K, LigMove &M, double ox, double oy, double oz, double u0, double u1, double u2, double u3, double u4, double u5) {
    const double amp = mul(K.ampB, u0);
    const double theta = mul(u1, K.pai), phai = mul(mul(u2, 2.0), K.pai);
    double st, ct, sp, cp; KMC_SINCOS(K, theta, &st, &ct); KMC_SINCOS(K, phai, &sp, &cp);
    M.shx = mul(mul(amp, st), cp); M.shy = mul(mul(amp, st), sp); M.shz = mul(amp, ct);
    const double c0x = add(ox, M.shx), c0y = add(oy, M.shy), c0z = add(oz, M.shz);
    M.PBx = wrap_offset(c0x, K.Lx); M.PBy = wrap_offset(c0y, K.Ly);
    M.reflect = c0z > K.Lz || c0z < 0;
    M.twoPBz = M.reflect ? mul(2.0, mul(K.Lz, round(dvd(c0z, K.Lz)))) : 0.0;      // main.cpp:925-931
    M.c[0] = sub(c0x, M.PBx); M.c[1] = sub(c0y, M.PBy); M.c[2] = M.reflect ? add(-c0z, M.twoPBz) : c0z;
    const double rt = mul(sub(mul(2.0, u3), 1.0), K.rotB), rp = mul(sub(mul(2.0, u4), 1.0), K.rotB),
                 rs = mul(sub(mul(2.0, u5), 1.0), K.rotB);
    M.R3 = euler(K, rt, rp, rs);
}
KD void lig_move_point(const LigMove &M, double x, double y, double z, double o[3]) {
    double pt[3] = {add(x, M.shx), add(y, M.shy), add(z, M.shz)};
    if (M.reflect) pt[2] = add(-pt[2], M.twoPBz);
    pt[0] = sub(pt[0], M.PBx); pt[1] = sub(pt[1], M.PBy);
    rot3_about(M.R3, pt, M.c, o);
}
#define LIG_TILE 128
KD uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
KD void mbar_init(uint64_t *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
KD void mbar_arrive_tx(uint64_t *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
KD void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
KD void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
KD void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(LIG_TILE, LIGMINB) k_propose_lig(const __grid_constant__ Args A) {
    KARGS
    const Consts &K = cK;
    __shared__ __align__(16) double2 tile[LIG_TILE][13];
    __shared__ __align__(8) uint64_t bar;
    const uint64_t step = D.step64[0];
    const unsigned stamp = (unsigned)D.scal[S_EPOCH];
    const int h0 = blockIdx.x * LIG_TILE, h = h0 + threadIdx.x, gid = K.NAt + h;
    if (h == 0) D.scal[S_TOPO_DIRTY] = 0;         // the gated rebuild kernel of this step is done
    const int nLive = nB_live(D);
    if (h0 >= nLive) return;                      // (whole CTA: the capacity padding of a strip costs nothing)
    const bool live = h < nLive;
    if (threadIdx.x == 0) { mbar_init(&bar, LIG_TILE); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    double *row = reinterpret_cast<double *>(&tile[threadIdx.x][0]);
    if (live) { mbar_arrive_tx(&bar, 192); bulk_g2s(row, D.lig + (size_t)h * 24, 192, &bar); }
    else mbar_arrive_tx(&bar, 0);
    // the scalar words of this thread's ligand travel in the same latency window as the tile
    int head = -1, csize = 0; float2 bc = make_float2(0.f, 0.f); uint32_t me = 0;
    if (live) { head = D.unitOf[gid]; csize = D.cxSize[h]; if (K.phase == 1) bc = D.bcen[gid]; me = ref_id(K, D, gid); }
    const bool act = live && head == gid && csize <= 1;     // a free ligand (complexes: the complex kernels)
    double u0, u1, u2, u3, u4, u5;
    if (act) {          // the draws do not need the pose: they overlap the copy
        const uint64_t seed = seed_of(cK, replica_of_gid(K, gid));
        keyed_uniform2(seed, me, 0, step, 0, u0, u1); keyed_uniform2(seed, me, 0, step, 2, u2, u3); keyed_uniform2(seed, me, 0, step, 4, u4, u5);
    }
    mbar_wait(&bar, 0);
    if (act) {
        const double ox = row[0], oy = row[1], oz = row[2];
        LigMove M; lig_move_setup(K, M, ox, oy, oz, u0, u1, u2, u3, u4, u5);
        const double *c = M.c;
#pragma unroll 1
        for (int q = 1; q < 8; q++) {
            double o[3]; lig_move_point(M, row[3 * q], row[3 * q + 1], row[3 * q + 2], o);
            row[3 * q] = o[0]; row[3 * q + 1] = o[1]; row[3 * q + 2] = o[2];
        }
        row[0] = c[0]; row[1] = c[1]; row[2] = c[2];          // t*(0)+c = c exactly (main.cpp:958-966)
        mark_far(cK, D, gid, ox, oy, c[0], c[1], oz, c[2], unit_key(K, gid, ox, oy), F_FREE_RL, stamp, bc);
        if (K.mode) D.ukey[gid] = unit_key(K, gid, ox, oy);
        D.unitRes[gid] = 0; D.pendCnt[gid] = 0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the row was written through the generic proxy
        bulk_s2g(D.lign + (size_t)h * 24, row, 192);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");    // the shared row must outlive the copy
    }
}


// ---- complexes: one WARP per ligand-rooted complex with more than one member ------------------------------------------
// The alignment code (S2e/S2f) is sequential by nature (order-dependent snaps, shuffles, the goto state machine) and runs on
// lane 0, but against a copy of the complex held in shared memory: member poses, the bonds inside the complex as member slots,
// the moved[] flags and the working row. All lanes load/translate/rotate/store the members in parallel; sums whose operand
// order matters (wrap centre, rotation centre: main.cpp:1007-1008, 1049-1067) are accumulated by lane 0 in row order.
// A member handle `m` is a slot of the cache (canonical breadth-first order); complexes with more members than the cache holds
// take the same code through accessors that go to global memory (handle = gid).
#define CX_CAP 40          // members cached per complex (one warp per complex: complexes of more than CX_SMALL members)
#define CX_WARPS 4         // of those, per CTA
#define CX_GCAP CX_SMALL   // members cached per small complex with several ligands: GROUPS of CX_G lanes, CX_GROUPS of them per CTA
#ifndef CX_G
#define CX_G 8
#endif
#ifndef CX_GROUPS
#define CX_GROUPS (32 / CX_G)         // (one warp per CTA: an idle launch -- a membrane without such complexes -- does not keep the streaming kernels' CTAs waiting for shared memory)
#endif
// a group of G consecutive lanes of a warp (G = 32: the warp) works on one complex
template <int G> KD unsigned grp_mask() { return G == 32 ? 0xffffffffu : ((G == 32 ? 0u : ((1u << (G & 31)) - 1u)) << ((threadIdx.x & 31) & ~(G - 1))); }
template <int G> KD void grp_sync() { __syncwarp(grp_mask<G>()); }
template <int G, class T> KD T grp_bcast(T v) { return __shfl_sync(grp_mask<G>(), v, 0, G); }      // the value of the group's lane 0

template <int CAP> struct CxSharedT {          // per group
    double pose[CAP][24];             // receptor: cx,cy,s2x,s2y,s3x,s3y ; ligand: 8 points x 3
    int gid[CAP];
    short lig[CAP], cis[CAP], site[CAP];   // receptor: slot of its ligand / cis partner (-1 none), ligand site 0..2
    short rec3[CAP][3];               // ligand: slot of the receptor on site s (-1 none)
    unsigned char moved[CAP];
    int row[CAP];                     // working order (handles), permuted by the shuffles
    int draw[CAP];                    // the rand() values of one shuffle, drawn by all lanes at once
};

template <int CAP> struct CxLocalT {           // accessors on the shared-memory copy; handle = slot
    CxSharedT<CAP> &S; const Consts &K; int NAt;
    KD bool is_rec(int m) const { return S.gid[m] < NAt; }
    KD int recLig(int m) const { return S.lig[m]; }
    KD int recSite(int m) const { return S.site[m]; }
    KD int recCis(int m) const { return S.cis[m]; }
    KD int ligRec(int m, int s) const { return S.rec3[m][s]; }
    KD bool moved(int m) const { return S.moved[m] != 0; }
    KD void set_moved(int m) const { S.moved[m] = 1; }
    KD Rec rec(int m) const { const double *p = S.pose[m]; Rec r = {p[0], p[1], p[2], p[3], p[4], p[5]}; return r; }
    KD void put(int m, const Rec &r) const { double *p = S.pose[m]; p[0] = r.cx; p[1] = r.cy; p[2] = r.s2x; p[3] = r.s2y; p[4] = r.s3x; p[5] = r.s3y; }
    KD void lig(int m, Lig &l) const { const double *p = S.pose[m]; for (int q = 0; q < 24; q++) (&l.p[0][0])[q] = p[q]; }
    KD void lig_site(int m, int s, double &sx, double &sy, double &bx, double &by) const { const double *p = S.pose[m]; sx = p[(5 + s) * 3]; sy = p[(5 + s) * 3 + 1]; bx = p[(1 + s) * 3]; by = p[(1 + s) * 3 + 1]; }
    KD void put(int m, const Lig &l) const { double *p = S.pose[m]; for (int q = 0; q < 24; q++) p[q] = (&l.p[0][0])[q]; }
};
struct CxGlobal {          // accessors straight on the nxt arrays; handle = gid
    const Dev &D; const Consts &K; int NAt;
    KD bool is_rec(int m) const { return m < NAt; }
    KD int recLig(int m) const { int h = D.recLig[m]; return h < 0 ? -1 : NAt + h; }
    KD int recSite(int m) const { return D.recSite[m]; }
    KD int recCis(int m) const { return D.recCis[m]; }
    KD int ligRec(int m, int s) const { return D.ligRec[(m - NAt) * 3 + s]; }
    KD bool moved(int m) const { return D.movedFlag[m] != 0; }
    KD void set_moved(int m) const { D.movedFlag[m] = 1; }
    KD Rec rec(int m) const { return load_rec(D.recCn, D.recS2n, D.recS3n, m); }
    KD void put(int m, const Rec &r) const { store_rec(D.recCn, D.recS2n, D.recS3n, m, r); }
    KD void lig(int m, Lig &l) const { load_lig(D.lign, m - NAt, l); }
    KD void lig_site(int m, int s, double &sx, double &sy, double &bx, double &by) const { const double *p = D.lign + (size_t)(m - NAt) * 24; sx = p[(5 + s) * 3]; sy = p[(5 + s) * 3 + 1]; bx = p[(1 + s) * 3]; by = p[(1 + s) * 3 + 1]; }
    KD void put(int m, const Lig &l) const { store_lig(D.lign, m - NAt, l); }
};

// main.cpp:1296-1328 (also 1511-1545, 1651-1683): re-snap receptor a onto its ligand if misaligned
template <class Cx> KD bool resnap_rec_to_its_ligand(const Cx &C, int a) {
    const int h = C.recLig(a); if (h < 0) return false;
    const int s = C.recSite(a);
    double sx, sy, bx, by; C.lig_site(h, s, sx, sy, bx, by);          // (the test and the snap only read the site point and its bead)
    Rec r = C.rec(a);
    if (!rl_misaligned(C.K, sx, sy, bx, by, r)) return false;
    snap_rec_to_lig(C.K, r, sx, sy, bx, by); C.put(a, r);
    return true;
}
// main.cpp:1548-1578, 1699-1728: re-snap the cis partner of a from a's axis if misaligned
template <class Cx> KD bool resnap_cis_partner(const Cx &C, int a, int a2) {
    Rec r1 = C.rec(a), r2 = C.rec(a2);
    if (!cis_misaligned(C.K, r1, r2)) return false;
    snap_cis(C.K, r2, r1); C.put(a2, r2);
    return true;
}
// body at `lable4`, main.cpp:1439-1585
template <class Cx> KD void reseat_ligand(const Cx &C, int h, int s, int a1) {
    const Consts &K = C.K;
    C.set_moved(h);
    Lig b; C.lig(h, b);
    const Rec r = C.rec(a1);
    const double zA = rec_bead_z(K, 3);
    for (int q = 0; q < 8; q++) b.p[q][2] = zA;
    b.p[4][2] = add(zA, K.rB);
    const double ax1 = K.ghost[1 + s][0], ay1 = K.ghost[1 + s][1];
    const double ax2 = sub(r.cx, r.s2x), ay2 = sub(r.cy, r.s2y);
    const double dot = add(mul(ax1, ax2), mul(ay1, ay2)), det = sub(mul(ax1, ay2), mul(ay1, ax2));
    const double angle = add(atan2(-det, -dot), K.pai);
    const double cx = add(mul(K.fSeat, sub(r.s2x, r.cx)), r.s2x), cy = add(mul(K.fSeat, sub(r.s2y, r.cy)), r.s2y);
    seat_ligand(K, b, angle, cx, cy);
    C.put(h, b);
    for (int m = 0; m < 3; m++) {
        const int am = C.ligRec(h, m);
        if (am < 0) continue;
        if (resnap_rec_to_its_ligand(C, am)) C.set_moved(am);
        const int a2 = C.recCis(am);
        if (a2 >= 0 && resnap_cis_partner(C, am, a2)) C.set_moved(a2);
    }
}
template <class Cx> KD bool bridge_candidate(const Cx &C, int h, int s) {        // main.cpp:1420-1423
    const int a1 = C.ligRec(h, s); if (a1 < 0) return false;
    const int a2 = C.recCis(a1); if (a2 < 0) return false;
    return C.recLig(a2) >= 0 && !C.moved(h);
}
template <class Cx> KD bool lig_site_misaligned(const Cx &C, int h, int s, int a1) {
    double sx, sy, bx, by; C.lig_site(h, s, sx, sy, bx, by);
    return rl_misaligned(C.K, sx, sy, bx, by, C.rec(a1));
}
// std::random_shuffle(&row[1], &row[size]) with rand() (libstdc++): the last member never moves (main.cpp:1285)
KD void shuffle_row(int *row, int size, uint64_t seed, uint32_t root, uint32_t &cnt, uint64_t step) {
    int n = size - 1;
    for (int i = 1; i < n; i++) {
        int j = keyed_rand31(seed, root, cnt++, step) % (i + 1);
        if (i != j) { int t = row[i]; row[i] = row[j]; row[j] = t; }
    }
}
// S2e (one ligand, handle hl) / S2f (several ligands): the sequential alignment of a complex after its rigid move
template <class Cx> KD void align_complex(const Cx &C, int *row, int size, int nB, int hl, uint64_t seed, uint32_t me, uint64_t step) {
    const Consts &K = C.K;
    if (nB == 1) {
        // ---- S2e (main.cpp:1138-1274) ----
        Lig b; C.lig(hl, b);
        if (b.p[4][2] != add(b.p[0][2], K.rB)) {                  // exact compare: first time only
            const double zA = rec_bead_z(K, 3);                   // R_z_new[last receptor][3][1]: the template value
            for (int q = 0; q < 8; q++) b.p[q][2] = zA;
            b.p[4][2] = add(zA, K.rB);
            const double angle = add(atan2(sub(b.p[1][0], b.p[0][0]), sub(b.p[1][1], b.p[0][1])), K.pai);
            seat_ligand(K, b, angle, b.p[0][0], b.p[0][1]);
            C.put(hl, b);
        }
        // (the ligand is not touched any more: one copy serves all its receptors; the receptors stay in registers for the cis pass)
        Rec rr[3]; int aa[3];
        for (int s = 0; s < 3; s++) {
            const int a1 = aa[s] = C.ligRec(hl, s);
            if (a1 < 0) continue;
            rr[s] = C.rec(a1);
            if (rl_misaligned(K, b, s, rr[s])) { snap_rec_to_lig(K, rr[s], b, s); C.put(a1, rr[s]); }      // main.cpp:1196-1233
        }
        for (int s = 0; s < 3; s++) {
            if (aa[s] < 0) continue;
            const int a2 = C.recCis(aa[s]);
            if (a2 < 0) continue;
            Rec r2 = C.rec(a2);
            if (cis_misaligned(K, rr[s], r2)) {                                                              // main.cpp:1237-1274
                snap_cis(K, r2, rr[s]); C.put(a2, r2);
                for (int t = 0; t < 3; t++) if (aa[t] == a2) rr[t] = r2;          // (a partner that sits on the same ligand: keep the register copy current)
            }
        }
        return;
    }
    // ---- S2f (main.cpp:1284-1732) ----
    uint32_t cnt = 0;
    shuffle_row(row, size, seed, me, cnt, step);                           // pass 0
    for (int i = 0; i < size; i++) { const int a = row[i]; if (C.is_rec(a) && resnap_rec_to_its_ligand(C, a)) C.set_moved(a); }
    shuffle_row(row, size, seed, me, cnt, step);                           // pass 1
    for (int i = 0; i < size; i++) {
        const int a = row[i];
        if (C.is_rec(a) && C.recLig(a) >= 0 && C.recCis(a) >= 0 && C.recLig(C.recCis(a)) >= 0 && !C.moved(a)) {
            const int a2 = C.recCis(a);
            C.set_moved(a); C.set_moved(a2);
            Rec r1 = C.rec(a), r2 = C.rec(a2);
            if (cis_misaligned(K, r1, r2)) { snap_cis(K, r1, r2); C.put(a, r1); }   // a rebuilt FROM a2, 1390-1400
        }
    }
    // passes 2 and 3; pass 3 re-enters pass 2's innermost block (goto lable4, main.cpp:1628 -> 1438)
    bool resume = false; int i = 0, s = 0, h = 0, a1 = 0;
    for (;;) {
        if (!resume) { shuffle_row(row, size, seed, me, cnt, step); i = 0; }
        for (; i < size; i++) {
            if (C.is_rec(row[i])) continue;
            if (!resume) { h = row[i]; s = 0; }
            for (; s < 3; s++) {
                bool run;
                if (resume) { run = true; resume = false; }
                else {
                    run = false;
                    if (bridge_candidate(C, h, s)) { a1 = C.ligRec(h, s); run = lig_site_misaligned(C, h, s, a1); }
                }
                if (run) reseat_ligand(C, h, s, a1);
            }
        }
        shuffle_row(row, size, seed, me, cnt, step);                       // pass 3
        for (i = 0; i < size; i++) {
            if (C.is_rec(row[i])) continue;
            h = row[i];
            for (s = 0; s < 3; s++)
                if (bridge_candidate(C, h, s)) {
                    a1 = C.ligRec(h, s);
                    if (lig_site_misaligned(C, h, s, a1)) { resume = true; break; }
                }
            if (resume) break;
        }
        if (!resume) break;
    }
    for (int q = 0; q < size; q++) { const int a = row[q]; if (C.is_rec(a) && resnap_rec_to_its_ligand(C, a)) C.set_moved(a); }   // pass 4
    for (int q = 0; q < size; q++) {                                                                                            // pass 5
        const int a = row[q];
        if (C.is_rec(a) && C.recLig(a) >= 0 && C.recCis(a) >= 0 && C.recLig(C.recCis(a)) < 0) resnap_cis_partner(C, a, C.recCis(a));
    }
}

// The same alignment for a cached complex, executed by the whole warp: the passes whose iterations cannot influence each other
// run one member per lane -- pass 0 and pass 4 (every receptor is re-snapped onto ITS ligand, ligands are not touched),
// pass 5 (every ligand-free cis partner is rebuilt from ITS one partner) -- and the Philox draws of a shuffle are made by all
// lanes; the order-dependent parts (the swaps of a shuffle, passes 1-3 with their moved[] flags and the goto) stay on lane 0.
template <int G, int CAP> KD void shuffle_row_warp(CxSharedT<CAP> &S, int size, uint64_t seed, uint32_t root, uint32_t &cnt, uint64_t step, int lane) {
    const int n = size - 1;
    for (int i = 1 + lane; i < n; i += G) S.draw[i] = keyed_rand31(seed, root, cnt + (uint32_t)(i - 1), step) % (i + 1);
    if (n > 1) cnt += (uint32_t)(n - 1);
    grp_sync<G>();
    if (lane == 0)
        for (int i = 1; i < n; i++) { const int j = S.draw[i]; if (i != j) { const int t = S.row[i]; S.row[i] = S.row[j]; S.row[j] = t; } }
    grp_sync<G>();
}
template <int G, int CAP> KD void align_complex_warp(const CxLocalT<CAP> &C, CxSharedT<CAP> &S, int size, int nB, uint64_t seed, uint32_t me, uint64_t step, int lane) {
    const Consts &K = C.K;
    if (nB == 1) { if (lane == 0) align_complex(C, S.row, size, nB, 0, seed, me, step); grp_sync<G>(); return; }
    uint32_t cnt = 0;
    shuffle_row_warp<G, CAP>(S, size, seed, me, cnt, step, lane);                   // pass 0 (order free: see above)
    for (int i = lane; i < size; i += G) { const int a = S.row[i]; if (C.is_rec(a) && resnap_rec_to_its_ligand(C, a)) C.set_moved(a); }
    grp_sync<G>();
    shuffle_row_warp<G, CAP>(S, size, seed, me, cnt, step, lane);                   // pass 1
    if (lane == 0)
        for (int i = 0; i < size; i++) {
            const int a = S.row[i];
            if (C.is_rec(a) && C.recLig(a) >= 0 && C.recCis(a) >= 0 && C.recLig(C.recCis(a)) >= 0 && !C.moved(a)) {
                const int a2 = C.recCis(a);
                C.set_moved(a); C.set_moved(a2);
                Rec r1 = C.rec(a), r2 = C.rec(a2);
                if (cis_misaligned(K, r1, r2)) { snap_cis(K, r1, r2); C.put(a, r1); }
            }
        }
    grp_sync<G>();
    // passes 2 and 3 (goto lable4, main.cpp:1628 -> 1438): lane 0 runs the state machine, every lane takes part in the shuffles
    int resume = 0; int i = 0, s = 0, h = 0, a1 = 0;
    for (;;) {
        if (!resume) { shuffle_row_warp<G, CAP>(S, size, seed, me, cnt, step, lane); i = 0; }
        if (lane == 0) {
            bool rs = resume != 0;
            for (; i < size; i++) {
                if (C.is_rec(S.row[i])) continue;
                if (!rs) { h = S.row[i]; s = 0; }
                for (; s < 3; s++) {
                    bool run;
                    if (rs) { run = true; rs = false; }
                    else {
                        run = false;
                        if (bridge_candidate(C, h, s)) { a1 = C.ligRec(h, s); run = lig_site_misaligned(C, h, s, a1); }
                    }
                    if (run) reseat_ligand(C, h, s, a1);
                }
            }
        }
        grp_sync<G>();
        shuffle_row_warp<G, CAP>(S, size, seed, me, cnt, step, lane);               // pass 3
        resume = 0;
        if (lane == 0) {
            for (i = 0; i < size; i++) {
                if (C.is_rec(S.row[i])) continue;
                h = S.row[i];
                for (s = 0; s < 3; s++)
                    if (bridge_candidate(C, h, s)) {
                        a1 = C.ligRec(h, s);
                        if (lig_site_misaligned(C, h, s, a1)) { resume = 1; break; }
                    }
                if (resume) break;
            }
        }
        resume = grp_bcast<G>(resume);
        if (!resume) break;
    }
    grp_sync<G>();
    for (int q = lane; q < size; q += G) { const int a = S.row[q]; if (C.is_rec(a) && resnap_rec_to_its_ligand(C, a)) C.set_moved(a); }   // pass 4
    grp_sync<G>();
    for (int q = lane; q < size; q += G) {                                                                                              // pass 5
        const int a = S.row[q];
        if (C.is_rec(a) && C.recLig(a) >= 0 && C.recCis(a) >= 0 && C.recLig(C.recCis(a)) < 0) resnap_cis_partner(C, a, C.recCis(a));
    }
    grp_sync<G>();
}

// rigid move of one member given the unit's shift / wrap / rotation (main.cpp:993-1026, 1035-1070, 1103-1128 for one molecule)
KD void shift_pose(double *p, bool isRec, double dx, double dy, bool subtract) {
    const int n = isRec ? 3 : 8, st = isRec ? 2 : 3;
    for (int q = 0; q < n; q++) {
        p[q * st] = subtract ? sub(p[q * st], dx) : add(p[q * st], dx);
        p[q * st + 1] = subtract ? sub(p[q * st + 1], dy) : add(p[q * st + 1], dy);
    }
}
KD void rotate_pose(double *p, bool isRec, double cs, double ss, double cmx, double cmy, double cmz) {
    const int n = isRec ? 3 : 8, st = isRec ? 2 : 3;
    for (int q = 0; q < n; q++) {
        double nx, ny; rotz(cs, ss, p[q * st], p[q * st + 1], cmx, cmy, nx, ny);
        p[q * st] = nx; p[q * st + 1] = ny;
        if (!isRec) p[q * st + 2] = add(sub(p[q * st + 2], cmz), cmz);        // 1*(z-c)+c, main.cpp:1123
    }
}

// One complex moved by ONE thread on global memory (handles = gids): S2d rigid move + S2e/S2f alignment. Used by
// k_propose_complex_small (a thread per small complex) and, for complexes beyond the shared-memory cache, by lane 0 of a warp.
__device__ __noinline__ void complex_move_serial(const Args &A, int h0, int size, int nB, uint64_t step, unsigned stamp, double u0, double u1, double u2) {
    KARGS
    const Consts &K = cK;
    const int rootGid = K.NAt + h0, nA = size - nB;
    const int *rowIn = D.members + D.cxOff[h0];
    int *rowOut = D.rowWork + D.cxOff[h0];
    const uint64_t seed = seed_of(cK, replica_of_gid(K, K.NAt + h0));
    const uint32_t me = ref_id(K, D, rootGid);
    const double amp = mul(nB == 1 ? K.ampBond : 0.0, u0);
    const double phai = mul(mul(u1, 2.0), K.pai);
    double sp, cp; KMC_SINCOS(K, phai, &sp, &cp);
    const double shx = mul(amp, cp), shy = mul(amp, sp);
    const double psai = mul(sub(mul(2.0, u2), 1.0), nB == 1 ? K.rotBond : 0.0);
    double ss, cs; KMC_SINCOS(K, psai, &ss, &cs);
    CxGlobal C{D, K, K.NAt};
    // wrap centre (main.cpp:1007-1008, 1022-1023) and rotation centre (1048-1068) are sums over the members in row order: two
    // read-only passes over centres / beads; then ONE pass that shifts, wraps and rotates every member and writes it once
    double PBx = 0, PBy = 0;
    for (int i = 0; i < size; i++) {
        const int m = rowIn[i]; rowOut[i] = m; D.movedFlag[m] = 0;
        double x, y;
        if (m < K.NAt) { const double2 c = D.recC[m]; x = c.x; y = c.y; } else { const double *p = D.lig + (size_t)(m - K.NAt) * 24; x = p[0]; y = p[1]; }
        PBx = add(PBx, add(x, shx)); PBy = add(PBy, add(y, shy));
    }
    PBx = mul(K.Lx, round(dvd(dvd(PBx, (double)(nA + nB)), K.Lx)));
    PBy = mul(K.Ly, round(dvd(dvd(PBy, (double)(nA + nB)), K.Ly)));
    double cmx = 0, cmy = 0, cmz = 0;
    for (int i = 0; i < size; i++) {
        const int m = rowOut[i];
        if (m < K.NAt) {
            const double2 c = D.recC[m];
            const double x = sub(add(c.x, shx), PBx), y = sub(add(c.y, shy), PBy);
            for (int j = 1; j <= 4; j++) { cmx = add(cmx, x); cmy = add(cmy, y); cmz = add(cmz, rec_bead_z(K, j)); }
        } else {
            const double *p = D.lig + (size_t)(m - K.NAt) * 24;
            for (int q = 0; q < 4; q++) { cmx = add(cmx, sub(add(p[q * 3], shx), PBx)); cmy = add(cmy, sub(add(p[q * 3 + 1], shy), PBy)); cmz = add(cmz, p[q * 3 + 2]); }
        }
    }
    const double nbeads = (double)(4 * nA + 4 * nB);
    cmx = dvd(cmx, nbeads); cmy = dvd(cmy, nbeads); cmz = dvd(cmz, nbeads);
    for (int i = 0; i < size; i++) {
        const int m = rowOut[i];
        if (m < K.NAt) {
            Rec r = load_rec(D.recC, D.recS2, D.recS3, m);
            shift_pose(&r.cx, true, shx, shy, false); shift_pose(&r.cx, true, PBx, PBy, true); rotate_pose(&r.cx, true, cs, ss, cmx, cmy, cmz);
            C.put(m, r);
        } else {
            Lig l; load_lig(D.lig, m - K.NAt, l);
            shift_pose(&l.p[0][0], false, shx, shy, false); shift_pose(&l.p[0][0], false, PBx, PBy, true); rotate_pose(&l.p[0][0], false, cs, ss, cmx, cmy, cmz);
            C.put(m, l);
        }
    }
    align_complex(C, rowOut, size, nB, rootGid, seed, me, step);
    const int ckey = unit_key(K, rootGid, D.lig[(size_t)h0 * 24], D.lig[(size_t)h0 * 24 + 1]);
    for (int q = 0; q < size; q++) {
        const int m = rowOut[q];
        if (K.mode) D.ukey[m] = ckey;
        if (m < K.NAt) { double2 o = D.recC[m], n = D.recCn[m]; mark_far(cK, D, m, o.x, o.y, n.x, n.y, 0.0, 0.0, ckey, (D.recLig[m] < 0 ? F_FREE_RL : 0) | (D.recCis[m] < 0 ? F_FREE_CIS : 0), stamp, entry_centre(K, D, m)); }
        else { const double *o = D.lig + (size_t)(m - K.NAt) * 24, *n = D.lign + (size_t)(m - K.NAt) * 24; const int *occ = D.ligRec + (size_t)(m - K.NAt) * 3; mark_far(cK, D, m, o[0], o[1], n[0], n[1], o[2], n[2], ckey, (occ[0] < 0 || occ[1] < 0 || occ[2] < 0) ? F_FREE_RL : 0, stamp, entry_centre(K, D, m)); }
    }
}
// Single-ligand complex (S2d + S2e), the commonest kind: ligand h0, the receptors on its sites and their ligand-free cis partners
// (a cis partner with a ligand of its own would make it a multi-ligand complex). The topology fixes the breadth-first member row --
// [ligand, receptors in site order, their partners in the same order] (main.cpp:528-560) -- so nothing has to be walked: the ligand
// and the bond words are requested at once, then the receptors, then the partners: three dependent rounds of loads instead of one
// per member and pass. Arithmetic and its order are those of complex_move_serial / align_complex. Returns false (nothing written)
// if the complex is not of this shape.
KD bool single_ligand_move(const Args &A, int h0, int size, uint64_t step, unsigned stamp, double u0, double u1, double u2) {
    KARGS
    const Consts &K = cK;
    const int rootGid = K.NAt + h0;
    Lig L; load_lig(D.lig, h0, L);
    int a[3], p[3];
    for (int s = 0; s < 3; s++) a[s] = D.ligRec[h0 * 3 + s];
    double2 ca[3], cp_[3];
    for (int s = 0; s < 3; s++) { p[s] = -1; if (a[s] >= 0) { ca[s] = D.recC[a[s]]; p[s] = D.recCis[a[s]]; } }
    int n = 1;
    for (int s = 0; s < 3; s++) {
        if (a[s] >= 0) n++;
        if (p[s] >= 0) { n++; cp_[s] = D.recC[p[s]]; if (p[s] == a[0] || p[s] == a[1] || p[s] == a[2]) return false; }
    }
    if (n != size) return false;
    const double ox = L.p[0][0], oy = L.p[0][1], oz = L.p[0][2];
    const double amp = mul(K.ampBond, u0);
    const double phai = mul(mul(u1, 2.0), K.pai);
    double sp, cp; KMC_SINCOS(K, phai, &sp, &cp);
    const double shx = mul(amp, cp), shy = mul(amp, sp);
    const double psai = mul(sub(mul(2.0, u2), 1.0), K.rotBond);
    double ss, cs; KMC_SINCOS(K, psai, &ss, &cs);
    // wrap centre and rotation centre: sums in row order (ligand, receptors, partners)
    double PBx = add(0.0, add(ox, shx)), PBy = add(0.0, add(oy, shy));
    for (int s = 0; s < 3; s++) if (a[s] >= 0) { PBx = add(PBx, add(ca[s].x, shx)); PBy = add(PBy, add(ca[s].y, shy)); }
    for (int s = 0; s < 3; s++) if (p[s] >= 0) { PBx = add(PBx, add(cp_[s].x, shx)); PBy = add(PBy, add(cp_[s].y, shy)); }
    PBx = mul(K.Lx, round(dvd(dvd(PBx, (double)size), K.Lx)));
    PBy = mul(K.Ly, round(dvd(dvd(PBy, (double)size), K.Ly)));
    double cmx = 0, cmy = 0, cmz = 0;
    for (int q = 0; q < 4; q++) { cmx = add(cmx, sub(add(L.p[q][0], shx), PBx)); cmy = add(cmy, sub(add(L.p[q][1], shy), PBy)); cmz = add(cmz, L.p[q][2]); }
    for (int pass = 0; pass < 2; pass++)
        for (int s = 0; s < 3; s++) {
            if ((pass ? p[s] : a[s]) < 0) continue;
            const double2 c = pass ? cp_[s] : ca[s];
            const double x = sub(add(c.x, shx), PBx), y = sub(add(c.y, shy), PBy);
            for (int j = 1; j <= 4; j++) { cmx = add(cmx, x); cmy = add(cmy, y); cmz = add(cmz, rec_bead_z(K, j)); }
        }
    const double nbeads = (double)(4 * size);
    cmx = dvd(cmx, nbeads); cmy = dvd(cmy, nbeads); cmz = dvd(cmz, nbeads);
    // the ligand: rigid move, then lay-down (main.cpp:1140-1189)
    shift_pose(&L.p[0][0], false, shx, shy, false); shift_pose(&L.p[0][0], false, PBx, PBy, true); rotate_pose(&L.p[0][0], false, cs, ss, cmx, cmy, cmz);
    if (L.p[4][2] != add(L.p[0][2], K.rB)) {
        const double zA = rec_bead_z(K, 3);
        for (int q = 0; q < 8; q++) L.p[q][2] = zA;
        L.p[4][2] = add(zA, K.rB);
        const double angle = add(atan2(sub(L.p[1][0], L.p[0][0]), sub(L.p[1][1], L.p[0][1])), K.pai);
        seat_ligand(K, L, angle, L.p[0][0], L.p[0][1]);
    }
    store_lig(D.lign, h0, L);
    const int ckey = unit_key(K, rootGid, ox, oy);
    const bool freeSite = a[0] < 0 || a[1] < 0 || a[2] < 0;
    mark_far(cK, D, rootGid, ox, oy, L.p[0][0], L.p[0][1], oz, L.p[0][2], ckey, freeSite ? F_FREE_RL : 0, stamp, entry_centre(K, D, rootGid));
    if (K.mode) D.ukey[rootGid] = ckey;
    // every receptor (rigid move, snap onto its site: 1196-1233), then its partner (rigid move, snap onto the receptor: 1237-1274)
    for (int s = 0; s < 3; s++) {
        if (a[s] < 0) continue;
        Rec r = load_rec(D.recC, D.recS2, D.recS3, a[s]);
        shift_pose(&r.cx, true, shx, shy, false); shift_pose(&r.cx, true, PBx, PBy, true); rotate_pose(&r.cx, true, cs, ss, cmx, cmy, cmz);
        if (rl_misaligned(K, L, s, r)) snap_rec_to_lig(K, r, L, s);
        store_rec(D.recCn, D.recS2n, D.recS3n, a[s], r);
        mark_far(cK, D, a[s], ca[s].x, ca[s].y, r.cx, r.cy, 0.0, 0.0, ckey, p[s] < 0 ? F_FREE_CIS : 0, stamp, entry_centre(K, D, a[s]));
        if (K.mode) D.ukey[a[s]] = ckey;
        if (p[s] < 0) continue;
        Rec r2 = load_rec(D.recC, D.recS2, D.recS3, p[s]);
        shift_pose(&r2.cx, true, shx, shy, false); shift_pose(&r2.cx, true, PBx, PBy, true); rotate_pose(&r2.cx, true, cs, ss, cmx, cmy, cmz);
        if (cis_misaligned(K, r, r2)) snap_cis(K, r2, r);
        store_rec(D.recCn, D.recS2n, D.recS3n, p[s], r2);
        mark_far(cK, D, p[s], cp_[s].x, cp_[s].y, r2.cx, r2.cy, 0.0, 0.0, ckey, F_FREE_RL, stamp, entry_centre(K, D, p[s]));
        if (K.mode) D.ukey[p[s]] = ckey;
    }
    return true;
}

// one complex moved by one thread (S2d + S2e/S2f): the three-round single-ligand path where the shape allows it, the generic serial code otherwise
KD void complex_move_thread(const Args &A, int h0, bool maybeSingle, uint64_t step, unsigned stamp) {
    KARGS
    const Consts &K = cK;
    const int rootGid = K.NAt + h0;
    const int size = D.cxSize[h0];
    const uint64_t seed = seed_of(cK, replica_of_gid(K, K.NAt + h0));
    const uint32_t me = ref_id(K, D, rootGid);
    double u0, u1; keyed_uniform2(seed, me, 0, step, 0, u0, u1);
    const double u2 = keyed_uniform(seed, me, 0, step, 2);
    if (!maybeSingle || !single_ligand_move(A, h0, size, step, stamp, u0, u1, u2)) {
        const int *rowIn = D.members + D.cxOff[h0];
        int nB = 0;
        for (int i = 0; i < size; i++) nB += rowIn[i] >= K.NAt;
        complex_move_serial(A, h0, size, nB, step, stamp, u0, u1, u2);
    }
    D.unitRes[rootGid] = 0; D.pendCnt[rootGid] = 0;
}
// small complexes (the bulk of an oligomerised membrane: 2-12 members): one THREAD per complex. The alignment is a serial,
// branchy algorithm; a warp per complex leaves 31 of 32 lanes idle in it, a thread per complex runs 32 of them per warp.
__global__ void __launch_bounds__(128) k_propose_complex_small(const __grid_constant__ Args A, int withMulti) {
    KARGS
    const Consts &K = cK;
    // two lists back to back: single-ligand complexes, padded to a whole number of warps, then (unless k_propose_complex_multi
    // takes them) the multi-ligand ones
    const int n1 = D.scal[S_NCX], n1pad = (n1 + 31) & ~31, ncx = n1pad + (withMulti ? D.scal[S_NCX_MULTI] : 0);
    const uint64_t step = D.step64[0];
    const unsigned stamp = (unsigned)D.scal[S_EPOCH];
    for (int ci = blockIdx.x * blockDim.x + threadIdx.x; ci < ncx; ci += gridDim.x * blockDim.x) {
        if (ci >= n1 && ci < n1pad) continue;
        complex_move_thread(A, ci < n1 ? D.cxRoots[ci] : D.cxRoots[K.NBt + ci - n1pad], ci < n1, step, stamp);
    }
}

// G lanes per complex, NG complexes per CTA, CAP members cached; LIST 2: the large complexes (one warp each), LIST 1: the small
// complexes with several ligands (S2f: shuffles and passes) -- a serial algorithm on dependent gathers when one thread runs it on
// global memory (402 us for 94 000 two-ligand complexes); here the group loads the members side by side into shared memory,
// lane 0 runs the order-dependent parts there, the order-free passes and the draws of a shuffle go one member per lane
template <int G, int NG, int CAP, int LIST>
KD void propose_complex_groups(const Args &A) {
    KARGS
    __shared__ CxSharedT<CAP> SH[NG];
    const uint64_t step = D.step64[0];
    const unsigned stamp = (unsigned)D.scal[S_EPOCH];
    const Consts &K = cK;
    const int ncx = D.scal[LIST == 2 ? S_NCX_BIG : S_NCX_MULTI];
    const int lane = threadIdx.x & (G - 1), wib = threadIdx.x / G;
    const int warp = blockIdx.x * NG + wib, nwarps = gridDim.x * NG;
    CxSharedT<CAP> &S = SH[wib];
    for (int ci = warp; ci < ncx; ci += nwarps) {
        const int h0 = LIST == 2 ? D.cxRoots[K.NBt - 1 - ci] : D.cxRoots[K.NBt + ci], rootGid = K.NAt + h0;
        const int size = D.cxSize[h0];
        const int *rowIn = D.members + D.cxOff[h0];
        int *rowOut = D.rowWork + D.cxOff[h0];
        const uint64_t seed = seed_of(cK, replica_of_gid(K, K.NAt + h0));
        const uint32_t me = ref_id(K, D, rootGid);
        const bool cached = size <= CAP;
        int nB = 0;
        for (int i = 0; i < size; i++) nB += rowIn[i] >= K.NAt;          // (uniform across the warp, tiny)
        const int nA = size - nB;
        double u0, u1; keyed_uniform2(seed, me, 0, step, 0, u0, u1);
        const double u2 = keyed_uniform(seed, me, 0, step, 2);
        // ---- S2d rigid move (main.cpp:974-1131); complexes with >= 2 ligands have D = 0 but still draw ----
        const double amp = mul(nB == 1 ? K.ampBond : 0.0, u0);
        const double phai = mul(mul(u1, 2.0), K.pai);
        double sp, cp; KMC_SINCOS(K, phai, &sp, &cp);
        const double shx = mul(amp, cp), shy = mul(amp, sp);
        const double psai = mul(sub(mul(2.0, u2), 1.0), nB == 1 ? K.rotBond : 0.0);
        double ss, cs; KMC_SINCOS(K, psai, &ss, &cs);
        grp_sync<G>();
        if (cached) {
            // load: every lane its members (old pose + the bonds inside the complex as slots), translated by the shift
            for (int i = lane; i < size; i += G) {
                const int m = rowIn[i];
                S.gid[i] = m; S.row[i] = i; S.moved[i] = 0;
                double *p = S.pose[i];
                if (m < K.NAt) {
                    const Rec r = load_rec(D.recC, D.recS2, D.recS3, m);
                    p[0] = r.cx; p[1] = r.cy; p[2] = r.s2x; p[3] = r.s2y; p[4] = r.s3x; p[5] = r.s3y;
                    const int l = D.recLig[m], c = D.recCis[m];
                    S.lig[i] = l >= 0 ? (short)D.rowPos[K.NAt + l] : (short)-1; S.site[i] = (short)D.recSite[m]; S.cis[i] = c >= 0 ? (short)D.rowPos[c] : (short)-1;
                } else {
                    const double *q = D.lig + (size_t)(m - K.NAt) * 24;
                    for (int t = 0; t < 24; t++) p[t] = q[t];
                    for (int s = 0; s < 3; s++) { const int r = D.ligRec[(m - K.NAt) * 3 + s]; S.rec3[i][s] = r >= 0 ? (short)D.rowPos[r] : (short)-1; }
                }
                shift_pose(p, m < K.NAt, shx, shy, false);
            }
            grp_sync<G>();
            double PBx = 0, PBy = 0;
            if (lane == 0) {                                             // wrap centre: sum in row order (main.cpp:1007-1008, 1022-1023)
                for (int i = 0; i < size; i++) { PBx = add(PBx, S.pose[i][0]); PBy = add(PBy, S.pose[i][1]); }
                PBx = mul(K.Lx, round(dvd(dvd(PBx, (double)(nA + nB)), K.Lx)));
                PBy = mul(K.Ly, round(dvd(dvd(PBy, (double)(nA + nB)), K.Ly)));
            }
            PBx = grp_bcast<G>(PBx); PBy = grp_bcast<G>(PBy);
            for (int i = lane; i < size; i += G) shift_pose(S.pose[i], S.gid[i] < K.NAt, PBx, PBy, true);
            grp_sync<G>();
            double cmx = 0, cmy = 0, cmz = 0;
            if (lane == 0) {                                             // rotation centre: bead by bead in row order (main.cpp:1048-1068)
                for (int i = 0; i < size; i++) {
                    const double *p = S.pose[i];
                    if (S.gid[i] < K.NAt) for (int j = 1; j <= 4; j++) { cmx = add(cmx, p[0]); cmy = add(cmy, p[1]); cmz = add(cmz, rec_bead_z(K, j)); }
                    else for (int q = 0; q < 4; q++) { cmx = add(cmx, p[q * 3]); cmy = add(cmy, p[q * 3 + 1]); cmz = add(cmz, p[q * 3 + 2]); }
                }
                const double nbeads = (double)(4 * nA + 4 * nB);
                cmx = dvd(cmx, nbeads); cmy = dvd(cmy, nbeads); cmz = dvd(cmz, nbeads);
            }
            cmx = grp_bcast<G>(cmx); cmy = grp_bcast<G>(cmy); cmz = grp_bcast<G>(cmz);
            for (int i = lane; i < size; i += G) rotate_pose(S.pose[i], S.gid[i] < K.NAt, cs, ss, cmx, cmy, cmz);
            grp_sync<G>();
            { CxLocalT<CAP> C{S, K, K.NAt}; align_complex_warp<G, CAP>(C, S, size, nB, seed, me, step, lane); }     // the root ligand is slot 0
            // store: new poses, the working row (cluster.log order), far flags + grid histogram, order keys
            const int ckey = unit_key(K, rootGid, D.lig[(size_t)h0 * 24], D.lig[(size_t)h0 * 24 + 1]);
            for (int i = lane; i < size; i += G) {
                const int m = S.gid[i]; const double *p = S.pose[i];
                rowOut[i] = S.gid[S.row[i]];
                if (K.mode) D.ukey[m] = ckey;
                if (m < K.NAt) {
                    const Rec r = {p[0], p[1], p[2], p[3], p[4], p[5]};
                    store_rec(D.recCn, D.recS2n, D.recS3n, m, r);
                    const double2 o = D.recC[m]; mark_far(cK, D, m, o.x, o.y, r.cx, r.cy, 0.0, 0.0, ckey, (S.lig[i] < 0 ? F_FREE_RL : 0) | (S.cis[i] < 0 ? F_FREE_CIS : 0), stamp, entry_centre(K, D, m));
                } else {
                    double *q = D.lign + (size_t)(m - K.NAt) * 24;
                    for (int t = 0; t < 24; t++) q[t] = p[t];
                    const double *o = D.lig + (size_t)(m - K.NAt) * 24; mark_far(cK, D, m, o[0], o[1], p[0], p[1], o[2], p[2], ckey, (S.rec3[i][0] < 0 || S.rec3[i][1] < 0 || S.rec3[i][2] < 0) ? F_FREE_RL : 0, stamp, entry_centre(K, D, m));
                }
            }
        } else if (lane == 0) complex_move_serial(A, h0, size, nB, step, stamp, u0, u1, u2);      // larger than the cache
        if (lane == 0) { D.unitRes[rootGid] = 0; D.pendCnt[rootGid] = 0; }
        grp_sync<G>();
    }
}
__global__ void __launch_bounds__(32 * CX_WARPS, 6) k_propose_complex(const __grid_constant__ Args A) { propose_complex_groups<32, CX_WARPS, CX_CAP, 2>(A); }
__global__ void __launch_bounds__(CX_G * CX_GROUPS) k_propose_complex_multi(const __grid_constant__ Args A) { propose_complex_groups<CX_G, CX_GROUPS, CX_GCAP, 1>(A); }

// ------------------------------------------------------------------------------------------------
// neighbour grid: counting sort of molecule centres by cell (old centre; far movers get a second,
// "ghost" entry at their proposed centre so that a fixed 3x3 search is exact)
// ------------------------------------------------------------------------------------------------
KD void centre_of(const Consts &cK, const Dev &D, int gid, bool nxt, double &x, double &y) {
    if (gid < cK.NAt) { double2 c = nxt ? D.recCn[gid] : D.recC[gid]; x = c.x; y = c.y; }
    else { const double *p = (nxt ? D.lign : D.lig) + (size_t)(gid - cK.NAt) * 24; x = p[0]; y = p[1]; }
}
__global__ void k_grid_scatter(const __grid_constant__ Args A) {
    KARGS
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool cells = D.scen != nullptr;         // the thread-per-entry resolve wants the centre and the cell next to each entry
    if (gid_live(cK, D, gid)) {
        int rep = replica_of_gid(cK, gid);
        double x, y; centre_of(cK, D, gid, false, x, y);
        int c = cell_of(cK, rep, x, y);
        const int e = D.cellStart[c] + D.molSlot[gid];
        D.sorted[e] = gid;
        if (cells) { D.scen[e] = make_float2((float)x, (float)y); D.scell[e] = c; }
    }
    if (gid < D.scal[S_NFAR]) {
        int4 f = D.farList[gid];
        const int e = D.cellStart[f.y] + f.z;
        D.sorted[e] = f.x | GHOST_BIT;
        if (cells) { double x, y; centre_of(cK, D, f.x, true, x, y); D.scen[e] = make_float2((float)x, (float)y); D.scell[e] = f.y; }
    }
}
// exclusive scan of cellCount into cellStart in three phases (tile sums, scan of the sums, per-tile scan). Arrays are padded
// to a multiple of SCAN_TILE; 8 consecutive ints per thread (two int4), warp shuffles, one shared word per warp. The last
// phase also zeroes cellCount: the histogram of the NEXT step is accumulated by the propose kernels.
#define SCAN_TILE 2048
KD int warp_incl_scan(int v) {
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v += t; }
    return v;
}
__global__ void __launch_bounds__(256) k_scan_reduce(const int4 *in, int *blockSums) {
    __shared__ int sh[8];
    const int4 a = in[(size_t)blockIdx.x * (SCAN_TILE / 4) + threadIdx.x * 2], b = in[(size_t)blockIdx.x * (SCAN_TILE / 4) + threadIdx.x * 2 + 1];
    int s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) blockSums[blockIdx.x] = sh[0] + sh[1] + sh[2] + sh[3] + sh[4] + sh[5] + sh[6] + sh[7];
}
__global__ void __launch_bounds__(1024) k_scan_sums(int *blockSums, int nb) {     // single block, exclusive, in place
    __shared__ int sh[32]; __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x, v = i < nb ? blockSums[i] : 0;
        int inc = warp_incl_scan(v);
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) sh[threadIdx.x] = warp_incl_scan(sh[threadIdx.x]);
        __syncthreads();
        const int wofs = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
        if (i < nb) blockSums[i] = carry + wofs + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += wofs + inc;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(256) k_scan_down(int4 *in, const int *blockSums, int4 *out) {
    __shared__ int sh[8];
    const size_t q = (size_t)blockIdx.x * (SCAN_TILE / 4) + threadIdx.x * 2;
    const int4 a = in[q], b = in[q + 1];
    const int tot = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    const int inc = warp_incl_scan(tot);
    if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
    __syncthreads();
    int ofs = blockSums[blockIdx.x] + inc - tot;
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) ofs += sh[w];
    int4 oa, ob;
    oa.x = ofs; oa.y = oa.x + a.x; oa.z = oa.y + a.y; oa.w = oa.z + a.z;
    ob.x = oa.w + a.w; ob.y = ob.x + b.x; ob.z = ob.y + b.y; ob.w = ob.z + b.z;
    out[q] = oa; out[q + 1] = ob;
    in[q] = make_int4(0, 0, 0, 0); in[q + 1] = make_int4(0, 0, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// S2g with ordering: decide accept/reject of every unit. The same neighbour walk also collects the (few) molecule
// pairs that can possibly react in S3, so the reaction stage never walks the grid again.
// ------------------------------------------------------------------------------------------------
KD bool hit_rec_beads(const Consts &K, double ax, double ay, const double b[3][3]) {
    for (int j = 0; j < 3; j++) {
        double d2 = add(sq(sub(b[j][0], ax)), sq(sub(b[j][1], ay)));
        for (int k = 1; k <= 4; k++)
            if (add(d2, sq(sub(b[j][2], rec_bead_z(K, k)))) < K.ovAB2) return true;
    }
    return false;
}
KD bool hit_beads_beads(const Consts &K, const double a[3][3], const double b[3][3]) {
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 3; k++)
            if (add(add(sq(sub(a[j][0], b[k][0])), sq(sub(a[j][1], b[k][1]))), sq(sub(a[j][2], b[k][2]))) < K.ovBB2) return true;
    return false;
}
KD void load_beads(const double *base, int h, double b[3][3]) {
    const double *q = base + (size_t)h * 24 + 3;
#pragma unroll
    for (int i = 0; i < 9; i++) (&b[0][0])[i] = q[i];
}
// ------------------------------------------------------------------------------------------------
// S2g pass 1, tile kernel: one CTA owns a TS x TS block of grid cells. The cell-sorted entries of the block and of
// the ring of cells around it are staged in shared memory once (cooperative gather: centre old/new, unit, flags),
// then every molecule whose entry lies in the block is tested against its 3x3 cell neighbourhood out of shared
// memory. Receptor-receptor tests (the bulk) never touch global memory again; ligand beads are fetched only for
// pairs whose centres are within reach. The same walk pre-selects the pairs that can react in S3.
//
// Walk geometry: the entry of a molecule sits in the cell of its OLD centre O; its proposal P is at most `skin` away
// (else it is a far mover with a ghost entry at cell(P)), and so is every neighbour's, hence everything P can touch
// has an entry within reach + 2*skin of O: the 3x3 cells around cell(O) suffice because edge >= reach + 2*skin.
// ------------------------------------------------------------------------------------------------
#ifndef TS
#define TS 16
#endif
#ifndef TTHREADS
#define TTHREADS 128
#endif
#ifndef TCAP
#define TCAP 384
#endif

struct TileRec { double ox, oy, nx, ny; float oz, nz; int gid, unit; int flg; };

// fused small-system step: the poses are resident in shared memory, the record is read from them (there are no ghost entries)
KD TileRec fetch_rec_small(const Consts &K, const Dev &D, int v) {
    const int2 w = D.small->meta[small_index(K, D, v)];
    TileRec r; r.gid = v; r.unit = w.x; r.flg = w.y;
    if (v < K.NAt) { const double2 o = D.recC[v], n = D.recCn[v]; r.ox = o.x; r.oy = o.y; r.nx = n.x; r.ny = n.y; r.oz = 0.f; r.nz = 0.f; }
    else {
        const double *p = D.lig + (size_t)(v - K.NAt) * 24, *q = D.lign + (size_t)(v - K.NAt) * 24;
        r.ox = p[0]; r.oy = p[1]; r.oz = (float)p[2]; r.nx = q[0]; r.ny = q[1]; r.nz = (float)q[2];
    }
    return r;
}
KD TileRec fetch_rec(const Consts &K, const Dev &D, int entry) {
    const int v = entry & ~GHOST_BIT;
    const double2 *nr = reinterpret_cast<const double2 *>(D.nrec) + (size_t)v * 3;
    const double2 o = nr[0], n = nr[1];
    const int4 w = reinterpret_cast<const int4 *>(nr)[2];
    TileRec r;
    r.ox = o.x; r.oy = o.y; r.nx = n.x; r.ny = n.y; r.oz = __int_as_float(w.x); r.nz = __int_as_float(w.w); r.gid = v; r.unit = w.y; r.flg = w.z | ((entry & GHOST_BIT) ? F_GHOST : 0);
    return r;
}

struct TileSmem {
    int cs[TS + 2][TS + 3];       // cellStart of the window: rows wy0..wy1, columns wx0..wx1+1
    int rowBase[TS + 3];          // index (in staged order) of the first entry of each window row
    int inBase[TS + 1];           // prefix of interior entries per interior row
    int rowStart[TS + 3];         // index in `sorted` of the first entry of each window row
    double ox[TCAP], oy[TCAP], nx[TCAP], ny[TCAP];
    float oz[TCAP], nz[TCAP];     // height of a ligand's centre, old / proposed
    float sx[TCAP], sy[TCAP];     // window-relative fp32 copy of the centre the entry stands for (distance cut only)
    int gid[TCAP], unit[TCAP];
    unsigned char flg[TCAP];
};

// precise overlap test of probe (proposed pose) against neighbour v at old / proposed pose, centres already known
KD bool tile_hits(const Consts &K, const Dev &D, bool prec, double px, double py, const double pb[3][3], int v, double vx, double vy, bool nxt) {
    const bool vrec = v < K.NAt;
    const double dx = vx - px, dy = vy - py, d2 = dx * dx + dy * dy;
    if (prec && vrec) return hit_rec_rec(K, px, py, vx, vy);
    const double r = (prec || vrec) ? K.reachRL : K.reachLL;
    if (d2 > r * r) return false;
    if (vrec) return hit_rec_beads(K, vx, vy, pb);
    double ob[3][3]; load_beads(nxt ? D.lign : D.lig, v - K.NAt, ob);
    return prec ? hit_rec_beads(K, px, py, ob) : hit_beads_beads(K, ob, pb);
}

// the molecule an interior entry stands for, as a probe
struct ProbeCtx {
    int m, u, flg; bool ghost, far, prec, pairsOnly, wantPairs;
    double px, py;            // proposed centre P
    float pz;                 // its height (ligand)
    double fax, fay, fbx, fby; // the two candidate FINAL centres for S3 pre-selection (normal entry: P and O)
};
KD ProbeCtx make_probe(const Consts &K, const TileRec &me) {
    ProbeCtx c;
    c.m = me.gid; c.u = me.unit; c.flg = me.flg;
    c.ghost = me.flg & F_GHOST; c.far = me.flg & F_FAR; c.prec = me.gid < K.NAt;
    c.pairsOnly = !c.ghost && c.far;                 // old entry of a far mover: it is only a reaction partner here
    c.px = me.nx; c.py = me.ny; c.pz = me.nz;
    c.fax = c.pairsOnly ? me.ox : c.px; c.fay = c.pairsOnly ? me.oy : c.py;
    c.fbx = c.ghost ? c.px : me.ox;     c.fby = c.ghost ? c.py : me.oy;
    c.wantPairs = c.prec && (me.flg & (F_FREE_RL | F_FREE_CIS));
    return c;
}
// probe against one neighbour entry that survived the distance cut: S3 pre-selection + overlap classification.
// returns bit0 definite overlap, bit1 overlap with exactly one pose of an earlier (still undecided) unit; in that case
// *conf = that unit (bit 30 set if the overlapping pose is its NEW one)
// S3 candidate pairs found by a CTA are collected in shared memory and appended to D.pairs with ONE atomicAdd per CTA
// (a per-pair atomicAdd on the single global counter serialises ~1e5 atomics per step at the L2)
// (k_pairs_eval does not collect at all: a list pair remembers in one byte which of its two directions is a candidate, and
// k_react_pairs runs over the list again: flag != nullptr, bit = 1 for (first, second), 2 for (second, first))
struct PairSink { unsigned long long *buf; int *cnt; int cap; int *flag; int bit; };
KD void append_pair_global(const Dev &D, unsigned long long pr) {
    int p = atomicAdd(&D.scal[S_NPAIR], 1);
    if (p < D.pairFastCap) D.pairsFast[p] = pr;
    else if (p - D.pairFastCap < D.pairCap) D.pairs[p - D.pairFastCap] = pr;
    else atomicOr(&D.scal[S_OVERFLOW], 4);
}
KD void sink_pair(const Dev &D, const PairSink &ps, unsigned long long pr) {
    if (ps.flag) { *ps.flag |= ps.bit; return; }
    if (!ps.cnt) { append_pair_global(D, pr); return; }
    const int i = atomicAdd(ps.cnt, 1);
    if (i < ps.cap) ps.buf[i] = pr; else append_pair_global(D, pr);
}
// all threads of the CTA, after a __syncthreads() that follows the last sink_pair; `base` is a shared scratch word
KD void flush_pairs(const Dev &D, const PairSink &ps, int *base) {
    const int n = min(*ps.cnt, ps.cap);
    if (threadIdx.x == 0 && n > 0) *base = atomicAdd(&D.scal[S_NPAIR], n);
    __syncthreads();
    if (n > 0) {
        const int b = *base;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            if (b + i < D.pairCap) D.pairs[b + i] = ps.buf[i];
            else atomicOr(&D.scal[S_OVERFLOW], 4);
        }
    }
}
// lower bounds on a separation in z from the fp32 heights kept in the records (0.01 covers their rounding)
KD double z_excess(double dz) { return fmax(0.0, dz - 0.01); }
KD double z_above_stack(const Consts &K, float z) {          // from a ligand centre at height z to the receptor's bead stack [0, 6 rA]
    return z_excess(fmax((double)z - rec_bead_z(K, 4), -(double)z));
}
// cheap part, from the two records alone: S3 pre-selection; returns true if beads have to be looked at (pair_detail)
KD bool pair_pre(const Consts &K, const Dev &D, const ProbeCtx &c, const TileRec &o, const PairSink &ps) {
    const int v = o.gid;
    if (v == c.m) return false;                          // the other (old/ghost) entry of the probe itself
    const bool vghost = o.flg & F_GHOST, vfar = o.flg & F_FAR;
    if (c.wantPairs) {
        const bool vlig = v >= K.NAt;
        if (vlig ? ((c.flg & F_FREE_RL) && (o.flg & F_FREE_RL)) : ((c.flg & F_FREE_CIS) && (o.flg & F_FREE_CIS))) {
            const double reach = vlig ? K.reachOn : K.reachCis;
            const double vx = vghost ? o.nx : o.ox, vy = vghost ? o.ny : o.oy;
            // a ligand binds at the receptor's bead 3 (z = 2*rA*2): its centre must also be within reach of that height
            const double zs = rec_bead_z(K, 3);
            const double z1 = vlig ? z_excess(fabs((double)(vghost ? o.nz : o.oz) - zs)) : 0.0, z2 = vlig ? z_excess(fabs((double)o.nz - zs)) : 0.0;
            double d2 = fmin((vx - c.fax) * (vx - c.fax) + (vy - c.fay) * (vy - c.fay), (vx - c.fbx) * (vx - c.fbx) + (vy - c.fby) * (vy - c.fby)) + z1 * z1;
            if (!vghost && !vfar)
                d2 = fmin(d2, fmin((o.nx - c.fax) * (o.nx - c.fax) + (o.ny - c.fay) * (o.ny - c.fay), (o.nx - c.fbx) * (o.nx - c.fbx) + (o.ny - c.fby) * (o.ny - c.fby)) + z2 * z2);
            if (d2 <= reach * reach) sink_pair(D, ps, ((unsigned long long)c.m << 32) | (unsigned)v);
        }
    }
    if (c.pairsOnly) return false;
    {   // nothing of v within reach of the probe at either of v's poses: done before any bead is fetched (most list pairs end here)
        // (3-D: a ligand high above the membrane is out of every receptor's reach, two ligands at different heights miss each other)
        const bool vrec = v < K.NAt;
        double r, zo = 0.0, zn = 0.0;
        if (c.prec && vrec) r = K.reachRR;
        else if (c.prec) { r = K.reachRL; zo = z_above_stack(K, o.oz); zn = z_above_stack(K, o.nz); }
        else if (vrec) { r = K.reachRL; zo = zn = z_above_stack(K, c.pz); }
        else { r = K.reachLL; zo = z_excess(fabs((double)o.oz - (double)c.pz)); zn = z_excess(fabs((double)o.nz - (double)c.pz)); }
        const double ax = o.ox - c.px, ay = o.oy - c.py, bx = o.nx - c.px, by = o.ny - c.py;
        if (fmin(ax * ax + ay * ay + zo * zo, bx * bx + by * by + zn * zn) > r * r) return false;
    }
    return true;
}
// exact classification of a pair that passed pair_pre
KD int pair_detail(const Consts &K, const Dev &D, const ProbeCtx &c, const TileRec &o, int *conf) {
    const int v = o.gid;
    const bool vghost = o.flg & F_GHOST, vfar = o.flg & F_FAR;
    double pb[3][3];
    if (!c.prec) load_beads(D.lign, c.m - K.NAt, pb);
    const int uv = o.unit;
    if (uv == c.u) {                                     // co-moving member: proposed pose, once (Q20)
        if (vghost != vfar) return 0;
        return tile_hits(K, D, c.prec, c.px, c.py, pb, v, o.nx, o.ny, true) ? 1 : 0;
    }
    if (!unit_before(uv, c.u))                           // later unit: still at its old pose
        return (!vghost && tile_hits(K, D, c.prec, c.px, c.py, pb, v, o.ox, o.oy, false)) ? 1 : 0;
    if (vghost) {                                        // earlier unit (undecided in pass 1): both poses possible
        if (!tile_hits(K, D, c.prec, c.px, c.py, pb, v, o.nx, o.ny, true)) return 0;
        *conf = (uv & UNIT_MASK) | 0x40000000; return 2;
    }
    const bool hitOld = tile_hits(K, D, c.prec, c.px, c.py, pb, v, o.ox, o.oy, false);
    const bool hitNew = !vfar && tile_hits(K, D, c.prec, c.px, c.py, pb, v, o.nx, o.ny, true);
    if (hitOld && hitNew) return 1;
    if (hitOld || hitNew) { *conf = (uv & UNIT_MASK) | (hitNew ? 0x40000000 : 0); return 2; }
    return 0;
}
KD int pair_eval(const Consts &K, const Dev &D, const ProbeCtx &c, const TileRec &o, int *conf, const PairSink &ps) {
    return pair_pre(K, D, c, o, ps) ? pair_detail(K, D, c, o, conf) : 0;
}
// publishes a probe's result into the unit head's word unitRes: 0 = nothing found (accept), bit0 = definite overlap (reject),
// bit1 = undecided. A finding that depends on ONE pose of an earlier unit (res == 2) is appended to the pending list as
// (unit, earlier unit | pose bit); after the pass the pending findings alone decide what is left (k_pend_resolve): no geometry
// is ever evaluated twice and nothing after the pass needs the neighbour grid.
KD void reject_unit(const Dev &D, int u) {              // whoever sets the bit first also lists the unit for the copy-back
    if (!(atomicOr(&D.unitRes[u], 1) & 1)) {
        const int i = atomicAdd(&D.scal[S_NREJ], 1);
        D.rejList[i] = u;
        D.rejPartner[i] = u < D.nAcap ? D.recCis[u] : -1;       // (S3 has not run yet: the cis word is the one the unit moved with)
    }
}
KD void publish(const Dev &D, int ukey, int res, int conf) {
    if (!res) return;
    const int u = ukey & UNIT_MASK;
    if (res & 1) { reject_unit(D, u); return; }
    atomicOr(&D.unitRes[u], 2);
    atomicAdd(&D.pendCnt[u], 1);
    const int i = atomicAdd(&D.scal[S_NPEND], 1);
    if (i < D.pendCap) D.pendList[i] = make_int2(u, conf);
    else atomicOr(&D.scal[S_OVERFLOW], 16);
}

#ifndef NSURV
#define NSURV 512
#endif
#ifndef TMINB
#define TMINB 8
#endif
#define TPAIRS 128
__global__ void __launch_bounds__(TTHREADS, TMINB) k_resolve_tiles(const __grid_constant__ Args A) {
    KARGS
    const Consts &K = cK;
    __shared__ TileSmem S;
    __shared__ unsigned int surv[NSURV];          // (staged index of probe << 16) | staged index of neighbour
    __shared__ int nsurv;
    __shared__ unsigned long long pbuf[TPAIRS];
    __shared__ int pcnt, pbase;
    const PairSink ps = {pbuf, &pcnt, TPAIRS, nullptr, 0};
    const int ts = K.tileEdge;        // tile edge in cells (<= TS), chosen from the mean cell occupancy so that a window fits the staging buffers
    const int ntx = (K.ncx + ts - 1) / ts, nty = (K.ncy + ts - 1) / ts;
    int b = blockIdx.x;
    const int tx = b % ntx; b /= ntx;
    const int ty = b % nty; const int rep = b / nty;
    const int x0 = tx * ts, x1 = min(x0 + ts - 1, K.ncx - 1), y0 = ty * ts, y1 = min(y0 + ts - 1, K.ncy - 1);
    const int wx0 = max(x0 - 1, 0), wx1 = min(x1 + 1, K.ncx - 1), wy0 = max(y0 - 1, 0), wy1 = min(y1 + 1, K.ncy - 1);
    const int nrow = wy1 - wy0 + 1, ncol = wx1 - wx0 + 2, nir = y1 - y0 + 1;
#ifdef KMC_TILE_TIMING
    long long tq[6]; int tqi = 0;
#define TICK() do { if (threadIdx.x == 0) tq[tqi++] = clock64(); } while (0)
#else
#define TICK() do {} while (0)
#endif
    TICK();
    if (threadIdx.x < 32) {
        // one warp: the extent of every window row / interior row in `sorted` (4 loads per row), then their prefixes
        const int r = threadIdx.x;
        int s0 = 0, s1 = 0, i0 = 0, i1 = 0;
        if (r < nrow) {
            const int *row = D.cellStart + (size_t)(rep * K.ncy + wy0 + r) * K.ncx;
            s0 = __ldg(row + wx0); s1 = __ldg(row + wx1 + 1);
            if (r >= y0 - wy0 && r < y0 - wy0 + nir) { i0 = __ldg(row + x0); i1 = __ldg(row + x1 + 1); }
        }
        const int len = s1 - s0, lin = i1 - i0;
        const int a = warp_incl_scan(len), bI = warp_incl_scan(lin);
        if (r <= nrow) { S.rowBase[r] = a - len; S.rowStart[r] = s0; }
        // interior rows are window rows y0-wy0 .. : store their prefix indexed by interior row
        const int ir = r - (y0 - wy0);
        if (ir >= 0 && ir <= nir) S.inBase[ir] = bI - lin;
        if (r == 0) { nsurv = 0; pcnt = 0; }
    }
    __syncthreads();
    TICK();
    const int total = S.rowBase[nrow], nst = min(total, TCAP);
    const double tox = K.gx0 + (double)wx0 / K.cellInv, toy = K.gy0 + (double)wy0 / K.cellInv;    // window origin
    // stage the window (first TCAP entries; the rest, if a tile is that crowded, is read from global memory on demand)
    // and, in the same latency window, the cellStart values of the window (column ranges of the 3x3 walks)
    for (int i = threadIdx.x; i < nst; i += TTHREADS) {
        int lo = 0, hi = nrow;                      // largest r with rowBase[r] <= i
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (S.rowBase[mid] <= i) lo = mid; else hi = mid; }
        const TileRec t = fetch_rec(K, D, __ldg(&D.sorted[S.rowStart[lo] + (i - S.rowBase[lo])]));
        S.ox[i] = t.ox; S.oy[i] = t.oy; S.nx[i] = t.nx; S.ny[i] = t.ny; S.oz[i] = t.oz; S.nz[i] = t.nz; S.gid[i] = t.gid; S.unit[i] = t.unit; S.flg[i] = (unsigned char)t.flg;
        const bool g = t.flg & F_GHOST;                                   // what this entry stands for, window-relative, fp32 (cut only)
        S.sx[i] = (float)(hash_x(K, g ? t.nx : t.ox) - tox); S.sy[i] = (float)((g ? t.ny : t.oy) - toy);
    }
    for (int r = threadIdx.x / 32; r < nrow; r += TTHREADS / 32)
        for (int c = threadIdx.x & 31; c < ncol; c += 32)
            S.cs[r][c] = __ldg(&D.cellStart[(size_t)(rep * K.ncy + wy0 + r) * K.ncx + wx0 + c]);
    __syncthreads();
    TICK();
    auto staged = [&](int i) -> TileRec { TileRec t; t.ox = S.ox[i]; t.oy = S.oy[i]; t.nx = S.nx[i]; t.ny = S.ny[i]; t.oz = S.oz[i]; t.nz = S.nz[i]; t.gid = S.gid[i]; t.unit = S.unit[i]; t.flg = S.flg[i]; return t; };
    const int nin = S.inBase[nir];
    const bool fits = total <= TCAP;
    // squared cut radii (fp32, with a safety margin): overlap reach + one skin (the neighbour's pose moves at most a skin from
    // its entry), S3 pre-selection reach + two skins (the probe's own final centre may be its old one)
    const float mg = 0.05f, sk = (float)K.skin;
    const float cRR = fmaxf((float)K.ovAA + sk, (float)K.reachCis + 2 * sk) + mg, cRL = fmaxf((float)K.reachRL + sk, (float)K.reachOn + 2 * sk) + mg,
                cLR = (float)K.reachRL + sk + mg, cLL = (float)K.reachLL + sk + mg;
    // phase 1: every interior entry scans its 3x3 cells with a cheap distance cut; survivors go to a shared list
    for (int q = threadIdx.x; q < nin; q += TTHREADS) {
        int lo = 0, hi = nir;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (S.inBase[mid] <= q) lo = mid; else hi = mid; }
        const int wr = y0 - wy0 + lo;
        const int eme = S.cs[wr][x0 - wx0] + (q - S.inBase[lo]);          // global entry index of this interior entry
        const int ime = S.rowBase[wr] + (eme - S.rowStart[wr]);            // its staged index
        if (fits) {
            const bool prec = S.gid[ime] < K.NAt;
            const int f = S.flg[ime];
            // centre the cut is measured from: P, or O for the old entry of a far mover (which is only a reaction partner)
            const bool pOnly = (f & F_FAR) && !(f & F_GHOST);
            const float fx = (float)(hash_x(K, pOnly ? S.ox[ime] : S.nx[ime]) - tox), fy = (float)((pOnly ? S.oy[ime] : S.ny[ime]) - toy);
            const int cxe = min(max((int)floor((hash_x(K, f & F_GHOST ? S.nx[ime] : S.ox[ime]) - K.gx0) * K.cellInv), 0), K.ncx - 1);
            const int cxl = max(cxe - 1, 0) - wx0, cxh = min(cxe + 1, K.ncx - 1) + 1 - wx0;
            const float cR = prec ? cRR : cLR, cL = prec ? cRL : cLL;
            const float cR2 = cR * cR, cL2 = cL * cL;
#pragma unroll
            for (int dr = -1; dr <= 1; dr++) {
                const int r = wr + dr;
                if (r < 0 || r >= nrow) continue;
                const int ib = S.rowBase[r] - S.rowStart[r];
                const int i0 = ib + S.cs[r][cxl], i1 = ib + S.cs[r][cxh];
                for (int i = i0; i < i1; i++) {
                    const float ex = S.sx[i] - fx, ey = S.sy[i] - fy;
                    const float lim = (S.gid[i] < K.NAt) ? cR2 : cL2;
                    if (ex * ex + ey * ey > lim || i == ime) continue;
                    const int slot = atomicAdd(&nsurv, 1);
                    if (slot < NSURV) surv[slot] = ((unsigned)ime << 16) | (unsigned)i;
                    else {                                                  // survivor list full: evaluate in place
                        const ProbeCtx c = make_probe(K, staged(ime));
                        int cf = -1; const int rr = pair_eval(K, D, c, staged(i), &cf, ps);
                        publish(D, c.u, rr, cf);
                    }
                }
            }
        } else {
            // crowded tile: generic path, entries beyond the staged part are fetched from global memory
            const TileRec me = ime < TCAP ? staged(ime) : fetch_rec(K, D, __ldg(&D.sorted[eme]));
            const ProbeCtx c = make_probe(K, me);
            const double wxp = c.ghost ? me.nx : me.ox;
            const int cxe = min(max((int)floor((hash_x(K, wxp) - K.gx0) * K.cellInv), 0), K.ncx - 1);
            const int cxl = max(cxe - 1, 0) - wx0, cxh = min(cxe + 1, K.ncx - 1) + 1 - wx0;
            const double cutR = (c.prec ? fmax(K.reachRL, K.reachOn) : K.reachLL) + 2 * K.skin, cut2 = cutR * cutR;
            for (int dr = -1; dr <= 1; dr++) {
                const int r = wr + dr;
                if (r < 0 || r >= nrow) continue;
                const int e0 = S.cs[r][cxl], e1 = S.cs[r][cxh], ib = S.rowBase[r] - S.rowStart[r];
                for (int e = e0; e < e1; e++) {
                    const int i = ib + e;
                    const TileRec o = i < TCAP ? staged(i) : fetch_rec(K, D, __ldg(&D.sorted[e]));
                    const bool g = o.flg & F_GHOST;
                    const double ex = (g ? o.nx : o.ox) - c.fax, ey = (g ? o.ny : o.oy) - c.fay;
                    if (ex * ex + ey * ey > cut2) continue;
                    int cf = -1; const int rr = pair_eval(K, D, c, o, &cf, ps);
                    publish(D, c.u, rr, cf);
                }
            }
        }
    }
    __syncthreads();
    TICK();
    // phase 2: one thread per surviving (probe, neighbour) pair
    const int ns = min(nsurv, NSURV);
    for (int s = threadIdx.x; s < ns; s += TTHREADS) {
        const unsigned w = surv[s];
        const ProbeCtx c = make_probe(K, staged((int)(w >> 16)));
        int cf = -1;
        const int rr = pair_eval(K, D, c, staged((int)(w & 0xffffu)), &cf, ps);
        publish(D, c.u, rr, cf);
    }
    __syncthreads();
    flush_pairs(D, ps, &pbase);
#ifdef KMC_TILE_TIMING
    __syncthreads();
    TICK();
    if (threadIdx.x == 0) { for (int i = 0; i < 4; i++) atomicAdd(&D.events[10 + i], (unsigned long long)(tq[i + 1] - tq[i])); atomicAdd(&D.events[15], 1ULL); }
#endif
}
// S2g pass 1 for SPARSE membranes (a fraction of a molecule per cell, the reference's own density), two kernels, no staging.
//
// k_cells_cut: one thread per grid entry in cell-sorted order. The scatter leaves, next to `sorted`, an fp32 copy of the
// centre every entry stands for (`scen`), so the distance cut touches only cell-sorted, contiguous data: the entry's own 12
// bytes, the six row extents of its 3x3 cells (shared with the neighbouring threads through L1) and 12 bytes per candidate.
// It is a light kernel (no fp64, few registers, full occupancy); pairs that survive the cut are collected per CTA and appended
// to a global list with one atomicAdd per CTA.
// k_pairs_eval: one thread per surviving (probe entry, neighbour entry) pair: the same pair_eval as the tile kernel, from the
// per-molecule records -- full warps instead of the few survivor lanes of a fused kernel.
// The tile kernel remains the path for crowded cells, where staging a window once pays for the many candidates per probe.
#ifndef CTHREADS
#define CTHREADS 256
#endif
#ifndef CSURV
#define CSURV 1024
#endif
#ifndef CMINB
#define CMINB 8
#endif
// classification of one surviving pair straight from the records (also the overflow path of k_cells_cut: not inlined there,
// so the cut keeps its small register footprint)
KD void eval_rec_pair(const Consts &K, const Dev &D, const TileRec &a, const TileRec &b, const PairSink &ps) {
#pragma unroll 1
    for (int dir = 0; dir < 2; dir++) {
        int cf = -1;
        const ProbeCtx pc = make_probe(K, dir ? b : a);
        PairSink pd = ps; pd.bit = 1 << dir;
        const int rr = pair_eval(K, D, pc, dir ? a : b, &cf, pd);
        publish(D, pc.u, rr, cf);
    }
}
// an unordered pair of grid entries: each entry is the probe once
KD void eval_entry_pair(const Consts &K, const Dev &D, int entry, int en, const PairSink &ps) {
    eval_rec_pair(K, D, fetch_rec(K, D, entry), fetch_rec(K, D, en), ps);
}
__device__ __noinline__ void eval_entry_pair_slow(const Args &A, int entry, int en) {
    const PairSink none = {nullptr, nullptr, 0, nullptr, 0};
    eval_entry_pair(A.K, A.D, entry, en, none);
}
__global__ void __launch_bounds__(CTHREADS, CMINB) k_cells_cut(const __grid_constant__ Args A) {
    KARGS
    const Consts &K = cK;
    __shared__ int2 surv[CSURV];                  // (entry of the probe, entry of the neighbour), ghost bits included
    __shared__ int nsurv, gbase;
    const int total = __ldg(&D.cellStart[D.ncell]);
    // cut radii, measured between the centres the two ENTRIES stand for: every pose of either molecule that matters here lies
    // within one skin of its entry (proposal of a near mover around its old entry; a far mover has one entry per pose), so
    // overlap reach + 2 skins and S3 reach + 2 skins bound every test pair_eval can make. fp32 with a margin that covers the
    // rounding of the stored coordinates (K.cutMargin, from the box size).
    const float mg = K.cutMargin, sk2 = 2 * (float)K.skin + 2 * (float)K.drift;     // (+ the drift allowed while the list is reused)
    // every UNORDERED pair of entries is listed once (by its lower entry index) and classified in both directions, so the
    // receptor-ligand radius is the larger of the two directions' needs
    const float cRR = fmaxf((float)K.ovAA, (float)K.reachCis) + sk2 + mg, cRL = fmaxf((float)K.reachRL, (float)K.reachOn) + sk2 + mg,
                cLR = cRL, cLL = (float)K.reachLL + sk2 + mg;
    const int *__restrict__ sorted = D.sorted;
    const float2 *__restrict__ scen = D.scen;
    for (int base = blockIdx.x * CTHREADS; base < total; base += gridDim.x * CTHREADS) {
        if (threadIdx.x == 0) nsurv = 0;
        __syncthreads();
        const int e = base + threadIdx.x;
        if (e < total) {
            const int entry = __ldg(&sorted[e]);
            const float2 w = __ldg(&scen[e]);                                   // centre this entry stands for (grid frame): cut centre
            const int cell = __ldg(&D.scell[e]);                                // its cell (exact, from the scatter)
            const bool prec = (entry & ~GHOST_BIT) < K.NAt;
            const int crow = cell / K.ncx, cx = cell - crow * K.ncx;            // crow = replica * ncy + cy
            const int cy = crow % K.ncy;
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, K.ncx - 1);
            // only partners with a HIGHER entry index (the pair is listed once): the rest of this entry's own row of cells,
            // then the row above; entries of the row below come earlier in the cell order and list this entry themselves
            int e0[2], e1[2];
            e0[0] = e + 1; e1[0] = __ldg(D.cellStart + (size_t)crow * K.ncx + x1 + 1);
            e0[1] = e1[1] = 0;
            if (cy < K.ncy - 1) { const int *row = D.cellStart + (size_t)(crow + 1) * K.ncx; e0[1] = __ldg(row + x0); e1[1] = __ldg(row + x1 + 1); }
            const float cR2 = prec ? cRR * cRR : cLR * cLR, cL2 = prec ? cRL * cRL : cLL * cLL;
#pragma unroll
            for (int r = 0; r < 2; r++)
                for (int i = e0[r]; i < e1[r]; i++) {
                    const int en = __ldg(&sorted[i]);
                    const float2 c = __ldg(&scen[i]);
                    const float ex = c.x - w.x, ey = c.y - w.y;
                    if (ex * ex + ey * ey > ((en & ~GHOST_BIT) < K.NAt ? cR2 : cL2)) continue;
                    const int slot = atomicAdd(&nsurv, 1);
                    if (slot < CSURV) surv[slot] = make_int2(entry, en);
                    else {                                                      // CTA list full (crowded spot): append one by one
                        const int g = atomicAdd(&D.scal[S_NSURV], 1);
                        if (g < D.survCap) D.surv[g] = make_int2(entry, en);
                        else { eval_entry_pair_slow(A, entry, en); atomicOr(&D.scal[S_OVERFLOW], 32); }
                    }
                }
        }
        __syncthreads();
        const int ns = min(nsurv, CSURV);
        if (threadIdx.x == 0 && ns > 0) gbase = atomicAdd(&D.scal[S_NSURV], ns);
        __syncthreads();
        if (ns > 0) {
            const int gb = gbase;
            for (int s = threadIdx.x; s < ns; s += CTHREADS) {
                if (gb + s < D.survCap) D.surv[gb + s] = surv[s];
                else { eval_entry_pair_slow(A, surv[s].x, surv[s].y); atomicOr(&D.scal[S_OVERFLOW], 32); }   // global list full: exact for this step, but the list cannot be reused (the host reports it unless it rebuilds every step)
            }
        }
        __syncthreads();
    }
}
#ifndef PTHREADS
#define PTHREADS 128
#endif
#ifndef PMINB
#define PMINB 5
#endif
// Two phases per chunk of the list. Phase A, one thread per list pair: the two records, S3 pre-selection and the 3-D early out
// for both directions -- every lane busy, no bead touched; the (few) directions that need beads are queued in shared memory.
// Phase B, one thread per queued direction: the exact classification, with full warps instead of the 2-3 lanes per warp
// that reach it when it is done in place.
#ifndef PE_CHUNK
#define PE_CHUNK 512
#endif
__global__ void __launch_bounds__(PTHREADS, PMINB) k_pairs_eval(const __grid_constant__ Args A) {
    KARGS
    const Consts &K = cK;
    __shared__ int items[2 * PE_CHUNK];          // (list index << 1) | direction
    __shared__ int ritems[2 * PE_CHUNK];         // the same coding: directions that may react in S3 (appended to D.reactList per chunk)
    __shared__ int nitems, nreact, rbase;
    const int ns = min(D.scal[S_NSURV], D.survCap);
    for (int base = blockIdx.x * PE_CHUNK; base < ns; base += gridDim.x * PE_CHUNK) {
        if (threadIdx.x == 0) { nitems = 0; nreact = 0; }
        __syncthreads();
        for (int s = base + threadIdx.x; s < min(base + PE_CHUNK, ns); s += PTHREADS) {
            const int2 w = D.surv[s];
            int flag = 0;
            // reuse step: ghost entries belong to the build step only; a molecule that is special this step (far mover or
            // displaced) is not where the list believes it to be and is handled by k_special_pairs instead
            if (K.phase == 0 || !((w.x | w.y) & GHOST_BIT)) {
                const TileRec a = fetch_rec(K, D, w.x), b = fetch_rec(K, D, w.y);
                if (K.phase == 0 || !((a.flg | b.flg) & (F_FAR | F_DISP))) {
                    PairSink ps = {nullptr, nullptr, 0, &flag, 1};
                    if (pair_pre(K, D, make_probe(K, a), b, ps)) items[atomicAdd(&nitems, 1)] = s << 1;
                    ps.bit = 2;
                    if (pair_pre(K, D, make_probe(K, b), a, ps)) items[atomicAdd(&nitems, 1)] = (s << 1) | 1;
                }
            }
            if (flag & 1) ritems[atomicAdd(&nreact, 1)] = s << 1;           // which directions of this pair may react in S3 (k_react_pairs)
            if (flag & 2) ritems[atomicAdd(&nreact, 1)] = (s << 1) | 1;
        }
        __syncthreads();
        const int ni = nitems, nr = nreact;
        if (threadIdx.x == 0 && nr) rbase = atomicAdd(&D.scal[S_NREACT], nr);          // one global atomic per chunk
        __syncthreads();
        for (int q = threadIdx.x; q < nr; q += PTHREADS) D.reactList[rbase + q] = ritems[q];
        for (int q = threadIdx.x; q < ni; q += PTHREADS) {
            const int it = items[q];
            const int2 w = D.surv[it >> 1];
            const TileRec a = fetch_rec(K, D, (it & 1) ? w.y : w.x), b = fetch_rec(K, D, (it & 1) ? w.x : w.y);
            const ProbeCtx pc = make_probe(K, a);
            int cf = -1;
            const int rr = pair_detail(K, D, pc, b, &cf);
            publish(D, pc.u, rr, cf);
        }
        __syncthreads();
    }
}
// List-reuse steps: the molecules the stale grid / pair list do not cover. A special entry stands for one centre X of its
// molecule f (a far mover has two: old centre, and proposal with the ghost bit; a displaced molecule one: its old centre,
// the proposal is within a skin of it). One warp per special entry walks the 3x3 cells around X in the STALE grid:
//   * every ordinary molecule n entered there is still within drift + skin of its entry, so the walk finds everything f can
//     touch from X (cell edge >= reach + 2 skins + 2 drifts); both directions are classified here, (f, n) and (n, f),
//     because the list pairs of f are skipped by k_pairs_eval;
//   * other special entries are found through the per-cell chains built by mark_far (direction (f, g); g's warp does (g, f)).
// pair_eval is exact for any two entries, so duplicates (a molecule with a leftover ghost entry in the stale grid) only repeat
// idempotent findings.
#define SP_WARPS 4
__global__ void __launch_bounds__(32 * SP_WARPS) k_special_pairs(const __grid_constant__ Args A) {
    KARGS
    const Consts &K = cK;
    __shared__ unsigned long long pbuf[128];
    __shared__ int pcnt, pbase;
    const PairSink ps = {pbuf, &pcnt, 128, nullptr, 0};
    const int ns = min(D.scal[S_NSPEC], 2 * K.NT);
    const unsigned stamp = (unsigned)D.scal[S_EPOCH];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    for (int base = blockIdx.x * SP_WARPS; base < ns; base += gridDim.x * SP_WARPS) {
        if (threadIdx.x == 0) pcnt = 0;
        __syncthreads();
        const int si = base + wib;
        if (si < ns) {
            const int ef = D.specList[si], f = ef & ~GHOST_BIT;
            const TileRec rf = fetch_rec(K, D, ef);
            const ProbeCtx pf = make_probe(K, rf);
            const double X = (ef & GHOST_BIT) ? rf.nx : rf.ox, Y = (ef & GHOST_BIT) ? rf.ny : rf.oy;
            const int rep = replica_of_gid(K, f);
            int cx = (int)floor((hash_x(K, X) - K.gx0) * K.cellInv), cy = (int)floor((Y - K.gy0) * K.cellInv);
            cx = min(max(cx, 0), K.ncx - 1); cy = min(max(cy, 0), K.ncy - 1);
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, K.ncx - 1), y0 = max(cy - 1, 0), y1 = min(cy + 1, K.ncy - 1);
            for (int yy = y0; yy <= y1; yy++) {
                const int rowc = (rep * K.ncy + yy) * K.ncx;
                const int e0 = D.cellStart[rowc + x0], e1 = D.cellStart[rowc + x1 + 1];
                for (int i = e0 + lane; i < e1; i += 32) {
                    const int n = D.sorted[i] & ~GHOST_BIT;
                    if (n == f) continue;
                    const TileRec rn = fetch_rec(K, D, n);
                    if (rn.flg & (F_FAR | F_DISP)) continue;              // special as well: met through the chains below
                    eval_rec_pair(K, D, rf, rn, ps);
                }
            }
            // the chains of special entries of the same 3x3 cells, one cell per lane
            const int ncolw = x1 - x0 + 1, ncw = ncolw * (y1 - y0 + 1);
            if (lane < ncw) {
                const int cell = (rep * K.ncy + y0 + lane / ncolw) * K.ncx + x0 + lane % ncolw;
                unsigned long long link = D.cellHead[cell];
                for (int guard = 0; guard < ns && (unsigned)(link >> 32) == stamp && (unsigned)link != 0; guard++) {
                    const int gi = (int)(unsigned)link - 1;
                    const int eg = D.specList[gi];
                    link = D.specNext[gi];
                    if ((eg & ~GHOST_BIT) == f) continue;
                    int cf = -1;
                    const int rr = pair_eval(K, D, pf, fetch_rec(K, D, eg), &cf, ps);
                    publish(D, pf.u, rr, cf);
                }
            }
        }
        __syncthreads();
        flush_pairs(D, ps, &pbase);
        __syncthreads();
    }
}

// Order dependence of the sweep (main.cpp:642: an overlap test sees the NEW pose of an earlier molecule that was accepted and
// the OLD pose of one that was reverted). After the pass every unit is accepted (unitRes 0), rejected (bit0) or waits on
// pending findings "overlaps earlier unit v at exactly one of v's two poses". Such a finding is a hit iff v ends up at that
// pose; a unit with a hit is rejected, a unit whose findings all turn out to be misses is accepted. One CTA sweeps the (short)
// list until nothing changes: the lowest undecided unit only waits on decided ones, so every sweep makes progress.
KD int unit_state(const Dev &D, int u) {                 // U_ACCEPT / U_REJECT / U_UNKNOWN from the unit's word
    const int r = ((volatile int *)D.unitRes)[u];
    return (r & 1) ? U_REJECT : (r == 0 ? U_ACCEPT : U_UNKNOWN);
}
KD void restore_pose(const Consts &cK, const Dev &D, int gid) {
    if (gid < cK.NAt) { D.recCn[gid] = D.recC[gid]; D.recS2n[gid] = D.recS2[gid]; D.recS3n[gid] = D.recS3[gid]; }
    else {
        const double2 *s = reinterpret_cast<const double2 *>(D.lig + (size_t)(gid - cK.NAt) * 24);
        double2 *d = reinterpret_cast<double2 *>(D.lign + (size_t)(gid - cK.NAt) * 24);
        for (int q = 0; q < 12; q++) d[q] = s[q];
    }
}
// one sweep over entries first, first + stride, ...: returns true if anything was decided
KD bool pend_sweep(const Dev &D, int n, int first, int stride) {
    bool changed = false;
    for (int i = first; i < n; i += stride) {
        const int2 rec = D.pendList[i];                              // (only this thread ever rewrites entry i)
        if (rec.x < 0) continue;                                        // settled in an earlier sweep
        const int u = rec.x, v = rec.y & UNIT_MASK; const bool onNew = rec.y & 0x40000000;
        bool done = false;
        if (unit_state(D, u) != U_UNKNOWN) done = true;                 // u already rejected through another finding
        else {
            const int sv = unit_state(D, v);
            if (sv != U_UNKNOWN) {
                done = true;
                if ((sv == U_ACCEPT) == onNew) { reject_unit(D, u); changed = true; }                   // v sits at the pose u overlaps
                else if (atomicSub(&D.pendCnt[u], 1) == 1) { atomicCAS(&D.unitRes[u], 2, 0); changed = true; }   // last finding, all misses
            }
        }
        if (done) D.pendList[i].x = -1;
    }
    return changed;
}
// the fixed point by ONE CTA (short lists: a few hundred findings per step on the 1.25e6-molecule membrane)
KD void pend_resolve_block(const Dev &D, int n) {
    for (int sweep = 0; sweep <= n; sweep++) {
        const bool changed = pend_sweep(D, n, threadIdx.x, blockDim.x);
        __threadfence();
        if (!__syncthreads_or(changed)) break;          // (one barrier per sweep: it also carries the "anything decided?" vote)
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {           // cannot happen: the pending findings always settle
        const int u = D.pendList[i].x;
        if (u >= 0 && unit_state(D, u) == U_UNKNOWN) atomicOr(&D.scal[S_OVERFLOW], 8);
    }
}

// ------------------------------------------------------------------------------------------------
// S3 reactions
// ------------------------------------------------------------------------------------------------
// one thread per pre-selected pair (receptor a, neighbour v), final poses: geometric tests of main.cpp:1882-1915 /
// 1960-1981 / 2014-2035; candidates whose keyed draw succeeds are appended (a failed draw never changes anything,
// main.cpp:1921/1987/2041); the ordered, first-come-first-served application happens in k_react_resolve
template <bool SMALL = false>
KD void react_pair(const Consts &K, const Dev &D, uint64_t step, int a, int v) {
    // everything that depends only on (a, v) is requested at once: unit words, bond words and the PROPOSED poses (the final
    // pose of a molecule is its proposal unless its unit was reverted -- the copy-back happens in k_finish --, which is rare:
    // the old pose is fetched only then)
    // (fused small-system step: the unit key written with this step's proposal carries the head)
    const int ua = SMALL ? (D.small->meta[small_index(K, D, a)].x & UNIT_MASK) : D.unitOf[a], uv = SMALL ? (D.small->meta[small_index(K, D, v)].x & UNIT_MASK) : D.unitOf[v];
    Rec ra = load_rec(D.recCn, D.recS2n, D.recS3n, a);
    if (v >= K.NAt) {
        const int h = v - K.NAt;
        const int la = D.recLig[a];
        const int occ[3] = {D.ligRec[h * 3], D.ligRec[h * 3 + 1], D.ligRec[h * 3 + 2]};
        Lig b; load_lig(D.lign, h, b);
        const bool anyRej = !SMALL || D.scal[S_NREJ] > 0;          // (fused small-system step: the replica's own count, in shared memory -- most steps reject nothing and skip the two global loads)
        const bool rejA = anyRej && (D.unitRes[ua] & 1), rejV = anyRej && (D.unitRes[uv] & 1);
        if (la >= 0) return;
        if (rejA) ra = load_rec(D.recC, D.recS2, D.recS3, a);
        if (rejV) load_lig(D.lig, h, b);
        int okMask = 0;          // (SMALL: the three sites as a loop, one copy of the geometry test in the instruction stream)
#pragma unroll (SMALL ? 1 : 3)
        for (int s = 0; s < 3; s++) if (occ[s] < 0 && rl_geometry_ok(K, ra, b, s)) okMask |= 1 << s;
        if (!okMask) return;
        const uint64_t seed = seed_of(K, replica_of_gid(K, a));
        const uint32_t me = ref_id(K, D, a), j = ref_id(K, D, v);
#pragma unroll (SMALL ? 1 : 3)
        for (int s = 0; s < 3; s++) {
            if (!(okMask >> s & 1)) continue;
            if (keyed_uniform<SMALL>(seed, me, 4 * j + (uint32_t)(s + 2), step, SLOT_RL_ON) < K.pOn) {
                int q = atomicAdd(&D.scal[S_NCAND_RL], 1);
                if (q < D.candCap) D.candRL[q] = ((unsigned long long)a << 32) | ((unsigned long long)h << 2) | (unsigned)s;
                else atomicOr(&D.scal[S_OVERFLOW], 1);
            }
        }
    } else {
        const int ca = D.recCis[a], cv = D.recCis[v];
        Rec rb = load_rec(D.recCn, D.recS2n, D.recS3n, v);
        const bool anyRej = !SMALL || D.scal[S_NREJ] > 0;          // (fused small-system step: the replica's own count, in shared memory -- most steps reject nothing and skip the two global loads)
        const bool rejA = anyRej && (D.unitRes[ua] & 1), rejV = anyRej && (D.unitRes[uv] & 1);
        if (ca >= 0 || cv >= 0) return;
        if (rejA) ra = load_rec(D.recC, D.recS2, D.recS3, a);
        if (rejV) rb = load_rec(D.recC, D.recS2, D.recS3, v);
        if (!cis_geometry_ok(K, ra, rb)) return;
        const uint64_t seed = seed_of(K, replica_of_gid(K, a));
        const uint32_t me = ref_id(K, D, a), j = ref_id(K, D, v);
        const bool okMono = keyed_uniform<SMALL>(seed, me, j, step, SLOT_MONO_CIS_ON) < K.pMonoCisOn;
        const bool okCis = keyed_uniform<SMALL>(seed, me, j, step, SLOT_CIS_ON) < K.pCisOn;
        if (okMono || okCis) {
            int q = atomicAdd(&D.scal[S_NCAND_CIS], 1);
            if (q < D.candCap) D.candCis[q] = ((unsigned long long)a << 32) | ((unsigned long long)v << 2) | (okMono ? 1u : 0u) | (okCis ? 2u : 0u);
            else atomicOr(&D.scal[S_OVERFLOW], 2);
        }
    }
}
// work items: the list pairs k_pairs_eval found within reaction reach (sparse path: D.reactList, a dense list, one thread per
// entry), then the pairs collected by the tile kernel / the special entries (D.pairs).
template <bool SMALL = false>
KD void react_pairs_body(const Consts &K, const Dev &D, int tid, int nth) {
    const uint64_t step = D.step64[0];
    const int nl = D.reactList ? min(D.scal[S_NREACT], 2 * D.survCap) : 0;
    for (int q = tid; q < nl; q += nth) {
        const int it = D.reactList[q];
        const int2 w = D.surv[it >> 1];
        const int a = w.x & ~GHOST_BIT, b = w.y & ~GHOST_BIT;
        if (it & 1) react_pair<SMALL>(K, D, step, b, a); else react_pair<SMALL>(K, D, step, a, b);
    }
    const int np = min(D.scal[S_NPAIR], D.pairFastCap + D.pairCap);
    for (int i = tid; i < np; i += nth) {
        const unsigned long long pr = i < D.pairFastCap ? D.pairsFast[i] : D.pairs[i - D.pairFastCap];
        react_pair<SMALL>(K, D, step, (int)(pr >> 32), (int)(pr & 0xffffffffu));
    }
}

// single CTA: order the successful candidates as the reference's loops would meet them, then apply them
// first-come-first-served (main.cpp:1877-1949, 1952-2003, 2007-2058)
KD void react_resolve_block(const Dev &D) {
    unsigned long long *rl = D.candRL, *cis = D.candCis;
    const int nRL = min(D.scal[S_NCAND_RL], D.candCap), nCis = min(D.scal[S_NCAND_CIS], D.candCap);
    // ---- rank sort (n = successful draws of one step, tiny in practice; O(n^2/threads)) ----
    for (int pass = 0; pass < 2; pass++) {
        unsigned long long *k = pass == 0 ? rl : cis;
        int n = pass == 0 ? nRL : nCis;
        unsigned long long *tmp = k + D.candCap;        // second half of the buffer is scratch
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            unsigned long long me = k[t]; int rank = 0;
            for (int q = 0; q < n; q++) { unsigned long long o = k[q]; rank += (o < me) || (o == me && q < t); }
            tmp[rank] = me;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += blockDim.x) k[t] = tmp[t];
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    int ev_rl = 0, ev_mono = 0, ev_cis = 0;
    for (int q = 0; q < nRL; q++) {
        unsigned long long key = rl[q];
        if (q && key == rl[q - 1]) continue;             // the same pair can be pre-selected twice (far movers)
        int a = (int)(key >> 32), h = (int)((key & 0xffffffffULL) >> 2), s = (int)(key & 3);
        if (D.recLig[a] < 0 && D.ligRec[h * 3 + s] < 0) {
            D.recLig[a] = h; D.recSite[a] = s; D.ligRec[h * 3 + s] = a; ev_rl++;
            touch_molecule(D, a);
        }
    }
    for (int variant = 1; variant <= 2; variant++)
        for (int q = 0; q < nCis; q++) {
            unsigned long long key = cis[q];
            if (!(key & (unsigned)variant) || (q && key == cis[q - 1])) continue;
            int a = (int)(key >> 32), b = (int)((key & 0xffffffffULL) >> 2);
            if (D.recCis[a] >= 0 || D.recCis[b] >= 0) continue;
            bool anyLig = D.recLig[a] >= 0 || D.recLig[b] >= 0;
            if ((variant == 1) == anyLig) continue;      // variant 1: both ligand-free; variant 2: at least one bound
            D.recCis[a] = b; D.recCis[b] = a;
            touch_molecule(D, a);
            if (variant == 1) ev_mono++; else ev_cis++;
        }
    // (atomic: in the fused small-system step every replica's CTA applies its own candidates)
    if (ev_rl) atomicAdd(&D.events[EV_RL_ON], (unsigned long long)ev_rl);
    if (ev_mono) atomicAdd(&D.events[EV_MONO_ON], (unsigned long long)ev_mono);
    if (ev_cis) atomicAdd(&D.events[EV_CIS_ON], (unsigned long long)ev_cis);
}

// S3c, main.cpp:2062-2141, for receptor a with ligand h / cis partner p (-1 none). Keyed draws make the three sequential loops order
// free: a thread owns the R-L bond of its receptor and the cis bond it is the lower index of; it re-derives the partner's R-L
// outcome from the partner's own keyed draw instead of waiting for it.
template <bool SMALL = false>
KD void dissociate(const Consts &K, const Dev &D, uint64_t step, int a, int h, int p) {
    const uint64_t seed = seed_of(K, replica_of_gid(K, a));
    const uint32_t me = ref_id(K, D, a);
    bool boundAfter = false;
    if (h >= 0) {
        if (keyed_uniform<SMALL>(seed, me, 0, step, SLOT_RL_OFF) < K.pOff) {
            int s = D.recSite[a];
            D.recLig[a] = -1; D.recSite[a] = -1; D.ligRec[h * 3 + s] = -1;
            atomicAdd(&D.events[EV_RL_OFF], 1ULL); touch_molecule(D, a); touch_molecule(D, K.NAt + h);
        } else boundAfter = true;
    }
    if (p > a) {
        const uint32_t pid = ref_id(K, D, p);
        // partner's ligand state after ITS R-L dissociation trial: -1 already cleared, else apply its draw
        bool pBoundAfter = D.recLig[p] >= 0 && !(keyed_uniform<SMALL>(seed, pid, 0, step, SLOT_RL_OFF) < K.pOff);
        bool inComplex = boundAfter || pBoundAfter;
        uint32_t slot = inComplex ? SLOT_CIS_OFF : SLOT_MONO_CIS_OFF;
        double P = inComplex ? K.pCisOff : K.pMonoCisOff;
        // drawn from both ends: the lower index first, the partner only if the bond survived
        if (keyed_uniform<SMALL>(seed, me, 0, step, slot) < P || keyed_uniform<SMALL>(seed, pid, 0, step, slot) < P) {
            D.recCis[a] = -1; D.recCis[p] = -1;
            atomicAdd(&D.events[inComplex ? EV_CIS_OFF : EV_MONO_OFF], 1ULL); touch_molecule(D, a); touch_molecule(D, p);
        }
    }
}
// Last kernel of a step.
// (1) The members of the rejected units get their old pose back (main.cpp:666-674, 851-863, 1831-1860), one thread per listed
//     unit. Membership is the one of THIS step's sweep: the member table for a ligand-headed unit, the cis partner recorded at
//     rejection time for a receptor-headed one -- whatever S3 did to the bonds since.
// (2) Dissociation: a thread takes four consecutive receptors; their bond words arrive as two 16-byte loads, and on a
//     membrane with few bonds that is all the kernel reads (8 bytes per receptor).
// PARTS: 1 = the restore of the rejected units, 2 = the dissociation trials, 3 = both. (The restore needs nothing of S3: the step
// graph runs it on a side branch next to k_react_pairs / k_react_resolve, the trials after them.)
template <bool RANGED, int PARTS = 3>          // RANGED: receptors [aBeg, aEnd) take their dissociation trials here (a replica, fused small-system step); else all live ones
KD void finish_body(const Consts &K, const Dev &D, int tid, int nth, int aBeg, int aEnd) {
    if (!RANGED) { aBeg = 0; aEnd = nA_live(D); }
    const int nrej = (PARTS & 1) ? min(D.scal[S_NREJ], K.NT) : 0;
    if (tid == 0 && nrej) atomicAdd(&D.events[EV_REVERTED], (unsigned long long)nrej);
    for (int i = tid; i < nrej; i += nth) {
        const int u = D.rejList[i];
        if (u < K.NAt) { restore_pose(K, D, u); const int q = D.rejPartner[i]; if (q >= 0) restore_pose(K, D, q); continue; }
        const int h = u - K.NAt, size = D.cxSize[h];
        if (size <= 1) restore_pose(K, D, u);
        else { const int *row = D.members + D.cxOff[h]; for (int q = 0; q < size; q++) restore_pose(K, D, row[q]); }
    }
    if (!(PARTS & 2)) return;
    const uint64_t step = D.step64[0];
    for (int a0 = (aBeg & ~3) + tid * 4; a0 < aEnd; a0 += nth * 4) {
        int hh[4] = {-1, -1, -1, -1}, pp[4] = {-1, -1, -1, -1};
        if (a0 + 3 < K.NAt) {
            const int4 hv = *reinterpret_cast<const int4 *>(D.recLig + a0), pv = *reinterpret_cast<const int4 *>(D.recCis + a0);
            hh[0] = hv.x; hh[1] = hv.y; hh[2] = hv.z; hh[3] = hv.w; pp[0] = pv.x; pp[1] = pv.y; pp[2] = pv.z; pp[3] = pv.w;
        } else for (int k = 0; k < 4; k++) if (a0 + k < K.NAt) { hh[k] = D.recLig[a0 + k]; pp[k] = D.recCis[a0 + k]; }
        if ((hh[0] & hh[1] & hh[2] & hh[3] & pp[0] & pp[1] & pp[2] & pp[3]) < 0) continue;      // all eight words negative: no bond on any of the four
        for (int k = 0; k < 4; k++)
            if ((!RANGED || a0 + k >= aBeg) && a0 + k < aEnd && (hh[k] >= 0 || pp[k] >= 0)) dissociate<RANGED>(K, D, step, a0 + k, hh[k], pp[k]);
    }
}

// The tail of a step: four dependent kernels on lists of a few hundred to a few ten thousand items. (Measured on the B200: the
// same four phases as ONE cooperative kernel with grid-wide barriers take 53 us instead of 47 -- the lists want many short-lived
// threads, not a persistent grid that spins at barriers while one block settles the pending findings.)
//   k_pend_resolve   the order dependence of the sweep is settled from the pending findings (one CTA, fixed point)
//   k_react_pairs    S3 candidates: geometric tests + keyed draws, one thread per pre-selected pair (main.cpp:1877-2058)
//   k_react_resolve  the successful candidates are applied in the reference's loop order, first come first served (one CTA)
//   k_finish         rejected units get their old pose back; dissociation trials (main.cpp:2062-2141)
__global__ void __launch_bounds__(1024) k_pend_resolve(const __grid_constant__ Args A) { KARGS pend_resolve_block(D, min(D.scal[S_NPEND], D.pendCap)); }
#ifndef RPTHREADS
#define RPTHREADS 64
#endif
__global__ void __launch_bounds__(RPTHREADS) k_react_pairs(const __grid_constant__ Args A) { KARGS react_pairs_body<false>(cK, D, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x); }
__global__ void k_react_resolve(const __grid_constant__ Args A) { KARGS react_resolve_block(D); }
__global__ void __launch_bounds__(256) k_finish(const __grid_constant__ Args A, int parts) {
    KARGS
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (parts == 3) finish_body<false, 3>(cK, D, tid, nth, 0, 0); else if (parts == 1) finish_body<false, 1>(cK, D, tid, nth, 0, 0); else finish_body<false, 2>(cK, D, tid, nth, 0, 0);
}

// ------------------------------------------------------------------------------------------------
// outputs (bond.dat columns, main.cpp:2251)
// ------------------------------------------------------------------------------------------------
__global__ void k_series(const __grid_constant__ Args A, int *out /*[R][4]: rl, mono, cis, -*/) {
    KARGS
    int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nA_live(D)) return;
    int rep = a / cK.NA;
    if (D.recLig[a] >= 0) atomicAdd(&out[rep * 4 + 0], 1);
    int p = D.recCis[a];
    if (p > a) atomicAdd(&out[rep * 4 + ((D.recLig[a] >= 0 || D.recLig[p] >= 0) ? 2 : 1)], 1);
}

// tot_cluster_num / tot_proteins_in_cluster (main.cpp:976-977) from the complex table of the last step: out[rep*2 + {0,1}]
__global__ void k_cx_stats(const __grid_constant__ Args A, int *out) {
    KARGS
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nB_live(D) || D.unitOf[cK.NAt + h] != cK.NAt + h) return;
    const int size = D.cxSize[h];
    if (size > 1) { atomicAdd(&out[(h / cK.NB) * 2], 1); atomicAdd(&out[(h / cK.NB) * 2 + 1], size); }
}

}  // namespace kmc
