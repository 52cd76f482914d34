// csrc/kmc_small.cu -- the whole time step (main.cpp:461-2202) of a SMALL system as one kernel: one CTA per replica, many
// steps per launch (BASELINE configs[2]: ensembles of the reference's default 200-molecule system).
//
// The general path (kmc_kernels.cu) spends one CUDA-graph node per stage: 13 dependent nodes per step, ~30 us for a system
// whose arithmetic takes a microsecond. Here the stages of a step are separated by __syncthreads() instead of kernel
// boundaries, a replica never leaves its SM, and nothing crosses between replicas, so the launch covers any number of steps.
//
// It is the SAME step: every stage calls the device functions of kmc_kernels.cu (propose_one_rec, lig_move_*,
// complex_move_thread, pair_eval/publish, pend_sweep, react_pair, react_resolve_block, finish_body) through a per-replica VIEW of
// the device state -- a copy of Args in shared memory whose work lists point at this replica's slices of the global lists and
// whose scalars (list lengths, step counter) live in shared memory. Only the neighbour search differs: a replica's CTA cuts ALL
// pairs of its molecules with an fp32 test on per-molecule search records (old centre, reach share + displacement of this
// step) held in shared memory, so there is no grid, no far-mover bookkeeping and no list reuse; the pairs that survive are
// classified by the same pair_eval as everywhere else. Results are bit-identical to the general path (tests/test_gpu_small.py)
// and to the oracle (every small-system test of tests/test_gpu_replay.py runs through this kernel).
#pragma once

namespace kmc {

#ifndef SMALL_T
#define SMALL_T 128           // threads per replica
#endif
#ifndef SMALL_MINB
#define SMALL_MINB 5          // CTAs per SM the register allocation aims at
#endif
#define SMALL_MAXN 512        // molecules per replica the shared-memory records hold
#define SMALL_MAXNB 128       // ligands per replica (complex work list)
#define SMALL_SURV 1024       // pairs that survive the cut, per step

struct SmallShared {
    Args view;                          // per-replica view of the device state (see above)
    int scal[S_COUNT];
    unsigned long long step64;
    float4 cen[SMALL_MAXN];             // search records: old centre (x, y), radius share of the molecule this step
    int touch[TOUCH_CAP];
    int cxList[SMALL_MAXNB];            // root ligands of the complexes with more than one member | bit 30: several ligands
    unsigned surv[SMALL_SURV];          // (local index a << 16) | local index b
    int ncx, nsurv, dirty;
};

KD int small_gid(const Consts &K, int rep, int m) { return m < K.NA ? rep * K.NA + m : K.NAt + rep * K.NB + (m - K.NA); }

// S1 for one replica (main.cpp:514-562): union-find over its bond graph, unit heads, sizes, breadth-first member rows; the member
// rows of replica r live in members[r*N, (r+1)*N)
KD void small_rebuild(SmallShared &sm, int rep) {
    const Dev &D = sm.view.D; const Consts &K = sm.view.K;
    const int N = K.NA + K.NB, tid = threadIdx.x;
    for (int m = tid; m < N; m += SMALL_T) {
        const int gid = small_gid(K, rep, m);
        const int uid = gid < K.NAt ? K.NBt + gid : gid - K.NAt;
        D.ufParent[uid] = uid; D.bfsMark[gid] = 0;
        if (gid >= K.NAt) { D.cxSize[gid - K.NAt] = 0; D.cxOff[gid - K.NAt] = -1; }
    }
    if (tid == 0) { sm.scal[S_MEMBER_CURSOR] = rep * N; sm.ncx = 0; }
    __syncthreads();
    for (int m = tid; m < K.NA; m += SMALL_T) {
        const int a = rep * K.NA + m, ua = K.NBt + a;
        const int l = D.recLig[a];
        if (l >= 0) uf_union(D.ufParent, ua, l);
        const int c = D.recCis[a];
        if (c > a) uf_union(D.ufParent, ua, K.NBt + c);
    }
    __syncthreads();
    for (int m = tid; m < N; m += SMALL_T) {
        const int gid = small_gid(K, rep, m);
        const int uid = gid < K.NAt ? K.NBt + gid : gid - K.NAt;
        const int r = uf_find(D.ufParent, uid);
        D.unitOf[gid] = r < K.NBt ? K.NAt + r : r - K.NBt;
        if (r < K.NBt) atomicAdd(&D.cxSize[r], 1);
    }
    __syncthreads();
    for (int b = tid; b < K.NB; b += SMALL_T) {
        const int h = rep * K.NB + b;
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
KD void small_propose_lig(const Args &A, uint64_t step, int gid) {
    KARGS
    const Consts &K = cK;
    const int h = gid - K.NAt;
    if (D.unitOf[gid] != gid || D.cxSize[h] > 1) return;          // member of a complex: moved by complex_move_thread
    const double *__restrict__ src = D.lig + (size_t)h * 24;
    double *__restrict__ dst = D.lign + (size_t)h * 24;
    const double ox = src[0], oy = src[1], oz = src[2];
    const uint64_t seed = seed_of(K, replica_of_gid(K, gid));
    const uint32_t me = ref_id(K, D, gid);
    double u0, u1, u2, u3, u4, u5;
    keyed_uniform2(seed, me, 0, step, 0, u0, u1); keyed_uniform2(seed, me, 0, step, 2, u2, u3); keyed_uniform2(seed, me, 0, step, 4, u4, u5);
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

__global__ void __launch_bounds__(SMALL_T, SMALL_MINB) k_small_step(const __grid_constant__ Args A, unsigned long long step0, int nsteps) {
    __shared__ SmallShared sm;
    const int tid = threadIdx.x, rep = blockIdx.x;
    const int NA = A.K.NA, NB = A.K.NB, N = NA + NB, R = A.K.R;
    if (tid == 0) {
        // the view: this replica's slices of the work lists, scalars in shared memory
        sm.view = A;
        Dev &V = sm.view.D;
        V.scal = sm.scal; V.step64 = &sm.step64; V.touchList = sm.touch; V.reactList = nullptr;
        V.smallCen = sm.cen; V.smallRecBase = rep * NA; V.smallLigBase = A.K.NAt + rep * NB - NA;
        const int pendPer = A.D.pendCap / R, pairPer = A.D.pairCap / R, candPer = A.D.candCap / R;
        V.pendList = A.D.pendList + (size_t)rep * pendPer; V.pendCap = pendPer;
        V.pairs = A.D.pairs + (size_t)rep * pairPer; V.pairCap = pairPer;
        V.candRL = A.D.candRL + (size_t)rep * 2 * candPer; V.candCis = A.D.candCis + (size_t)rep * 2 * candPer; V.candCap = candPer;
        V.rejList = A.D.rejList + (size_t)rep * N; V.rejPartner = A.D.rejPartner + (size_t)rep * N;
        for (int i = 0; i < S_COUNT; i++) sm.scal[i] = 0;
        sm.scal[S_NA_LIVE] = A.K.NAt; sm.scal[S_NB_LIVE] = A.K.NBt;
        sm.dirty = 1;                    // the complex tables are re-derived at the start of every launch
    }
    __syncthreads();
    const Args &V = sm.view;
    const Dev &D = V.D; const Consts &K = V.K;
    const PairSink none = {nullptr, nullptr, 0, nullptr, 0};
    for (int s = 0; s < nsteps; s++) {
        const uint64_t step = step0 + (uint64_t)s + 1;
        if (tid == 0) {
            sm.step64 = step;
            sm.scal[S_NPEND] = 0; sm.scal[S_NPAIR] = 0; sm.scal[S_NCAND_RL] = 0; sm.scal[S_NCAND_CIS] = 0; sm.scal[S_NREJ] = 0; sm.scal[S_NTOUCH] = 0; sm.scal[S_TOPO_DIRTY] = 0;
            sm.nsurv = 0;
        }
        // ---- S1 ----
        if (sm.dirty) small_rebuild(sm, rep);          // (uniform: sm.dirty was written before the last barrier)
        __syncthreads();
        // ---- S2 proposals: every molecule is proposed by exactly one thread (a unit head, or the thread of its complex) ----
        for (int m = tid; m < N; m += SMALL_T) {
            const int gid = small_gid(K, rep, m);
            if (m < NA) propose_one_rec(V, step, 0u, K.NAt, gid, D.unitOf[gid], D.recCis[gid], load_rec(D.recC, D.recS2, D.recS3, gid), make_float2(0.f, 0.f), ref_id(K, D, gid));
            else small_propose_lig(V, step, gid);
        }
        for (int ci = tid; ci < sm.ncx; ci += SMALL_T) complex_move_thread(V, sm.cxList[ci] & 0x3fffffff, !(sm.cxList[ci] & 0x40000000), step, 0u);
        __syncthreads();
        // ---- S2g: all pairs of the replica, fp32 cut on the search records; survivors are classified exactly ----
        {
            const int half = N / 2;
            for (int m = tid; m < N; m += SMALL_T) {
                const float4 me = sm.cen[m];
                const int dmax = (2 * half == N && m >= half) ? half - 1 : half;      // even N: the pair (m, m + N/2) is listed by its lower member
                int j = m;
                for (int d = 1; d <= dmax; d++) {
                    if (++j == N) j = 0;
                    const float4 o = sm.cen[j];
                    const float ex = o.x - me.x, ey = o.y - me.y, r = me.z + o.z;
                    if (ex * ex + ey * ey > r * r) continue;
                    const int slot = atomicAdd(&sm.nsurv, 1);
                    if (slot < SMALL_SURV) sm.surv[slot] = ((unsigned)m << 16) | (unsigned)j;
                    else eval_entry_pair(K, D, small_gid(K, rep, m), small_gid(K, rep, j), none);
                }
            }
        }
        __syncthreads();
        {
            const int ns = min(sm.nsurv, SMALL_SURV);
            for (int q = tid; q < ns; q += SMALL_T) {
                const unsigned w = sm.surv[q];
                eval_entry_pair(K, D, small_gid(K, rep, (int)(w >> 16)), small_gid(K, rep, (int)(w & 0xffffu)), none);
            }
        }
        __syncthreads();
        // ---- the order dependence of the sweep, from the pending findings ----
        if (sm.scal[S_NPEND] > 0) pend_resolve_block(D, min(sm.scal[S_NPEND], D.pendCap));
        // ---- S3 ----
        react_pairs_body(K, D, tid, SMALL_T);
        __syncthreads();
        if (sm.scal[S_NCAND_RL] | sm.scal[S_NCAND_CIS]) react_resolve_block(D);
        __syncthreads();
        finish_body(K, D, tid, SMALL_T, rep * NA, (rep + 1) * NA);
        __syncthreads();
        // ---- S4: the new buffers become the committed state ----
        if (tid == 0) {
            Dev &W = sm.view.D;
            double2 *t2; double *t1;
            t2 = W.recC; W.recC = W.recCn; W.recCn = t2; t2 = W.recS2; W.recS2 = W.recS2n; W.recS2n = t2; t2 = W.recS3; W.recS3 = W.recS3n; W.recS3n = t2;
            t1 = W.lig; W.lig = W.lign; W.lign = t1;
            sm.dirty = (sm.scal[S_NTOUCH] > 0) | sm.scal[S_TOPO_DIRTY];
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (sm.scal[S_OVERFLOW]) atomicOr(&A.D.scal[S_OVERFLOW], sm.scal[S_OVERFLOW]);
        if (rep == 0) A.D.step64[0] = step0 + (unsigned long long)nsteps;
    }
}

}  // namespace kmc
