// csrc/kmc_engine.cu -- host side of the C ABI (include/kmc_b200.h): device memory, the per-step launch
// sequence, state import/export in the reference's array shapes, and the reference's output records.
// There is no CPU compute path in this file: every stage of the sweep is a kernel in kmc_kernels.cu.
#include "../../include/kmc_b200.h"
#include "kmc_kernels.cu"
#include "kmc_small.cu"
#define KMC_NKERNELS 23
#define MON_EVERY 256

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace kmc;

static thread_local std::string g_create_error;

// ---- strip decomposition bookkeeping (csrc/kmc_strips.cu) ----
struct StripDev;
struct RecMsg { int32_t ref, ligRef, site, cisRef; double pose[6]; };      // 64 bytes; refs are reference ids (1-based), 0 = none
struct LigMsg { int32_t ref, recRef[3]; double pose[24]; };               // 208 bytes
static_assert(sizeof(RecMsg) == 64 && sizeof(LigMsg) == 208, "message records have no padding");

struct HostLocal {       // host mirror of the live local state
    std::vector<double> rec, lig; std::vector<int> rl, rs, rc, lr; std::vector<unsigned> refA, refB;
    int nA = 0, nB = 0;
};


struct kmc_handle {
    kmc_params P;
    Consts K;
    Dev D;
    cudaStream_t stream = nullptr, side[2] = {nullptr, nullptr};      // side streams: forked branches of the step (graph)
    cudaEvent_t evFork[3] = {nullptr, nullptr, nullptr}, evJoin[4] = {nullptr, nullptr, nullptr, nullptr};
    std::string err;
    int R = 1, NA = 0, NB = 0, N = 0, NAt = 0, NBt = 0, NT = 0;
    int64_t step_done = 0;
    bool stepped = false;            // complexes/accept data of a completed step are available
    std::vector<void *> allocs;
    int *d_series = nullptr;
    int scanBlocks = 0, nTiles = 0;
    bool useCells = false;      // pass 1 of the resolve: thread-per-entry kernel (sparse cells) instead of the tile kernel
    // list reuse (sparse path): every listEvery-th step rebuilds the neighbour grid and the pair list (phase 0), the steps in
    // between reuse them (phase 1); sinceBuild = steps taken since the last rebuild, 0 = the next step must rebuild
    int listEvery = 1, sinceBuild = 0, cellHeadCap = 0;
    bool adapt = true, adapted = false;      // kmc_sync may fall back to a rebuild every step (see there)
    int adaptLevel = 0; bool adaptSkip = false;      // back-off taken so far (0 none, 1 wide list every 4th step, 2 rebuild every step); ignore the next snapshot
    unsigned epoch = 0;
    cudaGraphExec_t gexec[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};      // [phase][buffer parity]
    int parity = 0, launches_per_step[2] = {0, 0};
    bool use_graph = true;
    double *stageRec = nullptr; int *stageInt = nullptr;      // device staging of kmc_set_packed / kmc_get_packed
    // kmc_get_packed_async: device-side snapshot (taken on the handle's stream) + copy stream that moves it to the host while stepping goes on
    double *snapRec = nullptr, *snapLig = nullptr; int *snapInt = nullptr;
    cudaStream_t copyStream = nullptr; cudaEvent_t evSnap = nullptr, evSnapDone = nullptr; bool snapPending = false;
    // strips
    bool strip_on = false; double strip_W = 0, strip_lo = 0, strip_hi = 0; int64_t strip_refreshes = 0;
    double strip_t[3] = {0, 0, 0};             // KMC_STRIP_TIMING: accumulated wall-clock of the refresh phases (us)
    int strip_every = 0, strip_since = 0;      // > 0: kmc_step refreshes the halos itself every strip_every steps (kmc_strip_comm_init)
    HostLocal strip_local; std::vector<char> strip_msg[3];
    struct StripDev *strip_dev = nullptr;
    int64_t launches = 0, passes = 0;
    int init_rounds = 0;             // rounds the GPU generator needed (diagnostics)
    int cxBlocks = 0;                // grid of the cooperative rebuild kernel
    unsigned long long *timeline = nullptr; int tlCount = 0, tlId[64]; cudaStream_t tlStream = nullptr;      // KMC_TIMELINE
    int nSM = 148;                   // multiprocessors of the device (cudaDeviceProp): persistent grids are sized from it
    int forkMask = 14;               // KMC_FORK, read once at kmc_create (bit 3: the restore of rejected units beside S3)
    bool cxGroups = true;            // small multi-ligand complexes by groups of 8 lanes on a shared-memory copy (KMC_CX_GROUPS=0: one thread each, on global memory)
    bool smallWide = false;          // fused step with 256 threads per replica (ensembles of at most one replica per SM)
    int smallSlots = 1;              // replicas per CTA of the fused step (1, or 4 in lockstep for ensembles that fill the device)
    int smallGrid = 0; int *smallQueue = nullptr;      // fused step: CTAs resident at once; ticket queue (1 + R ints) for ensembles larger than that
    bool fused = false;              // small replicas: the whole step is ONE kernel, one CTA per replica, many steps per launch (csrc/kmc_small.cu)
    // in-flight monitoring of long kmc_step calls: every MON_EVERY steps the device scalars are copied to pinned host memory
    // (asynchronously) and the copy of the PREVIOUS interval is examined -- capacity overflows are reported and the list-reuse
    // back-off is decided without ever stalling the stream
    int *monHost = nullptr; cudaEvent_t monEvent = nullptr; bool monPending = false; int64_t sinceMon = 0;
    // optional per-kernel timing with CUDA events on the handle's stream (bench.py roofline)
    bool profiling = false;
    struct Pending { int id; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> evpool;
    double kms[KMC_NKERNELS] = {0};
    int64_t kcount[KMC_NKERNELS] = {0};
};

static const char *const g_kernel_names[KMC_NKERNELS] = {
    "k_cx_rebuild", "k_uf_hook", "k_uf_flatten", "k_cx_build", "k_propose_rec",
    "k_propose_complex", "k_scan_reduce", "k_scan_sums", "k_scan_down", "k_grid_scatter",
    "k_resolve_tiles", "k_pend_resolve", "k_react_pairs", "k_react_resolve", "k_finish", "k_series", "k_pairs_eval", "k_special_pairs", "k_propose_lig", "k_propose_complex_small", "k_step_begin", "k_small_step", "k_propose_complex_multi"};
enum { KID_UF_INIT = 0, KID_UF_HOOK, KID_UF_FLATTEN, KID_CX_BUILD, KID_PROPOSE_SIMPLE, KID_PROPOSE_COMPLEX,
       KID_SCAN_REDUCE, KID_SCAN_SUMS, KID_SCAN_DOWN, KID_GRID_SCATTER, KID_RESOLVE, KID_PEND_RESOLVE,
       KID_REACT_PAIRS, KID_REACT_RESOLVE, KID_FINISH, KID_SERIES, KID_PAIRS_EVAL, KID_SPECIAL, KID_PROPOSE_LIG, KID_PROPOSE_COMPLEX_SMALL, KID_STEP_BEGIN, KID_SMALL_STEP, KID_PROPOSE_COMPLEX_MULTI };

static cudaEvent_t take_event(kmc_handle *h) {
    if (!h->evpool.empty()) { cudaEvent_t e = h->evpool.back(); h->evpool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
static void harvest(kmc_handle *h, bool all) {
    size_t keep = 0;
    for (auto &p : h->pending) {
        if (all ? (cudaEventSynchronize(p.b) == cudaSuccess) : (cudaEventQuery(p.b) == cudaSuccess)) {
            float ms = 0; cudaEventElapsedTime(&ms, p.a, p.b);
            h->kms[p.id] += ms; h->kcount[p.id]++;
            h->evpool.push_back(p.a); h->evpool.push_back(p.b);
        } else h->pending[keep++] = p;
    }
    h->pending.resize(keep);
}
// KMC_TIMELINE=1 (diagnostics): every kernel of the step graph is bracketed by two one-thread stamp kernels that write %globaltimer,
// kmc_timeline_print shows where each kernel of the LAST step ran on the device clock (the stamps add ~2 us per node)
__global__ void k_stamp(unsigned long long *buf, int slot) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); buf[slot] = t; }
// every kernel launch of the sweep goes through this macro: counts it, and brackets it with events when profiling
#define LAUNCH(id, ...)                                                                              \
    do {                                                                                             \
        h->launches++;                                                                               \
        if (h->timeline && h->tlCount < 64) {                                                        \
            cudaStream_t ts_ = h->tlStream ? h->tlStream : st;                                       \
            const int sl_ = h->tlCount++; h->tlId[sl_] = id;                                         \
            k_stamp<<<1, 1, 0, ts_>>>(h->timeline, 2 * sl_); __VA_ARGS__; k_stamp<<<1, 1, 0, ts_>>>(h->timeline, 2 * sl_ + 1); \
        } else if (h->profiling) {                                                                          \
            kmc_handle::Pending p_{id, take_event(h), take_event(h)};                                \
            cudaEventRecord(p_.a, st); __VA_ARGS__; cudaEventRecord(p_.b, st);                       \
            h->pending.push_back(p_);                                                                \
        } else { __VA_ARGS__; }                                                                      \
    } while (0)


#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return KMC_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

template <class T> static cudaError_t dalloc(kmc_handle *h, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) { h->allocs.push_back(q); e = cudaMemset(q, 0, std::max<size_t>(n, 1) * sizeof(T)); }
    *p = (T *)q;
    return e;
}

extern "C" int kmc_abi_version(void) { return KMC_ABI_VERSION; }

extern "C" void kmc_default_params(kmc_params *p) {
    memset(p, 0, sizeof *p);
    p->box[0] = 5773; p->box[1] = 5773; p->box[2] = 1000; p->dt = 10; p->pai = 3.1415926;
    p->rA = 20; p->DA = 1; p->DrotA = 0.0174; p->rB = 30; p->DB = 7.2614; p->DrotB = 0.0061209;
    p->mono_cis_on = 0.000047; p->mono_cis_off = 0.000000000000112;
    p->cis_D = 0.5; p->cis_Drot = 0.005; p->cis_on = 0.00096; p->cis_off = 0.000000000000112;
    p->bond_D = 0.5; p->bond_Drot = 0.005; p->on = 0.04; p->off = 0.000000000000348;
    p->bond_dist_cut = 18; p->thetapd_cut = 45; p->thetaot_cut = 90; p->cis_thetaot_cut = 10; p->cis_dist_cut = 15;
    p->n_receptor = 150; p->n_ligand = 50; p->n_replicas = 1; p->mode = KMC_MODE_REPLAY; p->seed = 1;
    p->cell_edge = 0; p->device = 0;
}

// smallest double T with sqrt(T) >= c, so that for every double d2:  sqrt(d2) < c  <=>  d2 < T  (IEEE sqrt is monotone and
// correctly rounded). Lets the overlap kernels compare squared distances and still decide exactly like main.cpp:646.
static double sq_threshold(double c) {
    double T = c * c;
    while (sqrt(T) >= c) T = nextafter(T, 0.0);
    while (sqrt(T) < c) T = nextafter(T, INFINITY);
    return T;
}

// AreSame(sqrt(q), t), i.e. fabs(sqrt(q) - t) < 1e-8 (main.cpp:2368-2371), as a window on q: sqrt is correctly rounded and monotone
// and the rounded difference is monotone in its first operand, so the q that pass form an interval; its two ends are found by
// stepping through the neighbouring doubles with the predicate itself. w[0] > w[1] (not available) if the search does not settle.
static void same_window(double t, double w[2]) {
    auto ok = [&](double q) { return fabs(sqrt(q) - t) < 1.0E-8; };
    w[0] = 1; w[1] = 0;
    if (!(t > 1.0E-6) || !ok(t * t)) return;
    double lo = (t - 1.0E-8) * (t - 1.0E-8), hi = (t + 1.0E-8) * (t + 1.0E-8);
    int guard = 0;
    while (ok(lo) && ++guard < 100000) lo = nextafter(lo, 0.0);
    while (!ok(lo) && ++guard < 200000) lo = nextafter(lo, INFINITY);
    while (ok(hi) && ++guard < 300000) hi = nextafter(hi, INFINITY);
    while (!ok(hi) && ++guard < 400000) hi = nextafter(hi, 0.0);
    if (guard >= 100000 || !ok(lo) || !ok(hi) || ok(nextafter(lo, 0.0)) || ok(nextafter(hi, INFINITY)) || !(lo <= t * t && t * t <= hi)) return;
    w[0] = lo; w[1] = hi;
}

// derived constants with the reference's own expressions (host doubles, no contraction: see build flags)
static void fill_consts(const kmc_params &P, Consts &K) {
    memset(&K, 0, sizeof K);
    K.Lx = P.box[0]; K.Ly = P.box[1]; K.Lz = P.box[2]; K.dt = P.dt; K.pai = P.pai; K.rA = P.rA; K.rB = P.rB;
    K.ampA = 2 * sqrt(P.DA * P.dt / 6); K.ampB = 2 * sqrt(P.DB * P.dt / 6);
    K.ampCis = 2 * sqrt(P.cis_D * P.dt / 6); K.ampBond = 2 * sqrt(P.bond_D * P.dt / 6);
    K.rotA = sqrt(P.DrotA * P.dt); K.rotB = sqrt(P.DrotB * P.dt); K.rotCis = sqrt(P.cis_Drot * P.dt); K.rotBond = sqrt(P.bond_Drot * P.dt);
    K.pOn = P.on * P.dt; K.pMonoCisOn = P.mono_cis_on * P.dt; K.pCisOn = P.cis_on * P.dt;
    K.pOff = P.off * P.dt; K.pMonoCisOff = P.mono_cis_off * P.dt; K.pCisOff = P.cis_off * P.dt;
    K.bondCut = P.bond_dist_cut; K.thetaPdCut = P.thetapd_cut; K.thetaOtCut = P.thetaot_cut;
    K.cisThetaCut = P.cis_thetaot_cut; K.cisCut = P.cis_dist_cut;
    K.ovAA = P.rA + P.rA; K.ovAB = P.rA + P.rB; K.ovBB = P.rB + P.rB;
    K.ovAA2 = sq_threshold(K.ovAA); K.ovAB2 = sq_threshold(K.ovAB); K.ovBB2 = sq_threshold(K.ovBB);
    K.rlD1 = P.bond_dist_cut / 2 + P.rA + P.rB; K.rlD2 = P.bond_dist_cut / 2;
    K.cisD1 = P.cis_dist_cut / 2 + P.rA + P.rA; K.cisD2 = P.cis_dist_cut / 2;
    same_window(K.rlD1, K.wRL1); same_window(K.rlD2, K.wRL2); same_window(K.cisD1, K.wCis1); same_window(K.cisD2, K.wCis2);
    K.fRL1 = (P.bond_dist_cut / 2 + P.rA) / P.rB; K.fRL3 = (P.bond_dist_cut / 2 + 2 * P.rA) / P.rB; K.fRL2 = (P.bond_dist_cut / 2) / P.rB;
    K.fC1 = (P.cis_dist_cut / 2 + P.rA) / P.rA; K.fC3 = (P.cis_dist_cut / 2) / P.rA; K.fC2 = (P.cis_dist_cut / 2 + 2 * P.rA) / P.rA;
    K.fSeat = (P.bond_dist_cut / 2 + P.rB * 2 / sqrt(3) + P.rB) / P.rA;
    const double rB = P.rB;
    const double g[8][2] = {{0, 0}, {0, rB * 2 / sqrt(3)}, {-rB, -rB / sqrt(3)}, {rB, -rB / sqrt(3)}, {0, 0},
                            {0, rB * (2 / sqrt(3) + 1)}, {-rB * (sqrt(3) / 2 + 1), -rB / sqrt(3) - rB / 2},
                            {rB * (sqrt(3) / 2 + 1), -rB / sqrt(3) - rB / 2}};
    memcpy(K.ghost, g, sizeof g);
    K.conv = 180 / 3.14159;
    const double rs = rB * 2 / sqrt(3.0);
    K.reachRR = 2 * P.rA + 1e-3; K.reachRL = P.rA + P.rB + rs + 1e-3; K.reachLL = 2 * P.rB + 2 * rs + 1e-3;
    K.reachOn = P.rA + P.bond_dist_cut + rs + P.rB + 1e-3; K.reachCis = 2 * P.rA + P.cis_dist_cut + 1e-3;
    // list reuse: the pair list of a build step serves the following (reuse - 1) steps; the cut and the cells grow by the
    // drift a molecule may accumulate meanwhile (free units move at most max(amp) per step; whatever moves further is handled
    // as a special entry). KMC_REUSE / KMC_SKIN / KMC_DRIFT override (tuning knobs, every setting is exact).
    int reuse = 6;
    if (const char *o = getenv("KMC_REUSE")) reuse = std::max(1, atoi(o));
    K.skin = reuse > 1 ? 12.0 : 24.0;            // far-mover threshold
    if (const char *sk = getenv("KMC_SKIN")) { double v = atof(sk); if (v > 0) K.skin = v; }
    K.drift = reuse > 1 ? (reuse - 1) * std::max({K.ampA, K.ampB, K.ampCis, K.ampBond}) + 0.25 : 0.0;
    if (const char *o = getenv("KMC_DRIFT")) { double v = atof(o); if (v >= 0 && reuse > 1) K.drift = v; }
    K.phase = 0;
    K.NA = P.n_receptor; K.NB = P.n_ligand; K.R = P.n_replicas; K.mode = P.mode;
    K.NAt = K.NA * K.R; K.NBt = K.NB * K.R; K.NT = K.NAt + K.NBt; K.seed = P.seed;
    K.strips = 1; K.stripRank = 0; K.stripXc = 0; K.stripHalf = INFINITY;
    double edge = std::max({K.reachLL, K.reachOn, K.reachCis}) + 2 * K.skin + 2 * K.drift + 1.0;   // walk around the entry: reach + 2 skins (+ 2 drifts)
    if (P.cell_edge > edge) edge = P.cell_edge;
    else if (P.cell_edge == 0) edge = std::max(edge, getenv("KMC_EDGE") ? atof(getenv("KMC_EDGE")) : 256.0);
    K.gx0 = -P.box[0] / 2 - edge; K.gy0 = -P.box[1] / 2 - edge;
    K.ncx = (int)ceil((P.box[0] + 2 * edge) / edge); K.ncy = (int)ceil((P.box[1] + 2 * edge) / edge);
    K.cellInv = 1.0 / edge;
    K.keyX0 = K.gx0; K.keyY0 = K.gy0; K.keyInv = K.cellInv;
    const float maxc = (float)(std::max(P.box[0], P.box[1]) + 2 * edge);
    K.cutMargin = 0.05f + 2 * (nextafterf(maxc, INFINITY) - maxc);
}

static void strip_dev_free(kmc_handle *h);
// tile edge of k_resolve_tiles: the largest of 16/12/8/6/4/2 cells whose window ((edge+2)^2 cells) is expected to hold at most
// ~300 entries (the staging buffers take TCAP = 384; a fuller window still works through the slower overflow path)
static void choose_tiles(kmc_handle *h) {
    Consts &K = h->K;
    const double perCell = (double)K.NT / std::max(1.0, (double)K.R * K.ncx * K.ncy);
    const int cand[6] = {16, 12, 8, 6, 4, 2};
    int ts = 2;
    for (int c : cand) if ((c + 2.0) * (c + 2.0) * perCell <= 300.0) { ts = c; break; }
    if (const char *o = getenv("KMC_TILE_EDGE")) { int v = atoi(o); if (v >= 1 && v <= TS) ts = v; }
    K.tileEdge = std::min(ts, TS);
    h->nTiles = K.R * ((K.ncx + K.tileEdge - 1) / K.tileEdge) * ((K.ncy + K.tileEdge - 1) / K.tileEdge);
    // sparse cells (the reference's own density: ~0.4 molecules per cell): no staging, one thread per grid entry
    h->useCells = perCell <= 2.0;
    if (const char *o = getenv("KMC_RESOLVE")) h->useCells = !strcmp(o, "cells");
    h->listEvery = 1;
    if (!h->useCells) { K.drift = 0; if (!getenv("KMC_SKIN")) K.skin = 24.0; }      // tile path: rebuilt every step (the cells are at least as large as this needs)
    if (h->useCells && K.drift > 0) { h->listEvery = 6; if (const char *o = getenv("KMC_REUSE")) h->listEvery = std::max(1, atoi(o)); }
    h->sinceBuild = 0;
    h->adapt = !(getenv("KMC_ADAPT") && atoi(getenv("KMC_ADAPT")) == 0);
}
// arrays of the sparse path (allocated on first need: strips can re-derive the grid and with it the choice of path)
static bool ensure_cells_arrays(kmc_handle *h) {
    Dev &D = h->D; const Consts &K = h->K;
    if (!h->useCells) return true;
    bool ok = true;
    if (!D.scen) {
        D.survCap = 8 * K.NT + 4096;
        ok = dalloc(h, &D.scen, (size_t)2 * K.NT) == cudaSuccess && dalloc(h, &D.scell, (size_t)2 * K.NT) == cudaSuccess && dalloc(h, &D.surv, (size_t)D.survCap) == cudaSuccess && dalloc(h, &D.reactList, (size_t)2 * D.survCap) == cudaSuccess &&
             dalloc(h, &D.bcen, (size_t)K.NT) == cudaSuccess && dalloc(h, &D.specList, (size_t)2 * K.NT) == cudaSuccess && dalloc(h, &D.specNext, (size_t)2 * K.NT) == cudaSuccess;
    }
    if (ok && h->cellHeadCap < D.ncell) { ok = dalloc(h, &D.cellHead, (size_t)D.ncell) == cudaSuccess; h->cellHeadCap = D.ncell; }
    return ok;
}

extern "C" void kmc_destroy(kmc_handle *h) {
    if (!h) return;
    strip_dev_free(h);
    cudaSetDevice(h->P.device);
    for (void *p : h->allocs) cudaFree(p);
    for (auto &p : h->pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    for (auto ev : h->evpool) cudaEventDestroy(ev);
    for (int p = 0; p < 4; p++) if (h->gexec[p >> 1][p & 1]) cudaGraphExecDestroy(h->gexec[p >> 1][p & 1]);
    if (h->copyStream) { cudaStreamSynchronize(h->copyStream); cudaStreamDestroy(h->copyStream); }
    if (h->evSnap) cudaEventDestroy(h->evSnap);
    if (h->evSnapDone) cudaEventDestroy(h->evSnapDone);
    if (h->monHost) cudaFreeHost(h->monHost);
    if (h->monEvent) cudaEventDestroy(h->monEvent);
    for (auto e : h->evFork) if (e) cudaEventDestroy(e);
    for (auto e : h->evJoin) if (e) cudaEventDestroy(e);
    for (auto q : h->side) if (q) cudaStreamDestroy(q);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" const char *kmc_last_error(const kmc_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int kmc_create(const kmc_params *p, kmc_handle **out) {
    if (!p || !out) { g_create_error = "null argument"; return KMC_ERR_INVALID; }
    *out = nullptr;
    if ((p->mode != KMC_MODE_REPLAY && p->mode != KMC_MODE_PRODUCTION) || p->n_receptor < 0 || p->n_ligand < 1 || p->n_replicas < 1 || p->box[0] <= 0 || p->box[1] <= 0 || p->box[2] <= 0 ||
        p->rA <= 0 || p->rB <= 0 || p->dt <= 0) { g_create_error = "invalid parameters"; return KMC_ERR_INVALID; }
    if ((int64_t)(p->n_receptor + p->n_ligand) * p->n_replicas >= (1LL << 30)) { g_create_error = "too many molecules"; return KMC_ERR_INVALID; }
    kmc_handle *h = new kmc_handle;
    h->P = *p;
    auto fail = [&](int code, const std::string &m) { g_create_error = m; kmc_destroy(h); return code; };
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(KMC_ERR_CUDA, std::string("no CUDA device: this library has no CPU path (") + cudaGetErrorString(e) + ")");
    if (p->device < 0 || p->device >= ndev) return fail(KMC_ERR_INVALID, "bad device ordinal");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess || cudaSetDevice(p->device) != cudaSuccess) return fail(KMC_ERR_CUDA, "cannot select device");
    if (prop.major != 10) return fail(KMC_ERR_CUDA, "device is not compute capability 10.x: the kernels are built for sm_100a only");
    if (p->min_image != 0) return fail(KMC_ERR_INVALID, "min_image = 1 is not implemented: the reference computes plain Euclidean distances (main.cpp:642-646), which is what 0 selects");
    h->nSM = std::max(prop.multiProcessorCount, 1);
    if (const char *o = getenv("KMC_FORK")) h->forkMask = atoi(o);
    if (const char *o = getenv("KMC_CX_GROUPS")) h->cxGroups = atoi(o) != 0;
    if (getenv("KMC_TIMELINE")) { void *q = nullptr; if (cudaMalloc(&q, 128 * sizeof(unsigned long long)) == cudaSuccess) { h->allocs.push_back(q); h->timeline = (unsigned long long *)q; } }
    fill_consts(*p, h->K);
    const Consts &K = h->K;
    h->R = K.R; h->NA = K.NA; h->NB = K.NB; h->N = K.NA + K.NB; h->NAt = K.NAt; h->NBt = K.NBt; h->NT = K.NT;
    if ((int64_t)K.R * K.ncx * K.ncy >= (1LL << 31) - 2) return fail(KMC_ERR_INVALID, "neighbour grid too large: raise cell_edge");
    Dev &D = h->D; memset(&D, 0, sizeof D);
    D.ncell = K.R * K.ncx * K.ncy; D.nAcap = K.NAt;
    // systems that fit a CTA's shared-memory records take the fused step (KMC_FUSED=0, or an explicit KMC_RESOLVE, keeps the general path)
    h->fused = K.NA + K.NB <= SMALL_MAXN && K.NB <= SMALL_MAXNB && std::max(p->box[0], p->box[1]) <= 2.0e5 &&
               !(getenv("KMC_FUSED") && atoi(getenv("KMC_FUSED")) == 0) && !getenv("KMC_RESOLVE");
    if (h->fused) {          // (poses, bond table, lists and views of a replica live in the shared memory of its CTA)
        const size_t per = sizeof(SmallShared) + small_dyn_bytes(K.NA, K.NB);
        int occ1 = 0, occ4 = 0;
        bool ok1 = per <= (size_t)prop.sharedMemPerBlockOptin && cudaFuncSetAttribute(k_small_step<1, SMALL_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per) == cudaSuccess &&
                   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_small_step<1, SMALL_T>, SMALL_T, per) == cudaSuccess && occ1 >= 1;
        bool ok4 = ok1 && 4 * per + 64 <= (size_t)prop.sharedMemPerBlockOptin && cudaFuncSetAttribute(k_small_step<4, SMALL_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * per)) == cudaSuccess &&
                   cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, k_small_step<4, SMALL_T>, 4 * SMALL_T, 4 * per) == cudaSuccess && occ4 >= 1;
        cudaGetLastError();
        if (!ok1) h->fused = false;
        // at most one replica per SM: a wide CTA (a molecule per thread, registers uncapped)
        h->smallWide = ok1 && K.R <= h->nSM && cudaFuncSetAttribute(k_small_step<1, 2 * SMALL_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per) == cudaSuccess &&
                       !(getenv("KMC_SMALL_WIDE") && atoi(getenv("KMC_SMALL_WIDE")) == 0);
        cudaGetLastError();
        h->smallGrid = occ1 * h->nSM;
        // ensembles that fill the device: four replicas per CTA in lockstep (one CTA per SM), see k_small_step
        h->smallSlots = (ok4 && 2 * K.R >= 7 * occ4 * h->nSM) ? 4 : 1;          // (measured: 444 replicas 19.1 vs 19.9 us per step, 560 replicas 21.7 vs 20.0)
        if (const char *o = getenv("KMC_SMALL_SLOTS")) { const int v = atoi(o); if (v == 1 || (v == 4 && ok4)) h->smallSlots = v; }
        if (h->smallSlots == 4) h->smallGrid = occ4 * h->nSM;
        if (const char *o = getenv("KMC_SMALL_GRID")) h->smallGrid = std::max(1, atoi(o));          // (tests: force the ticket path on a small ensemble)
        if (h->fused && dalloc(h, &h->smallQueue, (size_t)K.R + 1) != cudaSuccess) return fail(KMC_ERR_CUDA, "device allocation failed");
    }
    D.candCap = std::max({1 << 14, K.NAt / 8, h->fused ? 64 * K.R : 0});
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    // the side branches of the step graph carry the latency-bound kernels (complexes, special entries): few CTAs with long serial
    // chains. Highest priority, so that their CTAs are dispatched ahead of the streaming kernels' remaining CTAs and run underneath them
    int prLo = 0, prHi = 0; cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    for (auto &q : h->side) ok = ok && cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, getenv("KMC_NOPRIO") ? prLo : prHi) == cudaSuccess;
    for (auto &e : h->evFork) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    for (auto &e : h->evJoin) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->monEvent, cudaEventDisableTiming) == cudaSuccess && cudaMallocHost((void **)&h->monHost, sizeof(int) * S_COUNT) == cudaSuccess;
#define A(ptr, n) ok = ok && dalloc(h, &D.ptr, (size_t)(n)) == cudaSuccess
    A(recC, K.NAt); A(recS2, K.NAt); A(recS3, K.NAt); A(recCn, K.NAt); A(recS2n, K.NAt); A(recS3n, K.NAt);
    A(lig, (size_t)K.NBt * 24); A(lign, (size_t)K.NBt * 24);
    A(recLig, K.NAt); A(recSite, K.NAt); A(recCis, K.NAt); A(ligRec, (size_t)K.NBt * 3);
    A(ufParent, K.NT); A(unitOf, K.NT);
    if (K.mode == KMC_MODE_PRODUCTION) { A(ukey, K.NT); } else D.ukey = D.unitOf; A(cxSize, K.NBt); A(cxOff, K.NBt); A(cxRoots, (size_t)2 * K.NBt);
    A(members, K.NT); A(rowWork, K.NT); A(bfsMark, K.NT); A(rowPos, K.NT);
    A(cxStamp, K.NT); A(rootSlot, K.NBt); A(touchList, TOUCH_CAP); A(bfsQueue, K.NT);
    A(movedFlag, K.NT); A(nrec, (size_t)6 * K.NT);
    h->scanBlocks = (D.ncell + 1 + SCAN_TILE - 1) / SCAN_TILE;
    A(cellCount, (size_t)h->scanBlocks * SCAN_TILE); A(cellStart, (size_t)h->scanBlocks * SCAN_TILE);
    choose_tiles(h);
    A(scanTmp, (size_t)h->scanBlocks + 1);
    A(sorted, (size_t)2 * K.NT); A(molSlot, K.NT);
    ok = ok && ensure_cells_arrays(h); A(farList, K.NT);
    A(candRL, (size_t)2 * D.candCap); A(candCis, (size_t)2 * D.candCap);
    D.pairCap = std::max({1 << 16, 4 * K.NAt, h->fused ? 8 * K.NAt + 64 * K.R : 0});      // (the fused step deals every replica an equal slice of the lists)
    A(pairs, D.pairCap); A(unitRes, K.NT); A(pendCnt, K.NT); A(rejList, K.NT); A(rejPartner, K.NT); D.pendCap = 2 * K.NT + 4096;
    if (const char *o = getenv("KMC_TEST_PENDCAP")) D.pendCap = std::max(1, atoi(o));      // tests: force the overflow report
    A(pendList, D.pendCap); A(step64, 1);
    A(scal, S_COUNT); A(maxComplex, K.R); A(events, EV_COUNT);
#undef A
    ok = ok && dalloc(h, &h->d_series, (size_t)K.R * 6) == cudaSuccess;
    if (!ok) return fail(KMC_ERR_CUDA, std::string("device allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    // empty bond table
    ok = cudaMemset(D.recLig, 0xff, sizeof(int) * K.NAt) == cudaSuccess && cudaMemset(D.recSite, 0xff, sizeof(int) * K.NAt) == cudaSuccess &&
         cudaMemset(D.recCis, 0xff, sizeof(int) * K.NAt) == cudaSuccess && cudaMemset(D.ligRec, 0xff, sizeof(int) * 3 * (size_t)K.NBt) == cudaSuccess;
    int one = 1;
    ok = ok && cudaMemset(D.rootSlot, 0xff, sizeof(int) * (size_t)K.NBt) == cudaSuccess;
    ok = ok && cudaMemcpy(D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(D.scal + S_NA_LIVE, &K.NAt, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    ok = ok && cudaMemcpy(D.scal + S_NB_LIVE, &K.NBt, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) return fail(KMC_ERR_CUDA, std::string("device initialisation failed: ") + cudaGetErrorString(cudaGetLastError()));
    *out = h;
    return KMC_OK;
}

// ------------------------------------------------------------------------------------------------
// state exchange
// ------------------------------------------------------------------------------------------------
static inline size_t RI(int i, int j, int k) { return (size_t)i * 25 + j * 5 + k; }

static int select_device(kmc_handle *h) { CK(cudaSetDevice(h->P.device)); return KMC_OK; }

extern "C" int kmc_set_state(kmc_handle *h, int32_t rep, const double *Rx, const double *Ry, const double *Rz,
                             const int32_t *status, const int32_t *res_nei, int64_t step_done, int32_t max_complex) {
    if (!h) return KMC_ERR_INVALID;
    if (rep < 0 || rep >= h->R || !Rx || !Ry || !Rz || !status || !res_nei) { h->err = "kmc_set_state: bad argument"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    const int NA = h->NA, NB = h->NB, N = h->N;
    const double rA = h->P.rA;
    std::vector<double2> c(NA), s2(NA), s3(NA);
    std::vector<double> lig((size_t)NB * 24);
    std::vector<int> rl(NA), rs(NA), rc_(NA), lr((size_t)NB * 3);
    for (int a = 0; a < NA; a++) {
        const int i = a + 1;
        for (int j = 1; j <= 4; j++) {
            bool same = Rx[RI(i, j, 1)] == Rx[RI(i, 1, 1)] && Ry[RI(i, j, 1)] == Ry[RI(i, 1, 1)] && Rx[RI(i, j, 2)] == Rx[RI(i, 1, 2)] &&
                        Ry[RI(i, j, 2)] == Ry[RI(i, 1, 2)] && Rx[RI(i, j, 3)] == Rx[RI(i, 1, 3)] && Ry[RI(i, j, 3)] == Ry[RI(i, 1, 3)] &&
                        Rx[RI(i, j, 4)] == Rx[RI(i, 1, 1)] && Ry[RI(i, j, 4)] == Ry[RI(i, 1, 1)];
            bool zt = Rz[RI(i, j, 1)] == (j * 2 - 2) * rA && Rz[RI(i, j, 2)] == (j * 2 - 2) * rA && Rz[RI(i, j, 3)] == (j * 2 - 2) * rA &&
                      Rz[RI(i, j, 4)] == (j * 2 - 1) * rA;
            if (!same || !zt) {
                h->err = "kmc_set_state: receptor " + std::to_string(i) + " is not a stack of identical beads at z=(2j-2)*rA (main.cpp:298-316)";
                return KMC_ERR_STATE;
            }
        }
        c[a] = make_double2(Rx[RI(i, 1, 1)], Ry[RI(i, 1, 1)]);
        s2[a] = make_double2(Rx[RI(i, 1, 2)], Ry[RI(i, 1, 2)]);
        s3[a] = make_double2(Rx[RI(i, 1, 3)], Ry[RI(i, 1, 3)]);
        int l = res_nei[i * 7 + 2], st = res_nei[i * 7 + 4], cp = res_nei[i * 7 + 3];
        if ((l != 0) != (status[i * 5 + 2] != 0) || (cp != 0) != (status[i * 5 + 3] != 0) || (l != 0 && (l <= NA || l > N || st < 2 || st > 4)) ||
            (cp != 0 && (cp < 1 || cp > NA || cp == i))) { h->err = "kmc_set_state: bad bond entry for receptor " + std::to_string(i); return KMC_ERR_STATE; }
        if (l != 0 && res_nei[l * 7 + st] != i) { h->err = "kmc_set_state: asymmetric R-L bond at receptor " + std::to_string(i); return KMC_ERR_STATE; }
        if (cp != 0 && res_nei[cp * 7 + 3] != i) { h->err = "kmc_set_state: asymmetric cis bond at receptor " + std::to_string(i); return KMC_ERR_STATE; }
        rl[a] = l ? rep * NB + (l - NA - 1) : -1; rs[a] = l ? st - 2 : -1; rc_[a] = cp ? rep * NA + (cp - 1) : -1;
    }
    for (int b = 0; b < NB; b++) {
        const int i = NA + 1 + b;
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 2; k++) {
                int q = k == 1 ? (j == 1 ? 0 : j - 1) : (j == 1 ? 4 : j + 3);
                lig[(size_t)b * 24 + q * 3 + 0] = Rx[RI(i, j, k)]; lig[(size_t)b * 24 + q * 3 + 1] = Ry[RI(i, j, k)]; lig[(size_t)b * 24 + q * 3 + 2] = Rz[RI(i, j, k)];
            }
        for (int s = 0; s < 3; s++) {
            int r = res_nei[i * 7 + s + 2];
            if ((r != 0) != (status[i * 5 + s + 2] != 0) || (r != 0 && (r < 1 || r > NA || res_nei[r * 7 + 2] != i || res_nei[r * 7 + 4] != s + 2))) {
                h->err = "kmc_set_state: bad bond entry for ligand " + std::to_string(i); return KMC_ERR_STATE;
            }
            lr[(size_t)b * 3 + s] = r ? rep * NA + (r - 1) : -1;
        }
    }
    Dev &D = h->D;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(D.recC + (size_t)rep * NA, c.data(), sizeof(double2) * NA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.recS2 + (size_t)rep * NA, s2.data(), sizeof(double2) * NA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.recS3 + (size_t)rep * NA, s3.data(), sizeof(double2) * NA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.lig + (size_t)rep * NB * 24, lig.data(), sizeof(double) * 24 * NB, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.recLig + (size_t)rep * NA, rl.data(), sizeof(int) * NA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.recSite + (size_t)rep * NA, rs.data(), sizeof(int) * NA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.recCis + (size_t)rep * NA, rc_.data(), sizeof(int) * NA, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.ligRec + (size_t)rep * NB * 3, lr.data(), sizeof(int) * 3 * NB, cudaMemcpyHostToDevice));
    int one = 1;
    CK(cudaMemcpy(D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.maxComplex + rep, &max_complex, sizeof(int), cudaMemcpyHostToDevice));
    h->step_done = step_done; h->stepped = false; h->sinceBuild = 0;
    { unsigned long long s64 = (unsigned long long)step_done; CK(cudaMemcpy(D.step64, &s64, sizeof s64, cudaMemcpyHostToDevice)); }
    return KMC_OK;
}

extern "C" int kmc_get_state(kmc_handle *h, int32_t rep, double *Rx, double *Ry, double *Rz, int32_t *status, int32_t *res_nei) {
    if (!h) return KMC_ERR_INVALID;
    if (rep < 0 || rep >= h->R || !Rx || !Ry || !Rz || !status || !res_nei) { h->err = "kmc_get_state: bad argument"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    const int NA = h->NA, NB = h->NB, N = h->N;
    const double rA = h->P.rA;
    std::vector<double2> c(NA), s2(NA), s3(NA);
    std::vector<double> lig((size_t)NB * 24);
    std::vector<int> rl(NA), rs(NA), rcis(NA), lr((size_t)NB * 3);
    Dev &D = h->D;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(c.data(), D.recC + (size_t)rep * NA, sizeof(double2) * NA, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(s2.data(), D.recS2 + (size_t)rep * NA, sizeof(double2) * NA, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(s3.data(), D.recS3 + (size_t)rep * NA, sizeof(double2) * NA, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lig.data(), D.lig + (size_t)rep * NB * 24, sizeof(double) * 24 * NB, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rl.data(), D.recLig + (size_t)rep * NA, sizeof(int) * NA, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rs.data(), D.recSite + (size_t)rep * NA, sizeof(int) * NA, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rcis.data(), D.recCis + (size_t)rep * NA, sizeof(int) * NA, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lr.data(), D.ligRec + (size_t)rep * NB * 3, sizeof(int) * 3 * NB, cudaMemcpyDeviceToHost));
    memset(Rx, 0, sizeof(double) * 25 * (N + 1)); memset(Ry, 0, sizeof(double) * 25 * (N + 1)); memset(Rz, 0, sizeof(double) * 25 * (N + 1));
    memset(status, 0, sizeof(int32_t) * 5 * (N + 1)); memset(res_nei, 0, sizeof(int32_t) * 7 * (N + 1));
    for (int a = 0; a < NA; a++) {
        const int i = a + 1;
        for (int j = 1; j <= 4; j++) {
            double zj = (j * 2 - 2) * rA;
            Rx[RI(i, j, 1)] = c[a].x; Ry[RI(i, j, 1)] = c[a].y; Rz[RI(i, j, 1)] = zj;
            Rx[RI(i, j, 2)] = s2[a].x; Ry[RI(i, j, 2)] = s2[a].y; Rz[RI(i, j, 2)] = zj;
            Rx[RI(i, j, 3)] = s3[a].x; Ry[RI(i, j, 3)] = s3[a].y; Rz[RI(i, j, 3)] = zj;
            Rx[RI(i, j, 4)] = c[a].x; Ry[RI(i, j, 4)] = c[a].y; Rz[RI(i, j, 4)] = (j * 2 - 1) * rA;
        }
        if (rl[a] >= 0) { status[i * 5 + 2] = 1; res_nei[i * 7 + 2] = NA + 1 + (rl[a] - rep * NB); res_nei[i * 7 + 4] = rs[a] + 2; }
        if (rcis[a] >= 0) { status[i * 5 + 3] = 1; res_nei[i * 7 + 3] = 1 + (rcis[a] - rep * NA); }
    }
    for (int b = 0; b < NB; b++) {
        const int i = NA + 1 + b;
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 2; k++) {
                int q = k == 1 ? (j == 1 ? 0 : j - 1) : (j == 1 ? 4 : j + 3);
                Rx[RI(i, j, k)] = lig[(size_t)b * 24 + q * 3 + 0]; Ry[RI(i, j, k)] = lig[(size_t)b * 24 + q * 3 + 1]; Rz[RI(i, j, k)] = lig[(size_t)b * 24 + q * 3 + 2];
            }
        for (int s = 0; s < 3; s++)
            if (lr[(size_t)b * 3 + s] >= 0) { status[i * 5 + s + 2] = 1; res_nei[i * 7 + s + 2] = 1 + (lr[(size_t)b * 3 + s] - rep * NA); }
    }
    return KMC_OK;
}

static inline int nblk(int n, int b) { return (n + b - 1) / b; }

// ---- compact state exchange: the caller's buffers cross PCIe as they are (one copy per array, full speed from pinned memory);
// (de)interleaving, bond-table construction and validation run on the device ----
__global__ void k_pack_get(const __grid_constant__ Args A, double *recPose, int *site) {
    KARGS
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= cK.NAt) return;
    if (recPose) {
        const double2 c = D.recC[a], s2 = D.recS2[a], s3 = D.recS3[a];
        double *o = recPose + (size_t)a * 6;
        o[0] = c.x; o[1] = c.y; o[2] = s2.x; o[3] = s2.y; o[4] = s3.x; o[5] = s3.y;
    }
    if (site) { const int s = D.recSite[a]; site[a] = s >= 0 ? s + 2 : 0; }
}
// recPose: device copy of the caller's [n][6]; rl/rs/rc: device copies of the caller's bond words (or null = no bonds);
// builds recC/S2/S3, recLig/recSite/recCis and ligRec (pre-set to -1); flags[0] |= 1 R-L bond invalid, |= 2 cis bond invalid
__global__ void k_pack_set(const __grid_constant__ Args A, const double *recPose, const int *rl, const int *rs, const int *rc, int *flags) {
    KARGS
    const Consts &K = cK;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= K.NAt) return;
    const double *o = recPose + (size_t)a * 6;
    D.recC[a] = make_double2(o[0], o[1]); D.recS2[a] = make_double2(o[2], o[3]); D.recS3[a] = make_double2(o[4], o[5]);
    int l = rl ? rl[a] : -1, st = rs ? rs[a] : 0, cp = rc ? rc[a] : -1;
    if (l >= 0) {
        if (l >= K.NBt || l / K.NB != a / K.NA || st < 2 || st > 4 || atomicCAS(&D.ligRec[l * 3 + st - 2], -1, a) != -1) { atomicOr(flags, 1); l = -1; }
    } else l = -1;
    if (cp >= 0) { if (cp >= K.NAt || cp == a || cp / K.NA != a / K.NA || rc[cp] != a) { atomicOr(flags, 2); cp = -1; } } else cp = -1;
    D.recLig[a] = l; D.recSite[a] = l >= 0 ? st - 2 : -1; D.recCis[a] = cp;
}

static int stage_alloc(kmc_handle *h) {
    if (h->stageRec) return KMC_OK;
    bool ok = dalloc(h, &h->stageRec, (size_t)std::max(h->NAt, 1) * 6) == cudaSuccess && dalloc(h, &h->stageInt, (size_t)std::max(h->NAt, 1) * 3 + 4) == cudaSuccess;
    if (!ok) { h->err = "staging allocation failed"; return KMC_ERR_CUDA; }
    return KMC_OK;
}

extern "C" int kmc_get_packed(kmc_handle *h, double *rec_pose, double *lig_pose, int32_t *rec_lig, int32_t *rec_site, int32_t *rec_cis) {
    if (!h) return KMC_ERR_INVALID;
    int rc = select_device(h); if (rc) return rc;
    rc = stage_alloc(h); if (rc) return rc;
    Dev &D = h->D; const int NAt = h->NAt, NBt = h->NBt; cudaStream_t st = h->stream;
    const Args A{D, h->K};
    if (rec_pose || rec_site) {
        LAUNCH(KID_SERIES, (k_pack_get<<<nblk(std::max(NAt, 1), 256), 256, 0, st>>>(A, rec_pose ? h->stageRec : nullptr, rec_site ? h->stageInt : nullptr)));
        if (rec_pose) CK(cudaMemcpyAsync(rec_pose, h->stageRec, sizeof(double) * 6 * (size_t)NAt, cudaMemcpyDeviceToHost, st));
        if (rec_site) CK(cudaMemcpyAsync(rec_site, h->stageInt, sizeof(int) * NAt, cudaMemcpyDeviceToHost, st));
    }
    if (lig_pose) CK(cudaMemcpyAsync(lig_pose, D.lig, sizeof(double) * 24 * (size_t)NBt, cudaMemcpyDeviceToHost, st));
    if (rec_lig) CK(cudaMemcpyAsync(rec_lig, D.recLig, sizeof(int) * NAt, cudaMemcpyDeviceToHost, st));
    if (rec_cis) CK(cudaMemcpyAsync(rec_cis, D.recCis, sizeof(int) * NAt, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return KMC_OK;
}

// The committed state is copied into a device-side snapshot on the handle's stream (a few tens of microseconds; stepping may go
// on at once) and crosses PCIe from there on a copy stream of its own: the result of one batch travels while the next one computes.
extern "C" int kmc_get_packed_async(kmc_handle *h, double *rec_pose, double *lig_pose, int32_t *rec_lig, int32_t *rec_site, int32_t *rec_cis) {
    if (!h) return KMC_ERR_INVALID;
    int rc = select_device(h); if (rc) return rc;
    Dev &D = h->D; const int NAt = h->NAt, NBt = h->NBt; cudaStream_t st = h->stream;
    if (!h->copyStream) {
        bool ok = cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreateWithFlags(&h->evSnap, cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&h->evSnapDone, cudaEventDisableTiming) == cudaSuccess && dalloc(h, &h->snapRec, (size_t)std::max(NAt, 1) * 6) == cudaSuccess &&
                  dalloc(h, &h->snapLig, (size_t)std::max(NBt, 1) * 24) == cudaSuccess && dalloc(h, &h->snapInt, (size_t)std::max(NAt, 1) * 3) == cudaSuccess;
        if (!ok) { h->err = "kmc_get_packed_async: cannot set up the snapshot buffers"; return KMC_ERR_CUDA; }
    }
    if (h->snapPending) CK(cudaStreamWaitEvent(st, h->evSnapDone, 0));          // (the previous snapshot must have left the device buffers)
    const Args A{D, h->K};
    LAUNCH(KID_SERIES, (k_pack_get<<<nblk(std::max(NAt, 1), 256), 256, 0, st>>>(A, rec_pose ? h->snapRec : nullptr, rec_site ? h->snapInt : nullptr)));
    if (lig_pose) CK(cudaMemcpyAsync(h->snapLig, D.lig, sizeof(double) * 24 * (size_t)NBt, cudaMemcpyDeviceToDevice, st));
    if (rec_lig) CK(cudaMemcpyAsync(h->snapInt + NAt, D.recLig, sizeof(int) * NAt, cudaMemcpyDeviceToDevice, st));
    if (rec_cis) CK(cudaMemcpyAsync(h->snapInt + 2 * (size_t)NAt, D.recCis, sizeof(int) * NAt, cudaMemcpyDeviceToDevice, st));
    CK(cudaEventRecord(h->evSnap, st));
    CK(cudaStreamWaitEvent(h->copyStream, h->evSnap, 0));
    if (rec_pose) CK(cudaMemcpyAsync(rec_pose, h->snapRec, sizeof(double) * 6 * (size_t)NAt, cudaMemcpyDeviceToHost, h->copyStream));
    if (lig_pose) CK(cudaMemcpyAsync(lig_pose, h->snapLig, sizeof(double) * 24 * (size_t)NBt, cudaMemcpyDeviceToHost, h->copyStream));
    if (rec_site) CK(cudaMemcpyAsync(rec_site, h->snapInt, sizeof(int) * NAt, cudaMemcpyDeviceToHost, h->copyStream));
    if (rec_lig) CK(cudaMemcpyAsync(rec_lig, h->snapInt + NAt, sizeof(int) * NAt, cudaMemcpyDeviceToHost, h->copyStream));
    if (rec_cis) CK(cudaMemcpyAsync(rec_cis, h->snapInt + 2 * (size_t)NAt, sizeof(int) * NAt, cudaMemcpyDeviceToHost, h->copyStream));
    CK(cudaEventRecord(h->evSnapDone, h->copyStream));
    h->snapPending = true;
    return KMC_OK;
}
extern "C" int kmc_snapshot_wait(kmc_handle *h) {
    if (!h) return KMC_ERR_INVALID;
    if (!h->snapPending) return KMC_OK;
    CK(cudaSetDevice(h->P.device));
    CK(cudaEventSynchronize(h->evSnapDone));
    h->snapPending = false;
    return KMC_OK;
}

extern "C" int kmc_set_packed(kmc_handle *h, const double *rec_pose, const double *lig_pose, const int32_t *rec_lig,
                              const int32_t *rec_site, const int32_t *rec_cis, int64_t step_done) {
    if (!h || !rec_pose || !lig_pose) { if (h) h->err = "kmc_set_packed: null pose"; return KMC_ERR_INVALID; }
    if ((rec_lig != nullptr) != (rec_site != nullptr)) { h->err = "kmc_set_packed: rec_lig and rec_site go together"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    rc = stage_alloc(h); if (rc) return rc;
    Dev &D = h->D; const int NAt = h->NAt, NBt = h->NBt; cudaStream_t st = h->stream;
    int *dl = h->stageInt, *ds = h->stageInt + NAt, *dc = h->stageInt + 2 * (size_t)NAt, *dflag = h->stageInt + 3 * (size_t)NAt;
    CK(cudaMemcpyAsync(h->stageRec, rec_pose, sizeof(double) * 6 * (size_t)NAt, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.lig, lig_pose, sizeof(double) * 24 * (size_t)NBt, cudaMemcpyHostToDevice, st));
    if (rec_lig) { CK(cudaMemcpyAsync(dl, rec_lig, sizeof(int) * NAt, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(ds, rec_site, sizeof(int) * NAt, cudaMemcpyHostToDevice, st)); }
    if (rec_cis) CK(cudaMemcpyAsync(dc, rec_cis, sizeof(int) * NAt, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(D.ligRec, 0xff, sizeof(int) * 3 * (size_t)NBt, st));
    CK(cudaMemsetAsync(dflag, 0, sizeof(int), st));
    const Args A{D, h->K};
    LAUNCH(KID_SERIES, (k_pack_set<<<nblk(std::max(NAt, 1), 256), 256, 0, st>>>(A, h->stageRec, rec_lig ? dl : nullptr, rec_lig ? ds : nullptr, rec_cis ? dc : nullptr, dflag)));
    int one = 1, flags = 0;
    unsigned long long s64 = (unsigned long long)step_done;
    CK(cudaMemcpyAsync(D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(D.step64, &s64, sizeof s64, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(&flags, dflag, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    h->step_done = step_done; h->stepped = false; h->sinceBuild = 0;
    if (flags) { h->err = std::string("kmc_set_packed: inconsistent bond table (") + ((flags & 1) ? "R-L bond " : "") + ((flags & 2) ? "cis bond " : "") + "invalid or asymmetric)"; return KMC_ERR_STATE; }
    return KMC_OK;
}

// ------------------------------------------------------------------------------------------------
// the sweep
// ------------------------------------------------------------------------------------------------

// the gated parallel rebuild of the complex tables: one cooperative launch (grid = what fits on the device at once)
static void launch_cx_rebuild(kmc_handle *h, const Args &A, cudaStream_t st) {
    if (h->cxBlocks == 0) {
        int per = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_cx_rebuild, 256, 0);
        h->cxBlocks = std::max(1, std::min(per, 4)) * h->nSM;
    }
    void *args[] = {(void *)&A};
    LAUNCH(KID_UF_INIT, (cudaLaunchCooperativeKernel((const void *)k_cx_rebuild, dim3(h->cxBlocks), dim3(256), args, 0, st)));
}
// all launches of one time step (main.cpp:461-2202) on stream st; no host synchronisation anywhere.
// A.K.phase: 0 = the step rebuilds the neighbour grid (and, on the sparse path, the pair list), 1 = it reuses them.
static void issue_step(kmc_handle *h, const Args &A, cudaStream_t st) {
    const Dev &D = A.D;
    const int B = 128, NT = h->NT, NAt = h->NAt, NBt = h->NBt;
    const bool build = A.K.phase == 0;
    // step begin + S1: incremental update of the complexes the last step's reactions touched (one thread); the parallel rebuild of
    // the whole table is gated on a device flag (state loaded, strip refresh, too many changes at once)
    LAUNCH(KID_STEP_BEGIN, (k_step_begin<<<1, 32, 0, st>>>(A, 1)));
    launch_cx_rebuild(h, A, st);
    // S2 proposals: free receptors / cis dimers, free ligands and complexes are disjoint sets of molecules -- three kernels side
    // by side (forked branches of the graph; on one stream when per-kernel timing is on)
    const int forkMask = h->profiling ? 0 : h->forkMask;     // bit0: receptor/ligand proposals side by side (measured slower than back to back), bit1: special entries, bit2: complexes
    const bool fork = forkMask & 1, fork2 = forkMask & 2, forkC = forkMask & 5;                       // bit2: only the complexes on a side branch
    cudaStream_t s1 = fork ? h->side[0] : st, s2 = forkC ? h->side[1] : st;
    if (fork || forkC) { cudaEventRecord(h->evFork[0], st); cudaStreamWaitEvent(h->side[0], h->evFork[0], 0); cudaStreamWaitEvent(s2, h->evFork[0], 0); }
    LAUNCH(KID_PROPOSE_SIMPLE, (k_propose_rec<<<std::min(nblk(std::max(NAt, 1), REC_TILE), h->nSM * RECMINB), REC_TILE, 0, st>>>(A)));
    LAUNCH(KID_PROPOSE_LIG, (k_propose_lig<<<nblk(NBt, B), B, 0, s1>>>(A)));
    h->tlStream = s2;
    LAUNCH(KID_PROPOSE_COMPLEX, (k_propose_complex<<<std::min(nblk(NBt, CX_WARPS), h->nSM * 12), 32 * CX_WARPS, 0, s2>>>(A)));      // (large complexes: rare, first)
    LAUNCH(KID_PROPOSE_COMPLEX_SMALL, (k_propose_complex_small<<<std::min(nblk(NBt, 128), h->nSM * 16), 128, 0, s2>>>(A, h->cxGroups ? 0 : 1)));
    // small complexes with several ligands: 8 lanes each, next to the thread-per-complex kernel when the ligand proposals do not use that branch
    cudaStream_t s4 = (forkC && !fork) ? h->side[0] : s2;
    if (h->cxGroups) { h->tlStream = s4; LAUNCH(KID_PROPOSE_COMPLEX_MULTI, (k_propose_complex_multi<<<std::min(nblk(NBt, CX_GROUPS), h->nSM * 32), CX_G * CX_GROUPS, 0, s4>>>(A))); }
    h->tlStream = nullptr;
    if (fork || forkC) {
        cudaEventRecord(h->evJoin[0], h->side[0]); cudaEventRecord(h->evJoin[1], s2);
        cudaStreamWaitEvent(st, h->evJoin[0], 0); cudaStreamWaitEvent(st, h->evJoin[1], 0);
    }
    if (build) {
        // neighbour grid: the histogram was accumulated by the propose kernels (cellCount is zero at step start: k_scan_down clears it)
        LAUNCH(KID_SCAN_REDUCE, (k_scan_reduce<<<h->scanBlocks, 256, 0, st>>>((const int4 *)D.cellCount, D.scanTmp)));
        LAUNCH(KID_SCAN_SUMS, (k_scan_sums<<<1, 1024, 0, st>>>(D.scanTmp, h->scanBlocks)));
        LAUNCH(KID_SCAN_DOWN, (k_scan_down<<<h->scanBlocks, 256, 0, st>>>((int4 *)D.cellCount, D.scanTmp, (int4 *)D.cellStart)));
        LAUNCH(KID_GRID_SCATTER, (k_grid_scatter<<<nblk(NT, 256), 256, 0, st>>>(A)));
    }
    // S2g: every (probe, neighbour) pair within reach is classified once (+ reaction-pair pre-selection), then the order
    // dependence is settled from the pending findings
    const int gl = std::min(nblk(NT, B), h->nSM * 8);
    if (h->useCells) {
        if (build) LAUNCH(KID_RESOLVE, (k_cells_cut<<<std::min(nblk(NT + NT / 16 + 1, CTHREADS), h->nSM * CMINB * 16), CTHREADS, 0, st>>>(A)));
        cudaStream_t s3 = fork2 ? h->side[0] : st;
        if (!build && fork2) { cudaEventRecord(h->evFork[1], st); cudaStreamWaitEvent(s3, h->evFork[1], 0); }
        LAUNCH(KID_PAIRS_EVAL, (k_pairs_eval<<<std::min(nblk(NT / 2 + 1, PE_CHUNK) + h->nSM, h->nSM * 16), PTHREADS, 0, st>>>(A)));
        if (!build) {           // the special entries next to the list pairs (both only publish findings)
            h->tlStream = s3;
            LAUNCH(KID_SPECIAL, (k_special_pairs<<<h->nSM, 32 * SP_WARPS, 0, s3>>>(A)));
            h->tlStream = nullptr;
            if (fork2) { cudaEventRecord(h->evJoin[2], s3); cudaStreamWaitEvent(st, h->evJoin[2], 0); }
        }
    } else LAUNCH(KID_RESOLVE, (k_resolve_tiles<<<h->nTiles, TTHREADS, 0, st>>>(A)));
    LAUNCH(KID_PEND_RESOLVE, (k_pend_resolve<<<1, 1024, 0, st>>>(A)));
    // the rejected units get their old poses back on a side branch (it needs the settled decisions, nothing of S3), next to S3
    const bool forkR = (forkMask & 8) != 0;
    cudaStream_t s5 = forkR ? h->side[1] : st;
    if (forkR) {
        cudaEventRecord(h->evFork[2], st); cudaStreamWaitEvent(s5, h->evFork[2], 0);
        h->tlStream = s5;
        LAUNCH(KID_FINISH, (k_finish<<<std::min(nblk(NT / 256 + 1, 128), h->nSM * 2), 128, 0, s5>>>(A, 1)));
        h->tlStream = nullptr;
        cudaEventRecord(h->evJoin[3], s5);
    }
    // S3
    LAUNCH(KID_REACT_PAIRS, (k_react_pairs<<<std::min(nblk(NT / 16 + 1, RPTHREADS) + h->nSM, h->nSM * 32), RPTHREADS, 0, st>>>(A)));
    LAUNCH(KID_REACT_RESOLVE, (k_react_resolve<<<1, 1024, 0, st>>>(A)));
    LAUNCH(KID_FINISH, (k_finish<<<nblk(std::max((NAt + 3) / 4, 1), 256), 256, 0, st>>>(A, forkR ? 2 : 3)));
    if (forkR) cudaStreamWaitEvent(st, h->evJoin[3], 0);
}
static void swap_buffers(Dev &D) { std::swap(D.recC, D.recCn); std::swap(D.recS2, D.recS2n); std::swap(D.recS3, D.recS3n); std::swap(D.lig, D.lign); }

// the step as a CUDA graph, one per phase and buffer parity (the committed/new buffers alternate, S4 is a pointer swap)
static int ensure_graphs(kmc_handle *h) {
    if (h->gexec[0][0]) return KMC_OK;
    const int64_t saved = h->launches;
    for (int ph = 0; ph < (h->listEvery > 1 ? 2 : 1); ph++) {
        const int64_t before = h->launches;
        for (int p = 0; p < 2; p++) {
            Dev Dp = h->D;
            if (p != h->parity) swap_buffers(Dp);
            Args A{Dp, h->K};
            A.K.phase = ph;
            cudaGraph_t g = nullptr;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            h->tlCount = 0;
            issue_step(h, A, h->stream);
            CK(cudaStreamEndCapture(h->stream, &g));
            CK(cudaGraphInstantiate(&h->gexec[ph][p], g, 0));
            cudaGraphDestroy(g);
        }
        h->launches_per_step[ph] = (int)((h->launches - before) / 2);
    }
    h->launches = saved;
    return KMC_OK;
}

// What the host does with a snapshot of the device scalars (taken at kmc_sync, or asynchronously every MON_EVERY steps of a
// long kmc_step): report capacity overflows, and decide the list-reuse back-off.
// List reuse pays while few molecules outrun their grid entries. A state in which many do (most ligands bound: the members
// of a rotating complex swing 10-30 A per step) makes every reuse step walk the stale grid once per such molecule; from
// NT/64 special entries per step on, rebuilding every step is the cheaper exact path; the graphs are captured again with
// the new constants. KMC_ADAPT=0 disables it.
// coarser cells for a wider list: the arrays were sized for the finer grid the handle started with
static void regrid(kmc_handle *h, double edge) {
    Consts &K = h->K; const kmc_params &P = h->P;
    if (edge <= 1.0 / K.cellInv) return;
    K.gx0 = -P.box[0] / 2 - edge; K.gy0 = -P.box[1] / 2 - edge;
    K.ncx = (int)ceil((P.box[0] + 2 * edge) / edge); K.ncy = (int)ceil((P.box[1] + 2 * edge) / edge);
    K.cellInv = 1.0 / edge;
    const float maxc = (float)(std::max(P.box[0], P.box[1]) + 2 * edge);
    K.cutMargin = 0.05f + 2 * (nextafterf(maxc, INFINITY) - maxc);
    h->D.ncell = K.R * K.ncx * K.ncy;
    h->scanBlocks = (h->D.ncell + 1 + SCAN_TILE - 1) / SCAN_TILE;
}
static int examine_scalars(kmc_handle *h, const int *scal) {
    const int nspec = h->adaptSkip ? 0 : std::max(scal[S_NSPEC], scal[S_NSPEC_MAX]);
    h->adaptSkip = false;
    const int adaptMin = getenv("KMC_ADAPT_MIN") ? atoi(getenv("KMC_ADAPT_MIN")) : 64;          // (tests lower it to meet the back-off on a small system)
    if (h->listEvery > 1 && h->adapt && nspec > std::max(h->NT / 64, adaptMin)) {
        // Level 1 (a single membrane, not strips): many molecules outrun a 12 A skin -- members of rotating complexes swing 10-30 A
        // per step -- but a list built with a 30 A skin and 60 A of drift allowance, rebuilt every 4th step on coarser cells, still
        // beats rebuilding every step (oligomerised membrane: 0.47 against 0.52 ms per step). Level 2, if even that setting makes
        // more than NT/64 special entries per step: rebuild every step.
        const double wide = std::max({h->K.reachLL, h->K.reachOn, h->K.reachCis}) + 2 * 30.0 + 2 * 60.0;          // cut radius of the wide list
        const double pairsPerMolecule = 0.5 * 3.14159265 * wide * wide * h->NT / ((double)h->K.R * h->K.Lx * h->K.Ly);   // (the list holds 8 per molecule)
        const bool level1 = h->adaptLevel == 0 && !h->strip_on && pairsPerMolecule <= 3.0 && !getenv("KMC_SKIN") && !getenv("KMC_DRIFT") && !getenv("KMC_REUSE");
        if (level1) {
            h->adaptLevel = 1; h->listEvery = 4; h->K.skin = 30.0; h->K.drift = 60.0;
            regrid(h, std::max({h->K.reachLL, h->K.reachOn, h->K.reachCis}) + 2 * h->K.skin + 2 * h->K.drift + 1.0);
        } else {
            h->adaptLevel = 2; h->listEvery = 1;
            h->K.drift = 0; if (!getenv("KMC_SKIN")) h->K.skin = 24.0;
        }
        h->sinceBuild = 0; h->adapted = true; h->adaptSkip = true;          // (a snapshot taken before the switch says nothing about the new setting)
        cudaMemsetAsync(h->D.scal + S_NSPEC_MAX, 0, sizeof(int), h->stream); cudaMemsetAsync(h->D.scal + S_NSPEC, 0, sizeof(int), h->stream);
        for (int p = 0; p < 4; p++) if (h->gexec[p >> 1][p & 1]) { cudaGraphExecDestroy(h->gexec[p >> 1][p & 1]); h->gexec[p >> 1][p & 1] = nullptr; }
    }
    int ovf = scal[S_OVERFLOW];
    if (h->listEvery <= 1) ovf &= ~32;          // a pair list that is rebuilt every step may overflow into in-place evaluation
    if (ovf) {
        h->err = "device buffer overflow (mask " + std::to_string(ovf) + ((ovf & 32) ? "; pair list too small for list reuse: set KMC_REUSE=1" : "") +
                 ((ovf & 64) ? "; strips: a complex is wider than the halo budget (halo_width - refresh_every * reach)" : "") +
                 ((ovf & 128) ? "; strips: a message or the local capacity is too small" : "") +
                 ((ovf & 256) ? "; strips: a halo copy inside the exact zone differs from its owner's original" : "") + ")";
        return KMC_ERR_CAPACITY;
    }
    return KMC_OK;
}
static int monitor_poll(kmc_handle *h) {
    if (h->monPending) {
        CK(cudaEventSynchronize(h->monEvent));          // recorded MON_EVERY steps ago: long complete unless the host runs far ahead
        h->monPending = false;
        int rc = examine_scalars(h, h->monHost); if (rc) return rc;
    }
    CK(cudaMemcpyAsync(h->monHost, h->D.scal, sizeof(int) * S_COUNT, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaEventRecord(h->monEvent, h->stream));
    h->monPending = true; h->sinceMon = 0;
    return KMC_OK;
}

static int strip_auto_refresh(kmc_handle *h);

extern "C" int kmc_step(kmc_handle *h, int64_t n) {
    if (!h) return KMC_ERR_INVALID;
    if (n < 0) { h->err = "kmc_step: negative step count"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    cudaStream_t st = h->stream;
    if (h->fused) {
        // one launch advances every replica by up to 8192 steps; S4 is a pointer swap inside the kernel, so an odd count leaves
        // the committed state in the other buffer set
        for (int64_t left = n; left > 0;) {
            if (h->sinceMon >= MON_EVERY) { rc = monitor_poll(h); if (rc) return rc; }
            const int chunk = (int)std::min<int64_t>(left, 8192);
            Args A{h->D, h->K};
            A.K.phase = 2;
            const int slots = h->smallSlots, groups = (h->R + slots - 1) / slots;
            const size_t dyn = slots * (sizeof(SmallShared) + small_dyn_bytes(h->NA, h->NB));
            const bool tickets = groups > h->smallGrid;          // more replica groups than resident CTAs: persistent grid, groups dealt in chunks of 64 steps
            if (tickets) CK(cudaMemsetAsync(h->smallQueue, 0, sizeof(int) * (size_t)(1 + groups), st));
            const int grid = tickets ? h->smallGrid : groups, ch = tickets ? 64 : chunk;
            int *q = tickets ? h->smallQueue : nullptr;
            if (slots == 4) LAUNCH(KID_SMALL_STEP, (k_small_step<4, SMALL_T><<<grid, 4 * SMALL_T, dyn, st>>>(A, (unsigned long long)h->step_done, chunk, ch, q)));
            else if (h->smallWide && !tickets) LAUNCH(KID_SMALL_STEP, (k_small_step<1, 2 * SMALL_T><<<grid, 2 * SMALL_T, dyn, st>>>(A, (unsigned long long)h->step_done, chunk, ch, q)));
            else LAUNCH(KID_SMALL_STEP, (k_small_step<1, SMALL_T><<<grid, SMALL_T, dyn, st>>>(A, (unsigned long long)h->step_done, chunk, ch, q)));
            if (chunk & 1) { swap_buffers(h->D); h->parity ^= 1; }
            h->step_done += chunk; h->passes += chunk; h->sinceMon += chunk; left -= chunk;
            h->stepped = true;
        }
        CK(cudaGetLastError());
        return KMC_OK;
    }
    for (int64_t it = 0; it < n; it++) {
        if (h->sinceMon >= MON_EVERY) { rc = monitor_poll(h); if (rc) return rc; }
        if (!h->profiling && h->use_graph && !h->gexec[0][0]) { rc = ensure_graphs(h); if (rc) return rc; }
        const int phase = (h->sinceBuild == 0 || h->sinceBuild >= h->listEvery) ? 0 : 1;
        if (phase == 0) h->sinceBuild = 0;
        if (++h->epoch >= 0xfffffff0u) {         // the stamp of the per-cell chains is about to wrap: start a new era
            if (h->D.cellHead) CK(cudaMemsetAsync(h->D.cellHead, 0, sizeof(unsigned long long) * (size_t)h->cellHeadCap, st));
            CK(cudaMemsetAsync(h->D.scal + S_EPOCH, 0, sizeof(int), st));
            CK(cudaMemsetAsync(h->D.cxStamp, 0, sizeof(int) * (size_t)h->NT, st));          // (the stamps of the incremental complex update restart too:
            CK(cudaMemsetAsync(h->D.scal + S_TOPO_DIRTY, 1, 1, st));                         //  one full rebuild clears the marks)
            h->epoch = 1;
        }
        if (h->profiling || !h->use_graph) {
            Args A{h->D, h->K};
            A.K.phase = phase;
            issue_step(h, A, st);
            if (h->profiling && (it & 15) == 15) harvest(h, false);
        } else {
            CK(cudaGraphLaunch(h->gexec[phase][h->parity], st));
            h->launches += h->launches_per_step[phase];
        }
        // S4: the new buffers become the committed state
        swap_buffers(h->D); h->parity ^= 1;
        h->sinceBuild++; h->sinceMon++;
        h->passes += 1; h->step_done++; h->stepped = true;
        if (h->strip_every > 0 && ++h->strip_since >= h->strip_every) { rc = strip_auto_refresh(h); if (rc) return rc; }
    }
    CK(cudaGetLastError());
    return KMC_OK;
}

extern "C" int kmc_sync(kmc_handle *h) {
    if (!h) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    CK(cudaStreamSynchronize(h->stream));
    int scal[S_COUNT];
    CK(cudaMemcpy(scal, h->D.scal, sizeof scal, cudaMemcpyDeviceToHost));
    h->monPending = false; h->sinceMon = 0;
    return examine_scalars(h, scal);
}

// n steps bracketed by CUDA events on the handle's stream (the stream the kernels are launched on)
extern "C" int kmc_step_timed(kmc_handle *h, int64_t n, double *elapsed_ms) {
    if (!h || !elapsed_ms) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventRecord(a, h->stream));
    int rc = kmc_step(h, n);
    if (rc == KMC_OK) {
        CK(cudaEventRecord(b, h->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
        *elapsed_ms = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return rc;
}
extern "C" int kmc_timeline_print(kmc_handle *h) {
    if (!h || !h->timeline) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    CK(cudaStreamSynchronize(h->stream));
    unsigned long long t[128];
    CK(cudaMemcpy(t, h->timeline, sizeof t, cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < h->tlCount; i++) t0 = std::min(t0, t[2 * i]);
    for (int i = 0; i < h->tlCount; i++)
        fprintf(stderr, "timeline %-26s start %8.2f us  end %8.2f us  (%.2f)\n", g_kernel_names[h->tlId[i]], (t[2 * i] - t0) * 1e-3, (t[2 * i + 1] - t0) * 1e-3, (t[2 * i + 1] - t[2 * i]) * 1e-3);
    return KMC_OK;
}
extern "C" int kmc_profile(kmc_handle *h, int32_t enable) {
    if (!h) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    harvest(h, true);
    h->profiling = enable != 0;
    if (enable) { for (int i = 0; i < KMC_NKERNELS; i++) { h->kms[i] = 0; h->kcount[i] = 0; } }
    return KMC_OK;
}
extern "C" int kmc_profile_get(kmc_handle *h, int32_t idx, const char **name, double *total_ms, int64_t *launches) {
    if (!h || idx < 0) return KMC_ERR_INVALID;
    if (idx >= KMC_NKERNELS) return 1;
    CK(cudaSetDevice(h->P.device));
    harvest(h, true);
    if (name) *name = (idx == KID_RESOLVE && h->useCells) ? "k_cells_cut" : g_kernel_names[idx];
    if (total_ms) *total_ms = h->kms[idx];
    if (launches) *launches = h->kcount[idx];
    return KMC_OK;
}

// ------------------------------------------------------------------------------------------------
// outputs
// ------------------------------------------------------------------------------------------------
struct HostComplexes { std::vector<int> unitOf, cxSize, cxOff, rowWork; };
static int fetch_complexes(kmc_handle *h, HostComplexes &hc) {
    Dev &D = h->D;
    hc.unitOf.resize(h->NT); hc.cxSize.resize(h->NBt); hc.cxOff.resize(h->NBt); hc.rowWork.resize(h->NT);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(hc.unitOf.data(), D.unitOf, sizeof(int) * h->NT, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.cxSize.data(), D.cxSize, sizeof(int) * h->NBt, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.cxOff.data(), D.cxOff, sizeof(int) * h->NBt, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.rowWork.data(), D.rowWork, sizeof(int) * h->NT, cudaMemcpyDeviceToHost));
    return KMC_OK;
}

extern "C" int kmc_get_series(kmc_handle *h, int32_t rep, kmc_series *out) {
    if (!h || !out) return KMC_ERR_INVALID;
    if (rep < 0 || rep >= h->R) { h->err = "kmc_get_series: bad replica"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    Dev &D = h->D;
    CK(cudaMemsetAsync(h->d_series, 0, sizeof(int) * 6 * h->R, h->stream));
    const Args A{D, h->K};
    cudaStream_t st = h->stream;
    LAUNCH(KID_SERIES, (k_series<<<nblk(std::max(h->NAt, 1), 256), 256, 0, st>>>(A, h->d_series)));
    if (h->stepped) LAUNCH(KID_SERIES, (k_cx_stats<<<nblk(std::max(h->NBt, 1), 256), 256, 0, st>>>(A, h->d_series + 4 * h->R)));
    int s4[4], mx = 0;
    CK(cudaMemcpyAsync(s4, h->d_series + 4 * rep, sizeof s4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&mx, D.maxComplex + rep, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    memset(out, 0, sizeof *out);
    out->step = h->step_done; out->bond_num_rl = s4[0]; out->bond_num_mono_cis = s4[1]; out->bond_num_cis = s4[2];
    out->bond_num = s4[0] + s4[1] + s4[2]; out->max_complex = mx;
    if (h->stepped) {
        int cx[2] = {0, 0};
        CK(cudaMemcpy(cx, h->d_series + 4 * h->R + 2 * rep, sizeof cx, cudaMemcpyDeviceToHost));
        out->n_complexes = cx[0]; out->n_in_complexes = cx[1];
        if (out->n_complexes) out->cluster_size = (double)out->n_in_complexes / out->n_complexes;     // main.cpp:2200-2202
    }
    return KMC_OK;
}

extern "C" int64_t kmc_get_complexes(kmc_handle *h, int32_t rep, int32_t *row_len, int32_t *members, int64_t cap) {
    if (!h || !row_len) return KMC_ERR_INVALID;
    if (rep < 0 || rep >= h->R) { h->err = "kmc_get_complexes: bad replica"; return KMC_ERR_INVALID; }
    if (!h->stepped) { h->err = "kmc_get_complexes: no completed step yet (the reference lists complexes of the last step)"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    HostComplexes hc; rc = fetch_complexes(h, hc); if (rc) return rc;
    int64_t tot = 0;
    auto ref_id_host = [&](int gid) { return gid < h->NAt ? gid % h->NA + 1 : h->NA + (gid - h->NAt) % h->NB + 1; };
    for (int l = 0; l < h->NB; l++) {
        int b = rep * h->NB + l;
        if (hc.unitOf[h->NAt + b] != h->NAt + b) { row_len[l] = 0; continue; }
        int size = hc.cxSize[b]; row_len[l] = size;
        for (int i = 0; i < size; i++) {
            int gid = size > 1 ? hc.rowWork[hc.cxOff[b] + i] : h->NAt + b;
            if (members && tot < cap) members[tot] = ref_id_host(gid);
            tot++;
        }
    }
    return tot;
}

extern "C" int kmc_get_complex_labels(kmc_handle *h, int32_t rep, int32_t *root) {
    if (!h || !root) return KMC_ERR_INVALID;
    if (rep < 0 || rep >= h->R) { h->err = "kmc_get_complex_labels: bad replica"; return KMC_ERR_INVALID; }
    if (!h->stepped) { h->err = "kmc_get_complex_labels: no completed step yet (labels are those of the last step's sweep)"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    std::vector<int> unitOf(h->NT);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(unitOf.data(), h->D.unitOf, sizeof(int) * h->NT, cudaMemcpyDeviceToHost));
    auto ref_id_host = [&](int gid) { return gid < h->NAt ? gid % h->NA + 1 : h->NA + (gid - h->NAt) % h->NB + 1; };
    root[0] = 0;
    for (int a = 0; a < h->NA; a++) root[a + 1] = ref_id_host(unitOf[rep * h->NA + a]);
    for (int b = 0; b < h->NB; b++) root[h->NA + 1 + b] = ref_id_host(unitOf[h->NAt + rep * h->NB + b]);
    return KMC_OK;
}

extern "C" int kmc_get_oligomer_hist(kmc_handle *h, int32_t rep, int64_t *hist, int32_t nbins) {
    if (!h || !hist || nbins < 2) return KMC_ERR_INVALID;
    if (rep >= h->R) { h->err = "kmc_get_oligomer_hist: bad replica"; return KMC_ERR_INVALID; }
    if (!h->stepped) { h->err = "kmc_get_oligomer_hist: no completed step yet"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    HostComplexes hc; rc = fetch_complexes(h, hc); if (rc) return rc;
    for (int i = 0; i < nbins; i++) hist[i] = 0;
    int b0 = rep < 0 ? 0 : rep * h->NB, b1 = rep < 0 ? h->NBt : (rep + 1) * h->NB;
    for (int b = b0; b < b1; b++)
        if (hc.unitOf[h->NAt + b] == h->NAt + b) hist[std::min(hc.cxSize[b], nbins - 1)]++;
    return KMC_OK;
}

extern "C" int kmc_get_grid(kmc_handle *h, double *x0, double *y0, double *inv_edge, int32_t *ncx, int32_t *ncy) {
    if (!h) return KMC_ERR_INVALID;
    if (x0) *x0 = h->K.keyX0; if (y0) *y0 = h->K.keyY0; if (inv_edge) *inv_edge = h->K.keyInv;          // (the cells the production order colours)
    if (ncx) *ncx = h->K.ncx; if (ncy) *ncy = h->K.ncy;
    return KMC_OK;
}

extern "C" int kmc_get_accept(kmc_handle *h, int32_t rep, int32_t *accepted) {
    if (!h || !accepted) return KMC_ERR_INVALID;
    if (rep < 0 || rep >= h->R) { h->err = "kmc_get_accept: bad replica"; return KMC_ERR_INVALID; }
    int rc = select_device(h); if (rc) return rc;
    std::vector<int> unitOf(h->NT), st(h->NT);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(unitOf.data(), h->D.unitOf, sizeof(int) * h->NT, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), h->D.unitRes, sizeof(int) * h->NT, cudaMemcpyDeviceToHost));      // 0 = accepted, bit0 = reverted
    accepted[0] = 1;
    for (int a = 0; a < h->NA; a++) accepted[a + 1] = st[unitOf[rep * h->NA + a]] == 0;
    for (int b = 0; b < h->NB; b++) accepted[h->NA + 1 + b] = st[unitOf[h->NAt + rep * h->NB + b]] == 0;
    return KMC_OK;
}

extern "C" int kmc_alignment_window(double length, double *window) {
    if (!window) return KMC_ERR_INVALID;
    same_window(length, window);
    return window[0] <= window[1] ? KMC_OK : 1;
}

extern "C" int kmc_get_step_path(kmc_handle *h) { return h ? (h->fused ? 1 : 0) : KMC_ERR_INVALID; }

extern "C" int kmc_get_live_counts(kmc_handle *h, int32_t *n_rec, int32_t *n_lig) {
    if (!h) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    CK(cudaStreamSynchronize(h->stream));
    int live[2];
    CK(cudaMemcpy(live, h->D.scal + S_NA_LIVE, sizeof live, cudaMemcpyDeviceToHost));
    if (n_rec) *n_rec = live[0]; if (n_lig) *n_lig = live[1];
    return KMC_OK;
}

extern "C" int kmc_get_events(kmc_handle *h, int64_t *ev) {
    if (!h || !ev) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    CK(cudaStreamSynchronize(h->stream));
    unsigned long long d[EV_COUNT];
    CK(cudaMemcpy(d, h->D.events, sizeof d, cudaMemcpyDeviceToHost));
    for (int i = 0; i < EV_COUNT; i++) ev[i] = (int64_t)d[i];
    ev[EV_PASSES] = h->passes; ev[EV_LAUNCHES] = h->launches;
    int scal[S_COUNT];
    CK(cudaMemcpy(scal, h->D.scal, sizeof scal, cudaMemcpyDeviceToHost));
    ev[12] = scal[S_NSURV]; ev[13] = scal[S_NSPEC]; ev[14] = scal[S_NPEND]; ev[15] = scal[S_NPAIR];      // sizes of the last step's work lists
    return KMC_OK;
}

// ---- the reference's output records ----
// main.cpp:2249-2251: fixed, precision 3, widths 15 5 5 10 10 10 10. Pure host formatting (no device needed).
extern "C" int kmc_format_bond_dat(double dt, const kmc_series *s, char *buf, int32_t cap) {
    if (!s || !buf || cap < 1) return KMC_ERR_INVALID;
    int n = snprintf(buf, (size_t)cap, "%15.3f%5d%5d%10d%10d%10.3f%10d\n", (double)s->step * dt, s->bond_num_rl, s->bond_num_mono_cis,
                     s->bond_num_cis, s->bond_num, s->cluster_size, s->max_complex);
    return n < cap ? n : KMC_ERR_CAPACITY;
}
// main.cpp:2293-2301: header with the default ostream float format (%g), then one line per ligand: its row of `results`,
// every member followed by two blanks (an empty line for a ligand that is not a BFS root)
extern "C" int64_t kmc_format_cluster_log(double dt, int64_t step, int32_t n_ligand, const int32_t *row_len, const int32_t *members,
                                          char *buf, int64_t cap) {
    if (!row_len || !buf || cap < 1) return KMC_ERR_INVALID;
    std::string out;
    char tmp[64];
    snprintf(tmp, sizeof tmp, "Hello Cluster!, t=%g\n", (double)step * dt); out += tmp;
    int64_t o = 0;
    for (int l = 0; l < n_ligand; l++) {
        for (int i = 0; i < row_len[l]; i++) { snprintf(tmp, sizeof tmp, "%d  ", members[o + i]); out += tmp; }
        o += row_len[l];
        out += '\n';
    }
    if ((int64_t)out.size() + 1 > cap) return KMC_ERR_CAPACITY;
    memcpy(buf, out.c_str(), out.size() + 1);
    return (int64_t)out.size();
}
extern "C" int kmc_write_bond_dat(kmc_handle *h, int32_t rep, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    kmc_series s; int rc = kmc_get_series(h, rep, &s); if (rc) return rc;
    char line[160];
    rc = kmc_format_bond_dat(h->P.dt, &s, line, sizeof line); if (rc < 0) return rc;
    FILE *f = fopen(path, "a");
    if (!f) { h->err = std::string("cannot open ") + path; return KMC_ERR_IO; }
    fputs(line, f); fclose(f);
    return KMC_OK;
}
extern "C" int kmc_write_cluster_log(kmc_handle *h, int32_t rep, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    std::vector<int32_t> len(h->NB), mem(h->N + 1);
    int64_t tot = kmc_get_complexes(h, rep, len.data(), mem.data(), (int64_t)mem.size());
    if (tot < 0) return (int)tot;
    std::vector<char> buf((size_t)h->N * 14 + h->NB + 128);
    int64_t n = kmc_format_cluster_log(h->P.dt, h->step_done, h->NB, len.data(), mem.data(), buf.data(), (int64_t)buf.size());
    if (n < 0) return (int)n;
    FILE *f = fopen(path, "a");
    if (!f) { h->err = std::string("cannot open ") + path; return KMC_ERR_IO; }
    fwrite(buf.data(), 1, (size_t)n, f); fclose(f);
    return KMC_OK;
}
extern "C" int kmc_run(kmc_handle *h, int64_t n_steps, int32_t output_every, const char *dir) {
    if (!h || !dir || output_every < 1) return KMC_ERR_INVALID;
    std::string d(dir);
    const int64_t end = h->step_done + n_steps;
    while (h->step_done < end) {
        int64_t next = std::min<int64_t>(end, (h->step_done / output_every + 1) * output_every);
        int rc = kmc_step(h, next - h->step_done); if (rc) return rc;
        rc = kmc_sync(h); if (rc) return rc;                                    // capacity errors of the interval are reported here, never dropped
        if (h->step_done % output_every == 0)                                   // main.cpp:2247, 2291
            for (int r = 0; r < h->R; r++) {
                std::string suf = h->R > 1 ? "." + std::to_string(r) : "";
                rc = kmc_write_bond_dat(h, r, (d + "/bond.dat" + suf).c_str()); if (rc) return rc;
                rc = kmc_write_cluster_log(h, r, (d + "/cluster.log" + suf).c_str()); if (rc) return rc;
                rc = kmc_write_gro(h, r, (d + "/test.gro" + suf).c_str()); if (rc) return rc;                 // main.cpp:2258
                rc = kmc_write_checkpoint(h, r, (d + "/position.cpt" + suf).c_str()); if (rc) return rc;      // main.cpp:2206
            }
    }
    return KMC_OK;
}

// ------------------------------------------------------------------------------------------------
// scalable initial configuration (replaces the O(N^2) insertion of main.cpp:273-456)
// ------------------------------------------------------------------------------------------------
namespace {
// host-side cell grid (linked lists, about one molecule per cell) used only by the initial-configuration generator
struct HostGrid {
    double edge, x0, y0; int nx, ny; std::vector<int> head, next;
    HostGrid(double Lx, double Ly, double reach, int n) : x0(-Lx / 2), y0(-Ly / 2) {
        double e = std::max(reach, sqrt(Lx * Ly / std::max(n, 1)));
        nx = std::max(1, (int)(Lx / e)); ny = std::max(1, (int)(Ly / e));
        edge = std::max(Lx / nx, Ly / ny);
        head.assign((size_t)nx * ny, -1); next.assign(std::max(n, 1), -1);
    }
    int cx(double x) const { return std::min(std::max((int)((x - x0) / edge), 0), nx - 1); }
    int cy(double y) const { return std::min(std::max((int)((y - y0) / edge), 0), ny - 1); }
    template <class F> bool any_near(double x, double y, F f) const {
        int a = cx(x), b = cy(y);
        for (int yy = std::max(b - 1, 0); yy <= std::min(b + 1, ny - 1); yy++)
            for (int xx = std::max(a - 1, 0); xx <= std::min(a + 1, nx - 1); xx++)
                for (int id = head[(size_t)yy * nx + xx]; id >= 0; id = next[id]) if (f(id)) return true;
        return false;
    }
    void put(double x, double y, int id) { size_t c = (size_t)cy(y) * nx + cx(x); next[id] = head[c]; head[c] = id; }
};
}  // namespace

// pure host code (no device): fills rec[R*NA][6], lig[R*NB][24]; returns an error text or nullptr
static const char *generate_random(const kmc_params &P, const Consts &K, uint64_t init_seed, int32_t sort_cells, double *recOut, double *ligOut) {
    const int NA = P.n_receptor, NB = P.n_ligand, R = P.n_replicas;
    const double rs = P.rB * 2 / sqrt(3.0);
    const double exRR = P.rA + P.rA, exRL = P.rA + rs + P.rB, exLL = rs + rs + 2 * P.rB;      // main.cpp:293, 368, 380
    for (int rep = 0; rep < R; rep++) {
        const uint64_t seed = init_seed + (uint64_t)rep;
        uint32_t ctr = 0;
        auto U = [&](uint32_t mol) { return keyed_uniform(seed, mol, ctr++, 0, SLOT_INIT); };
        std::vector<double> ax(NA), ay(NA), bx(NB), by(NB), bz(NB);
        HostGrid ga(P.box[0], P.box[1], exRR, NA), gb(P.box[0], P.box[1], exLL, NB), gab(P.box[0], P.box[1], exRL, NA);
        for (int a = 0; a < NA; a++) {
            for (int tries = 0;; tries++) {
                if (tries > 100000) return "kmc_init_random: cannot place receptors (box too dense)";
                double x = U(a + 1) * P.box[0] - P.box[0] / 2, y = U(a + 1) * P.box[1] - P.box[1] / 2;
                bool clash = ga.any_near(x, y, [&](int j) { double dx = x - ax[j], dy = y - ay[j]; return sqrt(dx * dx + dy * dy) <= exRR; });
                if (!clash) { ax[a] = x; ay[a] = y; ga.put(x, y, a); gab.put(x, y, a); break; }
            }
        }
        for (int b = 0; b < NB; b++) {
            for (int tries = 0;; tries++) {
                if (tries > 100000) return "kmc_init_random: cannot place ligands (box too dense)";
                double x = U(NA + b + 1) * P.box[0] - P.box[0] / 2, y = U(NA + b + 1) * P.box[1] - P.box[1] / 2, z = U(NA + b + 1) * P.box[2];
                bool clash = gab.any_near(x, y, [&](int j) {
                    for (int k = 0; k < 4; k++) { double dx = x - ax[j], dy = y - ay[j], dz = z - 2 * k * P.rA; if (sqrt(dx * dx + dy * dy + dz * dz) <= exRL) return true; }
                    return false; });
                clash = clash || gb.any_near(x, y, [&](int j) { double dx = x - bx[j], dy = y - by[j], dz = z - bz[j]; return sqrt(dx * dx + dy * dy + dz * dz) <= exLL; });
                if (!clash) { bx[b] = x; by[b] = y; bz[b] = z; gb.put(x, y, b); break; }
            }
        }
        // optional cell-major renumbering (memory locality of neighbour gathers on large membranes)
        std::vector<int> oa(NA), ob(NB);
        for (int i = 0; i < NA; i++) oa[i] = i;
        for (int i = 0; i < NB; i++) ob[i] = i;
        if (sort_cells) {
            auto key = [&](double x, double y) {
                long cx = (long)floor((x - K.gx0) * K.cellInv), cy = (long)floor((y - K.gy0) * K.cellInv);
                return cy * (long)K.ncx + cx; };
            std::stable_sort(oa.begin(), oa.end(), [&](int p, int q) { return key(ax[p], ay[p]) < key(ax[q], ay[q]); });
            std::stable_sort(ob.begin(), ob.end(), [&](int p, int q) { return key(bx[p], by[p]) < key(bx[q], by[q]); });
        }
        for (int a = 0; a < NA; a++) {
            const int src = oa[a];
            const double psai = (2 * U(a + 1) - 1) * P.pai, x = ax[src], y = ay[src];
            double *o = recOut + ((size_t)rep * NA + a) * 6;
            o[0] = x; o[1] = y;
            o[2] = cos(psai) * P.rA + x; o[3] = sin(psai) * P.rA + y;
            o[4] = cos(psai) * (-P.rA) + x; o[5] = sin(psai) * (-P.rA) + y;
        }
        for (int b = 0; b < NB; b++) {
            const int src = ob[b];
            const double th = (2 * U(NA + b + 1) - 1) * P.pai, ph = (2 * U(NA + b + 1) - 1) * P.pai, ps = (2 * U(NA + b + 1) - 1) * P.pai;
            const double t[3][3] = {{cos(ps) * cos(ph) - cos(th) * sin(ph) * sin(ps), -sin(ps) * cos(ph) - cos(th) * sin(ph) * cos(ps), sin(th) * sin(ph)},
                                    {cos(ps) * sin(ph) + cos(th) * cos(ph) * sin(ps), -sin(ps) * sin(ph) + cos(th) * cos(ph) * cos(ps), -sin(th) * cos(ph)},
                                    {sin(ps) * sin(th), cos(ps) * sin(th), cos(th)}};
            const double rB = P.rB;
            const double tpl[8][3] = {{0, 0, 0}, {0, rB * 2 / sqrt(3), 0}, {-rB, -rB / sqrt(3), 0}, {rB, -rB / sqrt(3), 0}, {0, 0, rB},
                                      {0, rB * (2 / sqrt(3) + 1), 0}, {-rB * (sqrt(3) / 2 + 1), -rB / sqrt(3) - rB / 2, 0},
                                      {rB * (sqrt(3) / 2 + 1), -rB / sqrt(3) - rB / 2, 0}};
            double *o = ligOut + ((size_t)rep * NB + b) * 24;
            const double c[3] = {bx[src], by[src], bz[src]};
            for (int q = 0; q < 8; q++)
                for (int d = 0; d < 3; d++) o[q * 3 + d] = t[d][0] * tpl[q][0] + t[d][1] * tpl[q][1] + t[d][2] * tpl[q][2] + c[d];
        }
    }
    return nullptr;
}

#include "kmc_init.cu"

// the configuration is generated ON THE DEVICE (csrc/kmc_init.cu: the reference's sequential insertion as a parallel fixed point)
// straight into the handle's arrays; nothing crosses PCIe
extern "C" int kmc_init_random(kmc_handle *h, uint64_t init_seed, int32_t sort_cells) {
    if (!h) return KMC_ERR_INVALID;
    int rc = select_device(h); if (rc) return rc;
    rc = stage_alloc(h); if (rc) return rc;
    Dev &D = h->D; cudaStream_t st = h->stream;
    CK(cudaStreamSynchronize(st));
    if (const char *msg = generate_random_device(h->P, h->NA, h->NB, h->R, init_seed, sort_cells, h->stageRec, D.lig, st, &h->init_rounds)) { h->err = msg; return KMC_ERR_INVALID; }
    int *dflag = h->stageInt + 3 * (size_t)h->NAt;
    CK(cudaMemsetAsync(D.ligRec, 0xff, sizeof(int) * 3 * (size_t)h->NBt, st));
    CK(cudaMemsetAsync(dflag, 0, sizeof(int), st));
    const Args A{D, h->K};
    LAUNCH(KID_SERIES, (k_pack_set<<<nblk(std::max(h->NAt, 1), 256), 256, 0, st>>>(A, h->stageRec, nullptr, nullptr, nullptr, dflag)));
    CK(cudaMemsetAsync(D.step64, 0, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(D.maxComplex, 0, sizeof(int) * h->R, st));
    int one = 1;
    CK(cudaMemcpyAsync(D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    h->step_done = 0; h->stepped = false; h->sinceBuild = 0;
    return KMC_OK;
}

// ------------------------------------------------------------------------------------------------
// the reference's remaining text files: pure host formatting on reference-shaped arrays
// ------------------------------------------------------------------------------------------------
// test.gro frame, main.cpp:2258-2287 (fixed, precision 3; coordinates in nm)
extern "C" int kmc_gro_append_arrays(const char *path, int32_t NA, int32_t NB, const double *Rx, const double *Ry, const double *Rz,
                                     double dt, int64_t step, const double *box) {
    if (!path || !Rx || !Ry || !Rz || !box) return KMC_ERR_INVALID;
    FILE *f = fopen(path, "a");
    if (!f) return KMC_ERR_IO;
    fprintf(f, "Hello Gro!, t=%.3f\n%d\n", (double)step * dt, NA * 4 + NB * 3);
    for (int i = 1; i <= NA; i++)
        for (int j = 1; j <= 4; j++)
            fprintf(f, "%5dALA%7s%5d%8.3f%8.3f%8.3f\n", i, "CA", i, Rx[RI(i, j, 1)] / 10, Ry[RI(i, j, 1)] / 10, Rz[RI(i, j, 1)] / 10);
    for (int i = NA + 1; i <= NA + NB; i++)
        for (int j = 2; j <= 4; j++)
            fprintf(f, "%5dLEU%7s%5d%8.3f%8.3f%8.3f\n", i, "CA", i, Rx[RI(i, j, 1)] / 10, Ry[RI(i, j, 1)] / 10, Rz[RI(i, j, 1)] / 10);
    fprintf(f, "%8.3f%12.3f%12.3f\n", box[0] / 10, box[1] / 10, box[2] / 10);
    fclose(f);
    return KMC_OK;
}
// position.cpt, main.cpp:2206-2244 (truncates; fixed, precision 3: a restart from it is lossy by design)
// counters[6] = bond_num, bond_num_rl, bond_num_cis, bond_num_mono_cis, protein_num_in_Max_Complex, mc_time_step
extern "C" int kmc_checkpoint_write_arrays(const char *path, int32_t NA, int32_t NB, const double *Rx, const double *Ry, const double *Rz,
                                           const int32_t *status, const int32_t *res_nei, const int64_t *counters) {
    if (!path || !Rx || !Ry || !Rz || !status || !res_nei || !counters) return KMC_ERR_INVALID;
    FILE *f = fopen(path, "w");
    if (!f) return KMC_ERR_IO;
    for (int i = 1; i <= NA; i++) {
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++) fprintf(f, "%10.3f%10.3f%10.3f\n", Rx[RI(i, j, k)], Ry[RI(i, j, k)], Rz[RI(i, j, k)]);
        fprintf(f, "%8d%8d%8d%8d%8d\n", status[i * 5 + 2], status[i * 5 + 3], res_nei[i * 7 + 2], res_nei[i * 7 + 4], res_nei[i * 7 + 3]);
    }
    for (int i = NA + 1; i <= NA + NB; i++)
        for (int j = 1; j <= 4; j++) {
            for (int k = 1; k <= 2; k++) fprintf(f, "%10.3f%10.3f%10.3f\n", Rx[RI(i, j, k)], Ry[RI(i, j, k)], Rz[RI(i, j, k)]);
            fprintf(f, "%8d%8d\n", status[i * 5 + j], res_nei[i * 7 + j]);
        }
    for (int q = 0; q < 6; q++) fprintf(f, "%lld\n", (long long)counters[q]);
    fclose(f);
    return KMC_OK;
}
// reader of the same token stream, main.cpp:231-266 (any precision is parsed exactly)
extern "C" int kmc_checkpoint_read_arrays(const char *path, int32_t NA, int32_t NB, double *Rx, double *Ry, double *Rz, int32_t *status,
                                          int32_t *res_nei, int64_t *counters) {
    if (!path || !Rx || !Ry || !Rz || !status || !res_nei || !counters) return KMC_ERR_INVALID;
    FILE *f = fopen(path, "r");
    if (!f) return KMC_ERR_IO;
    const int N = NA + NB;
    memset(Rx, 0, sizeof(double) * 25 * (N + 1)); memset(Ry, 0, sizeof(double) * 25 * (N + 1)); memset(Rz, 0, sizeof(double) * 25 * (N + 1));
    memset(status, 0, sizeof(int32_t) * 5 * (N + 1)); memset(res_nei, 0, sizeof(int32_t) * 7 * (N + 1));
    bool ok = true;
    auto D = [&](double &v) { char tok[128]; if (fscanf(f, "%127s", tok) != 1) { ok = false; v = 0; return; } v = strtod(tok, nullptr); };
    auto I = [&](int32_t &v) { long long t = 0; if (fscanf(f, "%lld", &t) != 1) ok = false; v = (int32_t)t; };
    for (int i = 1; i <= NA; i++) {
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++) { D(Rx[RI(i, j, k)]); D(Ry[RI(i, j, k)]); D(Rz[RI(i, j, k)]); }
        I(status[i * 5 + 2]); I(status[i * 5 + 3]); I(res_nei[i * 7 + 2]); I(res_nei[i * 7 + 4]); I(res_nei[i * 7 + 3]);
    }
    for (int i = NA + 1; i <= N; i++)
        for (int j = 1; j <= 4; j++) {
            for (int k = 1; k <= 2; k++) { D(Rx[RI(i, j, k)]); D(Ry[RI(i, j, k)]); D(Rz[RI(i, j, k)]); }
            I(status[i * 5 + j]); I(res_nei[i * 7 + j]);
        }
    for (int q = 0; q < 6; q++) { long long t = 0; if (fscanf(f, "%lld", &t) != 1) ok = false; counters[q] = t; }
    fclose(f);
    return ok ? KMC_OK : KMC_ERR_IO;
}
// parameter.log, main.cpp:179-205 (default ostream float format = %g, widths 25/15)
extern "C" int kmc_parameter_log_write(const kmc_params *p, const char *path) {
    if (!p || !path) return KMC_ERR_INVALID;
    FILE *f = fopen(path, "w");
    if (!f) return KMC_ERR_IO;
    fprintf(f, "%25s%15g%7g%7g\n\n", "box size: x y z", p->box[0], p->box[1], p->box[2]);
    fprintf(f, "%25s%15d\n%25s%15d\n%25s%15d\n%25s%15d\n\n", "protein_A_tot_num", p->n_receptor, "RB_A_tot_num", p->n_receptor * 4,
            "protein_B_tot_num", p->n_ligand, "RB_B_tot_num", p->n_ligand * 4);
    fprintf(f, "%25s%15g\n%25s%15g\n%25s%15g\n%25s%15g\n\n", "RB_A_D", p->DA, "RB_A_rot_D", p->DrotA, "RB_B_D", p->DB, "RB_B_rot_D", p->DrotB);
    fprintf(f, "%25s\n%25s%15g\n%25s%15g\n%25s%15g\n%25s%15g\n\n", "R-L interaction:", "bond_D", p->bond_D, "bond_rot_D", p->bond_Drot, "Ass_Rate", p->on,
            "Diss_Rate", p->off);
    fprintf(f, "%25s\n%25s%15g\n%25s%15g\n%25s%15g\n%25s%15g\n\n%25s%15g\n%25s%15g\n\n", "Cis interaction:", "cis_D", p->cis_D, "cis_rot_D", p->cis_Drot,
            "mono_cis_Ass_Rate", p->mono_cis_on, "mono_cis_Diss_Rate", p->mono_cis_off, "cis_Ass_Rate", p->cis_on, "cis_Diss_Rate", p->cis_off);
    fclose(f);
    return KMC_OK;
}
// handle-level conveniences: state of one replica -> the reference's files, and restart from a position.cpt
extern "C" int kmc_write_gro(kmc_handle *h, int32_t rep, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    const size_t n = (size_t)(h->N + 1);
    std::vector<double> X(n * 25), Y(n * 25), Z(n * 25); std::vector<int32_t> st(n * 5), rn(n * 7);
    int rc = kmc_get_state(h, rep, X.data(), Y.data(), Z.data(), st.data(), rn.data()); if (rc) return rc;
    rc = kmc_gro_append_arrays(path, h->NA, h->NB, X.data(), Y.data(), Z.data(), h->P.dt, h->step_done, h->P.box);
    if (rc) h->err = std::string("cannot write ") + path;
    return rc;
}
extern "C" int kmc_write_checkpoint(kmc_handle *h, int32_t rep, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    const size_t n = (size_t)(h->N + 1);
    std::vector<double> X(n * 25), Y(n * 25), Z(n * 25); std::vector<int32_t> st(n * 5), rn(n * 7);
    int rc = kmc_get_state(h, rep, X.data(), Y.data(), Z.data(), st.data(), rn.data()); if (rc) return rc;
    kmc_series s; rc = kmc_get_series(h, rep, &s); if (rc) return rc;
    const int64_t c[6] = {s.bond_num, s.bond_num_rl, s.bond_num_cis, s.bond_num_mono_cis, s.max_complex, s.step};
    rc = kmc_checkpoint_write_arrays(path, h->NA, h->NB, X.data(), Y.data(), Z.data(), st.data(), rn.data(), c);
    if (rc) h->err = std::string("cannot write ") + path;
    return rc;
}
extern "C" int kmc_read_checkpoint(kmc_handle *h, int32_t rep, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    const size_t n = (size_t)(h->N + 1);
    std::vector<double> X(n * 25), Y(n * 25), Z(n * 25); std::vector<int32_t> st(n * 5), rn(n * 7); int64_t c[6];
    int rc = kmc_checkpoint_read_arrays(path, h->NA, h->NB, X.data(), Y.data(), Z.data(), st.data(), rn.data(), c);
    if (rc) { h->err = std::string("cannot read ") + path; return rc; }
    return kmc_set_state(h, rep, X.data(), Y.data(), Z.data(), st.data(), rn.data(), c[5], (int32_t)c[4]);    // main.cpp:261-267
}

// Lossless checkpoint: position.cpt keeps 3 decimals (main.cpp:2206-2244), so a restart from it is not
// bit-continuable. This one stores the whole handle (all replicas) in binary: header, running-max complex sizes, then the
// arrays of kmc_get_packed. A run restored from it continues bit for bit (the random stream is keyed, not stateful).
namespace { struct BinHeader { char magic[8]; int32_t n_receptor, n_ligand, n_replicas, reserved; int64_t step_done; }; }
extern "C" int kmc_write_checkpoint_bin(kmc_handle *h, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    std::vector<double> rec((size_t)h->NAt * 6), lig((size_t)h->NBt * 24);
    std::vector<int32_t> rl(h->NAt), rs(h->NAt), rc_(h->NAt), mx(h->R);
    int rc = kmc_get_packed(h, rec.data(), lig.data(), rl.data(), rs.data(), rc_.data()); if (rc) return rc;
    CK(cudaMemcpy(mx.data(), h->D.maxComplex, sizeof(int) * h->R, cudaMemcpyDeviceToHost));
    BinHeader H; memcpy(H.magic, "KMCB2001", 8); H.n_receptor = h->NA; H.n_ligand = h->NB; H.n_replicas = h->R; H.reserved = 0; H.step_done = h->step_done;
    FILE *f = fopen(path, "wb");
    bool ok = f != nullptr;
    ok = ok && fwrite(&H, sizeof H, 1, f) == 1 && fwrite(mx.data(), sizeof(int32_t), mx.size(), f) == mx.size();
    ok = ok && fwrite(rec.data(), sizeof(double), rec.size(), f) == rec.size() && fwrite(lig.data(), sizeof(double), lig.size(), f) == lig.size();
    ok = ok && fwrite(rl.data(), 4, rl.size(), f) == rl.size() && fwrite(rs.data(), 4, rs.size(), f) == rs.size() && fwrite(rc_.data(), 4, rc_.size(), f) == rc_.size();
    if (f) ok = (fclose(f) == 0) && ok;
    if (!ok) { h->err = std::string("cannot write ") + path; return KMC_ERR_IO; }
    return KMC_OK;
}
extern "C" int kmc_read_checkpoint_bin(kmc_handle *h, const char *path) {
    if (!h || !path) return KMC_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) { h->err = std::string("cannot read ") + path; return KMC_ERR_IO; }
    BinHeader H;
    bool ok = fread(&H, sizeof H, 1, f) == 1 && memcmp(H.magic, "KMCB2001", 8) == 0;
    if (ok && (H.n_receptor != h->NA || H.n_ligand != h->NB || H.n_replicas != h->R)) {
        fclose(f); h->err = "kmc_read_checkpoint_bin: the file holds " + std::to_string(H.n_receptor) + "+" + std::to_string(H.n_ligand) + " molecules x " +
                            std::to_string(H.n_replicas) + " replicas, the handle " + std::to_string(h->NA) + "+" + std::to_string(h->NB) + " x " + std::to_string(h->R);
        return KMC_ERR_STATE;
    }
    std::vector<double> rec((size_t)h->NAt * 6), lig((size_t)h->NBt * 24);
    std::vector<int32_t> rl(h->NAt), rs(h->NAt), rc_(h->NAt), mx(h->R);
    ok = ok && fread(mx.data(), sizeof(int32_t), mx.size(), f) == mx.size();
    ok = ok && fread(rec.data(), sizeof(double), rec.size(), f) == rec.size() && fread(lig.data(), sizeof(double), lig.size(), f) == lig.size();
    ok = ok && fread(rl.data(), 4, rl.size(), f) == rl.size() && fread(rs.data(), 4, rs.size(), f) == rs.size() && fread(rc_.data(), 4, rc_.size(), f) == rc_.size();
    fclose(f);
    if (!ok) { h->err = std::string("not a complete KMCB2001 checkpoint: ") + path; return KMC_ERR_IO; }
    int rc = kmc_set_packed(h, rec.data(), lig.data(), rl.data(), rs.data(), rc_.data(), H.step_done); if (rc) return rc;
    CK(cudaMemcpy(h->D.maxComplex, mx.data(), sizeof(int) * h->R, cudaMemcpyHostToDevice));
    return KMC_OK;
}

// the same generator without a handle or a device: host arrays out (strips: every rank generates the global start state)
extern "C" int kmc_generate_packed(const kmc_params *p, uint64_t init_seed, int32_t sort_cells, double *rec_pose, double *lig_pose) {
    if (!p || !rec_pose || !lig_pose || p->n_replicas < 1) return KMC_ERR_INVALID;
    Consts K; fill_consts(*p, K);
    return generate_random(*p, K, init_seed, sort_cells, rec_pose, lig_pose) ? KMC_ERR_INVALID : KMC_OK;
}

#include "kmc_strips.cu"
