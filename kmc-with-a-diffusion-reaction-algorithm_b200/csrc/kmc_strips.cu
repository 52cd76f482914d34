// csrc/kmc_strips.cu -- strip domain decomposition of ONE membrane across GPUs (included by kmc_engine.cu).
//
// The reference has no decomposition (single thread); this is new. Rank r of n owns the molecules of the
// units (free molecule, cis dimer, ligand-rooted complex) whose head molecule has its committed centre x in
// [-Lx/2 + r*Lx/n, -Lx/2 + (r+1)*Lx/n). Around its strip every rank also keeps HALO copies of the neighbours' units (whole
// units, width W) and simply simulates them too: the sweep is deterministic in (state, keyed draws) -- order keys and Philox
// keys use the reference (global) molecule ids, local arrays are kept sorted by that id -- so a halo copy evolves bit-identically
// to the owner's original as long as everything that can influence it is present locally. Influence travels at most one
// interaction range per step, so after k steps the outer k*D1 of the halo may be stale while the owned strip is still exact;
// every k steps the ranks REFRESH: ownership is re-derived from the current positions, owners send the units within W of each
// boundary to that neighbour (NCCL send/recv in the caller, kmc_b200/strips.py), halos are replaced.
// The periodic seam: the reference applies no minimum image (main.cpp:642-646), so nothing interacts across x = +-Lx/2 until a
// molecule is wrapped (main.cpp:597-605) -- the band beyond the seam is the halo of the first/last strip like any other band,
// kept at its true coordinates and hashed into the cell grid in the periodic frame of the strip (hash_x).
//
// Round-1 implementation: classification, packing and merging run on the HOST at refresh time (D2H, rebuild, H2D); the
// per-step path is untouched device code. Moving the refresh onto the device is the listed next step.
static double strip_hash_x(const kmc_handle *h, double x) {
    const double t = x - h->K.stripXc;
    return t > h->K.stripHalf ? x - h->K.Lx : (t < -h->K.stripHalf ? x + h->K.Lx : x);
}
static int strip_owner(const kmc_handle *h, double x) {
    const double L = h->K.Lx, xw = x - L * round(x / L);
    int r = (int)floor((xw + L / 2) / (L / h->K.strips));
    return std::min(std::max(r, 0), h->K.strips - 1);
}

static int strip_download(kmc_handle *h, HostLocal &s) {
    Dev &D = h->D;
    CK(cudaStreamSynchronize(h->stream));
    int live[2];
    CK(cudaMemcpy(live, D.scal + S_NA_LIVE, sizeof live, cudaMemcpyDeviceToHost));
    s.nA = live[0]; s.nB = live[1];
    std::vector<double2> c(s.nA), s2(s.nA), s3(s.nA);
    s.rec.resize((size_t)s.nA * 6); s.lig.resize((size_t)s.nB * 24); s.rl.resize(s.nA); s.rs.resize(s.nA); s.rc.resize(s.nA);
    s.lr.resize((size_t)s.nB * 3); s.refA.resize(s.nA); s.refB.resize(s.nB);
    if (s.nA) {
        CK(cudaMemcpy(c.data(), D.recC, sizeof(double2) * s.nA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s2.data(), D.recS2, sizeof(double2) * s.nA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s3.data(), D.recS3, sizeof(double2) * s.nA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s.rl.data(), D.recLig, sizeof(int) * s.nA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s.rs.data(), D.recSite, sizeof(int) * s.nA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s.rc.data(), D.recCis, sizeof(int) * s.nA, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s.refA.data(), D.refA, sizeof(unsigned) * s.nA, cudaMemcpyDeviceToHost));
    }
    if (s.nB) {
        CK(cudaMemcpy(s.lig.data(), D.lig, sizeof(double) * 24 * (size_t)s.nB, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s.lr.data(), D.ligRec, sizeof(int) * 3 * (size_t)s.nB, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(s.refB.data(), D.refB, sizeof(unsigned) * s.nB, cudaMemcpyDeviceToHost));
    }
    for (int a = 0; a < s.nA; a++) {
        double *o = &s.rec[(size_t)a * 6];
        o[0] = c[a].x; o[1] = c[a].y; o[2] = s2[a].x; o[3] = s2[a].y; o[4] = s3[a].x; o[5] = s3[a].y;
    }
    return KMC_OK;
}

static int strip_upload(kmc_handle *h, const HostLocal &s) {
    Dev &D = h->D;
    if (s.nA > h->NAt || s.nB > h->NBt) {
        h->err = "strip: local capacity exceeded (" + std::to_string(s.nA) + "/" + std::to_string(h->NAt) + " receptors, " + std::to_string(s.nB) + "/" +
                 std::to_string(h->NBt) + " ligands): create the handle with larger n_receptor/n_ligand"; return KMC_ERR_CAPACITY;
    }
    std::vector<double2> c(s.nA), s2(s.nA), s3(s.nA);
    for (int a = 0; a < s.nA; a++) {
        const double *o = &s.rec[(size_t)a * 6];
        c[a] = make_double2(o[0], o[1]); s2[a] = make_double2(o[2], o[3]); s3[a] = make_double2(o[4], o[5]);
    }
    CK(cudaStreamSynchronize(h->stream));
    if (s.nA) {
        CK(cudaMemcpy(D.recC, c.data(), sizeof(double2) * s.nA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.recS2, s2.data(), sizeof(double2) * s.nA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.recS3, s3.data(), sizeof(double2) * s.nA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.recLig, s.rl.data(), sizeof(int) * s.nA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.recSite, s.rs.data(), sizeof(int) * s.nA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.recCis, s.rc.data(), sizeof(int) * s.nA, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.refA, s.refA.data(), sizeof(unsigned) * s.nA, cudaMemcpyHostToDevice));
    }
    if (s.nB) {
        CK(cudaMemcpy(D.lig, s.lig.data(), sizeof(double) * 24 * (size_t)s.nB, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.ligRec, s.lr.data(), sizeof(int) * 3 * (size_t)s.nB, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.refB, s.refB.data(), sizeof(unsigned) * s.nB, cudaMemcpyHostToDevice));
    }
    int live[2] = {s.nA, s.nB}, one = 1;
    CK(cudaMemcpy(D.scal + S_NA_LIVE, live, sizeof live, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice));
    h->stepped = false;
    return KMC_OK;
}

// units of a local (or global) state: unit id = index of the head in a combined numbering (ligand h -> h, receptor a -> nB + a),
// head = lowest ligand of the component, or lowest receptor of a ligand-free one (same rule as k_uf_*: main.cpp:525, 682-688)
static void host_units(const HostLocal &s, std::vector<int> &unitOfRec, std::vector<int> &unitOfLig) {
    const int n = s.nA + s.nB;
    std::vector<int> parent(n);
    for (int i = 0; i < n; i++) parent[i] = i;
    auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto uni = [&](int a, int b) { a = find(a); b = find(b); if (a == b) return; if (a > b) std::swap(a, b); parent[b] = a; };
    for (int a = 0; a < s.nA; a++) {
        if (s.rl[a] >= 0) uni(s.nB + a, s.rl[a]);
        if (s.rc[a] > a) uni(s.nB + a, s.nB + s.rc[a]);
    }
    unitOfRec.resize(s.nA); unitOfLig.resize(s.nB);
    for (int a = 0; a < s.nA; a++) unitOfRec[a] = find(s.nB + a);
    for (int b = 0; b < s.nB; b++) unitOfLig[b] = find(b);
}
static double head_x(const HostLocal &s, int unit) { return unit < s.nB ? s.lig[(size_t)unit * 24] : s.rec[(size_t)(unit - s.nB) * 6]; }

static void append_rec(std::vector<char> &buf, const HostLocal &s, int a) {
    RecMsg m; m.ref = (int32_t)s.refA[a]; m.ligRef = s.rl[a] >= 0 ? (int32_t)s.refB[s.rl[a]] : 0; m.site = s.rs[a];
    m.cisRef = s.rc[a] >= 0 ? (int32_t)s.refA[s.rc[a]] : 0;
    memcpy(m.pose, &s.rec[(size_t)a * 6], sizeof m.pose);
    buf.insert(buf.end(), (const char *)&m, (const char *)&m + sizeof m);
}
static void append_lig(std::vector<char> &buf, const HostLocal &s, int b) {
    LigMsg m; m.ref = (int32_t)s.refB[b];
    for (int k = 0; k < 3; k++) m.recRef[k] = s.lr[(size_t)b * 3 + k] >= 0 ? (int32_t)s.refA[s.lr[(size_t)b * 3 + k]] : 0;
    memcpy(m.pose, &s.lig[(size_t)b * 24], sizeof m.pose);
    buf.insert(buf.end(), (const char *)&m, (const char *)&m + sizeof m);
}
// message = int64 nRec, int64 nLig, RecMsg[nRec], LigMsg[nLig]   (both sorted by ref)
static std::vector<char> make_msg(const std::vector<char> &recs, const std::vector<char> &ligs) {
    std::vector<char> out(16 + recs.size() + ligs.size());
    int64_t n[2] = {(int64_t)(recs.size() / sizeof(RecMsg)), (int64_t)(ligs.size() / sizeof(LigMsg))};
    memcpy(out.data(), n, 16);
    if (!recs.empty()) memcpy(out.data() + 16, recs.data(), recs.size());
    if (!ligs.empty()) memcpy(out.data() + 16 + recs.size(), ligs.data(), ligs.size());
    return out;
}

extern "C" int kmc_strip_configure(kmc_handle *h, int32_t rank, int32_t nranks, double halo_width) {
    if (!h) return KMC_ERR_INVALID;
    if (h->R != 1 || nranks < 1 || rank < 0 || rank >= nranks || halo_width <= 0) { h->err = "kmc_strip_configure: bad arguments (strips need n_replicas = 1)"; return KMC_ERR_INVALID; }
    const double width = h->K.Lx / nranks;
    if (nranks > 1 && 2 * halo_width >= width) { h->err = "kmc_strip_configure: halo must be narrower than half a strip"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    Consts &K = h->K; Dev &D = h->D;
    K.strips = nranks; K.stripRank = rank; K.stripXc = -K.Lx / 2 + (rank + 0.5) * width; K.stripHalf = nranks > 1 ? K.Lx / 2 : INFINITY;
    h->strip_on = true; h->strip_W = halo_width; h->strip_lo = -K.Lx / 2 + rank * width; h->strip_hi = h->strip_lo + width;
    if (nranks > 1) {        // the grid only has to cover the strip and its halos (in the periodic frame of the strip)
        const double edge = 1.0 / K.cellInv;
        K.gx0 = h->strip_lo - halo_width - 2 * edge;
        K.ncx = (int)ceil((width + 2 * halo_width + 4 * edge) / edge);
        const int ncell = K.ncx * K.ncy;
        h->scanBlocks = (ncell + 1 + SCAN_TILE - 1) / SCAN_TILE;
        if (ncell > D.ncell) {
            bool ok = dalloc(h, &D.cellCount, (size_t)h->scanBlocks * SCAN_TILE) == cudaSuccess && dalloc(h, &D.cellStart, (size_t)h->scanBlocks * SCAN_TILE) == cudaSuccess &&
                      dalloc(h, &D.scanTmp, (size_t)h->scanBlocks + 1) == cudaSuccess;
            if (!ok) { h->err = "kmc_strip_configure: grid allocation failed"; return KMC_ERR_CUDA; }
        }
        D.ncell = ncell;
        h->nTiles = ((K.ncx + TS - 1) / TS) * ((K.ncy + TS - 1) / TS);
    }
    if (!D.refA) {
        if (dalloc(h, &D.refA, std::max(h->NAt, 1)) != cudaSuccess || dalloc(h, &D.refB, std::max(h->NBt, 1)) != cudaSuccess) { h->err = "kmc_strip_configure: allocation failed"; return KMC_ERR_CUDA; }
    }
    int zero[2] = {0, 0};
    CK(cudaMemcpy(D.scal + S_NA_LIVE, zero, sizeof zero, cudaMemcpyHostToDevice));
    for (int p = 0; p < 2; p++) if (h->gexec[p]) { cudaGraphExecDestroy(h->gexec[p]); h->gexec[p] = nullptr; }
    return KMC_OK;
}

// is any part of the strip's neighbourhood [lo - W, hi + W) (periodic frame of the strip) touched by this x?
static bool in_reach(const kmc_handle *h, double x) { const double t = strip_hash_x(h, x); return t >= h->strip_lo - h->strip_W && t < h->strip_hi + h->strip_W; }

// Distribute a GLOBAL state (every rank passes the same arrays, kmc_get_packed layout; receptor a has reference id a+1, ligand
// b has n_rec + b + 1): this rank keeps the units it owns plus every other unit with a member within the halo width.
extern "C" int kmc_strip_load_global(kmc_handle *h, int32_t n_rec, int32_t n_lig, const double *rec_pose, const double *lig_pose,
                                     const int32_t *rec_lig, const int32_t *rec_site, const int32_t *rec_cis, int64_t step_done) {
    if (!h || !h->strip_on || !rec_pose || !lig_pose) { if (h) h->err = "kmc_strip_load_global: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    HostLocal g; g.nA = n_rec; g.nB = n_lig;
    g.rec.assign(rec_pose, rec_pose + (size_t)n_rec * 6); g.lig.assign(lig_pose, lig_pose + (size_t)n_lig * 24);
    g.rl.assign(n_rec, -1); g.rs.assign(n_rec, -1); g.rc.assign(n_rec, -1); g.lr.assign((size_t)n_lig * 3, -1);
    for (int a = 0; a < n_rec; a++) {
        if (rec_lig && rec_lig[a] >= 0) { g.rl[a] = rec_lig[a]; g.rs[a] = rec_site[a] - 2; g.lr[(size_t)rec_lig[a] * 3 + rec_site[a] - 2] = a; }
        if (rec_cis && rec_cis[a] >= 0) g.rc[a] = rec_cis[a];
    }
    g.refA.resize(n_rec); g.refB.resize(n_lig);
    for (int a = 0; a < n_rec; a++) g.refA[a] = a + 1;
    for (int b = 0; b < n_lig; b++) g.refB[b] = n_rec + b + 1;
    std::vector<int> uR, uL; host_units(g, uR, uL);
    // per unit: mine? touches my neighbourhood?
    std::vector<char> take(n_rec + n_lig, 0);
    for (int a = 0; a < n_rec; a++) if (in_reach(h, g.rec[(size_t)a * 6])) take[uR[a]] = 1;
    for (int b = 0; b < n_lig; b++) if (in_reach(h, g.lig[(size_t)b * 24])) take[uL[b]] = 1;
    for (int u = 0; u < n_rec + n_lig; u++) if (strip_owner(h, head_x(g, u)) == h->K.stripRank) take[u] |= 2;     // (only heads matter)
    HostLocal s;
    std::vector<int> mapA(n_rec, -1), mapB(n_lig, -1);
    for (int a = 0; a < n_rec; a++) if (take[uR[a]] & 1 || (take[uR[a]] & 2)) mapA[a] = s.nA++;
    for (int b = 0; b < n_lig; b++) if (take[uL[b]] & 1 || (take[uL[b]] & 2)) mapB[b] = s.nB++;
    s.rec.resize((size_t)s.nA * 6); s.lig.resize((size_t)s.nB * 24); s.rl.assign(s.nA, -1); s.rs.assign(s.nA, -1); s.rc.assign(s.nA, -1);
    s.lr.assign((size_t)s.nB * 3, -1); s.refA.resize(s.nA); s.refB.resize(s.nB);
    for (int a = 0; a < n_rec; a++) {
        const int q = mapA[a]; if (q < 0) continue;
        memcpy(&s.rec[(size_t)q * 6], &g.rec[(size_t)a * 6], 48); s.refA[q] = g.refA[a];
        if (g.rl[a] >= 0) { s.rl[q] = mapB[g.rl[a]]; s.rs[q] = g.rs[a]; }
        if (g.rc[a] >= 0) s.rc[q] = mapA[g.rc[a]];
    }
    for (int b = 0; b < n_lig; b++) {
        const int q = mapB[b]; if (q < 0) continue;
        memcpy(&s.lig[(size_t)q * 24], &g.lig[(size_t)b * 24], 192); s.refB[q] = g.refB[b];
        for (int k = 0; k < 3; k++) if (g.lr[(size_t)b * 3 + k] >= 0) s.lr[(size_t)q * 3 + k] = mapA[g.lr[(size_t)b * 3 + k]];
    }
    int rc = strip_upload(h, s); if (rc) return rc;
    h->step_done = step_done;
    unsigned long long s64 = (unsigned long long)step_done;
    CK(cudaMemcpy(h->D.step64, &s64, sizeof s64, cudaMemcpyHostToDevice));
    return KMC_OK;
}

// Refresh, part 1: re-derive ownership from the current positions and build the two messages for the neighbours
// (side 0 = towards lower x, side 1 = towards higher x; ranks 0 and n-1 are neighbours through the periodic seam).
extern "C" int kmc_strip_begin_refresh(kmc_handle *h) {
    if (!h || !h->strip_on) { if (h) h->err = "kmc_strip_begin_refresh: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    HostLocal &s = h->strip_local;
    int rc = strip_download(h, s); if (rc) return rc;
    std::vector<int> uR, uL; host_units(s, uR, uL);
    const int nU = s.nA + s.nB;
    std::vector<char> own(nU, 0), sl(nU, 0), sr(nU, 0);
    for (int u = 0; u < nU; u++) own[u] = strip_owner(h, head_x(s, u)) == h->K.stripRank;
    auto mark = [&](int u, double x) {
        if (!own[u]) return;
        const double t = strip_hash_x(h, x);
        if (t < h->strip_lo + h->strip_W) sl[u] = 1;
        if (t >= h->strip_hi - h->strip_W) sr[u] = 1;
    };
    for (int a = 0; a < s.nA; a++) mark(uR[a], s.rec[(size_t)a * 6]);
    for (int b = 0; b < s.nB; b++) mark(uL[b], s.lig[(size_t)b * 24]);
    std::vector<char> recs[3], ligs[3];      // 0 send left, 1 send right, 2 keep
    for (int a = 0; a < s.nA; a++) {
        const int u = uR[a];
        if (sl[u]) append_rec(recs[0], s, a);
        if (sr[u]) append_rec(recs[1], s, a);
        if (own[u]) append_rec(recs[2], s, a);
    }
    for (int b = 0; b < s.nB; b++) {
        const int u = uL[b];
        if (sl[u]) append_lig(ligs[0], s, b);
        if (sr[u]) append_lig(ligs[1], s, b);
        if (own[u]) append_lig(ligs[2], s, b);
    }
    for (int k = 0; k < 3; k++) h->strip_msg[k] = make_msg(recs[k], ligs[k]);
    return KMC_OK;
}
extern "C" int64_t kmc_strip_message(kmc_handle *h, int32_t side, const void **data) {
    if (!h || side < 0 || side > 2) return KMC_ERR_INVALID;
    if (data) *data = h->strip_msg[side].data();
    return (int64_t)h->strip_msg[side].size();
}

// Refresh, part 2: the new local set = the units this rank owns + what the two neighbours sent; sorted by reference id.
extern "C" int kmc_strip_rebuild(kmc_handle *h, const void *from_low, int64_t n_low, const void *from_high, int64_t n_high) {
    if (!h || !h->strip_on) { if (h) h->err = "kmc_strip_rebuild: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    const std::vector<char> &keep = h->strip_msg[2];
    const char *src[3] = {keep.data(), (const char *)from_low, (const char *)from_high};
    const int64_t len[3] = {(int64_t)keep.size(), n_low, n_high};
    std::vector<RecMsg> R; std::vector<LigMsg> L;
    for (int k = 0; k < 3; k++) {
        if (!src[k] || len[k] < 16) continue;
        int64_t n[2]; memcpy(n, src[k], 16);
        if (16 + n[0] * (int64_t)sizeof(RecMsg) + n[1] * (int64_t)sizeof(LigMsg) != len[k]) { h->err = "kmc_strip_rebuild: malformed message"; return KMC_ERR_INVALID; }
        const size_t r0 = R.size(), l0 = L.size();
        R.resize(r0 + n[0]); L.resize(l0 + n[1]);
        if (n[0]) memcpy(&R[r0], src[k] + 16, n[0] * sizeof(RecMsg));
        if (n[1]) memcpy(&L[l0], src[k] + 16 + n[0] * sizeof(RecMsg), n[1] * sizeof(LigMsg));
    }
    std::sort(R.begin(), R.end(), [](const RecMsg &a, const RecMsg &b) { return a.ref < b.ref; });
    std::sort(L.begin(), L.end(), [](const LigMsg &a, const LigMsg &b) { return a.ref < b.ref; });
    // a unit can arrive from both sides only when two ranks share both boundaries (n = 2): identical copies, keep one
    R.erase(std::unique(R.begin(), R.end(), [](const RecMsg &a, const RecMsg &b) { return a.ref == b.ref; }), R.end());
    L.erase(std::unique(L.begin(), L.end(), [](const LigMsg &a, const LigMsg &b) { return a.ref == b.ref; }), L.end());
    HostLocal s; s.nA = (int)R.size(); s.nB = (int)L.size();
    s.rec.resize((size_t)s.nA * 6); s.lig.resize((size_t)s.nB * 24); s.rl.assign(s.nA, -1); s.rs.assign(s.nA, -1); s.rc.assign(s.nA, -1);
    s.lr.assign((size_t)s.nB * 3, -1); s.refA.resize(s.nA); s.refB.resize(s.nB);
    for (int a = 0; a < s.nA; a++) s.refA[a] = (unsigned)R[a].ref;
    for (int b = 0; b < s.nB; b++) s.refB[b] = (unsigned)L[b].ref;
    auto findA = [&](int ref) { auto it = std::lower_bound(s.refA.begin(), s.refA.end(), (unsigned)ref); return (it != s.refA.end() && *it == (unsigned)ref) ? (int)(it - s.refA.begin()) : -1; };
    auto findB = [&](int ref) { auto it = std::lower_bound(s.refB.begin(), s.refB.end(), (unsigned)ref); return (it != s.refB.end() && *it == (unsigned)ref) ? (int)(it - s.refB.begin()) : -1; };
    for (int a = 0; a < s.nA; a++) {
        memcpy(&s.rec[(size_t)a * 6], R[a].pose, 48);
        if (R[a].ligRef) { s.rl[a] = findB(R[a].ligRef); s.rs[a] = R[a].site; if (s.rl[a] < 0) { h->err = "kmc_strip_rebuild: unit arrived incomplete (ligand missing)"; return KMC_ERR_STATE; } }
        if (R[a].cisRef) { s.rc[a] = findA(R[a].cisRef); if (s.rc[a] < 0) { h->err = "kmc_strip_rebuild: unit arrived incomplete (cis partner missing)"; return KMC_ERR_STATE; } }
    }
    for (int b = 0; b < s.nB; b++) {
        memcpy(&s.lig[(size_t)b * 24], L[b].pose, 192);
        for (int k = 0; k < 3; k++) if (L[b].recRef[k]) { s.lr[(size_t)b * 3 + k] = findA(L[b].recRef[k]); if (s.lr[(size_t)b * 3 + k] < 0) { h->err = "kmc_strip_rebuild: unit arrived incomplete (receptor missing)"; return KMC_ERR_STATE; } }
    }
    h->strip_refreshes++;
    return strip_upload(h, s);
}
