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
    h->stepped = false; h->sinceBuild = 0;
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
        choose_tiles(h);
        if (!ensure_cells_arrays(h)) { h->err = "kmc_strip_configure: allocation failed"; return KMC_ERR_CUDA; }
    }
    if (!D.refA) {
        if (dalloc(h, &D.refA, std::max(h->NAt, 1)) != cudaSuccess || dalloc(h, &D.refB, std::max(h->NBt, 1)) != cudaSuccess) { h->err = "kmc_strip_configure: allocation failed"; return KMC_ERR_CUDA; }
    }
    int zero[2] = {0, 0};
    CK(cudaMemcpy(D.scal + S_NA_LIVE, zero, sizeof zero, cudaMemcpyHostToDevice));
    for (int p = 0; p < 4; p++) if (h->gexec[p >> 1][p & 1]) { cudaGraphExecDestroy(h->gexec[p >> 1][p & 1]); h->gexec[p >> 1][p & 1] = nullptr; }
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

// ================================================================================================================
// Device-side refresh: the same three phases (classify + pack, exchange by the caller, merge) without the host round trip.
// Messages live in device buffers ([RecMsg x nRec][LigMsg x nLig], counts travel separately), so NCCL moves them GPU to GPU.
// The host path above is kept as the reference implementation (tests compare the two).
// ================================================================================================================
struct StripDev {
    unsigned char *flag = nullptr;            // [NT] bit0 owned, bit1 send to lower-x neighbour, bit2 send to higher-x neighbour
    int *tileCnt = nullptr, *tileOff = nullptr;   // per tile of ST_TILE molecules of one species: members of list 0 low, 1 high, 2 keep; their exclusive prefixes
    int *dcnt = nullptr, *hcnt = nullptr;         // totals [list*2 + species] on the device / in pinned host memory (+ [6] = overflow flag)
    int ntA = 0, ntB = 0;
    char *msg[3] = {nullptr}; size_t msgCap[3] = {0};     // packed messages (2 = the owned set)
    char *rcv[2] = {nullptr}; size_t rcvCap[2] = {0};     // what the neighbours sent (0 from lower x, 1 from higher x)
    int *bondRef = nullptr;                   // [NAt*3 + NBt*3] bonds of the merged molecules as reference ids, before translation
    int cnt[6] = {0};                         // totals of pos[]
    int padN = 0;
};

__device__ __forceinline__ int d_strip_owner(const Consts &K, double x) {
    const double L = K.Lx, xw = x - L * round(x / L);
    int r = (int)floor((xw + L / 2) / (L / K.strips));
    return min(max(r, 0), K.strips - 1);
}
__global__ void k_strip_prebuild(const __grid_constant__ Args A) {
    KARGS
    if (D.scal[S_TOPO_DIRTY]) { D.scal[S_MEMBER_CURSOR] = 0; D.scal[S_NCX] = 0; }
}
// one thread per molecule, unit heads act: owner of the unit from the head's centre; bands from every member
__global__ void k_strip_classify(const __grid_constant__ Args A, unsigned char *flag, double lo, double hi, double W) {
    KARGS
    const Consts &K = cK;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (!gid_live(K, D, gid) || D.unitOf[gid] != gid) return;
    double hx0, hy0; centre_of(K, D, gid, false, hx0, hy0);
    const bool own = d_strip_owner(K, hx0) == K.stripRank;
    int nmem = 1, m2 = -1; const int *row = nullptr;
    if (gid < K.NAt) { m2 = D.recCis[gid]; if (m2 >= 0) nmem = 2; }
    else if (D.cxSize[gid - K.NAt] > 1) { nmem = D.cxSize[gid - K.NAt]; row = D.members + D.cxOff[gid - K.NAt]; }
    int f = own ? 1 : 0;
    if (own)
        for (int i = 0; i < nmem; i++) {
            const int m = row ? row[i] : (i == 0 ? gid : m2);
            double x, y; centre_of(K, D, m, false, x, y);
            const double t = hash_x(K, x);
            if (t < lo + W) f |= 2;
            if (t >= hi - W) f |= 4;
        }
    for (int i = 0; i < nmem; i++) flag[row ? row[i] : (i == 0 ? gid : m2)] = (unsigned char)f;
}
// ---- the three messages in one pass ---------------------------------------------------------------------------------------
// Records must be id-sorted, i.e. in index order. A tile = ST_TILE consecutive molecules of one species (receptor tiles first),
// a thread 8 consecutive molecules: k_strip_count counts the members of the three lists per tile, k_strip_scan_tiles turns the
// counts into tile offsets and totals, k_strip_pack_all recomputes the ranks inside its tile and writes the records.
#define ST_TILE 2048
__device__ __forceinline__ void strip_tile_counts(const Consts &K, const Dev &D, const unsigned char *flag, int ntA, int &first, int &n, bool &lig, int c[3]) {
    const int t = blockIdx.x;
    lig = t >= ntA;
    first = (lig ? t - ntA : t) * ST_TILE + threadIdx.x * 8;          // index inside the species
    n = lig ? nB_live(D) : nA_live(D);
    c[0] = c[1] = c[2] = 0;
    for (int q = 0; q < 8; q++) {
        const int i = first + q;
        if (i < n) { const int f = flag[lig ? K.NAt + i : i]; c[0] += (f >> 1) & 1; c[1] += (f >> 2) & 1; c[2] += f & 1; }
    }
}
__global__ void __launch_bounds__(256) k_strip_count(const __grid_constant__ Args A, const unsigned char *flag, int ntA, int *tileCnt) {
    KARGS
    __shared__ int sh[3][8];
    int first, n, c[3]; bool lig;
    strip_tile_counts(cK, D, flag, ntA, first, n, lig, c);
    for (int l = 0; l < 3; l++) {
        int v = c[l];
        for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) sh[l][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) { int v = 0; for (int w = 0; w < 8; w++) v += sh[threadIdx.x][w]; tileCnt[blockIdx.x * 3 + threadIdx.x] = v; }
}
// one CTA: exclusive prefix over the tiles of each species for each list; totals to cnt6[list*2 + species]; capacity check of the messages
__global__ void __launch_bounds__(1024) k_strip_scan_tiles(const int *tileCnt, int *tileOff, int ntA, int ntB, int *cnt6, size_t cap0, size_t cap1, size_t cap2) {
    __shared__ int sh[32]; __shared__ int carry;
    for (int seq = 0; seq < 6; seq++) {
        const int sp = seq & 1, list = seq >> 1, t0 = sp ? ntA : 0, nt = sp ? ntB : ntA;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int base = 0; base < nt; base += 1024) {
            const int i = base + threadIdx.x, v = i < nt ? tileCnt[(t0 + i) * 3 + list] : 0;
            const int inc = warp_incl_scan(v);
            if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = inc;
            __syncthreads();
            if (threadIdx.x < 32) sh[threadIdx.x] = warp_incl_scan(sh[threadIdx.x]);
            __syncthreads();
            const int wofs = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
            if (i < nt) tileOff[(t0 + i) * 3 + list] = carry + wofs + inc - v;
            __syncthreads();
            if (threadIdx.x == 1023) carry += wofs + inc;
            __syncthreads();
        }
        if (threadIdx.x == 0) cnt6[seq] = carry;         // seq == list*2 + species
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const size_t cap[3] = {cap0, cap1, cap2};
        int bad = 0;
        for (int l = 0; l < 3; l++) if ((size_t)cnt6[l * 2] * sizeof(RecMsg) + (size_t)cnt6[l * 2 + 1] * sizeof(LigMsg) > cap[l]) bad |= 1 << l;
        cnt6[6] = bad;
    }
}
__device__ __forceinline__ void strip_write_record(const Consts &K, const Dev &D, int gid, char *out, int nRecOut, int pos) {
    if (gid < K.NAt) {
        RecMsg m; m.ref = (int)D.refA[gid];
        const int l = D.recLig[gid], c = D.recCis[gid];
        m.ligRef = l >= 0 ? (int)D.refB[l] : 0; m.site = D.recSite[gid]; m.cisRef = c >= 0 ? (int)D.refA[c] : 0;
        const double2 cc = D.recC[gid], s2 = D.recS2[gid], s3 = D.recS3[gid];
        m.pose[0] = cc.x; m.pose[1] = cc.y; m.pose[2] = s2.x; m.pose[3] = s2.y; m.pose[4] = s3.x; m.pose[5] = s3.y;
        reinterpret_cast<RecMsg *>(out)[pos] = m;
    } else {
        const int h = gid - K.NAt;
        LigMsg *o = reinterpret_cast<LigMsg *>(out + (size_t)nRecOut * sizeof(RecMsg)) + pos;
        o->ref = (int)D.refB[h];
        for (int k = 0; k < 3; k++) { const int r = D.ligRec[h * 3 + k]; o->recRef[k] = r >= 0 ? (int)D.refA[r] : 0; }
        const double *p = D.lig + (size_t)h * 24;
        for (int q = 0; q < 24; q++) o->pose[q] = p[q];
    }
}
__global__ void __launch_bounds__(256) k_strip_pack_all(const __grid_constant__ Args A, const unsigned char *flag, int ntA, const int *tileOff, const int *cnt6,
                                                        char *msg0, char *msg1, char *msg2) {
    KARGS
    const Consts &K = cK;
    __shared__ int sh[3][8];
    if (cnt6[6]) return;                                   // a message does not fit its buffer: the host reports it
    int first, n, c[3]; bool lig;
    strip_tile_counts(K, D, flag, ntA, first, n, lig, c);
    int rank[3];
    for (int l = 0; l < 3; l++) {                          // exclusive prefix of the thread counts inside the tile
        const int inc = warp_incl_scan(c[l]);
        if ((threadIdx.x & 31) == 31) sh[l][threadIdx.x >> 5] = inc;
        rank[l] = inc - c[l];
    }
    __syncthreads();
    for (int l = 0; l < 3; l++) {
        for (int w = 0; w < (int)(threadIdx.x >> 5); w++) rank[l] += sh[l][w];
        rank[l] += tileOff[blockIdx.x * 3 + l];
    }
    char *msg[3] = {msg0, msg1, msg2};
    for (int q = 0; q < 8; q++) {
        const int i = first + q;
        if (i >= n) break;
        const int gid = lig ? K.NAt + i : i, f = flag[gid];
        const int bits[3] = {(f >> 1) & 1, (f >> 2) & 1, f & 1};
        for (int l = 0; l < 3; l++) if (bits[l]) strip_write_record(K, D, gid, msg[l], cnt6[l * 2], rank[l]++);
    }
}
template <class T> __device__ __forceinline__ int d_lower_bound(const T *a, int n, int ref) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid].ref < ref) lo = mid + 1; else hi = mid; }
    return lo;
}
// merge of the three id-sorted lists (kept, from low, from high): every record knows its final index
struct MergeSrc { const RecMsg *r[3]; const LigMsg *l[3]; int nr[3], nl[3]; };
__global__ void k_strip_merge(const __grid_constant__ Args A, MergeSrc M, int *bondRef) {
    KARGS
    const Consts &K = cK;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int totR = M.nr[0] + M.nr[1] + M.nr[2], totL = M.nl[0] + M.nl[1] + M.nl[2];
    if (i < totR) {
        int k = 0; while (i >= M.nr[k]) { i -= M.nr[k]; k++; }
        const RecMsg m = M.r[k][i];
        int pos = i;
        for (int o = 0; o < 3; o++) if (o != k) pos += d_lower_bound(M.r[o], M.nr[o], m.ref);
        D.recC[pos] = make_double2(m.pose[0], m.pose[1]); D.recS2[pos] = make_double2(m.pose[2], m.pose[3]); D.recS3[pos] = make_double2(m.pose[4], m.pose[5]);
        D.refA[pos] = (unsigned)m.ref; D.recSite[pos] = m.site;
        bondRef[pos * 2] = m.ligRef; bondRef[pos * 2 + 1] = m.cisRef;
    } else if (i < totR + totL) {
        i -= totR;
        int k = 0; while (i >= M.nl[k]) { i -= M.nl[k]; k++; }
        const LigMsg *src = &M.l[k][i];
        int pos = i;
        for (int o = 0; o < 3; o++) if (o != k) pos += d_lower_bound(M.l[o], M.nl[o], src->ref);
        D.refB[pos] = (unsigned)src->ref;
        double *p = D.lig + (size_t)pos * 24;
        for (int q = 0; q < 24; q++) p[q] = src->pose[q];
        for (int q = 0; q < 3; q++) bondRef[2 * K.NAt + pos * 3 + q] = src->recRef[q];
    }
}
__device__ __forceinline__ int d_find_ref(const unsigned *a, int n, int ref) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < (unsigned)ref) lo = mid + 1; else hi = mid; }
    return (lo < n && a[lo] == (unsigned)ref) ? lo : -1;
}
__global__ void k_strip_fix_bonds(const __grid_constant__ Args A, const int *bondRef, int nA, int nB) {
    KARGS
    const Consts &K = cK;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nA) {
        const int lr = bondRef[i * 2], cr = bondRef[i * 2 + 1];
        const int l = lr ? d_find_ref(D.refB, nB, lr) : -1, c = cr ? d_find_ref(D.refA, nA, cr) : -1;
        if ((lr && l < 0) || (cr && c < 0)) atomicOr(&D.scal[S_OVERFLOW], 16);        // a unit arrived incomplete
        D.recLig[i] = l; D.recCis[i] = c; if (l < 0) D.recSite[i] = -1;
    } else if (i < nA + nB) {
        const int h = i - nA;
        for (int q = 0; q < 3; q++) {
            const int rr = bondRef[2 * K.NAt + h * 3 + q], r = rr ? d_find_ref(D.refA, nA, rr) : -1;
            if (rr && r < 0) atomicOr(&D.scal[S_OVERFLOW], 16);
            D.ligRec[h * 3 + q] = r;
        }
    }
}

static void strip_dev_free(kmc_handle *h) { if (h->strip_dev && h->strip_dev->hcnt) cudaFreeHost(h->strip_dev->hcnt); delete h->strip_dev; h->strip_dev = nullptr; }    // device buffers are in h->allocs

static int strip_dev_alloc(kmc_handle *h) {
    StripDev &S = *h->strip_dev;
    if (S.flag) return KMC_OK;
    const int NT = h->NT;
    S.ntA = (h->NAt + ST_TILE - 1) / ST_TILE; S.ntB = (h->NBt + ST_TILE - 1) / ST_TILE;
    bool ok = dalloc(h, &S.flag, NT) == cudaSuccess && dalloc(h, &S.tileCnt, (size_t)3 * (S.ntA + S.ntB) + 3) == cudaSuccess &&
              dalloc(h, &S.tileOff, (size_t)3 * (S.ntA + S.ntB) + 3) == cudaSuccess && dalloc(h, &S.dcnt, 8) == cudaSuccess &&
              dalloc(h, &S.bondRef, (size_t)2 * h->NAt + (size_t)3 * h->NBt + 8) == cudaSuccess &&
              cudaMallocHost((void **)&S.hcnt, 8 * sizeof(int)) == cudaSuccess;
    const size_t full = (size_t)h->NAt * sizeof(RecMsg) + (size_t)h->NBt * sizeof(LigMsg) + 64, band = full / 3 + 4096;
    for (int k = 0; k < 3 && ok; k++) { S.msgCap[k] = k == 2 ? full : band; ok = dalloc(h, &S.msg[k], S.msgCap[k]) == cudaSuccess; }
    for (int k = 0; k < 2 && ok; k++) { S.rcvCap[k] = band; ok = dalloc(h, &S.rcv[k], S.rcvCap[k]) == cudaSuccess; }
    if (!ok) { h->err = "strip: device buffer allocation failed"; return KMC_ERR_CUDA; }
    return KMC_OK;
}

// Device refresh, part 1. Afterwards kmc_strip_message_dev(side) gives the device address and the record counts of each message.
extern "C" int kmc_strip_begin_refresh_dev(kmc_handle *h) {
    if (!h || !h->strip_on) { if (h) h->err = "kmc_strip_begin_refresh_dev: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    if (!h->strip_dev) h->strip_dev = new StripDev;
    int rc = strip_dev_alloc(h); if (rc) return rc;
    StripDev &S = *h->strip_dev;
    cudaStream_t st = h->stream;
    const Args A{h->D, h->K};
    const int NT = h->NT, NAt = h->NAt, NBt = h->NBt, B = 128;
    // complexes of the CURRENT bond table (the last step's reactions may have changed it)
    LAUNCH(KID_UF_INIT, (k_strip_prebuild<<<1, 1, 0, st>>>(A)));
    LAUNCH(KID_UF_INIT, (k_uf_init<<<nblk(NT, 256), 256, 0, st>>>(A, 0)));
    LAUNCH(KID_UF_HOOK, (k_uf_hook<<<nblk(std::max(NAt, 1), 256), 256, 0, st>>>(A)));
    LAUNCH(KID_UF_FLATTEN, (k_uf_flatten<<<nblk(NT, 256), 256, 0, st>>>(A)));
    LAUNCH(KID_CX_BUILD, (k_cx_build<<<nblk(NBt, B), B, 0, st>>>(A)));
    CK(cudaMemsetAsync(S.flag, 0, NT, st));
    k_strip_classify<<<nblk(NT, 128), 128, 0, st>>>(A, S.flag, h->strip_lo, h->strip_hi, h->strip_W);
    const int nt = S.ntA + S.ntB;
    k_strip_count<<<nt, 256, 0, st>>>(A, S.flag, S.ntA, S.tileCnt);
    k_strip_scan_tiles<<<1, 1024, 0, st>>>(S.tileCnt, S.tileOff, S.ntA, S.ntB, S.dcnt, S.msgCap[0], S.msgCap[1], S.msgCap[2]);
    k_strip_pack_all<<<nt, 256, 0, st>>>(A, S.flag, S.ntA, S.tileOff, S.dcnt, S.msg[0], S.msg[1], S.msg[2]);
    CK(cudaMemcpyAsync(S.hcnt, S.dcnt, 7 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int q = 0; q < 6; q++) S.cnt[q] = S.hcnt[q];
    if (S.hcnt[6]) { h->err = "strip: message buffer too small (band holds more than a third of the local capacity)"; return KMC_ERR_CAPACITY; }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return KMC_OK;
}
extern "C" int kmc_strip_message_dev(kmc_handle *h, int32_t side, void **dev_ptr, int64_t *n_rec, int64_t *n_lig) {
    if (!h || !h->strip_dev || side < 0 || side > 2) return KMC_ERR_INVALID;
    StripDev &S = *h->strip_dev;
    if (dev_ptr) *dev_ptr = S.msg[side];
    if (n_rec) *n_rec = S.cnt[side * 2];
    if (n_lig) *n_lig = S.cnt[side * 2 + 1];
    return KMC_OK;
}
// device buffer the caller fills with the message coming from the lower-x (side 0) / higher-x (side 1) neighbour
extern "C" int kmc_strip_recv_dev(kmc_handle *h, int32_t side, int64_t n_rec, int64_t n_lig, void **dev_ptr) {
    if (!h || !h->strip_dev || side < 0 || side > 1 || !dev_ptr) return KMC_ERR_INVALID;
    StripDev &S = *h->strip_dev;
    const size_t need = (size_t)n_rec * sizeof(RecMsg) + (size_t)n_lig * sizeof(LigMsg);
    if (need > S.rcvCap[side]) { h->err = "strip: incoming message larger than the receive buffer"; return KMC_ERR_CAPACITY; }
    *dev_ptr = S.rcv[side];
    return KMC_OK;
}
// Device refresh, part 2: merge kept + received (counts of the two incoming messages given), translate bonds, new live counts
extern "C" int kmc_strip_rebuild_dev(kmc_handle *h, int64_t rec_low, int64_t lig_low, int64_t rec_high, int64_t lig_high) {
    if (!h || !h->strip_dev) return KMC_ERR_INVALID;
    CK(cudaSetDevice(h->P.device));
    StripDev &S = *h->strip_dev;
    cudaStream_t st = h->stream;
    MergeSrc M;
    const int nr[3] = {S.cnt[4], (int)rec_low, (int)rec_high}, nl[3] = {S.cnt[5], (int)lig_low, (int)lig_high};
    const char *base[3] = {S.msg[2], S.rcv[0], S.rcv[1]};
    for (int k = 0; k < 3; k++) {
        M.nr[k] = nr[k]; M.nl[k] = nl[k];
        M.r[k] = reinterpret_cast<const RecMsg *>(base[k]); M.l[k] = reinterpret_cast<const LigMsg *>(base[k] + (size_t)nr[k] * sizeof(RecMsg));
    }
    const int nA = nr[0] + nr[1] + nr[2], nB = nl[0] + nl[1] + nl[2];
    if (nA > h->NAt || nB > h->NBt) {
        h->err = "strip: local capacity exceeded (" + std::to_string(nA) + "/" + std::to_string(h->NAt) + " receptors, " + std::to_string(nB) + "/" +
                 std::to_string(h->NBt) + " ligands)"; return KMC_ERR_CAPACITY;
    }
    const Args A{h->D, h->K};
    k_strip_merge<<<nblk(std::max(nA + nB, 1), 128), 128, 0, st>>>(A, M, S.bondRef);
    k_strip_fix_bonds<<<nblk(std::max(nA + nB, 1), 128), 128, 0, st>>>(A, S.bondRef, nA, nB);
    int live[2] = {nA, nB}, one = 1;
    CK(cudaMemcpyAsync(h->D.scal + S_NA_LIVE, live, sizeof live, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    h->stepped = false; h->sinceBuild = 0; h->strip_refreshes++;
    return kmc_sync(h);
}
static int strip_auto_refresh(kmc_handle *h) { h->strip_since = 0; return KMC_OK; }
