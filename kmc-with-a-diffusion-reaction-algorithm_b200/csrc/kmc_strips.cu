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
    h->fused = false;          // (strips renumber and migrate molecules: the general path)
    h->strip_on = true; h->strip_W = halo_width; h->strip_lo = -K.Lx / 2 + rank * width; h->strip_hi = h->strip_lo + width;
    if (nranks > 1) {        // the grid only has to cover the strip and its halos (in the periodic frame of the strip)
        const double edge = 1.0 / K.cellInv;
        K.gx0 = h->strip_lo - halo_width - 2 * edge; K.keyX0 = K.gx0;
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
// Device-side refresh: classify + pack, exchange, merge -- every phase enqueued on the handle's stream, NO host synchronisation.
// Messages live in device buffers of FIXED capacity: a 64-byte header {nRec, nLig, overflow} followed by RecMsg[nRec] and
// LigMsg[nLig] (both id-sorted). The counts travel inside the message, so one exchange round suffices and the host never
// has to know them: it always moves the whole band buffer (a few MB over NVLink), the merge kernels read the headers.
//   exchange = NCCL (kmc_strip_refresh: ncclSend/ncclRecv to the two x-neighbours in one group, on the handle's stream), or
//              device-to-device copies between handles of one process (kmc_strip_refresh_local: K logical ranks on one GPU,
//              how the tests prove equality with the single-GPU run).
// By-products of every refresh: the owned-only bond counts, complex statistics and oligomer-size histogram of this rank
// (kmc_strip_get_series / kmc_strip_get_oligomer_hist all-reduce them: the bond.dat row of the WHOLE membrane).
// The host path above is kept as the reference implementation (tests compare the two).
// ================================================================================================================
#include <chrono>
#include <dlfcn.h>
#include <nccl.h>

#define MSG_HDR 64
#define STRIP_HIST_BINS 256

// NCCL is bound at run time (dlopen): the library stays loadable on a machine without NCCL, and inside a process that already
// carries a libnccl.so.2 (torch) the same copy is used.
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static const char *nccl_load() {
    if (g_nccl.lib) return nullptr;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 not found (strips over several GPUs need NCCL)";
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) return "libnccl: missing symbol " name
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.lib = lib;
    return nullptr;
}
#define NC(call)                                                                                   \
    do {                                                                                           \
        ncclResult_t r_ = (call);                                                                  \
        if (r_ != ncclSuccess) { h->err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return KMC_ERR_CUDA; }   \
    } while (0)

struct StripDev {
    unsigned char *flag = nullptr;            // [NT] bit0 owned, bit1 send to lower-x neighbour, bit2 send to higher-x neighbour
    int *tileCnt = nullptr, *tileOff = nullptr;   // per tile of ST_TILE molecules of one species: members of list 0 low, 1 high, 2 keep; their exclusive prefixes
    int *dcnt = nullptr;                      // [8] totals [list*2 + species], [6] = lists that do not fit their buffer (bit mask)
    int ntA = 0, ntB = 0;
    char *msg[3] = {nullptr}; size_t msgCap[3] = {0};     // packed messages: 0 to lower x, 1 to higher x (band capacity), 2 = the owned set (full capacity)
    char *rcv[2] = {nullptr};                 // what the neighbours sent (0 from lower x, 1 from higher x), band capacity
    size_t bandCap = 0;                       // bytes of a band message (header + records): the same on every rank
    int *bondRef = nullptr;                   // [NAt*3 + NBt*3] bonds of the merged molecules as reference ids, before translation
    int *keptRank = nullptr;                  // [NT] owned molecules of the same species with a lower index
    unsigned char *ownFlag = nullptr;         // [NT] after a merge: molecule of a unit this rank owns
    unsigned *refA2 = nullptr, *refB2 = nullptr;   // reference ids of the merged set (copied over refA / refB by k_strip_fix_bonds)
    int *cref = nullptr;                      // [2 NAt + 2 NBt] compact reference ids of the incoming lists
    int *series = nullptr;                    // [8] device, owned only: R-L, mono-cis, cis bonds, complexes, molecules in complexes, -, -, running-max complex
    unsigned long long *hist = nullptr;       // [STRIP_HIST_BINS] sizes of the owned ligand-rooted complexes (last step's tables)
    int *hostI = nullptr;                     // pinned: [0..15] header staging, [16..31] series read-back
    unsigned long long *hostH = nullptr;      // pinned: histogram read-back
    ncclComm_t comm = nullptr;
    bool fresh = false;                       // no step since the last refresh: by-products and msg[2] describe the current state
    double budget = INFINITY;                 // largest x-extent of a unit the halo width covers (halo - refresh_every * reach per step)
};

// by-product 1 (before the complexes are rebuilt: the tables are those of the last step's sweep, like kmc_get_series /
// kmc_get_oligomer_hist on one GPU): complexes rooted at a ligand this rank owns -- owner = strip of the root's centre
__global__ void k_strip_cx_owned(const __grid_constant__ Args A, int stepped, int *series, unsigned long long *hist) {
    KARGS
    const Consts &K = cK;
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    int size = 0;
    if (stepped && h < nB_live(D) && D.unitOf[K.NAt + h] == K.NAt + h && d_strip_owner(K, D.lig[(size_t)h * 24]) == K.stripRank) size = D.cxSize[h];
    const unsigned ones = __ballot_sync(0xffffffffu, size == 1);
    if ((threadIdx.x & 31) == 0 && ones) atomicAdd(&hist[1], (unsigned long long)__popc(ones));
    if (size > 1) { atomicAdd(&hist[min(size, STRIP_HIST_BINS - 1)], 1ULL); atomicAdd(&series[3], 1); atomicAdd(&series[4], size); }
}
// by-product 2 (after k_strip_classify): bonds of the receptors whose unit this rank owns (the bond.dat columns, main.cpp:2251)
__global__ void k_strip_bonds_owned(const __grid_constant__ Args A, const unsigned char *flag, int *series) {
    KARGS
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    int rl = 0, mono = 0, cis = 0;
    if (a < nA_live(D) && (flag[a] & 1)) {
        rl = D.recLig[a] >= 0;
        const int p = D.recCis[a];
        if (p > a) { if (rl || D.recLig[p] >= 0) cis = 1; else mono = 1; }
    }
    for (int o = 16; o; o >>= 1) { rl += __shfl_down_sync(0xffffffffu, rl, o); mono += __shfl_down_sync(0xffffffffu, mono, o); cis += __shfl_down_sync(0xffffffffu, cis, o); }
    if ((threadIdx.x & 31) == 0) { if (rl) atomicAdd(&series[0], rl); if (mono) atomicAdd(&series[1], mono); if (cis) atomicAdd(&series[2], cis); }
    if (a == 0) series[7] = D.maxComplex[0];
}
// one thread per molecule, unit heads act: owner of the unit from the head's centre; bands from every member.
// Guard of the exactness precondition: a unit wider (in x) than the halo budget could reach beyond what the neighbours sent.
__global__ void k_strip_classify(const __grid_constant__ Args A, unsigned char *flag, double lo, double hi, double W, double budget) {
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
    if (own) {
        double xmin = INFINITY, xmax = -INFINITY;
        for (int i = 0; i < nmem; i++) {
            const int m = row ? row[i] : (i == 0 ? gid : m2);
            double x, y; centre_of(K, D, m, false, x, y);
            const double t = hash_x(K, x);
            xmin = fmin(xmin, t); xmax = fmax(xmax, t);
            if (t < lo + W) f |= 2;
            if (t >= hi - W) f |= 4;
        }
        if (K.strips > 1 && xmax - xmin > budget) atomicOr(&D.scal[S_OVERFLOW], 64);
        if (K.strips == 2 && (f & 2)) f &= ~4;          // two ranks: both neighbours are the same peer -- a unit inside both bands travels once
    }
    for (int i = 0; i < nmem; i++) flag[row ? row[i] : (i == 0 ? gid : m2)] = (unsigned char)f;
}
__global__ void k_strip_flag_all(const __grid_constant__ Args A, unsigned char *flag) {       // every live molecule into list 2
    KARGS
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < cK.NT) flag[gid] = gid_live(cK, D, gid) ? 1 : 0;
}
// ---- the three messages in one pass ---------------------------------------------------------------------------------------
// Records must be id-sorted, i.e. in index order. A tile = ST_TILE consecutive molecules of one species (receptor tiles first),
// a thread 8 consecutive molecules: k_strip_count counts the members of the three lists per tile, k_strip_scan_tiles turns the
// counts into tile offsets and totals and writes the message headers, k_strip_pack_all recomputes the ranks inside its tile and
// writes the records.
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
// one CTA: exclusive prefix over the tiles of each species for each list; totals to cnt6[list*2 + species]; capacity check and
// header of each message (a list that does not fit travels empty with its overflow word set: sender and receiver both report it)
struct MsgBufs { char *p[3]; size_t cap[3]; };
__global__ void __launch_bounds__(1024) k_strip_scan_tiles(const int *tileCnt, int *tileOff, int ntA, int ntB, int *cnt6, MsgBufs M, int *scal) {
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
        int bad = 0;
        for (int l = 0; l < 3; l++) {
            const bool fits = MSG_HDR + (size_t)cnt6[l * 2] * sizeof(RecMsg) + (size_t)cnt6[l * 2 + 1] * sizeof(LigMsg) <= M.cap[l];
            if (!fits) bad |= 1 << l;
            int *hdr = reinterpret_cast<int *>(M.p[l]);
            hdr[0] = fits ? cnt6[l * 2] : 0; hdr[1] = fits ? cnt6[l * 2 + 1] : 0; hdr[2] = fits ? 0 : 1;
        }
        cnt6[6] = bad;
        if (bad) atomicOr(&scal[S_OVERFLOW], 128);
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
__global__ void __launch_bounds__(256) k_strip_pack_all(const __grid_constant__ Args A, const unsigned char *flag, int ntA, const int *tileOff, const int *cnt6, MsgBufs M, int lists, int *keptRank) {
    KARGS
    const Consts &K = cK;
    __shared__ int sh[3][8];
    const int bad = cnt6[6] | ~lists;          // lists: bit mask of the messages this launch writes
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
    for (int q = 0; q < 8; q++) {
        const int i = first + q;
        if (i >= n) break;
        const int gid = lig ? K.NAt + i : i, f = flag[gid];
        const int bits[3] = {(f >> 1) & 1, (f >> 2) & 1, f & 1};
        if (keptRank) keptRank[gid] = rank[2];             // owned molecules of this species with a lower index
        for (int l = 0; l < 3; l++) if (bits[l]) { if (!((bad >> l) & 1)) strip_write_record(K, D, gid, M.p[l] + MSG_HDR, cnt6[l * 2], rank[l]); rank[l]++; }
    }
}
__device__ __forceinline__ int d_lower_bound_i(const int *a, int n, int ref) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < ref) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int d_find_ref(const unsigned *a, int n, int ref) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < (unsigned)ref) lo = mid + 1; else hi = mid; }
    return (lo < n && a[lo] == (unsigned)ref) ? lo : -1;
}
// ---- merge: new local set = the molecules this rank keeps (owned units, still in the arrays) + the two incoming lists, id-sorted.
// The kept molecules never leave the arrays: they are copied from the committed pose buffers straight to their new index in the
// other ("next") pose buffers, which are scratch between two steps; the host then swaps the buffers like a step does. The lists
// are disjoint: a molecule has one owner, and with two ranks (both bands from the same peer) the sender lists a unit once.
struct MergeIn {
    const char *base[2];          // incoming lists: [from lower x, from higher x] (header + records)
    int *cref;                    // compact copies of their reference ids: rec list 0 at 0, rec list 1 at NAt, lig list 0 at 2 NAt, lig list 1 at 2 NAt + NBt
    const int *dcnt;              // [4] kept receptors, [5] kept ligands
    const int *keptRank;          // per old molecule: kept molecules of its species with a lower index
    const unsigned char *flag;    // bit0: kept
    unsigned *refA2, *refB2;      // new reference ids
    unsigned char *ownFlag;       // new: 1 = molecule of a unit this rank owns
    int *bond;                    // new bonds as reference ids: receptor [pos*3 + {lig, cis, site}], ligand [3 NAt + pos*3 + site]
};
__device__ __forceinline__ bool merge_counts(const Consts &K, const MergeIn &M, int nr[2], int nl[2], int &keptR, int &keptL) {
    bool bad = false;
    for (int k = 0; k < 2; k++) { const int *hdr = reinterpret_cast<const int *>(M.base[k]); nr[k] = hdr[0]; nl[k] = hdr[1]; bad |= hdr[2] != 0; }
    keptR = M.dcnt[4]; keptL = M.dcnt[5];
    return !bad && keptR + nr[0] + nr[1] <= K.NAt && keptL + nl[0] + nl[1] <= K.NBt;
}
__global__ void k_strip_refs(const __grid_constant__ Args A, MergeIn M) {
    KARGS
    const Consts &K = cK;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < 2; k++) {
        const int *hdr = reinterpret_cast<const int *>(M.base[k]);
        const int nr = min(hdr[0], K.NAt), nl = min(hdr[1], K.NBt);
        if (i < nr) { M.cref[k * K.NAt + i] = reinterpret_cast<const RecMsg *>(M.base[k] + MSG_HDR)[i].ref; return; }
        i -= nr;
        if (i < nl) { M.cref[2 * K.NAt + k * K.NBt + i] = reinterpret_cast<const LigMsg *>(M.base[k] + MSG_HDR + (size_t)nr * sizeof(RecMsg))[i].ref; return; }
        i -= nl;
    }
}
// lower bound of `ref` in the sorted array c[0..n) for all lanes of a warp whose refs ascend with the lane (sorted molecules):
// the bounds of the first and the last live lane are found by two warp-uniform binary searches (every lane probes the same
// address: one broadcast load per step), each lane then searches only the few entries between them
__device__ __forceinline__ int warp_lower_bound(const int *c, int n, int ref, unsigned live) {
    if (!live) return 0;
    const int first = __ffs(live) - 1, last = 31 - __clz(live);
    const int lo0 = d_lower_bound_i(c, n, __shfl_sync(0xffffffffu, ref, first));
    int hi = d_lower_bound_i(c, n, __shfl_sync(0xffffffffu, ref, last));
    int lo = lo0;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (c[mid] < ref) lo = mid + 1; else hi = mid; }
    return lo;
}
#define MERGE_B 128
__global__ void __launch_bounds__(MERGE_B) k_strip_merge(const __grid_constant__ Args A, MergeIn M, int nbA, int nbB) {
    KARGS
    const Consts &K = cK;
    int nr[2], nl[2], keptR, keptL;
    if (!merge_counts(K, M, nr, nl, keptR, keptL)) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&D.scal[S_OVERFLOW], 128); return; }      // a message or the local capacity is too small
    const int nAold = nA_live(D), nBold = nB_live(D);
    const int *cR[2] = {M.cref, M.cref + K.NAt}, *cL[2] = {M.cref + 2 * K.NAt, M.cref + 2 * K.NAt + K.NBt};
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x < nbA) {                                  // receptors that stay
        const int a = blockIdx.x * MERGE_B + threadIdx.x;
        const bool live = a < nAold;
        const int ref = live ? (int)D.refA[a] : 0x7fffffff;
        const unsigned lm = __ballot_sync(0xffffffffu, live);
        const int o0 = warp_lower_bound(cR[0], nr[0], ref, lm), o1 = warp_lower_bound(cR[1], nr[1], ref, lm);
        if (!live || !(M.flag[a] & 1)) return;
        const int pos = M.keptRank[a] + o0 + o1;
        D.recCn[pos] = D.recC[a]; D.recS2n[pos] = D.recS2[a]; D.recS3n[pos] = D.recS3[a];
        M.refA2[pos] = (unsigned)ref; M.ownFlag[pos] = 1;
        const int l = D.recLig[a], c = D.recCis[a];
        M.bond[pos * 3] = l >= 0 ? (int)D.refB[l] : 0; M.bond[pos * 3 + 1] = c >= 0 ? (int)D.refA[c] : 0; M.bond[pos * 3 + 2] = D.recSite[a];
        return;
    }
    if ((int)blockIdx.x < nbA + nbB) {                            // ligands that stay: the 192-byte poses are copied by the whole warp, 16 bytes per lane
        const int b = (blockIdx.x - nbA) * MERGE_B + threadIdx.x;
        const bool live = b < nBold;
        const int ref = live ? (int)D.refB[b] : 0x7fffffff;
        const unsigned lm = __ballot_sync(0xffffffffu, live);
        const int o0 = warp_lower_bound(cL[0], nl[0], ref, lm), o1 = warp_lower_bound(cL[1], nl[1], ref, lm);
        const bool keep = live && (M.flag[K.NAt + b] & 1);
        const int pos = keep ? M.keptRank[K.NAt + b] + o0 + o1 : -1;
        const int b0 = b - lane;
        const double2 *src = reinterpret_cast<const double2 *>(D.lig) + (size_t)b0 * 12;
        double2 *dst = reinterpret_cast<double2 *>(D.lign);
#pragma unroll
        for (int it = 0; it < 12; it++) {
            const int ch = it * 32 + lane, l = ch / 12, q = ch - l * 12;
            const int dp = __shfl_sync(0xffffffffu, pos, l);
            if (dp >= 0) dst[(size_t)dp * 12 + q] = src[ch];
        }
        if (!keep) return;
        M.refB2[pos] = (unsigned)ref; M.ownFlag[K.NAt + pos] = 1;
        for (int q = 0; q < 3; q++) { const int r = D.ligRec[b * 3 + q]; M.bond[3 * K.NAt + pos * 3 + q] = r >= 0 ? (int)D.refA[r] : 0; }
        return;
    }
    int i = (blockIdx.x - nbA - nbB) * MERGE_B + threadIdx.x;     // an incoming record
    if (i < nr[0] + nr[1]) {
        const int k = i < nr[0] ? 0 : 1; if (k) i -= nr[0];
        const RecMsg m = reinterpret_cast<const RecMsg *>(M.base[k] + MSG_HDR)[i];
        int q = 0; { int lo = 0, hi = nAold; while (lo < hi) { const int mid = (lo + hi) >> 1; if (D.refA[mid] < (unsigned)m.ref) lo = mid + 1; else hi = mid; } q = lo; }
        const int pos = i + (q < nAold ? M.keptRank[q] : keptR) + d_lower_bound_i(cR[1 - k], nr[1 - k], m.ref);
        D.recCn[pos] = make_double2(m.pose[0], m.pose[1]); D.recS2n[pos] = make_double2(m.pose[2], m.pose[3]); D.recS3n[pos] = make_double2(m.pose[4], m.pose[5]);
        M.refA2[pos] = (unsigned)m.ref; M.ownFlag[pos] = 0;
        M.bond[pos * 3] = m.ligRef; M.bond[pos * 3 + 1] = m.cisRef; M.bond[pos * 3 + 2] = m.site;
        return;
    }
    i -= nr[0] + nr[1];
    if (i < nl[0] + nl[1]) {
        const int k = i < nl[0] ? 0 : 1; if (k) i -= nl[0];
        const LigMsg *src = reinterpret_cast<const LigMsg *>(M.base[k] + MSG_HDR + (size_t)nr[k] * sizeof(RecMsg)) + i;
        const int ref = src->ref;
        int q = 0; { int lo = 0, hi = nBold; while (lo < hi) { const int mid = (lo + hi) >> 1; if (D.refB[mid] < (unsigned)ref) lo = mid + 1; else hi = mid; } q = lo; }
        const int pos = i + (q < nBold ? M.keptRank[K.NAt + q] : keptL) + d_lower_bound_i(cL[1 - k], nl[1 - k], ref);
        double *p = D.lign + (size_t)pos * 24;
        for (int t = 0; t < 24; t++) p[t] = src->pose[t];
        M.refB2[pos] = (unsigned)ref; M.ownFlag[K.NAt + pos] = 0;
        for (int t = 0; t < 3; t++) M.bond[3 * K.NAt + pos * 3 + t] = src->recRef[t];
    }
}
// bonds back from reference ids to local indices (new numbering); thread 0 publishes the new live counts (nobody in this kernel reads them)
__global__ void k_strip_fix_bonds(const __grid_constant__ Args A, MergeIn M) {
    KARGS
    const Consts &K = cK;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int nr[2], nl[2], keptR, keptL;
    if (!merge_counts(K, M, nr, nl, keptR, keptL)) return;
    const int nA = keptR + nr[0] + nr[1], nB = keptL + nl[0] + nl[1];
    if (i == 0) { D.scal[S_NA_LIVE] = nA; D.scal[S_NB_LIVE] = nB; D.scal[S_TOPO_DIRTY] = 1; }
    if (i < nA) {
        const int lr = M.bond[i * 3], cr = M.bond[i * 3 + 1];
        const int l = lr ? d_find_ref(M.refB2, nB, lr) : -1, c = cr ? d_find_ref(M.refA2, nA, cr) : -1;
        if ((lr && l < 0) || (cr && c < 0)) atomicOr(&D.scal[S_OVERFLOW], 16);        // a unit arrived incomplete
        D.recLig[i] = l; D.recCis[i] = c; D.recSite[i] = l < 0 ? -1 : M.bond[i * 3 + 2];
        D.refA[i] = M.refA2[i];
    } else if (i >= K.NAt && i - K.NAt < nB) {
        const int h = i - K.NAt;
        for (int q = 0; q < 3; q++) {
            const int rr = M.bond[3 * K.NAt + h * 3 + q], r = rr ? d_find_ref(M.refA2, nA, rr) : -1;
            if (rr && r < 0) atomicOr(&D.scal[S_OVERFLOW], 16);
            D.ligRec[h * 3 + q] = r;
        }
        D.refB[h] = M.refB2[h];
    }
}
__global__ void k_strip_flag_owned(const __grid_constant__ Args A, const unsigned char *ownFlag, unsigned char *flag) {       // the owned set into list 2
    KARGS
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < cK.NT) flag[gid] = (gid_live(cK, D, gid) && ownFlag[gid]) ? 1 : 0;
}

static void strip_dev_free(kmc_handle *h) {       // device buffers are in h->allocs
    if (!h->strip_dev) return;
    StripDev &S = *h->strip_dev;
    if (S.comm && g_nccl.CommDestroy) { cudaSetDevice(h->P.device); cudaStreamSynchronize(h->stream); g_nccl.CommDestroy(S.comm); }
    if (S.hostI) cudaFreeHost(S.hostI);
    if (S.hostH) cudaFreeHost(S.hostH);
    delete h->strip_dev; h->strip_dev = nullptr;
}

// reach of information per step (A): the largest overlap reach (ligand-ligand centres) plus twice the largest displacement of a
// molecule in one step (translation + the swing of a ligand site / complex member under the rotation)
static double strip_reach_per_step(const Consts &K) {
    const double rs = K.rB * 2 / sqrt(3.0);
    const double disp = std::max({K.ampA, K.ampB, K.ampCis, K.ampBond}) + 0.45 * (rs + K.rB);
    return K.reachLL + 2 * disp;
}
extern "C" double kmc_strip_halo_width(const kmc_params *p, int32_t refresh_every, double complex_extent) {
    if (!p || refresh_every < 1) return -1.0;
    Consts K; fill_consts(*p, K);
    return refresh_every * strip_reach_per_step(K) + complex_extent;
}

static int strip_dev_alloc(kmc_handle *h) {
    if (!h->strip_dev) h->strip_dev = new StripDev;
    StripDev &S = *h->strip_dev;
    if (S.flag) return KMC_OK;
    const int NT = h->NT;
    S.ntA = (h->NAt + ST_TILE - 1) / ST_TILE; S.ntB = (h->NBt + ST_TILE - 1) / ST_TILE;
    bool ok = dalloc(h, &S.flag, NT) == cudaSuccess && dalloc(h, &S.tileCnt, (size_t)3 * (S.ntA + S.ntB) + 3) == cudaSuccess &&
              dalloc(h, &S.tileOff, (size_t)3 * (S.ntA + S.ntB) + 3) == cudaSuccess && dalloc(h, &S.dcnt, 8) == cudaSuccess &&
              dalloc(h, &S.bondRef, (size_t)3 * h->NAt + (size_t)3 * h->NBt + 8) == cudaSuccess &&
              dalloc(h, &S.keptRank, NT) == cudaSuccess && dalloc(h, &S.ownFlag, NT) == cudaSuccess && dalloc(h, &S.refA2, std::max(h->NAt, 1)) == cudaSuccess &&
              dalloc(h, &S.refB2, std::max(h->NBt, 1)) == cudaSuccess && dalloc(h, &S.cref, (size_t)2 * h->NAt + (size_t)2 * h->NBt + 8) == cudaSuccess &&
              dalloc(h, &S.series, 8) == cudaSuccess && dalloc(h, &S.hist, STRIP_HIST_BINS) == cudaSuccess &&
              cudaMallocHost((void **)&S.hostI, 32 * sizeof(int)) == cudaSuccess && cudaMallocHost((void **)&S.hostH, STRIP_HIST_BINS * sizeof(unsigned long long)) == cudaSuccess;
    // band capacity: 1.5 x the share of the local capacity a band of width W is expected to hold (+ slack for small systems);
    // derived from the handle's capacities and the strip geometry only, so that every rank computes the same number
    const double width = h->K.Lx / h->K.strips, f = std::min(1.0, 1.5 * h->strip_W / (width + 2 * h->strip_W));
    const size_t bandRec = std::min<size_t>(h->NAt, (size_t)(h->NAt * f) + 2048), bandLig = std::min<size_t>(h->NBt, (size_t)(h->NBt * f) + 2048);
    S.bandCap = MSG_HDR + bandRec * sizeof(RecMsg) + bandLig * sizeof(LigMsg);
    const size_t full = MSG_HDR + (size_t)h->NAt * sizeof(RecMsg) + (size_t)h->NBt * sizeof(LigMsg);
    for (int k = 0; k < 3 && ok; k++) { S.msgCap[k] = k == 2 ? full : S.bandCap; ok = dalloc(h, &S.msg[k], S.msgCap[k]) == cudaSuccess; }
    for (int k = 0; k < 2 && ok; k++) ok = dalloc(h, &S.rcv[k], S.bandCap) == cudaSuccess;
    if (!ok) { h->err = "strip: device buffer allocation failed"; return KMC_ERR_CUDA; }
    return KMC_OK;
}

// count -> scan -> pack of the flagged lists (bit mask `lists`: 1 to lower x, 2 to higher x, 4 the owned set) into msg[0..2];
// the flags are already set. Counting and headers cover all three lists, whichever are written.
static void strip_pack_lists(kmc_handle *h, cudaStream_t st, int lists, bool count) {
    StripDev &S = *h->strip_dev;
    const Args A{h->D, h->K};
    const int nt = S.ntA + S.ntB;
    MsgBufs M; for (int k = 0; k < 3; k++) { M.p[k] = S.msg[k]; M.cap[k] = S.msgCap[k]; }
    if (count) {
        k_strip_count<<<nt, 256, 0, st>>>(A, S.flag, S.ntA, S.tileCnt);
        k_strip_scan_tiles<<<1, 1024, 0, st>>>(S.tileCnt, S.tileOff, S.ntA, S.ntB, S.dcnt, M, h->D.scal);
    }
    k_strip_pack_all<<<nt, 256, 0, st>>>(A, S.flag, S.ntA, S.tileOff, S.dcnt, M, lists, count ? S.keptRank : nullptr);
}
// Refresh, phase 1 (asynchronous): by-products, complexes of the CURRENT bond table, ownership + bands, the three messages
static int strip_pack(kmc_handle *h) {
    int rc = strip_dev_alloc(h); if (rc) return rc;
    StripDev &S = *h->strip_dev;
    cudaStream_t st = h->stream;
    const Args A{h->D, h->K};
    const int NT = h->NT, NAt = h->NAt, NBt = h->NBt, B = 128;
    CK(cudaMemsetAsync(S.series, 0, 8 * sizeof(int), st));
    CK(cudaMemsetAsync(S.hist, 0, STRIP_HIST_BINS * sizeof(unsigned long long), st));
    k_strip_cx_owned<<<nblk(std::max(NBt, 1), 256), 256, 0, st>>>(A, h->stepped ? 1 : 0, S.series, S.hist);
    // complexes of the CURRENT bond table (the last step's reactions may have changed it)
    LAUNCH(KID_STEP_BEGIN, (k_step_begin<<<1, 32, 0, st>>>(A, 0)));
    launch_cx_rebuild(h, A, st);
    CK(cudaMemsetAsync(S.flag, 0, NT, st));
    k_strip_classify<<<nblk(NT, 128), 128, 0, st>>>(A, S.flag, h->strip_lo, h->strip_hi, h->strip_W, S.budget);
    k_strip_bonds_owned<<<nblk(std::max(NAt, 1), 256), 256, 0, st>>>(A, S.flag, S.series);
    strip_pack_lists(h, st, 3, true);           // the two band messages (the owned molecules stay in the arrays: only their ranks are computed)
    return KMC_OK;
}
// Refresh, phase 2 (asynchronous): the new local set = the units this rank owns + the two incoming lists
static int strip_merge(kmc_handle *h, const char *fromLow, const char *fromHigh) {
    StripDev &S = *h->strip_dev;
    cudaStream_t st = h->stream;
    const Args A{h->D, h->K};
    MergeIn M; M.base[0] = fromLow; M.base[1] = fromHigh; M.cref = S.cref; M.dcnt = S.dcnt; M.keptRank = S.keptRank; M.flag = S.flag;
    M.refA2 = S.refA2; M.refB2 = S.refB2; M.ownFlag = S.ownFlag; M.bond = S.bondRef;
    k_strip_refs<<<nblk(std::max(h->NT, 1), 256), 256, 0, st>>>(A, M);
    const int nbA = nblk(std::max(h->NAt, 1), MERGE_B), nbB = nblk(std::max(h->NBt, 1), MERGE_B);
    k_strip_merge<<<nbA + nbB + nblk(std::max(h->NT, 1), MERGE_B), MERGE_B, 0, st>>>(A, M, nbA, nbB);
    k_strip_fix_bonds<<<nblk(std::max(h->NT, 1), 128), 128, 0, st>>>(A, M);
    CK(cudaGetLastError());
    swap_buffers(h->D); h->parity ^= 1;          // the merged poses were written into the "next" buffers
    h->stepped = false; h->sinceBuild = 0; h->strip_since = 0; h->strip_refreshes++;
    S.fresh = true;
    return KMC_OK;
}

extern "C" int kmc_strip_unique_id(void *id128) {
    if (!id128) return KMC_ERR_INVALID;
    if (const char *e = nccl_load()) { g_create_error = e; return KMC_ERR_CUDA; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return KMC_ERR_CUDA; }
    memcpy(id128, &id, sizeof id);
    return KMC_OK;
}
// One NCCL communicator over the ranks given to kmc_strip_configure (rank 0 creates the id with kmc_strip_unique_id, the caller
// hands the 128 bytes to every rank by whatever means it has). refresh_every > 0: kmc_step refreshes the halos itself every
// refresh_every steps. Collective: every rank must call it.
extern "C" int kmc_strip_comm_init(kmc_handle *h, const void *id128, int32_t refresh_every) {
    if (!h || !h->strip_on || !id128 || refresh_every < 0) { if (h) h->err = "kmc_strip_comm_init: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    if (const char *e = nccl_load()) { h->err = e; return KMC_ERR_CUDA; }
    int rc = strip_dev_alloc(h); if (rc) return rc;
    StripDev &S = *h->strip_dev;
    ncclUniqueId id; memcpy(&id, id128, sizeof id);
    NC(g_nccl.CommInitRank(&S.comm, h->K.strips, id, h->K.stripRank));
    // every rank must move band messages of the same size
    long long *d = reinterpret_cast<long long *>(S.hist);
    long long v[2] = {(long long)S.bandCap, -(long long)S.bandCap};
    CK(cudaMemcpyAsync(d, v, sizeof v, cudaMemcpyHostToDevice, h->stream));
    NC(g_nccl.AllReduce(d, d, 2, ncclInt64, ncclMax, S.comm, h->stream));
    CK(cudaMemcpyAsync(v, d, sizeof v, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (v[0] != -v[1]) { h->err = "kmc_strip_comm_init: ranks were created with different capacities (band messages must have one size)"; return KMC_ERR_INVALID; }
    h->strip_every = refresh_every; h->strip_since = 0;
    if (refresh_every > 0) S.budget = h->strip_W - refresh_every * strip_reach_per_step(h->K);
    if (getenv("KMC_STRIP_TIMING") && h->K.strips > 1) {          // diagnostics: the band exchange alone, CUDA events
        const int n = h->K.strips, r = h->K.stripRank, lo = (r + n - 1) % n, hi = (r + 1) % n;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int it = 0; it < 12; it++) {
            if (it == 2) cudaEventRecord(a, h->stream);
            NC(g_nccl.GroupStart());
            NC(g_nccl.Send(S.msg[0], S.bandCap, ncclChar, lo, S.comm, h->stream)); NC(g_nccl.Send(S.msg[1], S.bandCap, ncclChar, hi, S.comm, h->stream));
            NC(g_nccl.Recv(S.rcv[1], S.bandCap, ncclChar, hi, S.comm, h->stream)); NC(g_nccl.Recv(S.rcv[0], S.bandCap, ncclChar, lo, S.comm, h->stream));
            NC(g_nccl.GroupEnd());
        }
        cudaEventRecord(b, h->stream); cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        if (r == 0) fprintf(stderr, "band exchange alone: %.1f us per round (2 x %zu bytes each way)\n", 1e3 * ms / 10, S.bandCap);
        cudaEventDestroy(a); cudaEventDestroy(b);
    }
    return KMC_OK;
}
// The refresh over NCCL: pack -> one grouped send/recv round with the two x-neighbours -> merge, all on the handle's stream.
extern "C" int kmc_strip_refresh(kmc_handle *h) {
    if (!h || !h->strip_on) { if (h) h->err = "kmc_strip_refresh: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    // KMC_STRIP_TIMING=1 (diagnostics only): synchronise between the phases and accumulate their wall-clock times
    static const bool timing = getenv("KMC_STRIP_TIMING") != nullptr;
    auto now = [&]() { cudaStreamSynchronize(h->stream); return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t0 = timing ? now() : 0;
    int rc = strip_pack(h); if (rc) return rc;
    StripDev &S = *h->strip_dev;
    double t1 = timing ? now() : 0;
    const int n = h->K.strips, r = h->K.stripRank;
    if (n > 1) {
        if (!S.comm) { h->err = "kmc_strip_refresh: no communicator (kmc_strip_comm_init)"; return KMC_ERR_INVALID; }
        const int lo = (r + n - 1) % n, hi = (r + 1) % n;
        cudaStream_t xs = h->stream;
        if (xs != h->stream) { CK(cudaEventRecord(h->evFork[1], h->stream)); CK(cudaStreamWaitEvent(xs, h->evFork[1], 0)); }
        // untagged point-to-point operations between one pair of ranks are matched in posting order: sends [to low, to high],
        // receives [from high, from low] -- with two ranks (both neighbours the same peer) my first send meets its first receive
        NC(g_nccl.GroupStart());
        NC(g_nccl.Send(S.msg[0], S.bandCap, ncclChar, lo, S.comm, xs));
        NC(g_nccl.Send(S.msg[1], S.bandCap, ncclChar, hi, S.comm, xs));
        NC(g_nccl.Recv(S.rcv[1], S.bandCap, ncclChar, hi, S.comm, xs));
        NC(g_nccl.Recv(S.rcv[0], S.bandCap, ncclChar, lo, S.comm, xs));
        NC(g_nccl.GroupEnd());
        if (xs != h->stream) { CK(cudaEventRecord(h->evJoin[3], xs)); CK(cudaStreamWaitEvent(h->stream, h->evJoin[3], 0)); }
    } else { CK(cudaMemsetAsync(S.rcv[0], 0, MSG_HDR, h->stream)); CK(cudaMemsetAsync(S.rcv[1], 0, MSG_HDR, h->stream)); }
    double t2 = timing ? now() : 0;
    rc = strip_merge(h, S.rcv[0], S.rcv[1]);
    if (timing) {
        const double t3 = now();
        h->strip_t[0] += t1 - t0; h->strip_t[1] += t2 - t1; h->strip_t[2] += t3 - t2;
        if (h->strip_refreshes % 20 == 0 && r == 0)
            fprintf(stderr, "strip refresh timing over %lld refreshes (us each): classify+pack %.1f, exchange %.1f (%zu bytes per band), merge %.1f\n",
                    (long long)h->strip_refreshes, h->strip_t[0] / h->strip_refreshes, h->strip_t[1] / h->strip_refreshes, S.bandCap, h->strip_t[2] / h->strip_refreshes);
    }
    return rc;
}
static int strip_auto_refresh(kmc_handle *h) { return kmc_strip_refresh(h); }
// The same refresh between K handles of ONE process (logical ranks 0..K-1, all configured with nranks = K): device-to-device
// copies stand in for NCCL. refresh_every only sets the extent guard (0 = no guard).
extern "C" int kmc_strip_refresh_local(kmc_handle **hs, int32_t n, int32_t refresh_every) {
    if (!hs || n < 1) return KMC_ERR_INVALID;
    for (int i = 0; i < n; i++) {
        kmc_handle *h = hs[i];
        if (!h || !h->strip_on || h->K.strips != n || h->K.stripRank != i) { if (h) h->err = "kmc_strip_refresh_local: handle i must be rank i of n"; return KMC_ERR_INVALID; }
        CK(cudaSetDevice(h->P.device));
        int rc = strip_dev_alloc(h); if (rc) return rc;
        h->strip_dev->budget = refresh_every > 0 ? h->strip_W - refresh_every * strip_reach_per_step(h->K) : INFINITY;
        rc = strip_pack(h); if (rc) return rc;
    }
    for (int i = 0; i < n; i++) { kmc_handle *h = hs[i]; CK(cudaStreamSynchronize(h->stream)); }
    for (int i = 0; i < n; i++) {
        kmc_handle *h = hs[i]; StripDev &S = *h->strip_dev;
        if (n > 1) {
            StripDev &L = *hs[(i + n - 1) % n]->strip_dev, &H = *hs[(i + 1) % n]->strip_dev;
            if (L.bandCap != S.bandCap || H.bandCap != S.bandCap) { h->err = "kmc_strip_refresh_local: handles have different capacities"; return KMC_ERR_INVALID; }
            CK(cudaMemcpyAsync(S.rcv[0], L.msg[1], S.bandCap, cudaMemcpyDeviceToDevice, h->stream));      // what my lower neighbour sent upwards
            CK(cudaMemcpyAsync(S.rcv[1], H.msg[0], S.bandCap, cudaMemcpyDeviceToDevice, h->stream));      // what my upper neighbour sent downwards
        } else { CK(cudaMemsetAsync(S.rcv[0], 0, MSG_HDR, h->stream)); CK(cudaMemsetAsync(S.rcv[1], 0, MSG_HDR, h->stream)); }
        int rc = strip_merge(h, S.rcv[0], S.rcv[1]); if (rc) return rc;
    }
    for (int i = 0; i < n; i++) { kmc_handle *h = hs[i]; CK(cudaStreamSynchronize(h->stream)); }
    return KMC_OK;
}

// ---- global outputs of a strip-decomposed membrane (main.cpp:2247-2253, 2291-2305) ------------------------------------------
// Every complex is counted once, by the rank that owns its root ligand; every bond by the rank that owns the receptor's unit.
// reduce != 0: all-reduced over the communicator (every rank calls, every rank gets the row of the WHOLE membrane);
// reduce == 0: this rank's owned-only part (in-process ranks: the caller adds them up).
static int strip_fresh(kmc_handle *h, const char *who) {
    if (!h || !h->strip_on) { if (h) h->err = std::string(who) + ": configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    if (h->strip_dev && h->strip_dev->fresh) return KMC_OK;
    if (h->K.strips > 1 && !(h->strip_dev && h->strip_dev->comm)) { h->err = std::string(who) + ": the values are those of the last refresh; refresh first"; return KMC_ERR_INVALID; }
    return kmc_strip_refresh(h);            // refreshing more often than scheduled is always exact
}
extern "C" int kmc_strip_get_series(kmc_handle *h, int32_t reduce, kmc_series *out) {
    if (!out) return KMC_ERR_INVALID;
    const bool hadStep = h && h->stepped;
    int rc = strip_fresh(h, "kmc_strip_get_series"); if (rc) return rc;
    (void)hadStep;
    StripDev &S = *h->strip_dev; cudaStream_t st = h->stream;
    int *d = S.series;
    if (reduce && h->K.strips > 1) {
        if (!S.comm) { h->err = "kmc_strip_get_series: no communicator"; return KMC_ERR_INVALID; }
        int *tmp = reinterpret_cast<int *>(S.tileCnt);          // scratch (free between refreshes)
        CK(cudaMemcpyAsync(tmp, S.series, 8 * sizeof(int), cudaMemcpyDeviceToDevice, st));
        NC(g_nccl.AllReduce(tmp, tmp, 5, ncclInt32, ncclSum, S.comm, st));
        NC(g_nccl.AllReduce(tmp + 7, tmp + 7, 1, ncclInt32, ncclMax, S.comm, st));
        d = tmp;
    }
    CK(cudaMemcpyAsync(S.hostI + 16, d, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int *v = S.hostI + 16;
    memset(out, 0, sizeof *out);
    out->step = h->step_done; out->bond_num_rl = v[0]; out->bond_num_mono_cis = v[1]; out->bond_num_cis = v[2]; out->bond_num = v[0] + v[1] + v[2];
    out->n_complexes = v[3]; out->n_in_complexes = v[4]; out->max_complex = v[7];
    if (out->n_complexes) out->cluster_size = (double)out->n_in_complexes / out->n_complexes;     // main.cpp:2200-2202
    return KMC_OK;
}
extern "C" int kmc_strip_get_oligomer_hist(kmc_handle *h, int32_t reduce, int64_t *hist, int32_t nbins) {
    if (!hist || nbins < 2) return KMC_ERR_INVALID;
    int rc = strip_fresh(h, "kmc_strip_get_oligomer_hist"); if (rc) return rc;
    StripDev &S = *h->strip_dev; cudaStream_t st = h->stream;
    unsigned long long *d = S.hist;
    if (reduce && h->K.strips > 1) {
        if (!S.comm) { h->err = "kmc_strip_get_oligomer_hist: no communicator"; return KMC_ERR_INVALID; }
        unsigned long long *tmp = reinterpret_cast<unsigned long long *>(S.tileOff);      // scratch; 3*(tiles)+3 ints >= 512 ints? checked below
        if ((size_t)3 * (S.ntA + S.ntB) + 3 < 2 * STRIP_HIST_BINS) tmp = reinterpret_cast<unsigned long long *>(S.bondRef);
        CK(cudaMemcpyAsync(tmp, S.hist, STRIP_HIST_BINS * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        NC(g_nccl.AllReduce(tmp, tmp, STRIP_HIST_BINS, ncclUint64, ncclSum, S.comm, st));
        d = tmp;
    }
    CK(cudaMemcpyAsync(S.hostH, d, STRIP_HIST_BINS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < nbins; i++) hist[i] = 0;
    for (int i = 0; i < STRIP_HIST_BINS; i++) hist[std::min(i, nbins - 1)] += (int64_t)S.hostH[i];
    return KMC_OK;
}

// ---- per-rank state exchange with HOST buffers (the e2e path of a strip run: every rank moves only its own slab) ------------
// which = 2: the units this rank owns (the union over ranks is the membrane, every molecule once);
// which = 3: everything this rank holds (owned + halo copies): what kmc_strip_load_records takes to restore the rank.
// host_buf receives RecMsg[n_rec] then LigMsg[n_lig] (id-sorted), pinned memory recommended.
extern "C" int kmc_strip_get_records(kmc_handle *h, int32_t which, void *host_buf, int64_t cap_bytes, int64_t *n_rec, int64_t *n_lig) {
    if (!h || !h->strip_on || !host_buf || (which != 2 && which != 3)) { if (h) h->err = "kmc_strip_get_records: bad argument"; return KMC_ERR_INVALID; }
    int rc;
    if (which == 2) {
        rc = strip_fresh(h, "kmc_strip_get_records"); if (rc) return rc;
        const Args A{h->D, h->K};
        k_strip_flag_owned<<<nblk(std::max(h->NT, 1), 256), 256, 0, h->stream>>>(A, h->strip_dev->ownFlag, h->strip_dev->flag);
        strip_pack_lists(h, h->stream, 4, true);
    } else {
        CK(cudaSetDevice(h->P.device));
        rc = strip_dev_alloc(h); if (rc) return rc;
        const Args A{h->D, h->K};
        k_strip_flag_all<<<nblk(std::max(h->NT, 1), 256), 256, 0, h->stream>>>(A, h->strip_dev->flag);
        strip_pack_lists(h, h->stream, 4, true);
    }
    StripDev &S = *h->strip_dev; cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(S.hostI, S.msg[2], MSG_HDR, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int64_t nr = S.hostI[0], nl = S.hostI[1], bytes = nr * (int64_t)sizeof(RecMsg) + nl * (int64_t)sizeof(LigMsg);
    if (S.hostI[2]) { h->err = "kmc_strip_get_records: message buffer overflow"; return KMC_ERR_CAPACITY; }
    if (bytes > cap_bytes) { h->err = "kmc_strip_get_records: host buffer too small (" + std::to_string(bytes) + " bytes needed)"; return KMC_ERR_CAPACITY; }
    if (bytes) CK(cudaMemcpyAsync(host_buf, S.msg[2] + MSG_HDR, (size_t)bytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_rec) *n_rec = nr; if (n_lig) *n_lig = nl;
    return KMC_OK;
}
extern "C" int kmc_strip_load_records(kmc_handle *h, const void *host_buf, int64_t n_rec, int64_t n_lig, int64_t step_done) {
    if (!h || !h->strip_on || !host_buf || n_rec < 0 || n_lig < 0) { if (h) h->err = "kmc_strip_load_records: bad argument"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    int rc = strip_dev_alloc(h); if (rc) return rc;
    if (n_rec > h->NAt || n_lig > h->NBt) { h->err = "kmc_strip_load_records: more molecules than the handle's capacity"; return KMC_ERR_CAPACITY; }
    StripDev &S = *h->strip_dev; cudaStream_t st = h->stream;
    CK(cudaStreamSynchronize(st));
    memset(S.hostI, 0, MSG_HDR); S.hostI[0] = (int)n_rec; S.hostI[1] = (int)n_lig;
    unsigned long long *s64 = reinterpret_cast<unsigned long long *>(S.hostI + 24); *s64 = (unsigned long long)step_done;
    CK(cudaMemcpyAsync(S.msg[2], S.hostI, MSG_HDR, cudaMemcpyHostToDevice, st));
    const size_t bytes = (size_t)n_rec * sizeof(RecMsg) + (size_t)n_lig * sizeof(LigMsg);
    if (bytes) CK(cudaMemcpyAsync(S.msg[2] + MSG_HDR, host_buf, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(S.rcv[1], 0, MSG_HDR, st));
    CK(cudaMemsetAsync(S.dcnt, 0, 8 * sizeof(int), st));                       // nothing is kept: the records are the whole new set
    CK(cudaMemsetAsync(h->D.scal + S_NA_LIVE, 0, 2 * sizeof(int), st));
    CK(cudaMemcpyAsync(h->D.step64, s64, sizeof *s64, cudaMemcpyHostToDevice, st));
    rc = strip_merge(h, S.msg[2], S.rcv[1]); if (rc) return rc;
    S.fresh = false;                          // by-products describe nothing yet
    h->step_done = step_done; h->strip_refreshes--;
    CK(cudaStreamSynchronize(st));
    return KMC_OK;
}

// ---- start state of a strip run, generated on the GPU: every rank generates the SAME global configuration (n_rec + n_lig
// molecules in the handle's box, csrc/kmc_init.cu -- a few ms even for 1e7 molecules) in scratch device memory and keeps the
// molecules within reach of its strip (owned + halo). Nothing crosses PCIe, nothing is exchanged.
__global__ void k_strip_select_flags(const __grid_constant__ Args A, const double *rec, const double *lig, int nR, int nL, double lo, double hi, double W, int *flagA, int *flagB) {
    KARGS
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nR) { const double t = hash_x(cK, rec[(size_t)i * 6]); flagA[i] = (t >= lo - W && t < hi + W) ? 1 : 0; }
    else if (i < nR + nL) { const int b = i - nR; const double t = hash_x(cK, lig[(size_t)b * 24]); flagB[b] = (t >= lo - W && t < hi + W) ? 1 : 0; }
}
__global__ void k_strip_select_scatter(const __grid_constant__ Args A, const double *rec, const double *lig, int nR, int nL, double lo, double hi, double W, const int *offA, const int *offB) {
    KARGS
    const Consts &K = cK;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nR) {
        const double *o = rec + (size_t)i * 6;
        const double t = hash_x(K, o[0]);
        if (!(t >= lo - W && t < hi + W)) return;
        const int d = offA[i];
        if (d >= K.NAt) return;
        D.recC[d] = make_double2(o[0], o[1]); D.recS2[d] = make_double2(o[2], o[3]); D.recS3[d] = make_double2(o[4], o[5]);
        D.refA[d] = (unsigned)(i + 1); D.recLig[d] = -1; D.recSite[d] = -1; D.recCis[d] = -1;
    } else if (i < nR + nL) {
        const int b = i - nR;
        const double *o = lig + (size_t)b * 24;
        const double t = hash_x(K, o[0]);
        if (!(t >= lo - W && t < hi + W)) return;
        const int d = offB[b];
        if (d >= K.NBt) return;
        double *p = D.lig + (size_t)d * 24;
        for (int q = 0; q < 24; q++) p[q] = o[q];
        D.refB[d] = (unsigned)(nR + b + 1);
        for (int q = 0; q < 3; q++) D.ligRec[d * 3 + q] = -1;
    }
}
extern "C" int kmc_strip_init_random(kmc_handle *h, int32_t n_rec, int32_t n_lig, uint64_t seed, int32_t sort_cells) {
    if (!h || !h->strip_on || n_rec < 0 || n_lig < 1) { if (h) h->err = "kmc_strip_init_random: configure strips first"; return KMC_ERR_INVALID; }
    CK(cudaSetDevice(h->P.device));
    cudaStream_t st = h->stream;
    CK(cudaStreamSynchronize(st));
    double *rec = nullptr, *lig = nullptr; int *fl[2] = {nullptr, nullptr}, *off[2] = {nullptr, nullptr}, *tmp = nullptr;
    const size_t sbA = ((size_t)n_rec + 1 + SCAN_TILE - 1) / SCAN_TILE, sbB = ((size_t)n_lig + 1 + SCAN_TILE - 1) / SCAN_TILE;
    bool ok = cudaMalloc(&rec, sizeof(double) * 6 * (size_t)std::max(n_rec, 1)) == cudaSuccess && cudaMalloc(&lig, sizeof(double) * 24 * (size_t)n_lig) == cudaSuccess &&
              cudaMalloc(&fl[0], sizeof(int) * sbA * SCAN_TILE) == cudaSuccess && cudaMalloc(&off[0], sizeof(int) * sbA * SCAN_TILE) == cudaSuccess &&
              cudaMalloc(&fl[1], sizeof(int) * sbB * SCAN_TILE) == cudaSuccess && cudaMalloc(&off[1], sizeof(int) * sbB * SCAN_TILE) == cudaSuccess &&
              cudaMalloc(&tmp, sizeof(int) * (std::max(sbA, sbB) + 1)) == cudaSuccess;
    int rc = KMC_OK, tot[2] = {0, 0};
    if (!ok) { h->err = "kmc_strip_init_random: scratch allocation failed"; rc = KMC_ERR_CUDA; }
    if (!rc) if (const char *msg = generate_random_device(h->P, n_rec, n_lig, 1, seed, sort_cells, rec, lig, st, &h->init_rounds)) { h->err = msg; rc = KMC_ERR_INVALID; }
    if (!rc) {
        const Args A{h->D, h->K};
        const int n = n_rec + n_lig;
        cudaMemsetAsync(fl[0], 0, sizeof(int) * sbA * SCAN_TILE, st); cudaMemsetAsync(fl[1], 0, sizeof(int) * sbB * SCAN_TILE, st);
        k_strip_select_flags<<<nblk(n, 256), 256, 0, st>>>(A, rec, lig, n_rec, n_lig, h->strip_lo, h->strip_hi, h->strip_W, fl[0], fl[1]);
        const size_t sb[2] = {sbA, sbB};
        for (int s = 0; s < 2; s++) {
            k_scan_reduce<<<(unsigned)sb[s], 256, 0, st>>>((const int4 *)fl[s], tmp);
            k_scan_sums<<<1, 1024, 0, st>>>(tmp, (int)sb[s]);
            k_scan_down<<<(unsigned)sb[s], 256, 0, st>>>((int4 *)fl[s], tmp, (int4 *)off[s]);
        }
        k_strip_select_scatter<<<nblk(n, 256), 256, 0, st>>>(A, rec, lig, n_rec, n_lig, h->strip_lo, h->strip_hi, h->strip_W, off[0], off[1]);
        cudaMemcpyAsync(&tot[0], off[0] + n_rec, sizeof(int), cudaMemcpyDeviceToHost, st);        // exclusive prefix at index n = the total
        cudaMemcpyAsync(&tot[1], off[1] + n_lig, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { h->err = "kmc_strip_init_random: CUDA error"; rc = KMC_ERR_CUDA; }
    }
    cudaFree(rec); cudaFree(lig); cudaFree(fl[0]); cudaFree(fl[1]); cudaFree(off[0]); cudaFree(off[1]); cudaFree(tmp);
    if (rc) return rc;
    if (tot[0] > h->NAt || tot[1] > h->NBt) {
        h->err = "strip: local capacity exceeded (" + std::to_string(tot[0]) + "/" + std::to_string(h->NAt) + " receptors, " + std::to_string(tot[1]) + "/" +
                 std::to_string(h->NBt) + " ligands): create the handle with larger n_receptor/n_ligand"; return KMC_ERR_CAPACITY;
    }
    int one = 1;
    CK(cudaMemcpy(h->D.scal + S_NA_LIVE, tot, sizeof tot, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->D.scal + S_TOPO_DIRTY, &one, sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(h->D.step64, 0, sizeof(unsigned long long)));
    CK(cudaMemset(h->D.maxComplex, 0, sizeof(int)));
    h->step_done = 0; h->stepped = false; h->sinceBuild = 0; h->strip_since = 0;
    if (h->strip_dev) h->strip_dev->fresh = false;
    return KMC_OK;
}
