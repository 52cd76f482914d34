// oracle/kmc_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement of the reference's per-timestep sweep, /root/reference/main.cpp:461-2202 (plus the
// initial-configuration generator 273-456 and helpers 2329-2371), with run-time sizes so it can
// serve as the oracle at 1e5 molecules where the reference itself cannot run (its `results`
// matrix is (N+1)^2 ints, main.cpp:124-127).
//
// Pinning (see tests/test_oracle_vs_reference.py, tests/golden/): this file is proven BIT-EQUAL
// (positions, bond table, counters, draw counts) to the unmodified reference driven by
// oracle/ref_harness.cpp under identical sequential streams, at N=200 default, the dense
// oligomerising regime of SURVEY 8c (every branch incl. `goto lable4` fires) and N=2000; the
// committed golden vectors were produced by the reference itself.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
//
// Every arithmetic expression keeps the reference's operand order and association (no FMA:
// built with -ffp-contract=off and no -march), because replay parity is bit-level.
#include "kmc_oracle.h"
#include "philox.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <vector>

namespace {

struct Pt { double x, y, z; };
// one molecule's explicit points, indexed [bead j 1..4][site k 1..4] like R_*[i][j][k] (main.cpp:102-104)
struct Body { Pt p[5][5]; };

struct Rot { double t[3][3]; };

// Euler matrix exactly as written at main.cpp:613-623 (and 332-342, 728-738, 946-956, 1091-1101).
static Rot euler(double theta, double phi, double psai) {
    Rot r;
    r.t[0][0] = cos(psai) * cos(phi) - cos(theta) * sin(phi) * sin(psai);
    r.t[0][1] = -sin(psai) * cos(phi) - cos(theta) * sin(phi) * cos(psai);
    r.t[0][2] = sin(theta) * sin(phi);
    r.t[1][0] = cos(psai) * sin(phi) + cos(theta) * cos(phi) * sin(psai);
    r.t[1][1] = -sin(psai) * sin(phi) + cos(theta) * cos(phi) * cos(psai);
    r.t[1][2] = -sin(theta) * cos(phi);
    r.t[2][0] = sin(psai) * sin(theta);
    r.t[2][1] = cos(psai) * sin(theta);
    r.t[2][2] = cos(theta);
    return r;
}

// q = t*(s - c) + c, operand order of main.cpp:631-633
static inline Pt rot_about(const Rot &r, const Pt &s, const Pt &c) {
    Pt q;
    q.x = r.t[0][0] * (s.x - c.x) + r.t[0][1] * (s.y - c.y) + r.t[0][2] * (s.z - c.z) + c.x;
    q.y = r.t[1][0] * (s.x - c.x) + r.t[1][1] * (s.y - c.y) + r.t[1][2] * (s.z - c.z) + c.y;
    q.z = r.t[2][0] * (s.x - c.x) + r.t[2][1] * (s.y - c.y) + r.t[2][2] * (s.z - c.z) + c.z;
    return q;
}

static inline double dist3(const Pt &a, const Pt &b) {
    return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y) + (a.z - b.z) * (a.z - b.z));
}
static inline double dist2d(const Pt &a, const Pt &b) {
    return sqrt((a.x - b.x) * (a.x - b.x) + (a.y - b.y) * (a.y - b.y));
}
static inline bool are_same(double a, double b) { return fabs(a - b) < 1.0E-8; }  // main.cpp:171, 2368-2371

// main.cpp:2329-2366 with point[1] = 0 at every call site (1892-1894 ...): angle in degrees between
// -p0 and p2 using conv = 180/3.14159; note the x-term is written lx[1]*lx[0] (2355).
static double angle_deg(const Pt &p0, const Pt &p2) {
    double lx0 = 0 - p0.x, ly0 = 0 - p0.y, lz0 = 0 - p0.z;
    double lr0 = sqrt(lx0 * lx0 + ly0 * ly0 + lz0 * lz0);
    double lx1 = p2.x - 0, ly1 = p2.y - 0, lz1 = p2.z - 0;
    double lr1 = sqrt(lx1 * lx1 + ly1 * ly1 + lz1 * lz1);
    double conv = 180 / 3.14159;
    double doth1 = -(lx1 * lx0 + ly0 * ly1 + lz0 * lz1);
    double doth2 = doth1 / (lr1 * lr0);
    if (doth2 > 1) doth2 = 1;
    if (doth2 < -1) doth2 = -1;
    return acos(doth2) * conv;
}

struct Oracle {
    kmco_params P;
    int NA, NB, N;
    std::vector<Body> R, Rn, T;                    // committed, new, translated scratch (R_*, R_*_new, R_*_new0)
    std::vector<int> st, stn, nb, nbn;             // protein_status[N+1][5], res_nei[N+1][7] and *_new
    std::vector<int> visited, moved;
    std::vector<std::vector<int>> rows;            // rows[l] = results[NA+1+l][1..] (0-based inside)
    std::vector<int> accepted;
    int bond_num = 0, bond_rl = 0, bond_cis = 0, bond_mono = 0, max_complex = 0;
    int bond_num_n = 0, bond_rl_n = 0, bond_cis_n = 0, bond_mono_n = 0;
    int tot_cluster_num = 0, tot_proteins_in_cluster = 0;
    double cluster_size = 0.0;
    int64_t step_done = 0, cur_step = 0;
    uint64_t s2, sr, n_rand2 = 0, n_rand = 0;
    int64_t ev[16] = {0};

    // ---- O(N) neighbour grid (dynamic bins on the xy centre of R_new); results are order independent ----
    double ce = 0; int ncx = 0, ncy = 0; double gx0 = 0, gy0 = 0;
    std::vector<std::vector<int>> bins; std::vector<int> bin_of;

    int &S(std::vector<int> &v, int i, int j) { return v[(size_t)i * 5 + j]; }
    int &NBR(std::vector<int> &v, int i, int j) { return v[(size_t)i * 7 + j]; }

    explicit Oracle(const kmco_params &p) : P(p) {
        NA = p.n_receptor; NB = p.n_ligand; N = NA + NB;
        Body zero; memset(&zero, 0, sizeof zero);
        R.assign(N + 1, zero); Rn.assign(N + 1, zero); T.assign(N + 1, zero);
        st.assign((size_t)(N + 1) * 5, 0); stn = st; nb.assign((size_t)(N + 1) * 7, 0); nbn = nb;
        visited.assign(N + 1, 0); moved.assign(N + 1, 0); accepted.assign(N + 1, 1);
        rows.assign(NB, {});
        s2 = p.rand2_state; sr = p.rand_state;
    }

    // ---- random streams ----
    static inline uint64_t xs64(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
    double draw(int mol, int partner, int slot) {
        n_rand2++;
        if (P.stream_mode == 0) return (double)(xs64(s2) >> 11) * (1.0 / 9007199254740992.0);
        return kmco::keyed_uniform(P.seed, (uint32_t)mol, (uint32_t)partner, (uint64_t)cur_step, (uint32_t)slot);
    }
    int irand(int root, uint32_t &count) {
        n_rand++;
        if (P.stream_mode == 0) return (int)((xs64(sr) >> 33) & 0x7fffffff);
        return kmco::keyed_rand31(P.seed, (uint32_t)root, count++, (uint64_t)cur_step);
    }
    // libstdc++ std::random_shuffle(first,last) with rand(): for i in [first+1,last): swap(i, first + rand()%((i-first)+1)).
    // The reference passes (&row[1], &row[size]) i.e. the LAST member never moves (main.cpp:1285).
    void shuffle_row(std::vector<int> &row, int root, uint32_t &count) {
        int n = (int)row.size() - 1;  // elements taking part
        for (int i = 1; i < n; i++) {
            int j = irand(root, count) % (i + 1);
            if (i != j) std::swap(row[i], row[j]);
        }
    }

    // ---- grid ----
    void grid_setup() {
        double rs = P.rB * 2 / sqrt(3.0);
        double need = 2 * P.rB + 2 * rs + 1.0;                                  // ligand-ligand bead overlap reach
        need = std::max(need, P.rA + P.bond_dist_cut + rs + P.rB + 1.0);        // R-L association reach
        ce = std::max(need, 150.0);
        double margin = 4 * ce;
        gx0 = -P.box[0] / 2 - margin; gy0 = -P.box[1] / 2 - margin;
        ncx = (int)ceil((P.box[0] + 2 * margin) / ce); ncy = (int)ceil((P.box[1] + 2 * margin) / ce);
        bins.assign((size_t)ncx * ncy, {}); bin_of.assign(N + 1, -1);
    }
    int cell_of(double x, double y) const {
        int cx = (int)floor((x - gx0) / ce), cy = (int)floor((y - gy0) / ce);
        cx = std::min(std::max(cx, 0), ncx - 1); cy = std::min(std::max(cy, 0), ncy - 1);
        return cy * ncx + cx;
    }
    void grid_put(int id) {
        int c = cell_of(Rn[id].p[1][1].x, Rn[id].p[1][1].y);
        if (bin_of[id] == c) return;
        if (bin_of[id] >= 0) { auto &b = bins[bin_of[id]]; b.erase(std::find(b.begin(), b.end(), id)); }
        bins[c].push_back(id); bin_of[id] = c;
    }
    void grid_fill() { for (int i = 1; i <= N; i++) grid_put(i); }
    template <class F> void grid_near(double x, double y, double r, F f) const {
        int c = cell_of(x, y), cx = c % ncx, cy = c / ncx, w = (int)ceil(r / ce);
        for (int yy = std::max(cy - w, 0); yy <= std::min(cy + w, ncy - 1); yy++)
            for (int xx = std::max(cx - w, 0); xx <= std::min(cx + w, ncx - 1); xx++)
                for (int id : bins[(size_t)yy * ncx + xx]) f(id);
    }

    // ---- overlap tests (main.cpp:640-664, 806-849, 1768-1826) ----
    bool rec_hits_rec(int a, int i) const { return dist3(Rn[i].p[1][1], Rn[a].p[1][1]) < P.rA + P.rA; }
    bool rec_hits_lig(int a, int l) const {
        for (int j = 2; j <= 4; j++)
            for (int k = 1; k <= 4; k++)
                if (dist3(Rn[l].p[j][1], Rn[a].p[k][1]) < P.rA + P.rB) return true;
        return false;
    }
    bool lig_hits_lig(int b, int i) const {
        for (int j = 2; j <= 4; j++)
            for (int k = 2; k <= 4; k++)
                if (dist3(Rn[i].p[j][1], Rn[b].p[k][1]) < P.rB + P.rB) return true;
        return false;
    }
    double reach_RL() const { return P.rA + P.rB + P.rB * 2 / sqrt(3.0) + 1e-3; }
    double reach_LL() const { return 2 * P.rB + 2 * (P.rB * 2 / sqrt(3.0)) + 1e-3; }

    // does receptor a (at R_new) overlap anything else? `unit` lists the co-moving members, which the
    // grid may still hold at their pre-move bins, so they are tested directly.
    bool receptor_collides(int a, const int *unit, int nunit) {
        bool hit = false;
        if (!P.use_grid) {
            for (int i = 1; i <= NA; i++) if (i != a && rec_hits_rec(a, i)) hit = true;
            for (int i = NA + 1; i <= N; i++) if (rec_hits_lig(a, i)) hit = true;
            return hit;
        }
        auto in_unit = [&](int id) { for (int u = 0; u < nunit; u++) if (unit[u] == id) return true; return false; };
        for (int u = 0; u < nunit; u++) {
            int id = unit[u];
            if (id <= NA) { if (id != a && rec_hits_rec(a, id)) hit = true; }
            else if (rec_hits_lig(a, id)) hit = true;
        }
        grid_near(Rn[a].p[1][1].x, Rn[a].p[1][1].y, std::max(2 * P.rA + 1e-3, reach_RL()), [&](int id) {
            if (hit || id == a || in_unit(id)) return;
            if (id <= NA) { if (rec_hits_rec(a, id)) hit = true; }
            else if (rec_hits_lig(a, id)) hit = true;
        });
        return hit;
    }
    bool ligand_collides(int b, const int *unit, int nunit) {
        bool hit = false;
        if (!P.use_grid) {
            for (int i = NA + 1; i <= N; i++) if (i != b && lig_hits_lig(b, i)) hit = true;
            for (int i = 1; i <= NA; i++) if (rec_hits_lig(i, b)) hit = true;
            return hit;
        }
        auto in_unit = [&](int id) { for (int u = 0; u < nunit; u++) if (unit[u] == id) return true; return false; };
        for (int u = 0; u < nunit; u++) {
            int id = unit[u];
            if (id > NA) { if (id != b && lig_hits_lig(b, id)) hit = true; }
            else if (rec_hits_lig(id, b)) hit = true;
        }
        grid_near(Rn[b].p[1][1].x, Rn[b].p[1][1].y, reach_LL(), [&](int id) {
            if (hit || id == b || in_unit(id)) return;
            if (id > NA) { if (lig_hits_lig(b, id)) hit = true; }
            else if (rec_hits_lig(id, b)) hit = true;
        });
        return hit;
    }

    // ---- rigid-body building blocks ----
    // main.cpp:298-316: receptor template around (x,y,0): beads stacked in z, sites +R x, -R x, +R z
    void receptor_template(Body &b, double x, double y, double z) const {
        for (int j = 1; j <= 4; j++) {
            double zz = z + (j * 2 - 2) * P.rA;
            b.p[j][1] = {x, y, zz};
            b.p[j][2] = {x + P.rA, y, zz};
            b.p[j][3] = {x - P.rA, y, zz};
            b.p[j][4] = {x, y, z + (j * 2 - 1) * P.rA};
        }
    }
    // main.cpp:386-412 (3-D, around a centre) and 1157-1179 / 1453-1475 (xy ghost around the origin)
    void ligand_template(Body &b, double x, double y, double z) const {
        double rB = P.rB;
        b.p[1][1] = {x, y, z};
        b.p[1][2] = {x, y, z + rB};
        b.p[2][1] = {x, y + rB * 2 / sqrt(3), z};
        b.p[3][1] = {x - rB, y - rB / sqrt(3), z};
        b.p[4][1] = {x + rB, y - rB / sqrt(3), z};
        b.p[2][2] = {x, y + rB * (2 / sqrt(3) + 1), z};
        b.p[3][2] = {x - rB * (sqrt(3) / 2 + 1), y - rB / sqrt(3) - rB / 2, z};
        b.p[4][2] = {x + rB * (sqrt(3) / 2 + 1), y - rB / sqrt(3) - rB / 2, z};
    }
    void ligand_ghost(Body &g) const {
        double rB = P.rB;
        memset(&g, 0, sizeof g);
        g.p[2][1].y = rB * 2 / sqrt(3);
        g.p[2][2].y = rB * (2 / sqrt(3) + 1);
        g.p[3][1].x = -rB;                      g.p[3][1].y = -rB / sqrt(3);
        g.p[3][2].x = -rB * (sqrt(3) / 2 + 1);  g.p[3][2].y = -rB / sqrt(3) - rB / 2;
        g.p[4][1].x = rB;                       g.p[4][1].y = -rB / sqrt(3);
        g.p[4][2].x = rB * (sqrt(3) / 2 + 1);   g.p[4][2].y = -rB / sqrt(3) - rB / 2;
    }
    // main.cpp:1184-1189, 1496-1501: xy of every ligand point = Rz(angle) * ghost + (cx, cy)
    void seat_ligand(int b, const Body &g, double angle, double cx, double cy) {
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 2; k++) {
                Rn[b].p[j][k].x = g.p[j][k].x * cos(angle) - g.p[j][k].y * sin(angle) + cx;
                Rn[b].p[j][k].y = g.p[j][k].x * sin(angle) + g.p[j][k].y * cos(angle) + cy;
            }
    }
    // misalignment predicates (main.cpp:1205-1215 / 1245-1255 and their repeats)
    bool rl_misaligned(int b, int j, int a) const {
        double d2 = dist2d(Rn[b].p[j][2], Rn[a].p[3][2]);
        double d1 = dist2d(Rn[b].p[j][1], Rn[a].p[3][1]);
        return !are_same(d1, P.bond_dist_cut / 2 + P.rA + P.rB) || !are_same(d2, P.bond_dist_cut / 2);
    }
    bool cis_misaligned(int a1, int a2) const {
        double d2 = dist2d(Rn[a1].p[3][3], Rn[a2].p[3][3]);
        double d1 = dist2d(Rn[a1].p[3][1], Rn[a2].p[3][1]);
        return !are_same(d1, P.cis_dist_cut / 2 + P.rA + P.rA) || !are_same(d2, P.cis_dist_cut / 2);
    }
    // main.cpp:1216-1228: rebuild receptor a (xy only) on the bead->site axis of ligand b, site j
    void snap_receptor_to_ligand(int a, int b, int j) {
        double ux = Rn[b].p[j][2].x - Rn[b].p[j][1].x, uy = Rn[b].p[j][2].y - Rn[b].p[j][1].y;
        double sx = Rn[b].p[j][2].x, sy = Rn[b].p[j][2].y;
        for (int k = 1; k <= 4; k++) {
            Rn[a].p[k][1].x = (P.bond_dist_cut / 2 + P.rA) / P.rB * ux + sx;
            Rn[a].p[k][1].y = (P.bond_dist_cut / 2 + P.rA) / P.rB * uy + sy;
            Rn[a].p[k][4].x = (P.bond_dist_cut / 2 + P.rA) / P.rB * ux + sx;
            Rn[a].p[k][4].y = (P.bond_dist_cut / 2 + P.rA) / P.rB * uy + sy;
            Rn[a].p[k][3].x = (P.bond_dist_cut / 2 + 2 * P.rA) / P.rB * ux + sx;
            Rn[a].p[k][3].y = (P.bond_dist_cut / 2 + 2 * P.rA) / P.rB * uy + sy;
            Rn[a].p[k][2].x = (P.bond_dist_cut / 2) / P.rB * ux + sx;
            Rn[a].p[k][2].y = (P.bond_dist_cut / 2) / P.rB * uy + sy;
        }
    }
    // main.cpp:786-798, 1256-1268: rebuild `dst` (xy only) from the bead-3 axis (centre -> site 3) of `src`
    void snap_cis(int dst, int src) {
        double ux = Rn[src].p[3][3].x - Rn[src].p[3][1].x, uy = Rn[src].p[3][3].y - Rn[src].p[3][1].y;
        double sx = Rn[src].p[3][3].x, sy = Rn[src].p[3][3].y;
        for (int k = 1; k <= 4; k++) {
            Rn[dst].p[k][1].x = (P.cis_dist_cut / 2 + P.rA) / P.rA * ux + sx;
            Rn[dst].p[k][1].y = (P.cis_dist_cut / 2 + P.rA) / P.rA * uy + sy;
            Rn[dst].p[k][4].x = (P.cis_dist_cut / 2 + P.rA) / P.rA * ux + sx;
            Rn[dst].p[k][4].y = (P.cis_dist_cut / 2 + P.rA) / P.rA * uy + sy;
            Rn[dst].p[k][3].x = (P.cis_dist_cut / 2) / P.rA * ux + sx;
            Rn[dst].p[k][3].y = (P.cis_dist_cut / 2) / P.rA * uy + sy;
            Rn[dst].p[k][2].x = (P.cis_dist_cut / 2 + 2 * P.rA) / P.rA * ux + sx;
            Rn[dst].p[k][2].y = (P.cis_dist_cut / 2 + 2 * P.rA) / P.rA * uy + sy;
        }
    }
    int npts(int id) const { return id <= NA ? 4 : 2; }   // sites per bead: RB_A_res_num / RB_B_res_num

    // ---- initial configuration, main.cpp:273-456 (sequential stream) ----
    void init_reference() {
        double pai = P.pai;
        for (int i = 1; i <= NA; i++) {
            double x, y;
            for (;;) {
                x = draw(i, 0, 0) * P.box[0] - P.box[0] / 2;
                y = draw(i, 0, 1) * P.box[1] - P.box[1] / 2;
                bool clash = false;
                for (int j = 1; j <= i - 1 && !clash; j++) {
                    double d = sqrt((x - R[j].p[1][1].x) * (x - R[j].p[1][1].x) + (y - R[j].p[1][1].y) * (y - R[j].p[1][1].y));
                    if (d <= P.rA + P.rA) clash = true;
                }
                if (!clash) break;
            }
            Body t0; receptor_template(t0, x, y, 0);
            for (int j = 1; j <= 4; j++) R[i].p[j][1] = t0.p[j][1];
            Rot r = euler(0, 0, (2 * draw(i, 0, 2) - 1) * pai);
            for (int j = 1; j <= 4; j++)
                for (int k = 2; k <= 4; k++) R[i].p[j][k] = rot_about(r, t0.p[j][k], R[i].p[j][1]);
        }
        for (int i = NA + 1; i <= N; i++) {
            double x, y, z;
            for (;;) {
                x = draw(i, 0, 0) * P.box[0] - P.box[0] / 2;
                y = draw(i, 0, 1) * P.box[1] - P.box[0] / 2;      // sic: cell_range_x, main.cpp:358
                z = draw(i, 0, 2) * P.box[2];
                bool clash = false;
                Pt c{x, y, z};
                for (int j = 1; j <= NA && !clash; j++)
                    for (int k = 1; k <= 4 && !clash; k++)
                        if (dist3(c, R[j].p[k][1]) <= P.rA + P.rB * 2 / sqrt(3) + P.rB) clash = true;
                for (int j = NA + 1; j <= i - 1 && !clash; j++)
                    if (dist3(c, R[j].p[1][1]) <= P.rB * 2 / sqrt(3) + P.rB * 2 / sqrt(3) + 2 * P.rB) clash = true;
                if (!clash) break;
            }
            Body t0; memset(&t0, 0, sizeof t0); ligand_template(t0, x, y, z);
            R[i].p[1][1] = t0.p[1][1];
            double th = (2 * draw(i, 0, 3) - 1) * pai;
            double ph = (2 * draw(i, 0, 4) - 1) * pai;
            double ps = (2 * draw(i, 0, 5) - 1) * pai;
            Rot r = euler(th, ph, ps);
            for (int j = 1; j <= 4; j++)
                for (int k = 1; k <= 2; k++)
                    if (j != 1 || k != 1) R[i].p[j][k] = rot_about(r, t0.p[j][k], R[i].p[1][1]);
        }
        std::fill(st.begin(), st.end(), 0); std::fill(nb.begin(), nb.end(), 0);
        bond_num = bond_rl = bond_cis = bond_mono = 0; max_complex = 0; step_done = 0;
    }

    // ---- S1: complexes rooted at ligands, breadth first (main.cpp:514-562) ----
    void find_complexes() {
        std::fill(visited.begin(), visited.end(), 0);
        std::fill(moved.begin(), moved.end(), 0);
        for (auto &r : rows) r.clear();
        std::deque<int> q;
        for (int i = NA + 1; i <= N; i++) {
            if (visited[i]) continue;
            visited[i] = 1; q.push_back(i);
            auto &row = rows[i - NA - 1];
            while (!q.empty()) {
                int m = q.front(); q.pop_front();
                row.push_back(m);
                int cand[3], nc = 0;
                if (m <= NA) {
                    if (NBR(nb, m, 2) > 0) cand[nc++] = NBR(nb, m, 2);
                    if (NBR(nb, m, 3) > 0) cand[nc++] = NBR(nb, m, 3);
                } else {
                    for (int s = 2; s <= 4; s++) if (NBR(nb, m, s) > 0) cand[nc++] = NBR(nb, m, s);
                }
                for (int c = 0; c < nc; c++)
                    if (!visited[cand[c]]) { visited[cand[c]] = 1; q.push_back(cand[c]); }
            }
        }
    }

    void revert(int id) { Rn[id] = R[id]; }

    // ---- S2a: free receptor (main.cpp:584-677) ----
    void move_free_receptor(int a) {
        double amp = 2 * sqrt(P.DA * P.dt / 6) * draw(a, 0, 0);
        double phai = draw(a, 0, 1) * 2 * P.pai;
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++) {
                T[a].p[j][k].x = R[a].p[j][k].x + amp * cos(phai);
                T[a].p[j][k].y = R[a].p[j][k].y + amp * sin(phai);
                T[a].p[j][k].z = R[a].p[j][k].z;
            }
        double PBx = P.box[0] * round(T[a].p[1][1].x / P.box[0]);
        double PBy = P.box[1] * round(T[a].p[1][1].y / P.box[1]);
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++) { T[a].p[j][k].x = T[a].p[j][k].x - PBx; T[a].p[j][k].y = T[a].p[j][k].y - PBy; }
        Rot r = euler(0, 0, (2 * draw(a, 0, 2) - 1) * sqrt(P.DrotA * P.dt));
        for (int j = 1; j <= 4; j++) {
            Rn[a].p[j][1] = T[a].p[j][1];
            for (int k = 2; k <= 4; k++) Rn[a].p[j][k] = rot_about(r, T[a].p[j][k], Rn[a].p[j][1]);
        }
        ev[7]++;
        int unit[1] = {a};
        bool hit = receptor_collides(a, unit, 1);
        accepted[a] = !hit;
        if (hit) { revert(a); ev[6]++; } else if (P.use_grid) grid_put(a);
    }

    // ---- S2b: ligand-free cis dimer, moved once by its first-visited receptor (main.cpp:682-865) ----
    void move_cis_dimer(int a, int a2) {
        double amp = 2 * sqrt(P.cis_D * P.dt / 6) * draw(a, 0, 0);
        double phai = draw(a, 0, 1) * 2 * P.pai;
        const int pr[2] = {a, a2};
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++)
                for (int m : pr) {
                    T[m].p[j][k].x = R[m].p[j][k].x + amp * cos(phai);
                    T[m].p[j][k].y = R[m].p[j][k].y + amp * sin(phai);
                    T[m].p[j][k].z = R[m].p[j][k].z;
                }
        double PBx = P.box[0] * round((T[a].p[1][1].x + T[a2].p[1][1].x) / 2 / P.box[0]);
        double PBy = P.box[1] * round((T[a].p[1][1].y + T[a2].p[1][1].y) / 2 / P.box[1]);
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++)
                for (int m : pr) { T[m].p[j][k].x = T[m].p[j][k].x - PBx; T[m].p[j][k].y = T[m].p[j][k].y - PBy; }
        Rot r = euler(0, 0, (2 * draw(a, 0, 2) - 1) * sqrt(P.cis_Drot * P.dt));
        // rotation centre from R_new, which still holds the PRE-translation coordinates (745-747)
        Pt c{0, 0, 0};
        for (int j = 1; j <= 4; j++) {
            c.x = c.x + Rn[a].p[j][1].x + Rn[a2].p[j][1].x;
            c.y = c.y + Rn[a].p[j][1].y + Rn[a2].p[j][1].y;
            c.z = c.z + Rn[a].p[j][1].z + Rn[a2].p[j][1].z;
        }
        c.x = c.x / (4 * 2); c.y = c.y / (4 * 2); c.z = c.z / (4 * 2);
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++)
                for (int m : pr) Rn[m].p[j][k] = rot_about(r, T[m].p[j][k], c);
        if (cis_misaligned(a, a2)) snap_cis(a2, a);   // "relax", 770-799
        ev[7]++;
        bool hit = false;
        if (receptor_collides(a, pr, 2)) hit = true;
        if (receptor_collides(a2, pr, 2)) hit = true;
        accepted[a] = accepted[a2] = !hit;
        if (hit) { revert(a); revert(a2); ev[6]++; }
        else if (P.use_grid) { grid_put(a); grid_put(a2); }
    }

    // ---- S2c: free ligand (main.cpp:905-969) ----
    void move_free_ligand(int b) {
        double amp = 2 * sqrt(P.DB * P.dt / 6) * draw(b, 0, 0);
        double theta = draw(b, 0, 1) * P.pai;
        double phai = draw(b, 0, 2) * 2 * P.pai;
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 2; k++) {
                T[b].p[j][k].x = R[b].p[j][k].x + amp * sin(theta) * cos(phai);
                T[b].p[j][k].y = R[b].p[j][k].y + amp * sin(theta) * sin(phai);
                T[b].p[j][k].z = R[b].p[j][k].z + amp * cos(theta);
            }
        double PBx = P.box[0] * round(T[b].p[1][1].x / P.box[0]);
        double PBy = P.box[1] * round(T[b].p[1][1].y / P.box[1]);
        double PBz = P.box[2] * round(T[b].p[1][1].z / P.box[2]);
        if (T[b].p[1][1].z > P.box[2] || T[b].p[1][1].z < 0)
            for (int j = 1; j <= 4; j++)
                for (int k = 1; k <= 2; k++) T[b].p[j][k].z = -T[b].p[j][k].z + 2 * PBz;
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 2; k++) { T[b].p[j][k].x = T[b].p[j][k].x - PBx; T[b].p[j][k].y = T[b].p[j][k].y - PBy; }
        double th = (2 * draw(b, 0, 3) - 1) * sqrt(P.DrotB * P.dt);
        double ph = (2 * draw(b, 0, 4) - 1) * sqrt(P.DrotB * P.dt);
        double ps = (2 * draw(b, 0, 5) - 1) * sqrt(P.DrotB * P.dt);
        Rot r = euler(th, ph, ps);
        Rn[b].p[1][1] = T[b].p[1][1];
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 2; k++) Rn[b].p[j][k] = rot_about(r, T[b].p[j][k], Rn[b].p[1][1]);
    }

    // ---- S2d: rigid move of a whole complex (main.cpp:974-1131) ----
    // returns the last receptor / last ligand seen in row order (the reference keeps them in the
    // globals protein_A_index / protein_B_index, which S2e then reads, main.cpp:1141-1152)
    void move_complex(int root, const std::vector<int> &row, int nA, int nBc, int &lastA, int &lastB) {
        double D = (nBc == 1) ? P.bond_D : 0.0;       // 984-985 (still draws)
        double amp = 2 * sqrt(D * P.dt / 6) * draw(root, 0, 0);
        double phai = draw(root, 0, 1) * 2 * P.pai;
        double PBx = 0, PBy = 0;
        for (int m : row) {
            int np = npts(m);
            for (int j = 1; j <= 4; j++)
                for (int k = 1; k <= np; k++) {
                    T[m].p[j][k].x = R[m].p[j][k].x + amp * cos(phai);
                    T[m].p[j][k].y = R[m].p[j][k].y + amp * sin(phai);
                    T[m].p[j][k].z = R[m].p[j][k].z;
                }
            PBx = PBx + T[m].p[1][1].x; PBy = PBy + T[m].p[1][1].y;
            if (m <= NA) lastA = m; else lastB = m;
        }
        PBx = P.box[0] * round(PBx / (nA + nBc) / P.box[0]);
        PBy = P.box[1] * round(PBy / (nA + nBc) / P.box[1]);
        Pt c{0, 0, 0};
        for (int m : row) {
            int np = npts(m);
            for (int j = 1; j <= 4; j++)
                for (int k = 1; k <= np; k++) { T[m].p[j][k].x = T[m].p[j][k].x - PBx; T[m].p[j][k].y = T[m].p[j][k].y - PBy; }
            for (int j = 1; j <= 4; j++) { c.x = c.x + T[m].p[j][1].x; c.y = c.y + T[m].p[j][1].y; c.z = c.z + T[m].p[j][1].z; }
        }
        c.x = c.x / (4 * nA + 4 * nBc); c.y = c.y / (4 * nA + 4 * nBc); c.z = c.z / (4 * nA + 4 * nBc);
        double Dr = (nBc == 1) ? P.bond_Drot : 0.0;   // 1082-1083
        Rot r = euler(0, 0, (2 * draw(root, 0, 2) - 1) * sqrt(Dr * P.dt));
        for (int m : row) {
            int np = npts(m);
            for (int j = 1; j <= 4; j++)
                for (int k = 1; k <= np; k++) Rn[m].p[j][k] = rot_about(r, T[m].p[j][k], c);
        }
    }

    // ---- S2e: single-ligand complex: lay the ligand down, align receptors, align cis partners (1138-1274) ----
    void align_single_ligand(int b, int lastA) {
        if (Rn[b].p[1][2].z != (Rn[b].p[1][1].z + P.rB)) {      // exact compare, SURVEY Q11
            ev[8]++;
            double zA = Rn[lastA].p[3][1].z;
            for (int j = 1; j <= 4; j++) for (int k = 1; k <= 2; k++) Rn[b].p[j][k].z = zA;
            Rn[b].p[1][2].z = zA + P.rB;
            double angle = atan2(Rn[b].p[2][1].x - Rn[b].p[1][1].x, Rn[b].p[2][1].y - Rn[b].p[1][1].y) + P.pai;
            Body g; ligand_ghost(g);
            seat_ligand(b, g, angle, Rn[b].p[1][1].x, Rn[b].p[1][1].y);
        }
        for (int j = 2; j <= 4; j++) {
            int a1 = NBR(nbn, b, j);
            if (a1 != 0 && rl_misaligned(b, j, a1)) snap_receptor_to_ligand(a1, b, j);
        }
        for (int j = 2; j <= 4; j++) {
            int a1 = NBR(nbn, b, j);
            if (a1 != 0 && NBR(nbn, a1, 3) != 0) {
                int a2 = NBR(nbn, a1, 3);
                if (cis_misaligned(a1, a2)) snap_cis(a2, a1);
            }
        }
    }

    // body executed at `lable4` (main.cpp:1438-1585): re-seat ligand b against receptor a1 on its site j,
    // then re-snap every receptor of b and their cis partners
    void reseat_ligand(int b, int j, int a1) {
        moved[b] = 1;
        double zA = Rn[a1].p[3][1].z;
        for (int k = 1; k <= 2; k++) for (int jj = 1; jj <= 4; jj++) Rn[b].p[jj][k].z = zA;
        Rn[b].p[1][2].z = zA + P.rB;
        Body g; ligand_ghost(g);
        double ax1 = g.p[j][1].x, ay1 = g.p[j][1].y;
        double ax2 = Rn[a1].p[3][1].x - Rn[a1].p[3][2].x, ay2 = Rn[a1].p[3][1].y - Rn[a1].p[3][2].y;
        double dot = ax1 * ax2 + ay1 * ay2, det = ax1 * ay2 - ay1 * ax2;
        double angle = atan2(-det, -dot) + P.pai;
        double cx = (P.bond_dist_cut / 2 + P.rB * 2 / sqrt(3) + P.rB) / P.rA * (Rn[a1].p[3][2].x - Rn[a1].p[3][1].x) + Rn[a1].p[3][2].x;
        double cy = (P.bond_dist_cut / 2 + P.rB * 2 / sqrt(3) + P.rB) / P.rA * (Rn[a1].p[3][2].y - Rn[a1].p[3][1].y) + Rn[a1].p[3][2].y;
        seat_ligand(b, g, angle, cx, cy);
        for (int m = 2; m <= 4; m++) {
            int am = NBR(nbn, b, m);
            if (NBR(nbn, am, 2) != 0) {            // row 0 of res_nei is all zero
                int bb = NBR(nbn, am, 2), n = NBR(nbn, am, 4);
                if (rl_misaligned(bb, n, am)) { moved[am] = 1; snap_receptor_to_ligand(am, bb, n); }
                if (NBR(nbn, am, 3) != 0) {
                    int a2 = NBR(nbn, am, 3);
                    if (cis_misaligned(am, a2)) { moved[a2] = 1; snap_cis(a2, am); }
                }
            }
        }
    }
    bool bridge_candidate(int b, int j) {   // main.cpp:1420-1423 / 1604-1607
        int a1 = NBR(nbn, b, j);
        return a1 != 0 && NBR(nbn, a1, 3) != 0 && NBR(nbn, NBR(nbn, a1, 3), 2) != 0 && moved[b] == 0;
    }

    // ---- S2f: complexes with >= 2 ligands (main.cpp:1284-1732) ----
    void align_multi_ligand(int root, std::vector<int> &row) {
        uint32_t cnt = 0;
        const int size = (int)row.size();
        // pass 0 (1284-1332)
        shuffle_row(row, root, cnt);
        for (int s = 0; s < size; s++) {
            int a1 = row[s];
            if (a1 <= NA && NBR(nbn, a1, 2) != 0) {
                int b = NBR(nbn, a1, 2), j = NBR(nbn, a1, 4);
                if (rl_misaligned(b, j, a1)) { moved[a1] = 1; snap_receptor_to_ligand(a1, b, j); }
            }
        }
        // pass 1 (1343-1406)
        shuffle_row(row, root, cnt);
        for (int s = 0; s < size; s++) {
            int a = row[s];
            if (a <= NA && NBR(nbn, a, 2) != 0 && NBR(nbn, a, 3) != 0 && NBR(nbn, NBR(nbn, a, 3), 2) != 0 && moved[a] == 0) {
                int a2 = NBR(nbn, a, 3);
                moved[a] = 1; moved[a2] = 1;
                if (cis_misaligned(a, a2)) snap_cis(a, a2);      // a rebuilt FROM a2 (1390-1400)
            }
        }
        // pass 2 (1411-1590) and pass 3 (1595-1635). Pass 3 jumps back INTO pass 2's innermost block
        // (`goto lable4`, 1628 -> 1438) keeping its own loop position, ligand, site and receptor.
        bool resume = false; int s = 0, j = 2, b = 0, a1 = 0;
        for (;;) {
            if (!resume) { shuffle_row(row, root, cnt); s = 0; }
            for (; s < size; s++) {
                if (row[s] <= NA) continue;
                if (!resume) { b = row[s]; j = 2; }
                for (; j <= 4; j++) {
                    bool run_body;
                    if (resume) { run_body = true; resume = false; }
                    else {
                        run_body = false;
                        if (bridge_candidate(b, j)) { a1 = NBR(nbn, b, j); run_body = rl_misaligned(b, j, a1); }
                    }
                    if (run_body) reseat_ligand(b, j, a1);
                }
            }
            // pass 3
            shuffle_row(row, root, cnt);
            for (s = 0; s < size && !resume; s++) {
                if (row[s] <= NA) continue;
                b = row[s];
                for (j = 2; j <= 4; j++)
                    if (bridge_candidate(b, j)) {
                        a1 = NBR(nbn, b, j);
                        if (rl_misaligned(b, j, a1)) { resume = true; break; }
                    }
                if (resume) break;
            }
            if (!resume) break;
            ev[9]++;
        }
        // pass 4 (1645-1687)
        for (int q = 0; q < size; q++) {
            int a = row[q];
            if (a <= NA && NBR(nbn, a, 2) != 0) {
                int bb = NBR(nbn, a, 2), jj = NBR(nbn, a, 4);
                if (rl_misaligned(bb, jj, a)) { moved[a] = 1; snap_receptor_to_ligand(a, bb, jj); }
            }
        }
        // pass 5 (1691-1732)
        for (int q = 0; q < size; q++) {
            int a = row[q];
            if (a <= NA && NBR(nbn, a, 2) != 0 && NBR(nbn, a, 3) != 0 && NBR(nbn, NBR(nbn, a, 3), 2) == 0) {
                int a2 = NBR(nbn, a, 3);
                if (cis_misaligned(a, a2)) snap_cis(a2, a);
            }
        }
    }

    // ---- ligand-indexed part of S2 (main.cpp:879-1862) ----
    void ligand_unit(int b) {
        auto &row = rows[b - NA - 1];
        const int size = (int)row.size();
        if (size == 0) return;                       // not a BFS root: nothing happens
        int nA = 0, nBc = 0;
        for (int m : row) { if (m > NA) nBc++; else nA++; }
        if (size > max_complex) max_complex = size;  // 896-898, running max never reset (Q15)
        if (size == 1) move_free_ligand(row[0]);
        else {
            tot_cluster_num++; tot_proteins_in_cluster += size;
            int lastA = 0, lastB = 0;
            move_complex(b, row, nA, nBc, lastA, lastB);
            if (nBc == 1) align_single_ligand(lastB, lastA);
            else align_multi_ligand(b, row);
        }
        // S2g (1759-1860)
        ev[7]++;
        bool hit = false;
        for (int m : row) {
            if (m <= NA) { if (receptor_collides(m, row.data(), size)) hit = true; }
            else if (ligand_collides(m, row.data(), size)) hit = true;
        }
        for (int m : row) accepted[m] = !hit;
        if (hit) { for (int m : row) revert(m); ev[6]++; }
        else if (P.use_grid) for (int m : row) grid_put(m);
    }

    // ---- S3 reactions (main.cpp:1876-2141) ----
    bool rl_geometry_ok(int i, int j, int k) const {
        double d = dist3(Rn[j].p[k][2], Rn[i].p[3][2]);
        if (!(d < P.bond_dist_cut)) return false;
        Pt p0{Rn[i].p[3][1].x - Rn[i].p[3][2].x, Rn[i].p[3][1].y - Rn[i].p[3][2].y, Rn[i].p[3][1].z - Rn[i].p[3][2].z};
        Pt p2{Rn[j].p[k][1].x - Rn[j].p[k][2].x, Rn[j].p[k][1].y - Rn[j].p[k][2].y, Rn[j].p[k][1].z - Rn[j].p[k][2].z};
        double th_ot = angle_deg(p0, p2);
        Pt q0{Rn[i].p[3][1].x - Rn[i].p[3][4].x, Rn[i].p[3][1].y - Rn[i].p[3][4].y, Rn[i].p[3][1].z - Rn[i].p[3][4].z};
        Pt q2{Rn[j].p[1][1].x - Rn[j].p[1][2].x, Rn[j].p[1][1].y - Rn[j].p[1][2].y, Rn[j].p[1][1].z - Rn[j].p[1][2].z};
        double th_pd = angle_deg(q0, q2);
        return (std::fabs(th_pd) < P.thetapd_cut) && (std::fabs(th_ot - 180) < P.thetaot_cut);
    }
    bool cis_geometry_ok(int i, int j) const {
        double d = dist3(Rn[j].p[3][3], Rn[i].p[3][3]);
        if (!(d < P.cis_dist_cut)) return false;
        Pt p0{Rn[i].p[3][1].x - Rn[i].p[3][3].x, Rn[i].p[3][1].y - Rn[i].p[3][3].y, Rn[i].p[3][1].z - Rn[i].p[3][3].z};
        Pt p2{Rn[j].p[3][1].x - Rn[j].p[3][3].x, Rn[j].p[3][1].y - Rn[j].p[3][3].y, Rn[j].p[3][1].z - Rn[j].p[3][3].z};
        return std::fabs(angle_deg(p0, p2) - 180) < P.cis_thetaot_cut;
    }
    template <class F> void each_partner(int i, bool ligands, double reach, F f) {
        if (!P.use_grid) {
            if (ligands) for (int j = NA + 1; j <= N; j++) f(j);
            else for (int j = 1; j <= NA; j++) f(j);
            return;
        }
        std::vector<int> c;
        grid_near(Rn[i].p[1][1].x, Rn[i].p[1][1].y, reach, [&](int id) { if ((id > NA) == ligands) c.push_back(id); });
        std::sort(c.begin(), c.end());               // reference loop order is ascending j
        for (int j : c) f(j);
    }
    void reactions() {
        double rs = P.rB * 2 / sqrt(3.0);
        double reach_on = P.rA + P.bond_dist_cut + rs + P.rB + 1e-3;
        double reach_cis = 2 * P.rA + P.cis_dist_cut + 1e-3;
        // (1) R-L association, 1877-1949
        for (int i = 1; i <= NA; i++) {
            if (S(stn, i, 2) != 0) continue;
            each_partner(i, true, reach_on, [&](int j) {
                for (int k = 2; k <= 4; k++) {
                    if (S(stn, i, 2) == 0 && S(stn, j, k) == 0 && rl_geometry_ok(i, j, k)) {
                        double prob = draw(i, 4 * j + k, kmco::SLOT_RL_ON);
                        if (prob < P.on * P.dt) {
                            S(stn, i, 2) = 1; S(stn, j, k) = 1;
                            NBR(nbn, j, k) = i; NBR(nbn, i, 2) = j; NBR(nbn, i, 4) = k;
                            bond_num_n++; bond_rl_n++; ev[0]++;
                            int a2 = NBR(nbn, i, 3);
                            if (a2 != 0 && S(stn, a2, 2) == 0) { bond_mono_n--; bond_cis_n++; }
                        }
                    }
                }
            });
        }
        // (2) cis association between two ligand-free receptors, 1952-2003
        for (int i = 1; i <= NA; i++) {
            if (S(stn, i, 3) != 0 || S(stn, i, 2) != 0) continue;
            each_partner(i, false, reach_cis, [&](int j) {
                if (i != j && S(stn, i, 3) == 0 && S(stn, j, 3) == 0 && S(stn, i, 2) == 0 && S(stn, j, 2) == 0 &&
                    cis_geometry_ok(i, j)) {
                    double prob = draw(i, j, kmco::SLOT_MONO_CIS_ON);
                    if (prob < P.mono_cis_on * P.dt) {
                        S(stn, i, 3) = 1; S(stn, j, 3) = 1; bond_num_n++; bond_mono_n++; ev[1]++;
                        NBR(nbn, j, 3) = i; NBR(nbn, i, 3) = j;
                    }
                }
            });
        }
        // (3) cis association with at least one ligand-bound receptor, 2007-2058
        for (int i = 1; i <= NA; i++) {
            if (S(stn, i, 3) != 0) continue;
            each_partner(i, false, reach_cis, [&](int j) {
                if (i != j && S(stn, i, 3) == 0 && S(stn, j, 3) == 0 && (S(stn, j, 2) == 1 || S(stn, i, 2) == 1) &&
                    cis_geometry_ok(i, j)) {
                    double prob = draw(i, j, kmco::SLOT_CIS_ON);
                    if (prob < P.cis_on * P.dt) {
                        S(stn, i, 3) = 1; S(stn, j, 3) = 1; bond_num_n++; bond_cis_n++; ev[2]++;
                        NBR(nbn, j, 3) = i; NBR(nbn, i, 3) = j;
                    }
                }
            });
        }
        // (4) R-L dissociation, 2063-2092
        for (int i = 1; i <= NA; i++) {
            if (S(stn, i, 2) != 1) continue;
            int b = NBR(nbn, i, 2), sb = NBR(nbn, i, 4);
            double prob = draw(i, 0, kmco::SLOT_RL_OFF);
            if (prob < P.off * P.dt) {
                S(stn, i, 2) = 0; S(stn, b, sb) = 0;
                NBR(nbn, i, 2) = 0; NBR(nbn, i, 4) = 0; NBR(nbn, b, sb) = 0;
                bond_num_n--; bond_rl_n--; ev[3]++;
                int a2 = NBR(nbn, i, 3);
                if (a2 != 0 && S(stn, a2, 2) == 0) { bond_mono_n++; bond_cis_n--; }
            }
        }
        // (5) cis dissociation, ligand-free pair, 2097-2117; drawn from BOTH ends
        for (int i = 1; i <= NA; i++) {
            if (S(stn, i, 3) != 1) continue;
            int a2 = NBR(nbn, i, 3);
            if (S(stn, i, 2) == 0 && S(stn, a2, 2) == 0) {
                double prob = draw(i, 0, kmco::SLOT_MONO_CIS_OFF);
                if (prob < P.mono_cis_off * P.dt) {
                    S(stn, i, 3) = 0; S(stn, a2, 3) = 0; NBR(nbn, i, 3) = 0; NBR(nbn, a2, 3) = 0;
                    bond_num_n--; bond_mono_n--; ev[4]++;
                }
            }
        }
        // (6) cis dissociation inside complexes, 2120-2141
        for (int i = 1; i <= NA; i++) {
            if (S(stn, i, 3) != 1) continue;
            int a2 = NBR(nbn, i, 3);
            if (S(stn, i, 2) == 1 || S(stn, a2, 2) == 1) {
                double prob = draw(i, 0, kmco::SLOT_CIS_OFF);
                if (prob < P.cis_off * P.dt) {
                    S(stn, i, 3) = 0; S(stn, a2, 3) = 0; NBR(nbn, i, 3) = 0; NBR(nbn, a2, 3) = 0;
                    bond_num_n--; bond_cis_n--; ev[5]++;
                }
            }
        }
    }

    // ---- one time step, main.cpp:461-2202 ----
    void step() {
        cur_step = step_done + 1;
        // S0 (464-502)
        Rn = R; stn = st; nbn = nb;
        bond_num_n = bond_num; bond_rl_n = bond_rl; bond_cis_n = bond_cis; bond_mono_n = bond_mono;
        tot_cluster_num = 0; tot_proteins_in_cluster = 0; cluster_size = 0.0;
        find_complexes();                                         // S1
        if (P.use_grid) { if (bins.empty()) grid_setup(); grid_fill(); }
        std::fill(accepted.begin(), accepted.end(), 1);
        // S2 (577-1872), Gauss-Seidel in index order (or, order_mode 1, colour by colour of the head's cell)
        std::vector<int> seq(N);
        for (int m = 1; m <= N; m++) seq[m - 1] = m;
        if (P.order_mode == 1) {
            std::vector<unsigned> key(N + 1);
            for (int m = 1; m <= N; m++) {
                int cx = (int)floor((R[m].p[1][1].x - P.order_x0) * P.order_inv_edge), cy = (int)floor((R[m].p[1][1].y - P.order_y0) * P.order_inv_edge);
                key[m] = ((unsigned)((cx & 1) | ((cy & 1) << 1)) << 30) | (unsigned)m;
            }
            std::sort(seq.begin(), seq.end(), [&](int a, int b) { return key[a] < key[b]; });
        }
        for (int m : seq) {
            if (m <= NA) {
                if (S(stn, m, 2) == 0 && S(stn, m, 3) == 0) move_free_receptor(m);
                int p = NBR(nb, m, 3);
                if (visited[m] == 0 && NBR(nb, m, 2) == 0 && m == NBR(nb, p, 3) && NBR(nb, p, 2) == 0 && (P.order_mode == 0 || m < p)) {
                    visited[p] = 1;
                    move_cis_dimer(m, p);
                }
            } else ligand_unit(m);
        }
        reactions();                                              // S3
        // S4 (2164-2202)
        R = Rn; st = stn; nb = nbn;
        bond_num = bond_num_n; bond_rl = bond_rl_n; bond_cis = bond_cis_n; bond_mono = bond_mono_n;
        if (tot_cluster_num != 0) cluster_size = (double)tot_proteins_in_cluster / tot_cluster_num;
        step_done = cur_step;
    }
};

}  // namespace

extern "C" {

void kmco_default_params(kmco_params *p) {
    memset(p, 0, sizeof *p);
    p->box[0] = 5773; p->box[1] = 5773; p->box[2] = 1000; p->dt = 10; p->pai = 3.1415926;
    p->rA = 20; p->DA = 1; p->DrotA = 0.0174; p->rB = 30; p->DB = 7.2614; p->DrotB = 0.0061209;
    p->mono_cis_on = 0.000047; p->mono_cis_off = 0.000000000000112;
    p->cis_D = 0.5; p->cis_Drot = 0.005; p->cis_on = 0.00096; p->cis_off = 0.000000000000112;
    p->bond_D = 0.5; p->bond_Drot = 0.005; p->on = 0.04; p->off = 0.000000000000348;
    p->bond_dist_cut = 18; p->thetapd_cut = 45; p->thetaot_cut = 90; p->cis_thetaot_cut = 10; p->cis_dist_cut = 15;
    p->n_receptor = 150; p->n_ligand = 50; p->stream_mode = 0; p->use_grid = 0; p->seed = 1;
    p->rand2_state = 88172645463325252ULL; p->rand_state = 0x9E3779B97F4A7C15ULL;
    p->order_mode = 0; p->order_x0 = p->order_y0 = 0; p->order_inv_edge = 1.0 / 256;
}
void *kmco_create(const kmco_params *p) { return new Oracle(*p); }
void kmco_destroy(void *h) { delete (Oracle *)h; }
void kmco_init_reference(void *h) { ((Oracle *)h)->init_reference(); }

void kmco_set_state(void *h, const double *Rx, const double *Ry, const double *Rz, const int32_t *status,
                    const int32_t *res_nei, int64_t step_done, int32_t max_complex) {
    Oracle &o = *(Oracle *)h;
    for (int i = 0; i <= o.N; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 5; k++) {
                size_t q = (size_t)i * 25 + j * 5 + k;
                o.R[i].p[j][k] = {Rx[q], Ry[q], Rz[q]};
            }
    for (size_t q = 0; q < o.st.size(); q++) o.st[q] = status[q];
    for (size_t q = 0; q < o.nb.size(); q++) o.nb[q] = res_nei[q];
    // counters are functions of the bond table (every update site keeps them consistent)
    o.bond_rl = o.bond_cis = o.bond_mono = 0;
    for (int i = 1; i <= o.NA; i++) {
        if (o.NBR(o.nb, i, 2) != 0) o.bond_rl++;
        int p = o.NBR(o.nb, i, 3);
        if (p > i) { if (o.NBR(o.nb, i, 2) != 0 || o.NBR(o.nb, p, 2) != 0) o.bond_cis++; else o.bond_mono++; }
    }
    o.bond_num = o.bond_rl + o.bond_cis + o.bond_mono;
    o.step_done = step_done; o.max_complex = max_complex;
}
void kmco_get_state(void *h, double *Rx, double *Ry, double *Rz, int32_t *status, int32_t *res_nei) {
    Oracle &o = *(Oracle *)h;
    for (int i = 0; i <= o.N; i++)
        for (int j = 0; j < 5; j++)
            for (int k = 0; k < 5; k++) {
                size_t q = (size_t)i * 25 + j * 5 + k;
                Rx[q] = o.R[i].p[j][k].x; Ry[q] = o.R[i].p[j][k].y; Rz[q] = o.R[i].p[j][k].z;
            }
    for (size_t q = 0; q < o.st.size(); q++) status[q] = o.st[q];
    for (size_t q = 0; q < o.nb.size(); q++) res_nei[q] = o.nb[q];
}
void kmco_step(void *h, int64_t nsteps) { Oracle &o = *(Oracle *)h; for (int64_t s = 0; s < nsteps; s++) o.step(); }
double kmco_get_counts(void *h, int32_t *c, int64_t *step_done, uint64_t *n_rand2, uint64_t *n_rand) {
    Oracle &o = *(Oracle *)h;
    c[0] = o.bond_num; c[1] = o.bond_rl; c[2] = o.bond_cis; c[3] = o.bond_mono; c[4] = o.max_complex;
    c[5] = o.tot_cluster_num; c[6] = o.tot_proteins_in_cluster; c[7] = o.N;
    if (step_done) *step_done = o.step_done;
    if (n_rand2) *n_rand2 = o.n_rand2;
    if (n_rand) *n_rand = o.n_rand;
    return o.cluster_size;
}
int64_t kmco_get_results(void *h, int32_t *row_len, int32_t *members, int64_t cap) {
    Oracle &o = *(Oracle *)h; int64_t tot = 0;
    for (int l = 0; l < o.NB; l++) {
        row_len[l] = (int32_t)o.rows[l].size();
        for (int m : o.rows[l]) { if (tot < cap) members[tot] = m; tot++; }
    }
    return tot;
}
void kmco_get_accept(void *h, int32_t *accepted) {
    Oracle &o = *(Oracle *)h; for (int i = 0; i <= o.N; i++) accepted[i] = o.accepted[i];
}
void kmco_get_events(void *h, int64_t *ev) { Oracle &o = *(Oracle *)h; for (int i = 0; i < 16; i++) ev[i] = o.ev[i]; }
}
