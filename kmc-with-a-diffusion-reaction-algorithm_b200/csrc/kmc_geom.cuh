// csrc/kmc_geom.cuh -- device-side rigid-body geometry of the KMC sweep.
//
// Bit-level contract: every expression below reproduces the operand order and association of the
// reference statement it cites (main.cpp:NNN), with IEEE round-to-nearest adds/muls that can never be
// contracted into FMAs (explicit __dadd_rn/__dmul_rn), so that accept/reject, AreSame and threshold
// decisions are identical to the reference's x86-64 evaluation. Only libm calls (sin, cos, atan2,
// acos) may differ in the last ulp from glibc; DESIGN.md "Numerics" quantifies that.
//
// Compact poses (DESIGN.md "Data layout"):
//   receptor: centre c=(x,y) [= site 1 = site 4 in xy, all four beads], site 2, site 3   (6 doubles)
//             bead j sits at z = (2j-2)*rA, site 4 at z + rA: never stored (exact templates, main.cpp:301-315)
//   ligand  : 8 points x,y,z: 0 centre(1,1) 1..3 beads (2..4,1) 4 normal marker (1,2) 5..7 sites (2..4,2)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace kmc {

#define KD __device__ __forceinline__

KD double add(double a, double b) { return __dadd_rn(a, b); }
KD double sub(double a, double b) { return __dsub_rn(a, b); }
KD double mul(double a, double b) { return __dmul_rn(a, b); }
KD double dvd(double a, double b) { return __ddiv_rn(a, b); }
KD double sq(double a) { return __dmul_rn(a, a); }

// parameters + derived constants, filled on the host with the reference's own expressions
struct Consts {
    double Lx, Ly, Lz, dt, pai, rA, rB;
    double ampA, ampB, ampCis, ampBond;       // 2*sqrt(D*dt/6)          main.cpp:585, 909, 693, 990
    double rotA, rotB, rotCis, rotBond;       // sqrt(Drot*dt)           main.cpp:611, 942, 726, 1089
    double pOn, pMonoCisOn, pCisOn, pOff, pMonoCisOff, pCisOff;   // rate*dt     1918, 1984, 2038, 2069, 2104, 2127
    double bondCut, thetaPdCut, thetaOtCut, cisThetaCut, cisCut;
    double ovAA, ovAB, ovBB;                  // rA+rA, rA+rB, rB+rB     main.cpp:646, 659, 1806
    double ovAA2, ovAB2, ovBB2;               // exact squared thresholds: sqrt_rn(d2) < ov  <=>  d2 < ov2 (see sq_threshold)
    double rlD1, rlD2;                        // bondCut/2+rA+rB, bondCut/2      main.cpp:1215
    double cisD1, cisD2;                      // cisCut/2+rA+rA, cisCut/2        main.cpp:780-781
    double wRL1[2], wRL2[2], wCis1[2], wCis2[2];   // AreSame(sqrt(q), D) <=> w[0] <= q <= w[1]: the exact windows of the squared distance for the four alignment lengths above (same_window, host); w[0] > w[1] = not available, take the square root
    double fRL1, fRL3, fRL2;                  // (bondCut/2+rA)/rB, (bondCut/2+2rA)/rB, (bondCut/2)/rB   1217-1227
    double fC1, fC3, fC2;                     // (cisCut/2+rA)/rA, (cisCut/2)/rA, (cisCut/2+2rA)/rA      787-797
    double fSeat;                             // (bondCut/2 + rB*2/sqrt(3) + rB)/rA                      1491
    double ghost[8][2];                       // ligand template around the origin, main.cpp:1157-1179
    double conv;                              // 180/3.14159             main.cpp:2353
    double reachRR, reachRL, reachLL, reachOn, reachCis;   // centre-centre search bounds (with margin)
    double skin;                              // far-mover threshold
    double drift;                             // list reuse: a molecule further than this from its grid entry is 'displaced' (0 = grid rebuilt every step)
    int phase;                                // 0 = this step rebuilds the neighbour grid and the pair list, 1 = it reuses them, 2 = fused small-system step (csrc/kmc_small.cu)
    int smallRep;                             // phase 2: the replica this CTA's view stands for (spares the hot path every division by NA / NB)
    float cutMargin;                          // slack of the fp32 distance cut of k_cells_cut (covers the rounding of coordinates to fp32)
    double gx0, gy0, cellInv; int ncx, ncy;   // neighbour grid
    double keyX0, keyY0, keyInv;              // the cells whose 2x2 colouring orders the sweep in production mode: the neighbour grid as it was when the handle was set up (the neighbour grid itself may be re-laid later: list-reuse back-off)
    int tileEdge;                             // k_resolve_tiles: cells per tile edge
    int NA, NB, R, mode;                      // per-replica sizes, replicas
    int NAt, NBt, NT;                         // totals (capacities of the receptor / ligand blocks; live counts are device scalars)
    uint64_t seed;
    int strips, stripRank;                    // strip decomposition along x (1 = whole membrane on this GPU)
    double stripXc, stripHalf;                // centre of this rank's strip and Lx/2 (+inf without strips): cells are hashed in the periodic frame around the centre
};

struct Rec { double cx, cy, s2x, s2y, s3x, s3y; };
struct Lig { double p[8][3]; };

// L*round(x/L) (main.cpp:597-598 and its siblings). Away from the box edge the quotient rounds to (+-)0 and the product is
// (+-)0, which the following subtraction ignores: the division is skipped there (|x| < 0.4999 L: the correctly rounded quotient is then below 0.5 and rounds to 0; x != 0 keeps even the sign of a
// zero coordinate as the reference computes it).
KD double wrap_offset(double x, double L) {
    if (x != 0.0 && fabs(x) < 0.4999 * L) return 0.0;
    return mul(L, round(dvd(x, L)));
}
// receptor bead z (main.cpp:301): (j*2-2)*rA, j = 1..4
KD double rec_bead_z(const Consts &K, int j) { return mul((double)(j * 2 - 2), K.rA); }

// ---- z-axis rotation about (ox,oy): main.cpp:631-632 with theta = phi = 0 (t00=cos, t01=-sin, t10=sin, t11=cos;
// the t02/t12 terms are +-0 and drop out exactly) ----
KD void rotz(double c, double s, double x, double y, double ox, double oy, double &nx, double &ny) {
    double dx = sub(x, ox), dy = sub(y, oy);
    nx = add(add(mul(c, dx), mul(-s, dy)), ox);
    ny = add(add(mul(s, dx), mul(c, dy)), oy);
}

// sin and cos of one angle (inline at every call site: an out-of-line copy shared by all sites was measured SLOWER in the fused
// small-system kernel, 12.0 -> 15.2 us per step -- the independent sincos / Philox chains of a proposal no longer interleave)
#define KMC_SINCOS(K_, x, s, c) sincos((x), (s), (c))

// full Euler matrix, main.cpp:946-956
struct Rot3 { double t[3][3]; };
KD Rot3 euler_sc(double st, double ct, double sp, double cp, double ss, double cs) {
    Rot3 r;
    r.t[0][0] = sub(mul(cs, cp), mul(mul(ct, sp), ss));
    r.t[0][1] = sub(mul(-ss, cp), mul(mul(ct, sp), cs));
    r.t[0][2] = mul(st, sp);
    r.t[1][0] = add(mul(cs, sp), mul(mul(ct, cp), ss));
    r.t[1][1] = add(mul(-ss, sp), mul(mul(ct, cp), cs));
    r.t[1][2] = mul(-st, cp);
    r.t[2][0] = mul(ss, st);
    r.t[2][1] = mul(cs, st);
    r.t[2][2] = ct;
    return r;
}
KD Rot3 euler(const Consts &K, double theta, double phi, double psai) {
    double ct, st, cp, sp, cs, ss;
    KMC_SINCOS(K, theta, &st, &ct); KMC_SINCOS(K, phi, &sp, &cp); KMC_SINCOS(K, psai, &ss, &cs);
    return euler_sc(st, ct, sp, cp, ss, cs);
}
KD void rot3_about(const Rot3 &r, const double s[3], const double c[3], double q[3]) {
    double dx = sub(s[0], c[0]), dy = sub(s[1], c[1]), dz = sub(s[2], c[2]);
    q[0] = add(add(add(mul(r.t[0][0], dx), mul(r.t[0][1], dy)), mul(r.t[0][2], dz)), c[0]);
    q[1] = add(add(add(mul(r.t[1][0], dx), mul(r.t[1][1], dy)), mul(r.t[1][2], dz)), c[1]);
    q[2] = add(add(add(mul(r.t[2][0], dx), mul(r.t[2][1], dy)), mul(r.t[2][2], dz)), c[2]);
}

KD double dist2d(double ax, double ay, double bx, double by) {
    return sqrt(add(sq(sub(ax, bx)), sq(sub(ay, by))));
}
KD double dist3d(double ax, double ay, double az, double bx, double by, double bz) {
    return sqrt(add(add(sq(sub(ax, bx)), sq(sub(ay, by))), sq(sub(az, bz))));
}
KD bool are_same(double a, double b) { return fabs(sub(a, b)) < 1.0E-8; }   // main.cpp:2368-2371

// ---- overlap predicates (main.cpp:640-664, 1768-1826) ----
KD bool hit_rec_rec(const Consts &K, double ax, double ay, double bx, double by) {
    // bead-1 centres; both z are 0 so the z term adds an exact +0 (main.cpp:642-646)
    return add(sq(sub(bx, ax)), sq(sub(by, ay))) < K.ovAA2;      // == sqrt(...) < ovAA, without the sqrt
}
KD bool hit_rec_lig(const Consts &K, double ax, double ay, const Lig &l) {
    for (int j = 1; j <= 3; j++) {
        double d2 = add(sq(sub(l.p[j][0], ax)), sq(sub(l.p[j][1], ay)));
        for (int k = 1; k <= 4; k++) {
            if (add(d2, sq(sub(l.p[j][2], rec_bead_z(K, k)))) < K.ovAB2) return true;
        }
    }
    return false;
}
KD bool hit_lig_lig(const Consts &K, const Lig &a, const Lig &b) {
    for (int j = 1; j <= 3; j++)
        for (int k = 1; k <= 3; k++)
            if (add(add(sq(sub(a.p[j][0], b.p[k][0])), sq(sub(a.p[j][1], b.p[k][1]))), sq(sub(a.p[j][2], b.p[k][2]))) < K.ovBB2) return true;
    return false;
}

// ---- alignment predicates and snaps ----
// AreSame(dist2d(a, b), D) (main.cpp:2368-2371 on the distances of 1205-1215 / 1245-1255) without the square root: sqrt is
// correctly rounded and monotone and |s - D| < 1e-8 holds on an interval of s, so the test is a window on the squared distance
// q, computed exactly on the host (kmc_engine.cu: same_window). The alignment code evaluates it dozens of times per complex.
KD bool same_dist(const double w[2], double D, double ax, double ay, double bx, double by) {
    const double q = add(sq(sub(ax, bx)), sq(sub(ay, by)));
    if (w[0] > w[1]) return are_same(sqrt(q), D);
    return q >= w[0] && q <= w[1];
}
// main.cpp:1205-1215: ligand site s (reference j = s+2) given by its site point (sx, sy) and its bead (bx, by)
KD bool rl_misaligned(const Consts &K, double sx, double sy, double bx, double by, const Rec &a) {
    const bool same2 = same_dist(K.wRL2, K.rlD2, sx, sy, a.s2x, a.s2y);
    const bool same1 = same_dist(K.wRL1, K.rlD1, bx, by, a.cx, a.cy);
    return !same1 || !same2;
}
KD bool rl_misaligned(const Consts &K, const Lig &b, int s, const Rec &a) { return rl_misaligned(K, b.p[5 + s][0], b.p[5 + s][1], b.p[1 + s][0], b.p[1 + s][1], a); }
// main.cpp:1245-1255
KD bool cis_misaligned(const Consts &K, const Rec &a1, const Rec &a2) {
    const bool same2 = same_dist(K.wCis2, K.cisD2, a1.s3x, a1.s3y, a2.s3x, a2.s3y);
    const bool same1 = same_dist(K.wCis1, K.cisD1, a1.cx, a1.cy, a2.cx, a2.cy);
    return !same1 || !same2;
}
// main.cpp:1216-1228: receptor rebuilt on the bead->site axis of a ligand site (site point (sx, sy), bead (bx, by))
KD void snap_rec_to_lig(const Consts &K, Rec &a, double sx, double sy, double bx, double by) {
    double ux = sub(sx, bx), uy = sub(sy, by);
    a.cx = add(mul(K.fRL1, ux), sx);  a.cy = add(mul(K.fRL1, uy), sy);
    a.s3x = add(mul(K.fRL3, ux), sx); a.s3y = add(mul(K.fRL3, uy), sy);
    a.s2x = add(mul(K.fRL2, ux), sx); a.s2y = add(mul(K.fRL2, uy), sy);
}
KD void snap_rec_to_lig(const Consts &K, Rec &a, const Lig &b, int s) { snap_rec_to_lig(K, a, b.p[5 + s][0], b.p[5 + s][1], b.p[1 + s][0], b.p[1 + s][1]); }
// main.cpp:786-798 / 1256-1268: dst rebuilt from the centre->site-3 axis of src
KD void snap_cis(const Consts &K, Rec &dst, const Rec &src) {
    double ux = sub(src.s3x, src.cx), uy = sub(src.s3y, src.cy);
    dst.cx = add(mul(K.fC1, ux), src.s3x);  dst.cy = add(mul(K.fC1, uy), src.s3y);
    dst.s3x = add(mul(K.fC3, ux), src.s3x); dst.s3y = add(mul(K.fC3, uy), src.s3y);
    dst.s2x = add(mul(K.fC2, ux), src.s3x); dst.s2y = add(mul(K.fC2, uy), src.s3y);
}
// main.cpp:1184-1189 / 1496-1501: xy of all ligand points = Rz(angle)*ghost + (cx,cy)
KD int lig_point_of(int j, int k) { return k == 1 ? (j == 1 ? 0 : j - 1) : (j == 1 ? 4 : j + 3); }
KD void seat_ligand(const Consts &K, Lig &b, double angle, double cx, double cy) {
    double sa, ca; KMC_SINCOS(K, angle, &sa, &ca);
    for (int q = 0; q < 8; q++) {
        double gx = K.ghost[q][0], gy = K.ghost[q][1];
        b.p[q][0] = add(sub(mul(gx, ca), mul(gy, sa)), cx);
        b.p[q][1] = add(add(mul(gx, sa), mul(gy, ca)), cy);
    }
}

// main.cpp:2329-2366 with point[1] = 0: angle (degrees) between -p0 and p2
KD double angle_deg(const Consts &K, double p0x, double p0y, double p0z, double p2x, double p2y, double p2z) {
    double lx0 = sub(0.0, p0x), ly0 = sub(0.0, p0y), lz0 = sub(0.0, p0z);
    double lr0 = sqrt(add(add(sq(lx0), sq(ly0)), sq(lz0)));
    double lx1 = sub(p2x, 0.0), ly1 = sub(p2y, 0.0), lz1 = sub(p2z, 0.0);
    double lr1 = sqrt(add(add(sq(lx1), sq(ly1)), sq(lz1)));
    double doth1 = -add(add(mul(lx1, lx0), mul(ly0, ly1)), mul(lz0, lz1));
    double doth2 = dvd(doth1, mul(lr1, lr0));
    if (doth2 > 1) doth2 = 1;
    if (doth2 < -1) doth2 = -1;
    return mul(acos(doth2), K.conv);
}

// main.cpp:1882-1915: receptor i (bead 3: z = 4*rA... = rec_bead_z(3)) against site s of ligand b
KD bool rl_geometry_ok(const Consts &K, const Rec &a, const Lig &b, int s) {
    double z3 = rec_bead_z(K, 3);
    double d = dist3d(b.p[5 + s][0], b.p[5 + s][1], b.p[5 + s][2], a.s2x, a.s2y, z3);
    if (!(d < K.bondCut)) return false;
    double th_ot = angle_deg(K, sub(a.cx, a.s2x), sub(a.cy, a.s2y), sub(z3, z3),
                             sub(b.p[1 + s][0], b.p[5 + s][0]), sub(b.p[1 + s][1], b.p[5 + s][1]), sub(b.p[1 + s][2], b.p[5 + s][2]));
    // site 4 of bead 3 = centre xy, z + rA  (main.cpp:313-315)
    double z34 = mul((double)(3 * 2 - 1), K.rA);
    double th_pd = angle_deg(K, sub(a.cx, a.cx), sub(a.cy, a.cy), sub(z3, z34),
                             sub(b.p[0][0], b.p[4][0]), sub(b.p[0][1], b.p[4][1]), sub(b.p[0][2], b.p[4][2]));
    return (fabs(th_pd) < K.thetaPdCut) && (fabs(sub(th_ot, 180.0)) < K.thetaOtCut);
}
// main.cpp:1960-1981
KD bool cis_geometry_ok(const Consts &K, const Rec &a, const Rec &b) {
    double d = dist2d(b.s3x, b.s3y, a.s3x, a.s3y);     // both sites at z of bead 3: z term is an exact +0
    if (!(d < K.cisCut)) return false;
    double th = angle_deg(K, sub(a.cx, a.s3x), sub(a.cy, a.s3y), 0.0, sub(b.cx, b.s3x), sub(b.cy, b.s3y), 0.0);
    return fabs(sub(th, 180.0)) < K.cisThetaCut;
}

}  // namespace kmc
