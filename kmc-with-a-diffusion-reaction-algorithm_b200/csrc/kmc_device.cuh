// csrc/kmc_device.cuh -- device-resident state of one kmc_handle (SoA in HBM) and access helpers.
#pragma once
#include "kmc_geom.cuh"
#include "kmc_philox.cuh"

namespace kmc {

enum Scalar {
    S_MEMBER_CURSOR = 0,  // next free slot in members[]
    S_NCX,                // number of small complexes (1 < size <= CX_SMALL) at the front of cxRoots[] (one thread each)
    S_NFAR,               // far movers this step
    S_NPEND,              // pending findings of this step (pendList)
    S_EPOCH,              // steps taken by this handle (never reset): stamps the per-cell chains of special entries
    S_NSPEC,              // special entries of this step (specList): far movers and displaced molecules of a list-reuse step
    S_NPAIR,              // pre-selected reaction pairs of this step
    S_NCAND_RL, S_NCAND_CIS,
    S_TOPO_DIRTY,         // bond table changed: complexes must be rebuilt before the next sweep
    S_OVERFLOW,           // a device buffer overflowed (bitmask)
    S_NSURV,              // pairs in surv[] this step
    S_NREJ,               // rejected units of this step (rejList)
    S_NA_LIVE, S_NB_LIVE, // molecules actually present in the receptor / ligand blocks (<= NAt / NBt; strips change them)
    S_NSPEC_MAX,          // largest S_NSPEC of any step so far (the host decides the list-reuse back-off from it)
    S_NCX_BIG,            // complexes with more than CX_SMALL members: listed from the END of the first half of cxRoots[] (one warp each)
    S_NCX_MULTI,          // small complexes with several ligands: listed in the second half of cxRoots[] (S_NCX counts the single-ligand ones)
    S_NREACT,             // entries of reactList this step
    S_REC_TICKET,         // k_propose_rec: tiles handed out beyond the first one per CTA (dynamic schedule)
    S_NTOUCH,             // endpoints of the bonds formed / broken in the last step's S3 (touchList): their complexes are updated incrementally
    S_COUNT = 24
};
enum UnitState : unsigned char { U_UNKNOWN = 0, U_ACCEPT = 1, U_REJECT = 2 };
enum Event { EV_RL_ON = 0, EV_MONO_ON, EV_CIS_ON, EV_RL_OFF, EV_MONO_OFF, EV_CIS_OFF, EV_REVERTED, EV_TRIED, EV_FAR,
             EV_PASSES, EV_REBUILDS, EV_LAUNCHES, EV_COUNT = 16 };

// Molecule numbering on the device: receptors gid = [0, NAt), ligands gid = NAt + h. Replica r owns
// receptors [r*NA, (r+1)*NA) and ligands [r*NB, (r+1)*NB). Reference id (1-based, main.cpp) of receptor a
// is a%NA + 1, of ligand h it is NA + h%NB + 1.
struct Dev {
    // poses: cur = committed state of the previous step (R_*), nxt = proposals / new state (R_*_new)
    double2 *recC, *recS2, *recS3, *recCn, *recS2n, *recS3n;
    double *lig, *lign;                    // [NBt][24]
    // bond table (res_nei / protein_status, main.cpp:115-118): -1 = free
    int *recLig, *recSite, *recCis;        // ligand index, ligand site 0..2, cis partner
    int *ligRec;                           // [NBt][3] receptor on site s
    // complexes (S1, main.cpp:514-562)
    int *ufParent;                         // union-find over uid: ligands [0,NBt), receptors NBt + a
    int *unitOf;                           // [NT] gid of the head of the moving unit this molecule belongs to
    int *ukey;                             // [NT] sweep-order key of that unit: (colour << 30) | head; aliases unitOf in replay mode
    int *cxSize, *cxOff, *cxRoots;         // per root ligand: members, offset into members[]; list of roots with size>1
    int *members, *rowWork;                // member gids in BFS order / working copy permuted by the shuffles
    int *bfsMark;
    int *cxStamp;                          // [NT] incremental update: visited stamp (step epoch)
    int *rootSlot;                         // [NBt] where a root ligand sits in the work lists of cxRoots: (list << 28) | position, -1 = not listed
    int *touchList;                        // [TOUCH_CAP] molecules whose bonds changed in the last S3
    int *bfsQueue;                         // [NT] scratch of the (single-threaded) incremental update
    int *rowPos;                           // [NT] position of a complex member in its breadth-first member list
    unsigned char *movedFlag;
    double *nrec;                          // [NT][6] neighbour record per molecule: centre old xy, new xy, {old z (fp32), unit key, flags, new z (fp32)}
    // neighbour grid
    int *cellCount, *cellStart, *scanTmp;  // [ncell+1]
    int *sorted;                           // [2*NT] entries gid | ghost bit
    float2 *scen;                          // [2*NT] fp32 centre each entry stands for (old centre; proposed centre for a ghost), cell-sorted like `sorted`
    int *scell;                            // [2*NT] cell of each entry
    int2 *surv; int survCap;               // unordered entry pairs that passed the distance cut of k_cells_cut
    int *reactList;                        // [2*survCap] list pairs that may react in S3 this step: (list index << 1) | direction, appended per CTA by k_pairs_eval
    int *molSlot;                          // [NT]
    int4 *farList;                         // [NT] (gid, cell, slot, -)
    // list reuse (sparse path): the grid and the pair list of a build step serve the following steps as well
    float2 *bcen;                          // [NT] centre of the molecule's grid entry (fp32, what k_cells_cut measured from)
    int *specList;                         // [2*NT] entries (gid | ghost bit) that the stale grid/list do not cover this step
    unsigned long long *specNext;          // [2*NT] next special entry of the same cell: (stamp << 32) | (index + 1)
    unsigned long long *cellHead;          // [ncell] head of the per-cell chain of special entries, stamped with the step
    // reaction candidates (successful draws only)
    unsigned long long *candRL, *candCis;
    int candCap;
    unsigned long long *pairs; int pairCap;   // (receptor, neighbour) pairs that may react this step
    unsigned long long *pairsFast; int pairFastCap;   // fused small-system step: the first pairFastCap entries of that list live in shared memory (else null / 0)
    int *unitRes;                             // [NT] per unit head: 0 accepted, bit0 rejected (definite overlap), 2 = waits on pending findings
    int *pendCnt;                             // [NT] per unit head: pending findings not yet settled
    int *rejList;                             // [NT] heads of the units rejected this step (each once): their members are copied back
    int *rejPartner;                          // [NT] for a receptor-headed unit: its cis partner at rejection time (-1 none), i.e. the other member
    int2 *pendList; int pendCap;              // (unit head, earlier unit | bit30: overlap is with its NEW pose)
    unsigned long long *step64;               // [1] mc_time_step of the step being computed
    unsigned *refA, *refB;                    // reference (global, 1-based) ids of local receptors / ligands; null = a%NA+1, NA+h%NB+1
    int *scal;
    int *maxComplex;                       // [R]
    unsigned long long *events;
    int ncell;
    int nAcap;                             // = Consts::NAt (receptor gids are below it): for helpers that only get the Dev
    // fused small-system step (k_small_step, phase 2): the search records of the replica, in the shared memory of its CTA
    struct SmallSearch *small;
};

// Search records of one replica in the fused small-system step (csrc/kmc_small.cu), indexed by the molecule's number inside the
// replica (receptors first). Written by mark_far (phase 2) for every molecule once per step, read by the CTA's pair search.
#define SMALL_MAXN 256        // molecules per replica the records hold
#define SMALL_SPEC 32
struct SmallSearch {
    float4 cen[SMALL_MAXN];            // old centre (x, y), search radius of the molecule this step (its share of the reach + its displacement), displacement
    float2 ref[SMALL_MAXN];            // centre when the pair list was built
    int2 meta[SMALL_MAXN];             // unit key, free-site flags: with the centres of the resident poses, the neighbour record the general path keeps in D.nrec
    unsigned char ligFree[SMALL_MAXN]; // ligand b of the replica is a unit of its own (no receptor bound): moved by the ligand proposal, not by a complex
    unsigned char isSpec[SMALL_MAXN];  // this step the molecule may be further than dmax from its list centre: the list does not cover it
    int spec[SMALL_SPEC];              // those molecules
    int nspec;
    int recBase, ligBase, N;           // local index = gid - recBase (receptor) / gid - ligBase (ligand)
    float dmax;                        // drift the pair list allows for
    float share[2];                    // search_share of a ligand [0] / a receptor [1]
};
KD int small_index(const Consts &K, const Dev &D, int gid) { return gid < K.NAt ? gid - D.small->recBase : gid - D.small->ligBase; }

#define GHOST_BIT 0x40000000
#define TOUCH_CAP 256
// a bond of molecule gid changed: its complex is re-derived at the start of the next step (k_step_begin); beyond TOUCH_CAP
// changes per step the whole table is rebuilt instead
KD void touch_molecule(const Dev &D, int gid) {
    const int i = atomicAdd(&D.scal[S_NTOUCH], 1);
    if (i < TOUCH_CAP) D.touchList[i] = gid; else D.scal[S_TOPO_DIRTY] = 1;
}

KD Rec load_rec(const double2 *C, const double2 *S2, const double2 *S3, int a) {
    double2 c = C[a], s2 = S2[a], s3 = S3[a];
    Rec r; r.cx = c.x; r.cy = c.y; r.s2x = s2.x; r.s2y = s2.y; r.s3x = s3.x; r.s3y = s3.y; return r;
}
KD void store_rec(double2 *C, double2 *S2, double2 *S3, int a, const Rec &r) {
    C[a] = make_double2(r.cx, r.cy); S2[a] = make_double2(r.s2x, r.s2y); S3[a] = make_double2(r.s3x, r.s3y);
}
KD void load_lig(const double *base, int h, Lig &l) {
    const double2 *p = reinterpret_cast<const double2 *>(base + (size_t)h * 24);
    double *d = &l.p[0][0];
#pragma unroll
    for (int q = 0; q < 12; q++) { double2 v = p[q]; d[2 * q] = v.x; d[2 * q + 1] = v.y; }
}
// centre + three beads only (first 96 bytes): all an overlap test needs
KD void load_lig_beads(const double *base, int h, Lig &l) {
    const double2 *p = reinterpret_cast<const double2 *>(base + (size_t)h * 24);
    double *d = &l.p[0][0];
#pragma unroll
    for (int q = 0; q < 6; q++) { double2 v = p[q]; d[2 * q] = v.x; d[2 * q + 1] = v.y; }
}
KD void store_lig(double *base, int h, const Lig &l) {
    double2 *p = reinterpret_cast<double2 *>(base + (size_t)h * 24);
    const double *d = &l.p[0][0];
#pragma unroll
    for (int q = 0; q < 12; q++) p[q] = make_double2(d[2 * q], d[2 * q + 1]);
}

// x in the frame the cells are hashed in: with strips, the periodic image nearest to the strip centre, so that the band of the
// membrane on the far side of the periodic seam sits next to this strip in the grid (distances always use true coordinates:
// the reference has no minimum image, main.cpp:642-646; a molecule starts to interact across the seam only once it is wrapped)
KD double hash_x(const Consts &K, double x) {       // stripHalf = +inf on a single GPU: two compares, never shifts
    const double t = x - K.stripXc;
    return t > K.stripHalf ? x - K.Lx : (t < -K.stripHalf ? x + K.Lx : x);
}   // real branch: the single-GPU path pays nothing
// strip (rank) that owns a unit whose head molecule has its centre at x: strips of equal width along x, periodic
KD int d_strip_owner(const Consts &K, double x) {
    const double L = K.Lx, xw = x - L * round(x / L);
    int r = (int)floor((xw + L / 2) / (L / K.strips));
    return min(max(r, 0), K.strips - 1);
}
KD int nA_live(const Dev &D) { return D.scal[S_NA_LIVE]; }
KD int nB_live(const Dev &D) { return D.scal[S_NB_LIVE]; }
KD bool gid_live(const Consts &K, const Dev &D, int gid) { return gid < K.NAt ? gid < nA_live(D) : (gid < K.NT && gid - K.NAt < nB_live(D)); }
KD int cell_of(const Consts &K, int replica, double x, double y) {
    int cx = (int)floor((hash_x(K, x) - K.gx0) * K.cellInv), cy = (int)floor((y - K.gy0) * K.cellInv);
    cx = min(max(cx, 0), K.ncx - 1); cy = min(max(cy, 0), K.ncy - 1);
    return (replica * K.ncy + cy) * K.ncx + cx;
}
KD int replica_of_gid(const Consts &K, int gid) { return K.phase == 2 ? K.smallRep : (K.R == 1 ? 0 : (gid < K.NAt ? gid / K.NA : (gid - K.NAt) / K.NB)); }
// reference molecule id (1-based) used to key the random stream
KD uint32_t ref_id(const Consts &K, const Dev &D, int gid) {
    if (K.phase == 2) return gid < K.NAt ? (uint32_t)(gid - K.smallRep * K.NA + 1) : (uint32_t)(gid - K.NAt - K.smallRep * K.NB + K.NA + 1);
    if (D.refA) return gid < K.NAt ? D.refA[gid] : D.refB[gid - K.NAt];
    if (K.R == 1) return (uint32_t)(gid + 1);          // one replica: no integer division on the hot path
    return gid < K.NAt ? (uint32_t)(gid % K.NA + 1) : (uint32_t)(K.NA + (gid - K.NAt) % K.NB + 1);
}

}  // namespace kmc
