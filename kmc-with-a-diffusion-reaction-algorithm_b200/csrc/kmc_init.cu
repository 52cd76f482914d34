// csrc/kmc_init.cu -- initial configuration on the GPU (included by kmc_engine.cu). Replaces main.cpp:273-456.
//
// The reference inserts molecules one after the other: molecule i draws positions until one is free of every molecule placed
// before it (receptors: centre distance > 2 rA to earlier receptors, main.cpp:284-296; ligands: centre > rA + 2rB/sqrt3 + rB from
// every receptor bead and > 2(2rB/sqrt3 + rB) from earlier ligands, main.cpp:354-383), O(N^2) with goto-retries. With keyed draws
// -- candidate t of molecule i is a pure function of (seed, i, t) -- the same sequential rule has a parallel evaluation: the
// attempt number of molecule i is
//        t_i = min { t : candidate(i, t) clashes with no molecule j < i at candidate(j, t_j) },
// a fixed point that Jacobi iteration reaches in as many rounds as the longest chain of clashes (area fraction 0.6 %: a handful).
// Every round hashes the current candidates into a cell grid (linked lists) and re-evaluates t_i for all molecules at once.
// Result: exactly the configuration the sequential insertion with these draws would produce, in a few ms for 1e7 molecules.
// Orientations as in the reference: receptor psi ~ U(-pai, pai) (main.cpp:328-350), ligand three Euler angles ~ U(-pai, pai)
// (main.cpp:422-446). sort_cells renumbers molecules in cell-major order (deterministic: by generation id inside a cell).
namespace kmc {

struct GenGrid { double x0, y0, inv; int ncx, ncy; };
struct GenArgs {
    GenGrid G; double Lx, Ly, Lz, rA, rB, pai, exRR, exRL, exLL; uint64_t seed;
    int NA, NB, R;                      // per replica sizes, replicas
    double3 *pos; int *att, *attNew, *head, *next, *flag;
};
#define GEN_MAX_ATTEMPTS 256
#define SLOT_GEN_XY 32u
#define SLOT_GEN_Z 34u
#define SLOT_GEN_ROT 36u
#define SLOT_GEN_ROT2 38u

KD void gen_decode(const GenArgs &g, int i, bool &lig, int &rep, uint32_t &mol) {
    const int NAt = g.NA * g.R;
    lig = i >= NAt;
    const int q = lig ? i - NAt : i;
    rep = lig ? q / g.NB : q / g.NA;
    mol = lig ? (uint32_t)(g.NA + q % g.NB + 1) : (uint32_t)(q % g.NA + 1);
}
KD double3 gen_candidate(const GenArgs &g, bool lig, int rep, uint32_t mol, int t) {
    double u, v; keyed_uniform2(g.seed + (uint64_t)rep, mol, (uint32_t)t, 0, SLOT_GEN_XY, u, v);
    double3 p = make_double3(u * g.Lx - g.Lx / 2, v * g.Ly - g.Ly / 2, 0.0);
    if (lig) p.z = keyed_uniform(g.seed + (uint64_t)rep, mol, (uint32_t)t, 0, SLOT_GEN_Z) * g.Lz;
    return p;
}
KD int gen_cell(const GenArgs &g, int rep, double x, double y) {
    int cx = (int)floor((x - g.G.x0) * g.G.inv), cy = (int)floor((y - g.G.y0) * g.G.inv);
    cx = min(max(cx, 0), g.G.ncx - 1); cy = min(max(cy, 0), g.G.ncy - 1);
    return (rep * g.G.ncy + cy) * g.G.ncx + cx;
}
// does a molecule of index i (species lig) at p clash with any molecule j < i of its replica at its current position?
KD bool gen_clash(const GenArgs &g, int i, bool lig, int rep, double3 p) {
    const int NAt = g.NA * g.R;
    int cx = (int)floor((p.x - g.G.x0) * g.G.inv), cy = (int)floor((p.y - g.G.y0) * g.G.inv);
    cx = min(max(cx, 0), g.G.ncx - 1); cy = min(max(cy, 0), g.G.ncy - 1);
    for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.G.ncy - 1); yy++)
        for (int xx = max(cx - 1, 0); xx <= min(cx + 1, g.G.ncx - 1); xx++)
            for (int j = g.head[(rep * g.G.ncy + yy) * g.G.ncx + xx]; j >= 0; j = g.next[j]) {
                if (j >= i) continue;                            // only molecules inserted earlier (receptors come before ligands)
                const double3 q = g.pos[j];
                const double dx = p.x - q.x, dy = p.y - q.y;
                if (!lig) { if (sqrt(dx * dx + dy * dy) <= g.exRR) return true; }
                else if (j < NAt) {
                    for (int k = 0; k < 4; k++) { const double dz = p.z - 2 * k * g.rA; if (sqrt(dx * dx + dy * dy + dz * dz) <= g.exRL) return true; }
                } else { const double dz = p.z - q.z; if (sqrt(dx * dx + dy * dy + dz * dz) <= g.exLL) return true; }
            }
    return false;
}
__global__ void k_gen_place(GenArgs g, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool lig; int rep; uint32_t mol; gen_decode(g, i, lig, rep, mol);
    const double3 p = gen_candidate(g, lig, rep, mol, g.att[i]);
    g.pos[i] = p;
    g.next[i] = atomicExch(&g.head[gen_cell(g, rep, p.x, p.y)], i);
}
__global__ void k_gen_check(GenArgs g, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool lig; int rep; uint32_t mol; gen_decode(g, i, lig, rep, mol);
    const int cur = g.att[i];
    int t = 0;
    for (; t < GEN_MAX_ATTEMPTS; t++) {
        const double3 p = t == cur ? g.pos[i] : gen_candidate(g, lig, rep, mol, t);
        if (!gen_clash(g, i, lig, rep, p)) break;
    }
    if (t >= GEN_MAX_ATTEMPTS) { g.flag[1] = 1; t = cur; }          // box too dense
    g.attNew[i] = t;
    if (t != cur) g.flag[0] = 1;
}
// cell-major renumbering: index of molecule i inside its species = start of its cell + number of same-species molecules of the
// cell with a lower generation id
__global__ void k_gen_count(GenArgs g, int n, int *countA, int *countB) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool lig; int rep; uint32_t mol; gen_decode(g, i, lig, rep, mol);
    const double3 p = g.pos[i];
    atomicAdd(&(lig ? countB : countA)[gen_cell(g, rep, p.x, p.y)], 1);
}
// poses out: rec[dest][6], lig[dest][24]. startA/startB null = keep the generation order.
__global__ void k_gen_emit(GenArgs g, int n, const int *startA, const int *startB, double *recOut, double *ligOut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool lig; int rep; uint32_t mol; gen_decode(g, i, lig, rep, mol);
    const int NAt = g.NA * g.R;
    const double3 p = g.pos[i];
    int dest = lig ? i - NAt : i;
    if (startA) {
        const int c = gen_cell(g, rep, p.x, p.y);
        int lower = 0;
        for (int j = g.head[c]; j >= 0; j = g.next[j]) lower += (j < i) && ((j >= NAt) == lig);
        dest = (lig ? startB : startA)[c] + lower;
    }
    const uint64_t seed = g.seed + (uint64_t)rep;
    if (!lig) {
        const double psai = (2 * keyed_uniform(seed, mol, 0, 0, SLOT_GEN_ROT) - 1) * g.pai;
        double s, c; sincos(psai, &s, &c);
        double *o = recOut + (size_t)dest * 6;
        o[0] = p.x; o[1] = p.y; o[2] = c * g.rA + p.x; o[3] = s * g.rA + p.y; o[4] = c * (-g.rA) + p.x; o[5] = s * (-g.rA) + p.y;
    } else {
        double u0, u1; keyed_uniform2(seed, mol, 0, 0, SLOT_GEN_ROT, u0, u1);
        const double th = (2 * u0 - 1) * g.pai, ph = (2 * u1 - 1) * g.pai, ps = (2 * keyed_uniform(seed, mol, 0, 0, SLOT_GEN_ROT2) - 1) * g.pai;
        double st, ct, sp, cp, ss, cs; sincos(th, &st, &ct); sincos(ph, &sp, &cp); sincos(ps, &ss, &cs);
        const double t[3][3] = {{cs * cp - ct * sp * ss, -ss * cp - ct * sp * cs, st * sp},
                                {cs * sp + ct * cp * ss, -ss * sp + ct * cp * cs, -st * cp},
                                {ss * st, cs * st, ct}};
        const double rB = g.rB, q3 = sqrt(3.0);
        const double tpl[8][3] = {{0, 0, 0}, {0, rB * 2 / q3, 0}, {-rB, -rB / q3, 0}, {rB, -rB / q3, 0}, {0, 0, rB},
                                  {0, rB * (2 / q3 + 1), 0}, {-rB * (q3 / 2 + 1), -rB / q3 - rB / 2, 0}, {rB * (q3 / 2 + 1), -rB / q3 - rB / 2, 0}};
        const double c3[3] = {p.x, p.y, p.z};
        double *o = ligOut + (size_t)dest * 24;
        for (int q = 0; q < 8; q++)
            for (int d = 0; d < 3; d++) o[q * 3 + d] = t[d][0] * tpl[q][0] + t[d][1] * tpl[q][1] + t[d][2] * tpl[q][2] + c3[d];
    }
}

}  // namespace kmc

// Generates R replicas of (NA receptors + NB ligands) into device buffers recOut[R*NA][6], ligOut[R*NB][24] on stream st.
// Returns nullptr or an error text. Temporary device memory: ~40 B per molecule + 12 B per grid cell, freed before returning.
static const char *generate_random_device(const kmc_params &P, int NA, int NB, int R, uint64_t seed, int sort_cells, double *recOut, double *ligOut,
                                          cudaStream_t st, int *rounds_out) {
    using namespace kmc;
    const int n = (NA + NB) * R;
    if (n <= 0) return nullptr;
    const double rs = P.rB * 2 / sqrt(3.0);
    GenArgs g; memset(&g, 0, sizeof g);
    g.Lx = P.box[0]; g.Ly = P.box[1]; g.Lz = P.box[2]; g.rA = P.rA; g.rB = P.rB; g.pai = P.pai; g.seed = seed; g.NA = NA; g.NB = NB; g.R = R;
    g.exRR = P.rA + P.rA; g.exRL = P.rA + rs + P.rB; g.exLL = rs + rs + 2 * P.rB;               // main.cpp:293, 368, 380
    // cells at least one exclusion radius wide, about one molecule per cell
    const double edge = std::max(std::max({g.exRR, g.exRL, g.exLL}) + 1.0, sqrt(P.box[0] * P.box[1] / std::max(NA + NB, 1)));
    g.G.x0 = -P.box[0] / 2; g.G.y0 = -P.box[1] / 2; g.G.inv = 1.0 / edge;
    g.G.ncx = std::max(1, (int)ceil(P.box[0] / edge)); g.G.ncy = std::max(1, (int)ceil(P.box[1] / edge));
    const size_t ncell = (size_t)R * g.G.ncx * g.G.ncy;
    if (ncell >= (1ull << 31) - 4096) return "kmc_init_random: generator grid too large";
    const size_t scanBlocks = (ncell + 1 + SCAN_TILE - 1) / SCAN_TILE, padded = scanBlocks * SCAN_TILE;
    int *cnt[2] = {nullptr, nullptr}, *start[2] = {nullptr, nullptr}, *tmp = nullptr;
    bool ok = cudaMalloc(&g.pos, sizeof(double3) * (size_t)n) == cudaSuccess && cudaMalloc(&g.att, sizeof(int) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&g.attNew, sizeof(int) * (size_t)n) == cudaSuccess && cudaMalloc(&g.next, sizeof(int) * (size_t)n) == cudaSuccess &&
              cudaMalloc(&g.head, sizeof(int) * ncell) == cudaSuccess && cudaMalloc(&g.flag, 2 * sizeof(int)) == cudaSuccess;
    if (sort_cells)
        for (int s = 0; s < 2 && ok; s++) ok = cudaMalloc(&cnt[s], sizeof(int) * padded) == cudaSuccess && cudaMalloc(&start[s], sizeof(int) * padded) == cudaSuccess;
    if (sort_cells && ok) ok = cudaMalloc(&tmp, sizeof(int) * (scanBlocks + 1)) == cudaSuccess;
    const char *err = ok ? nullptr : "kmc_init_random: device allocation failed";
    int rounds = 0;
    if (ok) {
        cudaMemsetAsync(g.att, 0, sizeof(int) * (size_t)n, st);
        const int nb_ = (n + 255) / 256;
        for (;; rounds++) {
            if (rounds > 200) { err = "kmc_init_random: insertion did not converge"; break; }
            cudaMemsetAsync(g.head, 0xff, sizeof(int) * ncell, st);
            cudaMemsetAsync(g.flag, 0, 2 * sizeof(int), st);
            k_gen_place<<<nb_, 256, 0, st>>>(g, n);
            k_gen_check<<<nb_, 256, 0, st>>>(g, n);
            int flag[2] = {0, 0};
            cudaMemcpyAsync(flag, g.flag, sizeof flag, cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) { err = "kmc_init_random: CUDA error"; break; }
            if (flag[1]) { err = "kmc_init_random: cannot place molecules (box too dense)"; break; }
            if (!flag[0]) break;                               // fixed point: head/next/pos describe the final configuration
            std::swap(g.att, g.attNew);
        }
    }
    if (!err) {
        if (sort_cells) {
            for (int s = 0; s < 2; s++) cudaMemsetAsync(cnt[s], 0, sizeof(int) * padded, st);
            k_gen_count<<<(n + 255) / 256, 256, 0, st>>>(g, n, cnt[0], cnt[1]);
            for (int s = 0; s < 2; s++) {
                k_scan_reduce<<<(unsigned)scanBlocks, 256, 0, st>>>((const int4 *)cnt[s], tmp);
                k_scan_sums<<<1, 1024, 0, st>>>(tmp, (int)scanBlocks);
                k_scan_down<<<(unsigned)scanBlocks, 256, 0, st>>>((int4 *)cnt[s], tmp, (int4 *)start[s]);
            }
        }
        k_gen_emit<<<(n + 255) / 256, 256, 0, st>>>(g, n, sort_cells ? start[0] : nullptr, sort_cells ? start[1] : nullptr, recOut, ligOut);
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) err = "kmc_init_random: CUDA error";
    }
    if (rounds_out) *rounds_out = rounds;
    cudaFree(g.pos); cudaFree(g.att); cudaFree(g.attNew); cudaFree(g.next); cudaFree(g.head); cudaFree(g.flag);
    for (int s = 0; s < 2; s++) { cudaFree(cnt[s]); cudaFree(start[s]); }
    cudaFree(tmp);
    return err;
}
