// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference translation unit (/root/reference/main.cpp, compiled where it
// lies by oracle/Makefile; nothing of it is copied into this repository) so that its per-timestep
// sweep (main.cpp:461-2308) can be executed deterministically and its state read back.
//
// How: the reference object is post-processed with
//     objcopy --redefine-sym main=ref_main [--weaken-symbol=_Z5rand2v]
// All reference parameters and state are external-linkage globals (main.cpp:39-167), so this
// harness sets parameters through them and dumps R_x/R_y/R_z, protein_status, res_nei, counters.
// With rand2 weakened this file supplies  double rand2()  (main.cpp:2313 replacement) and
// extern "C" int rand()  (consumed by std::random_shuffle, main.cpp:1285/1345/1413/1597) from
// sequential xorshift64 streams, which makes a run bit-reproducible.
//
// An arbitrary start state (bonds included) is injected through the reference's own restart
// mechanism (main.cpp:226-268): the harness writes a full-precision position.cpt into a scratch
// working directory; operator>> parses it exactly (the reference's *writer* is 3-decimal lossy,
// its *reader* is not).
//
// Build variants (see oracle/Makefile):
//   kmcref_<tag>          rand2 injected  (deterministic oracle)
//   kmcref_<tag>_shipped  rand2 as shipped (CPU baseline timing; only main renamed)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <string>
#include <vector>
#include <chrono>
#include <unistd.h>
#include <sys/stat.h>

#ifndef KMC_NA
#define KMC_NA 150
#endif
#ifndef KMC_NB
#define KMC_NB 50
#endif
#define KMC_N (KMC_NA + KMC_NB)

// ---- reference globals (main.cpp:39-167) ----
extern int simu_step;
extern double time_step, cell_range_x, cell_range_y, cell_range_z;
extern double pai, RB_A_radius, RB_A_D, RB_A_rot_D, RB_B_radius, RB_B_D, RB_B_rot_D;
extern double mono_cis_Ass_Rate, mono_cis_Diss_Rate, cis_D, cis_rot_D, cis_Ass_Rate, cis_Diss_Rate;
extern double bond_D, bond_rot_D, Ass_Rate, Diss_Rate;
extern double bond_dist_cutoff, bond_thetapd_cutoff, bond_thetaot_cutoff, cis_thetaot_cutoff, cis_dist_cutoff;
extern double R_x[][5][5], R_y[][5][5], R_z[][5][5];
extern int protein_status[][5];
extern int res_nei[][7];
extern int results[][KMC_N + 1];
extern int bond_num, bond_num_rl, bond_num_cis, bond_num_mono_cis, protein_num_in_Max_Complex;
extern int tot_cluster_num, tot_proteins_in_cluster;
extern double cluster_size;
extern int mc_time_step, initial_simu_time;

extern "C" int ref_main();  // `main` is not mangled, so the objcopy-renamed symbol is plain C

// ---- deterministic streams ----
static uint64_t g_s2 = 88172645463325252ULL;        // rand2 stream state
static uint64_t g_sr = 0x9E3779B97F4A7C15ULL;       // rand  stream state
static uint64_t g_n_rand2 = 0, g_n_rand = 0;
static inline uint64_t xs64(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

// ---- frame dump ----
static FILE *g_out = nullptr;
static long g_frames_every = 0;
static int g_last_seen_step = -1;

static void write_frame(int64_t step) {
    if (!g_out) return;
    int32_t cnt[8] = {bond_num, bond_num_rl, bond_num_cis, bond_num_mono_cis, protein_num_in_Max_Complex,
                      tot_cluster_num, tot_proteins_in_cluster, KMC_N};
    fwrite(&step, 8, 1, g_out);
    fwrite(cnt, 4, 8, g_out);
    fwrite(&cluster_size, 8, 1, g_out);
    fwrite(R_x, sizeof(double), (size_t)(KMC_N + 1) * 25, g_out);
    fwrite(R_y, sizeof(double), (size_t)(KMC_N + 1) * 25, g_out);
    fwrite(R_z, sizeof(double), (size_t)(KMC_N + 1) * 25, g_out);
    fwrite(protein_status, sizeof(int), (size_t)(KMC_N + 1) * 5, g_out);
    fwrite(res_nei, sizeof(int), (size_t)(KMC_N + 1) * 7, g_out);
}

#ifndef KMC_SHIPPED_RAND2
// Replacement for main.cpp:2313-2326 (same distribution: iid U[0,1) with 53 random bits).
double rand2() {
    // Step-boundary hook: at the first draw of step s the committed arrays hold the state after
    // step s-1 (S0 copy and S1 BFS of step s have already run; they do not touch R/res_nei).
    if (g_frames_every > 0 && initial_simu_time >= 1 && mc_time_step != g_last_seen_step) {
        g_last_seen_step = mc_time_step;
        int64_t done = (int64_t)mc_time_step - 1;
        if (done > 0 && done % g_frames_every == 0) write_frame(done);
    }
    g_n_rand2++;
    return (double)(xs64(g_s2) >> 11) * (1.0 / 9007199254740992.0);
}
#endif

// Interposes libc rand() for std::random_shuffle (libstdc++ calls std::rand() % n).
extern "C" int rand(void) __THROW {
    g_n_rand++;
    return (int)((xs64(g_sr) >> 33) & 0x7fffffff);
}

struct NamedD { const char *name; double *p; };
static NamedD g_dbl[] = {
    {"time_step", &time_step}, {"cell_range_x", &cell_range_x}, {"cell_range_y", &cell_range_y},
    {"cell_range_z", &cell_range_z}, {"pai", &pai}, {"RB_A_radius", &RB_A_radius}, {"RB_A_D", &RB_A_D},
    {"RB_A_rot_D", &RB_A_rot_D}, {"RB_B_radius", &RB_B_radius}, {"RB_B_D", &RB_B_D},
    {"RB_B_rot_D", &RB_B_rot_D}, {"mono_cis_Ass_Rate", &mono_cis_Ass_Rate},
    {"mono_cis_Diss_Rate", &mono_cis_Diss_Rate}, {"cis_D", &cis_D}, {"cis_rot_D", &cis_rot_D},
    {"cis_Ass_Rate", &cis_Ass_Rate}, {"cis_Diss_Rate", &cis_Diss_Rate}, {"bond_D", &bond_D},
    {"bond_rot_D", &bond_rot_D}, {"Ass_Rate", &Ass_Rate}, {"Diss_Rate", &Diss_Rate},
    {"bond_dist_cutoff", &bond_dist_cutoff}, {"bond_thetapd_cutoff", &bond_thetapd_cutoff},
    {"bond_thetaot_cutoff", &bond_thetaot_cutoff}, {"cis_thetaot_cutoff", &cis_thetaot_cutoff},
    {"cis_dist_cutoff", &cis_dist_cutoff},
};

// Writes a position.cpt the reference reader (main.cpp:231-266) parses back exactly.
static bool write_cpt_from_frame(const char *frame_path) {
    FILE *f = fopen(frame_path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", frame_path); return false; }
    int64_t step; int32_t cnt[8]; double cs;
    std::vector<double> X((size_t)(KMC_N + 1) * 25), Y(X.size()), Z(X.size());
    std::vector<int> st((size_t)(KMC_N + 1) * 5), rn((size_t)(KMC_N + 1) * 7);
    bool ok = fread(&step, 8, 1, f) == 1 && fread(cnt, 4, 8, f) == 8 && fread(&cs, 8, 1, f) == 1 &&
              fread(X.data(), 8, X.size(), f) == X.size() && fread(Y.data(), 8, Y.size(), f) == Y.size() &&
              fread(Z.data(), 8, Z.size(), f) == Z.size() && fread(st.data(), 4, st.size(), f) == st.size() &&
              fread(rn.data(), 4, rn.size(), f) == rn.size();
    fclose(f);
    if (!ok || cnt[7] != KMC_N) { fprintf(stderr, "bad frame file (N mismatch?)\n"); return false; }
    FILE *c = fopen("position.cpt", "w");
    if (!c) return false;
    auto P = [&](int i, int j, int k) { return (size_t)i * 25 + j * 5 + k; };
    for (int i = 1; i <= KMC_NA; i++) {
        for (int j = 1; j <= 4; j++)
            for (int k = 1; k <= 4; k++)
                fprintf(c, "%.17g %.17g %.17g\n", X[P(i, j, k)], Y[P(i, j, k)], Z[P(i, j, k)]);
        fprintf(c, "%d %d %d %d %d\n", st[i * 5 + 2], st[i * 5 + 3], rn[i * 7 + 2], rn[i * 7 + 4], rn[i * 7 + 3]);
    }
    for (int i = KMC_NA + 1; i <= KMC_N; i++)
        for (int j = 1; j <= 4; j++) {
            for (int k = 1; k <= 2; k++)
                fprintf(c, "%.17g %.17g %.17g\n", X[P(i, j, k)], Y[P(i, j, k)], Z[P(i, j, k)]);
            fprintf(c, "%d %d\n", st[i * 5 + j], rn[i * 7 + j]);
        }
    fprintf(c, "%d\n%d\n%d\n%d\n%d\n%lld\n", cnt[0], cnt[1], cnt[2], cnt[3], cnt[4], (long long)step);
    fclose(c);
    return true;
}

int main(int argc, char **argv) {
    long steps = 1000;
    std::string out_path, in_path, workdir;
    bool timing_only = false;
    for (int a = 1; a < argc; a++) {
        std::string s = argv[a];
        auto next = [&]() -> std::string { return (a + 1 < argc) ? std::string(argv[++a]) : std::string(); };
        if (s == "--steps") steps = atol(next().c_str());
        else if (s == "--out") out_path = next();
        else if (s == "--in") in_path = next();
        else if (s == "--workdir") workdir = next();
        else if (s == "--frames-every") g_frames_every = atol(next().c_str());
        else if (s == "--rand2-state") g_s2 = strtoull(next().c_str(), nullptr, 0);
        else if (s == "--rand-state") g_sr = strtoull(next().c_str(), nullptr, 0);
        else if (s == "--timing") timing_only = true;
        else if (s == "--set") {
            std::string kv = next(); size_t eq = kv.find('=');
            if (eq == std::string::npos) { fprintf(stderr, "--set name=value\n"); return 2; }
            std::string k = kv.substr(0, eq); double v = atof(kv.c_str() + eq + 1); bool found = false;
            for (auto &d : g_dbl) if (k == d.name) { *d.p = v; found = true; }
            if (!found) { fprintf(stderr, "unknown global %s\n", k.c_str()); return 2; }
        } else if (s == "--scale") {  // multiply a global in place, e.g. cis_Ass_Rate*=20
            std::string kv = next(); size_t eq = kv.find('=');
            std::string k = kv.substr(0, eq); double v = atof(kv.c_str() + eq + 1); bool found = false;
            for (auto &d : g_dbl) if (k == d.name) { *d.p *= v; found = true; }
            if (!found) { fprintf(stderr, "unknown global %s\n", k.c_str()); return 2; }
        } else { fprintf(stderr, "unknown option %s\n", s.c_str()); return 2; }
    }
    // resolve paths before chdir
    auto absolutize = [](std::string p) {
        if (p.empty() || p[0] == '/') return p;
        char buf[4096]; if (!getcwd(buf, sizeof buf)) return p; return std::string(buf) + "/" + p;
    };
    out_path = absolutize(out_path); in_path = absolutize(in_path);
    // The reference reads/writes fixed file names in its CWD (main.cpp:169,226,275-278): always run
    // in a fresh empty directory.
    char tmpl[] = "/tmp/kmcref_XXXXXX";
    std::string wd = workdir.empty() ? std::string(mkdtemp(tmpl)) : workdir;
    if (!workdir.empty()) mkdir(wd.c_str(), 0777);
    if (chdir(wd.c_str()) != 0) { perror("chdir"); return 1; }
    remove("position.cpt");
    long start_step = 0;
    if (!in_path.empty()) {
        if (!write_cpt_from_frame(in_path.c_str())) return 1;
        FILE *f = fopen(in_path.c_str(), "rb"); int64_t st = 0; if (fread(&st, 8, 1, f) != 1) st = 0; fclose(f);
        start_step = (long)st;
    }
    simu_step = (int)(start_step + steps);
    if (!out_path.empty()) g_out = fopen(out_path.c_str(), "wb");
    auto t0 = std::chrono::steady_clock::now();
    ref_main();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    write_frame((int64_t)simu_step);
    if (g_out) fclose(g_out);
    printf("{\"N\": %d, \"NA\": %d, \"steps\": %ld, \"seconds\": %.6f, \"moves_per_s\": %.6g, "
           "\"rand2_draws\": %llu, \"rand_draws\": %llu, \"bond_num\": %d, \"bond_num_rl\": %d, "
           "\"bond_num_cis\": %d, \"bond_num_mono_cis\": %d, \"max_complex\": %d, \"cluster_size\": %.17g}\n",
           KMC_N, KMC_NA, steps, sec, (double)KMC_N * steps / sec, (unsigned long long)g_n_rand2,
           (unsigned long long)g_n_rand, bond_num, bond_num_rl, bond_num_cis, bond_num_mono_cis,
           protein_num_in_Max_Complex, cluster_size);
    (void)timing_only;
    if (workdir.empty()) {
        const char *files[] = {"position.cpt", "bond.dat", "test.gro", "cluster.log", "parameter.log"};
        for (auto fn : files) remove(fn);
        if (chdir("/") == 0) rmdir(wd.c_str());
    }
    return 0;
}
