// host/kmc_main.cpp -- the reference's program shape on top of the C ABI: same parameter set (shipped defaults of
// main.cpp:39-99, overridable on the command line since the reference needs a recompile for that), same start / restart
// behaviour (main.cpp:226-278: continue from position.cpt if there is one, else a fresh random start that truncates the
// output files), same outputs (parameter.log, and bond.dat, cluster.log, test.gro, position.cpt every `output_every` steps,
// main.cpp:179-205, 2206-2305) written into the working directory.
//
//   kmc_main [--steps N] [--output-every M] [--seed S] [--receptors NA --ligands NB --box LX LY LZ] [--replicas R]
//            [--set name=value ...]     names: dt DA DrotA DB DrotB on off cis_on cis_off mono_cis_on mono_cis_off ...
//
// build:  g++ -O2 -I../../include kmc_main.cpp -L.. -lkmc_b200 -Wl,-rpath,'$ORIGIN/..' -o kmc_main
#include "kmc_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main(int argc, char **argv) {
    kmc_params P; kmc_default_params(&P);
    long steps = 20000000;            // simu_step, main.cpp:39
    int every = 5000;                 // main.cpp:2206
    unsigned long long init_seed = 1;
    struct { const char *n; double *p; } D[] = {
        {"dt", &P.dt}, {"rA", &P.rA}, {"DA", &P.DA}, {"DrotA", &P.DrotA}, {"rB", &P.rB}, {"DB", &P.DB}, {"DrotB", &P.DrotB},
        {"mono_cis_on", &P.mono_cis_on}, {"mono_cis_off", &P.mono_cis_off}, {"cis_D", &P.cis_D}, {"cis_Drot", &P.cis_Drot},
        {"cis_on", &P.cis_on}, {"cis_off", &P.cis_off}, {"bond_D", &P.bond_D}, {"bond_Drot", &P.bond_Drot}, {"on", &P.on},
        {"off", &P.off}, {"bond_dist_cut", &P.bond_dist_cut}, {"thetapd_cut", &P.thetapd_cut}, {"thetaot_cut", &P.thetaot_cut},
        {"cis_thetaot_cut", &P.cis_thetaot_cut}, {"cis_dist_cut", &P.cis_dist_cut}};
    for (int a = 1; a < argc; a++) {
        std::string s = argv[a];
        auto next = [&]() { return a + 1 < argc ? argv[++a] : (char *)"0"; };
        if (s == "--steps") steps = atol(next());
        else if (s == "--output-every") every = atoi(next());
        else if (s == "--seed") { P.seed = strtoull(next(), 0, 0); init_seed = P.seed; }
        else if (s == "--receptors") P.n_receptor = atoi(next());
        else if (s == "--ligands") P.n_ligand = atoi(next());
        else if (s == "--replicas") P.n_replicas = atoi(next());
        else if (s == "--box") { P.box[0] = atof(next()); P.box[1] = atof(next()); P.box[2] = atof(next()); }
        else if (s == "--set") {
            std::string kv = next(); size_t eq = kv.find('='); bool ok = false;
            for (auto &d : D) if (eq != std::string::npos && kv.substr(0, eq) == d.n) { *d.p = atof(kv.c_str() + eq + 1); ok = true; }
            if (!ok) { fprintf(stderr, "unknown parameter %s\n", kv.c_str()); return 2; }
        } else { fprintf(stderr, "unknown option %s\n", s.c_str()); return 2; }
    }
    kmc_handle *h = nullptr;
    if (kmc_create(&P, &h)) { fprintf(stderr, "kmc_create: %s\n", kmc_last_error(nullptr)); return 1; }
    kmc_parameter_log_write(&P, "parameter.log");                    // main.cpp:169, 179-205 (truncated at every start, Q17)
    // main.cpp:226-278: a position.cpt in the working directory means restart (outputs are appended), none means a fresh start
    // (outputs truncated). An EMPTY position.cpt -- what a reference run killed before its first output leaves behind, Q21 --
    // is treated as none instead of being read as zeros.
    bool restart = false;
    if (FILE *f = fopen("position.cpt", "r")) { restart = fgetc(f) != EOF; fclose(f); }
    if (restart && P.n_replicas == 1) {
        if (kmc_read_checkpoint(h, 0, "position.cpt")) { fprintf(stderr, "%s\n", kmc_last_error(h)); return 1; }
        printf("CPT file exist\n");                                   // main.cpp:229
    } else {
        if (kmc_init_random(h, init_seed, P.n_receptor + P.n_ligand > 20000)) { fprintf(stderr, "%s\n", kmc_last_error(h)); return 1; }
        remove("bond.dat"); remove("cluster.log"); remove("test.gro"); remove("position.cpt");      // main.cpp:275-278
        printf("CPT file not exist\n");                               // main.cpp:274
    }
    kmc_series s; kmc_get_series(h, 0, &s);
    const long left = steps - (long)s.step;                           // the loop runs to simu_step in total (main.cpp:461)
    if (left > 0 && kmc_run(h, left, every, ".")) { fprintf(stderr, "%s\n", kmc_last_error(h)); return 1; }
    kmc_get_series(h, 0, &s);
    printf("step %lld  R-L %d  mono-cis %d  cis %d  bonds %d  mean complex %.3f  max complex %d\n", (long long)s.step, s.bond_num_rl,
           s.bond_num_mono_cis, s.bond_num_cis, s.bond_num, s.cluster_size, s.max_complex);
    kmc_destroy(h);
    return 0;
}
