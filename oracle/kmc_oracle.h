// oracle/kmc_oracle.h -- TEST INFRASTRUCTURE (C API of the CPU restatement, loaded with ctypes).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct kmco_params {
    // 1:1 with the reference globals, main.cpp:39-99
    double box[3];                 // cell_range_x/y/z              main.cpp:43-45
    double dt;                     // time_step                     main.cpp:40
    double pai;                    // main.cpp:71 (3.1415926, NOT M_PI)
    double rA, DA, DrotA;          // RB_A_radius, RB_A_D, RB_A_rot_D   72-74
    double rB, DB, DrotB;          // RB_B_*                            76-78
    double mono_cis_on, mono_cis_off;            // 80-81
    double cis_D, cis_Drot, cis_on, cis_off;     // 83-86
    double bond_D, bond_Drot, on, off;           // 88-91
    double bond_dist_cut, thetapd_cut, thetaot_cut, cis_thetaot_cut, cis_dist_cut;  // 93,95,97-99
    int32_t n_receptor, n_ligand;  // protein_A_tot_num, protein_B_tot_num  48, 57
    int32_t stream_mode;           // 0 = sequential xorshift64 streams (ref_harness.cpp), 1 = keyed Philox
    int32_t use_grid;              // 0 = all-pairs loops like the reference, 1 = O(N) cell grid (same results)
    uint64_t seed;                 // keyed mode
    uint64_t rand2_state, rand_state;  // sequential mode stream states
    // sweep order of S2 (main.cpp:577). 0 = molecule index, the reference's order. 1 = checkerboard: units are visited colour by
    // colour of the grid cell (2x2 colouring, cell = floor((x - order_x0)*order_inv_edge)) of their head molecule's committed centre,
    // by index inside a colour; a cis dimer is moved by its lower index. NOT in the reference: it restates the product's
    // production mode with the reference's own step functions, so that mode can be checked bit for bit as well.
    int32_t order_mode, pad_;
    double order_x0, order_y0, order_inv_edge;
} kmco_params;

void kmco_default_params(kmco_params *p);
void *kmco_create(const kmco_params *p);
void kmco_destroy(void *h);
// main.cpp:273-456 (sequential stream mode only)
void kmco_init_reference(void *h);
// reference-shaped arrays: R_x/R_y/R_z double[N+1][5][5], protein_status int[N+1][5], res_nei int[N+1][7]
void kmco_set_state(void *h, const double *Rx, const double *Ry, const double *Rz, const int32_t *status,
                    const int32_t *res_nei, int64_t step_done, int32_t max_complex);
void kmco_get_state(void *h, double *Rx, double *Ry, double *Rz, int32_t *status, int32_t *res_nei);
void kmco_step(void *h, int64_t nsteps);   // main.cpp:461-2202, nsteps times
// counts[8] = bond_num, bond_num_rl, bond_num_cis, bond_num_mono_cis, max_complex, tot_cluster_num,
//             tot_proteins_in_cluster, N;  returns cluster_size (main.cpp:2200-2202)
double kmco_get_counts(void *h, int32_t *counts, int64_t *step_done, uint64_t *n_rand2, uint64_t *n_rand);
// rows of `results` after the last step (post-shuffle member order, main.cpp:2294-2301):
// row_len[n_ligand]; members concatenated into `members` (capacity cap). returns total members.
int64_t kmco_get_results(void *h, int32_t *row_len, int32_t *members, int64_t cap);
// per-step decision log of the LAST step: accept flag per molecule (1 = its unit's move was kept, 0 = reverted)
void kmco_get_accept(void *h, int32_t *accepted /*[N+1]*/);
// event counters accumulated since creation: [0] R-L on, [1] mono-cis on, [2] cis on, [3] R-L off,
// [4] mono-cis off, [5] cis off, [6] unit moves reverted, [7] unit moves tried, [8] lay-downs, [9] goto-lable4 taken
void kmco_get_events(void *h, int64_t *ev /*[16]*/);
#ifdef __cplusplus
}
#endif
