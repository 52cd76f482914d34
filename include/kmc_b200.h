/* include/kmc_b200.h -- C ABI of the B200-native KMC sweep (drop-in boundary).
 *
 * The reference (xiaopuren/KMC-with-a-diffusion-reaction-algorithm, main.cpp) exposes no plugin or
 * FFI interface: its "API" is (i) the parameter globals main.cpp:39-99, (ii) the state arrays
 * main.cpp:102-118 and (iii) the output files written every 5000 steps (main.cpp:2206-2305).
 * This header carries exactly that parameter set, the state in the reference's own array shapes,
 * and the quantities the reference writes, so that main.cpp's time-step loop (main.cpp:461-2308)
 * can be replaced by
 *
 *     kmc_set_state(h, 0, R_x, R_y, R_z, protein_status, res_nei, step, protein_num_in_Max_Complex);
 *     kmc_step(h, n);
 *     kmc_get_state(h, 0, R_x, R_y, R_z, protein_status, res_nei);   kmc_get_series(h, 0, &s);
 *
 * (binding shown in INTEGRATION.md). Plain pointers and sizes only; no CUDA or torch types.
 * All functions return 0 on success or a negative kmc_status; kmc_last_error() gives the text.
 * A handle is owned by one host thread at a time. There is NO CPU fallback: every entry point that
 * computes requires a CUDA device of compute capability 10.0 (sm_100a cubin) and fails otherwise.
 */
#ifndef KMC_B200_H
#define KMC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KMC_ABI_VERSION 1

typedef enum kmc_status {
    KMC_OK = 0,
    KMC_ERR_INVALID = -1,     /* bad argument */
    KMC_ERR_CUDA = -2,        /* CUDA runtime error (no device, launch failure, ...) */
    KMC_ERR_CAPACITY = -3,    /* an internal device buffer overflowed (reported, never silently dropped) */
    KMC_ERR_STATE = -4,       /* inconsistent state handed in (bond table asymmetric, ...) */
    KMC_ERR_IO = -5
} kmc_status;

/* Update order of the diffusion sweep (main.cpp:577, S2). */
#define KMC_MODE_REPLAY 0      /* molecule-index order: reproduces the reference's sequential Gauss-Seidel
                                  sweep exactly (decisions bit-exact, positions to libm rounding) */
#define KMC_MODE_PRODUCTION 1  /* checkerboard order: units are swept colour by colour of their cell
                                  (2x2 colouring), statistically equivalent; same kernels */

typedef struct kmc_params {
    /* 1:1 with the reference globals ------------------------------------------ main.cpp */
    double box[3];                 /* cell_range_x, cell_range_y, cell_range_z      43-45 */
    double dt;                     /* time_step (ns)                                   40 */
    double pai;                    /* 3.1415926 (the reference's own pi)               71 */
    double rA, DA, DrotA;          /* RB_A_radius, RB_A_D, RB_A_rot_D               72-74 */
    double rB, DB, DrotB;          /* RB_B_radius, RB_B_D, RB_B_rot_D               76-78 */
    double mono_cis_on, mono_cis_off;             /* mono_cis_Ass_Rate / _Diss_Rate 80-81 */
    double cis_D, cis_Drot, cis_on, cis_off;      /* cis_*                          83-86 */
    double bond_D, bond_Drot, on, off;            /* bond_D, bond_rot_D, Ass_Rate, Diss_Rate 88-91 */
    double bond_dist_cut;          /* bond_dist_cutoff                                 93 */
    double thetapd_cut;            /* bond_thetapd_cutoff                              95 */
    double thetaot_cut;            /* bond_thetaot_cutoff                              97 */
    double cis_thetaot_cut;        /* cis_thetaot_cutoff                               98 */
    double cis_dist_cut;           /* cis_dist_cutoff                                  99 */
    int32_t n_receptor;            /* protein_A_tot_num (per replica)                  48 */
    int32_t n_ligand;              /* protein_B_tot_num (per replica)                  57 */
    /* build-side additions ------------------------------------------------------------- */
    int32_t n_replicas;            /* independent copies of the system advanced together (>=1) */
    int32_t mode;                  /* KMC_MODE_* */
    uint64_t seed;                 /* Philox key; replica r uses seed + r */
    double cell_edge;              /* neighbour-grid cell edge in Angstrom; 0 = automatic */
    int32_t device;                /* CUDA device ordinal */
    int32_t min_image;             /* 0 = plain Euclidean distances, the reference's behaviour (main.cpp:642-646: positions are wrapped,
                                      distances never use the minimum image). 1 is refused with KMC_ERR_INVALID (not implemented). */
} kmc_params;

/* the bond.dat columns, main.cpp:2251 (plus the step they belong to) */
typedef struct kmc_series {
    int64_t step;                  /* mc_time_step of the last completed step */
    int32_t bond_num_rl;           /* receptor-ligand bonds                  */
    int32_t bond_num_mono_cis;     /* cis bonds between two ligand-free receptors */
    int32_t bond_num_cis;          /* cis bonds with >=1 ligand-bound receptor */
    int32_t bond_num;              /* all bonds                              */
    int32_t max_complex;           /* protein_num_in_Max_Complex: running max, never reset (main.cpp:896-898) */
    int32_t n_complexes;           /* tot_cluster_num: ligand-rooted complexes of size > 1 at the last step */
    int32_t n_in_complexes;        /* tot_proteins_in_cluster */
    int32_t reserved;
    double cluster_size;           /* main.cpp:2200-2202 */
} kmc_series;

typedef struct kmc_handle kmc_handle;

int kmc_abi_version(void);
/* fills the reference's shipped defaults, main.cpp:39-99 */
void kmc_default_params(kmc_params *p);
int kmc_create(const kmc_params *p, kmc_handle **out);
void kmc_destroy(kmc_handle *h);
const char *kmc_last_error(const kmc_handle *h);     /* h may be NULL: error of the last failed kmc_create */

/* Scalable equivalent of the random sequential insertion main.cpp:273-456 (same exclusion radii and
 * orientation distributions; O(N) with a cell grid). Replica r is seeded with init_seed + r.
 * sort_cells != 0 numbers the molecules in cell-major order (memory locality for large membranes). */
int kmc_init_random(kmc_handle *h, uint64_t init_seed, int32_t sort_cells);
/* the same generator as pure host code (no handle, no device): rec_pose[n_replicas*n_receptor][6], lig_pose[n_replicas*n_ligand][24] */
int kmc_generate_packed(const kmc_params *p, uint64_t init_seed, int32_t sort_cells, double *rec_pose, double *lig_pose);

/* State in the reference's array shapes (main.cpp:102-118), one replica at a time:
 *   R_x, R_y, R_z  double[N+1][5][5]   (1-based; receptor uses [1..4][1..4], ligand [1..4][1..2])
 *   protein_status int[N+1][5],  res_nei int[N+1][7]       N = n_receptor + n_ligand
 * step_done = number of completed steps (mc_time_step of the last one), max_complex = running max. */
int kmc_set_state(kmc_handle *h, int32_t replica, const double *R_x, const double *R_y, const double *R_z,
                  const int32_t *protein_status, const int32_t *res_nei, int64_t step_done, int32_t max_complex);
int kmc_get_state(kmc_handle *h, int32_t replica, double *R_x, double *R_y, double *R_z,
                  int32_t *protein_status, int32_t *res_nei);

/* Compact pose exchange for large systems (all replicas): receptors double[n][6] = centre xy, site-2 xy,
 * site-3 xy; ligands double[n][24] = points (1,1),(2,1),(3,1),(4,1),(1,2),(2,2),(3,2),(4,2) x xyz;
 * bonds int32: rec_lig[n] (0-based ligand index or -1), rec_site[n] (2..4 or 0), rec_cis[n] (receptor index or -1). */
int kmc_get_packed(kmc_handle *h, double *rec_pose, double *lig_pose, int32_t *rec_lig, int32_t *rec_site, int32_t *rec_cis);
int kmc_set_packed(kmc_handle *h, const double *rec_pose, const double *lig_pose, const int32_t *rec_lig,
                   const int32_t *rec_site, const int32_t *rec_cis, int64_t step_done);
/* kmc_get_packed without waiting: the state as it stands after the steps enqueued so far is copied into a device-side snapshot on
 * the handle's stream -- kmc_step may be called again at once -- and moved to the caller's buffers (PINNED host memory, else the
 * copy is not asynchronous) on a copy stream of the handle. The buffers are complete when kmc_snapshot_wait returns. One snapshot
 * in flight per handle: a second call queues behind the first. (The reference writes its records every 5000 steps, main.cpp:2206:
 * this is how those leave the device while the next 5000 steps already run.) */
int kmc_get_packed_async(kmc_handle *h, double *rec_pose, double *lig_pose, int32_t *rec_lig, int32_t *rec_site, int32_t *rec_cis);
int kmc_snapshot_wait(kmc_handle *h);

/* Advance n time steps (main.cpp:461-2202 each). Asynchronous on the handle's stream except for the
 * conflict-resolution count it reads back each step; kmc_sync waits for completion. */
int kmc_step(kmc_handle *h, int64_t n);
int kmc_sync(kmc_handle *h);
/* kmc_step bracketed by CUDA events recorded on the handle's own stream; elapsed device time in ms */
int kmc_step_timed(kmc_handle *h, int64_t n, double *elapsed_ms);
/* per-kernel timing (CUDA events around every launch of the sweep). enable=1 resets and starts, 0 stops.
 * kmc_profile_get(idx) returns 1 past the last kernel. */
int kmc_profile(kmc_handle *h, int32_t enable);
int kmc_profile_get(kmc_handle *h, int32_t idx, const char **name, double *total_ms, int64_t *launches);
/* diagnostics: with KMC_TIMELINE=1 in the environment at kmc_create, prints (stderr) where every kernel of the last step graph ran on the device clock */
int kmc_timeline_print(kmc_handle *h);

/* Outputs the reference writes at its output cadence */
int kmc_get_series(kmc_handle *h, int32_t replica, kmc_series *out);
/* rows of `results` (main.cpp:2294-2301): for each ligand l (0-based) row_len[l] members (reference 1-based
 * molecule ids, member order as left by the last step incl. its shuffles), concatenated in `members`.
 * returns total number of members or a negative status. */
int64_t kmc_get_complexes(kmc_handle *h, int32_t replica, int32_t *row_len, int32_t *members, int64_t cap);
/* histogram of ligand-rooted complex sizes over replica (or all replicas if replica < 0): hist[s] = number of
 * complexes with s members, sizes >= nbins-1 are accumulated in the last bin */
int kmc_get_oligomer_hist(kmc_handle *h, int32_t replica, int64_t *hist, int32_t nbins);
/* component label of every molecule after the last step: root[i] (1-based, i = 1..N; root[0] = 0) = reference id of the head of the
 * unit molecule i moves with -- the BFS root ligand of its complex (main.cpp:525-560: the lowest-index ligand), or, for a ligand-free
 * receptor, itself / the lower index of its cis pair (main.cpp:682-688) */
int kmc_get_complex_labels(kmc_handle *h, int32_t replica, int32_t *root_of_molecule);
/* geometry of the neighbour grid (cell = floor((x - x0) * inv_edge)); the checkerboard order of KMC_MODE_PRODUCTION colours these cells */
int kmc_get_grid(kmc_handle *h, double *x0, double *y0, double *inv_edge, int32_t *ncx, int32_t *ncy);
/* replay diagnostics: accepted[N+1] (1-based): 1 if the molecule's unit kept its move in the last step */
int kmc_get_accept(kmc_handle *h, int32_t replica, int32_t *accepted);
/* ev[16] accumulated since creation: [0] R-L on, [1] mono-cis on, [2] cis on, [3] R-L off, [4] mono-cis off,
 * [5] cis off, [6] unit moves reverted, [7] unit moves tried, [8] far movers, [9] conflict-resolution passes,
 * [10] complex-table rebuilds, [11] kernels launched; sizes of the last step's work lists: [12] entry pairs in the neighbour
 * list, [13] special entries (far movers / drifted molecules of a list-reuse step), [14] pending findings, [15] pre-selected
 * reaction pairs */
int kmc_get_events(kmc_handle *h, int64_t *ev);
/* Pure host arithmetic (no device): the alignment passes test AreSame(distance, D), i.e. fabs(sqrt(q) - D) < 1e-8 (main.cpp:2368-2371
 * on the distances of 1205-1215 / 1245-1255), dozens of times per complex. sqrt is correctly rounded and monotone, so the squared
 * distances q that pass form an interval; the kernels compare q with its two ends instead of taking the root. window[0..1] = that
 * interval for the length D (ends included); returns 1 if it could not be established (the kernels then take the root). */
int kmc_alignment_window(double length, double *window);
/* which implementation of the step the handle uses: 0 = the general multi-kernel path (one CUDA graph per step), 1 = the fused
 * small-system step (csrc/kmc_small.cu: replicas of at most 512 molecules, one CTA per replica, the whole step in one kernel,
 * many steps per launch). Chosen at kmc_create; the two give bit-identical results. KMC_FUSED=0 in the environment forces 0. */
int kmc_get_step_path(kmc_handle *h);
/* molecules currently held by the handle (all of n_receptor / n_ligand x replicas unless strips made them capacities) */
int kmc_get_live_counts(kmc_handle *h, int32_t *n_rec, int32_t *n_lig);

/* Pure host formatting of the reference's records (no device needed): one bond.dat line (main.cpp:2249-2251) and one
 * cluster.log frame (main.cpp:2293-2301). Return the number of characters written (excluding the NUL) or a negative status. */
int kmc_format_bond_dat(double dt, const kmc_series *s, char *buf, int32_t cap);
int64_t kmc_format_cluster_log(double dt, int64_t step, int32_t n_ligand, const int32_t *row_len, const int32_t *members,
                               char *buf, int64_t cap);
/* Byte-compatible writers for the reference's output files (append one record, main.cpp:2247-2253, 2291-2305) */
int kmc_write_bond_dat(kmc_handle *h, int32_t replica, const char *path);
int kmc_write_cluster_log(kmc_handle *h, int32_t replica, const char *path);
/* The reference's other text files. Array-level functions are pure host code on reference-shaped arrays (no device):
 * test.gro frame (main.cpp:2258-2287), position.cpt writer/reader (main.cpp:2206-2244 / 226-268; counters[6] = bond_num,
 * bond_num_rl, bond_num_cis, bond_num_mono_cis, protein_num_in_Max_Complex, mc_time_step), parameter.log (main.cpp:179-205). */
int kmc_gro_append_arrays(const char *path, int32_t NA, int32_t NB, const double *R_x, const double *R_y, const double *R_z,
                          double dt, int64_t step, const double *box);
int kmc_checkpoint_write_arrays(const char *path, int32_t NA, int32_t NB, const double *R_x, const double *R_y, const double *R_z,
                                const int32_t *protein_status, const int32_t *res_nei, const int64_t *counters);
int kmc_checkpoint_read_arrays(const char *path, int32_t NA, int32_t NB, double *R_x, double *R_y, double *R_z,
                               int32_t *protein_status, int32_t *res_nei, int64_t *counters);
int kmc_parameter_log_write(const kmc_params *p, const char *path);
int kmc_write_gro(kmc_handle *h, int32_t replica, const char *path);
int kmc_write_checkpoint(kmc_handle *h, int32_t replica, const char *path);
int kmc_read_checkpoint(kmc_handle *h, int32_t replica, const char *path);      /* restart: main.cpp:226-268 */
/* Lossless binary checkpoint of the whole handle (all replicas): position.cpt keeps three decimals (main.cpp:2213), so a restart
 * from it cannot continue a run bit for bit; this one can (header "KMCB2001", sizes, step, running-max complex sizes, then the
 * arrays of kmc_get_packed). The handle must have been created with the same n_receptor / n_ligand / n_replicas. */
int kmc_write_checkpoint_bin(kmc_handle *h, const char *path);
int kmc_read_checkpoint_bin(kmc_handle *h, const char *path);
/* the reference's own main loop: n_steps steps with records every output_every steps into directory `dir` */
int kmc_run(kmc_handle *h, int64_t n_steps, int32_t output_every, const char *dir);

/* ---- strip decomposition of one membrane across GPUs (one handle per rank; see csrc/kmc_strips.cu) ------------------------
 * The handle is created with n_receptor / n_ligand = local CAPACITIES (owned + halo molecules), then configured as rank
 * `rank` of `nranks` strips along x with halo width `halo_width` (Angstrom). The caller steps all ranks by the same number of
 * steps and, every k steps, refreshes: begin_refresh on every rank, message(0) -> rank-1 (received there as from_high),
 * message(1) -> rank+1 (received as from_low), ranks 0 and nranks-1 being neighbours through the periodic seam; then rebuild.
 * Exactness needs halo_width >= k * (interaction reach + 2 max displacement per step) + largest complex extent.
 * message(2) after begin_refresh is the set this rank owns (gather of the global state). Message layout: int64 nRec, int64 nLig,
 * nRec x {int32 ref, ligRef, site(0..2,-1), cisRef; double pose[6]}, nLig x {int32 ref, recRef[3]; double pose[24]}; refs are
 * reference ids (1-based: receptors 1..NA, ligands NA+1..N), 0 = none; both lists sorted by ref. */
int kmc_strip_configure(kmc_handle *h, int32_t rank, int32_t nranks, double halo_width);
int kmc_strip_load_global(kmc_handle *h, int32_t n_rec, int32_t n_lig, const double *rec_pose, const double *lig_pose,
                          const int32_t *rec_lig, const int32_t *rec_site, const int32_t *rec_cis, int64_t step_done);
int kmc_strip_begin_refresh(kmc_handle *h);
int64_t kmc_strip_message(kmc_handle *h, int32_t side, const void **data);
int kmc_strip_rebuild(kmc_handle *h, const void *from_low, int64_t n_low, const void *from_high, int64_t n_high);
/* ---- the product path of the refresh: device resident, one exchange round, no host synchronisation (csrc/kmc_strips.cu) ----
 * Messages are fixed-capacity device buffers with the record counts in a 64-byte header; classify/pack kernels, the exchange and
 * the merge kernels are all enqueued on the handle's stream. Guards (reported by kmc_sync / kmc_step as KMC_ERR_CAPACITY): a unit
 * wider in x than halo_width - refresh_every * (reach per step), a band or the local capacity too small, a unit that arrives
 * incomplete. */
/* start state of a strip run generated on the GPU: every rank generates the same global configuration of n_rec + n_lig molecules (the
 * generator of kmc_init_random, a few ms for 1e7 molecules) and keeps what lies within reach of its strip; reference ids are the
 * global indices (receptor a -> a + 1, ligand b -> n_rec + b + 1) */
int kmc_strip_init_random(kmc_handle *h, int32_t n_rec, int32_t n_lig, uint64_t seed, int32_t sort_cells);
/* halo width that keeps the owned strip exact for refresh_every steps: refresh_every * (reach of information per step) + complex_extent */
double kmc_strip_halo_width(const kmc_params *p, int32_t refresh_every, double complex_extent);
/* NCCL communicator over the ranks of kmc_strip_configure: rank 0 fills id128 (128 bytes) with kmc_strip_unique_id, the caller
 * distributes it (MPI, torch.distributed, a file ...), every rank calls kmc_strip_comm_init (collective). refresh_every > 0 makes
 * kmc_step refresh the halos itself every refresh_every steps. All ranks must have been created with the same capacities. */
int kmc_strip_unique_id(void *id128);
int kmc_strip_comm_init(kmc_handle *h, const void *id128, int32_t refresh_every);
/* one refresh over NCCL (grouped ncclSend/ncclRecv with the two x-neighbours); asynchronous; collective */
int kmc_strip_refresh(kmc_handle *h);
/* the same refresh between n handles of ONE process (logical ranks 0..n-1 on one GPU; device-to-device copies instead of NCCL) */
int kmc_strip_refresh_local(kmc_handle **handles, int32_t n, int32_t refresh_every);
/* outputs of the WHOLE membrane (main.cpp:2247-2253, 2291-2305): every complex counted by the owner of its root ligand, every bond
 * by the owner of the receptor's unit. reduce != 0: ncclAllReduce over the communicator (collective, every rank gets the result);
 * reduce == 0: this rank's part. Values are those of the last refresh; a communicator-backed handle refreshes first if it has stepped since. */
int kmc_strip_get_series(kmc_handle *h, int32_t reduce, kmc_series *out);
int kmc_strip_get_oligomer_hist(kmc_handle *h, int32_t reduce, int64_t *hist, int32_t nbins);
/* per-rank state exchange with HOST buffers in the message record layout (64-byte receptor records, then 208-byte ligand records,
 * id-sorted). which = 2: the units this rank owns (union over ranks = the membrane); 3: everything it holds (owned + halo copies),
 * which is what kmc_strip_load_records takes to restore the rank. */
int kmc_strip_get_records(kmc_handle *h, int32_t which, void *host_buf, int64_t cap_bytes, int64_t *n_rec, int64_t *n_lig);
int kmc_strip_load_records(kmc_handle *h, const void *host_buf, int64_t n_rec, int64_t n_lig, int64_t step_done);

#ifdef __cplusplus
}
#endif
#endif
