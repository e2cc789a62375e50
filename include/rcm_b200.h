/*
 * rcm_b200.h - C ABI of the B200-native radiative-convective column solver.
 *
 * Drop-in boundary for the thermal hot path of pabloconrat/our_first_climate_model
 * (reference loop body main.cpp:531-583).  The reference has no plugin/FFI layer: its
 * boundary is a handful of C++ free functions.  Every entry point below names the
 * reference interface it replaces.  Plain pointers and sizes only; all host buffers are
 * caller-owned; every function returns an rcm_status (never exits, never throws).
 * One rcm_solver per GPU; calls on one solver are not re-entrant.
 *
 * Array conventions (identical to the reference's): layers/levels are TOP-DOWN (index 0 =
 * top of atmosphere), pressures in hPa, volume mixing ratios as fractions, tau[iwvl][ilyr].
 * The nine species are in read_tau's argument order (repwvl_thermal.h:3-7):
 *   0 H2O, 1 CO2, 2 O3, 3 N2O, 4 CO, 5 CH4, 6 O2, 7 HNO3, 8 N2.
 */
#ifndef RCM_B200_H
#define RCM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RCM_NLAYER 20 /* Consts::nlayer, main.cpp:78 */
#define RCM_NLEVEL 21 /* Consts::nlevel, main.cpp:79 */
#define RCM_NSPECIES 9

typedef enum {
    RCM_OK = 0,
    RCM_ERR_ARG = 1,      /* bad argument (NULL, size, unsupported dimension) */
    RCM_ERR_STATE = 2,    /* call order (no table / no columns loaded yet) */
    RCM_ERR_CUDA = 3,     /* CUDA runtime error; see rcm_last_error() */
    RCM_ERR_IO = 4,       /* file not found / unreadable */
    RCM_ERR_FORMAT = 5,   /* file is not a supported table / matrix */
    RCM_ERR_NOMEM = 6,
    RCM_ERR_NO_DEVICE = 7 /* no CUDA device: the solver has no CPU fallback */
} rcm_status;

typedef struct rcm_solver rcm_solver; /* opaque, one per GPU */
typedef struct rcm_table rcm_table;   /* opaque host copy of a repwvl lookup table */

/* Model constants; rcm_default_params() fills in the reference's `Consts` (main.cpp:66-92). */
typedef struct {
    int nangle;        /* 30  Consts::nangle (1..64) */
    int cloud_layer;   /* 17  Consts::cloud_layer; < 0 = no grey cloud (cloud_into_tau, main.cpp:266-274) */
    double cloud_tau;  /* 1.0 Consts::tau_s / 2 */
    double dp;         /* 50  hPa, 1000/nlayer (main.cpp:355) */
    double max_dT;     /* 5   K, Consts::max_dT (calculate_timestep, main.cpp:156-162) */
    double dt_cap;     /* 43200 s (main.cpp:158-160) */
    double solar_irr;  /* W/m2, from rcm_solar_setup() (236.882897... for the committed Consts) */
    double dT_converged; /* K, stationarity threshold used for the `converged` count (not in the reference) */
    unsigned species_mask; /* bit s set: species s contributes to tau. Species outside the mask must have
                              VMR == 0 (they then add exactly +0.0 as in repwvl_thermal.cpp:244).
                              Default 0x2F = H2O, CO2, O3, N2O, CH4 - what main.cpp:452-460 feeds. */
} rcm_params;

typedef struct {
    double tau_s, mu_s, g_asym, albedo, daytime, E_0;
    int doublings;
} rcm_solar_params; /* Consts of main.cpp:86-91 */

/* Per-step ensemble diagnostics: the only quantities that ever cross GPUs (one allreduce). */
typedef struct {
    double toa_net_sum; /* sum over columns of solar_irr - E_up[0]  (W/m2) */
    double max_dT;      /* max over columns, layers of |T_sorted(n) - T_sorted(n-1)|  (K) */
    double n_converged; /* number of columns with that change < dT_converged */
    double max_abs_dE;  /* max over columns, layers of |dE| (W/m2) */
} rcm_step_scalars;

/* ---- host-only helpers (no GPU needed) ------------------------------------------------- */
const char* rcm_status_string(int status);
int rcm_default_params(rcm_params* p);
int rcm_default_solar_params(rcm_solar_params* sp);
/* doubling_adding + solar_radiative_transfer_setup (main.cpp:214-264).
 * out7 = r_dir, s_dir, t_dir, r, t, r_total (planetary albedo), solar_irr. */
int rcm_solar_setup(const rcm_solar_params* sp, double* out7);
/* LowerPos (repwvl_thermal.cpp:19-45): index of the table interval used for x. */
long rcm_lowerpos(const double* nodes, int n, double x);

/* Lookup-table loader.  Replaces the NetCDF part of read_tau (repwvl_thermal.cpp:113-176):
 * reads Reduced{10,20,100}Forcing.nc (NetCDF-4/HDF5 subset) or the flat
 * .rcmtab form of the same data. */
int rcm_table_load(const char* path, rcm_table** out);
void rcm_table_free(rcm_table* t);
/* dims4 = n_tpert (xsec_nbooks), n_species (npages), n_wvl (nrows), n_p (ncols) */
int rcm_table_dims(const rcm_table* t, int* dims4);
/* which: 0 xsec, 1 wvl (ChosenWvls), 2 weight (ChosenWeights), 3 p_grid, 4 t_ref, 5 t_pert, 6 vmrs_ref */
const double* rcm_table_array(const rcm_table* t, int which);

/* 21-level atmosphere file (test.atm / fpda.lbl.atm layout, main.cpp:396-430): skips 4 header
 * lines, reads up to 9 whitespace-separated columns per row.  cols_out [9][max_rows]
 * (column-major per variable: z, p, T, air, H2O, O3, CO2, CH4, N2O); *ncols_out = columns found. */
int rcm_read_atm(const char* path, int max_rows, double* cols_out, int* nrows_out, int* ncols_out);

/* Level -> layer initialisation, the inline part of main() (main.cpp:439-479).
 * Tlevel [ncol][21]; vmr_ppm_level [ncol][5][21] in file order H2O, O3, CO2, CH4, N2O.
 * Outputs: Tlayer [ncol][20], vmr9 [ncol][9][20], rel_hum [ncol][20], player[20], conv[20]. */
int rcm_init_columns(int ncol, const double* plevel_hPa, const double* Tlevel, const double* vmr_ppm_level,
                     double co2_factor, double* Tlayer, double* vmr9, double* rel_hum, double* player,
                     double* conv);

/* Synthetic perturbed-profile ensemble; deterministic in `seed`.
 * base_* are one 21-level column; outputs Tlevel [ncol][21], vmr_ppm_level [ncol][5][21]. */
int rcm_make_ensemble(int ncol, unsigned long long seed, const double* plevel_hPa, const double* base_Tlevel,
                      const double* base_vmr_ppm_level, double* Tlevel, double* vmr_ppm_level);

/* Line-by-line table reader: result-identical replacement of ASCII_file2xy2D
 * (lbl.arts/ascii.cpp:1631-1691, ascii.h:63) with flat, caller-freed output.
 * *x [nx], *y [nx*ny] are malloc'ed; free with rcm_free().  Returns 0 or the reference's
 * negative codes: -1 not found, -2 no memory, -5 not rectangular (ascii.h:34-38). */
int rcm_ascii_file2xy2D(const char* filename, int* nx, int* ny, double** x, double** y);
void rcm_free(void* p);
/* Band-integrated Planck radiance (cplkavg.cpp:124-243) on the host; *status as below. */
double rcm_cplkavg_host(double wvllo_nm, double wvlhi_nm, double t, int* status);

/* ---- solver lifecycle ------------------------------------------------------------------ */
int rcm_device_count(void);
int rcm_create(int device, const rcm_params* p, rcm_solver** out);
int rcm_destroy(rcm_solver* s);
const char* rcm_last_error(const rcm_solver* s);
int rcm_set_params(rcm_solver* s, const rcm_params* p);
/* Solver options.  RCM_OPT_ANGLE_CUBES (default 1): visit the quadrature angles in chains
 * mu, mu/3, mu/9 so that two thirds of the transmissions need an exp and the rest are cubes
 * of the previous ones (same mu values, different summation order, ~1e-15 relative). */
#define RCM_OPT_ANGLE_CUBES 0
/* RCM_OPT_ANGLE_PAIRS (default 1, needs the cubes): two chain heads whose node numbers satisfy 3a = 5b, 5a = 7b or
 * 3a = 7b take their transmissions as powers of ONE exp at a virtual node (30 angles: 4 of the 20 exp's per layer
 * and wavelength become 3-4 multiplications; same mu values, ~1e-14 relative). */
#define RCM_OPT_ANGLE_PAIRS 4
/* RCM_OPT_PATH (default 0): 0 = split path ((tile, wavelength-split) units, results independent of shard / GPU count),
 * 1 = the fused tile kernel of round 1 (comparison).  RCM_OPT_MULTI_STEP (default 1, split path): rcm_advance(n >= 2) runs
 * its n steps as ONE persistent launch with per-tile step flags instead of kernel boundaries where a step is only a few
 * rounds of work units per CTA (up to ~34,000 columns per GPU); 0 = always three launches per step, 2 = always one launch.
 * Bit-identical either way. */
#define RCM_OPT_PATH 5
#define RCM_OPT_MULTI_STEP 6
int rcm_set_option(rcm_solver* s, int option, int value);
/* cudaStream_t to launch on (NULL = the solver's own stream). */
int rcm_set_stream(rcm_solver* s, void* cuda_stream);
int rcm_synchronize(rcm_solver* s);

/* Upload a repwvl table (arrays as read by rcm_table_load; xsec in the file's
 * [n_tpert][n_species][n_wvl][n_p] order - the solver re-lays it out for the GPU). */
int rcm_set_repwvl_table(rcm_solver* s, const double* xsec, const double* wvl, const double* weight,
                         const double* p_grid, const double* t_ref, const double* t_pert, int n_tpert,
                         int n_species, int n_wvl, int n_p);
int rcm_set_repwvl_table_from(rcm_solver* s, const rcm_table* t);
/* Wavelengths [nm] and spectral weights alone (what radiative_transfer, main.cpp:320-344, is handed):
 * enough for rcm_radiative_transfer with a caller-supplied tau. */
int rcm_set_spectral_grid(rcm_solver* s, const double* wvl, const double* weight, int n_wvl);

/* Upload line-by-line tables (lbl.arts/README:5-16) and switch the solver to the LBL path:
 * wvl [nwvl] nm ascending, tau5 [5][nwvl][20] in the order H2O, CO2, O3, CH4, N2O (top-down layers, as
 * read by rcm_ascii_file2xy2D), h2o_ref / o3_ref [20] = layer VMRs the H2O / O3 tables were computed for
 * (o3_ref NULL: O3 table used unscaled), co2_factor as `factor` in main.cpp:452-456.
 * tau = tau_H2O*(H2O/h2o_ref) + co2_factor*tau_CO2 + tau_O3*(O3/o3_ref) + tau_CH4 + tau_N2O; the source is
 * cplkavg() over bins bounded by the midpoints between adjacent wavelengths. */
int rcm_set_lbl_tables(rcm_solver* s, const double* wvl, const double* tau5, int nwvl, const double* h2o_ref,
                       const double* o3_ref, double co2_factor);
/* Synthetic LBL tables in the reference's format (the real lbl.*.asc are not distributed):
 * wvl [nwvl] log-spaced 4-100 um, tau5 [5][nwvl][20]; deterministic in seed. */
int rcm_make_lbl_tables(int nwvl, unsigned long long seed, const double* plevel_hPa, const double* h2o_vmr_layer,
                        const double* o3_vmr_layer, double* wvl, double* tau5);
/* Write one species table as text: "wavelength[nm] dtau(20 layers, top-down)" per line (lbl.arts/README:5-11). */
int rcm_write_lbl_asc(const char* path, int nwvl, const double* wvl, const double* tau);

/* Column state.  plevel_hPa [21] is shared by the ensemble.  Tlayer [ncol][20], Tsurf [ncol],
 * vmr9 [ncol][9][20], rel_hum [ncol][20].  Resets the step counter to 0 (the first step then
 * uses tau of the initial, unsorted profile exactly as main.cpp:500-504 does). */
int rcm_set_columns(rcm_solver* s, int ncol, const double* plevel_hPa, const double* Tlayer,
                    const double* Tsurf, const double* vmr9, const double* rel_hum);
/* Compact per-step upload for resident ensembles: only T, Tsurf and the H2O..CH4 rows named in
 * species_mask (vmr_active [ncol][n_active][20], ascending species index).  Keeps plevel/rel_hum. */
/* Per-column solar forcing, computed on the device: doubling_adding + solar_radiative_transfer_setup
 * (main.cpp:214-264) with one cloud optical depth / zenith cosine / surface albedo per column.  tau_s, mu_s, albedo:
 * host arrays [ncol] or NULL (then the scalar in *sp).  From the next step on, column c is heated by its own
 * solar_irr[c] instead of rcm_params::solar_irr (main.cpp:341); with cloud_from_tau_s != 0 the thermal grey cloud of
 * column c becomes tau_s[c] / 2 instead of rcm_params::cloud_tau (main.cpp:266-274).  Optional outputs (host, [ncol]):
 * solar_irr_out, r_total_out (planetary albedo).  sp == NULL switches back to the ensemble-wide constants.
 * rcm_set_columns() drops the per-column forcing: call this after it. */
int rcm_set_column_solar(rcm_solver* s, const rcm_solar_params* sp, const double* tau_s, const double* mu_s,
                         const double* albedo, int cloud_from_tau_s, double* solar_irr_out, double* r_total_out);
int rcm_update_columns(rcm_solver* s, const double* Tlayer, const double* Tsurf, const double* vmr_active);
int rcm_set_step_index(rcm_solver* s, long step_index);

/* K1 only: optical depth build = compute part of read_tau (repwvl_thermal.cpp:197-248) +
 * cloud_into_tau (main.cpp:266-274).  tau_out [ncol][nwvl][20] host (NULL = skip the copy);
 * lowpos_p / lowpos_t [ncol][20] in the reference's bottom-up layer order (NULL = skip). */
int rcm_build_tau(rcm_solver* s, double* tau_out, int* lowpos_p, int* lowpos_t);

/* K2-K4 for a GIVEN tau = radiative_transfer (main.cpp:320-344).  tau [ncol][nwvl][20] host,
 * NULL = use the tau of the last rcm_build_tau.  Outputs E_down/E_up [ncol][21], dE [ncol][20]. */
int rcm_radiative_transfer(rcm_solver* s, const double* tau, double* E_down, double* E_up, double* dE);

/* The fused time step (K1-K5), nsteps iterations of main.cpp:531-583 for every column, state
 * resident on the GPU.  scalars_out [nsteps] host (NULL = skip).  Asynchronous pieces are
 * synchronised before return when scalars_out != NULL. */
int rcm_advance(rcm_solver* s, int nsteps, rcm_step_scalars* scalars_out);
/* Same, but leaves the per-step scalars on the device for an allreduce:
 * *d_scalars -> device double[nsteps][4] (layout of rcm_step_scalars), valid until the next call. */
int rcm_advance_async(rcm_solver* s, int nsteps, double** d_scalars);

/* The RCE driver loop: iterate main.cpp:531-583 until the ensemble is stationary.  Advances in blocks of
 * `check_every` fused steps (one launch each) and stops after the first block whose last step moved every
 * column's sorted temperature profile by less than params.dT_converged (n_converged == ncol), or after
 * max_steps.  *last = scalars of the last step done, *steps_done = iterations run by this call (either may be
 * NULL).  One GPU; for N ranks use the same loop with an allreduce of the scalars between the blocks
 * (our_first_climate_model_b200/distributed.py: run_to_equilibrium). */
int rcm_run_to_equilibrium(rcm_solver* s, long max_steps, int check_every, rcm_step_scalars* last, long* steps_done);

/* Download state / last-step fluxes; any pointer may be NULL.  Tlayer [ncol][20], Tsurf [ncol],
 * h2o [ncol][20], time_h [ncol] (float hours, main.cpp:581), E_down/E_up [ncol][21],
 * dE [ncol][20], dt [ncol]. */
int rcm_get_state(rcm_solver* s, double* Tlayer, double* Tsurf, double* h2o, float* time_h, double* E_down,
                  double* E_up, double* dE, double* dt);

/* One call = one reference loop iteration with HOST buffers in and out (what bench.py's e2e
 * leg times): rcm_update_columns + 1 fused step + download of E_down, E_up, dE, Tlayer, Tsurf. */
/* Checkpoint / restart: everything the step reads or carries (T, Tsurf, active VMRs, rel_hum, the
 * previous sorted profile, model time, last dt / fluxes / dE, per-column forcing, step index, pressure grid) in one
 * flat little-endian file.  A solver with the same table and species_mask that loads it continues bit-identically.
 * Tables are not stored.  Errors: RCM_ERR_IO, RCM_ERR_FORMAT (not a checkpoint / truncated), RCM_ERR_STATE (other
 * species_mask or spectral path). */
int rcm_save_checkpoint(rcm_solver* s, const char* path);
int rcm_load_checkpoint(rcm_solver* s, const char* path);
int rcm_column_count(const rcm_solver* s); /* columns currently loaded (rcm_set_columns / rcm_load_checkpoint) */
/* Batched form of output_conv (main.cpp:102-114): for every column the 20 rows "layer,player,Tlayer,theta,time"
 * ("%d,%f,%f,%f,%f\n", theta = Tlayer * (1000/player)^(2/7) as t_to_theta, main.cpp:123-127; time in hours).
 * column_ids != 0 prefixes every row with the column number ("%d,"); with ncol == 1 and column_ids == 0 the rows are
 * byte-identical to the reference's.  append != 0 appends like the reference's freopen(..., "a"); header != 0 writes
 * the reference's header line (main.cpp:528) first.  Host only. */
int rcm_write_profiles(const char* path, int append, int header, int ncol, const double* plevel_hPa,
                       const double* Tlayer, const float* time_h, int column_ids);
/* One iteration with HOST buffers (what a caller that keeps its state on the host does per step of main.cpp:531-583):
 * uploads Tlayer_in [ncol][20], Tsurf_in [ncol], vmr_active_in [ncol][nactive][20] - each only when not NULL (no VMR upload: the
 * rows of the previous call stay, H2O following the feedback on the device as in the reference loop), runs the step, downloads
 * E_down / E_up [ncol][21], dE / Tlayer_out [ncol][20], Tsurf_out [ncol] (each may be NULL = not wanted); synchronous.
 * On the split path the columns travel in chunks through upload / compute / download streams; with page-locked buffers the
 * pipeline of a steady-state call (second call on, same pointers, no VMR upload) is replayed as one CUDA graph - results are
 * bit-identical to rcm_advance(1) either way.  Environment: RCM_NO_GRAPH=1 keeps the direct calls, RCM_PIPE_CHUNKS=n sets the
 * chunk count.  Errors: RCM_ERR_ARG (no solver), RCM_ERR_STATE without columns / table, RCM_ERR_CUDA. */
int rcm_step_host(rcm_solver* s, const double* Tlayer_in, const double* Tsurf_in, const double* vmr_active_in,
                  double* E_down, double* E_up, double* dE, double* Tlayer_out, double* Tsurf_out);

/* Device-side band Planck (K2 of the LBL path), for parity against cplkavg(): n triples. */
int rcm_cplkavg_device(rcm_solver* s, int n, const double* lo_nm, const double* hi_nm, const double* t,
                       double* out);

/* Introspection for the bench: kernels launched so far, and FP64-pipe microbenchmarks
 * (result in 1e9 thread-instructions/s: which = 0 DFMA, 1 exp(), 2 divide, 3 solver exp). */
long rcm_launch_count(const rcm_solver* s);
/* rcm_step_host in steady state replays its chunk pipeline as one CUDA graph: how often it was captured / replayed. */
int rcm_host_graph_stats(const rcm_solver* s, long* captures, long* replays);
int rcm_fp64_microbench(rcm_solver* s, int which, double* ginstr_per_s);
/* Average device time (ms) of the fused step kernel over the launches since the last reset,
 * measured with CUDA events on the launching stream. */
int rcm_kernel_time_ms(rcm_solver* s, int reset, double* avg_ms, long* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* RCM_B200_H */
