// Device side of the B200-native column solver: declarations shared with rcm_capi.cu.
#ifndef RCM_KERNELS_CUH
#define RCM_KERNELS_CUH

#include <cuda_runtime.h>

#include "rcm_internal.h"

constexpr int NLAY = RCM_NLAYER;
constexpr int NLEV = RCM_NLEVEL;
constexpr int MAX_ANGLE = 64;
constexpr int MAX_TPERT = 16;
constexpr int PLK_MAX = 128;       // wavelengths whose Planck factors the step kernel keeps in shared memory
constexpr int MAX_PAIR = 16;       // pair units of the angle schedule (DevConst)
constexpr int HALF = NLAY / 2;     // layers owned by each lane of a pair
constexpr int RCM_LBL_C = 16, RCM_LBL_NT = 128;  // tile shape of the LBL radiative-transfer kernel
// The solver's exp (exp_scaled): table 2^(j/EXP_TAB) in shared memory, EXP_REP copies of every entry side by side
// (lane l reads copy l & (EXP_REP-1)), Horner polynomial of degree EXP_DEG.  Two 8 KB configurations:
//   RCM_EXP_LOG2 = 10: 1024 entries, one copy, degree 2 (7 FP64 instructions per exp, approximation error 1.4e-16)
//   RCM_EXP_LOG2 = 7:   128 entries, eight copies, degree 3 (8 FP64 instructions, 7.6e-17)
#ifndef RCM_EXP_LOG2
#define RCM_EXP_LOG2 10
#endif
constexpr int EXP_LOG2 = RCM_EXP_LOG2, EXP_TAB = 1 << EXP_LOG2;
constexpr int EXP_REP = (EXP_LOG2 == 7) ? 8 : 1;
#ifdef RCM_EXP_DEG_OVERRIDE  // experiments only (tools/build_variants.sh): a shorter Horner polynomial = a less accurate exp
constexpr int EXP_DEG = RCM_EXP_DEG_OVERRIDE;
#else
constexpr int EXP_DEG = (EXP_LOG2 == 7) ? 3 : 2;
#endif
static_assert(EXP_LOG2 == 7 || EXP_LOG2 == 10, "exp table configurations: 128 x 8 copies or 1024 x 1");
constexpr double EXP_L2E = EXP_TAB * 1.4426950408889634;  // EXP_TAB / ln 2: argument scaling of exp_scaled

// Everything that is uniform over the ensemble.  Lives in __constant__ memory.
struct DevConst {
    int nwvl, nangle, nactive, n_tpert, n_species, n_p;
    int species[RCM_NSPECIES];  // active species, ascending
    int cloud_layer;
    double cloud_tau, dp, max_dT, dt_cap, solar_irr, dT_converged;
    // per layer, from the shared pressure grid - host computed with the reference's expressions
    // (repwvl_thermal.cpp:202, :214, :226, :240).  ip/player/conv are indexed by the top-down layer l;
    // ipcell/delP/numDens/tref_ip by the pair-order row (l for l<10, 29-l otherwise).
    int ip[NLAY];
    int ipcell[NLAY];  // row * (n_tpert-1): first cell of the layer's block in the per-layer coefficient table
    int cloud_row;     // pair-order row of the cloud layer, -1 if none
    double cloud_w[NLAY];  // 1.0 on that row, 0.0 elsewhere: tau += cloud_w[r] * cloud_tau as one FMA (exact either way)
    double delP[NLAY], numDens[NLAY], tref_ip[NLAY], player[NLAY], conv[NLAY];
    double t_pert[MAX_TPERT];
    // Angle schedule.  The quadrature nodes are visited in chains mu, mu/3, mu/9, ...: the head of a chain
    // is evaluated with exp, every further level is the cube of the one before (1/mu triples).  Slots are
    // numbered chain by chain.  nchain is even (a zero-weight exp(0) chain pads an odd count); entry
    // [nchain] of neg_inv_mu_l2e is 0 (the pipelined loop evaluates it and never uses the result).
    int nslot, nchain;
    int chain_len[MAX_ANGLE + 2];          // slots of chain i (>= 1)
    double neg_inv_mu_l2e[MAX_ANGLE + 2];  // per CHAIN: -1/mu of its head, times EXP_TAB/ln2 (argument scaling of exp_scaled)
    double cmu[MAX_ANGLE + 2];             // per SLOT: 2*pi*mu*dmu (0 for a padding slot)
    double csum;                           // sum of cmu over all nodes
    // Pair units (visited before the chains; their slots come first in cmu).  Two chain heads a > b whose node numbers
    // n = 2i+1 satisfy pa * a = pb * b = R with (pa, pb) = (3,5), (5,7) or (3,7) share ONE exp: x = t(R) at a virtual
    // node, t(a) = x^pa, t(b) = x^pb by 3-4 multiplications.  Every unit's root is evaluated during the last sweep of
    // the unit before it; pair_nim[p] is the root of unit p, pair_nim[npair] = 0.
    int npair;
    int pair_type[MAX_PAIR];               // 0: x^3, x^5   1: x^5, x^7   2: x^3, x^7
    int pair_lenA[MAX_PAIR], pair_lenB[MAX_PAIR];  // slots of the cube chains below the two heads
    double pair_nim[MAX_PAIR + 1];         // -1/mu of the virtual root, times EXP_TAB/ln2
    double expc[4];                        // Horner coefficients of exp_scaled, lowest order first
    // LBL band edges etc. live in global memory
};

// Per-launch arguments (pointers into the solver's device allocations).
struct StepArgs {
    int ncol;            // columns in this launch (all pointers below are offset to its first column)
    int diag_ncol;       // columns of the whole ensemble = row length of diag
    int C;               // columns per tile
    int nthreads;        // threads per CTA (2 * C * wavelength groups)
    int stage_rows;      // 1: the next wavelength's table rows travel into shared memory (cp.async) during the angle loop
    int clampk;          // 1: exp_scaled clamps its exponent itself (angle schedules where tau_clamp would bite)
    double tau_clamp;    // tau is clamped to this before the transmissions are evaluated (see exp_scaled)
    double T_floor;      // the Planck source is evaluated at max(T, T_floor): keeps exp_scaled's exponent in range (rcm_set_spectral_grid)
    int ntiles;
    int nsteps;          // time steps fused in this launch
    long step_index;     // global index of the first step (0 => initial-profile tau, main.cpp:500-504)
    // table as bilinear coefficients coef[cell][wvl][active species][4]
    const double* __restrict__ coef;
    const double* __restrict__ planck_c;  // [nwvl] h*c/(lambda*kB)   [K]
    const double* __restrict__ planck_k;  // [nwvl] weight*2*h*c^2/lambda^5/1e9
    // column state
    double* Tlayer;        // [ncol][20]
    double* Tsurf;         // [ncol]
    double* vmr;           // [ncol][nactive][20]
    const double* rel_hum; // [ncol][20]
    double* Tprev;         // [ncol][20] sorted profile of the previous step
    float* time_h;         // [ncol]
    // outputs of the last step
    double* E_down;        // [ncol][21]
    double* E_up;          // [ncol][21]
    double* dE;            // [ncol][20]
    double* dt;            // [ncol]
    double* diag;          // [nsteps][ncol][4]  toa_net, dT_stat, max|dE|, spare   (NULL = skip)
    // component paths
    double* tau_io;        // [ncol][nwvl][20]  (written by MODE_TAU, read by MODE_RT)
    int* lowpos_t;         // [ncol][20] bottom-up (MODE_TAU)
    const double* exp_tab; // [EXP_TAB] 2^(j/EXP_TAB), high words prepared for exp_scaled (rcm_create)
    int h2o_slot;          // position of H2O in the active list, -1 if absent
    // per-column solar forcing / grey-cloud optical depth (rcm_set_column_solar); NULL = the ensemble-wide constants
    const double* solar_col;  // [ncol] absorbed solar irradiance, W/m2 (main.cpp:255-264 per column)
    const double* cloud_col;  // [ncol] tau added to the cloud layer (main.cpp:266-274 per column)
};

// Line-by-line path: arguments of the three per-step kernels.
struct LblArgs {
    int ncol, ntiles, nwvl, nchunks, chunk_len, nact, h2o_slot, o3_slot;
    int clampk;          // as StepArgs
    long step_index;
    double co2_factor, tau_clamp;
    const double* __restrict__ wvl_lo;   // [nwvl] bin edges, nm
    const double* __restrict__ wvl_hi;
    const double* __restrict__ wn_hi;    // [nwvl] 1e7 / wvl_lo, 1e7 / wvl_hi (wavenumbers of the bin edges, cm-1)
    const double* __restrict__ wn_lo;
    const int* __restrict__ bin_ok;      // [nwvl] the bin takes cplkavg's narrow-band (Simpson) branch
    const double* __restrict__ tau3;     // [3][nwvl][20]  H2O, O3, and f_CO2 * CO2 + CH4 + N2O (the column-independent part)
    const double* __restrict__ h2o_ref;  // [20]
    const double* __restrict__ o3_ref;   // [20] or NULL
    const double* __restrict__ exp_tab;
    double* Tlayer; double* Tsurf; double* vmr; const double* rel_hum; double* Tprev; float* time_h;
    double* sH; double* sO; double* dTstat;  // [ncol][20], [ncol][20], [ncol]
    double* part;                            // [nchunks][ncol][42]
    double* E_down; double* E_up; double* dE; double* dt; double* diag;
    const double* solar_col; const double* cloud_col;  // as StepArgs
};

// Split path (rcm_split_kernels.cuh): the repwvl step as a per-tile K5 kernel and a (tile, wavelength split) K1-K4 kernel.
struct SplitArgs {
    int ncol;            // columns in this launch (pointers below are offset to its first column / tile)
    int diag_ncol;       // columns of the whole ensemble = row length of diag
    int ntiles, nsplit, ipu, nitem, nunits;  // tiles of 16 columns; splits per tile; wavelength rounds per split / per tile
    int stage_rows, clampk, h2o_slot;
    double tau_clamp, T_floor;  // as StepArgs
    const double* __restrict__ coef;      // the split path's rows: [cell][wvl][16] = {c0 + cP*delP, cT, cPT} x 5 species + pad
    const double* __restrict__ planck_c;
    const double* __restrict__ planck_k;
    const double* __restrict__ exp_tab;
    double* Tlayer; double* Tsurf; double* vmr; const double* rel_hum; double* Tprev; float* time_h;
    double* E_down; double* E_up; double* dE; double* dt; double* diag;
    const double* solar_col; const double* cloud_col;
    unsigned char* tile;  // [ntiles][TILE_BYTES] everything the unit kernel needs for a tile, written by the K5 kernel
    double* part;         // [ntiles][nsplit][42][16] partial fluxes of the units
    double* dTstat;       // [ncol] stationarity diagnostic of the step being computed
    unsigned* counter;    // work counter of the unit kernel (reset by the K5 kernel in front of it; NULL there: leave it alone)
    int quota;            // units a CTA of the unit kernel takes before it exits (persistent: a huge number)
};
struct SplitColFlags {
    int finish;         // sum the partial fluxes of a step, dE, time step, T update
    int prep;           // then prepare the next step (sort, feedback, indices, tile block)
    int first;          // the step being prepared is iteration 0: tau from the initial, unsorted profile (main.cpp:500-504)
    int write_all_vmr;  // (re)write every species into the tile block, not just H2O
    int write_out;      // store E_down, E_up, dE, dt of the finished step
    int diag_step;      // row of diag the finished step writes
};
struct SplitMultiArgs {  // the multi-step unit kernel (rcm_split_multi_kernel)
    int nsteps;          // steps of this launch; diag rows 0 .. nsteps-1
    unsigned* done;      // [ntiles] units finished per tile, all steps of the launch (zero at launch)
    unsigned* ready;     // [ntiles] step the tile is prepared for (zero at launch: the K5 prep in front has run)
};
size_t rcm_split_tile_bytes();
size_t rcm_split_part_doubles();
int rcm_split_ipu();
cudaError_t rcm_launch_split_col(const SplitArgs& a, const SplitColFlags& f, cudaStream_t st);
cudaError_t rcm_launch_split_rt(const SplitArgs& a, int grid, cudaStream_t st);
cudaError_t rcm_launch_split_multi(const SplitArgs& a, const SplitMultiArgs& m, int grid, cudaStream_t st);

enum { MODE_STEP = 0, MODE_TAU = 1, MODE_RT = 2 };

size_t rcm_step_smem_bytes(int C, int nactive, int nthreads);
cudaError_t rcm_upload_const(const DevConst& c);
cudaError_t rcm_launch_step(int mode, const StepArgs& a, int nactive, int grid, cudaStream_t st);
size_t rcm_reduce_scratch_doubles(int nsteps);
cudaError_t rcm_launch_reduce_diag(const double* diag, int nsteps, int ncol, double dT_converged, double* scratch,
                                   unsigned* ticket, double* scalars, cudaStream_t st);
cudaError_t rcm_launch_coef(const double* xsec_file, double* coef, int nt, int ns, int nw, int np, int nact,
                            const int* d_species, cudaStream_t st);
// coef4: rcm_launch_coef's table for five active species, nrows = 20 * (n_tpert - 1) * nwvl rows -> coef3 [nrows][16]
cudaError_t rcm_launch_coef3(const double* coef4, double* coef3, size_t nrows, cudaStream_t st);
cudaError_t rcm_launch_microbench(int which, double* out, const double* tab, long iters, int grid,
                                  cudaStream_t st);
size_t rcm_lbl_smem_bytes(int C, int nthreads);
// ev0 / ev1 (may be NULL): recorded around the radiative-transfer kernel, the dominant one of the three
cudaError_t rcm_launch_lbl_step(const LblArgs& a, cudaStream_t st, cudaEvent_t ev0, cudaEvent_t ev1);
// doubling_adding + solar_radiative_transfer_setup (main.cpp:214-264) for n columns, one thread each.  tau_s / mu_s /
// albedo: per-column arrays or NULL (then the scalar in sp).  Outputs (any may be NULL): solar_irr, r_total, and
// cloud_tau = tau_s / 2 (the thermal grey-cloud term of the same cloud, main.cpp:267).
struct SolarArgs {
    int n, doublings;
    double tau_s, mu_s, g_asym, albedo, daytime, E_0;
    const double* tau_s_col; const double* mu_s_col; const double* albedo_col;
    double* solar_irr; double* r_total; double* cloud_tau;
};
cudaError_t rcm_launch_solar(const SolarArgs& a, cudaStream_t st);
cudaError_t rcm_launch_cplkavg(int n, const double* lo, const double* hi, const double* t, double* out,
                               const double* exp_tab, int narrow, cudaStream_t st);

#endif
