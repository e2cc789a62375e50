/* TEST INFRASTRUCTURE ONLY (oracle).  Plain-C restatement of the reference's thermal hot
 * path; see rcm_oracle.c.  Nothing in the product path may include or link this. */
#ifndef RCM_ORACLE_H
#define RCM_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int n_tpert, n_species, n_wvl, n_p;
    const double* xsec;   /* [n_tpert][n_species][n_wvl][n_p] */
    const double* wvl;    /* [n_wvl] nm */
    const double* weight; /* [n_wvl] */
    const double* p_grid; /* [n_p] Pa, descending */
    const double* t_ref;  /* [n_p] K */
    const double* t_pert; /* [n_tpert] K */
} rcmo_table;

typedef struct {
    double tau_s, mu_s, g_asym, albedo, daytime, E_0;
    int doublings;
} rcmo_solar_params;

typedef struct {
    int nlayer, nangle;    /* 20, 30 */
    int cloud_layer;       /* 17; < 0 disables the grey cloud */
    double cloud_tau;      /* tau_s / 2 */
    double dp;             /* 1000 / nlayer  (hPa) */
    double max_dT;         /* 5 (float in the reference) */
    double dt_cap;         /* 3600 * 12 s */
    double solar_irr;      /* W/m2 */
} rcmo_params;

long rcmo_lowerpos(const double* a, int n, double x);
void rcmo_read_tau(const rcmo_table* t, int nlev, const double* plevel_hPa, const double* Tlayer,
                   const double* vmr9, double* tau, long* lowpos_p, long* lowpos_t);
void rcmo_cloud_into_tau(double* tau, int nwvl, int nlayer, int cloud_layer, double cloud_tau);
double rcmo_planck(double wvl_nm, double weight, double T);
void rcmo_radiative_transfer(const rcmo_params* p, int nwvl, const double* tau, const double* wvl,
                             const double* weight, const double* Tlayer, double T_surface, double* E_down,
                             double* E_up, double* dE);
double rcmo_timestep(const rcmo_params* p, const double* dE);
void rcmo_thermodynamics(const rcmo_params* p, double* Tlayer, const double* dE, double timestep,
                         double* T_surface, const double* conv);
void rcmo_theta_sort(int nlayer, double* Tlayer, const double* conv);
double rcmo_magnus(double T);
void rcmo_water_vapor_feedback(int nlayer, const double* Tlayer, const double* rel_hum, const double* player,
                               double* h2o_vmr);
void rcmo_solar_setup(const rcmo_solar_params* sp, double* out7);
void rcmo_init_columns(int ncol, int nlayer, const double* plevel_hPa, const double* Tlevel,
                       const double* vmr_ppm_level, double co2_factor, double* Tlayer, double* vmr9_layer,
                       double* rel_hum, double* player_out, double* conv_out);
int rcmo_advance(const rcmo_table* t, const rcmo_params* p, int ncol, int first_step, int nsteps,
                 const double* plevel_hPa, const double* rel_hum, double* Tlayer_io, double* Tsurf_io,
                 double* vmr9_io, float* time_io, double* E_down_out, double* E_up_out, double* dE_out,
                 double* dt_out, double* trace);
double rcmo_cplkavg(double wvllo, double wvlhi, double t, int* status);

/* line-by-line step (builder-defined composition; see rcm_oracle.c) */
void rcmo_lbl_bin_edges(int nwvl, const double* wvl, double* lo, double* hi);
void rcmo_lbl_tau(int nwvl, int nlayer, const double* tau_h2o, const double* tau_co2, const double* tau_o3,
                  const double* tau_ch4, const double* tau_n2o, const double* h2o_scale, double co2_factor,
                  const double* o3_scale, double* tau);
void rcmo_lbl_radiative_transfer(const rcmo_params* p, int nwvl, const double* tau, const double* wvl_lo,
                                 const double* wvl_hi, const double* Tlayer, double T_surface, double* E_down,
                                 double* E_up, double* dE);
int rcmo_lbl_advance(const rcmo_params* p, int nwvl, const double* wvl, const double* tau5, int ncol,
                     int first_step, int nsteps, const double* plevel_hPa, const double* rel_hum,
                     const double* h2o_ref, const double* o3_scale, double co2_factor, double* Tlayer_io,
                     double* Tsurf_io, double* h2o_io, double* E_down_out, double* E_up_out, double* dE_out,
                     double* dt_out);
#ifdef __cplusplus
}
#endif
#endif
