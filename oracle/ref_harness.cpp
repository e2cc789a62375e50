// TEST INFRASTRUCTURE ONLY (oracle).  Nothing in the product path may link this.
//
// Drives the UNMODIFIED reference sources as a library.  The reference's main.cpp is a
// single translation unit whose physics lives in free functions; this file pulls it in
// from /root/reference at BUILD time (never copied into the repo), renames its main(),
// and exposes C entry points that call the reference's own functions in the order of
// its time loop (main.cpp:531-583).  Only the glue the reference keeps inline in
// main() is re-expressed here, each piece citing the lines it mirrors.
//
// Build: see oracle/Makefile (target _ref/libref_oracle.so); needs /root/reference.
//
// Reference UB that is made in-bounds here without changing any value that is used
// (SURVEY.md Appendix C): the nine VMR arrays are allocated with nlevel elements
// (main.cpp:454-460 writes index 20) and tau rows are copied into rows of 32 doubles
// (main.cpp:207-211 reads tau[20..29]).
#define main reference_main
#include "main.cpp"
#undef main

#include <cstdlib>
#include <cstring>

// defined (with external linkage) in the reference's repwvl_thermal.cpp:19-45
size_t LowerPos(std::vector<double>& tempsOnLayer, double& currT);

namespace {
const int NLAY = 20, NLEV = 21, NANG = 30, TAUPAD = 32;

void free_tau(double**& tau, double*& wvl, double*& weight, int nwvl) {
    // the reference pairs calloc with delete[] (main.cpp:551-554); free() is the correct pair
    if (tau) {
        for (int i = 0; i < nwvl; ++i) std::free(tau[i]);
        std::free(tau);
    }
    std::free(wvl);
    std::free(weight);
    tau = nullptr;
    wvl = nullptr;
    weight = nullptr;
}
}  // namespace

extern "C" {

// Constants of the committed reference build (main.cpp:66-92).
void ref_consts(double* out8) {
    out8[0] = Consts::tau_s;
    out8[1] = Consts::mu_s;
    out8[2] = Consts::g_asym;
    out8[3] = Consts::albedo;
    out8[4] = Consts::daytime;
    out8[5] = Consts::E_0;
    out8[6] = Consts::doublings;
    out8[7] = Consts::cloud_layer;
}

// Solar setup, main.cpp:514-515 -> out = {r_dir, s_dir, t_dir, r, t, r_total, solar_irr}
void ref_solar(double* out7) {
    double r_dir, s_dir, t_dir, r, t, r_total;
    doubling_adding(r_dir, s_dir, t_dir, r, t);
    double solar_irr = solar_radiative_transfer_setup(r_total, r_dir, s_dir, t_dir, r, t);
    out7[0] = r_dir; out7[1] = s_dir; out7[2] = t_dir; out7[3] = r; out7[4] = t;
    out7[5] = r_total; out7[6] = solar_irr;
}

// LowerPos of the reference (repwvl_thermal.cpp:19-45).
long ref_lowerpos(const double* a, int n, double x) {
    std::vector<double> v(a, a + n);
    return (long)LowerPos(v, x);
}

// Level -> layer initialisation, the inline part of main() (main.cpp:439-479).
// vmr_ppm_level: [ncol][5][21] in the order H2O, O3, CO2, CH4, N2O (data[4..8], main.cpp:426-430).
// vmr9_layer:    [ncol][9][20] in read_tau's argument order H2O, CO2, O3, N2O, CO, CH4, O2, HNO3, N2.
void ref_init_columns(int ncol, const double* plevel_hPa, const double* Tlevel, const double* vmr_ppm_level,
                      double co2_factor, double* Tlayer, double* vmr9_layer, double* rel_hum, double* player_out,
                      double* conv_out) {
    for (int c = 0; c < ncol; ++c) {
        vector<double> plevel(plevel_hPa, plevel_hPa + NLEV), Tlev(Tlevel + c * NLEV, Tlevel + (c + 1) * NLEV);
        vector<double> lev[5];
        for (int s = 0; s < 5; ++s)
            lev[s].assign(vmr_ppm_level + (c * 5 + s) * NLEV, vmr_ppm_level + (c * 5 + s + 1) * NLEV);
        double H2O[NLEV], O3[NLEV], CO2[NLEV], CH4[NLEV], N2O[NLEV];
        H2O[NLAY] = O3[NLAY] = CO2[NLAY] = CH4[NLAY] = N2O[NLAY] = 0.0;  // slot the reference writes out of bounds
        VMR_level_to_layer(lev[0], H2O);  // main.cpp:439-443
        VMR_level_to_layer(lev[1], O3);
        VMR_level_to_layer(lev[2], CO2);
        VMR_level_to_layer(lev[3], CH4);
        VMR_level_to_layer(lev[4], N2O);
        double factor = co2_factor;  // main.cpp:452
        for (int i = 0; i < NLEV; ++i) {  // main.cpp:454-457
            H2O[i] *= 1E-6; O3[i] *= 1E-6; CO2[i] *= factor * 1E-6; CH4[i] *= 1E-6; N2O[i] *= 1E-6;
        }
        for (int i = 0; i < NLAY; ++i) {  // main.cpp:467-468 (layer VMR x level p / e_sat(level T))
            double e_sat = magnus(Tlev[i]);
            rel_hum[c * NLAY + i] = H2O[i] * plevel[i] / e_sat;
        }
        for (int i = 0; i < NLAY; ++i) {  // main.cpp:472-474
            double player = (plevel[i] + plevel[i + 1]) / 2.0;
            Tlayer[c * NLAY + i] = (Tlev[i] + Tlev[i + 1]) / 2.0;
            if (c == 0) {
                player_out[i] = player;
                conv_out[i] = pow(1000.0 / player, Consts::kappa);
            }
        }
        double* v = vmr9_layer + (size_t)c * 9 * NLAY;
        for (int i = 0; i < NLAY; ++i) {
            v[0 * NLAY + i] = H2O[i]; v[1 * NLAY + i] = CO2[i]; v[2 * NLAY + i] = O3[i]; v[3 * NLAY + i] = N2O[i];
            v[4 * NLAY + i] = 0.0;    v[5 * NLAY + i] = CH4[i]; v[6 * NLAY + i] = 0.0;   v[7 * NLAY + i] = 0.0;
            v[8 * NLAY + i] = 0.0;
        }
    }
}

// read_tau of the reference, one column (repwvl_thermal.cpp:49-262), plus optional cloud (main.cpp:266-274).
// tau_out [nwvl][20] (may be NULL), wvl_out/weight_out [nwvl] (may be NULL).  Returns nwvl.
int ref_read_tau(const char* table, const double* plevel_hPa, const double* Tlayer, const double* vmr9, int cloud_on,
                 double* tau_out, double* wvl_out, double* weight_out) {
    vector<double> plevel(plevel_hPa, plevel_hPa + NLEV), T(Tlayer, Tlayer + NLAY);
    double sp[9][NLEV];
    for (int s = 0; s < 9; ++s) {
        std::memcpy(sp[s], vmr9 + s * NLAY, NLAY * sizeof(double));
        sp[s][NLAY] = 0.0;
    }
    int nwvl = 0;
    double *wvl = NULL, *weight = NULL, **tau = NULL;
    read_tau(table, NLEV, plevel, T, sp[0], sp[1], sp[2], sp[3], sp[4], sp[5], sp[6], sp[7], sp[8], &tau, &wvl,
             &weight, &nwvl, 0);
    if (cloud_on) cloud_into_tau(tau, nwvl);
    for (int i = 0; i < nwvl; ++i) {
        if (tau_out) std::memcpy(tau_out + (size_t)i * NLAY, tau[i], NLAY * sizeof(double));
        if (wvl_out) wvl_out[i] = wvl[i];
        if (weight_out) weight_out[i] = weight[i];
    }
    int n = nwvl;
    free_tau(tau, wvl, weight, nwvl);
    return n;
}

// radiative_transfer of the reference for a GIVEN tau (main.cpp:320-344); tau [nwvl][20].
void ref_radiative_transfer(int nwvl, const double* tau_in, const double* wvl_in, const double* weight_in,
                            const double* Tlayer_in, double T_surface, double solar_irr, double* E_down,
                            double* E_up, double* dE_out) {
    double dmu = 1.0 / (double)Consts::nangle;  // main.cpp:356
    vector<double> mu(NANG), B(NLAY, 0.0), alpha(NANG, 0.0), Ed(NLEV, 0.0), Eu(NLEV, 0.0), dE(NLAY, 0.0);
    for (int i = 0; i < NANG; ++i) mu[i] = dmu / 2.0 + dmu * (double)i;  // main.cpp:482
    vector<double> T(Tlayer_in, Tlayer_in + NLAY);
    vector<double> pad((size_t)nwvl * TAUPAD, 0.0), wv(wvl_in, wvl_in + nwvl), wt(weight_in, weight_in + nwvl);
    vector<double*> rows(nwvl);
    for (int i = 0; i < nwvl; ++i) {
        std::memcpy(&pad[(size_t)i * TAUPAD], tau_in + (size_t)i * NLAY, NLAY * sizeof(double));
        rows[i] = &pad[(size_t)i * TAUPAD];
    }
    radiative_transfer(B, alpha, Ed, Eu, dE, solar_irr, mu, dmu, T, T_surface, rows.data(), wt.data(), nwvl,
                       wv.data());
    std::memcpy(E_down, Ed.data(), NLEV * sizeof(double));
    std::memcpy(E_up, Eu.data(), NLEV * sizeof(double));
    std::memcpy(dE_out, dE.data(), NLAY * sizeof(double));
}

// The reference time loop (main.cpp:531-583) for `ncol` independent columns, iterations
// first_step .. first_step+nsteps-1 of the reference's loop counter i.  State in/out per column:
// Tlayer[20], T_surface, vmr9[9][20] (H2O row is rewritten by water_vapor_feedback), time_h (float).
// Outputs of the LAST iteration per column: E_down[21], E_up[21], dE[20], dt.  trace (may be NULL):
// [ncol][nsteps][24] = Tlayer[20] after the step, T_surface, dt, E_up[0], E_down[20].
int ref_advance(const char* table, int ncol, int first_step, int nsteps, const double* plevel_hPa,
                const double* rel_hum, double solar_irr, int cloud_on, double* Tlayer_io, double* Tsurf_io,
                double* vmr9_io, float* time_io, double* E_down_out, double* E_up_out, double* dE_out,
                double* dt_out, double* trace) {
    const double dp = 1000.0 / (double)Consts::nlayer;  // main.cpp:355
    const double dmu = 1.0 / (double)Consts::nangle;    // main.cpp:356
    vector<double> plevel(plevel_hPa, plevel_hPa + NLEV), player(NLAY), conv(NLAY), mu(NANG);
    for (int i = 0; i < NLAY; ++i) {
        player[i] = (plevel[i] + plevel[i + 1]) / 2.0;      // main.cpp:472
        conv[i] = pow(1000.0 / player[i], Consts::kappa);   // main.cpp:474
    }
    for (int i = 0; i < NANG; ++i) mu[i] = dmu / 2.0 + dmu * (double)i;  // main.cpp:482
    int nwvl_ret = 0;

    for (int c = 0; c < ncol; ++c) {
        vector<double> Tlayer(Tlayer_io + c * NLAY, Tlayer_io + (c + 1) * NLAY), theta(NLAY), B(NLAY, 0.0),
            alpha(NANG, 0.0), dE(NLAY, 0.0), E_down(NLEV, 0.0), E_up(NLEV, 0.0), e_sat(NLEV, 0.0),
            rh(rel_hum + c * NLAY, rel_hum + (c + 1) * NLAY);
        rh.push_back(0.0);
        double sp[9][NLEV];
        for (int s = 0; s < 9; ++s) {
            std::memcpy(sp[s], vmr9_io + ((size_t)c * 9 + s) * NLAY, NLAY * sizeof(double));
            sp[s][NLAY] = 0.0;
        }
        double T_surface = Tsurf_io[c], timestep = 0.0;
        float time = time_io ? time_io[c] : 0.0f;
        int nwvl = 0;
        double *wvl = NULL, *weight = NULL, **tau = NULL;

        if (first_step == 0) {  // main.cpp:500-504: tau from the initial, unsorted profile
            read_tau(table, NLEV, plevel, Tlayer, sp[0], sp[1], sp[2], sp[3], sp[4], sp[5], sp[6], sp[7], sp[8],
                     &tau, &wvl, &weight, &nwvl, 0);
            if (cloud_on) cloud_into_tau(tau, nwvl);
        }
        for (int k = 0; k < nsteps; ++k) {
            const int i = first_step + k;
            t_to_theta(Tlayer, theta, conv);                       // main.cpp:536
            sort(theta.begin(), theta.end(), greater<double>());   // main.cpp:539
            theta_to_t(theta, Tlayer, conv);                       // main.cpp:540
            if (i != 0) {                                          // main.cpp:549-572
                free_tau(tau, wvl, weight, nwvl);
                nwvl = 0;
                water_vapor_feedback(e_sat, Tlayer, rh, player, sp[0]);
                read_tau(table, NLEV, plevel, Tlayer, sp[0], sp[1], sp[2], sp[3], sp[4], sp[5], sp[6], sp[7],
                         sp[8], &tau, &wvl, &weight, &nwvl, 0);
                if (cloud_on) cloud_into_tau(tau, nwvl);
            }
            vector<double> pad((size_t)nwvl * TAUPAD, 0.0);
            vector<double*> rows(nwvl);
            for (int w = 0; w < nwvl; ++w) {
                std::memcpy(&pad[(size_t)w * TAUPAD], tau[w], NLAY * sizeof(double));
                rows[w] = &pad[(size_t)w * TAUPAD];
            }
            radiative_transfer(B, alpha, E_down, E_up, dE, solar_irr, mu, dmu, Tlayer, T_surface, rows.data(),
                               weight, nwvl, wvl);                 // main.cpp:574
            calculate_timestep(dp, dE, timestep);                  // main.cpp:576
            thermodynamics(Tlayer, dp, dE, timestep, T_surface, conv);  // main.cpp:578
            time += (float)timestep / 3600;                        // main.cpp:581
            if (trace) {
                double* t = trace + ((size_t)c * nsteps + k) * 24;
                std::memcpy(t, Tlayer.data(), NLAY * sizeof(double));
                t[20] = T_surface; t[21] = timestep; t[22] = E_up[0]; t[23] = E_down[NLEV - 1];
            }
        }
        nwvl_ret = nwvl;
        free_tau(tau, wvl, weight, nwvl);
        std::memcpy(Tlayer_io + c * NLAY, Tlayer.data(), NLAY * sizeof(double));
        Tsurf_io[c] = T_surface;
        for (int s = 0; s < 9; ++s) std::memcpy(vmr9_io + ((size_t)c * 9 + s) * NLAY, sp[s], NLAY * sizeof(double));
        if (time_io) time_io[c] = time;
        if (E_down_out) std::memcpy(E_down_out + c * NLEV, E_down.data(), NLEV * sizeof(double));
        if (E_up_out) std::memcpy(E_up_out + c * NLEV, E_up.data(), NLEV * sizeof(double));
        if (dE_out) std::memcpy(dE_out + c * NLAY, dE.data(), NLAY * sizeof(double));
        if (dt_out) dt_out[c] = timestep;
    }
    return nwvl_ret;
}

}  // extern "C"
