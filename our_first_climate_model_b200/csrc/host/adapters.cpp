// Reference-signature adapters (include/rcm_b200_adapters.hpp) over the C ABI.
//   read_tau            <- repwvl_V2.01_cpp/repwvl_thermal.cpp:49-262
//   radiative_transfer  <- main.cpp:320-344
// One process-wide solver per adapter (single column, like the reference driver).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/rcm_b200_adapters.hpp"
#include "rcm_internal.h"

namespace {

struct TauCtx {
    rcm_solver* s = nullptr;
    rcm_table* tab = nullptr;
    std::string path;
};
TauCtx g_tau;
rcm_solver* g_rt = nullptr;
std::vector<double> g_rt_wvl, g_rt_weight;

bool trace() {
    static int on = -1;
    if (on < 0) on = std::getenv("RCM_ADAPTER_LOG") ? 1 : 0;
    return on == 1;
}

// The reference opens a literal path (main.cpp:500).  When that file is absent, RCM_TABLE_DIR may
// point at a directory holding the same table as <basename>.rcmtab or <basename>.nc.
int load_table(const char* path, rcm_table** out) {
    int st = rcm_table_load(path, out);
    if (st != RCM_ERR_IO) return st;
    const char* dir = std::getenv("RCM_TABLE_DIR");
    if (!dir) return st;
    std::string base(path);
    size_t slash = base.find_last_of('/');
    if (slash != std::string::npos) base = base.substr(slash + 1);
    size_t dot = base.find_last_of('.');
    std::string stem = dot == std::string::npos ? base : base.substr(0, dot);
    for (const char* ext : {".rcmtab", ".nc"}) {
        st = rcm_table_load((std::string(dir) + "/" + stem + ext).c_str(), out);
        if (st == RCM_OK) return st;
    }
    return st;
}

}  // namespace

void read_tau(const char* reducedLkpPath, int nLev, std::vector<double>& plevel, std::vector<double>& Tvector,
              double* H20_VMR, double* CO2_VMR, double* O3_VMR, double* N2O_VMR, double* CO_VMR, double* CH4_VMR,
              double* O2_VMR, double* HNO3_VMR, double* N2_VMR, double*** tau, double** wvl, double** weight,
              int* nWvl, int prop_at_Lev) {
    *nWvl = 0;
    *tau = nullptr;
    *wvl = *weight = nullptr;
    if (nLev != RCM_NLEVEL) {
        std::fprintf(stderr, "rcm read_tau: only %d levels are supported (got %d)\n", RCM_NLEVEL, nLev);
        return;
    }
    int st = RCM_OK;
    if (!g_tau.s) {
        rcm_params p;
        rcm_default_params(&p);
        p.cloud_layer = -1;  // the driver adds the cloud itself (cloud_into_tau, main.cpp:504/568)
        p.species_mask = 0x1FF;  // all nine species as passed in
        st = rcm_create(0, &p, &g_tau.s);
        if (st != RCM_OK) {
            std::fprintf(stderr, "rcm read_tau: %s\n", rcm_status_string(st));
            return;
        }
    }
    if (g_tau.path != reducedLkpPath) {
        rcm_table* t = nullptr;
        st = load_table(reducedLkpPath, &t);
        if (st == RCM_OK) st = rcm_set_repwvl_table_from(g_tau.s, t);
        if (st != RCM_OK) {
            std::fprintf(stderr, "rcm read_tau: cannot load %s: %s\n", reducedLkpPath, rcm_status_string(st));
            if (t) rcm_table_free(t);
            return;
        }
        if (g_tau.tab) rcm_table_free(g_tau.tab);
        g_tau.tab = t;
        g_tau.path = reducedLkpPath;
    }
    const int nl = RCM_NLAYER;
    const double* sp[RCM_NSPECIES] = {H20_VMR, CO2_VMR, O3_VMR, N2O_VMR, CO_VMR, CH4_VMR, O2_VMR, HNO3_VMR, N2_VMR};
    double T[RCM_NLAYER], vmr9[RCM_NSPECIES * RCM_NLAYER], zeros[RCM_NLAYER] = {0};
    for (int l = 0; l < nl; ++l) {
        // prop_at_Lev != 0: properties given at levels are averaged to layers (repwvl_thermal.cpp:219, :224)
        T[l] = prop_at_Lev ? (Tvector[l + 1] + Tvector[l]) / 2 : Tvector[l];
        for (int k = 0; k < RCM_NSPECIES; ++k)
            vmr9[k * nl + l] = prop_at_Lev ? (sp[k][l] + sp[k][l + 1]) / 2 : sp[k][l];
    }
    double Ts = T[nl - 1];
    st = rcm_set_columns(g_tau.s, 1, plevel.data(), T, &Ts, vmr9, zeros);
    int dims[4];
    rcm_table_dims(g_tau.tab, dims);
    const int nw = dims[2];
    std::vector<double> flat((size_t)nw * nl);
    if (st == RCM_OK) st = rcm_build_tau(g_tau.s, flat.data(), nullptr, nullptr);
    if (st != RCM_OK) {
        std::fprintf(stderr, "rcm read_tau: %s: %s\n", rcm_status_string(st), rcm_last_error(g_tau.s));
        return;
    }
    // same allocation pattern as the reference (repwvl_thermal.cpp:163-168): caller frees
    *nWvl = nw;
    *wvl = (double*)std::calloc(nw, sizeof(double));
    *weight = (double*)std::calloc(nw, sizeof(double));
    *tau = (double**)std::calloc(nw, sizeof(double*));
    std::memcpy(*wvl, rcm_table_array(g_tau.tab, 1), nw * sizeof(double));
    std::memcpy(*weight, rcm_table_array(g_tau.tab, 2), nw * sizeof(double));
    for (int i = 0; i < nw; ++i) {
        (*tau)[i] = (double*)std::calloc(nLev - 1, sizeof(double));
        std::memcpy((*tau)[i], &flat[(size_t)i * nl], nl * sizeof(double));
    }
    if (trace()) std::fprintf(stderr, "rcm adapter read_tau: nwvl=%d tau[0][19]=%.17g\n", nw, (*tau)[0][nl - 1]);
}

void radiative_transfer(std::vector<double>& B, std::vector<double>& alpha, std::vector<double>& E_down,
                        std::vector<double>& E_up, std::vector<double>& dE, const double solar_irr,
                        std::vector<double>& mu, const double& dmu, std::vector<double>& Tlayer,
                        const double& T_surface, double** tau, double* weight, int& nwvl, double* wvl) {
    (void)B;
    (void)alpha;
    const int nl = RCM_NLAYER, na = (int)mu.size();
    auto bail = [&](const char* why, int st) {
        std::fprintf(stderr, "rcm radiative_transfer: %s (%s)\n", why, rcm_status_string(st));
    };
    if ((int)Tlayer.size() != nl || (int)E_down.size() != nl + 1 || (int)E_up.size() != nl + 1 || (int)dE.size() != nl)
        return bail("vector sizes do not match 20 layers", RCM_ERR_ARG);
    for (int i = 0; i < na; ++i)  // the solver evaluates the reference's quadrature nodes (main.cpp:482)
        if (std::fabs(mu[i] - (dmu / 2.0 + dmu * (double)i)) > 1e-15 || std::fabs(dmu * na - 1.0) > 1e-12)
            return bail("mu is not the midpoint grid of main.cpp:482", RCM_ERR_ARG);
    rcm_params p;
    rcm_default_params(&p);
    p.cloud_layer = -1;  // tau arrives with the cloud already added by the driver
    p.nangle = na;
    p.solar_irr = solar_irr;
    int st = RCM_OK;
    if (!g_rt) st = rcm_create(0, &p, &g_rt); else st = rcm_set_params(g_rt, &p);
    if (st != RCM_OK) return bail("no solver", st);
    if ((int)g_rt_wvl.size() != nwvl || std::memcmp(g_rt_wvl.data(), wvl, nwvl * sizeof(double)) != 0 ||
        std::memcmp(g_rt_weight.data(), weight, nwvl * sizeof(double)) != 0) {
        st = rcm_set_spectral_grid(g_rt, wvl, weight, nwvl);
        if (st != RCM_OK) return bail("spectral grid", st);
        g_rt_wvl.assign(wvl, wvl + nwvl);
        g_rt_weight.assign(weight, weight + nwvl);
    }
    double plevel[RCM_NLEVEL], vmr9[RCM_NSPECIES * RCM_NLAYER] = {0}, zeros[RCM_NLAYER] = {0};
    for (int i = 0; i < RCM_NLEVEL; ++i) plevel[i] = 1000.0 * i / nl;  // not used by K2-K4
    double Ts = T_surface;
    st = rcm_set_columns(g_rt, 1, plevel, Tlayer.data(), &Ts, vmr9, zeros);
    std::vector<double> flat((size_t)nwvl * nl);
    for (int i = 0; i < nwvl; ++i) std::memcpy(&flat[(size_t)i * nl], tau[i], nl * sizeof(double));
    if (st == RCM_OK) st = rcm_radiative_transfer(g_rt, flat.data(), E_down.data(), E_up.data(), dE.data());
    if (st != RCM_OK) {
        std::fprintf(stderr, "rcm radiative_transfer: %s: %s\n", rcm_status_string(st), rcm_last_error(g_rt));
        return;
    }
    if (trace()) std::fprintf(stderr, "rcm adapter radiative_transfer: OLR=%.17g E_down_sfc=%.17g dE19=%.17g\n", E_up[0], E_down[nl], dE[nl - 1]);
}
