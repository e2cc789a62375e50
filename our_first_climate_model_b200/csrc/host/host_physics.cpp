// Host-side, one-time pieces of the reference driver that stay C++ (include/rcm_b200.h):
// constants, solar setup, level->layer initialisation, LowerPos, the synthetic ensemble
// generator and the band-integrated Planck function used to check the device version.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "rcm_internal.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

long rcm_lowerpos_impl(const double* a, int n, double x) {
    // reference repwvl_thermal.cpp:19-45: walk the nodes until the sign of (node - x) changes;
    // no change (x outside the grid on either side, or n < 2) ends in the last interval.
    auto sgn = [](double v) { return (0.0 < v) - (v < 0.0); };
    int prev = sgn(a[0] - x);
    for (int k = 1; k < n; ++k) {
        int cur = sgn(a[k] - x);
        if (cur != prev || k == n - 1) return k - 1;
        prev = cur;
    }
    return 0;
}

extern "C" {

const char* rcm_status_string(int status) {
    switch (status) {
        case RCM_OK: return "ok";
        case RCM_ERR_ARG: return "bad argument";
        case RCM_ERR_STATE: return "call out of order (table or columns not loaded)";
        case RCM_ERR_CUDA: return "CUDA error";
        case RCM_ERR_IO: return "file not found or unreadable";
        case RCM_ERR_FORMAT: return "unsupported file format";
        case RCM_ERR_NOMEM: return "out of memory";
        case RCM_ERR_NO_DEVICE: return "no CUDA device (this solver has no CPU fallback)";
        default: return "unknown status";
    }
}

int rcm_default_params(rcm_params* p) {
    if (!p) return RCM_ERR_ARG;
    p->nangle = 30;            // main.cpp:80
    p->cloud_layer = 17;       // main.cpp:92
    p->cloud_tau = 2.0 / 2.0;  // main.cpp:86, :267
    p->dp = 1000.0 / 20;       // main.cpp:355
    p->max_dT = 5;             // main.cpp:82
    p->dt_cap = 3600 * 12;     // main.cpp:158
    rcm_solar_params sp;
    rcm_default_solar_params(&sp);
    double o[7];
    rcm_solar_setup(&sp, o);
    p->solar_irr = o[6];
    p->dT_converged = 1e-5;
    p->species_mask = 0x2F;
    return RCM_OK;
}

int rcm_default_solar_params(rcm_solar_params* sp) {
    if (!sp) return RCM_ERR_ARG;
    sp->tau_s = 2.0;                           // main.cpp:86
    sp->mu_s = std::cos(60 * M_PI / 180.0);    // main.cpp:87
    sp->g_asym = 0.85;                         // main.cpp:89
    sp->albedo = 0.12;                         // main.cpp:90
    sp->daytime = 0.5;                         // main.cpp:91
    sp->E_0 = 1361.0;                          // main.cpp:69
    sp->doublings = 20;                        // main.cpp:88
    return RCM_OK;
}

int rcm_solar_setup(const rcm_solar_params* sp, double* out7) {
    if (!sp || !out7 || sp->doublings < 0 || sp->doublings > 60) return RCM_ERR_ARG;
    // doubling_adding, main.cpp:214-253: start from an optically thin layer of the
    // delta-scaled cloud and double it `doublings` times.
    const double tau = (1 - sp->g_asym) * sp->tau_s;
    const double dtau = tau / std::pow(2, sp->doublings);
    const double thin = dtau / sp->mu_s;
    double r = 0.5 * thin, t = 1.0 - r;
    double r_dir = thin * 0.5, s_dir = r_dir, t_dir = 1 - thin;
    for (int i = 0; i < sp->doublings; ++i) {
        const double denom = 1 - r * r;
        const double r2 = r + (r * t * t) / denom;
        const double t2 = (t * t) / denom;
        const double s2 = (t * s_dir + t_dir * r_dir * r * t) / denom + t_dir * s_dir;
        const double rd2 = (t * s_dir * r + t * t_dir * r) / denom + r_dir;
        t_dir = std::pow(t_dir, 2);
        s_dir = s2;
        r_dir = rd2;
        r = r2;
        t = t2;
    }
    // solar_radiative_transfer_setup, main.cpp:255-264
    const double r_total = r_dir + (t_dir + s_dir) / (1 - sp->albedo * r) * t * sp->albedo;
    out7[0] = r_dir;
    out7[1] = s_dir;
    out7[2] = t_dir;
    out7[3] = r;
    out7[4] = t;
    out7[5] = r_total;
    out7[6] = (float)sp->daytime * sp->E_0 * sp->mu_s * (1 - r_total);
    return RCM_OK;
}

long rcm_lowerpos(const double* nodes, int n, double x) {
    if (!nodes || n < 1) return -1;
    return rcm_lowerpos_impl(nodes, n, x);
}

static inline double magnus_hPa(double T) {  // main.cpp:277-279
    return 6.1094 * std::exp(17.625 * (T - 273.15) / (T - 273.15 + 243.04));
}

int rcm_init_columns(int ncol, const double* plevel, const double* Tlevel, const double* vmr_ppm_level,
                     double co2_factor, double* Tlayer, double* vmr9, double* rel_hum, double* player,
                     double* conv) {
    if (ncol < 0 || !plevel || !Tlevel || !vmr_ppm_level || !Tlayer || !vmr9 || !rel_hum) return RCM_ERR_ARG;
    const int NL = RCM_NLAYER, NV = RCM_NLEVEL;
    for (int i = 0; i < NL; ++i) {
        const double pm = (plevel[i] + plevel[i + 1]) / 2.0;     // main.cpp:472
        if (player) player[i] = pm;
        if (conv) conv[i] = std::pow(1000.0 / pm, 2.0 / 7.0);    // main.cpp:66, :474
    }
    // file order H2O, O3, CO2, CH4, N2O (main.cpp:426-430) -> read_tau slots (repwvl_thermal.h:3-7)
    static const int slot[5] = {0, 2, 1, 5, 3};
    for (long c = 0; c < ncol; ++c) {
        const double* T = Tlevel + c * NV;
        const double* lev = vmr_ppm_level + c * 5 * NV;
        double* v = vmr9 + c * RCM_NSPECIES * NL;
        std::memset(v, 0, sizeof(double) * RCM_NSPECIES * NL);  // CO, O2, HNO3, N2 (main.cpp:459)
        for (int s = 0; s < 5; ++s) {
            const double unit = (s == 2) ? co2_factor * 1E-6 : 1E-6;  // main.cpp:456
            for (int i = 0; i < NL; ++i) {
                double m = (lev[s * NV + i] + lev[s * NV + i + 1]) / 2.0;  // main.cpp:144
                m *= unit;
                v[slot[s] * NL + i] = m;
            }
        }
        for (int i = 0; i < NL; ++i) {
            // main.cpp:467-468 mixes the LAYER vmr with the LEVEL pressure and temperature (quirk C2)
            rel_hum[c * NL + i] = v[i] * plevel[i] / magnus_hPa(T[i]);
            Tlayer[c * NL + i] = (T[i] + T[i + 1]) / 2.0;  // main.cpp:473
        }
    }
    return RCM_OK;
}

// ---- synthetic ensemble ------------------------------------------------------------------
// SplitMix64 stream + Box-Muller: reproducible everywhere (no std:: distribution objects,
// whose output is implementation-defined).
namespace {
struct Rng {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
    double uni(double a, double b) { return a + (b - a) * uni(); }
    double normal() {
        double u1 = 1.0 - uni(), u2 = uni();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(2.0 * M_PI * u2);
    }
};
}  // namespace

int rcm_make_ensemble(int ncol, unsigned long long seed, const double* plevel, const double* baseT,
                      const double* base_vmr, double* Tlevel, double* vmr_ppm_level) {
    if (ncol < 0 || !plevel || !baseT || !base_vmr || !Tlevel || !vmr_ppm_level) return RCM_ERR_ARG;
    const int NV = RCM_NLEVEL;
    for (long c = 0; c < ncol; ++c) {
        Rng g{seed * 0x100000001B3ull + (uint64_t)c * 0xD1342543DE82EF95ull + 1};
        // member 0 is the unperturbed base column
        const double a = c ? g.uni(-8, 8) : 0, b = c ? g.uni(-6, 6) : 0;
        double amp[3], ph[3];
        for (int m = 0; m < 3; ++m) {
            amp[m] = c ? g.normal() / std::sqrt(3.0) : 0;
            ph[m] = g.uni(0, 2 * M_PI);
        }
        const double fh2o = c ? std::exp(0.4 * g.normal()) : 1, fo3 = c ? g.uni(0.8, 1.2) : 1;
        for (int i = 0; i < NV; ++i) {
            const double x = plevel[i] / 1000.0;
            double noise = 0;
            for (int m = 0; m < 3; ++m) noise += amp[m] * std::sin((m + 1) * M_PI * x + ph[m]);
            Tlevel[c * NV + i] = baseT[i] + a + b * x + noise;
            double* v = vmr_ppm_level + c * 5 * NV;
            v[0 * NV + i] = base_vmr[0 * NV + i] * fh2o;
            v[1 * NV + i] = base_vmr[1 * NV + i] * fo3;
            v[2 * NV + i] = base_vmr[2 * NV + i];
            v[3 * NV + i] = base_vmr[3 * NV + i];
            v[4 * NV + i] = base_vmr[4 * NV + i];
        }
    }
    return RCM_OK;
}

// ---- band-integrated Planck on the host (checks the device version; cplkavg.cpp:124-243) ----
double rcm_cplkavg_host(double wvllo, double wvlhi, double t, int* status) {
    const double c2 = 1.438786, sigma = 5.67032E-8, vcut = 1.5;               // cplkavg.cpp:120-122
    const double a1 = 1. / 3., a2 = -1. / 8., a3 = 1. / 60., a4 = -1. / 5040., a5 = 1. / 272160.,
                 a6 = -1. / 13305600.;                                         // :114-119
    static const double vcp[7] = {10.25, 5.7, 3.9, 2.9, 2.3, 1.9, 0.0};        // :128
    const double vmax = std::log(DBL_MAX), sigdpi = sigma / M_PI, conc = 15. / std::pow(M_PI, 4.);
    auto plkf = [](double x) { return x * x * x / (std::exp(x) - 1.); };       // :109-113
    int st = 0;
    double ans = 0;
    const double whi = 1.0E7 / wvllo, wlo = 1.0E7 / wvlhi;                     // :141-142
    if (t < 0. || whi <= wlo || wlo < 0.) {                                    // :144-146 (reference exits)
        if (status) *status = 1;
        return NAN;
    }
    if (t < 1.e-4) {
        if (status) *status = 0;
        return 0.;
    }
    const double v0 = c2 * wlo / t, v1 = c2 * whi / t;
    if (v0 > DBL_EPSILON && v1 < vmax && (whi - wlo) / whi < 1.e-2) {          // Simpson branch :155-182
        const double hh = v1 - v0, ends = plkf(v0) + plkf(v1);
        double prev = 0., val = 0.;
        bool conv = false;
        for (int n = 1; n <= 10 && !conv; ++n) {
            const double del = hh / (2 * n);
            val = ends;
            for (int k = 1; k <= 2 * n - 1; ++k) val += (double)(2 * (1 + k % 2)) * plkf(v0 + (double)k * del);
            val *= del * a1;
            conv = std::fabs((val - prev) / val) <= 1.e-6;
            prev = val;
        }
        if (!conv) st = 2;
        ans = sigdpi * std::pow(t, 4.0) * conc * val;
    } else {                                                                    // general case :187-237
        const double v[2] = {v0, v1};
        double d[2] = {0, 0}, p[2] = {0, 0};
        int smallv = 0;
        for (int i = 0; i < 2; ++i) {
            if (v[i] < vcut) {
                ++smallv;
                const double vsq = v[i] * v[i];
                p[i] = conc * vsq * v[i] * (a1 + v[i] * (a2 + v[i] * (a3 + vsq * (a4 + vsq * (a5 + vsq * a6)))));
            } else {
                int mmax = 1;
                while (v[i] < vcp[mmax - 1]) ++mmax;
                const double ex = std::exp(-v[i]);
                double exm = 1.;
                for (int m = 1; m <= mmax; ++m) {
                    const double mv = (double)m * v[i];
                    exm = ex * exm;
                    d[i] += exm * (6. + mv * (6. + mv * (3. + mv))) / (m * m * m * m);
                }
                d[i] *= conc;
            }
        }
        ans = (smallv == 2) ? p[1] - p[0] : (smallv == 1) ? 1. - p[0] - d[1] : d[0] - d[1];
        ans *= sigdpi * std::pow(t, 4.0);
        if (ans == 0.) st = 3;
    }
    if (status) *status = st;
    return ans;
}

}  // extern "C"

// ---- synthetic line-by-line tables (the reference's lbl.*.asc are not distributed) -------------
// Smooth band envelopes (Gaussians in ln lambda at the main thermal bands of each gas) times a
// lognormal "line" factor per wavelength, times the absorber amount of the layer.  Only meant to
// have the right shape, dynamic range (1e-8 .. 1e3) and format; documented in DESIGN.md.
extern "C" int rcm_make_lbl_tables(int nwvl, unsigned long long seed, const double* plevel, const double* h2o,
                                   const double* o3, double* wvl, double* tau5) {
    if (nwvl < 2 || !plevel || !h2o || !o3 || !wvl || !tau5) return RCM_ERR_ARG;
    struct Band { double center_um, width, strength; };
    static const Band bands[5][3] = {
        {{6.3, 0.12, 300.0}, {40.0, 0.55, 1500.0}, {12.0, 0.6, 0.3}},   // H2O: nu2, rotation, continuum-like
        {{15.0, 0.07, 300.0}, {4.3, 0.03, 3000.0}, {10.4, 0.03, 0.02}}, // CO2
        {{9.6, 0.03, 3.0}, {14.2, 0.04, 0.3}, {4.75, 0.02, 0.2}},        // O3
        {{7.7, 0.04, 1.0}, {0, 0, 0}, {0, 0, 0}},                        // CH4
        {{7.8, 0.03, 0.6}, {4.5, 0.02, 3.0}, {17.0, 0.03, 0.1}}};        // N2O
    const int NL = RCM_NLAYER;
    const double lo = std::log(4000.0), hi = std::log(100000.0);
    for (int i = 0; i < nwvl; ++i) wvl[i] = std::exp(lo + (hi - lo) * (double)i / (double)(nwvl - 1));
    for (int sp = 0; sp < 5; ++sp) {
        Rng g{seed * 0x9E3779B97F4A7C15ull + 77777ull * (sp + 1)};
        for (int i = 0; i < nwvl; ++i) {
            const double lx = std::log(wvl[i] / 1000.0);
            double env = 0.0;
            for (const Band& b : bands[sp])
                if (b.width > 0) {
                    const double z = (lx - std::log(b.center_um)) / b.width;
                    env += b.strength * std::exp(-0.5 * z * z);
                }
            const double line = std::exp(1.5 * g.normal());
            for (int l = 0; l < NL; ++l) {
                const double dp = (plevel[l + 1] - plevel[l]) / 50.0;
                const double pm = (plevel[l + 1] + plevel[l]) / 2000.0;
                double amount = dp;
                if (sp == 0) amount *= h2o[l] / 7.0e-3;
                if (sp == 2) amount *= o3[l] / 1.0e-6;
                // pressure broadening: weaker absorption aloft in the band wings
                tau5[((size_t)sp * nwvl + i) * NL + l] = (env * line * (0.3 + 0.7 * pm) + 1e-8) * amount / NL;
            }
        }
    }
    return RCM_OK;
}

extern "C" int rcm_write_lbl_asc(const char* path, int nwvl, const double* wvl, const double* tau) {
    if (!path || !wvl || !tau || nwvl < 1) return RCM_ERR_ARG;
    FILE* f = std::fopen(path, "w");
    if (!f) return RCM_ERR_IO;
    std::fprintf(f, "# wavelength [nm], delta_tau for the 20 layers, sorted top-down (lbl.arts/README)\n");
    for (int i = 0; i < nwvl; ++i) {
        std::fprintf(f, "%.17g", wvl[i]);
        for (int l = 0; l < RCM_NLAYER; ++l) std::fprintf(f, " %.17g", tau[(size_t)i * RCM_NLAYER + l]);
        std::fprintf(f, "\n");
    }
    std::fclose(f);
    return RCM_OK;
}
