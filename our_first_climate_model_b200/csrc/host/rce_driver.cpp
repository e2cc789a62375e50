// rcm_rce - the RCE driver for ensembles, host C++ over the C ABI (include/rcm_b200.h).
//
// The reference's driver is main() of main.cpp: one hard-coded column (main.cpp:396-430), hard-coded table
// (main.cpp:500), a fixed number of iterations (main.cpp:531) and rows appended to output.txt (main.cpp:102-114).
// This is the same run for an ensemble of columns on one GPU: read the .atm file, build ncol perturbed members
// (member 0 is the file's own column), solar setup, upload, iterate until every column is stationary or
// max_steps is reached, write the profiles in the reference's row format and, optionally, a checkpoint.
//
//   rcm_rce --atm test.atm --table Reduced100Forcing.nc [--ncol N] [--seed S] [--max-steps M] [--check-every K]
//           [--dT 1e-3] [--device D] [--out output.txt] [--checkpoint file] [--resume file] [--steps-exact]
//   rcm_rce --atm fpda.lbl.atm --lbl DIR [--co2-factor F] ...     line-by-line: DIR/lbl.{h2o,co2,o3,ch4,n2o}.asc
//
// The line-by-line form reads the five tables with the drop-in of ASCII_file2xy2D (lbl.arts/testlblarts.cpp:24-26 is the
// reference's only call of that reader); a six-column .atm (lbl.arts/README:1-3) gets the well-mixed CO2 / CH4 / N2O of
// lbl.arts/README:13-16.  The tables hold the optical depths of the file's own column: member 0's H2O and O3.
//
// --steps-exact runs exactly max_steps iterations (the reference's n_steps semantics, main.cpp:83) instead of
// stopping at stationarity.  Exit code 0, or 1 with the library's error text on stderr.  No CPU fallback.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rcm_b200.h"

namespace {

int die(const char* what, int st, const rcm_solver* s) {
    std::fprintf(stderr, "rcm_rce: %s: %s%s%s\n", what, rcm_status_string(st), s ? " - " : "", s ? rcm_last_error(s) : "");
    return 1;
}

}  // namespace

int main(int argc, char** argv) {
    std::string atm_path, table_path, lbl_dir, out_path = "output.txt", ckpt_path, resume_path;
    double co2_factor = 1.0;
    int ncol = 1, device = 0, check_every = 250, steps_exact = 0;
    long max_steps = 6000;
    unsigned long long seed = 12345;
    double dT = 1e-3;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
        if (a == "--atm") atm_path = val();
        else if (a == "--table") table_path = val();
        else if (a == "--lbl") lbl_dir = val();
        else if (a == "--co2-factor") co2_factor = std::atof(val());
        else if (a == "--ncol") ncol = std::atoi(val());
        else if (a == "--seed") seed = std::strtoull(val(), nullptr, 10);
        else if (a == "--max-steps") max_steps = std::atol(val());
        else if (a == "--check-every") check_every = std::atoi(val());
        else if (a == "--dT") dT = std::atof(val());
        else if (a == "--device") device = std::atoi(val());
        else if (a == "--out") out_path = val();
        else if (a == "--checkpoint") ckpt_path = val();
        else if (a == "--resume") resume_path = val();
        else if (a == "--steps-exact") steps_exact = 1;
        else {
            std::fprintf(stderr, "rcm_rce: unknown argument %s\n", a.c_str());
            return 1;
        }
    }
    if (atm_path.empty() || (table_path.empty() == lbl_dir.empty()) || ncol < 1 || max_steps < 1 || check_every < 1) {
        std::fprintf(stderr, "usage: rcm_rce --atm FILE (--table FILE | --lbl DIR [--co2-factor F]) [--ncol N] [--seed S] [--max-steps M] [--check-every K] "
                             "[--dT K/step] [--device D] [--out FILE] [--checkpoint FILE] [--resume FILE] [--steps-exact]\n");
        return 1;
    }

    // ---- the column of the .atm file (main.cpp:396-430) and its ensemble ---------------------------
    constexpr int NLEV = RCM_NLEVEL, NLAY = RCM_NLAYER;
    std::vector<double> cols(9 * NLEV);
    int nrows = 0, nc = 0;
    int st = rcm_read_atm(atm_path.c_str(), NLEV, cols.data(), &nrows, &nc);
    if (st != RCM_OK) return die(atm_path.c_str(), st, nullptr);
    const bool lbl = !lbl_dir.empty();
    if (nrows != NLEV || nc < (lbl ? 6 : 9)) {
        std::fprintf(stderr, "rcm_rce: %s: need 21 levels x %d columns (z p T air H2O O3%s), got %d x %d\n", atm_path.c_str(),
                     lbl ? 6 : 9, lbl ? "" : " CO2 CH4 N2O", nrows, nc);
        return 1;
    }
    if (nc < 9) {  // lbl.arts/README:13-16
        const double well_mixed[3] = {400.0, 1.7, 0.315};
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < NLEV; ++i) cols[(6 + k) * NLEV + i] = well_mixed[k];
    }
    const double* plevel = &cols[1 * NLEV];
    const double* Tbase = &cols[2 * NLEV];
    const double* vbase = &cols[4 * NLEV];  // H2O, O3, CO2, CH4, N2O: five consecutive columns
    const size_t n = (size_t)ncol;
    std::vector<double> Tlevel(n * NLEV), vlev(n * 5 * NLEV);
    st = rcm_make_ensemble(ncol, seed, plevel, Tbase, vbase, Tlevel.data(), vlev.data());
    if (st != RCM_OK) return die("rcm_make_ensemble", st, nullptr);
    std::vector<double> Tlayer(n * NLAY), vmr9(n * RCM_NSPECIES * NLAY), rel_hum(n * NLAY), player(NLAY), conv(NLAY);
    st = rcm_init_columns(ncol, plevel, Tlevel.data(), vlev.data(), 1.0, Tlayer.data(), vmr9.data(), rel_hum.data(),
                          player.data(), conv.data());  // main.cpp:439-479
    if (st != RCM_OK) return die("rcm_init_columns", st, nullptr);
    std::vector<double> Tsurf(n, 288.2);  // main.cpp:357

    // ---- constants, solar setup (main.cpp:214-264), solver, table ---------------------------------
    rcm_params p;
    rcm_default_params(&p);
    rcm_solar_params sp;
    rcm_default_solar_params(&sp);
    double sol[7];
    rcm_solar_setup(&sp, sol);
    p.solar_irr = sol[6];
    p.dT_converged = dT;
    rcm_solver* s = nullptr;
    st = rcm_create(device, &p, &s);
    if (st != RCM_OK) return die("rcm_create", st, nullptr);
    if (lbl) {
        // the five species tables, README order of the composition: H2O, CO2, O3, CH4, N2O (rcm_set_lbl_tables)
        const char* names[5] = {"h2o", "co2", "o3", "ch4", "n2o"};
        std::vector<double> wvl, tau5;
        int nwvl = 0;
        for (int k = 0; k < 5; ++k) {
            const std::string path = lbl_dir + "/lbl." + names[k] + ".asc";
            int nx = 0, ny = 0;
            double *x = nullptr, *y = nullptr;
            const int rc = rcm_ascii_file2xy2D(path.c_str(), &nx, &ny, &x, &y);  // 0, or the reference's -1 / -2 / -5
            if (rc != 0 || ny != NLAY || nx < 2 || (k > 0 && nx != nwvl)) {
                std::fprintf(stderr, "rcm_rce: %s: reader status %d, %d wavelengths x %d layers (need the same wavelengths in "
                                     "all five files and 20 layers)\n", path.c_str(), rc, nx, ny);
                return 1;
            }
            if (k == 0) {
                nwvl = nx;
                wvl.assign(x, x + nx);
                tau5.resize((size_t)5 * nx * NLAY);
            } else if (std::memcmp(wvl.data(), x, (size_t)nx * sizeof(double)) != 0) {
                std::fprintf(stderr, "rcm_rce: %s: wavelength grid differs from lbl.h2o.asc\n", path.c_str());
                return 1;
            }
            std::memcpy(&tau5[(size_t)k * nx * NLAY], y, (size_t)nx * NLAY * sizeof(double));
            rcm_free(x);
            rcm_free(y);
        }
        // reference profiles of the tables = the .atm file's own column = member 0 (layer VMRs, species 0 and 2 of vmr9)
        st = rcm_set_lbl_tables(s, wvl.data(), tau5.data(), nwvl, &vmr9[0 * NLAY], &vmr9[2 * NLAY], co2_factor);
        if (st != RCM_OK) return die("rcm_set_lbl_tables", st, s);
    } else {
        rcm_table* t = nullptr;
        st = rcm_table_load(table_path.c_str(), &t);
        if (st != RCM_OK) return die(table_path.c_str(), st, s);
        st = rcm_set_repwvl_table_from(s, t);
        rcm_table_free(t);
        if (st != RCM_OK) return die("rcm_set_repwvl_table_from", st, s);
    }
    if (!resume_path.empty()) {
        st = rcm_load_checkpoint(s, resume_path.c_str());
        if (st != RCM_OK) return die(resume_path.c_str(), st, s);
        ncol = rcm_column_count(s);
    } else {
        st = rcm_set_columns(s, ncol, plevel, Tlayer.data(), Tsurf.data(), vmr9.data(), rel_hum.data());
        if (st != RCM_OK) return die("rcm_set_columns", st, s);
    }

    // ---- the time loop (main.cpp:531-583), in blocks of check_every fused iterations ---------------
    rcm_step_scalars last{};
    long done = 0;
    if (steps_exact) {
        std::vector<rcm_step_scalars> sc((size_t)check_every);
        while (done < max_steps) {
            const int k = (int)((max_steps - done < check_every) ? (max_steps - done) : check_every);
            st = rcm_advance(s, k, sc.data());
            if (st != RCM_OK) return die("rcm_advance", st, s);
            last = sc[k - 1];
            done += k;
        }
    } else {
        st = rcm_run_to_equilibrium(s, max_steps, check_every, &last, &done);
        if (st != RCM_OK) return die("rcm_run_to_equilibrium", st, s);
    }

    // ---- results: the reference's rows for every member, optional checkpoint -----------------------
    const size_t m = (size_t)ncol;
    std::vector<double> T(m * NLAY), Ts(m), Eu(m * NLEV);
    std::vector<float> time_h(m);
    st = rcm_get_state(s, T.data(), Ts.data(), nullptr, time_h.data(), nullptr, Eu.data(), nullptr, nullptr);
    if (st != RCM_OK) return die("rcm_get_state", st, s);
    st = rcm_write_profiles(out_path.c_str(), /*append*/ 0, /*header*/ 1, ncol, plevel, T.data(), time_h.data(),
                            /*column ids*/ ncol > 1);
    if (st != RCM_OK) return die(out_path.c_str(), st, s);
    if (!ckpt_path.empty()) {
        st = rcm_save_checkpoint(s, ckpt_path.c_str());
        if (st != RCM_OK) return die(ckpt_path.c_str(), st, s);
    }
    double ts_mean = 0.0;
    for (size_t c = 0; c < m; ++c) ts_mean += Ts[c];
    std::printf("rcm_rce: %d columns, %ld iterations, %d/%d stationary (< %g K per step), max dT %.3e K, mean TOA net %.4f W/m2, "
                "mean T_surface %.4f K, member 0: T_surface %.6f K, OLR %.6f W/m2, time %.2f h\n",
                ncol, done, (int)last.n_converged, ncol, dT, last.max_dT, last.toa_net_sum / ncol, ts_mean / ncol, Ts[0], Eu[0],
                (double)time_h[0]);
    rcm_destroy(s);
    return 0;
}
