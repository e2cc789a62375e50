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
//   rcm_rce ... --gpus N                                          the same ensemble sharded over N GPUs of this node
//
// --gpus N (N > 1): one process, one host thread and one solver per GPU, contiguous blocks of columns per GPU (tables
// replicated), NCCL directly: after every block of --check-every fused iterations ONE pair of ncclAllReduce calls (sum,
// max) over the block's four scalars per step decides - identically on every GPU - whether the whole ensemble is
// stationary.  Nothing else crosses GPUs.  Per-column results do not depend on N (the solver's order of additions is
// fixed per column): the profile rows of an N-GPU run are byte-identical to the 1-GPU run's.
//
// The line-by-line form reads the five tables with the drop-in of ASCII_file2xy2D (lbl.arts/testlblarts.cpp:24-26 is the
// reference's only call of that reader); a six-column .atm (lbl.arts/README:1-3) gets the well-mixed CO2 / CH4 / N2O of
// lbl.arts/README:13-16.  The tables hold the optical depths of the file's own column: member 0's H2O and O3.
//
// --steps-exact runs exactly max_steps iterations (the reference's n_steps semantics, main.cpp:83) instead of
// stopping at stationarity.  Exit code 0, or 1 with the library's error text on stderr.  No CPU fallback.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>
#include <nccl.h>

#include "rcm_b200.h"

namespace {

int die(const char* what, int st, const rcm_solver* s) {
    std::fprintf(stderr, "rcm_rce: %s: %s%s%s\n", what, rcm_status_string(st), s ? " - " : "", s ? rcm_last_error(s) : "");
    return 1;
}

// ---- N GPUs: one thread per GPU, columns [lo, hi) each ---------------------------------------------------------
struct Shard {
    int dev = 0, lo = 0, hi = 0;
    rcm_solver* s = nullptr;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    double *d_sum = nullptr, *d_max = nullptr;  // [check_every][4] reduced scalars of a block
    long done = 0;
    rcm_step_scalars last{};
    double loop_s = 0.0;  // wall time of the time loop (first launch to last block decided)
    std::string err;
};

struct Job {  // what every shard needs; host arrays cover the WHOLE ensemble
    const rcm_params* p;
    const rcm_table* table;  // repwvl, or NULL
    const std::vector<double>*wvl, *tau5;
    int lbl_nwvl;
    double co2_factor;
    const double *plevel, *Tlayer, *Tsurf, *vmr9, *rel_hum;
    int ncol;
    long max_steps;
    int check_every, steps_exact;
    std::string resume, ckpt;
    double *T_out, *Ts_out, *Eu_out;
    float* time_out;
};

std::atomic<int> g_failed{0};

void run_shard(Shard* sh, const Job* j) {
    auto fail = [&](const char* what, int st) {
        sh->err = std::string(what) + ": " + rcm_status_string(st) + (sh->s ? std::string(" - ") + rcm_last_error(sh->s) : "");
        g_failed.store(1);
    };
    constexpr int NLAY = RCM_NLAYER, NLEV = RCM_NLEVEL;
    const size_t lo = (size_t)sh->lo, n = (size_t)(sh->hi - sh->lo);
    int st = rcm_create(sh->dev, j->p, &sh->s);
    if (st != RCM_OK) return fail("rcm_create", st);
    if (cudaSetDevice(sh->dev) != cudaSuccess || cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&sh->d_sum, (size_t)j->check_every * 4 * sizeof(double)) != cudaSuccess ||
        cudaMalloc((void**)&sh->d_max, (size_t)j->check_every * 4 * sizeof(double)) != cudaSuccess)
        return fail("cuda setup", RCM_ERR_CUDA);
    rcm_set_stream(sh->s, sh->stream);  // the solver's kernels and the NCCL calls share one stream: ordered without events
    st = j->table ? rcm_set_repwvl_table_from(sh->s, j->table)
                  : rcm_set_lbl_tables(sh->s, j->wvl->data(), j->tau5->data(), j->lbl_nwvl, j->vmr9 + 0 * NLAY, j->vmr9 + 2 * NLAY,
                                       j->co2_factor);
    if (st != RCM_OK) return fail("tables", st);
    if (!j->resume.empty()) {
        st = rcm_load_checkpoint(sh->s, (j->resume + ".gpu" + std::to_string(sh->dev)).c_str());
        if (st == RCM_OK && rcm_column_count(sh->s) != (int)n) st = RCM_ERR_STATE;
    } else {
        st = rcm_set_columns(sh->s, (int)n, j->plevel, j->Tlayer + lo * NLAY, j->Tsurf + lo, j->vmr9 + lo * RCM_NSPECIES * NLAY,
                             j->rel_hum + lo * NLAY);
    }
    if (st != RCM_OK) return fail("columns", st);
    // ---- the time loop (main.cpp:531-583) in blocks; one allreduce pair per block ------------------------------
    std::vector<double> hs(4), hm(4);
    // warm-up outside the clock: the first launch loads the kernels, the first collective builds NCCL's channels; the
    // ensemble then restarts from its initial state (a resumed run keeps its state: only the collective is warmed up)
    if (j->resume.empty()) {
        st = rcm_advance(sh->s, 1, nullptr);
        if (st == RCM_OK)
            st = rcm_set_columns(sh->s, (int)n, j->plevel, j->Tlayer + lo * NLAY, j->Tsurf + lo, j->vmr9 + lo * RCM_NSPECIES * NLAY,
                                 j->rel_hum + lo * NLAY);
        if (st != RCM_OK) return fail("warm-up", st);
    }
    cudaMemsetAsync(sh->d_sum, 0, 4 * sizeof(double), sh->stream);
    if (ncclAllReduce(sh->d_sum, sh->d_max, 4, ncclDouble, ncclSum, sh->comm, sh->stream) != ncclSuccess)
        return fail("ncclAllReduce (warm-up)", RCM_ERR_CUDA);
    cudaStreamSynchronize(sh->stream);
    const auto t0 = std::chrono::steady_clock::now();
    while (sh->done < j->max_steps && !g_failed.load()) {
        const int k = (int)((j->max_steps - sh->done < j->check_every) ? (j->max_steps - sh->done) : j->check_every);
        double* d_sc = nullptr;
        st = rcm_advance_async(sh->s, k, &d_sc);
        if (st != RCM_OK) return fail("rcm_advance_async", st);
        if (ncclAllReduce(d_sc, sh->d_sum, (size_t)k * 4, ncclDouble, ncclSum, sh->comm, sh->stream) != ncclSuccess ||
            ncclAllReduce(d_sc, sh->d_max, (size_t)k * 4, ncclDouble, ncclMax, sh->comm, sh->stream) != ncclSuccess)
            return fail("ncclAllReduce", RCM_ERR_CUDA);
        cudaMemcpyAsync(hs.data(), sh->d_sum + (size_t)(k - 1) * 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, sh->stream);
        cudaMemcpyAsync(hm.data(), sh->d_max + (size_t)(k - 1) * 4, 4 * sizeof(double), cudaMemcpyDeviceToHost, sh->stream);
        if (cudaStreamSynchronize(sh->stream) != cudaSuccess) return fail("block", RCM_ERR_CUDA);
        sh->done += k;
        sh->last = {hs[0], hm[1], hs[2], hm[3]};  // sums: TOA net, converged count; maxima: dT, |dE|
        if (!j->steps_exact && sh->last.n_converged >= (double)j->ncol) break;  // the same numbers on every GPU
    }
    sh->loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (g_failed.load()) return;
    st = rcm_get_state(sh->s, j->T_out + lo * NLAY, j->Ts_out + lo, nullptr, j->time_out + lo, nullptr, j->Eu_out + lo * NLEV, nullptr,
                       nullptr);
    if (st != RCM_OK) return fail("rcm_get_state", st);
    if (!j->ckpt.empty()) {
        st = rcm_save_checkpoint(sh->s, (j->ckpt + ".gpu" + std::to_string(sh->dev)).c_str());
        if (st != RCM_OK) return fail("rcm_save_checkpoint", st);
    }
}

}  // namespace

int main(int argc, char** argv) {
    std::string atm_path, table_path, lbl_dir, out_path = "output.txt", ckpt_path, resume_path;
    double co2_factor = 1.0;
    int ncol = 1, device = 0, check_every = 250, steps_exact = 0, gpus = 1;
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
        else if (a == "--gpus") gpus = std::atoi(val());
        else {
            std::fprintf(stderr, "rcm_rce: unknown argument %s\n", a.c_str());
            return 1;
        }
    }
    if (atm_path.empty() || (table_path.empty() == lbl_dir.empty()) || ncol < 1 || max_steps < 1 || check_every < 1 || gpus < 1 ||
        gpus > ncol) {
        std::fprintf(stderr, "usage: rcm_rce --atm FILE (--table FILE | --lbl DIR [--co2-factor F]) [--ncol N] [--seed S] [--max-steps M] [--check-every K] "
                             "[--dT K/step] [--device D] [--out FILE] [--checkpoint FILE] [--resume FILE] [--steps-exact] [--gpus N]\n");
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
    if (gpus > 1) {
        // ---- one thread and one solver per GPU, NCCL for the block scalars --------------------------------------
        if (rcm_device_count() < gpus) {
            std::fprintf(stderr, "rcm_rce: --gpus %d but %d CUDA device(s) visible\n", gpus, rcm_device_count());
            return 1;
        }
        rcm_table* table = nullptr;
        std::vector<double> wvl, tau5;
        int nwvl = 0;
        if (lbl) {
            const char* names[5] = {"h2o", "co2", "o3", "ch4", "n2o"};
            for (int k = 0; k < 5; ++k) {
                const std::string path = lbl_dir + "/lbl." + names[k] + ".asc";
                int nx = 0, ny = 0;
                double *x = nullptr, *y = nullptr;
                const int rc = rcm_ascii_file2xy2D(path.c_str(), &nx, &ny, &x, &y);
                if (rc != 0 || ny != NLAY || nx < 2 || (k > 0 && (nx != nwvl || std::memcmp(wvl.data(), x, (size_t)nx * sizeof(double)) != 0))) {
                    std::fprintf(stderr, "rcm_rce: %s: reader status %d, %d wavelengths x %d layers\n", path.c_str(), rc, nx, ny);
                    return 1;
                }
                if (k == 0) {
                    nwvl = nx;
                    wvl.assign(x, x + nx);
                    tau5.resize((size_t)5 * nx * NLAY);
                }
                std::memcpy(&tau5[(size_t)k * nx * NLAY], y, (size_t)nx * NLAY * sizeof(double));
                rcm_free(x);
                rcm_free(y);
            }
        } else {
            st = rcm_table_load(table_path.c_str(), &table);
            if (st != RCM_OK) return die(table_path.c_str(), st, nullptr);
        }
        std::vector<double> T(n * NLAY), Ts(n), Eu(n * NLEV);
        std::vector<float> time_h(n);
        Job job{&p, table, &wvl, &tau5, nwvl, co2_factor, plevel, Tlayer.data(), Tsurf.data(), vmr9.data(), rel_hum.data(), ncol,
                max_steps, check_every, steps_exact, resume_path, ckpt_path, T.data(), Ts.data(), Eu.data(), time_h.data()};
        std::vector<Shard> shards((size_t)gpus);
        std::vector<ncclComm_t> comms((size_t)gpus);
        std::vector<int> devs((size_t)gpus);
        for (int d = 0; d < gpus; ++d) devs[d] = d;
        if (ncclCommInitAll(comms.data(), gpus, devs.data()) != ncclSuccess) {
            std::fprintf(stderr, "rcm_rce: ncclCommInitAll failed for %d GPUs\n", gpus);
            return 1;
        }
        const int base = ncol / gpus, rem = ncol % gpus;  // contiguous blocks, sizes differ by at most one
        for (int d = 0, lo = 0; d < gpus; ++d) {
            shards[d].dev = d;
            shards[d].lo = lo;
            shards[d].hi = lo += base + (d < rem ? 1 : 0);
            shards[d].comm = comms[d];
        }
        std::vector<std::thread> th;
        for (int d = 0; d < gpus; ++d) th.emplace_back(run_shard, &shards[d], &job);
        for (auto& t : th) t.join();
        int bad = 0;
        for (auto& sh : shards)
            if (!sh.err.empty()) {
                std::fprintf(stderr, "rcm_rce: GPU %d: %s\n", sh.dev, sh.err.c_str());
                bad = 1;
            }
        for (auto& sh : shards) {
            if (sh.s) rcm_destroy(sh.s);
            if (sh.comm) ncclCommDestroy(sh.comm);
        }
        if (table) rcm_table_free(table);
        if (bad) return 1;
        st = rcm_write_profiles(out_path.c_str(), 0, 1, ncol, plevel, T.data(), time_h.data(), ncol > 1);
        if (st != RCM_OK) return die(out_path.c_str(), st, nullptr);
        const rcm_step_scalars& last = shards[0].last;
        double ts_mean = 0.0;
        for (size_t c = 0; c < n; ++c) ts_mean += Ts[c];
        double loop_s = 0.0;
        for (auto& sh : shards) loop_s = sh.loop_s > loop_s ? sh.loop_s : loop_s;
        std::printf("rcm_rce: %d columns on %d GPUs, %ld iterations, %d/%d stationary (< %g K per step), max dT %.3e K, mean TOA net %.4f W/m2, "
                    "mean T_surface %.4f K, member 0: T_surface %.6f K, OLR %.6f W/m2, time %.2f h\n",
                    ncol, gpus, shards[0].done, (int)last.n_converged, ncol, dT, last.max_dT, last.toa_net_sum / ncol, ts_mean / ncol,
                    Ts[0], Eu[0], (double)time_h[0]);
        std::printf("rcm_rce: time loop %.4f s on the slowest GPU = %.4f ms per iteration\n", loop_s, 1e3 * loop_s / (double)shards[0].done);
        return 0;
    }
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
    rcm_synchronize(s);
    const auto t0 = std::chrono::steady_clock::now();
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

    rcm_synchronize(s);
    const double loop_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
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
    std::printf("rcm_rce: time loop %.4f s = %.4f ms per iteration\n", loop_s, 1e3 * loop_s / (double)done);
    rcm_destroy(s);
    return 0;
}
