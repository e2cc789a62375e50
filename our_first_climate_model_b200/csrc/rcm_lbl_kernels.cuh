// rcm_lbl_kernels.cuh - band-integrated Planck function and the three line-by-line kernels
// Included by rcm_kernels.cu inside its anonymous namespace (one translation unit: the kernels share the
// __constant__ bank `cst` and the device functions are force-inlined).
#pragma once

// ------------------------------------------------------------------------------------------
// Band-integrated Planck radiance on the device (K2 of the line-by-line path): libRadtran's
// c_planck_func1 as vendored by the reference (cplkavg.cpp:124-243), branch for branch.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double plkf(double x) { return x * x * x / (exp(x) - 1.); }

__device__ __noinline__ double cplkavg_dev(double wvllo, double wvlhi, double t) {
    const double c2 = 1.438786, sigma = 5.67032E-8, vcut = 1.5;
    const double a1 = 1. / 3., a2 = -1. / 8., a3 = 1. / 60., a4 = -1. / 5040., a5 = 1. / 272160.,
                 a6 = -1. / 13305600.;
    const double vcp[7] = {10.25, 5.7, 3.9, 2.9, 2.3, 1.9, 0.0};
    const double pi = 3.14159265358979323846;
    const double vmax = 709.782712893384, sigdpi = sigma / pi, conc = 15. / (pi * pi * pi * pi);
    const double whi = 1.0E7 / wvllo, wlo = 1.0E7 / wvlhi;
    if (t < 0. || whi <= wlo || wlo < 0.) return __longlong_as_double(0x7ff8000000000000ULL);
    if (t < 1.e-4) return 0.;
    const double v0 = c2 * wlo / t, v1 = c2 * whi / t;
    const double t4 = (t * t) * (t * t);
    if (v0 > DBL_EPSILON && v1 < vmax && (whi - wlo) / whi < 1.e-2) {
        const double hh = v1 - v0, ends = plkf(v0) + plkf(v1);
        double prev = 0., val = 0.;
        for (int n = 1; n <= 10; ++n) {
            const double del = hh / (2 * n);
            val = ends;
            for (int k = 1; k <= 2 * n - 1; ++k) val += (double)(2 * (1 + k % 2)) * plkf(v0 + (double)k * del);
            val *= del * a1;
            if (fabs((val - prev) / val) <= 1.e-6) break;
            prev = val;
        }
        return sigdpi * t4 * conc * val;
    }
    double d[2] = {0., 0.}, p[2] = {0., 0.};
    int smallv = 0;
    const double v[2] = {v0, v1};
    for (int i = 0; i < 2; ++i) {
        if (v[i] < vcut) {
            ++smallv;
            const double vsq = v[i] * v[i];
            p[i] = conc * vsq * v[i] * (a1 + v[i] * (a2 + v[i] * (a3 + vsq * (a4 + vsq * (a5 + vsq * a6)))));
        } else {
            int mmax = 1;
            while (v[i] < vcp[mmax - 1]) ++mmax;
            const double ex = exp(-v[i]);
            double exm = 1.;
            for (int m = 1; m <= mmax; ++m) {
                const double mv = (double)m * v[i];
                exm = ex * exm;
                d[i] += exm * (6. + mv * (6. + mv * (3. + mv))) / (double)(m * m * m * m);
            }
            d[i] *= conc;
        }
    }
    const double ans = (smallv == 2) ? p[1] - p[0] : (smallv == 1) ? 1. - p[0] - d[1] : d[0] - d[1];
    return ans * (sigdpi * t4);
}

// The same function for the LBL kernel's inner loop.  LBL bins are narrow ((hi-lo)/hi < 1e-2), which is the
// Simpson branch (cplkavg.cpp:155-182): 2 + 1 + 3 evaluations of x^3/(exp(x)-1), converged at n = 2.  Here
// with the solver's exp (exp_scaled, <= 1 ulp like libm's) and division (div_fast, <= 1 ulp) instead of the
// library routines, and with the two wavenumbers 1e7/lambda taken once per wavelength by the caller; every other
// case goes to cplkavg_dev.  Same control flow and summation order, results within a few ulp of it.
// bin_ok: whi > wlo && wlo >= 0 && (whi - wlo) / whi < 1e-2 - a property of the bin, evaluated once per wavelength.
__device__ __forceinline__ double cplkavg_narrow(double wvllo, double wvlhi, double whi, double wlo, bool bin_ok, double t,
                                                 unsigned tab_lane) {
    const double c2 = 1.438786, sigma = 5.67032E-8, pi = 3.14159265358979323846;
    const double vmax = 709.782712893384, sigdpi = sigma / pi, conc = 15. / (pi * pi * pi * pi);
    const double v0 = div_fast(c2 * wlo, t), v1 = div_fast(c2 * whi, t);
    if (!(bin_ok && t >= 1.e-4 && v0 > DBL_EPSILON && v1 < vmax)) return cplkavg_dev(wvllo, wvlhi, t);
    auto f = [&](double x) { return div_fast(x * x * x, exp_scaled<false>(x, L2E64, tab_lane) - 1.); };
    const double hh = v1 - v0;
    const double t4 = (t * t) * (t * t);
    // n = 1 and n = 2 in straight-line code (five evaluations instead of a data-dependent loop): the midpoint
    // v0 + 2 * (hh / 4) of n = 2 is bit-identical to v0 + 1 * (hh / 2) of n = 1 (exact scaling by powers of two), so its
    // value is reused; same summation order as the loop below.  n = 1 never passes the convergence test (prev = 0),
    // n = 2 nearly always does for LBL bins.
    // The five abscissae are equidistant, so their exponentials are exp(v0) * exp(hh/4)^k: two exp's and four
    // products instead of five exp's (a few ulp each; exp(x) - 1 amplifies that by at most 1/x, hence only for
    // v0 >= 1/4 - thermal LBL bins have x between 0.5 and 18).
    const double del1 = hh * 0.5, del2 = hh * 0.25;
    const double x1 = v0 + del2, x2 = v0 + del1, x3 = v0 + 3.0 * del2;
    double fa, fb, fm, fq1, fq3;
    if (v0 >= 0.25) {
        auto gx = [&](double x, double e) { return div_fast(x * x * x, e - 1.); };
        const double e0 = exp_scaled<false>(v0, L2E64, tab_lane), r = exp_scaled<false>(del2, L2E64, tab_lane);
        const double e1 = e0 * r, e2 = e1 * r, e3 = e2 * r, e4 = e3 * r;
        fa = gx(v0, e0); fq1 = gx(x1, e1); fm = gx(x2, e2); fq3 = gx(x3, e3); fb = gx(v1, e4);
    } else {
        fa = f(v0); fq1 = f(x1); fm = f(x2); fq3 = f(x3); fb = f(v1);
    }
    const double ends = fa + fb;
    double prev = (ends + 4.0 * fm) * (del1 * (1. / 3.));
    double val = (((ends + 4.0 * fq1) + 2.0 * fm) + 4.0 * fq3) * (del2 * (1. / 3.));
    // convergence test of cplkavg.cpp:172 without its division: |val - prev| <= 1e-6 |val| decides the same way except when
    // the quotient rounds across 1e-6 exactly, and then both Simpson orders agree to 1e-6 anyway (an IEEE division here is
    // ~25 instructions with its special-case paths, 3 % of the LBL kernel)
    if (fabs(val - prev) <= 1.e-6 * fabs(val)) return sigdpi * t4 * conc * val;
    prev = val;
    for (int n = 3; n <= 10; ++n) {
        const double del = hh / (2 * n);
        val = ends;
        for (int k = 1; k <= 2 * n - 1; ++k) val += (double)(2 * (1 + k % 2)) * f(v0 + (double)k * del);
        val *= del * (1. / 3.);
        if (fabs((val - prev) / val) <= 1.e-6) break;
        prev = val;
    }
    return sigdpi * t4 * conc * val;
}

// ------------------------------------------------------------------------------------------
// Line-by-line path (BASELINE configs 3 and 5).  The reference ships the table format
// (lbl.arts/README:5-16), the reader and cplkavg() but no driver; the composition below is the one
// documented in DESIGN.md section 5 (and restated on the CPU for the tests):
//   tau = tau_H2O*s_H2O(l) + f_CO2*tau_CO2 + tau_O3*s_O3(l) + tau_CH4 + tau_N2O   (left to right)
//   source = cplkavg(lo_w, hi_w, T) with unit spectral weight, sweeps as main.cpp:297-341.
// Three kernels per step: prep (theta-sort, feedback, scale factors), rt (tau, source, sweeps,
// partial fluxes per wavelength chunk), finish (sum of the chunks, dE, time step, T update).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rcm_lbl_prep_kernel(const LblArgs a) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.ncol) return;
    double th[NLAY];
#pragma unroll
    for (int l = 0; l < NLAY; ++l) th[l] = a.Tlayer[(size_t)col * NLAY + l] * cst.conv[l];  // main.cpp:536
#pragma unroll
    for (int pass = 0; pass < NLAY; ++pass) {
#pragma unroll
        for (int l = (pass & 1); l + 1 < NLAY; l += 2) cex(th[l], th[l + 1]);
    }
    double dmax = 0.0;
#pragma unroll
    for (int l = 0; l < NLAY; ++l) {
        const size_t gi = (size_t)col * NLAY + l;
        const double Tn = th[l] / cst.conv[l];  // main.cpp:540
        a.Tlayer[gi] = Tn;
        dmax = fmax(dmax, fabs(Tn - a.Tprev[gi]));
        a.Tprev[gi] = Tn;
        double h2o = a.vmr[((size_t)col * a.nact + a.h2o_slot) * NLAY + l];
        if (a.step_index != 0) {  // water_vapor_feedback, main.cpp:281-289
            const double Tc = Tn - 273.15;
            h2o = a.rel_hum[gi] * (6.1094 * exp(17.625 * Tc / (Tc + 243.04))) / cst.player[l];
            a.vmr[((size_t)col * a.nact + a.h2o_slot) * NLAY + l] = h2o;
        }
        a.sH[gi] = h2o / a.h2o_ref[l];
        a.sO[gi] = (a.o3_slot >= 0 && a.o3_ref) ? a.vmr[((size_t)col * a.nact + a.o3_slot) * NLAY + l] / a.o3_ref[l] : 1.0;
    }
    a.dTstat[col] = dmax;
}

template <int C, int NT, bool CLAMPK>
__global__ void __launch_bounds__(NT, 384 / NT) rcm_lbl_rt_kernel(const LblArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int G = NT / (2 * C), GC = G * C;
    double* p = reinterpret_cast<double*>(smem_raw);
    double* s_tab = p; p += EXP_TAB * EXP_REP;
    static_assert(C == 16, "tbd() is laid out for 16-column tiles");
    double* s_T = p;   p += TBD_LEN;  // [20][C] in pair order, rows 10..19 half a bank row further on (tbd)
    double* s_sH = p;  p += TBD_LEN;
    double* s_sO = p;  p += TBD_LEN;
    double* s_Ts = p;  p += C;
    double* s_cl = p;  p += C;
    double* s_B = p;   p += HALF * NT;  // [10][NT] Planck source of the thread's ten layers (written by a rolled loop)
    double* s_Ep = p;  // [21][GC]
    const int tid = threadIdx.x, lane = tid & 31;
    const int h = tid & 1, q = tid >> 1, c = q % C, g = q / C;
    const int sb = tbd(h * HALF, c);
    for (int i = tid; i < EXP_TAB * EXP_REP; i += NT) s_tab[i] = a.exp_tab[i / EXP_REP];
    const unsigned tab_lane = (unsigned)__cvta_generic_to_shared(s_tab + (lane & (EXP_REP - 1)));
    const int tile = blockIdx.x % a.ntiles, chunk = blockIdx.x / a.ntiles;
    const int col0 = tile * C, ncl = min(C, a.ncol - col0);
    for (int i = tid; i < NLAY * C; i += NT) {
        const int l = i / C, cc = i % C, r = tbd(prow(l), cc);
        const bool ok = cc < ncl;
        const size_t gi = (size_t)(col0 + cc) * NLAY + l;
        s_T[r] = ok ? a.Tlayer[gi] : 250.0;
        s_sH[r] = ok ? a.sH[gi] : 1.0;
        s_sO[r] = ok ? a.sO[gi] : 1.0;
    }
    if (tid < C) {
        s_Ts[tid] = (tid < ncl) ? a.Tsurf[col0 + tid] : 250.0;
        s_cl[tid] = a.cloud_col ? a.cloud_col[col0 + (tid < ncl ? tid : 0)] : cst.cloud_tau;
    }
    __syncthreads();

    double E1[HALF], E2[HALF], Eu20 = 0.0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) E1[j] = E2[j] = 0.0;
    // chunk_len is a multiple of G: every thread runs chunk_len / G items (uniform trip count, see the step
    // kernel); items beyond the last wavelength repeat it with a zero source.
    const int w_lo = chunk * a.chunk_len;
    const int tau_clamp_hi = __double2hiint(a.tau_clamp);
    const size_t plane = (size_t)a.nwvl * NLAY;
#pragma unroll 1
    for (int item = 0; item < a.chunk_len / G; ++item) {
        const int w_any = w_lo + g + item * G;
        const bool real = w_any < a.nwvl;
        const int w = real ? w_any : a.nwvl - 1;
        double tau[HALF], Bo[HALF];
        // the bin: edges [nm], wavenumbers 1e7 / lambda (cplkavg.cpp:141-142) and the narrow-band flag, all taken once per
        // wavelength on the host (rcm_set_lbl_tables): two IEEE divisions per wavelength and thread otherwise
        const double lo = __ldg(a.wvl_lo + w), hi = __ldg(a.wvl_hi + w);
        const double whi = __ldg(a.wn_hi + w), wlo = __ldg(a.wn_lo + w);
        const bool bin_ok = __ldg(a.bin_ok + w) != 0;
        // tau = tau_H2O * s_H2O + tau_O3 * s_O3 + (f_CO2 * tau_CO2 + tau_CH4 + tau_N2O): the bracket does not depend on the
        // column and is summed once when the tables are uploaded (rcm_set_lbl_tables); two FMAs per layer and wavelength
        const double* t3 = a.tau3 + (size_t)w * NLAY;
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const int l = h ? (NLAY - 1 - j) : j;
            double v = fma(__ldg(t3 + l), s_sH[sb + j * C], fma(__ldg(t3 + plane + l), s_sO[sb + j * C], __ldg(t3 + 2 * plane + l)));
            v = fma(cst.cloud_w[h * HALF + j], s_cl[c], v);  // + cloud tau on the cloud layer, + 0 (exact) elsewhere
            tau[j] = CLAMPK ? v : clamp_hi(v, tau_clamp_hi);
        }
        // The band-integrated Planck function of the ten layers in a ROLLED loop through shared memory: inlined ten
        // times it made the kernel 9,900 instructions long and instruction fetch 6 % of its stalls.
#pragma unroll 1
        for (int j = 0; j < HALF; ++j) s_B[j * NT + tid] = cplkavg_narrow(lo, hi, whi, wlo, bin_ok, s_T[sb + j * C], tab_lane);
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const double B = s_B[j * NT + tid];
            Bo[j] = real ? B : 0.0;
        }
        const double Bsurf = cplkavg_narrow(lo, hi, whi, wlo, bin_ok, s_Ts[c], tab_lane);
        sweep_item<CLAMPK>(tau, Bo, real ? Bsurf : 0.0, h, tab_lane, E1, E2, Eu20);
    }
    // partial fluxes of this wavelength chunk: part[chunk][col][0..20] = E_down, [21..41] = E_up
    double* part = a.part + ((size_t)chunk * a.ncol + col0) * 42;
#pragma unroll
    for (int j = 0; j < HALF; ++j) s_Ep[(h ? (NLAY - 1 - j) : j) * GC + g * C + c] = h ? E2[j] : E1[j];
    __syncthreads();
    for (int i = tid; i < NLAY * C; i += NT) {
        const int l = i / C, cc = i % C;
        double sum = 0.0;
        for (int gg = 0; gg < G; ++gg) sum += s_Ep[l * GC + gg * C + cc];
        if (cc < ncl) part[(size_t)cc * 42 + l + 1] = sum;
    }
    if (tid < ncl) part[(size_t)tid * 42] = 0.0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < HALF; ++j) s_Ep[(h ? (NLAY - 1 - j) : j) * GC + g * C + c] = h ? E1[j] : E2[j];
    if (h) s_Ep[NLAY * GC + g * C + c] = Eu20;
    __syncthreads();
    for (int i = tid; i < NLEV * C; i += NT) {
        const int l = i / C, cc = i % C;
        double sum = 0.0;
        for (int gg = 0; gg < G; ++gg) sum += s_Ep[l * GC + gg * C + cc];
        if (cc < ncl) part[(size_t)cc * 42 + 21 + l] = sum;
    }
}

// One 64-thread CTA per column: thread r < 42 sums row r of the column's chunk partials (chunk order, from 0.0: the order of
// the thread-per-column version this replaces - that one kept 512 columns on four SMs and took 0.2 ms per step, as long as
// 3 % of the 8-GPU LBL step), then threads l < 20 finish the layer l of the step.
constexpr int LBL_FIN_NT = 64;
__global__ void __launch_bounds__(LBL_FIN_NT) rcm_lbl_finish_kernel(const LblArgs a) {
    __shared__ double sE[2 * NLEV], sdE[NLAY], sdt;
    const int col = blockIdx.x, r = threadIdx.x;
    if (r < 2 * NLEV) {
        const double* pp = a.part + (size_t)col * 42 + r;
        const size_t stride = (size_t)a.ncol * 42;
        double sum = 0.0;
#pragma unroll 8
        for (int k = 0; k < a.nchunks; ++k) sum += pp[(size_t)k * stride];  // fixed order: deterministic
        sE[r] = sum;
    }
    __syncthreads();
    const double* Ed = sE;
    const double* Eu = sE + NLEV;
    const double solar = a.solar_col ? a.solar_col[col] : cst.solar_irr;
    if (r < NLAY) {
        double d = Ed[r] - Ed[r + 1] + Eu[r + 1] - Eu[r];                      // main.cpp:338
        if (r == NLAY - 1) d += solar + Ed[NLAY] - Eu[NLAY];                   // main.cpp:341
        sdE[r] = d;
    }
    __syncthreads();
    if (r == 0) {
        double mx = -1e300, mabs = 0.0;
#pragma unroll
        for (int l = 0; l < NLAY; ++l) {
            const double d = sdE[l];
            if (mx < d) mx = d;
            mabs = fmax(mabs, fabs(d));
        }
        double dt = (double)(float)cst.max_dT / mx * (1004.0 * cst.dp * 100.0) / 9.80665;  // main.cpp:157
        if (dt > cst.dt_cap) dt = cst.dt_cap;
        sdt = dt;
        a.dt[col] = dt;
        a.time_h[col] += (float)dt / 3600;
        if (a.diag) {
            double* dg = a.diag + (size_t)col * 4;
            dg[0] = solar - Eu[0];
            dg[1] = a.dTstat[col];
            dg[2] = mabs;
            dg[3] = dt;
        }
    }
    __syncthreads();
    if (r < NLAY) {
        const size_t gi = (size_t)col * NLAY + r;
        const double Tl = a.Tlayer[gi] + sdE[r] * sdt * 9.80665 / (1004.0 * cst.dp * 100.0);  // main.cpp:169
        a.Tlayer[gi] = Tl;
        a.dE[gi] = sdE[r];
        if (r == NLAY - 1) a.Tsurf[col] = Tl * cst.conv[NLAY - 1];  // main.cpp:173
    }
    if (r < NLEV) {
        a.E_down[(size_t)col * NLEV + r] = Ed[r];
        a.E_up[(size_t)col * NLEV + r] = Eu[r];
    }
}
