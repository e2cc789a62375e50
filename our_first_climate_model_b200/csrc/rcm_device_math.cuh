// rcm_device_math.cuh - the solver's exp, division, LowerPos and small helpers
// Included by rcm_kernels.cu inside its anonymous namespace (one translation unit: the kernels share the
// __constant__ bank `cst` and the device functions are force-inlined).
#pragma once

// ------------------------------------------------------------------------------------------
// exp of (a*b) for the transmissions t = exp(-tau/mu) and the Planck exponent.  With N = EXP_TAB table entries per
// octave, z = a*b*N/ln2 (the caller passes b already scaled by N/ln2), k = round(z), f = z - k:
//   exp = 2^(k div N) * 2^((k mod N)/N) * exp(f*ln2/N),  |f| <= 1/2,   exp(f*c) - 1 = f*h(f), c = ln2/N.
// Default (rcm_kernels.cuh): N = 1024, ONE copy of the table (8 KB), h of degree 2 - the Taylor polynomial with its
// f^3 term economised onto the linear one (Chebyshev), 1.4e-16 relative: 7 FP64-pipe instructions + 4 others
// (LOP3, IMAD, LDS.64, IMAD).  Alternative: N = 128, eight copies side by side (8 KB), degree 3, 7.6e-17, 8 + 4.
// What the instructions around the FP64 ones cost was measured in isolation (tools/probe/exp_probe.cu, ten
// independent exp's at the solver's occupancy): the four "others" cost 7.2 cycles per exp on top of the 16 of its
// FP64 instructions - the table lookup alone 6.7 - while bank conflicts of an unreplicated table cost only 0.3.
// Hence one Horner step less (2 cycles) at the price of conflicts is a gain, and:
//  * the power of two is applied to the TABLE VALUE with one integer multiply-add on its high word,
//    hi += k << (20 - log2 N).  Since k = N m + j, that is (m << 20) + (j << (20 - log2 N)): the table entries are
//    stored with j << (20 - log2 N) pre-subtracted from their high word (rcm_create), so k needs no shift or mask;
//  * the Horner coefficients come from the constant bank (as literals they were re-materialised into uniform
//    registers in every block);
//  * no clamp of the exponent: the caller guarantees |k| / N <= 1000 (tau is clamped once per layer,
//    StepArgs::tau_clamp), unless CLAMPK, which clamps here for angle schedules that need it.
// ------------------------------------------------------------------------------------------
template <bool CLAMPK>
__device__ __forceinline__ double exp_scaled(double a, double b_l2e, unsigned tab_lane) {
    const double SHIFT = 6755399441055744.0;  // 1.5 * 2^52: the add leaves round(z) in the low word
    const double t = fma(a, b_l2e, SHIFT);
    // CLAMPK: the clamp is taken on the double (with 1024 table entries per octave the integer itself can leave int32)
    // (as a comparison, not fmax: nvcc 12.9 folds fmax(t, constant) of this expression into the constant)
    const int k = (CLAMPK && t < SHIFT - 1000.0 * EXP_TAB) ? -1000 * EXP_TAB : __double2loint(t);
    const double kd = t - SHIFT;
    const double f = fma(a, b_l2e, -kd);  // exact product minus an integer: one rounding
    double Ts;  // tab_lane: shared-window byte address of this lane's copy of entry 0 (entries are EXP_REP * 8 bytes apart)
    asm("{\n\t.reg .b32 j, ad;\n\tand.b32 j, %1, %4;\n\tmad.lo.u32 ad, j, %3, %2;\n\tld.shared.f64 %0, [ad];\n\t}"
        : "=d"(Ts)
        : "r"(k), "r"(tab_lane), "n"(EXP_REP * 8), "n"(EXP_TAB - 1));
    const double T = __hiloint2double(__double2hiint(Ts) + (k << (20 - EXP_LOG2)), __double2loint(Ts));  // 2^(k/128)
    // Horner coefficients from the constant bank: as literals each block of ten exp's would re-materialise them
    // into uniform registers (10 UMOV per block)
    double h = fma(f, cst.expc[EXP_DEG], cst.expc[EXP_DEG - 1]);
#pragma unroll
    for (int d = EXP_DEG - 2; d >= 0; --d) h = fma(f, h, cst.expc[d]);
    const double u = T * f;
    return fma(u, h, T);
}

constexpr double L2E64 = EXP_L2E;  // EXP_TAB / ln2  (name kept: "scaled log2(e)")

// a / d for normal, finite d: hardware reciprocal seed (>= 20 bits) + one Newton step (40 bits) + one
// residual correction of the quotient (<= 1 ulp).  5 FP64-pipe instructions, no special-case branches
// (the IEEE division routine costs ~45 instructions with its slow-path checks).
__device__ __forceinline__ double div_fast(double a, double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    const double q = a * r;
    return fma(fma(-d, q, a), r, q);
}

// Upper clamp of a tau in front of the transmissions (exp(-tau_clamp / mu) ~ 1e-100 for every mu) as ONE integer minimum on
// the high word - for non-negative doubles the order of the high words is the order of the numbers, negative ones stay as
// they are - instead of DSETP + two FSEL; the result may exceed the clamp by less than 2^-20 of it, irrelevant here.
__device__ __forceinline__ double clamp_hi(double v, int hi_max) {
    return __hiloint2double(min(__double2hiint(v), hi_max), __double2loint(v));
}

// descending compare-exchange
__device__ __forceinline__ void cex(double& a, double& b) {
    const double hi = fmax(a, b), lo = fmin(a, b);
    a = hi;
    b = lo;
}

// LowerPos (repwvl_thermal.cpp:19-45) on the perturbed temperatures of one pressure node.
__device__ __forceinline__ int lowerpos_t(double tref, double x, int n) {
    auto sgn = [](double v) { return (0.0 < v) - (v < 0.0); };
    int prev = sgn((tref + cst.t_pert[0]) - x);
    int res = n - 2;
    bool done = false;
    for (int k = 1; k < n; ++k) {
        const int cur = sgn((tref + cst.t_pert[k]) - x);
        if (!done && cur != prev) {
            res = k - 1;
            done = true;
        }
        prev = cur;
    }
    return res;
}

