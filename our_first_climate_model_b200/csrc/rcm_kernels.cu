// Hand-written sm_100a kernels of the radiative-convective column solver.
//
// One fused kernel does a complete reference time step (main.cpp:531-583) per column:
//   K5a  theta-sort adjustment + water-vapour feedback + table indices     (main.cpp:536-540, :281-289,
//                                                                          repwvl_thermal.cpp:226-240)
//   K1   optical depth tau[lambda][layer]                                  (repwvl_thermal.cpp:197-248,
//                                                                          main.cpp:266-274)
//   K2   Planck source per wavelength and layer                            (main.cpp:186-204)
//   K3   30-angle Schwarzschild down/up recurrences                        (main.cpp:291-318)
//   K4   spectral + angular reduction to E_down, E_up, heating dE          (main.cpp:326-341)
//   K5b  adaptive time step, temperature update, surface temperature       (main.cpp:156-176)
//
// Mapping.  A CTA owns a tile of C consecutive columns for all fused steps.  Its threads are
// C columns x 2 halves x G wavelength groups: the lane pair (c, g) walks wavelengths g, g+G, ... of
// column c, one lane per half of the atmosphere, so lanes of a warp hold the same wavelength for
// consecutive columns (table rows and Planck constants are warp-uniform, per-column scalars come from
// shared memory without bank conflicts).  The vertical problem of one (column, wavelength, half) lives in
// registers: tau[10], source differences[10], two sets of transmissions[10], 20 flux accumulators.
// The path is FP64-pipe bound: no tensor cores, HBM traffic ~1.6 KB per column-step.
#include <cfloat>
#include <cstdio>

#include "rcm_kernels.cuh"

__constant__ DevConst cst;

cudaError_t rcm_upload_const(const DevConst& c) { return cudaMemcpyToSymbol(cst, &c, sizeof(DevConst)); }

namespace {

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

// ------------------------------------------------------------------------------------------
// K3 + K4 for one (column, wavelength, half): from the optical depths tau[j] and the Planck source
// Bo[j] of the ten owned layers (Bs: surface source) accumulate the fluxes over all angles.
// Written for the deviation of the radiance from the source of the NEXT layer,
//   down: N_{lev+1} = L_{lev+1} - B_{lev+1} = t_lev N_lev + (B_lev - B_{lev+1}),    N_0  = -B_0
//   up:   V_lev     = U_lev - B_{lev-1}     = t_lev V_{lev+1} + (B_lev - B_{lev-1}), V_20 = B_s - B_19
// which is the reference's L = (1-alpha) L + alpha B, alpha = 1 - t (main.cpp:307/312), at one FMA per
// layer and sweep; the angle-independent parts sum_mu cmu*B (and main.cpp:302) are added up front.
// Lane h=0 runs the down sweep through its layers 0..9 while lane h=1 runs the up sweep through 19..10;
// they swap the radiance at level 10 and each finishes the other's sweep through its own layers.  Both
// lanes execute identical code.
// The angle loop is software-pipelined by hand: while the two dependent 10-step recurrences of one angle
// run (latency-bound on their own), the ten independent transmissions of the next chain head are evaluated
// in the same basic block, so a warp always has independent FP64 work in flight.  Two register sets
// ping-pong (loop over chains unrolled by two); cubes are taken in place.
// ------------------------------------------------------------------------------------------
template <bool CLAMPK>
__device__ __forceinline__ void sweep_item(const double (&tau)[HALF], const double (&Bo)[HALF], double Bs, int h,
                                           unsigned tab_lane, double (&E1)[HALF], double (&E2)[HALF],
                                           double& Eu20) {
    double D1[HALF], Dx, X0;
    {
        const double Bnb = __shfl_xor_sync(0xffffffffu, Bo[HALF - 1], 1);  // partner's boundary layer
        const double cs = cst.csum;
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const double Bnext = (j < HALF - 1) ? Bo[j + 1] : Bnb;
            D1[j] = Bo[j] - Bnext;
            E1[j] = fma(cs, Bnext, E1[j]);
            if (j > 0) E2[j] = fma(cs, Bo[j - 1], E2[j]);
        }
        Dx = Bo[0];
        const double Bstart = h ? Bs : 0.0;  // down sweep starts with L=0, up sweep with B(T_surface)
        X0 = Bstart - Bo[0];
        Eu20 = fma(cs, Bstart, Eu20);  // main.cpp:302 summed over the angles (h=1 only)
    }
    // both sweeps of one angle with the transmissions tc
    auto sweep = [&](const double (&tc)[HALF], double cm) {
        double X = X0;
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            X = fma(tc[j], X, D1[j]);
            E1[j] = fma(cm, X, E1[j]);
        }
        double Y = __shfl_xor_sync(0xffffffffu, X, 1);
#pragma unroll
        for (int j = HALF - 1; j >= 1; --j) {
            Y = fma(tc[j], Y, -D1[j - 1]);
            E2[j] = fma(cm, Y, E2[j]);
        }
        Y = fma(tc[0], Y, Dx);
        E2[0] = fma(cm, Y, E2[0]);
    };
    // One chain of angles mu, mu/3, mu/9, ...: the head's transmissions tc were produced during the previous
    // chain; every further level is the cube of the one before (in place).  While the last level is swept,
    // the transmissions of the NEXT unit's head (or virtual root) are evaluated into tn (ten independent exp's that
    // fill the issue slots the two dependent recurrences leave empty).
    int slot = 0;
    auto chain = [&](double (&tc)[HALF], double (&tn)[HALF], int len, double nim) {
#pragma unroll 1
        for (int k = 1; k < len; ++k) {
            sweep(tc, cst.cmu[slot++]);
#pragma unroll
            for (int j = 0; j < HALF; ++j) tc[j] = tc[j] * tc[j] * tc[j];
        }
#pragma unroll
        for (int j = 0; j < HALF; ++j) tn[j] = exp_scaled<CLAMPK>(tau[j], nim, tab_lane);
        sweep(tc, cst.cmu[slot++]);
    };
    const int nchain = cst.nchain;  // even (a zero-weight exp(0) chain pads an odd count)
    const int npair = cst.npair;
    double tA[HALF], tB[HALF];
    {
        const double nim = npair ? cst.pair_nim[0] : cst.neg_inv_mu_l2e[0];
#pragma unroll
        for (int j = 0; j < HALF; ++j) tA[j] = exp_scaled<CLAMPK>(tau[j], nim, tab_lane);
    }
    // Pair units: tA holds x = t(R) of a virtual node R shared by two chain heads a > b (pa * a = pb * b = R):
    // t(a) = x^pa into tA, t(b) = x^pb into tB by 3-4 multiplications, then the two chains; the second one evaluates
    // the next unit's root into tA again.
#pragma unroll 1
    for (int p = 0; p < npair; ++p) {
        const int type = cst.pair_type[p];
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            tB[j] = tA[j] * tA[j];  // x^2
            tA[j] = tA[j] * tB[j];  // x^3
        }
        if (type == 1) {
#pragma unroll
            for (int j = 0; j < HALF; ++j) tA[j] = tA[j] * tB[j];  // x^5
        } else if (type == 2) {
#pragma unroll
            for (int j = 0; j < HALF; ++j) tB[j] = tB[j] * tB[j];  // x^4
        }
#pragma unroll
        for (int j = 0; j < HALF; ++j) tB[j] = tB[j] * tA[j];  // x^5 (type 0), x^7 (types 1, 2)
        const int lenA = cst.pair_lenA[p];
        int lenB = cst.pair_lenB[p];
#pragma unroll 1
        for (int k = 1; k < lenA; ++k) {
            sweep(tA, cst.cmu[slot++]);
#pragma unroll
            for (int j = 0; j < HALF; ++j) tA[j] = tA[j] * tA[j] * tA[j];
        }
        if (lenB > 1) {
            // the last angle of the first chain and the first one of the second in ONE block: four independent
            // recurrences (a sweep on its own is latency-bound: 2.2x the time per instruction of the other blocks)
            sweep(tA, cst.cmu[slot]);
            sweep(tB, cst.cmu[slot + 1]);
            slot += 2;
#pragma unroll
            for (int j = 0; j < HALF; ++j) tB[j] = tB[j] * tB[j] * tB[j];
            --lenB;
        } else {
            sweep(tA, cst.cmu[slot++]);
        }
        chain(tB, tA, lenB, cst.pair_nim[p + 1]);
    }
    for (int ic = 0; ic < nchain; ic += 2) {
        chain(tA, tB, cst.chain_len[ic], cst.neg_inv_mu_l2e[ic + 1]);
        chain(tB, tA, cst.chain_len[ic + 1], cst.neg_inv_mu_l2e[ic + 2]);
    }
}

// Layer split.  The two lanes of a pair share one (column, wavelength): lane h=0 owns layers 0..9
// top-down, lane h=1 owns layers 19..10 (bottom-up), both as local index j=0..9.  Per-layer
// arrays are stored in this order: row(l) = l for l<10, 29-l otherwise (= 10*h + j).
__device__ __forceinline__ constexpr int prow(int l) { return l < HALF ? l : 29 - l; }

constexpr int ROWB = 5 * 32;             // bytes of one table row: 5 active species x {c0, cT, cP, cPT}
constexpr int NCAND = 3;                 // candidate rows per layer: temperature intervals it_min .. it_min + 2 of the tile
constexpr int ROWBUF = NLAY * NCAND * ROWB;  // per warp: 60 rows, 9600 bytes

// 16-byte asynchronous global -> shared copies (LDGSTS) for the row staging.  (cp.async.bulk was tried first: its
// operands live in uniform registers, so 40 per-lane row copies became a 40-trip ELECT/R2UR/UBLKCP waterfall.)
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// STAGE (16-column tiles, the five default species): every warp owns a 9600-byte buffer that receives, by
// cp.async while the previous wavelength's angles run, the table rows its next wavelength needs; after the
// wavelength loop the first 5376 bytes carry the warp's flux partials into the spectral reduction and the tails of
// the first three buffers hold the tile's reduced fluxes (Ed, Eu, dE).
template <int C, int NT, int NACT>
struct Smem {
    static constexpr bool STAGE = (C == 16 && NACT == 5);
    static constexpr int G = NT / (2 * C), NW = NT / 32;
    static constexpr size_t PART = 2 * (size_t)NLEV * C;  // doubles of one wavelength group's partial fluxes [42][C]
    static constexpr size_t EP_BYTES = STAGE ? (size_t)NW * ROWBUF : (size_t)G * PART * sizeof(double);
    static_assert(!STAGE || (G == NW && NW >= 3 && (PART + NLEV * C) * sizeof(double) <= (size_t)ROWBUF), "one warp per wavelength group");
    double* exp_tab;  // [EXP_TAB][EXP_REP]
    double* plk;      // [2][PLK_MAX] Planck factors per wavelength (tables of up to PLK_MAX wavelengths)
    double* T;        // [20][C] layer temperature (sorted), rows in pair order
    double* invT;     // [20][C]
    double* delT;     // [20][C]
    double* vmr;      // [nactive][20][C]
    double* Ts;       // [C]
    double* invTs;    // [C]
    double* dt;       // [C]
    double* solar;    // [C] absorbed solar irradiance of the column
    double* cloudc;   // [C] grey-cloud tau of the column
    double* Ed;       // [21][C]   natural level order
    double* Eu;       // [21][C]
    double* dE;       // [20][C]   natural layer order
    unsigned char* ep;  // partial fluxes of group gg at ep + gg * ep_stride (STAGE: = that warp's row buffer)
    int* it;          // [20][C]
    int* rowsel;      // [20][C] byte offset of the (layer, column)'s row inside the warp's row buffer
    int* rowoff;      // [20][NCAND] first row (cell * nwvl) of the candidates of every layer (pair order)
    int* itmin;       // [20]
    int* outside;     // [20] then [10]: some column of the tile needs a row beyond the two candidates
    static constexpr size_t ep_stride = STAGE ? (size_t)ROWBUF : PART * sizeof(double);
    static size_t bytes(int nactive) {
        return ((size_t)EXP_TAB * EXP_REP + 2 * PLK_MAX + 3 * (size_t)NLAY * C + (size_t)nactive * NLAY * C + 5 * (size_t)C +
                (STAGE ? 0 : 2 * (size_t)NLEV * C + (size_t)NLAY * C)) * sizeof(double) + EP_BYTES +
               (2 * (size_t)NLAY * C + (NCAND + 3) * NLAY + HALF + 2) * sizeof(int);
    }
    __device__ __forceinline__ Smem(unsigned char* base, int nactive) {
        double* p = reinterpret_cast<double*>(base);
        exp_tab = p; p += EXP_TAB * EXP_REP;
        ep = reinterpret_cast<unsigned char*>(p); p += EP_BYTES / sizeof(double);  // 128-byte aligned: 8 KB into the block
        T = p;       p += NLAY * C;
        invT = p;    p += NLAY * C;
        delT = p;    p += NLAY * C;
        vmr = p;     p += nactive * NLAY * C;
        plk = p;     p += 2 * PLK_MAX;
        Ts = p;      p += C;
        invTs = p;   p += C;
        dt = p;      p += C;
        solar = p;   p += C;
        cloudc = p;  p += C;
        if (STAGE) {  // tails of the row buffers (free while the partials are reduced and until the next request)
            Ed = reinterpret_cast<double*>(ep + 0 * ep_stride) + PART;
            Eu = reinterpret_cast<double*>(ep + 1 * ep_stride) + PART;
            dE = reinterpret_cast<double*>(ep + 2 * ep_stride) + PART;
        } else {
            Ed = p;      p += NLEV * C;
            Eu = p;      p += NLEV * C;
            dE = p;      p += NLAY * C;
        }
        it = reinterpret_cast<int*>(p);
        rowsel = it + NLAY * C;
        rowoff = rowsel + NLAY * C;
        itmin = rowoff + NCAND * NLAY;
        outside = itmin + NLAY;
    }
};

// Table indices and interpolation weights in T for every (layer, column) of the tile, from the
// temperatures currently in s.T (repwvl_thermal.cpp:229-239).
template <int C, int NT, int NACT>
__device__ __forceinline__ void prep_tau_indices(const Smem<C, NT, NACT>& s, int tid) {
    for (int i = tid; i < NLAY * C; i += NT) {
        const int r = i / C;  // pair-order row; tref_ip is stored in the same order
        const double midT = s.T[i];
        const double tref = cst.tref_ip[r];
        const int it = lowerpos_t(tref, midT, cst.n_tpert);
        const double t0 = tref + cst.t_pert[it], t1 = tref + cst.t_pert[it + 1];
        s.it[i] = it;
        s.delT[i] = (midT - t0) / (t1 - t0);
    }
}

constexpr int min_ctas(int NT) { return (NT == 256 || NT == 512) ? 512 / NT : 384 / NT; }  // 168 registers per thread

template <int MODE, int NACT, int C, int NT, bool CLAMPK>
__global__ void __launch_bounds__(NT, min_ctas(NT)) rcm_step_kernel(const StepArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int G = NT / (2 * C);  // wavelength groups
    const int tid = threadIdx.x, lane = tid & 31;
    const int h = tid & 1, q = tid >> 1, c = q % C, g = q / C;
    const int nact = (NACT > 0) ? NACT : cst.nactive;
    using SM = Smem<C, NT, NACT>;
    const SM s(smem_raw, nact);
    const int sb = h * HALF * C + c;  // this thread's row block in the per-layer arrays

    for (int i = tid; i < EXP_TAB * EXP_REP; i += NT) s.exp_tab[i] = a.exp_tab[i / EXP_REP];
    const unsigned tab_lane = (unsigned)__cvta_generic_to_shared(s.exp_tab + (lane & (EXP_REP - 1)));
    const int nwvl = cst.nwvl;
    // Planck factors of a repwvl-sized table live in shared memory (the per-wavelength global loads sat on the long
    // scoreboard in front of K2); bigger spectral grids (rcm_set_spectral_grid) read them from global memory
    const bool plk_smem = nwvl <= PLK_MAX;
    if (plk_smem)
        for (int i = tid; i < nwvl; i += NT) {
            s.plk[i] = a.planck_c[i];
            s.plk[PLK_MAX + i] = a.planck_k[i];
        }
    // row staging (STAGE): this warp's buffer
    const bool stage = SM::STAGE && MODE == MODE_STEP && a.stage_rows;
    const int warp = tid >> 5;
    unsigned char* const rows = s.ep + (size_t)warp * SM::ep_stride;
    const unsigned rows_addr = (unsigned)__cvta_generic_to_shared(rows);
    // Start the copies of the 60 rows of wavelength w (three candidates per layer, 160 bytes each): lane q copies
    // rows q and q + 32 in 16-byte pieces.  The buffer must be free: all lanes have consumed the previous fill.
    auto request_rows = [&](int w) {
        __syncwarp();
        const char* base = reinterpret_cast<const char*>(a.coef) + (size_t)w * ROWB;
        for (int row = lane; row < NCAND * NLAY; row += 32) {
            const char* src = base + (size_t)s.rowoff[row] * ROWB;
            const unsigned dst = rows_addr + row * ROWB;
#pragma unroll
            for (int part = 0; part < ROWB / 16; ++part) cp_async16(dst + part * 16, src + part * 16);
        }
        cp_async_commit();
    };
    auto wait_rows = [&] {
        cp_async_wait_all();
        __syncwarp();
    };

    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const int col0 = tile * C;
        const int ncl = min(C, a.ncol - col0);  // columns really present in this tile
        const bool live = c < ncl;
        __syncthreads();
        // ---- load the tile's state: T [20][C], surface T, active VMRs -----------------------
        for (int i = tid; i < NLAY * C; i += NT) {
            const int l = i / C, cc = i % C;
            s.T[prow(l) * C + cc] = a.Tlayer[(size_t)(col0 + (cc < ncl ? cc : 0)) * NLAY + l];  // padding = column 0
        }
        for (int i = tid; i < nact * NLAY * C; i += NT) {
            const int cc = i % C, l = (i / C) % NLAY, sp = i / (C * NLAY);
            s.vmr[(sp * NLAY + prow(l)) * C + cc] =
                (cc < ncl) ? a.vmr[((size_t)(col0 + cc) * nact + sp) * NLAY + l] : 0.0;
        }
        if (tid < C) {
            const int cc = col0 + (tid < ncl ? tid : 0);
            s.Ts[tid] = (tid < ncl) ? a.Tsurf[col0 + tid] : 250.0;
            s.solar[tid] = a.solar_col ? a.solar_col[cc] : cst.solar_irr;
            s.cloudc[tid] = a.cloud_col ? a.cloud_col[cc] : cst.cloud_tau;
        }
        __syncthreads();

        // K1 for one owned layer j (local index) and wavelength w: bilinear (p,T) interpolation of the cross
        // sections in the reference's operation order, no FMA contraction -> tau is bit-identical to
        // read_tau's for identical inputs.  The four bilinear coefficients c0, cT, cP, cPT
        // (repwvl_thermal.cpp:235-238) depend on the table alone and are precomputed per cell (rcm_coef_kernel).
        auto tau_from = [&](int j, const double2* cf, double cl) -> double {
            const int r = h * HALF + j;
            const double dT = s.delT[sb + j * C], dP = cst.delP[r];
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < (NACT > 0 ? NACT : RCM_NSPECIES); ++k) {
                if (NACT == 0 && k >= nact) break;
                // two 128-bit loads (one 256-bit LDG.E.ENL2.256 was measured 8% slower for the whole step)
                const double2 c0T = cf[2 * k], cPPT = cf[2 * k + 1];
                double v = __dadd_rn(c0T.x, __dmul_rn(c0T.y, dT));
                v = __dadd_rn(v, cPPT.x);  // cP * delP of this layer, rounded once when the table was built
                v = __dadd_rn(v, __dmul_rn(__dmul_rn(cPPT.y, dT), dP));
                acc = __dadd_rn(acc, __dmul_rn(v, s.vmr[k * NLAY * C + sb + j * C]));
            }
            acc = __dmul_rn(acc, cst.numDens[r]);
            if (cst.cloud_row == r) acc = __dadd_rn(acc, cl);  // main.cpp:270, cl: the column's cloud tau
            return acc;
        };
        // ... with the coefficients read from the table in global memory
        auto tau_cell = [&](int j, int w, double cl) -> double {
            const int cell = cst.ipcell[h * HALF + j] + s.it[sb + j * C];
            return tau_from(j, reinterpret_cast<const double2*>(a.coef) + (size_t)(cell * nwvl + w) * 2 * nact, cl);
        };
        // ... or from the rows staged in this warp's buffer
        auto tau_staged = [&](int j, double cl) -> double {
            return tau_from(j, reinterpret_cast<const double2*>(rows + s.rowsel[sb + j * C]), cl);
        };
        // tau of owned layer j at wavelength w as the transmissions will use it (w is clamped by the caller)
        auto tau_use = [&](int j, int w, double cl) -> double {
            double v;
            if (MODE == MODE_RT) {
                const int l = h ? (NLAY - 1 - j) : j;
                v = live ? a.tau_io[((size_t)(col0 + c) * nwvl + w) * NLAY + l] : 0.0;
            } else {
                v = tau_cell(j, w, cl);
            }
            if (!CLAMPK) v = fmin(v, a.tau_clamp);  // exp(-tau_clamp/mu) ~ 1e-100: same fluxes, see exp_scaled
            return v;
        };

        for (int step = 0; step < a.nsteps; ++step) {
            const bool first = (MODE == MODE_STEP) && (a.step_index + step == 0);
            // ---------------- K5a: adjustment, feedback, table indices ------------------------
            if (MODE == MODE_STEP) {
                if (first) {  // tau of the initial profile is built BEFORE the first sort (main.cpp:500-504)
                    prep_tau_indices(s, tid);
                    __syncthreads();
                }
                // theta-sort (main.cpp:536-540) by ranking, all threads: element (layer l, column c) goes to layer
                // rank = #{l' : theta[l'] > theta[l], or equal and l' < l} (descending; any correct sort gives the
                // reference's values).  One thread per column running a sorting network took 4 % of all warp-time - the
                // other three warps of the CTA waiting at the barrier behind it.  s.invT (rebuilt below) holds theta,
                // s.dE (rebuilt by K4) the change against the previous sorted profile.
                for (int i = tid; i < NLAY * C; i += NT) {
                    const int r = i / C;
                    s.invT[i] = s.T[i] * cst.conv[r < HALF ? r : 29 - r];
                }
                __syncthreads();
                for (int i = tid; i < NLAY * C; i += NT) {
                    const int r = i / C, cc = i % C, l = r < HALF ? r : 29 - r;
                    const double my = s.invT[i];
                    int rank = 0;
#pragma unroll
                    for (int l2 = 0; l2 < NLAY; ++l2) {
                        const double v = s.invT[prow(l2) * C + cc];
                        rank += (v > my || (v == my && l2 < l)) ? 1 : 0;
                    }
                    const double Tn = my / cst.conv[rank];
                    s.T[prow(rank) * C + cc] = Tn;
                    double d = 0.0;
                    if (cc < ncl) {
                        const size_t gi = (size_t)(col0 + cc) * NLAY + rank;
                        d = fabs(Tn - a.Tprev[gi]);
                        a.Tprev[gi] = Tn;
                    }
                    s.dE[rank * C + cc] = d;
                }
                __syncthreads();
                if (tid < C) {
                    double dmax = 0.0;
#pragma unroll
                    for (int l = 0; l < NLAY; ++l) dmax = fmax(dmax, s.dE[l * C + tid]);
                    s.dt[tid] = dmax;  // parked here until the diagnostics are written
                }
                __syncthreads();
                if (!first) {
                    // water_vapor_feedback (main.cpp:281-289) then indices from the sorted profile
                    if (a.h2o_slot >= 0) {
                        for (int i = tid; i < NLAY * C; i += NT) {
                            const int l = i / C, cc = i % C;
                            if (cc < ncl) {
                                const int r = prow(l) * C + cc;
                                const double Tc = s.T[r] - 273.15;
                                const double e_sat = 6.1094 * exp(17.625 * Tc / (Tc + 243.04));
                                const double rh = a.rel_hum[(size_t)(col0 + cc) * NLAY + l];
                                s.vmr[a.h2o_slot * NLAY * C + r] = rh * e_sat / cst.player[l];
                            }
                        }
                    }
                    prep_tau_indices(s, tid);
                }
            } else if (MODE == MODE_TAU) {
                prep_tau_indices(s, tid);
            }
            for (int i = tid; i < NLAY * C; i += NT) s.invT[i] = 1.0 / s.T[i];
            if (tid < C) s.invTs[tid] = 1.0 / s.Ts[tid];
            __syncthreads();
            if (stage) {
                // the candidate rows of every layer: temperature intervals it_min .. it_min + 2 of the tile's columns
                if (tid < NLAY) {
                    int mn = s.it[tid * C], mx = mn;
                    for (int cc = 1; cc < C; ++cc) {
                        mn = min(mn, s.it[tid * C + cc]);
                        mx = max(mx, s.it[tid * C + cc]);
                    }
                    s.itmin[tid] = mn;
                    s.outside[tid] = (mx - mn >= NCAND);
                    for (int k = 0; k < NCAND; ++k)
                        s.rowoff[NCAND * tid + k] = (cst.ipcell[tid] + min(mn + k, cst.n_tpert - 2)) * nwvl;
                }
                __syncthreads();
                for (int i = tid; i < NLAY * C; i += NT) {
                    const int r = i / C;
                    s.rowsel[i] = (NCAND * r + min(s.it[i] - s.itmin[r], NCAND - 1)) * ROWB;
                }
                if (tid == 0) {
                    int any = 0;
                    for (int r = 0; r < NLAY; ++r) any |= s.outside[r];
                    s.outside[NLAY] = any;
                }
                __syncthreads();
            }

            if (MODE == MODE_TAU) {  // K1 alone: the compute part of read_tau + cloud_into_tau
                if (a.lowpos_t) {
                    for (int i = tid; i < NLAY * C; i += NT) {
                        const int l = i / C, cc = i % C;
                        if (cc < ncl)
                            a.lowpos_t[(size_t)(col0 + cc) * NLAY + (NLAY - 1 - l)] = s.it[prow(l) * C + cc];
                    }
                }
                const double cl = s.cloudc[c];
                for (int w = g; w < nwvl; w += G) {
#pragma unroll
                    for (int j = 0; j < HALF; ++j) {
                        const double t = tau_cell(j, w, cl);
                        const int l = h ? (NLAY - 1 - j) : j;
                        if (live) a.tau_io[((size_t)(col0 + c) * nwvl + w) * NLAY + l] = t;
                    }
                }
                continue;
            }

            // ---------------- K1-K4: per (column, wavelength, half) work in registers -----------
            // E1[j]: flux of the first sweep  (h=0: E_down[j+1],   h=1: E_up[19-j])
            // E2[j]: flux of the second sweep (h=0: E_up[j],       h=1: E_down[20-j])
            double E1[HALF], E2[HALF], Eu20 = 0.0;
#pragma unroll
            for (int j = 0; j < HALF; ++j) E1[j] = E2[j] = 0.0;
            // Every thread runs the same number of wavelength items, so the loop and the shuffles inside are
            // provably warp-uniform: a thread whose last item does not exist (w >= nwvl) repeats the last
            // wavelength with a zero Planck factor, which adds exactly 0 to every flux.
            const int nitem = (nwvl + G - 1) / G;
            if (stage) request_rows(min(g, nwvl - 1));
#pragma unroll 1
            for (int item = 0; item < nitem; ++item) {
                const int w_any = g + item * G;
                const bool real = w_any < nwvl;
                const int w = real ? w_any : nwvl - 1;
                double tau[HALF], Bo[HALF];
                const double cl = (MODE == MODE_RT) ? 0.0 : s.cloudc[c];  // read per item: not live across the angle loop
                if (stage) wait_rows();  // the rows of this wavelength were requested one wavelength ago
                // two straight-line versions of K1 (a branch per layer would cut the block the loads are scheduled in);
                // a tile where some column needs a row beyond the two candidates takes the global one for every layer
                if (stage && !s.outside[NLAY]) {
#pragma unroll
                    for (int j = 0; j < HALF; ++j) {
                        const double v = tau_staged(j, cl);
                        tau[j] = CLAMPK ? v : fmin(v, a.tau_clamp);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < HALF; ++j) tau[j] = tau_use(j, w, cl);
                }
                // K1 has consumed the buffer: the rows of the NEXT wavelength travel while this one's angles run
                // (nothing is requested after the last one: the buffer then carries the flux partials)
                if (stage && item + 1 < nitem) request_rows(min(w_any + G, nwvl - 1));
                // K2: Planck source B = k_w / (exp(c_w / T) - 1) (main.cpp:188-191 regrouped so that everything
                // that depends on the wavelength alone is precomputed on the host); surface: main.cpp:301
                const double pc = plk_smem ? s.plk[w] : __ldg(a.planck_c + w);
                const double pk = !real ? 0.0 : plk_smem ? s.plk[PLK_MAX + w] : __ldg(a.planck_k + w);
#pragma unroll
                for (int j = 0; j < HALF; ++j)
                    Bo[j] = div_fast(pk, exp_scaled<false>(pc, s.invT[sb + j * C] * L2E64, tab_lane) - 1.0);
                const double Bs = div_fast(pk, exp_scaled<false>(pc, s.invTs[c] * L2E64, tab_lane) - 1.0);
                sweep_item<CLAMPK>(tau, Bo, Bs, h, tab_lane, E1, E2, Eu20);
            }

            // ---------------- K4: reduce the G wavelength groups of every column ---------------
            // every group leaves its partial fluxes in its own buffer [42][C] (row l: E_down[l+1] for l < 20, row 21+l:
            // E_up[l]); the groups are then summed in a fixed order
            {
                double* part = reinterpret_cast<double*>(s.ep + (size_t)g * SM::ep_stride);
#pragma unroll
                for (int j = 0; j < HALF; ++j) {
                    const int l = h ? (NLAY - 1 - j) : j;
                    part[l * C + c] = h ? E2[j] : E1[j];
                    part[(NLEV + l) * C + c] = h ? E1[j] : E2[j];
                }
                if (h) part[(NLEV + NLAY) * C + c] = Eu20;
            }
            __syncthreads();
            for (int i = tid; i < 2 * NLEV * C; i += NT) {
                const int row = i / C;
                if (row == NLAY) {
                    s.Ed[i % C] = 0.0;  // E_down at the top of the atmosphere stays 0 (main.cpp:300)
                    continue;
                }
                double sum = 0.0;
                for (int gg = 0; gg < G; ++gg) sum += reinterpret_cast<const double*>(s.ep + (size_t)gg * SM::ep_stride)[i];
                if (row < NLAY) s.Ed[i + C] = sum; else s.Eu[i - NLEV * C] = sum;
            }
            __syncthreads();
            // heating rates (main.cpp:337-341)
            for (int i = tid; i < NLAY * C; i += NT) {
                const int l = i / C, cc = i % C;
                double d = s.Ed[l * C + cc] - s.Ed[(l + 1) * C + cc] + s.Eu[(l + 1) * C + cc] - s.Eu[l * C + cc];
                if (l == NLAY - 1) d += s.solar[cc] + s.Ed[NLAY * C + cc] - s.Eu[NLAY * C + cc];
                s.dE[i] = d;
            }
            __syncthreads();

            const bool last = (step == a.nsteps - 1);
            if (MODE == MODE_STEP) {
                // ------------- K5b: time step and temperature update (main.cpp:156-176) ---------
                // the column's time step by one thread per column, the update of its 20 layers by all threads
                if (tid < C) {
                    double mx = s.dE[tid], mabs = 0.0;
#pragma unroll
                    for (int l = 0; l < NLAY; ++l) {
                        const double d = s.dE[l * C + tid];
                        if (mx < d) mx = d;
                        mabs = fmax(mabs, fabs(d));
                    }
                    double dt = (double)(float)cst.max_dT / mx * (1004.0 * cst.dp * 100.0) / 9.80665;
                    if (dt > cst.dt_cap) dt = cst.dt_cap;
                    const double dT_stat = s.dt[tid];
                    s.dt[tid] = dt;
                    if (tid < ncl) {
                        const int col = col0 + tid;
                        a.time_h[col] += (float)dt / 3600;  // main.cpp:581
                        if (a.diag) {
                            double* dg = a.diag + ((size_t)step * a.diag_ncol + col) * 4;
                            dg[0] = s.solar[tid] - s.Eu[tid];
                            dg[1] = dT_stat;
                            dg[2] = mabs;
                            dg[3] = dt;
                        }
                    }
                }
                __syncthreads();
                for (int i = tid; i < NLAY * C; i += NT) {
                    const int l = i / C, cc = i % C;
                    const double Tn = s.T[prow(l) * C + cc] + s.dE[i] * s.dt[cc] * 9.80665 / (1004.0 * cst.dp * 100.0);
                    s.T[prow(l) * C + cc] = Tn;
                    if (l == NLAY - 1) s.Ts[cc] = Tn * cst.conv[NLAY - 1];  // main.cpp:173
                }
            }
            __syncthreads();
            if (last) {
                // fluxes of the last step: the tile's block of each output array is contiguous
                for (int i = tid; i < NLEV * ncl; i += NT) {
                    const int cc = i / NLEV, l = i % NLEV;
                    a.E_down[(size_t)col0 * NLEV + i] = s.Ed[l * C + cc];
                    a.E_up[(size_t)col0 * NLEV + i] = s.Eu[l * C + cc];
                }
                for (int i = tid; i < NLAY * ncl; i += NT) {
                    const int cc = i / NLAY, l = i % NLAY;
                    a.dE[(size_t)col0 * NLAY + i] = s.dE[l * C + cc];
                    if (MODE == MODE_STEP) {
                        a.Tlayer[(size_t)col0 * NLAY + i] = s.T[prow(l) * C + cc];
                        if (a.h2o_slot >= 0)
                            a.vmr[((size_t)(col0 + cc) * nact + a.h2o_slot) * NLAY + l] =
                                s.vmr[(a.h2o_slot * NLAY + prow(l)) * C + cc];
                    }
                }
                if (MODE == MODE_STEP && tid < ncl) {
                    a.Tsurf[col0 + tid] = s.Ts[tid];
                    a.dt[col0 + tid] = s.dt[tid];
                }
            }
        }
    }
}

// Bilinear coefficients per LAYER, temperature interval and active species, in the reference's operation order
// (repwvl_thermal.cpp:235-238):  coef[(r*(nt-1)+it)][w][k] = {c0, cT, cP * delP_r, cPT}, r = pair-order layer row.
// The layer's pressure interval ip and its weight delP come from the shared pressure grid (constant bank), so the
// product cP * delP - one rounding in the reference too - is taken here once instead of once per column and step.
// src is the file-order table xsec[it][species][w][ip].
__global__ void rcm_coef_kernel(const double* __restrict__ src, double* __restrict__ dst, int nt, int ns, int nw,
                                int np, int nact, const int* __restrict__ species) {
    const size_t n = (size_t)NLAY * (nt - 1) * nw * nact;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int k = r % nact; r /= nact;
        const int w = r % nw; r /= nw;
        const int it = r % (nt - 1); r /= (nt - 1);
        const int row = (int)r;                             // pair-order layer row
        const int l = row < HALF ? row : 29 - row;          // top-down layer
        const int ip = cst.ip[l];
        const int sp = species[k];
        auto X = [&](int b, int cc) { return src[(((size_t)b * ns + sp) * nw + w) * np + cc]; };
        const double c0 = X(it, ip);
        const double cT = __dsub_rn(X(it + 1, ip), c0);
        const double cP = __dsub_rn(X(it, ip + 1), c0);
        const double cPT = __dsub_rn(__dsub_rn(__dsub_rn(X(it + 1, ip + 1), cP), cT), c0);
        dst[4 * i + 0] = c0;
        dst[4 * i + 1] = cT;
        dst[4 * i + 2] = __dmul_rn(cP, cst.delP[row]);
        dst[4 * i + 3] = cPT;
    }
}

// ------------------------------------------------------------------------------------------
// Per-step ensemble scalars from the per-column diagnostics.  RED_BLOCKS CTAs per step reduce fixed slices
// of the columns with a fixed-order tree into partial[step][block][4]; the CTA that finishes last (ticket
// counter, reset for the next launch) folds the partials in block order.  The result does not depend on
// scheduling.  out[step] = {sum toa, max dT, #converged, max|dE|} (layout of rcm_step_scalars).
// ------------------------------------------------------------------------------------------
constexpr int RED_BLOCKS = 64, RED_THREADS = 256;

__device__ __forceinline__ void red4(double& sum, double& mx, double& cnt, double& mde, double (*sh)[RED_THREADS / 32]) {
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mde = fmax(mde, __shfl_xor_sync(0xffffffffu, mde, o));
    }
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    __syncthreads();
    if (ln == 0) {
        sh[0][w] = sum; sh[1][w] = mx; sh[2][w] = cnt; sh[3][w] = mde;
    }
    __syncthreads();
    constexpr int NW = RED_THREADS / 32;
    sum = ln < NW ? sh[0][ln] : 0.0;
    mx = ln < NW ? sh[1][ln] : 0.0;
    cnt = ln < NW ? sh[2][ln] : 0.0;
    mde = ln < NW ? sh[3][ln] : 0.0;
    for (int o = NW / 2; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mde = fmax(mde, __shfl_xor_sync(0xffffffffu, mde, o));
    }
}

__global__ void __launch_bounds__(RED_THREADS) rcm_reduce_diag_kernel(const double* __restrict__ diag, int ncol,
                                                                      double dT_conv, double* __restrict__ partial,
                                                                      unsigned* __restrict__ ticket,
                                                                      double* __restrict__ out) {
    __shared__ double sh[4][RED_THREADS / 32];
    __shared__ bool last;
    const int step = blockIdx.y, b = blockIdx.x;
    const double* d = diag + (size_t)step * ncol * 4;
    const int per = (ncol + RED_BLOCKS - 1) / RED_BLOCKS, lo = b * per, hi = min(ncol, lo + per);
    double sum = 0.0, mx = 0.0, cnt = 0.0, mde = 0.0;
    for (int i = lo + threadIdx.x; i < hi; i += RED_THREADS) {
        const double4 v = *reinterpret_cast<const double4*>(d + (size_t)i * 4);
        sum += v.x;
        mx = fmax(mx, v.y);
        cnt += (v.y < dT_conv) ? 1.0 : 0.0;
        mde = fmax(mde, v.z);
    }
    red4(sum, mx, cnt, mde, sh);
    if (threadIdx.x == 0) {
        double* pp = partial + ((size_t)step * RED_BLOCKS + b) * 4;
        pp[0] = sum; pp[1] = mx; pp[2] = cnt; pp[3] = mde;
        __threadfence();
        last = (atomicAdd(ticket + step, 1u) == RED_BLOCKS - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    sum = mx = cnt = mde = 0.0;
    if (threadIdx.x < 32) {  // 64 partials, two per lane, folded in block order by the fixed tree
        for (int k = threadIdx.x; k < RED_BLOCKS; k += 32) {
            const double* pp = partial + ((size_t)step * RED_BLOCKS + k) * 4;
            sum += pp[0];
            mx = fmax(mx, pp[1]);
            cnt += pp[2];
            mde = fmax(mde, pp[3]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            mde = fmax(mde, __shfl_xor_sync(0xffffffffu, mde, o));
        }
        if (threadIdx.x == 0) {
            out[step * 4 + 0] = sum; out[step * 4 + 1] = mx; out[step * 4 + 2] = cnt; out[step * 4 + 3] = mde;
            ticket[step] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// FP64-pipe microbenchmarks: the measured denominators of the roofline.  Each
// thread runs `iters` rounds of 8 independent dependency chains.
// ------------------------------------------------------------------------------------------
template <int WHICH>
__global__ void __launch_bounds__(256) rcm_microbench_kernel(double* out, long iters, const double* tab) {
    __shared__ double stab[EXP_TAB * EXP_REP];
    for (int i = threadIdx.x; i < EXP_TAB * EXP_REP; i += blockDim.x) stab[i] = tab[i / EXP_REP];
    __syncthreads();
    const unsigned tl = (unsigned)__cvta_generic_to_shared(stab + (threadIdx.x & (EXP_REP - 1)));
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = -1.0 - 0.001 * (threadIdx.x + k);
    const double a = 0.999999, b = -1e-7;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (WHICH == 0) v[k] = fma(v[k], a, b);
            if (WHICH == 1) v[k] = exp(v[k]) - 1.5;
            if (WHICH == 2) v[k] = -1.0 / v[k] - 1.7;
            if (WHICH == 3) v[k] = exp_scaled<false>(v[k], L2E64, tl) - 1.5;
        }
    }
    double sacc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) sacc += v[k];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = sacc;
}

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
__device__ __forceinline__ double cplkavg_narrow(double wvllo, double wvlhi, double whi, double wlo, double t,
                                                 unsigned tab_lane) {
    const double c2 = 1.438786, sigma = 5.67032E-8, pi = 3.14159265358979323846;
    const double vmax = 709.782712893384, sigdpi = sigma / pi, conc = 15. / (pi * pi * pi * pi);
    const double v0 = div_fast(c2 * wlo, t), v1 = div_fast(c2 * whi, t);
    if (!(t >= 1.e-4 && whi > wlo && wlo >= 0. && v0 > DBL_EPSILON && v1 < vmax && (whi - wlo) / whi < 1.e-2))
        return cplkavg_dev(wvllo, wvlhi, t);
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
    if (fabs((val - prev) / val) <= 1.e-6) return sigdpi * t4 * conc * val;
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
    double* s_T = p;   p += NLAY * C;
    double* s_sH = p;  p += NLAY * C;
    double* s_sO = p;  p += NLAY * C;
    double* s_Ts = p;  p += C;
    double* s_cl = p;  p += C;
    double* s_B = p;   p += HALF * NT;  // [10][NT] Planck source of the thread's ten layers (written by a rolled loop)
    double* s_Ep = p;  // [21][GC]
    const int tid = threadIdx.x, lane = tid & 31;
    const int h = tid & 1, q = tid >> 1, c = q % C, g = q / C;
    const int sb = h * HALF * C + c;
    for (int i = tid; i < EXP_TAB * EXP_REP; i += NT) s_tab[i] = a.exp_tab[i / EXP_REP];
    const unsigned tab_lane = (unsigned)__cvta_generic_to_shared(s_tab + (lane & (EXP_REP - 1)));
    const int tile = blockIdx.x % a.ntiles, chunk = blockIdx.x / a.ntiles;
    const int col0 = tile * C, ncl = min(C, a.ncol - col0);
    for (int i = tid; i < NLAY * C; i += NT) {
        const int l = i / C, cc = i % C, r = prow(l) * C + cc;
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
    const size_t plane = (size_t)a.nwvl * NLAY;
#pragma unroll 1
    for (int item = 0; item < a.chunk_len / G; ++item) {
        const int w_any = w_lo + g + item * G;
        const bool real = w_any < a.nwvl;
        const int w = real ? w_any : a.nwvl - 1;
        double tau[HALF], Bo[HALF];
        const double lo = __ldg(a.wvl_lo + w), hi = __ldg(a.wvl_hi + w);
        const double whi = 1.0E7 / lo, wlo = 1.0E7 / hi;  // cplkavg.cpp:141-142, once per wavelength
        const double* t5 = a.tau5 + (size_t)w * NLAY;
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const int l = h ? (NLAY - 1 - j) : j;
            double v = __dmul_rn(__ldg(t5 + l), s_sH[sb + j * C]);
            v = __dadd_rn(v, __dmul_rn(a.co2_factor, __ldg(t5 + plane + l)));
            v = __dadd_rn(v, __dmul_rn(__ldg(t5 + 2 * plane + l), s_sO[sb + j * C]));
            v = __dadd_rn(v, __ldg(t5 + 3 * plane + l));
            v = __dadd_rn(v, __ldg(t5 + 4 * plane + l));
            if (cst.cloud_row == h * HALF + j) v = __dadd_rn(v, s_cl[c]);
            tau[j] = CLAMPK ? v : fmin(v, a.tau_clamp);
        }
        // The band-integrated Planck function of the ten layers in a ROLLED loop through shared memory: inlined ten
        // times it made the kernel 9,900 instructions long and instruction fetch 6 % of its stalls.
#pragma unroll 1
        for (int j = 0; j < HALF; ++j) s_B[j * NT + tid] = cplkavg_narrow(lo, hi, whi, wlo, s_T[sb + j * C], tab_lane);
#pragma unroll
        for (int j = 0; j < HALF; ++j) {
            const double B = s_B[j * NT + tid];
            Bo[j] = real ? B : 0.0;
        }
        const double Bsurf = cplkavg_narrow(lo, hi, whi, wlo, s_Ts[c], tab_lane);
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

__global__ void __launch_bounds__(128) rcm_lbl_finish_kernel(const LblArgs a) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= a.ncol) return;
    double Ed[NLEV], Eu[NLEV];
#pragma unroll
    for (int l = 0; l < NLEV; ++l) Ed[l] = Eu[l] = 0.0;
    for (int k = 0; k < a.nchunks; ++k) {  // fixed order: deterministic
        const double* pp = a.part + ((size_t)k * a.ncol + col) * 42;
#pragma unroll
        for (int l = 0; l < NLEV; ++l) {
            Ed[l] += pp[l];
            Eu[l] += pp[21 + l];
        }
    }
    const double solar = a.solar_col ? a.solar_col[col] : cst.solar_irr;
    double dE[NLAY], mx = -1e300, mabs = 0.0;
#pragma unroll
    for (int l = 0; l < NLAY; ++l) {
        double d = Ed[l] - Ed[l + 1] + Eu[l + 1] - Eu[l];                      // main.cpp:338
        if (l == NLAY - 1) d += solar + Ed[NLAY] - Eu[NLAY];                   // main.cpp:341
        dE[l] = d;
        if (mx < d) mx = d;
        mabs = fmax(mabs, fabs(d));
    }
    double dt = (double)(float)cst.max_dT / mx * (1004.0 * cst.dp * 100.0) / 9.80665;  // main.cpp:157
    if (dt > cst.dt_cap) dt = cst.dt_cap;
    double Tl = 0.0;
#pragma unroll
    for (int l = 0; l < NLAY; ++l) {
        const size_t gi = (size_t)col * NLAY + l;
        Tl = a.Tlayer[gi] + dE[l] * dt * 9.80665 / (1004.0 * cst.dp * 100.0);  // main.cpp:169
        a.Tlayer[gi] = Tl;
        a.dE[gi] = dE[l];
    }
    a.Tsurf[col] = Tl * cst.conv[NLAY - 1];  // main.cpp:173
    a.dt[col] = dt;
    a.time_h[col] += (float)dt / 3600;
#pragma unroll
    for (int l = 0; l < NLEV; ++l) {
        a.E_down[(size_t)col * NLEV + l] = Ed[l];
        a.E_up[(size_t)col * NLEV + l] = Eu[l];
    }
    if (a.diag) {
        double* dg = a.diag + (size_t)col * 4;
        dg[0] = solar - Eu[0];
        dg[1] = a.dTstat[col];
        dg[2] = mabs;
        dg[3] = dt;
    }
}

// ------------------------------------------------------------------------------------------
// Solar setup per column: doubling_adding + solar_radiative_transfer_setup
// (main.cpp:214-264) with per-column cloud optical depth, zenith cosine and surface albedo.  One thread per
// column, the reference's operation order, no FMA contraction (explicit round-to-nearest intrinsics, IEEE
// division); pow(2, doublings) is an exact power of two, pow(t_dir, 2) the correctly rounded square.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rcm_solar_kernel(const SolarArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const double tau_s = a.tau_s_col ? a.tau_s_col[i] : a.tau_s;
    const double mu_s = a.mu_s_col ? a.mu_s_col[i] : a.mu_s;
    const double albedo = a.albedo_col ? a.albedo_col[i] : a.albedo;
    auto mul = [](double x, double y) { return __dmul_rn(x, y); };
    auto add = [](double x, double y) { return __dadd_rn(x, y); };
    auto sub = [](double x, double y) { return __dsub_rn(x, y); };
    auto dvd = [](double x, double y) { return __ddiv_rn(x, y); };
    const double tau = mul(sub(1.0, a.g_asym), tau_s);                       // main.cpp:216
    const double dtau = dvd(tau, scalbn(1.0, a.doublings));                  // :217
    const double thin = dvd(dtau, mu_s);
    double r = mul(0.5, thin), t = sub(1.0, r);                              // :223-224
    double r_dir = mul(thin, 0.5), s_dir = r_dir, t_dir = sub(1.0, thin);    // :225-227
    for (int k = 0; k < a.doublings; ++k) {                                  // :232-250
        const double denom = sub(1.0, mul(r, r));
        const double r2 = add(r, dvd(mul(mul(r, t), t), denom));
        const double t2 = dvd(mul(t, t), denom);
        const double s2 = add(dvd(add(mul(t, s_dir), mul(mul(mul(t_dir, r_dir), r), t)), denom), mul(t_dir, s_dir));
        const double rd2 = add(dvd(add(mul(mul(t, s_dir), r), mul(mul(t, t_dir), r)), denom), r_dir);
        t_dir = mul(t_dir, t_dir);
        s_dir = s2;
        r_dir = rd2;
        r = r2;
        t = t2;
    }
    // solar_radiative_transfer_setup, main.cpp:258-260
    const double r_total = add(r_dir, mul(mul(dvd(add(t_dir, s_dir), sub(1.0, mul(albedo, r))), t), albedo));
    if (a.r_total) a.r_total[i] = r_total;
    if (a.solar_irr) a.solar_irr[i] = mul(mul(mul(a.daytime, a.E_0), mu_s), sub(1.0, r_total));
    if (a.cloud_tau) a.cloud_tau[i] = dvd(tau_s, 2.0);                       // main.cpp:267
}

__global__ void __launch_bounds__(256) rcm_cplkavg_kernel(int n, const double* lo, const double* hi, const double* t,
                                                          double* out, const double* tab, int narrow) {
    __shared__ double stab[EXP_TAB * EXP_REP];
    for (int i = threadIdx.x; i < EXP_TAB * EXP_REP; i += blockDim.x) stab[i] = tab[i / EXP_REP];
    __syncthreads();
    const unsigned tl = (unsigned)__cvta_generic_to_shared(stab + (threadIdx.x & (EXP_REP - 1)));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        out[i] = narrow ? cplkavg_narrow(lo[i], hi[i], 1.0E7 / lo[i], 1.0E7 / hi[i], t[i], tl) : cplkavg_dev(lo[i], hi[i], t[i]);
}

}  // namespace

size_t rcm_step_smem_bytes(int C, int nactive, int nthreads) {
#define RCM_SHAPE(CC, TT) \
    if (C == CC && nthreads == TT) return nactive == 5 ? Smem<CC, TT, 5>::bytes(5) : Smem<CC, TT, 0>::bytes(nactive);
    RCM_SHAPE(16, 128)
    RCM_SHAPE(8, 128)
    RCM_SHAPE(4, 128)
    RCM_SHAPE(32, 192)
    RCM_SHAPE(16, 96)
#undef RCM_SHAPE
    return 0;
}

template <int MODE, int NACT, int C, int NT>
static cudaError_t launch_t(const StepArgs& a, int nactive, int grid, cudaStream_t st) {
    const size_t smem = Smem<C, NT, NACT>::bytes(nactive);
    auto kern = a.clampk ? rcm_step_kernel<MODE, NACT, C, NT, true> : rcm_step_kernel<MODE, NACT, C, NT, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, NT, smem, st>>>(a);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_m(const StepArgs& a, int nactive, int grid, cudaStream_t st) {
    const bool five = (nactive == 5);
#define RCM_SHAPE(CC, TT)                  \
    if (a.C == CC && a.nthreads == TT)     \
        return five ? launch_t<MODE, 5, CC, TT>(a, nactive, grid, st) : launch_t<MODE, 0, CC, TT>(a, nactive, grid, st);
    RCM_SHAPE(16, 128)
    RCM_SHAPE(8, 128)
    RCM_SHAPE(4, 128)
    RCM_SHAPE(32, 192)
    RCM_SHAPE(16, 96)
#undef RCM_SHAPE
    return cudaErrorInvalidValue;
}

cudaError_t rcm_launch_step(int mode, const StepArgs& a, int nactive, int grid, cudaStream_t st) {
    switch (mode) {
        case MODE_STEP: return launch_m<MODE_STEP>(a, nactive, grid, st);
        case MODE_TAU: return launch_m<MODE_TAU>(a, nactive, grid, st);
        case MODE_RT: return launch_m<MODE_RT>(a, nactive, grid, st);
    }
    return cudaErrorInvalidValue;
}

size_t rcm_reduce_scratch_doubles(int nsteps) { return (size_t)nsteps * (RED_BLOCKS * 4 + 1); }

// scratch: rcm_reduce_scratch_doubles(nsteps) doubles, zeroed once when allocated (the tickets live at its end)
cudaError_t rcm_launch_reduce_diag(const double* diag, int nsteps, int ncol, double dT_converged, double* scratch,
                                   double* scalars, cudaStream_t st) {
    unsigned* ticket = reinterpret_cast<unsigned*>(scratch + (size_t)nsteps * RED_BLOCKS * 4);
    rcm_reduce_diag_kernel<<<dim3(RED_BLOCKS, nsteps), RED_THREADS, 0, st>>>(diag, ncol, dT_converged, scratch, ticket, scalars);
    return cudaGetLastError();
}

cudaError_t rcm_launch_coef(const double* xsec_file, double* coef, int nt, int ns, int nw, int np, int nact,
                            const int* d_species, cudaStream_t st) {
    rcm_coef_kernel<<<296, 256, 0, st>>>(xsec_file, coef, nt, ns, nw, np, nact, d_species);
    return cudaGetLastError();
}

cudaError_t rcm_launch_microbench(int which, double* out, const double* tab, long iters, int grid,
                                  cudaStream_t st) {
    switch (which) {
        case 0: rcm_microbench_kernel<0><<<grid, 256, 0, st>>>(out, iters, tab); break;
        case 1: rcm_microbench_kernel<1><<<grid, 256, 0, st>>>(out, iters, tab); break;
        case 2: rcm_microbench_kernel<2><<<grid, 256, 0, st>>>(out, iters, tab); break;
        case 3: rcm_microbench_kernel<3><<<grid, 256, 0, st>>>(out, iters, tab); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t rcm_launch_solar(const SolarArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    rcm_solar_kernel<<<(a.n + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}

cudaError_t rcm_launch_cplkavg(int n, const double* lo, const double* hi, const double* t, double* out,
                               const double* exp_tab, int narrow, cudaStream_t st) {
    rcm_cplkavg_kernel<<<148, 256, 0, st>>>(n, lo, hi, t, out, exp_tab, narrow);
    return cudaGetLastError();
}

size_t rcm_lbl_smem_bytes(int C, int nthreads) {
    return ((size_t)EXP_TAB * EXP_REP + (size_t)NLAY * C * 3 + 2 * C + (size_t)HALF * nthreads +
            (size_t)NLEV * (nthreads / 2)) * sizeof(double);
}

cudaError_t rcm_launch_lbl_step(const LblArgs& a, cudaStream_t st) {
    constexpr int C = RCM_LBL_C, NT = RCM_LBL_NT;
    rcm_lbl_prep_kernel<<<(a.ncol + 127) / 128, 128, 0, st>>>(a);
    const size_t smem = rcm_lbl_smem_bytes(C, NT);
    auto kern = a.clampk ? rcm_lbl_rt_kernel<C, NT, true> : rcm_lbl_rt_kernel<C, NT, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<a.ntiles * a.nchunks, NT, smem, st>>>(a);
    rcm_lbl_finish_kernel<<<(a.ncol + 127) / 128, 128, 0, st>>>(a);
    return cudaGetLastError();
}
