// rcm_step_kernel.cuh - the fused repwvl step kernel: K1-K5 (sweep_item is shared with the LBL kernel)
// Included by rcm_kernels.cu inside its anonymous namespace (one translation unit: the kernels share the
// __constant__ bank `cst` and the device functions are force-inlined).
#pragma once

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
// [20][16] shared-memory arrays read by (column c, half h) lanes at rows j and 10 + j in the same instruction keep their
// rows 10..19 half a bank row further on (8 doubles resp. 16 ints of padding after row 9): at a distance of exactly 10 rows
// (1280 bytes) the two lanes of a pair would hit the same bank every time.
constexpr int TBD_LEN = NLAY * 16 + 8, TBI_LEN = NLAY * 16 + 16;
__host__ __device__ constexpr int tbd(int r, int cc) { return r * 16 + cc + (r >= HALF ? 8 : 0); }    // doubles
__host__ __device__ constexpr int tbix(int r, int cc) { return r * 16 + cc + (r >= HALF ? 16 : 0); }  // ints

constexpr double TAU_FLOOR = -8.0;       // lower clamp of tau in front of the transmissions: exp(8 * 60) is still in range
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
        // outside the table's temperature nodes the cross sections are extrapolated and can come out negative: such a
        // tile takes the global-memory K1, the variant that clamps tau from below (benign race: everybody writes 1)
        if (midT < tref + cst.t_pert[0] || midT > tref + cst.t_pert[cst.n_tpert - 1]) s.outside[NLAY + 1] = 1;
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
    if (tid == 0) s.outside[NLAY + 1] = 0;
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
            return fmax(v, TAU_FLOOR);  // a negative tau (cross sections extrapolated far outside the table) must not overflow the exp
        };
        auto tau_staged_use = [&](int j, double cl) -> double {
            const double v = tau_staged(j, cl);
            return CLAMPK ? v : fmin(v, a.tau_clamp);
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
            for (int i = tid; i < NLAY * C; i += NT) s.invT[i] = 1.0 / fmax(s.T[i], a.T_floor);
            if (tid < C) s.invTs[tid] = 1.0 / fmax(s.Ts[tid], a.T_floor);
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
                    int any = s.outside[NLAY + 1];  // extrapolation flag of prep_tau_indices, consumed here
                    s.outside[NLAY + 1] = 0;
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
                    for (int j = 0; j < HALF; ++j) tau[j] = tau_staged_use(j, cl);
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

