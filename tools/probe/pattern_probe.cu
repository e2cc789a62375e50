// Pattern probe: achievable FP64-pipe rate of the solver's instruction patterns in isolation, at the solver's
// occupancy (192 threads x 2 CTAs/SM, __launch_bounds__ as the step kernel).  Includes the solver's kernels
// file to use the very same device functions.  Build: see tools/probe/Makefile.
#include "../../our_first_climate_model_b200/csrc/rcm_kernels.cu"

namespace {
template <int WHICH>
__global__ void __launch_bounds__(192, 2) pat(double* out, int iters, const double* tab) {
    __shared__ double stab[EXP_TAB * EXP_REP];
    for (int i = threadIdx.x; i < EXP_TAB * EXP_REP; i += blockDim.x) stab[i] = tab[i / EXP_REP];
    __syncthreads();
    const unsigned tl = (unsigned)__cvta_generic_to_shared(stab + (threadIdx.x & (EXP_REP - 1)));
    double tau[HALF], D1[HALF], E1[HALF], E2[HALF], tA[HALF], tB[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) {
        tau[j] = 0.01 * (j + 1) + 1e-4 * threadIdx.x;
        D1[j] = 0.1 * j - 0.3;
        E1[j] = E2[j] = 0.0;
        tA[j] = 0.9 - 0.01 * j;
        tB[j] = 0.8 - 0.01 * j;
    }
    const double X0 = 1.0 + threadIdx.x, Dx = 0.5;
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
    for (int i = 0; i < iters; ++i) {
        const double cm = cst.cmu[i & 7], nim = cst.neg_inv_mu_l2e[i & 7];
        if (WHICH == 0) {  // sweep only: 40 DFMA, two dependent 10-chains
            sweep(tA, cm);
        } else if (WHICH == 1) {  // ten exp only: 90 FP64
#pragma unroll
            for (int j = 0; j < HALF; ++j) tB[j] = exp_scaled<false>(tau[j], nim, tl);
#pragma unroll
            for (int j = 0; j < HALF; ++j) E1[j] += tB[j];
        } else if (WHICH == 2) {  // the solver's block: sweep + ten exp: 130 FP64
#pragma unroll
            for (int j = 0; j < HALF; ++j) tB[j] = exp_scaled<false>(tau[j], nim, tl);
            sweep(tA, cm);
#pragma unroll
            for (int j = 0; j < HALF; ++j) tA[j] = tB[j];
        } else if (WHICH == 3) {  // cube + sweep: 60 FP64
            sweep(tA, cm);
#pragma unroll
            for (int j = 0; j < HALF; ++j) tA[j] = tA[j] * tA[j] * tA[j] + 0.5;
        } else if (WHICH == 4) {  // two angles swept together (round-1 schedule): 80 DFMA, four chains
            sweep(tA, cm);
            sweep(tB, nim);
        } else if (WHICH == 5) {  // 40 independent 3-register DFMAs
#pragma unroll
            for (int j = 0; j < HALF; ++j) {
                E1[j] = fma(tA[j], D1[j], E1[j]);
                E2[j] = fma(tB[j], tau[j], E2[j]);
            }
#pragma unroll
            for (int j = 0; j < HALF; ++j) {
                E1[j] = fma(tB[j], tau[j], E1[j]);
                E2[j] = fma(tA[j], D1[j], E2[j]);
            }
        }
    }
    double sacc = 0;
#pragma unroll
    for (int j = 0; j < HALF; ++j) sacc += E1[j] + E2[j] + tA[j] + tB[j];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = sacc;
}

template <int WHICH>
void run(const char* name, int fp64_per_iter, double* d, const double* tab) {
    const int iters = 3000, grid = 148 * 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    pat<WHICH><<<grid, 192>>>(d, 100, tab);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        pat<WHICH><<<grid, 192>>>(d, iters, tab);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double n = (double)grid * 192 * iters * fp64_per_iter;
    const double cyc = best * 1e-3 * 1.965e9 / iters;  // cycles per iteration per SM sub-partition (3 warps resident)
    printf("%-44s %8.1f G FP64 instr/s  = %5.1f%% of 18490   %7.1f cycles/iter/SMSP (3 warps) -> %6.1f per warp-iter\n", name,
           n / (best * 1e-3) / 1e9, 100.0 * n / (best * 1e-3) / 1e9 / 18490.0, cyc, cyc / 3.0 * 1.0);
}
}  // namespace

int main() {
    double *d, *tab;
    cudaMalloc(&d, 148 * 2 * 192 * sizeof(double));
    cudaMalloc(&tab, EXP_TAB * sizeof(double));
    double htab[EXP_TAB];
    for (int j = 0; j < EXP_TAB; ++j) {
        const double v = exp2((double)j / EXP_TAB);
        unsigned long long bits;
        memcpy(&bits, &v, 8);
        bits -= (unsigned long long)j << (20 - EXP_LOG2 + 32);
        memcpy(&htab[j], &bits, 8);
    }
    cudaMemcpy(tab, htab, sizeof(htab), cudaMemcpyHostToDevice);
    DevConst dc{};
    const double ec[4] = {0x1.62e42fefa3685p-8, 0x1.ebfbdff82c58fp-17, 0x1.c6b09b1799fcbp-26, 0x1.3b2ab6fba4e77p-35};
    for (int k = 0; k < 4; ++k) dc.expc[k] = ec[k];
    for (int k = 0; k < 8; ++k) {
        dc.cmu[k] = 0.01 * (k + 1);
        dc.neg_inv_mu_l2e[k] = -(1.0 + 0.2 * k) * 184.66496523378733;
    }
    rcm_upload_const(dc);
    run<0>("sweep only (40 DFMA, 2 chains of 10)", 40, d, tab);
    run<4>("two sweeps (80 DFMA, 4 chains of 10)", 80, d, tab);
    run<5>("40 independent 3-register DFMA", 40, d, tab);
    run<1>("ten exp (80 FP64 + 40 int) + 10 DADD", 90, d, tab);
    run<2>("sweep + ten exp (120 FP64)", 120, d, tab);
    run<3>("sweep + cube (70 FP64)", 70, d, tab);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
