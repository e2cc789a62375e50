// FP64 issue-port probe for sm_100a: how much DFMA throughput survives when integer / LDS instructions are
// interleaved, and how many warps x ILP the FP64 pipe needs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o issue_probe issue_probe.cu ; run on the GPU box.  Output feeds DESIGN.md section 3.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int NINT, int NLDS>
__global__ void __launch_bounds__(1024) probe(double* out, int iters, int* iout) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x * 1e-9;
    __syncthreads();
    double v[ILP];
    int q[NINT > 0 ? NINT : 1];
#pragma unroll
    for (int k = 0; k < ILP; ++k) v[k] = 1.0 + 1e-3 * (threadIdx.x + k);
#pragma unroll
    for (int k = 0; k < (NINT > 0 ? NINT : 1); ++k) q[k] = threadIdx.x + k;
    const double a = 0.999999, b = 1e-7;
    double ls = 0.0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) v[k] = fma(v[k], a, b);
#pragma unroll
            for (int k = 0; k < NINT; ++k) q[k] = (q[k] ^ (q[k] >> 3)) + i;   // SHF+LOP3/IADD: 2-3 int instr
#pragma unroll
            for (int k = 0; k < NLDS; ++k) { ls += sh[(idx + k * 32 + r) & 1023]; }
        }
    }
    double s = ls;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += v[k];
    int t = 0;
#pragma unroll
    for (int k = 0; k < (NINT > 0 ? NINT : 1); ++k) t += q[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int ILP, int NINT, int NLDS>
void run(const char* name, int threads, int ctas_per_sm, double* d, int* di) {
    const int iters = 4000;
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<ILP, NINT, NLDS><<<grid, threads>>>(d, 200, di);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<ILP, NINT, NLDS><<<grid, threads>>>(d, iters, di);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double dfma = (double)grid * threads * 8.0 * ILP * iters;
    printf("%-34s thr/CTA %4d CTA/SM %d  warps/SMSP %4.1f : %8.1f G DFMA/s  (%.3f ms)\n", name, threads, ctas_per_sm,
           threads * ctas_per_sm / 128.0, dfma / (best * 1e-3) / 1e9, best);
}

int main() {
    double* d; int* di;
    cudaMalloc(&d, 148 * 8 * 1024 * sizeof(double));
    cudaMalloc(&di, 148 * 8 * 1024 * sizeof(int));
    for (int cfg = 0; cfg < 3; ++cfg) {
        const int thr = cfg == 0 ? 192 : (cfg == 1 ? 256 : 1024), cps = cfg == 0 ? 2 : (cfg == 1 ? 4 : 2);
        run<1, 0, 0>("ILP1", thr, cps, d, di);
        run<2, 0, 0>("ILP2", thr, cps, d, di);
        run<4, 0, 0>("ILP4", thr, cps, d, di);
        run<8, 0, 0>("ILP8", thr, cps, d, di);
        run<8, 1, 0>("ILP8 + 3 int/8 DFMA", thr, cps, d, di);
        run<8, 2, 0>("ILP8 + 6 int/8 DFMA", thr, cps, d, di);
        run<8, 4, 0>("ILP8 + 12 int/8 DFMA", thr, cps, d, di);
        run<8, 8, 0>("ILP8 + 24 int/8 DFMA", thr, cps, d, di);
        run<8, 0, 1>("ILP8 + 1 LDS/8 DFMA", thr, cps, d, di);
        run<8, 0, 2>("ILP8 + 2 LDS/8 DFMA", thr, cps, d, di);
        run<4, 4, 0>("ILP4 + 12 int/4 DFMA", thr, cps, d, di);
    }
    return 0;
}
