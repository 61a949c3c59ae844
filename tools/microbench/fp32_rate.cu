// Microbenchmark: issue rate of FFMA / FADD / FMUL with register operands on sm_100a
// (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rate fp32_rate.cu; run on the GPU box)
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(float* out, float b, float c, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = fmaf(a[i], b, c);          // FFMA R, R, R, R
                else if (OP == 1) a[i] = a[i] + b;             // FADD
                else if (OP == 2) a[i] = a[i] * b;             // FMUL
                else if (OP == 3) a[i] = fmaf(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);   // FFMA, 3 distinct varying regs
                else a[i] = fmaf(a[i], 1.0001f, 0.5f);         // FFMA with immediates
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
static void run(const char* name, float* out) {
    const int iters = 4096, grid = 148 * 8, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<OP><<<grid, block>>>(out, 1.0001f, 0.5f, 16);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_kernel<OP><<<grid, block>>>(out, 1.0001f, 0.5f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)grid * (block / 32) * iters * 32.0;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double ipc = warp_instr / (ms * 1e-3) / (148.0 * 4.0) / (clk * 1e3);
    printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SMSP (at %d MHz nominal)  %7.2f Tinstr-lanes/s\n", name, ms, ipc,
           clk / 1000, warp_instr * 32 / (ms * 1e-3) / 1e12);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("FFMA a=a*b+c (b,c uniform)", out);
    run<3>("FFMA 3 varying registers", out);
    run<1>("FADD", out);
    run<2>("FMUL", out);
    run<4>("FFMA immediates", out);
    return 0;
}
