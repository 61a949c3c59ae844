// Dependent-issue latency (1 warp) and throughput (16 warps x 4 independent chains) of the fp64 operations the small
// factorisations of eigh.cu are made of.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fp64_lat.cu -o fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

#define CHAIN 512

template <int OP>
__device__ __forceinline__ double step(double x, double a) {
    if (OP == 0) return fma(x, a, a);                       // DFMA
    if (OP == 1) return x * a;                              // DMUL
    if (OP == 2) return rsqrt(x) + a;                       // rsqrt(double) (+ DADD)
    if (OP == 3) return sqrt(x) + a;                        // sqrt(double)
    if (OP == 4) return a / x;                              // divide
    if (OP == 5) { double y = (double)__frcp_rn((float)x); y = fma(y, fma(-x, y, 1.0), y); return fma(y, fma(-x, y, 1.0), y) + a; }
    if (OP == 6) return (double)((float)x) + a;             // F2F round trip (+ DADD)
    if (OP == 7) return x + a;                              // DADD
    if (OP == 8) { float f = (float)x; f = rsqrtf(f); return (double)f + a; }   // F2F + MUFU.RSQ + F2F + DADD
    if (OP == 9) { double lo = __shfl_xor_sync(0xffffffffu, x, 1); return lo + a; }   // 64-bit shuffle + DADD
    return x;
}

template <int OP>
__global__ void lat_kernel(double* out, long long* cyc, double a, int ilp) {
    double x0 = 1.0 + threadIdx.x * 1e-9, x1 = 1.1, x2 = 1.2, x3 = 1.3;
    const long long t0 = clock64();
    if (ilp == 1) {
#pragma unroll 8
        for (int i = 0; i < CHAIN; ++i) x0 = step<OP>(x0, a);
    } else {
#pragma unroll 4
        for (int i = 0; i < CHAIN; ++i) { x0 = step<OP>(x0, a); x1 = step<OP>(x1, a); x2 = step<OP>(x2, a); x3 = step<OP>(x3, a); }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, double a) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
    lat_kernel<OP><<<1, 32>>>(out, cyc, a, 1); cudaDeviceSynchronize();
    lat_kernel<OP><<<1, 32>>>(out, cyc, a, 1); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double lat = (double)h / CHAIN;
    lat_kernel<OP><<<148, 512>>>(out, cyc, a, 4); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double thr = (double)h / (CHAIN * 4 * 16);        // cycles per warp-op per SM at 16 warps x ILP 4
    printf("%-34s dependent latency %7.1f cycles | 16 warps x ILP4: %6.2f cycles per warp-op per SM\n", name, lat, thr);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("DFMA", 0.999999);
    run<1>("DMUL", 0.999999);
    run<7>("DADD", 1e-9);
    run<2>("rsqrt(double) + DADD", 0.5);
    run<3>("sqrt(double) + DADD", 0.5);
    run<4>("a / x (double)", 1.5);
    run<5>("frcp seed + 2 Newton + DADD", 0.5);
    run<6>("F2F.f32.f64 + F2F.f64.f32 + DADD", 1e-9);
    run<8>("F2F + MUFU.RSQ + F2F + DADD", 0.5);
    run<9>("shfl 64-bit + DADD", 1e-9);
    return 0;
}
