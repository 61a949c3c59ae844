// Measured-peak probe: a register-only FFMA stream used by bench.py to report the FP32 FMA rate of THIS GPU at
// its current clocks next to the FFT roofline of the derotation (the nominal figure is 148 SMs x 128 lanes x
// 2 flop x 1.965 GHz = 74.4 TFLOP/s; tools/microbench/fp32_rate.cu measures 0.95 warp-instructions/clk/SMSP).
#include "common.cuh"

namespace vb {

__global__ void __launch_bounds__(256) fp32_probe_kernel(float* out, float b, float c, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// out: blocks * 256 floats.  Executes blocks * 256 * iters * 32 FMAs.
int fp32_probe(float* out, int blocks, int iters, cudaStream_t st) {
    VB_REQUIRE(out != nullptr && blocks > 0 && iters > 0, "fp32_probe: bad arguments");
    fp32_probe_kernel<<<blocks, 256, 0, st>>>(out, 1.0001f, 0.5f, iters);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
