// Detection helpers and FITS decode.
//
//  * local_max_mask: the peak mask of the blob finder -- scikit-image peak_local_max as called by
//    vip_hci.metrics.detection (src/vip_hci/metrics/detection.py:277-279): a pixel is kept when no pixel of its
//    (2 d + 1)^2 neighbourhood (edge-replicated, scipy maximum_filter mode='nearest') is larger, it exceeds the
//    threshold and lies farther than d pixels from the border.  NaNs compare false both ways: never a peak, never
//    hide one.  One thread per pixel; the window rows come from L1/L2 (frames are <= 4 MB).
//  * fits_decode: the data unit of a FITS image HDU (big-endian, BITPIX 8 / 16 / 32 / 64 / -32 / -64) byte-swapped,
//    scaled (BSCALE, BZERO) and converted to fp32 on the device, so that a cube goes disk -> pinned staging -> HBM
//    without a host pass over it (vip_hci.fits.open_fits, src/vip_hci/fits/fits.py:23-117, reads through astropy and
//    converts on the host).
#include "common.cuh"

namespace vb {

__global__ void __launch_bounds__(256)
local_max_kernel(const float* __restrict__ img, int H, int W, int d, float thr, unsigned char* __restrict__ mask) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const float v = img[(size_t)y * W + x];
    bool keep = v > thr && x >= d && y >= d && x < W - d && y < H - d;
    if (keep) {
        for (int dy = -d; dy <= d && keep; ++dy) {
            const int yy = min(max(y + dy, 0), H - 1);
            const float* row = img + (size_t)yy * W;
            for (int dx = -d; dx <= d; ++dx) {
                const int xx = min(max(x + dx, 0), W - 1);
                if (row[xx] > v) { keep = false; break; }
            }
        }
    }
    mask[(size_t)y * W + x] = keep ? 1 : 0;
}

int local_max_mask(const float* img, int H, int W, int d, float thr, unsigned char* mask, cudaStream_t st) {
    VB_REQUIRE(H > 0 && W > 0 && d >= 0, "local_max_mask: bad arguments");
    local_max_kernel<<<dim3(ceil_div(W, 32), ceil_div(H, 8)), 256, 0, st>>>(img, H, W, d, thr, mask);
    VB_CHECK_LAUNCH();
    return 0;
}

__device__ __forceinline__ unsigned int bswap32(unsigned int v) { return __byte_perm(v, 0, 0x0123); }

template <int BITPIX>
__global__ void __launch_bounds__(256)
fits_decode_kernel(const unsigned char* __restrict__ raw, size_t count, double bscale, double bzero, int scaled,
                   float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= count) return;
    double v;
    if (BITPIX == -32) {
        const unsigned int w = bswap32(reinterpret_cast<const unsigned int*>(raw)[i]);
        if (!scaled) { out[i] = __uint_as_float(w); return; }
        v = (double)__uint_as_float(w);
    } else if (BITPIX == -64) {
        const uint2 w = reinterpret_cast<const uint2*>(raw)[i];
        v = __hiloint2double((int)bswap32(w.x), (int)bswap32(w.y));
    } else if (BITPIX == 8) {
        v = (double)raw[i];
    } else if (BITPIX == 16) {
        const unsigned short w = reinterpret_cast<const unsigned short*>(raw)[i];
        v = (double)(short)((w >> 8) | (w << 8));
    } else if (BITPIX == 32) {
        v = (double)(int)bswap32(reinterpret_cast<const unsigned int*>(raw)[i]);
    } else {
        const uint2 w = reinterpret_cast<const uint2*>(raw)[i];
        v = (double)(long long)(((unsigned long long)bswap32(w.x) << 32) | bswap32(w.y));
    }
    if (scaled) v = fma(v, bscale, bzero);
    out[i] = (float)v;
}

int fits_decode(const void* raw, int bitpix, size_t count, double bscale, double bzero, float* out, cudaStream_t st) {
    if (count == 0) return 0;
    const int scaled = !(bscale == 1.0 && bzero == 0.0);
    const unsigned grid = (unsigned)ceil_div(count, (size_t)256);
    const unsigned char* r = reinterpret_cast<const unsigned char*>(raw);
    switch (bitpix) {
        case -32: fits_decode_kernel<-32><<<grid, 256, 0, st>>>(r, count, bscale, bzero, scaled, out); break;
        case -64: fits_decode_kernel<-64><<<grid, 256, 0, st>>>(r, count, bscale, bzero, scaled, out); break;
        case 8:   fits_decode_kernel<8><<<grid, 256, 0, st>>>(r, count, bscale, bzero, scaled, out); break;
        case 16:  fits_decode_kernel<16><<<grid, 256, 0, st>>>(r, count, bscale, bzero, scaled, out); break;
        case 32:  fits_decode_kernel<32><<<grid, 256, 0, st>>>(r, count, bscale, bzero, scaled, out); break;
        case 64:  fits_decode_kernel<64><<<grid, 256, 0, st>>>(r, count, bscale, bzero, scaled, out); break;
        default: VB_REQUIRE(false, "fits_decode: BITPIX %d is not a FITS image type", bitpix);
    }
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
