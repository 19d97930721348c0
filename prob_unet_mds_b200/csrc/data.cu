// Device-side ClimEx sample preparation and its inverse: what `climex2torch.__getitem__`
// (climex_utils.py:122-162) does per item on the CPU -- AvgPool2d(scale) -> bilinear upsample back (align_corners=False)
// -> standardise -> residual -- and `residual_to_hr` (climex_utils.py:198-211) after sampling.  fp32 NCHW, HBM-bound.
#include "../../include/probunet_b200.h"
#include "common.cuh"

namespace pu {

// lr[n][c][i][j] = mean of the scale x scale window (row-major summation order, then one division, like at::avg_pool2d)
__global__ void climex_pool_kernel(const float* __restrict__ hr, float* __restrict__ lr, int NC, int H, int W, int scale) {
    const int h = H / scale, w = W / scale;
    const long long total = (long long)NC * h * w;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t % w);
        const int i = (int)((t / w) % h);
        const long long nc = t / ((long long)w * h);
        const float* src = hr + (nc * H + (long long)i * scale) * W + (long long)j * scale;
        float s = 0.f;
        for (int a = 0; a < scale; ++a)
            for (int b = 0; b < scale; ++b) s += src[(long long)a * W + b];
        lr[t] = s / (float)(scale * scale);
    }
}

struct StatRef {
    const float* s0;
    const float* s1;
    int mode;       // PU_STAND_*
    float eps;
};

// (shift, denominator) of the standardisation at (n, c, y, x)
__device__ __forceinline__ void stand_consts(const StatRef& st, int n, int c, long long pix, int C, long long HW, float& shift,
                                             float& den) {
    if (st.mode == PU_STAND_NONE) {
        shift = 0.f;
        den = 1.f;
    } else if (st.mode == PU_STAND_PERPIXEL) {
        const long long k = (long long)c * HW + pix;
        shift = st.s0[k];
        den = st.s1[k] + st.eps;
    } else {
        const int k = n * C + c;
        shift = st.s0[k];
        den = (st.mode == PU_STAND_MINMAX) ? (st.s1[k] - st.s0[k] + st.eps) : (st.s1[k] + st.eps);
    }
}

__global__ void climex_prepare_kernel(const float* __restrict__ hr, const float* __restrict__ lr, StatRef st, int N, int C,
                                      int H, int W, int scale, float* __restrict__ lrinterp, float* __restrict__ inputs,
                                      float* __restrict__ targets) {
    const int h = H / scale, w = W / scale;
    const float r = 1.f / (float)scale;
    const long long HW = (long long)H * W;
    const long long total = (long long)N * C * HW;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(t % W);
        const int y = (int)((t / W) % H);
        const long long nc = t / HW;
        const int c = (int)(nc % C), n = (int)(nc / C);
        // at::upsample_bilinear2d, align_corners = False, explicit scale factor
        const float sy = fmaxf(r * ((float)y + 0.5f) - 0.5f, 0.f);
        const float sx = fmaxf(r * ((float)x + 0.5f) - 0.5f, 0.f);
        const int y0 = (int)sy, x0 = (int)sx;
        const int yp = (y0 < h - 1) ? 1 : 0, xp = (x0 < w - 1) ? 1 : 0;
        const float ly = sy - (float)y0, lx = sx - (float)x0;
        const float* p = lr + (nc * h + y0) * w + x0;
        const float v = (1.f - ly) * ((1.f - lx) * p[0] + lx * p[xp]) +
                        ly * ((1.f - lx) * p[(long long)yp * w] + lx * p[(long long)yp * w + xp]);
        lrinterp[t] = v;
        float shift, den;
        stand_consts(st, n, c, (long long)y * W + x, C, HW, shift, den);
        if (st.mode == PU_STAND_NONE) {
            inputs[t] = v;
            targets[t] = hr[t] - v;
        } else {
            const float a = (v - shift) / den;
            const float b = (hr[t] - shift) / den;
            inputs[t] = a;
            targets[t] = b - a;
        }
    }
}

__global__ void climex_residual_to_hr_kernel(const float* __restrict__ residual, const float* __restrict__ lrinterp,
                                             StatRef st, int N, int C, int H, int W, float* __restrict__ out) {
    const long long HW = (long long)H * W;
    const long long total = (long long)N * C * HW;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long nc = t / HW;
        const int c = (int)(nc % C), n = (int)(nc / C);
        float shift, den;
        stand_consts(st, n, c, t % HW, C, HW, shift, den);
        out[t] = lrinterp[t] + (st.mode == PU_STAND_NONE ? residual[t] : residual[t] * den);
    }
}

static unsigned data_grid(long long total) {
    long long g = cdivll(total, 256);
    if (g > 148LL * 16) g = 148LL * 16;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace pu

extern "C" {
using namespace pu;

int pu_climex_prepare(const float* hr, const float* s0, const float* s1, int stand_mode, float eps, int N, int C, int H,
                      int W, int scale, float* lr, float* lrinterp, float* inputs, float* targets, void* stream) {
    PU_REQUIRE(hr && lr && lrinterp && inputs && targets, "pu_climex_prepare: null pointer");
    PU_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && scale >= 1 && H % scale == 0 && W % scale == 0,
               "pu_climex_prepare: H, W must be positive multiples of the low-resolution scale");
    PU_REQUIRE(stand_mode >= PU_STAND_NONE && stand_mode <= PU_STAND_MINMAX, "pu_climex_prepare: bad standardisation mode");
    PU_REQUIRE(stand_mode == PU_STAND_NONE || (s0 && s1), "pu_climex_prepare: statistics missing");
    cudaStream_t st = (cudaStream_t)stream;
    const long long nlr = (long long)N * C * (H / scale) * (W / scale);
    climex_pool_kernel<<<data_grid(nlr), 256, 0, st>>>(hr, lr, N * C, H, W, scale);
    int rc = check_launch("climex_pool");
    if (rc) return rc;
    StatRef sr{s0, s1, stand_mode, eps};
    climex_prepare_kernel<<<data_grid((long long)N * C * H * W), 256, 0, st>>>(hr, lr, sr, N, C, H, W, scale, lrinterp, inputs,
                                                                             targets);
    return check_launch("climex_prepare");
}

int pu_climex_residual_to_hr(const float* residual, const float* lrinterp, const float* s0, const float* s1, int stand_mode,
                             float eps, int N, int C, int H, int W, float* hr_pred, void* stream) {
    PU_REQUIRE(residual && lrinterp && hr_pred && N > 0 && C > 0 && H > 0 && W > 0, "pu_climex_residual_to_hr: bad arguments");
    PU_REQUIRE(stand_mode >= PU_STAND_NONE && stand_mode <= PU_STAND_MINMAX, "pu_climex_residual_to_hr: bad mode");
    PU_REQUIRE(stand_mode == PU_STAND_NONE || (s0 && s1), "pu_climex_residual_to_hr: statistics missing");
    StatRef sr{s0, s1, stand_mode, eps};
    climex_residual_to_hr_kernel<<<data_grid((long long)N * C * H * W), 256, 0, (cudaStream_t)stream>>>(
        residual, lrinterp, sr, N, C, H, W, hr_pred);
    return check_launch("climex_residual_to_hr");
}
}
