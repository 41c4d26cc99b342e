// Inference-side kernels: BatchNorm folding for the eval-mode forward, the image-quality metrics the
// reference evaluates reconstructions with (/root/reference/train/train_utils.py:267-293 PearsonCorrelation, :295-425
// StructuralSimilarity), and the device-side input pipeline of the training step's left edge (8f-3:
// /root/reference/train/train_vgan_stage1.py:161-171 transforms.Normalize / GreyToColor / RandomHorizontalFlip,
// /root/reference/data_preprocessing/data_loader.py:93-111 Normalization, :187-217 RandomShift).
#pragma once
#include "ptx.cuh"

namespace fmri {

// Eval-mode BatchNorm folded into the preceding bias-free conv / linear layer:
//   y = gamma * (conv(x, w) - running_mean) / sqrt(running_var + eps) + beta = conv(x, w * s) + (beta - running_mean * s)
// with s = gamma / sqrt(running_var + eps) per OUTPUT channel. The output channel of element i is (i / inner) % C
// (Conv2d [Cout,Cin,5,5]: inner = Cin*25; ConvTranspose2d [Cin,Cout,5,5]: inner = 25; Linear [out,in]: inner = in).
__global__ void bn_fold_kernel(const float* __restrict__ w, long long n, long long inner, int C,
                               const float* __restrict__ rm, const float* __restrict__ rv, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float eps, float* __restrict__ w_out,
                               float* __restrict__ b_out) {
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (long long i = i0; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / inner) % C);
        w_out[i] = w[i] * (__ldg(gamma + c) * rsqrtf(__ldg(rv + c) + eps));
    }
    if (i0 < C) {
        const float s = gamma[i0] * rsqrtf(rv[i0] + eps);
        b_out[i0] = beta[i0] - rm[i0] * s;
    }
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums[0..4] += sum a, sum b, sum a*b, sum a*a, sum b*b  (fp64; one atomic per warp and quantity)
__global__ void __launch_bounds__(256) pearson_sums_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                           long long n, double* __restrict__ sums) {
    double s[5] = {0, 0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = __ldg(a + i), y = __ldg(b + i);
        s[0] += x; s[1] += y; s[2] += x * y; s[3] += x * x; s[4] += y * y;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double v = warp_sum_d(s[k]);
        if ((threadIdx.x & 31) == 0) atomicAdd(sums + k, v);
    }
}
// PearsonCorrelation.forward: sum(vx * vy) / (sqrt(sum vx^2) * sqrt(sum vy^2)), vx = a - mean(a), vy = b - mean(b)
__global__ void pearson_final_kernel(const double* __restrict__ sums, double n, float* __restrict__ out) {
    const double sa = sums[0], sb = sums[1];
    const double cov = sums[2] - sa * sb / n, va = sums[3] - sa * sa / n, vb = sums[4] - sb * sb / n;
    out[0] = (float)(cov / (sqrt(va) * sqrt(vb)));
}

// StructuralSimilarity.forward (mean of the local SSIM map; Gaussian window sigma 1.5, `win` taps per axis, zero padding
// win / 2, depthwise over the channels, C1 = 0.01^2, C2 = 0.03^2 -- train_utils.py:343-425). One thread per output pixel;
// acc[0] += sum of the SSIM map (fp64).
struct SsimWindow { float g[11]; };
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int planes, int H,
                                                   int W, int win, const __grid_constant__ SsimWindow wd,
                                                   double* __restrict__ acc) {
    const long long total = (long long)planes * H * W;
    const int pad = win / 2;
    double local = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const long long pl = i / ((long long)W * H);
        const float* pa = a + pl * H * W;
        const float* pb = b + pl * H * W;
        float m1 = 0.f, m2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
        for (int ky = 0; ky < win; ++ky) {
            const int yy = y + ky - pad;
            if (yy < 0 || yy >= H) continue;
            for (int kx = 0; kx < win; ++kx) {
                const int xx = x + kx - pad;
                if (xx < 0 || xx >= W) continue;
                const float wgt = wd.g[ky] * wd.g[kx];
                const float u = __ldg(pa + yy * W + xx), v = __ldg(pb + yy * W + xx);
                m1 += wgt * u; m2 += wgt * v; s11 += wgt * u * u; s22 += wgt * v * v; s12 += wgt * u * v;
            }
        }
        const float mu1_sq = m1 * m1, mu2_sq = m2 * m2, mu12 = m1 * m2;
        const float sg1 = s11 - mu1_sq, sg2 = s22 - mu2_sq, sg12 = s12 - mu12;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        local += (double)(((2.f * mu12 + C1) * (2.f * sg12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sg1 + sg2 + C2)));
    }
    local = warp_sum_d(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, local);
}
__global__ void scale_d2f_kernel(const double* __restrict__ acc, double scale, float* __restrict__ out) {
    out[0] = (float)(acc[0] * scale);
}

// ---- device-side input pipeline: uint8 HWC (or CHW) batch -> normalised fp32 NCHW training batch ---------------------------
// dst[n, c, y, x] = (src[n, ys, xs, cs] / 255 - mean[c]) / std[c] with
//   cs = c for 3-channel sources, 0 for grey ones (GreyToColor: the grey plane replicated, data_loader.py:374-400),
//   x' = flip[n] ? W - 1 - x : x        (RandomHorizontalFlip, train_vgan_stage1.py:161),
//   (xs, ys) = clamp((x', y) - shift[n], 0, W/H - 1)   (RandomShift with mode='nearest', data_loader.py:187-217; the
//   reference never combines the two: shift-then-flip is the order defined here);
// flip / shift are per-image int arrays drawn by the caller (either may be NULL).
__global__ void __launch_bounds__(256) image_pipeline_kernel(const uint8_t* __restrict__ src, int N, int H, int W, int Csrc,
                                                             const int* __restrict__ flip, const int* __restrict__ shift_xy,
                                                             float m0, float m1, float m2, float s0, float s1, float s2,
                                                             float* __restrict__ dst) {
    const long long total = (long long)N * 3 * H * W;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const int y = (int)((i / W) % H);
        const int c = (int)((i / ((long long)W * H)) % 3);
        const int n = (int)(i / ((long long)3 * W * H));
        int xs = x, ys = y;
        if (flip && flip[n]) xs = W - 1 - xs;   // the flip acts on the (already shifted) image
        if (shift_xy) {   // scipy.ndimage.shift(img, [dy, dx, 0], order=0, mode='nearest'): out[y, x] = in[y - dy, x - dx], clamped
            ys = min(max(ys - shift_xy[2 * n], 0), H - 1);
            xs = min(max(xs - shift_xy[2 * n + 1], 0), W - 1);
        }
        const int cs = Csrc == 3 ? c : 0;
        const float v = (float)src[(((long long)n * H + ys) * W + xs) * Csrc + cs] * (1.f / 255.f);
        const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
        dst[i] = (v - mean) / sd;
    }
}

}  // namespace fmri
