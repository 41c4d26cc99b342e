// "C -> 3" 5x5 stride-1 convolution as ONE small GEMM per pixel chunk plus a shift-and-add gather ("cto3").
//
//   out[n, co, y, x] = act(bias[co] + sum_{ky,kx} sum_c X[n, y+ky-2, x+kx-2, c] * Wt[ky*5+kx][c][co])          co < 3
//
// The halo-tile kernel (hconv_kernel<16>) issues one N = 16 MMA per (tap, 16 channels): 50 MMAs per 128 pixels whose N is 3
// real columns. A probe with its MMAs switched off (FMRI_HC_SKIP=4) halves its run time: it is bound by tcgen05.mma issue.
// Here the contraction over the C channels is done FIRST, for all 25 taps and 3 outputs at once:
//   Y[q, (tap, co)] = sum_c X[q, c] * Wt[tap][c][co]        -- a plain GEMM, M = 128 pixels, N = 80 (75 used), K = C:
//                                                              C / 16 = 2 MMAs per 128 pixels instead of 50
//   out[p, co]      = sum_tap Y[p + off(tap), (tap, co)]    -- 75 shared-memory reads + adds per output pixel
// A CTA walks whole images, 128 consecutive pixels (two image rows) per chunk: TMA brings the [128][C] bf16 tile, the MMA
// warp leaves Y in TMEM (double buffered), eight "writer" warps move it to a shared-memory ring of 5 chunks stored as 75
// planes (structure of arrays: lanes = consecutive pixels, so every access is bank-conflict free), and two groups of four
// "gather" warps (alternating iterations) finalise the 128 output pixels whose 5x5 neighbourhood is complete (pixels [128 j - 130, 128 j - 2) after chunk j), apply
// bias + activation and write coalesced NCHW rows. Zero padding costs nothing: taps that fall outside the image are skipped.
// fp32 from the accumulator to the output (Y is never rounded).
//
// Used for Decoder.conv[3] forward (Conv2d(32, 3, 5, s1, p2) + bias + tanh, /root/reference/models/vae_gan.py:118-121) and the
// data gradient of Discriminator.conv[0] (Conv2d(3, 32, 5, s1, p2), :145-147; flipped taps) when C = 32 and W = 64.
#pragma once
#include "ptx.cuh"
#include "hconv_kernels.cuh"

namespace fmri {

constexpr int C3_RING_CHUNKS = 5;
constexpr int C3_RING = C3_RING_CHUNKS * 128;       // pixels held by the Y ring
constexpr int C3_PLANES = 75;                        // (tap, co)
constexpr int C3_NB = 80;                            // GEMM N (75 padded to a multiple of 16)
constexpr int C3_W = 64;                             // image width this kernel is specialised for (two rows per chunk)
constexpr int C3_THREADS = 576;                      // TMA warp, MMA warp, 8 writer warps, 2 x 4 gather warps
constexpr int C3_A_BYTES = 128 * 32 * 2;             // one [128 pixels][32 channels] bf16 tile
constexpr int C3_W_BYTES = 6144;                     // [80][32] bf16 = 5120, padded
constexpr int C3_Y_OFF = 2 * C3_A_BYTES + C3_W_BYTES;
constexpr int C3_BAR_OFF = C3_Y_OFF + C3_PLANES * C3_RING * 4;
constexpr int C3_SMEM = C3_BAR_OFF + 256 + 1024;

// explicit shared-space accesses: through a generic `float*` derived from the dynamic shared-memory base nvcc emits generic
// LD.E / ST.E (address-space check per access, longer latency) for the ring traffic, which is this kernel's inner loop
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

struct C3Params {
    CUtensorMap mapX;   // (C, H*W, 1, N) box (32, 128, 1, 1), 64-byte swizzle
    CUtensorMap mapW;   // (C, 80) box (32, 80), 64-byte swizzle
    int N, H;
    float* out;         // [N][3][H][64] fp32
    const float* bias;  // [3] or null
    int act;
};

__global__ void __launch_bounds__(C3_THREADS, 1) cto3_kernel(const __grid_constant__ C3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem + 2 * C3_A_BYTES;
    const uint32_t ys_u32 = smem_u32(smem + C3_Y_OFF);             // Y ring: [75 planes][C3_RING] fp32
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C3_BAR_OFF);
    uint64_t* a_full = bars;        // [2] TMA tile landed
    uint64_t* a_empty = bars + 2;   // [2] MMAs that read the tile retired
    uint64_t* t_full = bars + 4;    // [2] accumulator buffer complete
    uint64_t* t_empty = bars + 6;   // [2] drained by the 8 writer warps
    uint64_t* y_full = bars + 8;    // [2] chunk G is in the ring (8 writer warps)
    uint64_t* y_done = bars + 10;   // [2] gather G finished (the 4 warps of gather group G & 1)
    uint64_t* w_full = bars + 12;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = p.H * C3_W;
    const int chunks = (HW + 127) >> 7;
    const int iters = chunks + 2;            // two gather-only iterations flush the last 130 pixels of an image

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&p.mapX);
        tma_prefetch_desc(&p.mapW);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 8);
            mbar_init(&y_full[i], 8);
            mbar_init(&y_done[i], 4);
        }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer: the weights once, then one [128][32] tile per real chunk =================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_full, C3_NB * 32 * 2);
            tma_load_2d(sW, &p.mapW, w_full, 0, 0);
        }
        __syncwarp();
        int gr = 0;
        for (int img = blockIdx.x; img < p.N; img += gridDim.x)
            for (int j = 0; j < chunks; ++j, ++gr) {
                const int s = gr & 1;
                mbar_wait(&a_empty[s], ((gr >> 1) & 1) ^ 1);
                mbar_arrive_expect_tx_elect(&a_full[s], C3_A_BYTES);
                tma_load_4d_elect(smem + s * C3_A_BYTES, &p.mapX, &a_full[s], 0, j * 128, 0, img);
            }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer: Y[128][80] = X[128][32] * Wt[80][32]^T, two K = 16 steps =================
        constexpr uint32_t idesc = umma_idesc_bf16(128, C3_NB, false, false);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        mbar_wait(w_full, 0);
        tc_fence_after();
        const uint64_t bdesc = umma_smem_desc(smem_u32(sW), 16, 512, UMMA_SW64);
        int gr = 0;
        for (int img = blockIdx.x; img < p.N; img += gridDim.x)
            for (int j = 0; j < chunks; ++j, ++gr) {
                const int s = gr & 1;
                mbar_wait(&a_full[s], (gr >> 1) & 1);
                mbar_wait(&t_empty[s], ((gr >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint64_t adesc = umma_smem_desc(smem_u32(smem + s * C3_A_BYTES), 16, 512, UMMA_SW64);
                umma_bf16_elect(tmem_u + s * 128, adesc, bdesc, idesc, 0);
                umma_bf16_elect(tmem_u + s * 128, adesc + 2, bdesc + 2, idesc, 1);
                umma_commit_elect(&a_empty[s]);
                umma_commit_elect(&t_full[s]);
            }
        __syncwarp();
    } else if (warp < 10) {
        // ================= writers: TMEM -> registers -> the ring's 75 planes =================
        // Every role here is a single-warp latency chain (~10 cycles per issued instruction), so the work is spread: two
        // warps per TMEM lane quarter split the 75 columns (0..39 | 40..74).
        const int q = warp & 3;                 // TMEM lane quarter of this warp
        const int half = (warp - 2) >> 2;
        const int r = q * 32 + lane;            // pixel of the chunk
        int gr = 0, G = 0;
        for (int img = blockIdx.x; img < p.N; img += gridDim.x)
            for (int j = 0; j < iters; ++j, ++G) {
                mbar_wait(&y_done[G & 1], ((G >> 1) & 1) ^ 1);      // gather G - 2 no longer reads the slot chunk G overwrites
                if (j < chunks) {
                    const int s = gr & 1;
                    mbar_wait(&t_full[s], (gr >> 1) & 1);
                    tc_fence_after();
                    const uint32_t dst = ys_u32 + ((G % C3_RING_CHUNKS) * 128 + r) * 4;
                    const uint32_t tad = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * 128;
                    uint32_t v[32], u[16];
                    tmem_ld32(tad + half * 32, v);
                    tmem_ld16(tad + 32 + half * 32, u);
                    tmem_ld_wait();
                    if (half == 0) {            // columns 0..31 from v, 32..39 from u
#pragma unroll
                        for (int c = 0; c < 32; ++c) sts_f32(dst + c * C3_RING * 4, __uint_as_float(v[c]));
#pragma unroll
                        for (int c = 0; c < 8; ++c) sts_f32(dst + (32 + c) * C3_RING * 4, __uint_as_float(u[c]));
                    } else {                    // columns 40..63 from v[8..31], 64..74 from u[0..10]
#pragma unroll
                        for (int c = 8; c < 32; ++c) sts_f32(dst + (32 + c) * C3_RING * 4, __uint_as_float(v[c]));
#pragma unroll
                        for (int c = 0; c < 11; ++c) sts_f32(dst + (64 + c) * C3_RING * 4, __uint_as_float(u[c]));
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[s]);
                    ++gr;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&y_full[G & 1]);
            }
    } else {
        // ================= gather: out[p] = sum over the 5x5 neighbourhood of p, bias, activation, NCHW store =================
        // two groups of four warps: group G & 1 finalises the output block of iteration G, so consecutive blocks overlap
        const int grp = (warp - 10) >> 2;
        const int gt = (threadIdx.x - 320) & 127;       // 0..127 inside the group
        const float b0 = p.bias ? __ldg(p.bias) : 0.f, b1 = p.bias ? __ldg(p.bias + 1) : 0.f, b2 = p.bias ? __ldg(p.bias + 2) : 0.f;
        int G = 0, L = 0;
        for (int img = blockIdx.x; img < p.N; img += gridDim.x, ++L)
            for (int j = 0; j < iters; ++j, ++G) {
                if ((G & 1) != grp) continue;
                mbar_wait(&y_full[G & 1], (G >> 1) & 1);
                const int lo = max(0, 128 * j - 130), hi = min(HW, 128 * j - 2);
                const int pix = lo + gt;
                if (pix < hi) {
                    const int y = pix >> 6, x = pix & 63;
                    float a0 = b0, a1 = b1, a2 = b2;
                    // branch-free: out-of-image taps become predicated-off loads (no divergent control flow, all 75 loads of a
                    // pixel independent and in flight together)
                    const bool vx0 = x >= 2, vx1 = x >= 1, vx3 = x < C3_W - 1, vx4 = x < C3_W - 2;
                    const int g0 = L * iters;
#pragma unroll
                    for (int ky = 0; ky < 5; ++ky) {
                        const int yy = y + ky - 2;
                        const bool vy = yy >= 0 && yy < p.H;
                        // row yy lies in chunk yy >> 1 of this image = global chunk L * iters + (yy >> 1)
                        const int slot = (g0 + (yy >> 1) + C3_RING_CHUNKS) % C3_RING_CHUNKS;
                        const uint32_t src = ys_u32 + (slot * 128 + (yy & 1) * 64 + x - 2 + (ky * 15) * C3_RING) * 4;
#pragma unroll
                        for (int kx = 0; kx < 5; ++kx) {
                            const bool v = vy && (kx == 0 ? vx0 : kx == 1 ? vx1 : kx == 3 ? vx3 : kx == 4 ? vx4 : true);
                            if (v) {     // a predicate on three independent ld.shared, not a divergent branch
                                a0 += lds_f32(src + ((kx * 3 + 0) * C3_RING + kx) * 4);
                                a1 += lds_f32(src + ((kx * 3 + 1) * C3_RING + kx) * 4);
                                a2 += lds_f32(src + ((kx * 3 + 2) * C3_RING + kx) * 4);
                            }
                        }
                    }
                    float* o = p.out + ((size_t)img * 3) * HW + pix;
                    o[0] = hc_act(a0, p.act);
                    o[HW] = hc_act(a1, p.act);
                    o[2 * (size_t)HW] = hc_act(a2, p.act);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&y_done[G & 1]);
            }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// Wt[(tap*3 + co)][c] = w[co*s_co + c*s_c + (flip ? 24 - tap : tap)], bf16, rows 75..79 zero.
__global__ void c3_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, long long s_co,
                                       long long s_c, int flip, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C3_NB * C) return;
    const int row = i / C, c = i - row * C;
    float v = 0.f;
    if (row < C3_PLANES) {
        const int tap = row / 3, co = row - tap * 3;
        v = __ldg(w + co * s_co + c * s_c + (flip ? 24 - tap : tap));
    }
    dst[i] = __float2bfloat16_rn(v);
}

}  // namespace fmri
