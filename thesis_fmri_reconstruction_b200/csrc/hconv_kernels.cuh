// Halo-tile implicit-GEMM convolution ("hconv") for the thin-channel 5x5 convolutions (C_in in {32, 64} per tap).
//
// Why: the tap-per-TMA-box kernel (igemm_kernel) re-fetches the activation tile from L2 once per filter tap. With 32
// input channels a (tap, chunk) step carries only 128x32 MACs per 8 KB of activation, and the launches sit at 9-27 %
// tensor-pipe utilisation, bound by L2 -> shared-memory bandwidth (profiles/r1_ncu_full_gemm_kernels_B512.md). Here the
// activation HALO tile of an output tile is staged in shared memory ONCE and every tap reads it through a shifted
// shared-memory descriptor:
//   * the halo tile is stored as "16-byte column slabs": slab j holds channels [8j, 8j+8) of every halo pixel, pixels
//     linearised row-major over the padded tile width PW -> a tcgen05 no-swizzle K-major operand whose 8x16B core matrices
//     are contiguous 128 B (SBO = 128 B between 8-row groups, LBO = slab size between the two K halves of an MMA);
//   * output pixel r = yy*PW + xx of the (virtual, padded-width) tile reads halo row r + dy*PW + dx for tap (dy, dx): a tap
//     is just a start-address offset of 16*(dy*PW+dx) bytes. Columns xx >= OW of the virtual tile are computed and dropped.
//   * weights of ALL taps stay resident in shared memory for the life of the (persistent) CTA.
// Roles (384 threads): warps 0-2 = MMA issuers (the taps are dealt round-robin over G <= 3 accumulator groups; warp g issues
// group g -- an N = 16 MMA is ~8 clocks of tensor work, so the kernel is bound by how fast tcgen05.mma can be ISSUED and by
// the accumulator read-modify-write latency; several issuing warps and independent accumulators attack both), warps 3-7 =
// halo producers (cp.async 16 B with zero fill = conv padding), warps 8-11 = epilogue (TMEM -> registers -> global, adds the
// groups). Halo slabs and TMEM accumulators are double buffered, so producer, tensor pipe and epilogue of consecutive tiles
// overlap.
//
// Reference ops: Decoder.conv[3] Conv2d(C,3,5,s1,p2)+bias+tanh (/root/reference/models/vae_gan.py:118-121) and the data
// gradient of Discriminator.conv[0] Conv2d(3,C,5,s1,p2) (:145).
#pragma once
#include "ptx.cuh"

namespace fmri {

struct HcTap {
    int16_t plane;    // which halo plane the tap reads
    int16_t pad_;
    int32_t row_off;  // dy*PW + dx inside that plane (rows of 16 B)
};

struct HcParams {
    const __nv_bfloat16* X;  // NHWC input [N][H][W][C]
    int N, H, W, C;
    int num_planes;
    int pl_ys[4], pl_xs[4];      // input step per halo row / column (1, or 2 for stride-parity planes)
    int pl_yoff[4], pl_xoff[4];  // input y of halo row sy is (oy0 + sy) * ys + yoff, likewise x
    int PW, PH;                  // halo tile width / height (pixels)
    int slab_rows;               // allocated rows per slab (>= MT*128 + max tap offset + 1, multiple of 8)
    int num_taps;
    HcTap taps[25];
    int THt, tiles_y;            // output rows per tile, tiles per image
    int OH, OW;                  // output grid
    int MT;                      // 128-row MMA sub-tiles per tile
    int G;                       // independent accumulator groups the taps are dealt over (summed in the epilogue)
    const __nv_bfloat16* Bslab;  // weights in slab layout [C/8][num_taps*BN][8]
    float* img;                  // output NCHW fp32 [N][n_out][OH][OW]
    const float* bias;           // [n_out] or null
    int act;
    int n_out;                   // real output channels (<= BN)
    int accumulate;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ float hc_act(float v, int act) {
    if (act == 1) return fmaxf(v, 0.f);
    if (act == 2) return tanhf(v);
    if (act == 3) return 1.f / (1.f + __expf(-v));
    return v;
}

constexpr int HC_ISSUERS = 3;                       // MMA-issuing warps (warp g issues the taps of accumulator group g)
constexpr int HC_PRODUCERS = 160;                   // warps 3-7: halo producers
constexpr int HC_THREADS = 32 * HC_ISSUERS + HC_PRODUCERS + 128;  // + 4 epilogue warps (8-11, TMEM lane quarter = warp % 4)

// shared-memory plan (host and device agree through these helpers)
__host__ __device__ inline int hc_slab_bytes(const HcParams& p) { return p.slab_rows * 16; }
__host__ __device__ inline int hc_a_buffer_bytes(const HcParams& p) { return p.num_planes * (p.C / 8) * hc_slab_bytes(p); }
__host__ __device__ inline int hc_b_bytes(const HcParams& p, int BN) { return (p.C / 8) * p.num_taps * BN * 16; }
__host__ __device__ inline int hc_smem_bytes(const HcParams& p, int BN) {
    return 2 * hc_a_buffer_bytes(p) + hc_b_bytes(p, BN) + 512 + 1024;
}

template <int BN>
__global__ void __launch_bounds__(HC_THREADS) hconv_kernel(const __grid_constant__ HcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int a_buf = hc_a_buffer_bytes(p);
    const int slab = hc_slab_bytes(p);
    const int chunks = p.C / 8;
    uint8_t* sA = smem;                      // [2][planes][chunks][slab_rows][16]
    uint8_t* sB = smem + 2 * a_buf;          // [chunks][num_taps*BN][16]
    const int b_slab = p.num_taps * BN * 16;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + hc_b_bytes(p, BN));
    uint64_t* slab_full = bars;       // [2] count = producers
    uint64_t* slab_empty = bars + 2;  // [2] count = G (one tcgen05.commit per issuing warp)
    uint64_t* tmem_full = bars + 4;   // [2] count = G
    uint64_t* tmem_empty = bars + 6;  // [2] count = 4 (epilogue warps)
    uint64_t* b_full = bars + 8;      // count = producers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    uint32_t* s_tapoff = tmem_slot + 2;  // [25] tap offset inside a halo buffer, in 16-byte units

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.N * p.tiles_y;
    const int acc_cols = p.G * p.MT * BN;  // TMEM columns of one accumulator buffer: [group][sub-tile][BN]
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * acc_cols) tmem_cols <<= 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&slab_full[i], HC_PRODUCERS);
            mbar_init(&slab_empty[i], p.G);
            mbar_init(&tmem_full[i], p.G);
            mbar_init(&tmem_empty[i], 4);
        }
        mbar_init(b_full, HC_PRODUCERS);
        fence_barrier_init();
    }
    if (threadIdx.x < p.num_taps)
        s_tapoff[threadIdx.x] = (uint32_t)((p.taps[threadIdx.x].plane * chunks * slab) >> 4) + (uint32_t)p.taps[threadIdx.x].row_off;
    if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < HC_ISSUERS) {
        // ======================================================== MMA issuers
        // The whole warp runs this loop converged with warp-uniform operands and one ELECTED lane issues each tcgen05
        // instruction (umma_bf16_elect): the MMAs here are small (N = 16: ~8 clocks of tensor work), so the issue path must be
        // a handful of uniform-datapath instructions per MMA. Descriptors are formed by adding a (byte offset >> 4) to a base
        // descriptor -- the 14-bit start-address field never carries into the LBO field (shared memory < 256 KB).
        if (warp < p.G) {
            const uint32_t idesc = umma_idesc_bf16(128, BN, false, false);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            mbar_wait(b_full, 0);
            tc_fence_after();
            const uint64_t b_base = umma_smem_desc(smem_u32(sB), b_slab, 128, 0);
            const uint32_t a_kstep = (2 * slab) >> 4, b_kstep = (2 * b_slab) >> 4;  // two 8-channel slabs per K = 16 MMA
            const uint32_t plane_step = (chunks * slab) >> 4;
            const int ksteps = chunks / 2;
            int it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
                const int sb = it & 1;
                const uint32_t par = (it >> 1) & 1;
                mbar_wait(&slab_full[sb], par);
                mbar_wait(&tmem_empty[sb], par ^ 1);
                tc_fence_after();
                const uint64_t a_base = umma_smem_desc(smem_u32(sA + sb * a_buf), slab, 128, 0);
                // An N = 16 MMA is ~8 clocks of tensor work but accumulating into the SAME TMEM columns serialises on the
                // accumulator's read-modify-write latency. So consecutive MMAs target different accumulators: sub-tile index
                // innermost, and the taps are dealt round-robin over G accumulator groups that the epilogue adds up.
                const uint32_t dg = tmem_u + sb * acc_cols + warp * (p.MT * BN);
                uint32_t acc = 0;  // the first MMA of this group's accumulators overwrites
#pragma unroll 1
                for (int tp = warp; tp < p.num_taps; tp += p.G) {
                    const uint64_t ad = a_base + (uint32_t)(p.taps[tp].plane * plane_step + p.taps[tp].row_off);
                    const uint64_t bd = b_base + (uint32_t)(tp * BN);
                    for (int j = 0; j < ksteps; ++j) {
                        const uint64_t adj = ad + j * a_kstep, bdj = bd + j * b_kstep;
                        for (int m = 0; m < p.MT; ++m)
                            umma_bf16_elect(dg + m * BN, adj + (uint32_t)(m * 128), bdj, idesc, acc);
                        acc = 1;
                    }
                }
                umma_commit_elect(&slab_empty[sb]);  // halo buffer reusable once these MMAs retire
                umma_commit_elect(&tmem_full[sb]);   // accumulators ready for the epilogue
            }
        }
        __syncwarp();
    } else if (warp < HC_ISSUERS + HC_PRODUCERS / 32) {
        // ======================================================== halo producers (cp.async, zero fill = padding)
        const int ptid = threadIdx.x - 32 * HC_ISSUERS;
        {   // weights once: the global pack already has the slab layout -> straight 16 B copies
            const int n16 = hc_b_bytes(p, BN) / 16;
            const uint32_t b0 = smem_u32(sB);
            for (int i = ptid; i < n16; i += HC_PRODUCERS)
                cp_async16(b0 + i * 16, reinterpret_cast<const uint8_t*>(p.Bslab) + (size_t)i * 16, 16);
            cp_async_wait_all();
            fence_proxy_async_smem();
            mbar_arrive(b_full);
        }
        const int halo_px = p.PH * p.PW;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int sb = it & 1;
            const uint32_t par = (it >> 1) & 1;
            const int n = t / p.tiles_y;
            const int oy0 = (t - n * p.tiles_y) * p.THt;
            mbar_wait(&slab_empty[sb], par ^ 1);
            const uint32_t a0 = smem_u32(sA + sb * a_buf);
            const __nv_bfloat16* Xn = p.X + (size_t)n * p.H * p.W * p.C;
            for (int pl = 0; pl < p.num_planes; ++pl) {
                const int ys = p.pl_ys[pl], xs = p.pl_xs[pl], yo = p.pl_yoff[pl], xo = p.pl_xoff[pl];
                const uint32_t pbase = a0 + pl * chunks * slab;
                // item i = row * chunks + j; HC_PRODUCERS is a multiple of chunks, so j is fixed per thread and the halo
                // pixel advances by a constant number of rows per iteration (no division in the loop)
                const int j = ptid % chunks;
                const int row_step = HC_PRODUCERS / chunks;
                int row = ptid / chunks;
                int sy = row / p.PW, sx = row - sy * p.PW;
                const uint32_t dst0 = pbase + j * slab;
                for (; row < halo_px; row += row_step) {
                    const int iy = (oy0 + sy) * ys + yo, ix = sx * xs + xo;
                    const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                    const __nv_bfloat16* src = ok ? Xn + ((size_t)iy * p.W + ix) * p.C + j * 8 : Xn;
                    cp_async16(dst0 + row * 16, src, ok ? 16u : 0u);
                    sx += row_step;
                    while (sx >= p.PW) { sx -= p.PW; ++sy; }
                }
            }
            cp_async_wait_all();
            fence_proxy_async_smem();
            mbar_arrive(&slab_full[sb]);
        }
    } else {
        // ======================================================== epilogue
        const int q = warp & 3;
        int it = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
            const int sb = it & 1;
            const uint32_t par = (it >> 1) & 1;
            const int n = t / p.tiles_y;
            const int oy0 = (t - n * p.tiles_y) * p.THt;
            mbar_wait(&tmem_full[sb], par);
            tc_fence_after();
            for (int m = 0; m < p.MT; ++m) {
                const int r = m * 128 + q * 32 + lane;
                const int yy = r / p.PW, xx = r - yy * p.PW;
                const bool valid = yy < p.THt && (oy0 + yy) < p.OH && xx < p.OW;
                uint32_t v[16];
                const uint32_t tad = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + sb * acc_cols + m * BN;
                tmem_ld16(tad, v);
                tmem_ld_wait();
                for (int g = 1; g < p.G; ++g) {  // add the other accumulator groups
                    uint32_t u[16];
                    tmem_ld16(tad + g * (p.MT * BN), u);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) v[c] = __float_as_uint(__uint_as_float(v[c]) + __uint_as_float(u[c]));
                }
                if (valid) {
                    float* o = p.img + (((size_t)n * p.n_out) * p.OH + (oy0 + yy)) * p.OW + xx;
                    const size_t cs = (size_t)p.OH * p.OW;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        if (c < p.n_out) {
                            float f = __uint_as_float(v[c]) + (p.bias ? __ldg(p.bias + c) : 0.f);
                            f = hc_act(f, p.act);
                            if (p.accumulate) f += o[c * cs];
                            o[c * cs] = f;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[sb]);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// weights -> slab layout: dst[(j*rows + tp*BN + co)*8 + e] = co < n_out ? w[co*s_co + (j*8+e)*s_c + tap_src] : 0,
// tap_src = flip ? num_taps-1-tp : tp
__global__ void hc_pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int C, int BN,
                                       int n_out, int num_taps, long long s_co, long long s_c, int flip) {
    const int rows = num_taps * BN;
    const int total = (C / 8) * rows * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i & 7;
        const int rr = (i >> 3) % rows;
        const int j = (i >> 3) / rows;
        const int tp = rr / BN, co = rr - tp * BN;
        float v = 0.f;
        if (co < n_out) v = __ldg(w + co * s_co + (j * 8 + e) * s_c + (flip ? num_taps - 1 - tp : tp));
        dst[i] = __float2bfloat16_rn(v);
    }
}

}  // namespace fmri
