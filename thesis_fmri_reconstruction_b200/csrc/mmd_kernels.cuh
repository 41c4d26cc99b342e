// WAE-MMD latent penalty with the inverse-multiquadratic kernel (Tolstikhin et al., "Wasserstein Auto-Encoders", the
// estimator of the repositories README.md:287,293 of the reference cites). The reference itself ships NO MMD code
// (SURVEY.md 0-3): this is an extension named by the north star, pinned only against our own fp64 statement in
// oracle/mmd.py ("parity unpinned").
//
//   k(a, b) = sum_s C_s / (C_s + |a - b|^2),  C_s = 2 * Z * sigma2 * s,  s in {.1, .2, .5, 1, 2, 5, 10}
//   MMD     = [sum_{i != j} k(q_i, q_j) + sum_{i != j} k(p_i, p_j)] / (B (B - 1))  -  2 / B^2 * sum_{i, j} k(q_i, p_j)
//
// One kernel, two modes. A CTA owns a 64-row block of one operand and walks 64-row tiles of the other operand (a
// blockIdx.z slice of them), both staged in shared memory as [rows][129] slabs; every thread owns a 4 x 4 register tile
// of pairs (rows ti + 16 ii, columns tj + 16 jj: conflict-free / broadcast shared-memory reads, 8 loads per 16 pair
// updates). Squared distances are accumulated as sum (a - b)^2, not as |a|^2 + |b|^2 - 2ab: no cancellation, so fp32
// holds 1e-6 against the fp64 oracle; at B = 4096, Z = 128 the pairwise work is 13 GFLOP forward + 26 GFLOP backward
// (0.1 % of a step's FLOPs), which is why this stays on the fp32 pipe instead of rounding the latents to bf16 for a
// tensor-core Gram matrix.
//   forward : per-pair kernel sums -> block reduction -> three fp64 atomics (qq, pp, qp).
//   backward: dq_i = lambda * [ 2 / (B (B - 1)) * sum_{j != i} k'(q_i, q_j) 2 (q_i - q_j) - 2 / B^2 * sum_j k'(q_i, p_j) 2 (q_i - p_j) ],
//             k'(d) = - sum_s C_s / (C_s + d)^2; the coefficient tile goes through shared memory and each thread
//             accumulates a 4-row x 8-column slice of the gradient block (one 128-column chunk per blockIdx.z group);
//             with more than one j-slice the slices combine through fp32 atomics into a zeroed (or accumulated) dq.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int MMD_T = 64;     // rows per CTA block and per tile of the other operand
constexpr int MMD_KC = 128;   // latent columns per shared-memory chunk
constexpr int MMD_LD = MMD_KC + 1;
constexpr int MMD_MAXCH = 4;  // Z <= 512
constexpr int MMD_THREADS = 256;

struct MmdParams {
    const float* q;  // [B, Z] encoded latents, row stride ldq (floats)
    const float* p;  // [B, Z] prior samples, row stride ldp
    int ldq, ldp, B, Z;
    int nsplit;      // j-tile slices (blockIdx.z % nsplit)
    float cbase;     // 2 * Z * sigma2
    double* stat;    // forward: [3] = sum_{i != j} k(q,q), sum_{i != j} k(p,p), sum k(q,p)
    float* dq;       // backward: [B, Z] fp32, row stride ldd
    int ldd, accumulate;
    float w_same, w_cross;  // backward pair weights: lambda * 4 / (B (B - 1)), -lambda * 4 / B^2
};

__device__ __forceinline__ void mmd_scales(float cbase, float (&c)[7]) {
    const float s[7] = {0.1f, 0.2f, 0.5f, 1.f, 2.f, 5.f, 10.f};
#pragma unroll
    for (int i = 0; i < 7; ++i) c[i] = cbase * s[i];
}

// stage rows [r0, r0 + 64) x columns [k0, k0 + 128) of an operand into slab[64][129] (zero beyond B / Z)
__device__ __forceinline__ void mmd_stage(float* slab, const float* src, int ld, int r0, int k0, int B, int Z) {
    for (int e = threadIdx.x; e < MMD_T * MMD_KC; e += MMD_THREADS) {
        const int r = e >> 7, k = e & (MMD_KC - 1);
        const int gr = r0 + r, gk = k0 + k;
        slab[r * MMD_LD + k] = (gr < B && gk < Z) ? src[(size_t)gr * ld + gk] : 0.f;
    }
}

template <bool BWD>
__global__ void __launch_bounds__(MMD_THREADS, 2) mmd_imq_kernel(MmdParams P) {
    extern __shared__ float sm[];
    float* si = sm;                       // [64][129] this CTA's rows (current K chunk)
    float* sj = si + MMD_T * MMD_LD;      // [64][129] tile of the other operand
    float* sc = sj + MMD_T * MMD_LD;      // [64][65] coefficient tile (backward)
    __shared__ double red[2][MMD_THREADS / 32];
    const int t = threadIdx.x, ti = t >> 4, tj = t & 15;
    const int i0 = blockIdx.x * MMD_T;
    // blockIdx.y == 0: rows i from q against q (same) then p (cross); blockIdx.y == 1 (forward only): rows i from p against p
    const bool i_from_p = blockIdx.y == 1;
    const float* isrc = i_from_p ? P.p : P.q;
    const int i_ld = i_from_p ? P.ldp : P.ldq;
    const int nch = (P.Z + MMD_KC - 1) / MMD_KC;
    const int ntj = (P.B + MMD_T - 1) / MMD_T;
    const int split = blockIdx.z % P.nsplit;
    const int och = blockIdx.z / P.nsplit;   // backward: the 128-column chunk of dq this CTA produces
    float cs[7];
    mmd_scales(P.cbase, cs);
    float acc_same = 0.f, acc_cross = 0.f;
    float g[BWD ? 4 : 1][BWD ? 8 : 1];
    if (BWD) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 8; ++k) g[a][k] = 0.f;
    }
    const int nsrc = i_from_p ? 1 : 2;
    for (int src = 0; src < nsrc; ++src) {
        const bool same = i_from_p || src == 0;          // the j operand is the same set as i
        const bool j_from_p = i_from_p || src == 1;
        const float* jsrc = j_from_p ? P.p : P.q;
        const int j_ld = j_from_p ? P.ldp : P.ldq;
        for (int jt = split; jt < ntj; jt += P.nsplit) {
            const int j0 = jt * MMD_T;
            float d[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) d[a][b] = 0.f;
            for (int ch = 0; ch < nch; ++ch) {
                __syncthreads();
                mmd_stage(si, isrc, i_ld, i0, ch * MMD_KC, P.B, P.Z);
                mmd_stage(sj, jsrc, j_ld, j0, ch * MMD_KC, P.B, P.Z);
                __syncthreads();
#pragma unroll 4
                for (int k = 0; k < MMD_KC; ++k) {
                    float av[4], bv[4];
#pragma unroll
                    for (int a = 0; a < 4; ++a) av[a] = si[(ti + 16 * a) * MMD_LD + k];
#pragma unroll
                    for (int b = 0; b < 4; ++b) bv[b] = sj[(tj + 16 * b) * MMD_LD + k];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const float df = av[a] - bv[b];
                            d[a][b] = fmaf(df, df, d[a][b]);
                        }
                }
            }
            if (!BWD) {
                float s = 0.f;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int irow = i0 + ti + 16 * a, jrow = j0 + tj + 16 * b;
                        if (irow < P.B && jrow < P.B && !(same && irow == jrow)) {
#pragma unroll
                            for (int q = 0; q < 7; ++q) s += __fdividef(cs[q], cs[q] + d[a][b]);
                        }
                    }
                if (same) acc_same += s; else acc_cross += s;
            } else {
                const float w = same ? P.w_same : P.w_cross;
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int irow = i0 + ti + 16 * a, jrow = j0 + tj + 16 * b;
                        float c = 0.f;
                        if (irow < P.B && jrow < P.B && !(same && irow == jrow)) {
#pragma unroll
                            for (int q = 0; q < 7; ++q) {
                                const float r = __fdividef(1.f, cs[q] + d[a][b]);
                                c = fmaf(-cs[q] * r, r, c);
                            }
                        }
                        sc[(ti + 16 * a) * (MMD_T + 1) + tj + 16 * b] = c * w;
                    }
                __syncthreads();   // coefficient tile complete; the distance loop no longer reads the slabs
                if (nch > 1) {     // the slabs hold the LAST chunk: bring back the chunk this CTA differentiates
                    mmd_stage(si, isrc, i_ld, i0, och * MMD_KC, P.B, P.Z);
                    mmd_stage(sj, jsrc, j_ld, j0, och * MMD_KC, P.B, P.Z);
                    __syncthreads();
                }
                // g[i][k] += sum_j c_ij (x_i[k] - x_j[k]) for rows ti + 16 a, columns tj + 16 kk
                float av[4][8];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int k = 0; k < 8; ++k) av[a][k] = si[(ti + 16 * a) * MMD_LD + tj + 16 * k];
#pragma unroll 2
                for (int j = 0; j < MMD_T; ++j) {
                    float cv[4], bv[8];
#pragma unroll
                    for (int a = 0; a < 4; ++a) cv[a] = sc[(ti + 16 * a) * (MMD_T + 1) + j];
#pragma unroll
                    for (int k = 0; k < 8; ++k) bv[k] = sj[j * MMD_LD + tj + 16 * k];
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int k = 0; k < 8; ++k) g[a][k] = fmaf(cv[a], av[a][k] - bv[k], g[a][k]);
                }
            }
        }
    }
    if (!BWD) {
        // block reduction of the two partial sums -> fp64 atomics
        double v0 = (double)acc_same, v1 = (double)acc_cross;
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_down_sync(0xffffffffu, v0, o);
            v1 += __shfl_down_sync(0xffffffffu, v1, o);
        }
        if ((t & 31) == 0) { red[0][t >> 5] = v0; red[1][t >> 5] = v1; }
        __syncthreads();
        if (t == 0) {
            double a = 0, b = 0;
            for (int w = 0; w < MMD_THREADS / 32; ++w) { a += red[0][w]; b += red[1][w]; }
            atomicAdd(P.stat + (i_from_p ? 1 : 0), a);
            if (!i_from_p) atomicAdd(P.stat + 2, b);
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int irow = i0 + ti + 16 * a;
            if (irow >= P.B) continue;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int col = och * MMD_KC + tj + 16 * k;
                if (col >= P.Z) continue;
                float* o = P.dq + (size_t)irow * P.ldd + col;
                if (P.nsplit > 1) atomicAdd(o, g[a][k]);
                else *o = P.accumulate ? *o + g[a][k] : g[a][k];
            }
        }
    }
}

// mmd[0] = lambda * { (stat_qq + stat_pp) / (B (B - 1)) - 2 stat_qp / B^2 }
__global__ void mmd_finalize_kernel(const double* stat, int B, float lambda, float* mmd) {
    const double b = (double)B;
    const double v = (stat[0] + stat[1]) / (b * (b - 1.0)) - 2.0 * stat[2] / (b * b);
    mmd[0] = (float)(lambda * v);
}
