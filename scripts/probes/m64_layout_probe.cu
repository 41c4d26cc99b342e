// Which TMEM lanes hold which accumulator rows for tcgen05.mma cta_group::1 with M = 64? (No network, no PTX manual in the image:
// measured.) D1 = an M = 128 MMA that writes -1 everywhere, then D2 = an M = 64 MMA (accumulate = 0) with A[row][0] = row + 1,
// B[n][0] = 1: lane l of TMEM column 0 then shows row + 1 where the M = 64 instruction wrote, -1 where it did not.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I thesis_fmri_reconstruction_b200/csrc -o gpurun_out/m64_probe scripts/probes/m64_layout_probe.cu -lcuda
#include <cstdio>
#include "ptx.cuh"
#include "hconv_kernels.cuh"
using namespace fmri;

__global__ void probe(float* out) {
    __shared__ __align__(1024) uint8_t smem[16384];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __nv_bfloat16* A128 = reinterpret_cast<__nv_bfloat16*>(smem);           // [128 rows][16 k], no swizzle, K-major
    __nv_bfloat16* A64 = reinterpret_cast<__nv_bfloat16*>(smem + 4096);     // [64 rows][16 k]
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(smem + 8192);       // [16 n][16 k]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // element (r, k) at (r/8)*256 + (k/8)*128 + (r%8)*16 + (k%8)*2 bytes  (LBO = 128 between K core matrices, SBO = 256)
    auto at = [](__nv_bfloat16* base, int r, int k) -> __nv_bfloat16& {
        return *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(base) + (r / 8) * 256 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2);
    };
    for (int i = tid; i < 16384 / 2; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(0.f);
    __syncthreads();
    if (tid < 128) at(A128, tid, 0) = __float2bfloat16(-1.f);
    if (tid < 64) at(A64, tid, 0) = __float2bfloat16((float)(tid + 1));
    if (tid < 16) at(B, tid, 0) = __float2bfloat16(1.f);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&slot, 32);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 0) {
        const uint64_t a128 = umma_smem_desc(smem_u32(A128), 128, 256, 0), a64 = umma_smem_desc(smem_u32(A64), 128, 256, 0);
        const uint64_t b = umma_smem_desc(smem_u32(B), 128, 256, 0);
        umma_bf16_elect(tm, a128, b, umma_idesc_bf16(128, 16, false, false), 0);
        umma_bf16_elect(tm, a64, b, umma_idesc_bf16(64, 16, false, false), 0);
        umma_commit_elect(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    if (warp < 4) {
        uint32_t v[16];
        tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16), v);
        tmem_ld_wait();
        out[warp * 32 + lane] = __uint_as_float(v[0]);
        out[128 + warp * 32 + lane] = __uint_as_float(v[5]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 32);
}

int main() {
    float *d, h[256];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0, sizeof(h));
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int c = 0; c < 2; ++c) {
        printf("column %d: TMEM lane -> value (row + 1 where the M = 64 MMA wrote, -1 elsewhere)\n", c ? 5 : 0);
        for (int l = 0; l < 128; ++l) printf("%s%4.0f", (l % 16) ? " " : (l ? "\n" : ""), h[c * 128 + l]);
        printf("\n");
    }
    return 0;
}
