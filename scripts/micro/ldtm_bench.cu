// Microbenchmark: tcgen05.ld cost per 32x32b.x32 load, alone and with a concurrent tcgen05.mma stream.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../fusion_b200/csrc/ptx.cuh"
using namespace fz;

__global__ void __launch_bounds__(256, 1) k(int iters, int with_mma, int n_mma_cols, unsigned long long* out) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_slot;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 1 && lane == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (warp == 2) { ptx::tmem_alloc(&tmem_slot, 512); ptx::tmem_relinquish(); }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (warp == 1 && with_mma) {
        if (lane == 0) {
            const uint32_t sa = ptx::smem_u32(smem), sb = sa + 16384;
            const uint32_t idesc = ptx::make_idesc_bf16(128, n_mma_cols);
            long long t0 = clock64();
            for (int it = 0; it < iters * 4; ++it) {
                for (int k2 = 0; k2 < 4; ++k2)
                    ptx::mma_bf16_ss(tmem_base + 256, ptx::make_smem_desc_sw128(sa + k2 * 32), ptx::make_smem_desc_sw128(sb + k2 * 32), idesc, 1);
                if ((it & 7) == 7) { ptx::mma_commit(&bar); ptx::mbar_wait(&bar, (it >> 3) & 1); }
            }
            out[blockIdx.x * 8 + 4] = clock64() - t0;
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16);
        uint32_t r[32];
        float acc = 0.f;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            for (int c = 0; c < 8; ++c) {
                ptx::tmem_ld_32x32(t_row + c * 32, r);
                ptx::tmem_ld_wait(r);
                acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
            }
        }
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x * 8 + ew] = t1 - t0;
        if (acc == 123.f) out[7] = 1;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

int main() {
    unsigned long long* d;
    cudaMalloc(&d, 148 * 8 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int iters = 2000;
    for (int with_mma = 0; with_mma <= 1; ++with_mma)
        for (int n = 64; n <= 256; n *= 2) {
            if (!with_mma && n > 64) continue;
            cudaMemset(d, 0, 148 * 8 * 8);
            k<<<148, 256, 64 * 1024>>>(iters, with_mma, n, d);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[8];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("with_mma=%d N=%d: %s  cycles per x32 ld+wait: w0 %.1f w1 %.1f w2 %.1f w3 %.1f | cycles per MMA %.1f\n", with_mma, n,
                   cudaGetErrorString(e), h[0] / (iters * 8.0), h[1] / (iters * 8.0), h[2] / (iters * 8.0), h[3] / (iters * 8.0),
                   h[4] / (iters * 16.0));
        }
    return 0;
}
