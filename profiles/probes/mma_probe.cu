// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) issued back to back by one thread per SM,
// for N in {64,128,256}, operands in 128B-swizzled K-major shared-memory tiles.  Variants:
//   mode 0: every MMA reads the same A and B tile          mode 1: cycles through 4 stage buffers (pipeline-like addresses)
//   mode 2: mode 1 + four other warps stream st.shared into a fifth buffer (TMA-like write traffic)
//   mode 3: mode 1 + four warps run tcgen05.ld on the other accumulator (epilogue-like TMEM reads)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../chessvision-3lc_b200/csrc mma_probe.cu -o mma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"
using namespace cvb;

template <int N>
__global__ void __launch_bounds__(192, 1) probe(int reps, int mode, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - raw);
    constexpr int kA = 128 * 128, kB = N * 128, kStage = kA + kB;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (4 * kStage + 16384) / 16; i += blockDim.x) reinterpret_cast<uint4*>(bp)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t idesc = umma_idesc_f16(128, N, 0);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (elect_one()) {
                const int st = mode == 0 ? 0 : (r & 3);
                const uint64_t a = umma_desc_sw128(base + st * kStage), b = umma_desc_sw128(base + st * kStage + kA);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16(tmem, a + 2 * k, b + 2 * k, idesc, 1u);
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit(smem_u32(&bar));
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    } else if (warp >= 2 && mode == 2) {
        uint4* dst = reinterpret_cast<uint4*>(bp + 4 * kStage);
        volatile int* flag = reinterpret_cast<volatile int*>(&bar);
        for (int r = 0; r < reps * 4; ++r) {   // 128 threads x 16 B = 2 KB per iteration
            dst[(threadIdx.x - 64) + 128 * (r & 7)] = make_uint4(r, r, r, r);
            (void)flag;
        }
    } else if (warp >= 2 && mode == 3) {
        uint32_t v[32];
        uint32_t acc = 0;
        for (int r = 0; r < reps / 2; ++r) {
            tmem_ld_32x32(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + (r & 1) * 32, v);
            tmem_ld_wait();
            acc += v[r & 31];
        }
        if (acc == 0x12345678u) out[0] = 0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N>
void run(int mode) {
    const int reps = 4096, sms = 148;
    long long* d; cudaMalloc(&d, sms * sizeof(long long));
    const int smem = 4 * (128 * 128 + N * 128) + 16384 + 1024;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 2; ++it) probe<N><<<sms, 192, smem>>>(reps, mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1LL << 60; for (int i = 0; i < sms; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    printf("N=%3d mode=%d: cycles/MMA min %.1f max %.1f (ideal %.0f) %s\n", N, mode, (double)mn / (reps * 4), (double)mx / (reps * 4), N / 2.0,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int mode = 0; mode < 4; ++mode) { run<64>(mode); run<128>(mode); run<256>(mode); }
    return 0;
}
