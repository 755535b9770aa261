// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, cta_group::1) issued back to back by one thread per SM,
// for N in {64,128,256}, operands in 128B-swizzled K-major shared-memory tiles.  Variants:
//   mode 0: every MMA reads the same A and B tile          mode 1: cycles through 4 stage buffers (pipeline-like addresses)
//   mode 2: mode 1 + four other warps stream st.shared into a fifth buffer (TMA-like write traffic)
//   mode 3: mode 1 + four warps run tcgen05.ld on the other accumulator (epilogue-like TMEM reads)
//   mode 5/7: accumulator window slides by 64 columns every 12/24 MMAs (overlapping, different D: row-streaming conv)
//   mode 6: two disjoint accumulator rings alternate every 12 MMAs, each sliding by 64 columns
//   mode 4: mode 1 + accumulator window sliding over the TMEM columns and A start addresses shifted by 128/256 B (row-streaming conv)
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
                for (int k = 0; k < 4; ++k) {
                    uint32_t d = tmem;
                    if (mode == 4) d += ((r & 7) * 64u) % (512u - N + 64u) / 64u * 64u;
                    if (mode == 5) d += ((r / 3) % 5) * 64u;                              // window slides by 64 columns every 12 MMAs
                    if (mode == 6) d += ((r / 3) & 1) * 256u + (((r / 3) >> 1) & 1) * 64u;   // two rings alternate every 12 MMAs
                    if (mode == 7) d += ((r / 6) % 5) * 64u;                              // slides every 24 MMAs
                    umma_f16(d, a + (mode == 4 ? 8u * (r % 3) : 0u) + 2 * k, b + 2 * k, idesc, 1u);
                }
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

// Row-streaming stream (conv3x3_rs_kernel): per input row 12 K steps into a 3-slot window of a ring of eight 64-column
// slots that slides by one slot per row.  variant bit 0: split the first K step (N=64 overwrite + N=128), bit 1: split the
// window where it wraps around the ring (otherwise the ring is 6 windows long and never wraps), bit 2: commit per row.
__global__ void __launch_bounds__(384, 1) probe_rs(int rows, int variant, long long* out, int fill) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - raw);
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { done = 0; mbar_init(smem_u32(&bar3), 1); }
    for (int i = threadIdx.x; i < (4 * 17408 + 3 * 24576) / 4; i += blockDim.x) {
        // fill 0: zeros; fill 1: pseudo-random fp16 pairs in [-1, 1) (exponent bits 0x3800..0x3bff, random sign and mantissa)
        uint32_t h = (i + 1) * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const uint32_t v = fill ? ((h & 0x83FF83FFu) | 0x38003800u) : 0u;
        reinterpret_cast<uint32_t*>(bp)[i] = v;
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t idesc0 = umma_idesc_f16(128, 64, 0) & ~(0x3Fu << 17);
    if (warp == 0) {
        if (elect_one()) {
            const long long t0 = clock64();
            const uint32_t wbase = base + 4 * 17408;
            for (int r = 0; r < rows; ++r) {
                const uint32_t s_base = (variant & 2) ? ((0u - r) & 7u) : (5u - (r % 6));
                const uint64_t a0 = umma_desc_sw128(base + (r & 3) * 17408);
                bool fresh = (variant & 1) != 0;
                for (int dd = 0; dd < 3; ++dd) {
                    const uint64_t b0 = umma_desc_sw128(wbase + dd * 24576);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t a = a0 + 8u * dd + 2u * k, b = b0 + 2u * k;
                        int lo = 0;
                        if (fresh) {
                            umma_f16(tmem + s_base * 64u, a, b, idesc0 | (8u << 17), 0u);
                            lo = 1;
                            fresh = false;
                        }
                        while (lo <= 2) {
                            const uint32_t s0 = (s_base + lo) & 7u;
                            int n = 3 - lo;
                            if (s0 + n > 8) n = 8 - s0;
                            umma_f16(tmem + s0 * 64u, a, b + lo * 512u, idesc0 | (static_cast<uint32_t>(n * 8) << 17), 1u);
                            lo += n;
                        }
                    }
                }
                if (variant & 4) { umma_commit(smem_u32(&bar2)); umma_commit(smem_u32(&bar2)); }
                if (variant & 8) { mbar_try_wait(smem_u32(&bar3), 1); tc_fence_after(); }
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            out[blockIdx.x] = clock64() - t0;
            done = 1;
        }
        __syncwarp();
    } else if (variant & 16) {
        while (!done) mbar_try_wait(smem_u32(&bar3), 0);   // never completes: spinning waiters like idle epilogue warps
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

void run_rs(int variant, int fill) {
    const int rows = 2048, sms = 148;
    long long* d; cudaMalloc(&d, sms * sizeof(long long));
    const int smem = 4 * 17408 + 3 * 24576 + 1024;
    cudaFuncSetAttribute(probe_rs, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int it = 0; it < 2; ++it) probe_rs<<<sms, 384, smem>>>(rows, variant, d, fill);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("row-streaming stream, variant %d (1 split first, 2 ring wraps, 4 commit/row), %s operands: %.0f cycles/row (ideal 1152) %s\n", variant, fill ? "random" : "zero",
           (double)mx / rows, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
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
    for (int mode = 0; mode < 5; ++mode) { run<64>(mode); run<128>(mode); run<192>(mode); run<256>(mode); }
    for (int mode = 5; mode < 8; ++mode) { run<64>(mode); run<128>(mode); run<192>(mode); }
    for (int v : {0, 7, 15, 23, 31}) run_rs(v, 1);
    return 0;
}
