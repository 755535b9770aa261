// First layers of both networks on tcgen05 tensor cores.
//
// Their contraction depth is tiny (UNet: 3x3x3 = 27, ResNet-18: 7x7x1 = 49), so no TMA box describes the im2col row:
// producer warps assemble the A operand in shared memory themselves (K padded to 64 fp16 = one 128-byte swizzled row
// per output pixel), hand it to the single MMA-issuing thread through mbarriers, and an epilogue warpgroup drains the
// TMEM accumulators.  The integer preprocessing of the reference is fused in front and kept exact: pixels enter the
// MMA as v/256 (exact in fp16) and the 256/255 factor of the reference's `/255` is folded into the fp16 weights.
//
//   k_unet_stem_tc    cv2.resize INTER_AREA 2x (core.py:212) -> /255 (core.py:215-216, BGR kept) -> Conv3x3(3->64)
//                     + BN + ReLU (unet_parts.py:16-18)                            -> fp16 NHWC [N,256,256,64]
//   k_resnet_stem_tc  extract_squares (core.py:420-439) -> /255 (core.py:236-237) -> Conv7x7 s2 p3 (1->64) + BN + ReLU
//                     -> MaxPool3x3 s2 p1 (timm resnet18 conv1/bn1/act1/maxpool)   -> fp16 NHWC [N*64,16,16,64]
//
// Warp roles: warps 0-3 producers (input staging + im2col rows), then epilogue warpgroups (TMEM lanes 32*(w&3); the
// ResNet stem has two, one per 32-channel half, because its pooling epilogue is the long pole), last warp TMEM allocation
// + MMA issue.  One CTA per SM, persistent over squares / image blocks.
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace cvb {
namespace {

constexpr int kThreads = 416;      // UNet stem: 4 producer + 2 x 4 epilogue (alternate tiles) + 1 MMA warps
constexpr int kRsThreads = 544;    // ResNet stem: 2 x 4 producer + 8 epilogue + 1 MMA warps
constexpr int kStages = 4;
constexpr int kABytes = 128 * 128;   // 128 im2col rows x 64 fp16
constexpr int kBBytes = 64 * 128;    // 64 output channels x 64 fp16

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
// byte offset of the 16-byte chunk j of row r inside a K-major tile with 128-byte rows and the 128-byte swizzle
__device__ __forceinline__ uint32_t sw128(int r, int j) { return static_cast<uint32_t>(r * 128 + ((j ^ (r & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int bytes = valid ? 16 : 0;   // src-size 0 => the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// two bytes of `w` (selected by `sel`) -> half2 (b_lo/256, b_hi/256): 0x6400|b is the fp16 value 1024+b
__device__ __forceinline__ __half2 bytes_to_h2(uint32_t w, uint32_t sel) {
    const uint32_t bits = __byte_perm(w, 0x64646464u, sel);
    return __hfma2(*reinterpret_cast<const __half2*>(&bits), __float2half2_rn(1.0f / 256.0f), __float2half2_rn(-4.0f));
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 bits_h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

struct Bars {
    uint32_t full, empty, tfull, tempty;
};

// Common prologue: barriers + TMEM (kAcc 64-column accumulators).  Called by all threads.
template <int kMmaWarp, int kEpiThreads, int kAcc>
__device__ __forceinline__ uint32_t setup(uint64_t* bars, uint32_t* tmem_slot, Bars& b, int warp, int lane) {
    b.full = smem_u32(bars);
    b.empty = b.full + 8 * kStages;
    b.tfull = b.empty + 8 * kStages;
    b.tempty = b.tfull + 8 * kAcc;
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int i = 0; i < kStages; ++i) {
                mbar_init(b.full + 8 * i, 128);
                mbar_init(b.empty + 8 * i, 1);
            }
            for (int i = 0; i < kAcc; ++i) {
                mbar_init(b.tfull + 8 * i, 1);
                mbar_init(b.tempty + 8 * i, kEpiThreads);
            }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 64 * kAcc);
    }
    fence_proxy_async();   // weight tile / zero chunks written with ordinary stores are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return *tmem_slot;
}

// MMA issue for one 128x64x64 tile held in `stage`.
__device__ __forceinline__ void mma_tile(uint32_t a_addr, uint32_t b_addr, uint32_t d_tmem, uint32_t idesc) {
    const uint64_t a_desc = umma_desc_sw128(a_addr);
    const uint64_t b_desc = umma_desc_sw128(b_addr);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0 ? 1u : 0u);
}

// =====================================================================================================================
// ResNet-18 stem.  The 3x3 / stride-2 max-pool is folded into the GEMM's row order: an M tile is 8 x 16 POOLED pixels (half a
// square) and is multiplied nine times, once per window position (dy, dx), with the im2col rows of conv pixel
// (2py + dy - 1, 2px + dx - 1).  The nine accumulators of a pooled pixel then sit in the SAME TMEM lane, so the pool is a
// per-thread running maximum (3-input FMNMX3) over accumulator blocks: no shuffles, no shared-memory row ring, no named
// barriers, and ReLU + the fp16 conversion are applied once per pooled pixel instead of once per conv pixel (max commutes
// with the monotone conversion, so the bytes equal pool(relu(conv))).
// The im2col rows never touch shared memory: the A operand lives in tensor memory, written by tcgen05.st from the producer
// thread that owns the lane; only the 8 KB weight tile is read from shared memory, so the N = 64 MMAs run at the tensor rate
// (32 clocks; 52 with an A operand in shared memory, and writing + reading 16 KB per tile kept the first version of this
// kernel bound by shared-memory bandwidth).
// Unit of the pipeline = one "batch": the three window rows dy = 0..2 of one window column dx.  Their im2col rows overlap:
// with K laid out as 7 input rows x 8 columns, K step i of tile dy is K step dy + i of an 11-row column strip -- so the
// producers write ONE strip of 48 TMEM columns per batch and the three tiles read it at column offsets 8 dy.  (K columns
// 56..63 of a tile are then the next input row of the strip; their weights are zero.)  Bias and validity ride in a fifth K
// step against a second weight tile: its A operand is one of four constant 8-column blocks (written once), rows
// (1, 1, 0, ..) for a conv pixel -> + bias_hi + bias_lo, rows (0, 0, 1, ..) for a window position outside the 32 x 32 conv
// image (dy = 0 at py = 0, dx = 0 at px = 0) -> -30000, which loses every maximum it takes part in.
// TMEM: two sets of three 64-column accumulators (columns 0..383), the four marker blocks (384..415), two strips (416..511).
// Two producer groups fill the two strips alternately; the hand-over latencies (tcgen05.st -> wait -> arrive -> MMA issue ->
// commit) are per batch of 15 MMAs, not per tile of 4 as in the previous version, which they bounded.
// =====================================================================================================================
constexpr int kRsInHalfs = 70 * 72;                      // staged square: s_in[r][c] = px(r-3, c-3) / 256
constexpr int kRsInBytes = ((kRsInHalfs * 2 + 1023) / 1024) * 1024;
constexpr int kRsOffIn = 2 * kBBytes;                    // after the two weight tiles: 2 buffers x 2 copies (the second one word to the left)
constexpr int kRsOffBars = kRsOffIn + 4 * kRsInBytes;
constexpr int kRsSmem = kRsOffBars + 256 + 1024;
constexpr uint32_t kRsColMarker = 384, kRsColStrip = 416, kRsStripCols = 48;

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

__global__ void __launch_bounds__(kRsThreads, 1) k_resnet_stem_tc(const uint8_t* __restrict__ board, const uint4* __restrict__ wsw,
                                                               __half* __restrict__ out, int n_squares) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (base_addr - raw_addr);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kRsOffBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 16, bar_tfull = bar_full + 32, bar_tempty = bar_full + 48;

    for (int i = tid; i < 2 * kBBytes / 16; i += kRsThreads) reinterpret_cast<uint4*>(base)[i] = __ldg(wsw + i);
    for (int i = tid; i < 4 * kRsInBytes / 16; i += kRsThreads) reinterpret_cast<uint4*>(base + kRsOffIn)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 16) {
        if (lane == 0) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(bar_full + 8 * i, 128);
                mbar_init(bar_empty + 8 * i, 1);
                mbar_init(bar_tfull + 8 * i, 1);
                mbar_init(bar_tempty + 8 * i, 256);
            }
            mbar_fence_init();
        }
        __syncwarp();
        tmem_alloc(smem_u32(tmem_slot), 512);
    }
    fence_proxy_async();   // the weight tiles written with ordinary stores are read by the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp < 4) {
        // marker blocks (constant): variant v = (dy == 0 && upper half square) + 2 * (dx == 0); lane = pooled pixel (tid >> 4, tid & 15)
        uint32_t mk[32];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const bool bad = ((v & 1) && tid < 16) || ((v & 2) && (tid & 15) == 0);
#pragma unroll
            for (int i = 0; i < 8; ++i) mk[8 * v + i] = 0u;
            mk[8 * v] = bad ? 0u : 0x3C003C00u;        // K columns 0, 1 = 1.0: + bias_hi + bias_lo
            mk[8 * v + 1] = bad ? 0x00003C00u : 0u;    // K column 2 = 1.0: - 30000
        }
        tmem_st_32x32(tmem_base + kRsColMarker + (static_cast<uint32_t>(warp * 32) << 16), mk);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    griddep_launch();
    griddep_wait();   // weights, zero padding, barriers and TMEM were set up under the previous kernel's tail

    if (warp < 8) {
        // ------------------------------------------------------------------------------------------------ producers
        // Two groups of four warps (a warp writes the TMEM lane quadrant warp % 4); group pg fills strip pg for the batches
        // of its parity.  Group 0 also stages the square.
        const int pg = warp >> 2;
        const int p = tid & 127;           // TMEM lane = pooled pixel (p >> 4, p & 15) of the half square;
        const int row = p >> 1, half = p & 1;   // also (row, half) of the staged square
        const uint32_t strip = tmem_base + kRsColStrip + kRsStripCols * pg + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        int batch = 0, buf = 0;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        auto fetch = [&](int sq) {
            const int n = sq >> 6, q = sq & 63;
            const uint4* src = reinterpret_cast<const uint4*>(board + (static_cast<size_t>(n) * 512 + (q >> 3) * 64 + row) * 512 +
                                                             (q & 7) * 64 + half * 32);
            r0 = __ldg(src);
            r1 = __ldg(src + 1);
        };
        int sq = blockIdx.x;
        if (pg == 0 && sq < n_squares) fetch(sq);
        for (; sq < n_squares; sq += gridDim.x) {
            // 32 pixels of one row -> fp16/256, written at element offset 3 (the 7x7 window of conv column x starts at 2x-3);
            // copy 1 holds the same words one position to the left, so that every 4-word window starts on an 8-byte boundary
            // in one of the two copies
            uint32_t* in0 = reinterpret_cast<uint32_t*>(base + kRsOffIn + (2 * buf) * kRsInBytes);
            uint32_t* in1 = reinterpret_cast<uint32_t*>(base + kRsOffIn + (2 * buf + 1) * kRsInBytes);
            if (pg == 0) {
                const uint32_t wsrc[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                __half2 h[16];   // h[k] = (x_2k, x_2k+1)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    h[2 * k] = bytes_to_h2(wsrc[k], 0x4140);
                    h[2 * k + 1] = bytes_to_h2(wsrc[k], 0x4342);
                }
                // words of the staged row hold (x_odd, x_even): shift the pair stream by one pixel
                const uint32_t last = h2_bits(h[15]) >> 16;                                  // x_31
                const uint32_t left = __shfl_up_sync(0xffffffffu, last, 1);                  // x_31 of the left half-row
                const int w0 = (row + 3) * 36 + 1 + 16 * half;
                uint32_t prev = half ? left : 0u;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t cur = h2_bits(h[k]);
                    const uint32_t v = prev | (cur << 16);
                    in0[w0 + k] = v;
                    in1[w0 + k - 1] = v;
                    prev = cur >> 16;
                }
                if (half) {                                                                  // (x_63, 0)
                    in0[w0 + 16] = prev;
                    in1[w0 + 15] = prev;
                }
                const int nsq = sq + gridDim.x;
                if (nsq < n_squares) fetch(nsq);
            }
            named_bar(1, 256);
            const int px = p & 15;
            for (int T = 0; T < 2; ++T) {
                const int py = 8 * T + (p >> 4);
                for (int dx = 0; dx < 3; ++dx, ++batch) {
                    if ((batch & 1) != pg) continue;
                    // column strip of conv column cx = 2px + dx - 1: words cx .. cx+3 of staged rows 4py - 2 + j, j = 0..10
                    const int cx = 2 * px + dx - 1;
                    const bool x_ok = cx >= 0;
                    const uint32_t* src = ((cx & 1) ? in1 + (cx - 1) : in0 + cx) + (4 * py - 2) * 36;
                    uint32_t R[48];
#pragma unroll
                    for (int j = 0; j < 11; ++j) {
                        if (x_ok && (j >= 2 || py > 0)) {
                            const uint2 a = *reinterpret_cast<const uint2*>(src + j * 36);
                            const uint2 c = *reinterpret_cast<const uint2*>(src + j * 36 + 2);
                            R[4 * j] = a.x; R[4 * j + 1] = a.y; R[4 * j + 2] = c.x; R[4 * j + 3] = c.y;
                        } else {
                            R[4 * j] = R[4 * j + 1] = R[4 * j + 2] = R[4 * j + 3] = 0u;
                        }
                    }
                    R[44] = R[45] = R[46] = R[47] = 0u;
                    mbar_wait(bar_empty + 8 * pg, ((batch >> 1) & 1) ^ 1);
                    tc_fence_after();
                    tmem_st_32x32(strip, reinterpret_cast<const uint32_t(&)[32]>(R[0]));
                    tmem_st_32x16(strip + 32, R + 32);
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(bar_full + 8 * pg);
                }
            }
            buf ^= 1;
        }
    } else if (warp < 16) {
        // ------------------------------------------------------------------------------------------------ epilogue
        // warpgroup g (warps 8-11 / 12-15) owns channels 32g .. 32g+31; thread <-> TMEM lane <-> pooled pixel of the half square
        const int e = warp & 3, g = (warp - 8) >> 2;
        const int prow = e * 32 + lane;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(e * 32) << 16) + g * 32;
        int batch = 0;
        for (int sq = blockIdx.x; sq < n_squares; sq += gridDim.x) {
            for (int T = 0; T < 2; ++T) {
                float run[32];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx, ++batch) {
                    const int bb = batch & 1;
                    mbar_wait(bar_tfull + 8 * bb, (batch >> 1) & 1);
                    tc_fence_after();
                    uint32_t v0[32], v1[32];
                    tmem_ld_32x32(lane_addr + bb * 192, v0);
                    tmem_ld_32x32(lane_addr + bb * 192 + 64, v1);
                    tmem_ld_wait();
                    float t[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i]));
                    tmem_ld_32x32(lane_addr + bb * 192 + 128, v0);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(bar_tempty + 8 * bb);
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        run[i] = dx == 0 ? fmaxf(t[i], __uint_as_float(v0[i])) : fmax3(run[i], t[i], __uint_as_float(v0[i]));
                }
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = pack_h2_relu(run[2 * i], run[2 * i + 1]);   // the bias is part of the accumulator
                __half* dst = out + ((static_cast<size_t>(sq) * 16 + 8 * T + (prow >> 4)) * 16 + (prow & 15)) * 64 + 32 * g;
#pragma unroll
                for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + 8 * c) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_f16(128, 64, 0);
        const uint64_t b_desc = umma_desc_sw128(base_addr), bm_desc = umma_desc_sw128(base_addr + kBBytes);
        const bool leader = elect_one();
        int batch = 0;
        for (int sq = blockIdx.x; sq < n_squares; sq += gridDim.x) {
            for (int T = 0; T < 2; ++T) {
                for (int dx = 0; dx < 3; ++dx, ++batch) {
                    const int bb = batch & 1;
                    const uint32_t ph = (batch >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * bb, ph ^ 1);
                    mbar_wait(bar_full + 8 * bb, ph);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t strip = tmem_base + kRsColStrip + kRsStripCols * bb;
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint32_t acc = tmem_base + bb * 192 + dy * 64;
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_ts(acc, strip + 8 * (dy + k), b_desc + 2 * k, idesc, k != 0 ? 1u : 0u);
                            const uint32_t variant = (dy == 0 && T == 0 ? 1u : 0u) + (dx == 0 ? 2u : 0u);
                            umma_f16_ts(acc, tmem_base + kRsColMarker + 8 * variant, bm_desc, idesc, 1u);
                        }
                        umma_commit(bar_empty + 8 * bb);
                        umma_commit(bar_tfull + 8 * bb);
                    }
                    __syncwarp();
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// =====================================================================================================================
// UNet stem.  Work unit = 32 output rows x 64 output columns of one image (16 tiles of 2 rows x 64 columns).
// =====================================================================================================================
constexpr int kUsRawRow = 432;                           // source bytes [6*x0-16, 6*x0+416) of one 512-pixel BGR row
constexpr int kUsRawBytes = ((68 * kUsRawRow + 1023) / 1024) * 1024;
constexpr int kUsRedBytes = ((34 * 68 * 8 + 1023) / 1024) * 1024;   // reduced patch [34][68] x (B,G,R,0) fp16
constexpr int kUsOffA = kBBytes;
constexpr int kUsOffRaw = kUsOffA + kStages * kABytes;
constexpr int kUsOffRed = kUsOffRaw + 2 * kUsRawBytes;
constexpr int kUsOffOut = kUsOffRed + kUsRedBytes;          // two 16 KB staging buffers for the TMA tile store
constexpr int kUsOffBias = kUsOffOut + 2 * kABytes;
constexpr int kUsOffBars = kUsOffBias + 256;
constexpr int kUsSmem = kUsOffBars + 256 + 1024;

__global__ void __launch_bounds__(kThreads, 1) k_unet_stem_tc(const uint8_t* __restrict__ img, const uint4* __restrict__ wsw,
                                                             const float* __restrict__ bias, const __grid_constant__ CUtensorMap omap,
                                                             int n_images) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (base_addr - raw_addr);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kUsOffBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_units = n_images * 32;

    for (int i = tid; i < kBBytes / 16; i += kThreads) reinterpret_cast<uint4*>(base)[i] = __ldg(wsw + i);
    for (int i = tid; i < kStages * 128 * 2; i += kThreads)   // K columns 48..63 are written once: 48, 49 = 1.0 (x the bias rows of B), rest 0
        *reinterpret_cast<uint4*>(base + kUsOffA + (i >> 8) * kABytes + sw128((i >> 1) & 127, 6 + (i & 1))) = make_uint4((i & 1) ? 0u : 0x3C003C00u, 0, 0, 0);
    if (tid < 64) reinterpret_cast<float*>(base + kUsOffBias)[tid] = __ldg(bias + tid);
    if (tid == 0) tma_prefetch_desc(&omap);
    Bars b;
    const uint32_t tmem_base = setup<12, 128, 4>(bars, tmem_slot, b, warp, lane);
    griddep_launch();
    griddep_wait();

    if (warp < 4) {
        // ------------------------------------------------------------------------------------------------ producers
        const int p = tid;
        int stage = 0, buf = 0;
        uint32_t phase = 0;
        auto prefetch = [&](int unit, int bf) {
            const int n = unit >> 5, y0 = ((unit >> 2) & 7) * 32, x0 = (unit & 3) * 64;
            const uint8_t* src = img + static_cast<size_t>(n) * 512 * 1536;
            const uint32_t dst = base_addr + kUsOffRaw + bf * kUsRawBytes;
            for (int i = p; i < 68 * 27; i += 128) {
                const int r = i / 27, v = i - r * 27;
                const int gy = 2 * (y0 - 1) + r, gb = 6 * x0 - 16 + 16 * v;
                const bool ok = gy >= 0 && gy < 512 && gb >= 0 && gb < 1536;
                cp_async16(dst + r * kUsRawRow + v * 16, ok ? src + static_cast<size_t>(gy) * 1536 + gb : src, ok);
            }
            cp_async_commit();
        };
        int unit = blockIdx.x;
        if (unit < n_units) prefetch(unit, 0);
        for (; unit < n_units; unit += gridDim.x) {
            cp_async_wait_all();
            named_bar(1, 128);
            // INTER_AREA 2x reduction: (a+b+c+d+2)>>2 per channel, stored as (B,G,R,0)/256 in fp16
            const uint8_t* raw = base + kUsOffRaw + buf * kUsRawBytes;
            uint2* red = reinterpret_cast<uint2*>(base + kUsOffRed);
            for (int i = p; i < 34 * 67; i += 128) {
                const int r = i / 67, c = i - r * 67;
                const uint16_t* s0 = reinterpret_cast<const uint16_t*>(raw + (2 * r) * kUsRawRow + 6 * c + 10);
                const uint16_t* s1 = reinterpret_cast<const uint16_t*>(raw + (2 * r + 1) * kUsRawRow + 6 * c + 10);
                const uint32_t a0 = s0[0], a1 = s0[1], a2 = s0[2];   // (B0,G0) (R0,B1) (G1,R1)
                const uint32_t c0 = s1[0], c1 = s1[1], c2 = s1[2];
                const uint32_t bsum = (a0 & 255u) + (a1 >> 8) + (c0 & 255u) + (c1 >> 8) + 2u;
                const uint32_t gsum = (a0 >> 8) + (a2 & 255u) + (c0 >> 8) + (c2 & 255u) + 2u;
                const uint32_t rsum = (a1 & 255u) + (a2 >> 8) + (c1 & 255u) + (c2 >> 8) + 2u;
                const uint32_t bg = (bsum >> 2) | ((gsum >> 2) << 8), rz = rsum >> 2;
                uint2 o;
                o.x = h2_bits(bytes_to_h2(bg, 0x4140));
                o.y = h2_bits(bytes_to_h2(rz, 0x4140)) & 0x0000ffffu;   // channel 3 = 0 (bytes_to_h2 of byte 0 is already 0)
                red[r * 68 + c] = o;
            }
            const int next = unit + gridDim.x;
            if (next < n_units) prefetch(next, buf ^ 1);
            named_bar(1, 128);
            const int ty = p >> 6, x = p & 63;
            for (int t = 0; t < 16; ++t) {
                mbar_wait(b.empty + 8 * stage, phase ^ 1);
                uint8_t* a = base + kUsOffA + stage * kABytes;
                const uint2* src = red + (2 * t + ty) * 68 + x;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint2 q0 = src[ky * 68], q1 = src[ky * 68 + 1], q2 = src[ky * 68 + 2], q3 = src[ky * 68 + 3];
                    *reinterpret_cast<uint4*>(a + sw128(p, 2 * ky)) = make_uint4(q0.x, q0.y, q1.x, q1.y);
                    *reinterpret_cast<uint4*>(a + sw128(p, 2 * ky + 1)) = make_uint4(q2.x, q2.y, q3.x, q3.y);
                }
                fence_proxy_async();
                mbar_arrive(b.full + 8 * stage);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            buf ^= 1;
        }
    } else if (warp < 12) {
        // ------------------------------------------------------------------------------------------------ epilogue
        // TMEM -> bias + ReLU -> fp16 -> swizzled staging tile -> one TMA store of the 2-row x 64-column x 64-channel box.
        // Two groups of four warps take alternate tiles (the chain wait -> TMEM load -> convert -> stage -> store is
        // latency-bound, two tiles in flight hide it); four accumulators, one staging buffer per group.
        const int e = warp & 3, g = (warp - 4) >> 2, etid = (tid - 128) & 127;
        const int row = e * 32 + lane;
        uint8_t* dst = base + kUsOffOut + g * kABytes;
        int iter = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int n = unit >> 5, y0 = ((unit >> 2) & 7) * 32, x0 = (unit & 3) * 64;
            for (int t = 0; t < 16; ++t, ++iter) {
                if ((iter & 1) != g) continue;
                const int acc = iter & 3;
                mbar_wait(b.tfull + 8 * acc, (iter >> 2) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(e * 32) << 16) + acc * 64;
                uint32_t v0[32], v1[32];
                tmem_ld_32x32(taddr, v0);
                tmem_ld_32x32(taddr + 32, v1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(b.tempty + 8 * acc);
                uint32_t o[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {   // the bias is part of the accumulator (two K columns of ones x bias_hi, bias_lo)
                    o[i] = pack_h2_relu(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
                    o[16 + i] = pack_h2_relu(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
                }
                if (etid == 0) bulk_wait_read<0>();   // this group's previous store has finished reading its buffer
                named_bar(2 + g, 128);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(dst + sw128(row, j)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                fence_proxy_async();
                named_bar(2 + g, 128);
                if (etid == 0) {
                    tma_store_4d(&omap, base_addr + kUsOffOut + g * kABytes, 0, x0, y0 + 2 * t, n);
                    bulk_commit();
                }
            }
        }
        if (etid == 0) bulk_wait_all();
    } else {
        // ------------------------------------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_f16(128, 64, 0);
        int stage = 0, iter = 0;
        uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            for (int t = 0; t < 16; ++t, ++iter) {
                const int acc = iter & 3;
                mbar_wait(b.tempty + 8 * acc, ((iter >> 2) & 1) ^ 1);
                mbar_wait(b.full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    mma_tile(base_addr + kUsOffA + stage * kABytes, base_addr, tmem_base + acc * 64, idesc);
                    umma_commit(b.empty + 8 * stage);
                    umma_commit(b.tfull + 8 * acc);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace

cudaError_t configure_stems_tc() {
    cudaError_t e = cudaFuncSetAttribute(k_resnet_stem_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kRsSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_unet_stem_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kUsSmem);
}

cudaError_t launch_resnet_stem_tc(const uint8_t* board, const void* wsw, const float* bias, __half* out, int n_boards, int sm_count,
                                  cudaStream_t s) {
    const int n_squares = n_boards * 64;
    if (n_squares == 0) return cudaSuccess;
    (void)bias;   // part of the packed weight tile (two K columns of ones x bias_hi, bias_lo)
    return launch_k(k_resnet_stem_tc, n_squares < sm_count ? n_squares : sm_count, kRsThreads, kRsSmem, s, true, board, static_cast<const uint4*>(wsw),
                    out, n_squares);
}

cudaError_t launch_unet_stem_tc(const uint8_t* img, const void* wsw, const float* bias, const CUtensorMap* omap, int N, int sm_count,
                                cudaStream_t s) {
    const int n_units = N * 32;
    if (n_units == 0) return cudaSuccess;
    return launch_k(k_unet_stem_tc, n_units < sm_count ? n_units : sm_count, kThreads, kUsSmem, s, true, img, static_cast<const uint4*>(wsw), bias, *omap, N);
}

}  // namespace cvb
