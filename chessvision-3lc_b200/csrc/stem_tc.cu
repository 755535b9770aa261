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
constexpr int kRsThreads = 416;    // ResNet stem: 4 producer + 8 epilogue + 1 MMA warps
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
// ResNet-18 stem
// =====================================================================================================================
constexpr int kRsInHalfs = 70 * 72;                      // staged square: s_in[r][c] = px(r-3, c-3) / 256
constexpr int kRsInBytes = ((kRsInHalfs * 2 + 1023) / 1024) * 1024;
constexpr int kRsOffA = kBBytes;
constexpr int kRsOffIn = kRsOffA + kStages * kABytes;
constexpr int kRsOffH = kRsOffIn + 2 * kRsInBytes;       // ring of 16 horizontally pooled conv rows [16 px][64 ch]
constexpr int kRsOffBias = kRsOffH + 16 * 2048;
constexpr int kRsOffBars = kRsOffBias + 256;
constexpr int kRsSmem = kRsOffBars + 256 + 1024;

__global__ void __launch_bounds__(kRsThreads, 1) k_resnet_stem_tc(const uint8_t* __restrict__ board, const uint4* __restrict__ wsw,
                                                               const float* __restrict__ bias, __half* __restrict__ out,
                                                               int n_squares) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (base_addr - raw_addr);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + kRsOffBars);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < kBBytes / 16; i += kRsThreads) reinterpret_cast<uint4*>(base)[i] = __ldg(wsw + i);
    for (int i = tid; i < 2 * kRsInBytes / 16; i += kRsThreads) reinterpret_cast<uint4*>(base + kRsOffIn)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kStages * 128; i += kRsThreads)   // the padding chunk (K columns 56..63) is written once: columns 56, 57 = 1.0
        *reinterpret_cast<uint4*>(base + kRsOffA + (i >> 7) * kABytes + sw128(i & 127, 7)) = make_uint4(0x3C003C00u, 0, 0, 0);   // x (bias_hi, bias_lo) rows of B
    if (tid < 64) reinterpret_cast<float*>(base + kRsOffBias)[tid] = __ldg(bias + tid);
    Bars b;
    const uint32_t tmem_base = setup<12, 256, 2>(bars, tmem_slot, b, warp, lane);
    griddep_launch();
    griddep_wait();   // weights, zero padding, barriers and TMEM were set up under the previous kernel's tail

    if (warp < 4) {
        // ------------------------------------------------------------------------------------------------ producers
        const int p = tid;                 // im2col row of every tile; also (row, half) of the staged square
        const int row = p >> 1, half = p & 1;
        int stage = 0, buf = 0;
        uint32_t phase = 0;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0;
        auto fetch = [&](int sq) {
            const int n = sq >> 6, q = sq & 63;
            const uint4* src = reinterpret_cast<const uint4*>(board + (static_cast<size_t>(n) * 512 + (q >> 3) * 64 + row) * 512 +
                                                             (q & 7) * 64 + half * 32);
            r0 = __ldg(src);
            r1 = __ldg(src + 1);
        };
        int sq = blockIdx.x;
        if (sq < n_squares) fetch(sq);
        for (; sq < n_squares; sq += gridDim.x) {
            // 32 pixels of one row -> fp16/256, written at element offset 3 (the 7x7 window of output x starts at 2x-3)
            __half* in = reinterpret_cast<__half*>(base + kRsOffIn + buf * kRsInBytes);
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
            uint32_t* dst = reinterpret_cast<uint32_t*>(in) + (row + 3) * 36 + 1 + 16 * half;
            uint32_t prev = half ? left : 0u;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t cur = h2_bits(h[k]);
                dst[k] = prev | (cur << 16);
                prev = cur >> 16;
            }
            if (half) dst[16] = prev;                                                    // (x_63, 0)
            const int nsq = sq + gridDim.x;
            if (nsq < n_squares) fetch(nsq);
            named_bar(1, 128);
            const int x = p & 31, yq = p >> 5;
            for (int t = 0; t < 8; ++t) {
                mbar_wait(b.empty + 8 * stage, phase ^ 1);
                uint8_t* a = base + kRsOffA + stage * kABytes;
                const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + (2 * (4 * t + yq)) * 36 + x;
#pragma unroll
                for (int ky = 0; ky < 7; ++ky) {
                    const uint32_t* s = src + ky * 36;
                    *reinterpret_cast<uint4*>(a + sw128(p, ky)) = make_uint4(s[0], s[1], s[2], s[3]);
                }
                fence_proxy_async();
                mbar_arrive(b.full + 8 * stage);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            buf ^= 1;
        }
    } else if (warp < 12) {
        // ------------------------------------------------------------------------------------------------ epilogue
        // warpgroup g (warps 4-7 / 8-11) owns channels 32g .. 32g+31 of every conv pixel
        const int e = warp & 3, g = (warp - 4) >> 2, etid = (tid - 128) & 127;
        uint8_t* sH = base + kRsOffH;
        int iter = 0;
        for (int sq = blockIdx.x; sq < n_squares; sq += gridDim.x) {
            for (int t = 0; t < 8; ++t, ++iter) {
                const int acc = iter & 1;
                mbar_wait(b.tfull + 8 * acc, (iter >> 1) & 1);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(e * 32) << 16) + acc * 64 + g * 32, v);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(b.tempty + 8 * acc);
                // conv pixel (y = 4t + e, x = lane): bias + ReLU -> fp16, then the horizontal half of the 3x3 max pool
                uint32_t hv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i)   // the bias is part of the accumulator (two K columns of ones x bias_hi, bias_lo)
                    hv[i] = pack_h2_relu(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t up = __shfl_up_sync(0xffffffffu, hv[i], 1);       // lane 0 keeps its own value
                    const uint32_t dn = __shfl_down_sync(0xffffffffu, hv[i], 1);
                    hv[i] = h2_bits(__hmax2(bits_h2(hv[i]), __hmax2(bits_h2(up), bits_h2(dn))));
                }
                if ((lane & 1) == 0) {
                    const int px = lane >> 1;
                    uint8_t* dst = sH + ((iter * 4 + e) & 15) * 2048 + px * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(dst + (((4 * g + c) ^ (px & 7)) << 4)) = make_uint4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
                }
                named_bar(2 + g, 128);
                // vertical half: pooled rows 2t (conv rows 4t-1..4t+1) and 2t+1 (conv rows 4t+1..4t+3); one 16-byte item per thread
                {
                    const int k = etid >> 6, px = (etid >> 2) & 15, c = 4 * g + (etid & 3);
                    const int first = k ? 1 : (t == 0 ? 0 : -1);
                    uint4 m = make_uint4(0, 0, 0, 0);   // activations are >= 0, so 0 is the identity of max
                    for (int dy = first; dy <= (k ? 3 : 1); ++dy) {
                        const uint4 r = *reinterpret_cast<const uint4*>(sH + ((iter * 4 + dy) & 15) * 2048 + px * 128 + ((c ^ (px & 7)) << 4));
                        m.x = h2_bits(__hmax2(bits_h2(m.x), bits_h2(r.x)));
                        m.y = h2_bits(__hmax2(bits_h2(m.y), bits_h2(r.y)));
                        m.z = h2_bits(__hmax2(bits_h2(m.z), bits_h2(r.z)));
                        m.w = h2_bits(__hmax2(bits_h2(m.w), bits_h2(r.w)));
                    }
                    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(sq) * 16 + 2 * t + k) * 16 + px) * 64 + c * 8) = m;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------------ MMA issuer
        const uint32_t idesc = umma_idesc_f16(128, 64, 0);
        int stage = 0, iter = 0;
        uint32_t phase = 0;
        for (int sq = blockIdx.x; sq < n_squares; sq += gridDim.x) {
            for (int t = 0; t < 8; ++t, ++iter) {
                const int acc = iter & 1;
                mbar_wait(b.tempty + 8 * acc, ((iter >> 1) & 1) ^ 1);
                mbar_wait(b.full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    mma_tile(base_addr + kRsOffA + stage * kABytes, base_addr, tmem_base + acc * 64, idesc);
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
        tmem_dealloc(tmem_base, 128);
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
    return launch_k(k_resnet_stem_tc, n_squares < sm_count ? n_squares : sm_count, kRsThreads, kRsSmem, s, true, board, static_cast<const uint4*>(wsw), bias,
                    out, n_squares);
}

cudaError_t launch_unet_stem_tc(const uint8_t* img, const void* wsw, const float* bias, const CUtensorMap* omap, int N, int sm_count,
                                cudaStream_t s) {
    const int n_units = N * 32;
    if (n_units == 0) return cudaSuccess;
    return launch_k(k_unet_stem_tc, n_units < sm_count ? n_units : sm_count, kThreads, kUsSmem, s, true, img, static_cast<const uint4*>(wsw), bias, *omap, N);
}

}  // namespace cvb
