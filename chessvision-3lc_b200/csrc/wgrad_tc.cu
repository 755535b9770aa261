// tcgen05 weight-gradient GEMM for the UNet training step (reference: the autograd backward of
// nn.Conv2d / nn.ConvTranspose2d inside scripts/train/train_unet.py:319, executed by cuDNN wgrad there).
//
//   dW[co][tap][ci] = sum over pixels p of  dz[p][co] * x[p + tap][ci]
//
// is a GEMM whose K dimension is the PIXEL index, so with NHWC activations both operands are "MN-major" (the 64
// channels of a pixel are contiguous, consecutive K rows are consecutive pixels).  A TMA box {64 ch, 16 w, 4 h} with the
// 128-byte swizzle lands in shared memory as 64 rows (pixels) of 128 bytes: exactly the canonical MN-major SWIZZLE_128B
// operand layout of tcgen05.mma (8-row groups 1024 B apart = stride byte offset; the next 64 channels are a separate
// box = leading byte offset).  The 3x3 taps and the padding are shifted box coordinates + TMA zero fill, as in the
// forward kernel; ConvTranspose uses the stride-2 parity views of its output gradient.
//
// One persistent, warp-specialised CTA per SM: warp 0 TMA producer, warp 1 MMA issuer (4 x tcgen05.mma M=128,
// N=64*NB, K=16 per 64-pixel stage, fp32 accumulators in TMEM, two accumulator buffers), warp 2 TMEM allocator,
// warps 4-7 epilogue (tcgen05.ld -> scale -> red.global.add.f32 into the fp32 gradient buffer).  Work item = (tile of
// the weight matrix, slice of the pixel range); slices of one tile are combined by the atomics.
#include "common.cuh"
#include "conv_tc.h"
#include "wgrad_tc.h"

#include <string.h>

namespace cvb {

namespace {

constexpr int kBlkBytes = 64 * 128;   // one operand block of one stage: 64 pixels x 64 channels fp16

template <int NB>
struct WgCfg {
    static constexpr int kStageBytes = (2 + NB) * kBlkBytes;
    static constexpr int kStages = NB == 1 ? 8 : (NB == 2 ? 6 : (NB == 3 ? 5 : 4));
    static constexpr int kAccCols = 64 * NB;
    static constexpr int kTmemCols = NB == 1 ? 128 : (NB == 2 ? 256 : 512);
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// MN-major operand, 128-byte swizzle: 8-pixel groups 1024 B apart (SBO), 64-channel blocks `lbo` bytes apart (LBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}

template <int NB>
__global__ void __launch_bounds__(256, 1) wgrad_tc_kernel(const __grid_constant__ WgParams p) {
    using Cfg = WgCfg<NB>;
    constexpr int S = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* tiles_ptr = smem_raw + (tiles_addr - raw_addr);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tiles_ptr + S * Cfg::kStageBytes);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * S;
    const uint32_t bar_tfull = bar_full + 16 * S;
    const uint32_t bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 6; ++i) tma_prefetch_desc(&p.maps[i]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 128);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = p.n_tiles * p.splits;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int tile = item / p.splits, sp = item - tile * p.splits;
            const int pt0 = static_cast<int>(static_cast<long long>(sp) * p.p_tiles / p.splits);
            const int pt1 = static_cast<int>(static_cast<long long>(sp + 1) * p.p_tiles / p.splits);
            const WgTile T = p.tiles[tile];
            uint32_t tx = 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) tx += T.a[i].map >= 0 ? kBlkBytes : 0;
#pragma unroll
            for (int j = 0; j < NB; ++j) tx += T.b[j].map >= 0 ? kBlkBytes : 0;
            for (int pt = pt0; pt < pt1; ++pt) {
                const int w0 = (pt % p.tiles_w) * p.tw;
                const int h0 = ((pt / p.tiles_w) % p.tiles_h) * p.th;
                const int n0 = pt / (p.tiles_w * p.tiles_h);
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    const uint32_t dst = tiles_addr + stage * Cfg::kStageBytes;
                    const uint32_t bar = bar_full + 8 * stage;
                    mbar_expect_tx(bar, tx);
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (T.a[i].map >= 0)
                            tma_load_4d(dst + i * kBlkBytes, &p.maps[T.a[i].map], bar, T.a[i].c0, w0 + T.a[i].dx, h0 + T.a[i].dy, n0);
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        if (T.b[j].map >= 0)
                            tma_load_4d(dst + (2 + j) * kBlkBytes, &p.maps[T.b[j].map], bar, T.b[j].c0, w0 + T.b[j].dx, h0 + T.b[j].dy, n0);
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
            const int tile = item / p.splits, sp = item - tile * p.splits;
            const int pt0 = static_cast<int>(static_cast<long long>(sp) * p.p_tiles / p.splits);
            const int pt1 = static_cast<int>(static_cast<long long>(sp + 1) * p.p_tiles / p.splits);
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * Cfg::kAccCols;
            for (int pt = pt0; pt < pt1; ++pt) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = tiles_addr + stage * Cfg::kStageBytes;
                    const uint64_t a_desc = umma_desc_mn_sw128(a_addr, kBlkBytes);
                    const uint64_t b_desc = umma_desc_mn_sw128(a_addr + 2 * kBlkBytes, kBlkBytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // 16 pixels (K) further = two 8-row swizzle groups = 2048 B = 128 units of 16 B
                        umma_f16(d_tmem, a_desc + 128 * k, b_desc + 128 * k, p.idesc, (pt != pt0 || k != 0) ? 1u : 0u);
                    }
                    umma_commit(bar_empty + 8 * stage);
                    if (pt == pt1 - 1) umma_commit(bar_tfull + 8 * acc);
                }
                __syncwarp();
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const int ablk = m >> 6, r = m & 63;
        int iter = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++iter) {
            const int tile = item / p.splits;
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const WgTile& T = p.tiles[tile];
            const bool a_ok = T.a[ablk].map >= 0;
            float* row_base = p.out + T.a[ablk].off + static_cast<long long>(r) * p.row_stride;
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * Cfg::kAccCols;
#pragma unroll 1
            for (int c = 0; c < 2 * NB; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c * 32, v);
                tmem_ld_wait(v);
                const WgBlock& bb = T.b[c >> 1];
                if (a_ok && bb.map >= 0) {
                    float* dst = row_base + bb.off + static_cast<long long>((c & 1) * 32) * p.col_stride;
                    if (p.col_stride == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            red_add_v4(dst + 4 * j, __uint_as_float(v[4 * j]) * p.scale, __uint_as_float(v[4 * j + 1]) * p.scale,
                                       __uint_as_float(v[4 * j + 2]) * p.scale, __uint_as_float(v[4 * j + 3]) * p.scale);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) red_add(dst + j * p.col_stride, __uint_as_float(v[j]) * p.scale);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * acc);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

template <int NB>
cudaError_t launch_nb(const WgParams& p, int grid, cudaStream_t s) {
    wgrad_tc_kernel<NB><<<grid, 256, WgCfg<NB>::kSmemBytes, s>>>(p);
    return cudaGetLastError();
}

template <int NB>
cudaError_t configure_nb() {
    return cudaFuncSetAttribute(wgrad_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgCfg<NB>::kSmemBytes);
}

}  // namespace

cudaError_t wgrad_configure() {
    cudaError_t e;
    if ((e = configure_nb<1>()) != cudaSuccess) return e;
    if ((e = configure_nb<2>()) != cudaSuccess) return e;
    if ((e = configure_nb<3>()) != cudaSuccess) return e;
    return configure_nb<4>();
}

int wgrad_tmap(CUtensorMap* m, const void* base, int C, int Wv, int Hv, int Nv, int64_t sW, int64_t sH, int64_t sN) {
    return tmap_act(m, base, C, Wv, Hv, Nv, sW, sH, sN, 16, 4, 1);
}

int wgrad_build(WgLaunch& L, const std::vector<WgBlock>& U, const std::vector<WgBlock>& V, int nb, WgTile* d_tiles, int tile_capacity,
                int N, int H, int W, float* out, long long row_stride, long long col_stride, float scale, int sm_count) {
    if (nb < 1 || nb > 4 || W % 16 || H % 4 || U.empty() || V.empty()) return -5;
    std::vector<WgTile> tiles;
    const WgBlock none = {-1, 0, 0, 0, 0, 0};
    for (size_t u = 0; u < U.size(); u += 2)
        for (size_t v = 0; v < V.size(); v += nb) {
            WgTile t;
            for (int i = 0; i < 2; ++i) t.a[i] = u + i < U.size() ? U[u + i] : none;
            for (int j = 0; j < 4; ++j) t.b[j] = (j < nb && v + j < V.size()) ? V[v + j] : none;
            tiles.push_back(t);
        }
    if (static_cast<int>(tiles.size()) > tile_capacity) return -5;
    if (cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(WgTile), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
    WgParams& p = L.p;
    p.tiles = d_tiles;
    p.n_tiles = static_cast<int>(tiles.size());
    p.tw = 16;
    p.th = 4;
    p.tiles_w = W / 16;
    p.tiles_h = H / 4;
    p.p_tiles = N * p.tiles_w * p.tiles_h;
    int splits = (3 * sm_count + p.n_tiles - 1) / p.n_tiles;   // about three items per SM
    if (splits > p.p_tiles) splits = p.p_tiles;
    if (splits < 1) splits = 1;
    p.splits = splits;
    p.out = out;
    p.row_stride = row_stride;
    p.col_stride = col_stride;
    p.scale = scale;
    p.idesc = umma_idesc_f16(128, 64 * nb, 0) | (1u << 15) | (1u << 16);   // both operands MN-major
    L.nb = nb;
    return 0;
}

cudaError_t wgrad_launch(const WgLaunch& L, int sm_count, cudaStream_t stream) {
    const WgParams& p = L.p;
    const long long items = 1LL * p.n_tiles * p.splits;
    if (items <= 0) return cudaSuccess;
    const int grid = static_cast<int>(items < sm_count ? items : sm_count);
    switch (L.nb) {
        case 1: return launch_nb<1>(p, grid, stream);
        case 2: return launch_nb<2>(p, grid, stream);
        case 3: return launch_nb<3>(p, grid, stream);
        case 4: return launch_nb<4>(p, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cvb
