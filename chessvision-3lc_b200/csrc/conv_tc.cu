// tcgen05 implicit-GEMM convolution for the UNet board extractor and the ResNet-18 piece classifier.
//
// Replaces (reference, all executed by ATen/cuDNN there):
//   Conv3x3+BN+ReLU            chessvision/pytorch_unet/unet/unet_parts.py:16-21
//   ConvTranspose2d k2 s2      chessvision/pytorch_unet/unet/unet_parts.py:53,57   (+ the concat of :67 via out_c_off)
//   Conv1x1 head + sigmoid>thr chessvision/pytorch_unet/unet/unet_parts.py:74, core.py:273, utils.py:109-112
//   ResNet BasicBlock convs    timm resnet18 (utils.py:35-39)
//
// One persistent, warp-specialised CTA per SM:
//   warp 0  TMA producer  : per K step one 4-D box {64 ch, tw, th, tn} of the NHWC activations (the 3x3 halo and the
//                           padding come from shifted box coordinates + TMA zero fill) and one {64, BLOCK_N} weight box
//   warp 1  MMA issuer    : 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per stage, fp32 accumulators in TMEM, two
//                           accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1
//   warp 2  TMEM allocator
//   warps 4-7 epilogue    : tcgen05.ld 32x32b.x32 -> bias / residual / ReLU -> fp16 NHWC (or convT scatter, or the
//                           fused 1x1 head producing logits + mask)
#include "common.cuh"
#include "conv_tc.h"
#include "launch.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace cvb {

constexpr int kEpiConstBytes = 256 * 4 + 64 * 4;   // per-tile bias (<= 256 columns) + the 1x1 head weights

// PAIR = 1: the CTA pair form (cta_group::2, see common.cuh): a stage holds this CTA's 128 pixels of A and HALF of the weight tile
template <int BLOCK_N, int PAIR = 0>
struct ConvCfg {
    static constexpr int kABytes = 128 * 128;
    static constexpr int kBBytes = BLOCK_N * 128 / (PAIR ? 2 : 1);
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = PAIR ? (BLOCK_N == 128 ? 8 : 6) : (BLOCK_N == 64 ? 8 : (BLOCK_N == 128 ? 6 : 4));
    static constexpr int kAcc = PAIR && BLOCK_N <= 128 ? 4 : 2;   // accumulator ring (see conv3x3_vr_kernel for the four-deep pair form)
    static constexpr int kTmemCols = kAcc * BLOCK_N;
    static constexpr int kOutBytes = 2 * 128 * 128;   // two staging buffers for the TMA tile store
    static constexpr int kSmemBytes = kStages * kStageBytes + kOutBytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + kEpiConstBytes;
};

constexpr int kOutBufBytes = 128 * 128;   // one staged store group: 128 pixels x 64 channels fp16

__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// 64 channels (128 B) of one pixel of the residual tensor -> 32 registers, as four 256-bit loads (every request moves whole
// 32-byte sectors).  Issued well before the accumulator is ready so that the L2 latency is off the epilogue's critical path.
__device__ __forceinline__ void res_load64(const __half* res, uint32_t (&r)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[8 * i]), "=r"(r[8 * i + 1]), "=r"(r[8 * i + 2]), "=r"(r[8 * i + 3]), "=r"(r[8 * i + 4]), "=r"(r[8 * i + 5]),
                       "=r"(r[8 * i + 6]), "=r"(r[8 * i + 7])
                     : "l"(res + 16 * i));
}

// 32 accumulator columns of one pixel -> bias (+ residual: 16 preloaded fp16 pairs) (+ ReLU) -> 16 packed fp16 pairs.
template <int RES_OFF>
__device__ __forceinline__ void pack_chunk(const uint32_t (&v)[32], const float* s_bias, const uint32_t (&res)[32], bool has_res, bool relu,
                                           uint32_t (&o)[16]) {
    const float4* b4 = reinterpret_cast<const float4*>(s_bias);
    float f[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 b = b4[i];
        f[4 * i + 0] = __uint_as_float(v[4 * i + 0]) + b.x;
        f[4 * i + 1] = __uint_as_float(v[4 * i + 1]) + b.y;
        f[4 * i + 2] = __uint_as_float(v[4 * i + 2]) + b.z;
        f[4 * i + 3] = __uint_as_float(v[4 * i + 3]) + b.w;
    }
    if (has_res) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const uint32_t rj = res[RES_OFF + j];
            const float2 rf = __half22float2(*reinterpret_cast<const __half2*>(&rj));
            f[2 * j] += rf.x;
            f[2 * j + 1] += rf.y;
        }
    }
    if (relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_h2_relu(f[2 * i], f[2 * i + 1]);
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = pack_h2(f[2 * i], f[2 * i + 1]);
    }
}

// Epilogue of one 128-row accumulator tile (thread <-> TMEM lane <-> output pixel `row` of the tile).
//   EPI_STORE / EPI_CONVT: 64 columns at a time are converted, staged in shared memory in the 128-byte-swizzled box
//   layout and written by one TMA tile store (full 128-byte lines; the pixel-shuffle of the transposed convolution is a
//   strided output view per (dy,dx)).  EPI_OUTC: fused 1x1 head, one logit + mask byte per pixel.
//   The TMEM load of chunk c+1 is in flight while chunk c is processed.
// COLS / col_base / bar_id: an epilogue group of four warps may own only the columns [col_base, col_base + COLS) of the tile
// (taddr, s_out and s_out_addr then point at that group's columns / staging buffer, bar_id is its named barrier).
template <int BLOCK_N, int EPI, int COLS = BLOCK_N>
__device__ __forceinline__ void epilogue_tile(const ConvParams& p, uint32_t taddr, int row, int n0, int h0, int w0, int n, int h, int w,
                                              bool valid, int n_tile, const float* s_bias, const float* s_outw, uint8_t* s_out,
                                              uint32_t s_out_addr, int& store_count, int etid, const uint32_t (&res0)[32], int col_base = 0,
                                              int bar_id = 1) {
    constexpr int NC = COLS / 32;
    s_bias += col_base;
    const size_t pix = (static_cast<size_t>(n) * p.H + h) * p.W + w;
    uint32_t va[32], vb[32];
    tmem_ld_32x32(taddr, va);
    if constexpr (EPI == EPI_OUTC) {
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            tmem_ld_wait(va);
            tmem_ld_32x32(taddr + (c + 1) * 32, vb);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint32_t(&v)[32] = half ? vb : va;
                if (half) {
                    tmem_ld_wait(vb);
                    if (c + 2 < NC) tmem_ld_32x32(taddr + (c + 2) * 32, va);
                }
                const float4* b4 = reinterpret_cast<const float4*>(s_bias + (c + half) * 32);
                const float4* w4 = reinterpret_cast<const float4*>(s_outw + (c + half) * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 b = b4[i];
                    const float4 wv = w4[i];
                    dot = fmaf(fmaxf(__uint_as_float(v[4 * i + 0]) + b.x, 0.f), wv.x, dot);
                    dot = fmaf(fmaxf(__uint_as_float(v[4 * i + 1]) + b.y, 0.f), wv.y, dot);
                    dot = fmaf(fmaxf(__uint_as_float(v[4 * i + 2]) + b.z, 0.f), wv.z, dot);
                    dot = fmaf(fmaxf(__uint_as_float(v[4 * i + 3]) + b.w, 0.f), wv.w, dot);
                }
            }
        }
        if (valid) {
            const float logit = dot + p.outc_b;
            p.logits[pix] = logit;
            const float prob = 1.0f / (1.0f + expf(-logit));
            p.mask[pix] = prob > p.thr ? 255 : 0;
        }
    } else {
        const bool relu = p.relu != 0;
        const bool has_res = EPI == EPI_STORE && p.res != nullptr && valid;
        const __half* res_px = has_res ? p.res + pix * p.res_c_stride + n_tile * BLOCK_N + col_base : nullptr;
        uint32_t rcur[32], rnext[32];
        if (has_res) {
#pragma unroll
            for (int i = 0; i < 32; ++i) rcur[i] = res0[i];
        }
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
            const int col0 = n_tile * BLOCK_N + col_base + c * 32;   // first of the 64 columns of this store group
            if (c + 2 < NC && has_res) res_load64(res_px + (c + 2) * 32, rnext);   // one group ahead
            uint32_t o[32];
            tmem_ld_wait(va);
            tmem_ld_32x32(taddr + (c + 1) * 32, vb);
            pack_chunk<0>(va, s_bias + c * 32, rcur, has_res, relu, reinterpret_cast<uint32_t(&)[16]>(o[0]));
            tmem_ld_wait(vb);
            if (c + 2 < NC) tmem_ld_32x32(taddr + (c + 2) * 32, va);
            pack_chunk<16>(vb, s_bias + c * 32 + 32, rcur, has_res, relu, reinterpret_cast<uint32_t(&)[16]>(o[16]));
            if (c + 2 < NC && has_res) {
#pragma unroll
                for (int i = 0; i < 32; ++i) rcur[i] = rnext[i];
            }
            if (EPI == EPI_STORE && p.pool_out != nullptr) {
                // fused MaxPool2d(2): the 2x2 window lives in one warp (rows of the tile are tw = 8 or 16 lanes apart)
                uint32_t m[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const uint32_t up = __shfl_xor_sync(0xffffffffu, o[i], p.tw);
                    const __half2 v2 = __hmax2(*reinterpret_cast<const __half2*>(&o[i]), *reinterpret_cast<const __half2*>(&up));
                    const uint32_t vu = *reinterpret_cast<const uint32_t*>(&v2);
                    const uint32_t side = __shfl_xor_sync(0xffffffffu, vu, 1);
                    const __half2 h2 = __hmax2(v2, *reinterpret_cast<const __half2*>(&side));
                    m[i] = *reinterpret_cast<const uint32_t*>(&h2);
                }
                if (valid && ((h | w) & 1) == 0) {
                    const size_t ppix = (static_cast<size_t>(n) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                    __half* pdst = p.pool_out + ppix * p.pool_c_stride + col0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) st_global_v8(pdst + 16 * q, m + 8 * q);
                }
            }
            // staging buffer: free once the store issued `out_bufs` groups ago has read it
            const int buf = p.out_bufs == 2 ? (store_count & 1) : 0;
            if (etid == 0) {
                if (p.out_bufs == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
            }
            named_bar_sync(bar_id, 128);
            uint8_t* dst = s_out + buf * kOutBufBytes + row * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            fence_proxy_async_smem();
            named_bar_sync(bar_id, 128);
            if (etid == 0) {
                if constexpr (EPI == EPI_CONVT) {
                    const int q = col0 / p.convt_cout;
                    tma_store_4d(&p.o_map[q], s_out_addr + buf * kOutBufBytes, p.out_c_off + col0 - q * p.convt_cout, w0, h0, n0);
                } else {
                    tma_store_4d(&p.o_map[0], s_out_addr + buf * kOutBufBytes, p.out_c_off + col0, w0, h0, n0);
                }
                bulk_commit();
            }
            ++store_count;
        }
    }
}

// Epilogue warps: (re)load the per-column constants of `n_tile` into shared memory (all 128 epilogue threads call this).
template <int BLOCK_N, int EPI>
__device__ __forceinline__ void epilogue_consts(const ConvParams& p, int n_tile, float* s_bias, float* s_outw, int etid) {
    named_bar_sync(1, 128);   // nobody still reads the previous tile's constants
    // transposed convolution: the four (dy,dx) column blocks share one per-channel bias (index modulo Cout)
    for (int i = etid; i < BLOCK_N; i += 128)
        s_bias[i] = __ldg(p.bias + (EPI == EPI_CONVT ? (n_tile * BLOCK_N + i) % p.convt_cout : n_tile * BLOCK_N + i));
    if constexpr (EPI == EPI_OUTC) {
        if (etid < 64) s_outw[etid] = __ldg(p.outc_w + etid);
    }
    named_bar_sync(1, 128);
}

// EPG = 2: two epilogue groups of four warps, each draining half of the tile's columns through its own staging buffer (for
// launches whose K loop is shorter than the epilogue of a whole tile: transposed convolutions, 1x1 downsamples).
template <int BLOCK_N, int EPI, int PAIR = 0, int EPG = 1>
__global__ void __launch_bounds__(128 + 128 * EPG, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BLOCK_N, PAIR>;
    constexpr int S = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t tiles_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* tiles_ptr = smem_raw + (tiles_addr - raw_addr);
    uint8_t* s_out = tiles_ptr + S * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + Cfg::kOutBytes);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * S;
    const uint32_t bar_tfull = bar_full + 16 * S;
    constexpr int kAcc = Cfg::kAcc;
    const uint32_t bar_tempty = bar_tfull + 8 * kAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * kAcc);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // CTA pair: the even CTA of the cluster issues the MMAs for both; tile index and stride count clusters, not CTAs
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int sched0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int sched_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.a_map[0]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAcc; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, PAIR ? 2 * EPG : 128 * EPG);   // pair: one arrival per CTA and group (its warps sync first)
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) tmem_alloc_pair(smem_u32(tmem_slot), Cfg::kTmemCols);
        else tmem_alloc(smem_u32(tmem_slot), Cfg::kTmemCols);
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();
    griddep_wait();   // everything above overlapped the previous kernel's tail (programmatic dependent launch)

    // pair: a scheduled tile is two consecutive M tiles (this CTA takes 2 * pair + rank) x one N tile
    const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
    const int m_units = PAIR ? (m_tiles + 1) >> 1 : m_tiles;
    const int total_tiles = m_units * p.n_tiles;
    const int k_steps = p.taps * p.c_chunks;

    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        // transaction bytes of both CTAs are counted on the leader's barrier
        const uint32_t full_sig = PAIR ? mapa_cluster(bar_full, 0) : bar_full;
        for (int t = sched0; t < total_tiles; t += sched_step) {
            const int n_tile = t % p.n_tiles;
            const int m_tile = PAIR ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
            const int w0 = (m_tile % p.tiles_w) * p.tw;
            const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * p.th;
            const int n0 = (m_tile / (p.tiles_w * p.tiles_h)) * p.tn;   // beyond the last image for the odd tile out: zero fill
            for (int tap = 0; tap < p.taps; ++tap) {
                const CUtensorMap* amap = &p.a_map[p.tap_map[tap]];
                const int hh = h0 + p.tap_dy[tap];
                const int ww = w0 + p.tap_dx[tap];
                for (int kc = 0; kc < p.c_chunks; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        const uint32_t a_dst = tiles_addr + stage * Cfg::kStageBytes;
                        const uint32_t b_dst = a_dst + Cfg::kABytes;
                        if constexpr (PAIR) {
                            if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * Cfg::kStageBytes);
                            tma_load_4d_pair(a_dst, amap, full_sig + 8 * stage, p.a_c_off + kc * 64, ww, hh, n0);
                            tma_load_2d_pair(b_dst, &p.b_map, full_sig + 8 * stage, (tap * p.c_chunks + kc) * 64,
                                             n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
                        } else {
                            mbar_expect_tx(bar_full + 8 * stage, Cfg::kStageBytes);
                            tma_load_4d(a_dst, amap, bar_full + 8 * stage, p.a_c_off + kc * 64, ww, hh, n0);
                            tma_load_2d(b_dst, &p.b_map, bar_full + 8 * stage, (tap * p.c_chunks + kc) * 64, n_tile * BLOCK_N);
                        }
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // warp-uniform role loop, tcgen05 instructions predicated on one lane elected once; descriptor words are
        // base + stage * pitch (uniform registers), so a stage costs a handful of instructions besides its four MMAs
        const bool leader = elect_one();
        const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
        const uint32_t a_lo0 = static_cast<uint32_t>(umma_desc_sw128(0)) + ((tiles_addr & 0x3FFFFu) >> 4);
        const uint32_t idesc = p.idesc;
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int t = sched0; t < total_tiles; t += sched_step, ++iter) {
            const int acc = iter % kAcc;
            const uint32_t acc_phase = (iter / kAcc) & 1;
            mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kb = 0; kb < k_steps; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (leader) {
                    const uint32_t a_lo = a_lo0 + stage * (Cfg::kStageBytes >> 4);
                    const uint32_t b_lo = a_lo + (Cfg::kABytes >> 4);
                    // advance 16 elements (32 B) along K inside the 128-byte swizzle atom: +2 in 16-byte units
                    if constexpr (PAIR) {
                        umma_f16_pair(d_tmem, desc_hi | a_lo, desc_hi | b_lo, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) umma_f16_pair(d_tmem, desc_hi | (a_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, 1u);
                        umma_commit_pair(bar_empty + 8 * stage);
                        if (kb == k_steps - 1) umma_commit_pair(bar_tfull + 8 * acc);
                    } else {
                        umma_f16(d_tmem, desc_hi | a_lo, desc_hi | b_lo, idesc, kb != 0 ? 1u : 0u);
#pragma unroll
                        for (int k = 1; k < 4; ++k) umma_f16(d_tmem, desc_hi | (a_lo + 2 * k), desc_hi | (b_lo + 2 * k), idesc, 1u);
                        umma_commit(bar_empty + 8 * stage);
                        if (kb == k_steps - 1) umma_commit(bar_tfull + 8 * acc);
                    }
                }
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        const int g = EPG == 2 ? (warp - 4) >> 2 : 0;          // epilogue group: columns [g * BLOCK_N / 2, (g + 1) * BLOCK_N / 2) when there are two
        constexpr int kGroupCols = BLOCK_N / EPG;
        const int etid = (threadIdx.x - 128) & 127;
        const int row = quarter * 32 + lane;
        const int rn = row / (p.th * p.tw);
        const int rh = (row / p.tw) % p.th;
        const int rw = row % p.tw;
        float* s_bias = reinterpret_cast<float*>(s_out + Cfg::kOutBytes + 256);
        int store_count = 0;
        float* s_outw = s_bias + 256;
        int iter = 0, cur_nt = -1;
        const uint32_t tempty_sig = PAIR ? mapa_cluster(bar_tempty, 0) : bar_tempty;
        for (int t = sched0; t < total_tiles; t += sched_step, ++iter) {
            const int acc = iter % kAcc;
            const uint32_t acc_phase = (iter / kAcc) & 1;
            const int n_tile = t % p.n_tiles;
            const int m_tile = PAIR ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
            if (n_tile != cur_nt) {
                if constexpr (EPG == 2) {   // both groups share the per-column constants
                    named_bar_sync(3, 256);
                    for (int i = threadIdx.x - 128; i < BLOCK_N; i += 256)
                        s_bias[i] = __ldg(p.bias + (EPI == EPI_CONVT ? (n_tile * BLOCK_N + i) % p.convt_cout : n_tile * BLOCK_N + i));
                    named_bar_sync(3, 256);
                } else {
                    epilogue_consts<BLOCK_N, EPI>(p, n_tile, s_bias, s_outw, threadIdx.x - 128);
                }
                cur_nt = n_tile;
            }
            const int w = (m_tile % p.tiles_w) * p.tw + rw;
            const int h = ((m_tile / p.tiles_w) % p.tiles_h) * p.th + rh;
            const int n = (m_tile / (p.tiles_w * p.tiles_h)) * p.tn + rn;
            const bool valid = n < p.N;
            uint32_t res0[32];   // first 64 residual channels of this pixel, requested before the accumulator is waited for
            if (EPI == EPI_STORE && p.res != nullptr && valid)
                res_load64(p.res + ((static_cast<size_t>(n) * p.H + h) * p.W + w) * p.res_c_stride + n_tile * BLOCK_N + g * kGroupCols, res0);
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N + g * kGroupCols;
            if constexpr (EPG == 2)
                epilogue_tile<BLOCK_N, EPI, kGroupCols>(p, taddr, row, n - rn, h - rh, w - rw, n, h, w, valid, n_tile, s_bias, s_outw,
                                                        s_out + g * kOutBufBytes, tiles_addr + S * Cfg::kStageBytes + g * kOutBufBytes, store_count, etid,
                                                        res0, g * kGroupCols, 1 + g);
            else
                epilogue_tile<BLOCK_N, EPI>(p, taddr, row, n - rn, h - rh, w - rw, n, h, w, valid, n_tile, s_bias, s_outw, s_out,
                                            tiles_addr + S * Cfg::kStageBytes, store_count, etid, res0);
            tc_fence_before();
            if constexpr (PAIR) {
                named_bar_sync(1 + g, 128);   // all four warps have read their lanes of the accumulator
                if (etid == 0) mbar_arrive_cluster(tempty_sig + 8 * acc);
            } else {
                mbar_arrive(bar_tempty + 8 * acc);
            }
        }
        if (etid == 0) bulk_wait_all();   // the staged tiles must have left shared memory before the CTA exits
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();   // neither CTA leaves while the other may still signal its barriers or read its B half
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
        else tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// conv3x3(Cin -> 128) + BN + ReLU fused with the ConvTranspose2d(128 -> 64, k2, s2) that consumes it (UNet: up3.conv.3 and
// up4.up, unet_parts.py:16-21,53).  Unfused, the 128-channel tensor is written to HBM (4 MB per board) only to be read back by a
// transposed-convolution kernel that is bound by exactly that traffic.  Here the epilogue converts the accumulator tile
// (128 pixels x 128 channels) to fp16 in the 128-byte-swizzled K-major layout -- the same bytes the unfused kernel would have
// stored -- and leaves it in shared memory as the A operand of a second MMA against the transposed-conv weights
// (N = 4 taps x 64 channels = 256, K = 128); its accumulator (256 TMEM columns next to the two 128-column main accumulators)
// is drained, biased and scattered by the same four epilogue warps through the strided output views of EPI_CONVT.  The
// results equal those of the two separate kernels (same fp16 rounding point, same accumulation order as conv3x3_vr_kernel).
//   warp 0 TMA producer, warp 1 MMA issue (main K loop of tile i, then the second MMA of tile i-1, whose A operand the
//   epilogue produced meanwhile), warp 2 TMEM, warps 4-11 epilogue (two groups).
// ---------------------------------------------------------------------------------------------------------------------
struct FusedCfg {
    // main loop in the vertical-reuse form (see conv3x3_vr_kernel): one activation box {64 ch, 8 w, 18 h} per (K chunk, dx)
    // serves the three vertical taps, whose weight tiles flow through a ring of their own.  The transposed-conv weights
    // (64 KB) travel through the same ring as four more 16 KB tiles per output tile instead of staying resident: that buys
    // a ring of nine tiles (2,300 tensor-clocks of look-ahead, what the TMA pipeline needs to cover the L2 latency).
    static constexpr int kAStages = 2, kBStages = 9;
    static constexpr int kABytes = 18 * 1024, kBBytes = 128 * 128;
    static constexpr int kA2Bytes = 2 * 128 * 128;   // two K blocks of [128 pixels][64 k]; reused as the two store staging buffers
    static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBBytes + kA2Bytes + 1024 + 256 + (128 + 64) * 4;
};
constexpr int kFusedThreads = 384;   // TMA, MMA, TMEM, idle warp + two epilogue groups of four warps

__global__ void __launch_bounds__(kFusedThreads, 1) conv_convt_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = FusedCfg;
    constexpr int SA = Cfg::kAStages, SB = Cfg::kBStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t a_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (a_addr - raw_addr);
    const uint32_t b_addr = a_addr + SA * Cfg::kABytes;
    const uint32_t a2_addr = b_addr + SB * Cfg::kBBytes;
    uint8_t* a2_ptr = base_ptr + SA * Cfg::kABytes + SB * Cfg::kBBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a2_ptr + Cfg::kA2Bytes);
    const uint32_t bar_afull = smem_u32(bars);
    const uint32_t bar_aempty = bar_afull + 8 * SA;
    const uint32_t bar_bfull = bar_aempty + 8 * SA;
    const uint32_t bar_bempty = bar_bfull + 8 * SB;
    const uint32_t bar_tfull = bar_bempty + 8 * SB;
    const uint32_t bar_tempty = bar_tfull + 16;
    const uint32_t bar_a2full = bar_tfull + 32;
    const uint32_t bar_d2full = bar_tfull + 40;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 8);
    float* s_bias = reinterpret_cast<float*>(a2_ptr + Cfg::kA2Bytes + 256);
    float* s_bias2 = s_bias + 128;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.b2_map);
        tma_prefetch_desc(&p.a_map[0]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) {
            mbar_init(bar_afull + 8 * i, 1);
            mbar_init(bar_aempty + 8 * i, 1);
        }
        for (int i = 0; i < SB; ++i) {
            mbar_init(bar_bfull + 8 * i, 1);
            mbar_init(bar_bempty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 256);
        }
        mbar_init(bar_a2full, 256);
        mbar_init(bar_d2full, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    if (warp == 3) {
        for (int i = lane; i < 128; i += 32) s_bias[i] = __ldg(p.bias + i);
        for (int i = lane; i < 64; i += 32) s_bias2[i] = __ldg(p.bias2 + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d2_tmem = tmem_base + 256;
    griddep_launch();
    griddep_wait();

    const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w;   // one N tile: all 128 channels of a pixel in this CTA
    // Order of the weight ring, identical in the producer and the MMA warp: the 18 tiles of the main loop of tile i, then
    // (from the second tile on) the four transposed-conv tiles (row half rh, K block kb) for the second MMA of tile i-1;
    // after the last tile four more for its own second MMA.

    if (warp == 0) {
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        auto load_w2 = [&]() {
            for (int j = 0; j < 4; ++j) {   // j = rh * 2 + kb: rows [128 rh, 128 rh + 128) of the 256 (dy,dx,co) rows, K block kb
                mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(bar_bfull + 8 * bs, Cfg::kBBytes);
                    tma_load_2d(b_addr + bs * Cfg::kBBytes, &p.b2_map, bar_bfull + 8 * bs, (j & 1) * 64, (j >> 1) * 128);
                }
                __syncwarp();
                if (++bs == SB) { bs = 0; bph ^= 1; }
            }
        };
        int iter = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
            const int w0 = (t % p.tiles_w) * 8;
            const int h0 = ((t / p.tiles_w) % p.tiles_h) * 16;
            const int n0 = t / (p.tiles_w * p.tiles_h);
            for (int kc = 0; kc < p.c_chunks; ++kc) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(bar_aempty + 8 * as, aph ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(bar_afull + 8 * as, Cfg::kABytes);
                        tma_load_4d(a_addr + as * Cfg::kABytes, &p.a_map[0], bar_afull + 8 * as, p.a_c_off + kc * 64, w0 + dxi - 1, h0 - 1, n0);
                    }
                    __syncwarp();
                    if (++as == SA) { as = 0; aph ^= 1; }
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
                        if (elect_one()) {
                            mbar_expect_tx(bar_bfull + 8 * bs, Cfg::kBBytes);
                            tma_load_2d(b_addr + bs * Cfg::kBBytes, &p.b_map, bar_bfull + 8 * bs, ((dy * 3 + dxi) * p.c_chunks + kc) * 64, 0);
                        }
                        __syncwarp();
                        if (++bs == SB) { bs = 0; bph ^= 1; }
                    }
                }
            }
            if (iter > 0) load_w2();
        }
        if (iter > 0) load_w2();
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_sw128(0));
        const uint32_t a_lo0 = desc_lo0 + ((a_addr & 0x3FFFFu) >> 4);
        const uint32_t b_lo0 = desc_lo0 + ((b_addr & 0x3FFFFu) >> 4);
        const uint32_t a2_lo = desc_lo0 + ((a2_addr & 0x3FFFFu) >> 4);
        const uint32_t idesc = umma_idesc_f16(128, 128, 0);
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        // second MMA of tile j: D2[128 px][(dy,dx,co)] = A2[128 px][128 ci] . W2^T as four N = 128 pieces (row half rh of W2,
        // K block kb), each fed by one tile of the weight ring
        auto mma2 = [&](int j) {
            mbar_wait(bar_a2full, j & 1);
            for (int q = 0; q < 4; ++q) {
                mbar_wait(bar_bfull + 8 * bs, bph);
                tc_fence_after();
                if (leader) {
                    const int rh = q >> 1, kb = q & 1;
                    const uint32_t b_lo = b_lo0 + bs * (Cfg::kBBytes >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(d2_tmem + rh * 128, desc_hi | (a2_lo + kb * (128 * 128 >> 4) + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(bar_bempty + 8 * bs);
                    if (q == 3) umma_commit(bar_d2full);
                }
                __syncwarp();
                if (++bs == SB) { bs = 0; bph ^= 1; }
            }
        };
        int iter = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
            const int acc = iter & 1;
            mbar_wait(bar_tempty + 8 * acc, ((iter >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 128;
            for (int kc = 0; kc < p.c_chunks; ++kc) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(bar_afull + 8 * as, aph);
                    const uint32_t a_lo = a_lo0 + as * (Cfg::kABytes >> 4);
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(bar_bfull + 8 * bs, bph);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t b_lo = b_lo0 + bs * (Cfg::kBBytes >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // tap dy: the same box one 8-row swizzle group (1024 B) further
                                umma_f16(d_tmem, desc_hi | (a_lo + dy * 64 + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (kc | dxi | dy | k) != 0 ? 1u : 0u);
                            umma_commit(bar_bempty + 8 * bs);
                        }
                        __syncwarp();
                        if (++bs == SB) { bs = 0; bph ^= 1; }
                    }
                    if (leader) {
                        umma_commit(bar_aempty + 8 * as);
                        if (kc == p.c_chunks - 1 && dxi == 2) umma_commit(bar_tfull + 8 * acc);
                    }
                    __syncwarp();
                    if (++as == SA) { as = 0; aph ^= 1; }
                }
            }
            if (iter > 0) mma2(iter - 1);
        }
        if (iter > 0) mma2(iter - 1);
    } else if (warp >= 4) {
        // Two epilogue groups (warps 4-7 / 8-11) work on the SAME tile: group g converts channels 64g .. 64g+63 of the
        // accumulator into K block g of the A operand, then drains taps q = 2g, 2g+1 of the transposed-conv accumulator and
        // stores them from "its" K block (free again once the second MMA has completed).  That halves the chain
        // accumulator -> A operand -> second MMA -> scatter, which otherwise outlasts the main K loop of the next tile.
        const int quarter = warp & 3, g = (warp - 4) >> 2, etid = (threadIdx.x - 128) & 127;
        const int row = quarter * 32 + lane;   // = 8 * (row of the 16 x 8 tile) + column
        const uint32_t no_res[32] = {0};
        uint8_t* my_buf = a2_ptr + g * (128 * 128);
        const uint32_t my_buf_addr = a2_addr + g * (128 * 128);
        const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
        int iter = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++iter) {
            const int acc = iter & 1;
            const int w0 = (t % p.tiles_w) * 8;
            const int h0 = ((t / p.tiles_w) % p.tiles_h) * 16;
            const int n0 = t / (p.tiles_w * p.tiles_h);
            mbar_wait(bar_tfull + 8 * acc, (iter >> 1) & 1);
            tc_fence_after();
            // ---- part 1: accumulator -> bias + ReLU -> fp16 -> K block g of the second MMA's A operand
            if (etid == 0) bulk_wait_read<0>();   // this group's stores of the previous tile have finished reading the block
            named_bar_sync(1 + g, 128);
            {
                uint32_t va[32], vb[32], o[32];
                tmem_ld_32x32(tmem_base + lane_addr + acc * 128 + g * 64, va);
                tmem_ld_32x32(tmem_base + lane_addr + acc * 128 + g * 64 + 32, vb);
                tmem_ld_wait(va);
                pack_chunk<0>(va, s_bias + g * 64, no_res, false, true, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                tmem_ld_wait(vb);
                tc_fence_before();
                mbar_arrive(bar_tempty + 8 * acc);          // the accumulator half is in registers: release it early
                pack_chunk<16>(vb, s_bias + g * 64 + 32, no_res, false, true, reinterpret_cast<uint32_t(&)[16]>(o[16]));
                uint8_t* dst = my_buf + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_a2full);
            // ---- part 2: taps q = 2g, 2g+1 of the transposed-conv accumulator -> + bias -> fp16 -> strided tile stores
            mbar_wait(bar_d2full, iter & 1);
            tc_fence_after();
            {
                uint32_t va[32], vb[32];
                tmem_ld_32x32(d2_tmem + lane_addr + (2 * g) * 64, va);
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const int q = 2 * g + qq;
                    uint32_t o[32];
                    tmem_ld_wait(va);
                    tmem_ld_32x32(d2_tmem + lane_addr + q * 64 + 32, vb);
                    pack_chunk<0>(va, s_bias2, no_res, false, false, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                    tmem_ld_wait(vb);
                    if (qq == 0) tmem_ld_32x32(d2_tmem + lane_addr + (q + 1) * 64, va);
                    pack_chunk<16>(vb, s_bias2 + 32, no_res, false, false, reinterpret_cast<uint32_t(&)[16]>(o[16]));
                    // block free: qq = 0 -> the second MMA has consumed it (d2full); qq = 1 -> this group's first store has read it
                    if (qq == 1 && etid == 0) bulk_wait_read<0>();
                    named_bar_sync(1 + g, 128);
                    uint8_t* dst = my_buf + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    fence_proxy_async_smem();
                    named_bar_sync(1 + g, 128);
                    if (etid == 0) {
                        tma_store_4d(&p.o_map[q], my_buf_addr, p.out_c_off, w0, h0, n0);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
        }
        if (etid == 0) bulk_wait_all();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// conv_convt_kernel as a CTA pair (cta_group::2): two 16 x 8 tiles per MMA, each CTA holds its own activation boxes, its own
// A operand of the second MMA, and HALF of every weight tile (64 of the 128 rows; the transposed-conv pieces likewise).  The
// single-CTA form pulls 66 KB per 768 tensor-clocks from L2 into every SM (~19 TB/s over the chip -- the limit; its MMA warp
// spent half of its time waiting for weight tiles, ncu r02); here it is 42 KB, and an MMA reads 4 + 2 KB of operands from
// shared memory instead of 4 + 4.  Barriers the leader's MMA warp waits on (afull, bfull, a2full, tempty) live in the leader
// and are signalled by both CTAs; barriers the two CTAs' producers / epilogues wait on are signalled by multicast commits.
// ---------------------------------------------------------------------------------------------------------------------
template <int SA_, int SB_>
struct FusedPairCfg {
    static constexpr int kAStages = SA_, kBStages = SB_;
    static constexpr int kABytes = 18 * 1024, kBBytes = 64 * 128;
    static constexpr int kA2Bytes = 2 * 128 * 128;
    static constexpr int kBarBytes = 512;
    static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBBytes + kA2Bytes + 1024 + kBarBytes + (128 + 64) * 4;
};

template <int SA_, int SB_>
__global__ void __launch_bounds__(kFusedThreads, 1) conv_convt_pair_kernel(const __grid_constant__ ConvParams p) {
    using Cfg = FusedPairCfg<SA_, SB_>;
    constexpr int SA = Cfg::kAStages, SB = Cfg::kBStages;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t a_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (a_addr - raw_addr);
    const uint32_t b_addr = a_addr + SA * Cfg::kABytes;
    const uint32_t a2_addr = b_addr + SB * Cfg::kBBytes;
    uint8_t* a2_ptr = base_ptr + SA * Cfg::kABytes + SB * Cfg::kBBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a2_ptr + Cfg::kA2Bytes);
    const uint32_t bar_afull = smem_u32(bars);
    const uint32_t bar_aempty = bar_afull + 8 * SA;
    const uint32_t bar_bfull = bar_aempty + 8 * SA;
    const uint32_t bar_bempty = bar_bfull + 8 * SB;
    const uint32_t bar_tfull = bar_bempty + 8 * SB;
    const uint32_t bar_tempty = bar_tfull + 16;
    const uint32_t bar_a2full = bar_tfull + 32;
    const uint32_t bar_d2full = bar_tfull + 40;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 8);
    float* s_bias = reinterpret_cast<float*>(a2_ptr + Cfg::kA2Bytes + Cfg::kBarBytes);
    float* s_bias2 = s_bias + 128;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.b2_map);
        tma_prefetch_desc(&p.a_map[0]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < SA; ++i) {
            mbar_init(bar_afull + 8 * i, 1);
            mbar_init(bar_aempty + 8 * i, 1);
        }
        for (int i = 0; i < SB; ++i) {
            mbar_init(bar_bfull + 8 * i, 1);
            mbar_init(bar_bempty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 512);   // the 256 epilogue threads of both CTAs
        }
        mbar_init(bar_a2full, 512);
        mbar_init(bar_d2full, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc_pair(smem_u32(tmem_slot), 512);
    if (warp == 3) {
        for (int i = lane; i < 128; i += 32) s_bias[i] = __ldg(p.bias + i);
        for (int i = lane; i < 64; i += 32) s_bias2[i] = __ldg(p.bias2 + i);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d2_tmem = tmem_base + 256;
    griddep_launch();
    griddep_wait();

    const int total_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
    const int total_units = (total_tiles + 1) >> 1;
    const int unit0 = static_cast<int>(blockIdx.x >> 1), unit_step = static_cast<int>(gridDim.x >> 1);

    if (warp == 0) {
        const uint32_t afull_sig = mapa_cluster(bar_afull, 0), bfull_sig = mapa_cluster(bar_bfull, 0);
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        auto load_w2 = [&]() {
            for (int j = 0; j < 4; ++j) {   // j = rh * 2 + kb: this CTA's 64 of the rows [128 rh, 128 rh + 128) of W2, K block kb
                mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(bar_bfull + 8 * bs, 2 * Cfg::kBBytes);
                    tma_load_2d_pair(b_addr + bs * Cfg::kBBytes, &p.b2_map, bfull_sig + 8 * bs, (j & 1) * 64, (j >> 1) * 128 + static_cast<int>(rank) * 64);
                }
                __syncwarp();
                if (++bs == SB) { bs = 0; bph ^= 1; }
            }
        };
        int iter = 0;
        for (int u = unit0; u < total_units; u += unit_step, ++iter) {
            const int t = 2 * u + static_cast<int>(rank);
            const int w0 = (t % p.tiles_w) * 8;
            const int h0 = ((t / p.tiles_w) % p.tiles_h) * 16;
            const int n0 = t / (p.tiles_w * p.tiles_h);
            for (int kc = 0; kc < p.c_chunks; ++kc) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(bar_aempty + 8 * as, aph ^ 1);
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(bar_afull + 8 * as, 2 * Cfg::kABytes);
                        tma_load_4d_pair(a_addr + as * Cfg::kABytes, &p.a_map[0], afull_sig + 8 * as, p.a_c_off + kc * 64, w0 + dxi - 1, h0 - 1, n0);
                    }
                    __syncwarp();
                    if (++as == SA) { as = 0; aph ^= 1; }
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(bar_bempty + 8 * bs, bph ^ 1);
                        if (elect_one()) {
                            if (rank == 0) mbar_expect_tx(bar_bfull + 8 * bs, 2 * Cfg::kBBytes);
                            tma_load_2d_pair(b_addr + bs * Cfg::kBBytes, &p.b_map, bfull_sig + 8 * bs, ((dy * 3 + dxi) * p.c_chunks + kc) * 64,
                                             static_cast<int>(rank) * 64);
                        }
                        __syncwarp();
                        if (++bs == SB) { bs = 0; bph ^= 1; }
                    }
                }
            }
            if (iter > 0) load_w2();
        }
        if (iter > 0) load_w2();
    } else if (warp == 1 && rank == 0) {
        const bool leader = elect_one();
        const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_sw128(0));
        const uint32_t a_lo0 = desc_lo0 + ((a_addr & 0x3FFFFu) >> 4);
        const uint32_t b_lo0 = desc_lo0 + ((b_addr & 0x3FFFFu) >> 4);
        const uint32_t a2_lo = desc_lo0 + ((a2_addr & 0x3FFFFu) >> 4);
        const uint32_t idesc = umma_idesc_f16(256, 128, 0);
        int as = 0, bs = 0;
        uint32_t aph = 0, bph = 0;
        // second MMA of unit j: D2[256 px][(dy,dx,co)] = A2[256 px][128 ci] . W2^T as four N = 128 pieces (row half rh of W2,
        // K block kb), each fed by one tile of the weight ring (64 rows per CTA)
        auto mma2 = [&](int j) {
            mbar_wait_cluster(bar_a2full, j & 1);
            for (int q = 0; q < 4; ++q) {
                mbar_wait(bar_bfull + 8 * bs, bph);
                tc_fence_after();
                if (leader) {
                    const int rh = q >> 1, kb = q & 1;
                    const uint32_t b_lo = b_lo0 + bs * (Cfg::kBBytes >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_pair(d2_tmem + rh * 128, desc_hi | (a2_lo + kb * (128 * 128 >> 4) + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_pair(bar_bempty + 8 * bs);
                    if (q == 3) umma_commit_pair(bar_d2full);
                }
                __syncwarp();
                if (++bs == SB) { bs = 0; bph ^= 1; }
            }
        };
        int iter = 0;
        for (int u = unit0; u < total_units; u += unit_step, ++iter) {
            const int acc = iter & 1;
            mbar_wait(bar_tempty + 8 * acc, ((iter >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 128;
            for (int kc = 0; kc < p.c_chunks; ++kc) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(bar_afull + 8 * as, aph);
                    const uint32_t a_lo = a_lo0 + as * (Cfg::kABytes >> 4);
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(bar_bfull + 8 * bs, bph);
                        tc_fence_after();
                        if (leader) {
                            const uint32_t b_lo = b_lo0 + bs * (Cfg::kBBytes >> 4);
#pragma unroll
                            for (int k = 0; k < 4; ++k)   // tap dy: the same box one 8-row swizzle group (1024 B) further
                                umma_f16_pair(d_tmem, desc_hi | (a_lo + dy * 64 + 2 * k), desc_hi | (b_lo + 2 * k), idesc, (kc | dxi | dy | k) != 0 ? 1u : 0u);
                            umma_commit_pair(bar_bempty + 8 * bs);
                        }
                        __syncwarp();
                        if (++bs == SB) { bs = 0; bph ^= 1; }
                    }
                    if (leader) {
                        umma_commit_pair(bar_aempty + 8 * as);
                        if (kc == p.c_chunks - 1 && dxi == 2) umma_commit_pair(bar_tfull + 8 * acc);
                    }
                    __syncwarp();
                    if (++as == SA) { as = 0; aph ^= 1; }
                }
            }
            if (iter > 0) mma2(iter - 1);
        }
        if (iter > 0) mma2(iter - 1);
    } else if (warp >= 4) {
        // Two epilogue groups (warps 4-7 / 8-11) work on this CTA's tile exactly as in conv_convt_kernel.
        const int quarter = warp & 3, g = (warp - 4) >> 2, etid = (threadIdx.x - 128) & 127;
        const int row = quarter * 32 + lane;   // = 8 * (row of the 16 x 8 tile) + column
        const uint32_t no_res[32] = {0};
        uint8_t* my_buf = a2_ptr + g * (128 * 128);
        const uint32_t my_buf_addr = a2_addr + g * (128 * 128);
        const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tempty_sig = mapa_cluster(bar_tempty, 0), a2full_sig = mapa_cluster(bar_a2full, 0);
        int iter = 0;
        for (int u = unit0; u < total_units; u += unit_step, ++iter) {
            const int acc = iter & 1;
            const int t = 2 * u + static_cast<int>(rank);
            const int w0 = (t % p.tiles_w) * 8;
            const int h0 = ((t / p.tiles_w) % p.tiles_h) * 16;
            const int n0 = t / (p.tiles_w * p.tiles_h);
            mbar_wait(bar_tfull + 8 * acc, (iter >> 1) & 1);
            tc_fence_after();
            // ---- part 1: accumulator -> bias + ReLU -> fp16 -> K block g of the second MMA's A operand
            if (etid == 0) bulk_wait_read<0>();   // this group's stores of the previous tile have finished reading the block
            named_bar_sync(1 + g, 128);
            {
                uint32_t va[32], vb[32], o[32];
                tmem_ld_32x32(tmem_base + lane_addr + acc * 128 + g * 64, va);
                tmem_ld_32x32(tmem_base + lane_addr + acc * 128 + g * 64 + 32, vb);
                tmem_ld_wait(va);
                pack_chunk<0>(va, s_bias + g * 64, no_res, false, true, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                tmem_ld_wait(vb);
                tc_fence_before();
                mbar_arrive_cluster(tempty_sig + 8 * acc);          // the accumulator half is in registers: release it early
                pack_chunk<16>(vb, s_bias + g * 64 + 32, no_res, false, true, reinterpret_cast<uint32_t(&)[16]>(o[16]));
                uint8_t* dst = my_buf + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            }
            fence_proxy_async_smem();
            mbar_arrive_cluster(a2full_sig);
            // ---- part 2: taps q = 2g, 2g+1 of the transposed-conv accumulator -> + bias -> fp16 -> strided tile stores
            mbar_wait(bar_d2full, iter & 1);
            tc_fence_after();
            {
                uint32_t va[32], vb[32];
                tmem_ld_32x32(d2_tmem + lane_addr + (2 * g) * 64, va);
#pragma unroll
                for (int qq = 0; qq < 2; ++qq) {
                    const int q = 2 * g + qq;
                    uint32_t o[32];
                    tmem_ld_wait(va);
                    tmem_ld_32x32(d2_tmem + lane_addr + q * 64 + 32, vb);
                    pack_chunk<0>(va, s_bias2, no_res, false, false, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                    tmem_ld_wait(vb);
                    if (qq == 0) tmem_ld_32x32(d2_tmem + lane_addr + (q + 1) * 64, va);
                    pack_chunk<16>(vb, s_bias2 + 32, no_res, false, false, reinterpret_cast<uint32_t(&)[16]>(o[16]));
                    // block free: qq = 0 -> the second MMA has consumed it (d2full); qq = 1 -> this group's first store has read it
                    if (qq == 1 && etid == 0) bulk_wait_read<0>();
                    named_bar_sync(1 + g, 128);
                    uint8_t* dst = my_buf + row * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(dst + ((j ^ (row & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
                    fence_proxy_async_smem();
                    named_bar_sync(1 + g, 128);
                    if (etid == 0) {
                        tma_store_4d(&p.o_map[q], my_buf_addr, p.out_c_off, w0, h0, n0);
                        bulk_commit();
                    }
                }
            }
            tc_fence_before();
        }
        if (etid == 0) bulk_wait_all();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 variant with vertical-tap reuse (and optionally stationary weights) for the wide, shallow layers
// (Cout = 64 or 128 at 256^2 / 128^2), which are L2->SM bandwidth bound in the generic kernel: every tap re-fetches
// the same activations.  Here one TMA box {64 ch, 8 w, 18 h} (18 swizzle groups of 1024 B) per horizontal offset dx
// serves the three vertical taps: tap dy is the same smem tile addressed 1024*dy bytes further (a whole 8-row group,
// so the 128-byte swizzle phase is unchanged).  Output tile = 16 rows x 8 columns.  When the layer's packed weights
// (9*Cin*Cout*2 B) fit beside the pipeline they are loaded once per CTA and stay in shared memory.
// L2->SM bytes per 128x64 output tile, Cin = 64:  generic 216 KB  ->  54 KB.
// ---------------------------------------------------------------------------------------------------------------------
// EPG = 2 (N = 128): two epilogue groups of four warps, each draining 64 of the 128 columns of every tile through its own
// staging buffer.  With a short K loop (Cin <= 128) one group cannot drain a 128-column tile in the time the MMAs need for
// the next one (3,360 clocks per tile against 2,304 of MMA work at Cin = 64; 6,700 with the fused max-pool at Cin = 128).
template <int BLOCK_N, int EPI, bool W_STAT, int PAIR = 0, int EPG = 1>
__global__ void __launch_bounds__(128 + 128 * EPG, 1) conv3x3_vr_kernel(const __grid_constant__ ConvParams p) {
    // PAIR = 1: CTA pair (cta_group::2, see conv_tc_kernel): two 16 x 8 output tiles per MMA, each CTA holds half of every
    // weight tile (BLOCK_N / 2 rows), so an MMA reads 4 KB of A + 2 KB of B per SM instead of 4 + 4 (the N = 128 layers are
    // bound by exactly that shared-memory operand bandwidth) and twice the input channels fit as resident weights.
    constexpr int kABytes = 18 * 1024;
    constexpr int kBBytes = BLOCK_N * 128 / (PAIR ? 2 : 1);
    constexpr int kStageBytes = kABytes + (W_STAT ? 0 : 3 * kBBytes);
    // accumulator ring (a four-deep ring for N = 128 pairs was measured: no gain on Cin >= 128, and Cin = 64 -- where the
    // cross-CTA hand-over of an accumulator outlasts the short K loop -- stays slower than the single-CTA form either way)
    constexpr int kAcc = 2;
    constexpr int kTmemCols = kAcc * BLOCK_N;
    const int S = p.vr_stages;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base_addr - raw_addr);
    const uint32_t w_bytes = W_STAT ? 9u * p.c_chunks * kBBytes : 0u;
    const uint32_t stages_addr = base_addr + w_bytes;
    uint8_t* s_out = base_ptr + w_bytes + S * kStageBytes;
    const uint32_t out_bytes = EPI == EPI_OUTC ? 0u : static_cast<uint32_t>(EPG * kOutBufBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_out + out_bytes);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * S;
    const uint32_t bar_tfull = bar_full + 16 * S;
    const uint32_t bar_tempty = bar_tfull + 8 * kAcc;
    const uint32_t bar_w = bar_tfull + 16 * kAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 2 * kAcc + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int sched0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int sched_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.a_map[0]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < kAcc; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, PAIR ? 2 * EPG : 128 * EPG);
        }
        mbar_init(bar_w, 1);
        mbar_fence_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) tmem_alloc_pair(smem_u32(tmem_slot), kTmemCols);
        else tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();
    if (warp != 0) griddep_wait();   // warp 0 first requests the resident weights (no kernel writes them)

    const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
    const int m_units = PAIR ? (m_tiles + 1) >> 1 : m_tiles;
    const int total_tiles = m_units * p.n_tiles;

    if (warp == 0) {
        // pair: the bytes of both CTAs are counted on the leader's barriers (its MMA warp is the only consumer)
        const uint32_t full_sig = PAIR ? mapa_cluster(bar_full, 0) : bar_full;
        if (W_STAT && elect_one()) {
            if constexpr (PAIR) {
                const uint32_t w_sig = mapa_cluster(bar_w, 0);
                if (rank == 0) mbar_expect_tx(bar_w, 2 * w_bytes);
                for (int i = 0; i < 9 * p.c_chunks; ++i)
                    tma_load_2d_pair(base_addr + i * kBBytes, &p.b_map, w_sig, i * 64, static_cast<int>(rank) * (BLOCK_N / 2));
            } else {
                mbar_expect_tx(bar_w, w_bytes);
                for (int i = 0; i < 9 * p.c_chunks; ++i) tma_load_2d(base_addr + i * kBBytes, &p.b_map, bar_w, i * 64, 0);
            }
        }
        __syncwarp();
        griddep_wait();
        int stage = 0;
        uint32_t phase = 0;
        for (int t = sched0; t < total_tiles; t += sched_step) {
            const int n_tile = t % p.n_tiles;
            const int m_tile = PAIR ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
            const int w0 = (m_tile % p.tiles_w) * 8;
            const int h0 = ((m_tile / p.tiles_w) % p.tiles_h) * 16;
            const int n0 = m_tile / (p.tiles_w * p.tiles_h);
            for (int kc = 0; kc < p.c_chunks; ++kc) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (elect_one()) {
                        const uint32_t a_dst = stages_addr + stage * kStageBytes;
                        if constexpr (PAIR) {
                            if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * kStageBytes);
                            tma_load_4d_pair(a_dst, &p.a_map[0], full_sig + 8 * stage, p.a_c_off + kc * 64, w0 + dxi - 1, h0 - 1, n0);
                            if (!W_STAT) {
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy)
                                    tma_load_2d_pair(a_dst + kABytes + dy * kBBytes, &p.b_map, full_sig + 8 * stage,
                                                     ((dy * 3 + dxi) * p.c_chunks + kc) * 64,
                                                     n_tile * BLOCK_N + static_cast<int>(rank) * (BLOCK_N / 2));
                            }
                        } else {
                            mbar_expect_tx(bar_full + 8 * stage, kStageBytes);
                            tma_load_4d(a_dst, &p.a_map[0], bar_full + 8 * stage, p.a_c_off + kc * 64, w0 + dxi - 1, h0 - 1, n0);
                            if (!W_STAT) {
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy)
                                    tma_load_2d(a_dst + kABytes + dy * kBBytes, &p.b_map, bar_full + 8 * stage,
                                                ((dy * 3 + dxi) * p.c_chunks + kc) * 64, n_tile * BLOCK_N);
                            }
                        }
                    }
                    __syncwarp();
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        const bool leader = elect_one();   // warp-uniform loop, tcgen05 instructions predicated on one lane (see conv_tc_kernel)
        if (W_STAT) {
            mbar_wait(bar_w, 0);
            tc_fence_after();
        }
        const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_sw128(0));
        const uint32_t a_lo0 = desc_lo0 + ((stages_addr & 0x3FFFFu) >> 4);
        const uint32_t w_lo0 = desc_lo0 + ((base_addr & 0x3FFFFu) >> 4);
        // tap dy: the A tile 1024 B (one 8-row swizzle group) further, the weight tile 3*c_chunks (or 1) tiles further
        const uint32_t b_step = (W_STAT ? 3u * p.c_chunks * kBBytes : static_cast<uint32_t>(kBBytes)) >> 4;
        const uint32_t idesc = p.idesc;
        int stage = 0;
        uint32_t phase = 0;
        int iter = 0;
        for (int t = sched0; t < total_tiles; t += sched_step, ++iter) {
            const int acc = iter % kAcc;
            const uint32_t acc_phase = (iter / kAcc) & 1;
            mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            for (int kc = 0; kc < p.c_chunks; ++kc) {
                for (int dxi = 0; dxi < 3; ++dxi) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t a_lo = a_lo0 + stage * (kStageBytes >> 4);
                        const uint32_t b_lo = W_STAT ? w_lo0 + (dxi * p.c_chunks + kc) * (kBBytes >> 4) : a_lo + (kABytes >> 4);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if constexpr (PAIR)
                                    umma_f16_pair(d_tmem, desc_hi | (a_lo + dy * 64 + 2 * k), desc_hi | (b_lo + dy * b_step + 2 * k), idesc,
                                                  (kc | dxi | dy | k) != 0 ? 1u : 0u);
                                else
                                    umma_f16(d_tmem, desc_hi | (a_lo + dy * 64 + 2 * k), desc_hi | (b_lo + dy * b_step + 2 * k), idesc,
                                             (kc | dxi | dy | k) != 0 ? 1u : 0u);
                            }
                        }
                        if constexpr (PAIR) {
                            umma_commit_pair(bar_empty + 8 * stage);
                            if (kc == p.c_chunks - 1 && dxi == 2) umma_commit_pair(bar_tfull + 8 * acc);
                        } else {
                            umma_commit(bar_empty + 8 * stage);
                            if (kc == p.c_chunks - 1 && dxi == 2) umma_commit(bar_tfull + 8 * acc);
                        }
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        const int quarter = warp & 3;
        const int g = EPG == 2 ? (warp - 4) >> 2 : 0;          // epilogue group: columns [64 g, 64 g + 64) when there are two
        const int etid = (threadIdx.x - 128) & 127;
        const int row = quarter * 32 + lane;
        const int rh = row >> 3, rw = row & 7;
        float* s_bias = reinterpret_cast<float*>(s_out + out_bytes + 256);
        int store_count = 0;
        float* s_outw = s_bias + 256;
        int iter = 0, cur_nt = -1;
        const uint32_t tempty_sig = PAIR ? mapa_cluster(bar_tempty, 0) : bar_tempty;
        for (int t = sched0; t < total_tiles; t += sched_step, ++iter) {
            const int acc = iter % kAcc;
            const uint32_t acc_phase = (iter / kAcc) & 1;
            const int n_tile = t % p.n_tiles;
            const int m_tile = PAIR ? 2 * (t / p.n_tiles) + static_cast<int>(rank) : t / p.n_tiles;
            if (n_tile != cur_nt) {
                if constexpr (EPG == 2) {   // both groups share the per-column constants
                    named_bar_sync(3, 256);
                    for (int i = threadIdx.x - 128; i < BLOCK_N; i += 256) s_bias[i] = __ldg(p.bias + n_tile * BLOCK_N + i);
                    named_bar_sync(3, 256);
                } else {
                    epilogue_consts<BLOCK_N, EPI>(p, n_tile, s_bias, s_outw, threadIdx.x - 128);
                }
                cur_nt = n_tile;
            }
            const int w = (m_tile % p.tiles_w) * 8 + rw;
            const int h = ((m_tile / p.tiles_w) % p.tiles_h) * 16 + rh;
            const int n = m_tile / (p.tiles_w * p.tiles_h);
            uint32_t res0[32];
            if (EPI == EPI_STORE && p.res != nullptr && n < p.N)
                res_load64(p.res + ((static_cast<size_t>(n) * p.H + h) * p.W + w) * p.res_c_stride + n_tile * BLOCK_N + 64 * g, res0);
            mbar_wait(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BLOCK_N + 64 * g;
            if constexpr (EPG == 2)
                epilogue_tile<BLOCK_N, EPI, 64>(p, taddr, row, n, h - rh, w - rw, n, h, w, n < p.N, n_tile, s_bias, s_outw, s_out + g * kOutBufBytes,
                                                stages_addr + S * kStageBytes + g * kOutBufBytes, store_count, etid, res0, 64 * g, 1 + g);
            else
                epilogue_tile<BLOCK_N, EPI>(p, taddr, row, n, h - rh, w - rw, n, h, w, n < p.N, n_tile, s_bias, s_outw, s_out,
                                            stages_addr + S * kStageBytes, store_count, etid, res0);
            tc_fence_before();
            if constexpr (PAIR) {
                named_bar_sync(1 + g, 128);
                if (etid == 0) mbar_arrive_cluster(tempty_sig + 8 * acc);
            } else {
                mbar_arrive(bar_tempty + 8 * acc);
            }
        }
        if (etid == 0) bulk_wait_all();
    }

    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();
    else __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1, Cout = 64 "row streaming" variant for the two full-resolution UNet levels (inc.3, up4.conv0,
// up4.conv3 + head).  With BLOCK_N = 64 one tcgen05.mma 128x64x16 reads 4 KB of A + 2 KB of B from shared memory for 32
// tensor-clocks of work, so the kernels above are bound by shared-memory operand bandwidth at ~45 % tensor utilisation
// (profiles/README.md).  Here the three VERTICAL taps of one horizontal offset become one MMA with N = 192:
//     D[128 pixels of input row y][dy*64 + co] = A(row y, shifted by dx) . [W(dy=0,dx) | W(dy=1,dx) | W(dy=2,dx)]
// so A is read once for three taps.  Column block dy is the contribution of input row y to OUTPUT row y + 1 - dy; the
// TMEM accumulator is a ring of eight 64-column slots, slot(out row r) = (-r) & 7, so the three blocks of an input row
// land on three consecutive slots and consecutive input rows slide the window down by one slot.  An output row is
// complete after input row r + 1 has been issued; its slot is then drained by the epilogue warps while the MMA warp
// carries on with the other slots.  Where the window wraps around the ring (2 of 8 rows) and on the very first K step
// of a slot (accumulate = 0 for block dy = 0 only) the MMA is split into N = 64 / N = 128 pieces.
// One CTA streams a strip of rs_rows output rows x 128 columns of one image; weights (9*Cin*64*2 B) stay resident.
// The MMA-issuing thread is the critical resource of this kernel (12 MMAs per row, each only 96 tensor-clocks long): the
// TMEM runs and descriptor words of a row are computed once, the K loop is fully unrolled, and one elected thread runs
// the role loop runs warp-uniformly with only the tcgen05 instructions predicated on the elected lane.  Two epilogue
// groups (warps 4-7 / 8-11) drain alternate output ROW PAIRS and write fp16 NHWC with 256-bit global stores (one full
// 32-byte sector per lane and instruction), so the epilogue adds no shared-memory traffic; a thread therefore sees both
// rows of a 2x2 window and the encoder's max-pool is fused here.  The same kernel runs the 16x16 ResNet level: a streamed
// row is then the same image row of eight squares (box {64 ch, 16 px, 1 row, 8 images}, MODE 0).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kRsBarBytes = 512;
constexpr int kRsThreads = 384;   // TMA warp, MMA warp, TMEM warp, one idle warp, two epilogue groups of four warps

// Contiguous runs of TMEM slots covered by column blocks dy in [lo, hi] of one input row (at most two: the ring wraps
// after slot 7).  col = first TMEM column, boff = offset of the first block inside the resident N = 192 weight operand
// (16-byte units), idesc = instruction descriptor with N = 64 * run length.
struct RsRuns {
    uint32_t col[2], boff[2], idesc[2];
    int n;
};
__device__ __forceinline__ void rs_runs(uint32_t s_base, int lo, int hi, uint32_t idesc0, RsRuns& r) {
    r.n = 0;
    r.col[1] = r.boff[1] = r.idesc[1] = 0;
    if (lo > hi) return;
    const uint32_t s0 = (s_base + static_cast<uint32_t>(lo)) & 7u;
    int n0 = hi - lo + 1;
    const int room = 8 - static_cast<int>(s0);
    const int n1 = n0 > room ? n0 - room : 0;
    n0 -= n1;
    r.col[0] = s0 * 64u;
    r.boff[0] = static_cast<uint32_t>(lo) * (64u * 128u >> 4);
    r.idesc[0] = idesc0 | (static_cast<uint32_t>(n0 * 8) << 17);
    r.n = 1;
    if (n1) {
        r.col[1] = 0;
        r.boff[1] = static_cast<uint32_t>(lo + n0) * (64u * 128u >> 4);
        r.idesc[1] = idesc0 | (static_cast<uint32_t>(n1 * 8) << 17);
        r.n = 2;
    }
}


// All K steps of one pipeline stage for an interior input row (its three column blocks land on slots s_base .. s_base+2
// of the ring).  NA = slots before the ring wraps (3: no wrap; 2 or 1: the remaining 3 - NA blocks go to slot 0 on).
// Straight-line code: every N and every operand offset is a compile-time constant.
template <int NA, int ADX>
__device__ __forceinline__ void rs_issue_row(uint32_t tmem_base, uint32_t s_base, uint64_t desc_hi, uint32_t a_lo, uint32_t b_lo,
                                             uint32_t b_dx_stride, uint32_t idesc0, bool first) {
    constexpr int NB = 3 - NA;
    constexpr uint32_t kBlk = 64u * 128u >> 4;   // one 64-row weight block, 16-byte units
    const uint32_t d_a = tmem_base + s_base * 64u;
    const uint32_t i_a = idesc0 | (static_cast<uint32_t>(NA * 8) << 17);
    const uint32_t i_b = idesc0 | (static_cast<uint32_t>(NB * 8) << 17);
#pragma unroll
    for (int dd = 0; dd < 3; ++dd) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t a_desc = desc_hi | (a_lo + static_cast<uint32_t>(ADX) * dd + 2u * k);
            const uint32_t b_k = b_lo + dd * b_dx_stride + 2u * k;
            if (dd == 0 && k == 0 && first) {
                umma_f16(d_a, a_desc, desc_hi | b_k, idesc0 | (8u << 17), 0u);
                if constexpr (NA > 1) umma_f16(d_a + 64u, a_desc, desc_hi | (b_k + kBlk), idesc0 | (static_cast<uint32_t>((NA - 1) * 8) << 17), 1u);
            } else {
                umma_f16(d_a, a_desc, desc_hi | b_k, i_a, 1u);
            }
            if constexpr (NB > 0) umma_f16(tmem_base, a_desc, desc_hi | (b_k + NA * kBlk), i_b, 1u);
        }
    }
}

// Same for the first / last input rows of a strip, whose window is partial (kept out of line so that its bookkeeping is
// not hoisted into the hot loop).  fresh_blocks: the first K step overwrites blocks dy < fresh_blocks (new output rows).
template <int ADX>
__device__ __noinline__ void rs_issue_partial(uint32_t tmem_base, uint32_t s_base, int dy_lo, int dy_hi, uint64_t desc_hi, uint32_t a_lo,
                                              uint32_t b_lo, uint32_t b_dx_stride, uint32_t idesc0, int fresh_blocks) {
    RsRuns all;
    rs_runs(s_base, dy_lo, dy_hi, idesc0, all);
    for (int dd = 0; dd < 3; ++dd) {
        for (int k = 0; k < 4; ++k) {
            const uint64_t a_desc = desc_hi | (a_lo + static_cast<uint32_t>(ADX) * dd + 2u * k);
            const uint32_t b_k = b_lo + dd * b_dx_stride + 2u * k;
            if (fresh_blocks > 0) {   // first K step of the row: block by block, overwriting the new output rows
                for (int dy = dy_lo; dy <= dy_hi; ++dy)
                    umma_f16(tmem_base + ((s_base + dy) & 7u) * 64u, a_desc, desc_hi | (b_k + dy * (64u * 128u >> 4)), idesc0 | (8u << 17),
                             dy < fresh_blocks ? 0u : 1u);
                fresh_blocks = 0;
            } else {
                umma_f16(tmem_base + all.col[0], a_desc, desc_hi | (b_k + all.boff[0]), all.idesc[0], 1u);
                if (all.n > 1) umma_f16(tmem_base + all.col[1], a_desc, desc_hi | (b_k + all.boff[1]), all.idesc[1], 1u);
            }
        }
    }
}

// A pipeline stage holds everything one input row (and one 64-channel chunk) contributes: all three horizontal taps.
// MODE 0: three TMA boxes {64 ch, 128 px}, one per dx, 16 KB apart (also the 8 x 16 form: pixels of different images are
// not contiguous).  MODE 1: one box {64 ch, 130 px}, tap dx addressed 128*dx bytes further (the 128-byte swizzle is a
// function of the shared-memory address, so a start address that is a multiple of 128 B inside a 1024-byte group keeps
// the pattern the TMA wrote).
// A strip that covers the whole image height (the 16x16 squares) skips input rows -1 and rs_rows: they are all zero.
// POOL: the 2x2 max-pool of the output is written as well (p.pool_out); a template parameter so that the row-pair
// registers of the pooling form and the residual registers of the other do not share one register budget.
template <int EPI, int MODE, bool POOL>
__global__ void __launch_bounds__(kRsThreads, 1) conv3x3_rs_kernel(const __grid_constant__ ConvParams p) {
    constexpr uint32_t kBBytes = 64 * 128;
    constexpr uint32_t a_bytes = MODE == 0 ? 3u * 128u * 128u : 130u * 128u;   // bytes the TMA delivers per stage
    constexpr uint32_t stage_bytes = MODE == 0 ? 49152u : 17408u;              // 1024-aligned stage pitch
    constexpr int ADX = MODE == 0 ? 1024 : 8;                                  // A descriptor step per horizontal tap (16-byte units)
    const int S = p.vr_stages;
    const int R = p.rs_rows;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = smem_u32(smem_raw);
    const uint32_t base_addr = (raw_addr + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base_addr - raw_addr);
    const uint32_t w_bytes = 9u * p.c_chunks * kBBytes;
    const uint32_t stages_addr = base_addr + w_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + w_bytes + S * stage_bytes);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * S;
    const uint32_t bar_tfull = bar_full + 16 * S;     // 8 slots
    const uint32_t bar_tempty = bar_tfull + 64;       // 8 slots
    const uint32_t bar_w = bar_tempty + 64;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 17);
    float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + kRsBarBytes);
    float* s_outw = s_bias + 256;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.b_map);
        tma_prefetch_desc(&p.a_map[MODE == 0 ? 0 : 1]);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        for (int i = 0; i < 8; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);   // one arrival per epilogue warp of the group that drained the slot
        }
        mbar_init(bar_w, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc(smem_u32(tmem_slot), 512);
    if (warp == 3) {   // per-column constants (Cout = 64: one N tile)
        for (int i = lane; i < 64; i += 32) {
            s_bias[i] = __ldg(p.bias + i);
            if constexpr (EPI == EPI_OUTC) s_outw[i] = __ldg(p.outc_w + i);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_launch();
    if (warp != 0) griddep_wait();   // warp 0 first requests the resident weights (no kernel writes them)

    const int strips_per_image = p.tiles_h * p.tiles_w;
    const int total_strips = p.tiles_n * strips_per_image;
    const int loads_per_row = p.c_chunks;
    const bool skip_halo = R == p.H;   // the strip is the whole image: input rows -1 and R are outside it
    const int j_lo = skip_halo ? 0 : -1, j_hi = skip_halo ? R - 1 : R;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(bar_w, w_bytes);
            // resident order: ((dx * c_chunks + kc) * 3 + dy), so the three vertical taps of one (dx, kc) are one N = 192 operand
            for (int dxi = 0; dxi < 3; ++dxi)
                for (int kc = 0; kc < p.c_chunks; ++kc)
                    for (int dy = 0; dy < 3; ++dy)
                        tma_load_2d(base_addr + ((dxi * p.c_chunks + kc) * 3 + dy) * kBBytes, &p.b_map, bar_w,
                                    ((dy * 3 + dxi) * p.c_chunks + kc) * 64, 0);
            griddep_wait();
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < total_strips; t += gridDim.x) {
                const int w0 = (t % p.tiles_w) * p.tw;
                const int h0 = ((t / p.tiles_w) % p.tiles_h) * R;
                const int n0 = (t / strips_per_image) * p.tn;
                for (int j = j_lo; j <= j_hi; ++j) {
                    for (int kc = 0; kc < loads_per_row; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        mbar_expect_tx(bar_full + 8 * stage, a_bytes);
                        if (MODE == 0) {
                            for (int dxi = 0; dxi < 3; ++dxi)
                                tma_load_4d(stages_addr + stage * stage_bytes + dxi * 16384u, &p.a_map[0], bar_full + 8 * stage,
                                            p.a_c_off + kc * 64, w0 + dxi - 1, h0 + j, n0);
                        } else {
                            tma_load_4d(stages_addr + stage * stage_bytes, &p.a_map[1], bar_full + 8 * stage, p.a_c_off + kc * 64, w0 - 1, h0 + j, n0);
                        }
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // The role loop runs warp-uniformly (so descriptors and barrier addresses live in uniform registers) and only
        // the tcgen05 instructions are predicated on the elected lane.  The issuing thread is the critical resource:
        // the tensor pipe drains 12 queued MMAs in ~1150 clocks, so a row must not cost more than that to set up.
        const bool leader = elect_one();
        mbar_wait(bar_w, 0);
        tc_fence_after();
        const uint32_t idesc0 = p.idesc & ~(0x3Fu << 17);
        const uint64_t desc_hi = umma_desc_sw128(0) & 0xFFFFFFFF00000000ull;
        const uint32_t desc_lo0 = static_cast<uint32_t>(umma_desc_sw128(0));
        const uint32_t w_lo = desc_lo0 + ((base_addr & 0x3FFFFu) >> 4);
        const uint32_t a_lo0 = desc_lo0 + ((stages_addr & 0x3FFFFu) >> 4);
        const uint32_t b_dx_stride = static_cast<uint32_t>(p.c_chunks) * 3u * (kBBytes >> 4);   // one horizontal tap further
        int stage = 0;
        uint32_t phase = 0;
        uint32_t rho0 = 0;   // output rows this CTA has started before the current strip
        for (int t = blockIdx.x; t < total_strips; t += gridDim.x, rho0 += R) {
            for (int j = j_lo; j <= j_hi; ++j) {
                // input row j feeds output rows j + 1 - dy, dy in [dy_lo, dy_hi]
                const int dy_lo = j + 2 - R > 0 ? j + 2 - R : 0;
                const int dy_hi = j + 1 < 2 ? j + 1 : 2;
                const uint32_t rho_new = rho0 + static_cast<uint32_t>(j + 1);
                const uint32_t s_base = (0u - rho_new) & 7u;   // slot of block dy = (s_base + dy) & 7
                // output rows this input row starts (their first K step overwrites): block 0 always, and block 1 too when
                // the zero row above the image is skipped and this is input row 0
                const int fresh_blocks = dy_lo == 0 ? (skip_halo && j == 0 ? 2 : 1) : 0;
                const bool interior = dy_lo == 0 && dy_hi == 2;
                for (int f = 0; f < fresh_blocks; ++f) {   // a new output row's slot must have been drained
                    const uint32_t rho = rho_new - static_cast<uint32_t>(f);
                    mbar_wait(bar_tempty + 8 * ((s_base + f) & 7u), ((rho >> 3) & 1u) ^ 1u);
                }
                tc_fence_after();
                for (int kc = 0; kc < loads_per_row; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint32_t a_lo = a_lo0 + stage * (stage_bytes >> 4);
                    const uint32_t b_lo = w_lo + static_cast<uint32_t>(kc * 3) * (kBBytes >> 4);
                    const bool first = kc == 0 && fresh_blocks > 0;   // overwrite on the first K step of a new output row
                    if (leader) {
                        if (interior) {
                            if (s_base <= 5u) rs_issue_row<3, ADX>(tmem_base, s_base, desc_hi, a_lo, b_lo, b_dx_stride, idesc0, first);
                            else if (s_base == 6u) rs_issue_row<2, ADX>(tmem_base, s_base, desc_hi, a_lo, b_lo, b_dx_stride, idesc0, first);
                            else rs_issue_row<1, ADX>(tmem_base, s_base, desc_hi, a_lo, b_lo, b_dx_stride, idesc0, first);
                        } else {   // first / last input rows of a strip: a partial window
                            rs_issue_partial<ADX>(tmem_base, s_base, dy_lo, dy_hi, desc_hi, a_lo, b_lo, b_dx_stride, idesc0, first ? fresh_blocks : 0);
                        }
                        umma_commit(bar_empty + 8 * stage);
                        if (kc == loads_per_row - 1) {
                            // after the last K step of input row j, output row j - 1 is complete; so is row R - 1 when
                            // there is no input row R to wait for
                            if (j >= 1) umma_commit(bar_tfull + 8 * ((s_base + 2u) & 7u));
                            if (skip_halo && j == R - 1) umma_commit(bar_tfull + 8 * ((s_base + 1u) & 7u));
                        }
                    }
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // two epilogue groups: group g drains the output row PAIRS with ((rho >> 1) & 1) == g, so two rows are in flight
        // and a thread sees both rows of a 2x2 max-pool window (the fused MaxPool2d of unet_parts.py:34)
        const int group = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        uint32_t rho0 = 0;
        for (int t = blockIdx.x; t < total_strips; t += gridDim.x, rho0 += R) {
            // a streamed row is tn images x tw pixels (1 x 128 for the wide UNet levels, 8 x 16 for 16x16 squares)
            const int w0 = (t % p.tiles_w) * p.tw + row % p.tw;
            const int h0 = ((t / p.tiles_w) % p.tiles_h) * R;
            const int n = (t / strips_per_image) * p.tn + row / p.tw;
            const bool has_res = EPI == EPI_STORE && !POOL && p.res != nullptr && n < p.N;
            for (int r2 = 2 * group; r2 < R; r2 += 4) {
                uint32_t o_prev[POOL ? 32 : 1];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r = r2 + half;
                    const uint32_t rho = rho0 + static_cast<uint32_t>(r);
                    const uint32_t slot = (0u - rho) & 7u;
                    const size_t pix = (static_cast<size_t>(n) * p.H + h0 + r) * p.W + w0;
                    // requested before the accumulator is waited for (a one-row look-ahead costs 32 more live registers and
                    // spills: measured slower)
                    uint32_t res_cur[32];
                    if (has_res) res_load64(p.res + pix * p.res_c_stride, res_cur);
                    mbar_wait(bar_tfull + 8 * slot, (rho >> 3) & 1u);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + slot * 64u;
                    if constexpr (EPI == EPI_OUTC) {
                        int unused = 0;
                        epilogue_tile<64, EPI>(p, taddr, row, n, h0 + r, w0, n, h0 + r, w0, n < p.N, 0, s_bias, s_outw, nullptr, 0u, unused, 0, res_cur);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);
                    } else {
                        uint32_t va[32], vb[32], o[32];
                        tmem_ld_32x32(taddr, va);
                        tmem_ld_32x32(taddr + 32, vb);
                        tmem_ld_wait(va);
                        pack_chunk<0>(va, s_bias, res_cur, has_res, p.relu != 0, reinterpret_cast<uint32_t(&)[16]>(o[0]));
                        tmem_ld_wait(vb);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * slot);   // the accumulator is in registers: release the slot early
                        pack_chunk<16>(vb, s_bias + 32, res_cur, has_res, p.relu != 0, reinterpret_cast<uint32_t(&)[16]>(o[16]));
                        if (n < p.N) {
                            __half* dst = p.out + pix * p.out_c_stride + p.out_c_off;
#pragma unroll
                            for (int q = 0; q < 4; ++q) st_global_v8(dst + 16 * q, o + 8 * q);
                        }
                        if constexpr (POOL) {
                            if (half == 0) {
#pragma unroll
                                for (int i = 0; i < 32; ++i) o_prev[i] = o[i];
                            } else {
                                // vertical max with the row above, horizontal max with the neighbouring pixel (lane ^ 1)
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&o[i]), *reinterpret_cast<const __half2*>(&o_prev[i]));
                                    uint32_t mu = *reinterpret_cast<const uint32_t*>(&m);
                                    const uint32_t nb = __shfl_xor_sync(0xffffffffu, mu, 1);
                                    const __half2 mm = __hmax2(m, *reinterpret_cast<const __half2*>(&nb));
                                    o[i] = *reinterpret_cast<const uint32_t*>(&mm);
                                }
                                if (n < p.N && (lane & 1) == 0) {
                                    const size_t ppix = (static_cast<size_t>(n) * (p.H >> 1) + ((h0 + r) >> 1)) * (p.W >> 1) + (w0 >> 1);
                                    __half* pdst = p.pool_out + ppix * p.pool_c_stride;
#pragma unroll
                                    for (int q = 0; q < 4; ++q) st_global_v8(pdst + 16 * q, o + 8 * q);
                                }
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------------------ host

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

int tmap_init() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || fn == nullptr) return -1;
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
    return 0;
}

int tmap_act(CUtensorMap* m, const void* base, int C, int Wv, int Hv, int Nv, int64_t sW, int64_t sH, int64_t sN, int tw,
             int th, int tn) {
    if (tmap_init()) return -1;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wv, (cuuint64_t)Hv, (cuuint64_t)Nv};
    cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)tn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

static int g_pair_clusters = 0;   // see configure_pair

int tmap_weights(CUtensorMap* m, const void* base, int K_total, int rows, int block_n) {
    if (tmap_init()) return -1;
    cuuint64_t dims[2] = {(cuuint64_t)K_total, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K_total * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}

int tmap_act_vr(CUtensorMap* m, const void* base, int C, int Wv, int Hv, int Nv, int64_t sW, int64_t sH, int64_t sN) {
    return tmap_act(m, base, C, Wv, Hv, Nv, sW, sH, sN, 8, 18, 1);
}

static void pick_tile(int Ho, int Wo, int& tn, int& th, int& tw) {
    tw = Wo < 16 ? Wo : 16;
    th = 128 / tw;
    if (th > Ho) th = Ho;
    tn = 128 / (tw * th);
}

// Describe one conv (ksize 1|3, stride 1|2, pad ksize/2) over an NHWC fp16 buffer whose pixel stride is in_c_stride.
// Returns 0, -5 (shape not supported) or a cuTensorMapEncodeTiled error (< -1000).
int conv_build(ConvLaunch& L, const __half* in, int Nmax, int Hin, int Win, int in_c_stride, int in_c_off, int Cin, const __half* w,
               const float* bias, int rows, int K, int ksize, int stride, int epilogue, bool use_vr) {
    memset(&L, 0, sizeof L);
    ConvParams& p = L.p;
    const int Ho = Hin / stride, Wo = Win / stride;
    if (Cin % 64 || K != ksize * ksize * Cin) return -5;
    pick_tile(Ho, Wo, p.tn, p.th, p.tw);
    if (Wo % p.tw || Ho % p.th || p.tn * p.th * p.tw != 128) return -5;
    p.H = Ho;
    p.W = Wo;
    p.tiles_w = Wo / p.tw;
    p.tiles_h = Ho / p.th;
    p.taps = ksize * ksize;
    p.c_chunks = Cin / 64;
    p.a_c_off = in_c_off;
    const int64_t sW = in_c_stride, sH = static_cast<int64_t>(Win) * in_c_stride, sN = static_cast<int64_t>(Hin) * Win * in_c_stride;
    int rc = 0;
    if (stride == 1) {
        rc = tmap_act(&p.a_map[0], in, in_c_stride, Win, Hin, Nmax, sW, sH, sN, p.tw, p.th, p.tn);
        for (int r = 0; r < ksize; ++r)
            for (int s = 0; s < ksize; ++s) {
                p.tap_map[r * ksize + s] = 0;
                p.tap_dy[r * ksize + s] = static_cast<int8_t>(r - ksize / 2);
                p.tap_dx[r * ksize + s] = static_cast<int8_t>(s - ksize / 2);
            }
    } else {
        // stride 2: four parity views (py,px) of the input, each with doubled strides; tap r reads parity (r+1)&1 at
        // view offset -1 (r == 0) or 0, and the zero fill of the view at -1 is exactly the padding row/column.
        for (int py = 0; py < 2 && !rc; ++py)
            for (int px = 0; px < 2 && !rc; ++px)
                rc = tmap_act(&p.a_map[py * 2 + px], in + (static_cast<int64_t>(py) * Win + px) * in_c_stride, in_c_stride,
                              Win / 2, Hin / 2, Nmax, 2 * sW, 2 * sH, sN, p.tw, p.th, p.tn);
        for (int r = 0; r < ksize; ++r)
            for (int s = 0; s < ksize; ++s) {
                const int rr = ksize == 1 ? 1 : r, ss = ksize == 1 ? 1 : s;  // 1x1: the centre tap
                const int py = (rr + 1) & 1, px = (ss + 1) & 1;
                p.tap_map[r * ksize + s] = static_cast<int8_t>(py * 2 + px);
                p.tap_dy[r * ksize + s] = static_cast<int8_t>(rr == 0 ? -1 : 0);
                p.tap_dx[r * ksize + s] = static_cast<int8_t>(ss == 0 ? -1 : 0);
            }
    }
    if (rc) return rc;
    L.block_n = rows % 256 == 0 ? 256 : (rows % 128 == 0 ? 128 : 64);
    if (epilogue == EPI_OUTC) L.block_n = 64;
    if (rows % L.block_n) return -5;
    p.n_tiles = rows / L.block_n;
    rc = tmap_weights(&p.b_map, w, K, rows, L.block_n);
    if (rc) return rc;
    p.bias = bias;
    L.epilogue = epilogue;
    L.n_max = Nmax;
    L.pdl = 1;
    if (epilogue == EPI_FUSED_CONVT) {
        if (L.block_n != 128 || p.n_tiles != 1 || ksize != 3 || stride != 1 || Ho % 16 || Wo % 8) return -5;
        p.tn = 1; p.th = 16; p.tw = 8;   // the vertical-reuse tile: 16 rows x 8 columns, activation box {64, 8, 18, 1}
        p.tiles_w = Wo / 8;
        p.tiles_h = Ho / 16;
        const char* np = getenv("CVB_NO_PAIR_CONVT");
        L.pair = g_pair_clusters > 0 && !(np && np[0] == '1') ? 1 : 0;   // conv_convt_pair_kernel: 64 weight rows per CTA
        if (L.pair && (rc = tmap_weights(&p.b_map, w, K, rows, 64))) return rc;
        return tmap_act_vr(&p.a_map[0], in, in_c_stride, Win, Hin, Nmax, sW, sH, sN);   // conv_set_fused_convt completes the launch
    }
    if (use_vr && conv_try_rs(L, ksize, stride, Ho, Wo, Cin)) {
        rc = tmap_act(&p.a_map[0], in, in_c_stride, Win, Hin, Nmax, sW, sH, sN, p.tw, 1, p.tn);
        if (!rc && p.rs_mode == 1) rc = tmap_act(&p.a_map[1], in, in_c_stride, Win, Hin, Nmax, sW, sH, sN, 130, 1, 1);
        if (rc) return rc;
    } else if (use_vr && conv_try_vr(L, ksize, stride, Ho, Wo, Cin)) {
        rc = tmap_act_vr(&p.a_map[0], in, in_c_stride, Win, Hin, Nmax, sW, sH, sN);
        if (!rc && L.pair) rc = tmap_weights(&p.b_map, w, K, rows, L.block_n / 2);
        if (rc) return rc;
    }
    return conv_try_pair(L, w, K, rows);
}

// Generic kernel, BLOCK_N = 128 / 256, plain or transposed-conv store: run it as CTA pairs (each CTA then loads half of the
// weight tile: the weight view gets a box of BLOCK_N / 2 rows).
int conv_try_pair(ConvLaunch& L, const __half* w, int K, int rows) {
    // generic kernel, N >= 128, at most four K steps (up3.up, the 1x1 downsamples): two epilogue groups, each with one of the
    // two staging buffers.  Measured per 148 boards: up3.up 205 -> 190 us, layer2 downsample 52 -> 45 us; with 8 or 9 K steps
    // (up2.up, layer2.0.conv1) it is 2-3 % slower, so the bound is 4.  CVB_EPG2=0 turns it off, CVB_EPG2_MAX_K sets the bound.
    if (L.variant == 0 && (L.block_n == 128 || L.block_n == 256) && (L.epilogue == EPI_STORE || L.epilogue == EPI_CONVT)) {
        const char* off = getenv("CVB_EPG2");
        const char* mk = getenv("CVB_EPG2_MAX_K");
        if (!(off && off[0] == '0') && L.p.taps * L.p.c_chunks <= (mk ? atoi(mk) : 4)) L.epg = 2;
    }
    if (g_pair_clusters <= 0 || L.variant != 0 || (L.block_n != 128 && L.block_n != 256)) return 0;
    if (L.epilogue != EPI_STORE && L.epilogue != EPI_CONVT) return 0;
    // Measured (profiles/README.md, round 2): the pair form wins 7-14 % on N = 256 tiles with long K loops and loses on short
    // ones (1x1 downsamples, transposed convs with Cin <= 512: the epilogue is the bottleneck there and two CTAs in lock step
    // wait for the slower of the two) and on N = 128 tiles.  CVB_PAIR_MIN_N / CVB_PAIR_MIN_K override the thresholds (A/B runs).
    const char* min_n = getenv("CVB_PAIR_MIN_N");
    const char* min_k = getenv("CVB_PAIR_MIN_K");
    if (L.block_n < (min_n ? atoi(min_n) : 256)) return 0;
    if (L.p.taps * L.p.c_chunks < (min_k ? atoi(min_k) : 16)) return 0;
    L.pair = 1;
    return tmap_weights(&L.p.b_map, w, K, rows, L.block_n / 2);
}

// "Convolution" with kernel 2x2, stride 2, no padding over an NHWC buffer [Nmax, 2*Ho, 2*Wo, in_c_stride]: tap q = (dy,dx)
// reads the parity view q at the output pixel itself.  This is the data gradient of ConvTranspose2d(k=2, s=2)
// (unet_parts.py:53): dx[p][ci] = sum_{q,co} dy[2p+q][co] * W[ci][co][q], weights packed [rows = Cin_t][K = 4*C] with
// k = q*C + co.  C = channels read per tap (from in_c_off).
int conv_build_k2s2(ConvLaunch& L, const __half* in, int Nmax, int Ho, int Wo, int in_c_stride, int in_c_off, int C, const __half* w,
                    const float* bias, int rows, int K) {
    memset(&L, 0, sizeof L);
    ConvParams& p = L.p;
    if (C % 64 || K != 4 * C) return -5;
    pick_tile(Ho, Wo, p.tn, p.th, p.tw);
    if (Wo % p.tw || Ho % p.th || p.tn * p.th * p.tw != 128) return -5;
    p.H = Ho;
    p.W = Wo;
    p.tiles_w = Wo / p.tw;
    p.tiles_h = Ho / p.th;
    p.taps = 4;
    p.c_chunks = C / 64;
    p.a_c_off = in_c_off;
    const int Win = 2 * Wo, Hin = 2 * Ho;
    const int64_t sW = in_c_stride, sH = static_cast<int64_t>(Win) * in_c_stride, sN = static_cast<int64_t>(Hin) * Win * in_c_stride;
    int rc = 0;
    for (int q = 0; q < 4 && !rc; ++q) {
        rc = tmap_act(&p.a_map[q], in + (static_cast<int64_t>(q >> 1) * Win + (q & 1)) * in_c_stride, in_c_stride, Wo, Ho, Nmax, 2 * sW,
                      2 * sH, sN, p.tw, p.th, p.tn);
        p.tap_map[q] = static_cast<int8_t>(q);
        p.tap_dy[q] = p.tap_dx[q] = 0;
    }
    if (rc) return rc;
    L.block_n = rows % 256 == 0 ? 256 : (rows % 128 == 0 ? 128 : 64);
    if (rows % L.block_n) return -5;
    p.n_tiles = rows / L.block_n;
    rc = tmap_weights(&p.b_map, w, K, rows, L.block_n);
    if (rc) return rc;
    p.bias = bias;
    L.epilogue = EPI_STORE;
    L.n_max = Nmax;
    L.pdl = 1;
    return conv_try_pair(L, w, K, rows);
}

// Output side of a launch: NHWC fp16 buffer with `out_c_stride` channels per pixel, first output channel `out_c_off`.
// Builds the TMA store views (box {64 ch, tw, th, tn}); the transposed convolution gets one stride-2 view per (dy,dx).
int conv_set_store(ConvLaunch& L, __half* out, int out_c_stride, int out_c_off, int relu, const __half* res, int res_c_stride) {
    ConvParams& p = L.p;
    p.out = out;
    p.out_c_stride = out_c_stride;
    p.out_c_off = out_c_off;
    p.relu = relu;
    p.res = res;
    p.res_c_stride = res_c_stride;
    if (L.epilogue == EPI_OUTC) {
        p.out_bufs = 0;
        return 0;
    }
    if (L.variant == 0) p.out_bufs = L.epg == 2 ? 1 : 2;   // two groups: one of the two staging buffers each
    int rc = 0;
    const int64_t C = out_c_stride;
    if (L.epilogue == EPI_CONVT) {
        const int64_t W2 = 2 * p.W, H2 = 2 * p.H;
        for (int q = 0; q < 4 && !rc; ++q)
            rc = tmap_act(&p.o_map[q], out + ((q >> 1) * W2 + (q & 1)) * C, out_c_stride, p.W, p.H, L.n_max, 2 * C, 2 * W2 * C, H2 * W2 * C,
                          p.tw, p.th, p.tn);
    } else {
        rc = tmap_act(&p.o_map[0], out, out_c_stride, p.W, p.H, L.n_max, C, p.W * C, static_cast<int64_t>(p.H) * p.W * C, p.tw, p.th, p.tn);
    }
    return rc;
}

int conv_set_fused_convt(ConvLaunch& L, const __half* w2, const float* bias2, int cout2, __half* out, int out_c_stride, int out_c_off) {
    ConvParams& p = L.p;
    if (L.epilogue != EPI_FUSED_CONVT || L.block_n != 128 || p.n_tiles != 1 || cout2 != 64) return -5;
    if (tmap_init()) return -1;
    // transposed-conv weights [4 * cout2 rows][128 k]: box {64 k, 128 rows} (one tile of the weight ring)
    int rc = tmap_weights(&p.b2_map, w2, 128, 4 * cout2, L.pair ? 64 : 128);
    if (rc) return rc;
    p.bias2 = bias2;
    p.convt_cout = cout2;
    p.out = out;
    p.out_c_stride = out_c_stride;
    p.out_c_off = out_c_off;
    p.relu = 1;
    p.out_bufs = 2;
    const int64_t C = out_c_stride, W2 = 2 * p.W, H2 = 2 * p.H;
    for (int q = 0; q < 4 && !rc; ++q)
        rc = tmap_act(&p.o_map[q], out + ((q >> 1) * W2 + (q & 1)) * C, out_c_stride, p.W, p.H, L.n_max, 2 * C, 2 * W2 * C, H2 * W2 * C, p.tw, p.th, p.tn);
    return rc;
}

constexpr int kVrMaxSmem = 227 * 1024;

template <int BN, int EPI, bool WS>
static cudaError_t configure_vr() {
    return cudaFuncSetAttribute(conv3x3_vr_kernel<BN, EPI, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem);
}

template <int BN, int EPI>
static cudaError_t configure_one() {
    return cudaFuncSetAttribute(conv_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<BN>::kSmemBytes);
}

// CTA-pair form of the generic kernel: clusters of two CTAs.  g_pair_clusters = how many such clusters the device runs at once
// (74 on a B200: one per TPC); 0 turns the pair form off (CVB_NO_PAIR=1, or a device that cannot co-schedule them).
template <int BN, int EPI>
static cudaError_t configure_epg2() {
    return cudaFuncSetAttribute(conv_tc_kernel<BN, EPI, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<BN>::kSmemBytes);
}

template <int BN, int EPI>
static cudaError_t configure_pair() {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, EPI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvCfg<BN, 1>::kSmemBytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = ConvCfg<BN, 1>::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<BN, EPI, 1>, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (n < g_pair_clusters) g_pair_clusters = n;
    return cudaSuccess;
}

int conv_pair_clusters() { return g_pair_clusters; }

cudaError_t conv_configure() {
    cudaError_t e;
    if ((e = configure_one<64, EPI_STORE>()) != cudaSuccess) return e;
    if ((e = configure_one<128, EPI_STORE>()) != cudaSuccess) return e;
    if ((e = configure_one<256, EPI_STORE>()) != cudaSuccess) return e;
    if ((e = configure_one<64, EPI_CONVT>()) != cudaSuccess) return e;
    if ((e = configure_one<128, EPI_CONVT>()) != cudaSuccess) return e;
    if ((e = configure_one<256, EPI_CONVT>()) != cudaSuccess) return e;
    if ((e = configure_one<64, EPI_OUTC>()) != cudaSuccess) return e;
    if ((e = configure_epg2<128, EPI_STORE>()) != cudaSuccess) return e;
    if ((e = configure_epg2<256, EPI_STORE>()) != cudaSuccess) return e;
    if ((e = configure_epg2<128, EPI_CONVT>()) != cudaSuccess) return e;
    if ((e = configure_epg2<256, EPI_CONVT>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv_convt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg::kSmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv_convt_pair_kernel<3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedPairCfg<3, 16>::kSmemBytes)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv_convt_pair_kernel<4, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedPairCfg<4, 12>::kSmemBytes)) != cudaSuccess) return e;
    {
        const char* off = getenv("CVB_NO_PAIR");
        g_pair_clusters = off && off[0] == '1' ? 0 : 1 << 20;
        if (g_pair_clusters) {
            if ((e = configure_pair<256, EPI_STORE>()) != cudaSuccess) return e;
            if ((e = configure_pair<128, EPI_STORE>()) != cudaSuccess) return e;
            if ((e = configure_pair<256, EPI_CONVT>()) != cudaSuccess) return e;
            if ((e = configure_pair<128, EPI_CONVT>()) != cudaSuccess) return e;
            if (g_pair_clusters == 1 << 20) g_pair_clusters = 0;
        }
    }
    if ((e = configure_vr<64, EPI_STORE, true>()) != cudaSuccess) return e;
    if ((e = configure_vr<64, EPI_STORE, false>()) != cudaSuccess) return e;
    if ((e = configure_vr<128, EPI_STORE, true>()) != cudaSuccess) return e;
    if ((e = configure_vr<128, EPI_STORE, false>()) != cudaSuccess) return e;
    if ((e = configure_vr<64, EPI_OUTC, true>()) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<128, EPI_STORE, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<128, EPI_STORE, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<256, EPI_STORE, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<128, EPI_STORE, true, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<128, EPI_STORE, false, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<128, EPI_STORE, true, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_vr_kernel<128, EPI_STORE, false, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_rs_kernel<EPI_STORE, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_rs_kernel<EPI_STORE, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_rs_kernel<EPI_STORE, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_rs_kernel<EPI_STORE, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_rs_kernel<EPI_OUTC, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(conv3x3_rs_kernel<EPI_OUTC, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kVrMaxSmem)) != cudaSuccess) return e;
    return cudaSuccess;
}

template <int BN, int EPI>
static cudaError_t launch_one(const ConvParams& p, int grid, cudaStream_t s, bool pdl) {
    return launch_k(conv_tc_kernel<BN, EPI>, grid, 256, ConvCfg<BN>::kSmemBytes, s, pdl, p);
}
template <int BN, int EPI>
static cudaError_t launch_one_epg2(const ConvParams& p, int grid, cudaStream_t s, bool pdl) {
    return launch_k(conv_tc_kernel<BN, EPI, 0, 2>, grid, 384, ConvCfg<BN>::kSmemBytes, s, pdl, p);
}
template <int BN, int EPI>
static cudaError_t launch_pair(const ConvParams& p, int clusters, cudaStream_t s, bool pdl) {
    return launch_kc(conv_tc_kernel<BN, EPI, 1>, 2 * clusters, 256, ConvCfg<BN, 1>::kSmemBytes, s, pdl, 2, p);
}

// Decide whether the vertical-reuse kernel applies to this launch and size its pipeline.
bool conv_try_vr(ConvLaunch& L, int ksize, int stride, int Ho, int Wo, int Cin) {
    ConvParams& p = L.p;
    if (ksize != 3 || stride != 1 || Ho % 16 || Wo % 8) return false;
    // N = 256 (the deep UNet levels): only as a CTA pair, where a stage (one activation box + three half weight tiles) is 66 KB
    // and three of them fit; against the generic pair kernel that is a third less L2 -> SM traffic (the activations are fetched
    // once per horizontal tap instead of once per tap).  CVB_VR256=0 keeps those layers on the generic kernel.
    const char* vr256 = getenv("CVB_VR256");
    const bool wide = L.block_n == 256 && L.epilogue == EPI_STORE && g_pair_clusters > 0 && !(vr256 && vr256[0] == '0');
    if (L.block_n != 64 && L.block_n != 128 && !wide) return false;
    if (L.epilogue != EPI_STORE && L.epilogue != EPI_OUTC) return false;
    // N = 128 with a plain store: CTA pairs (half of each weight tile per CTA); CVB_NO_PAIR_VR=1 keeps the single-CTA form
    const char* no_pair = getenv("CVB_NO_PAIR_VR");
    const char* min_cc = getenv("CVB_PAIR_VR_MIN_CC");   // A/B: pair form only from this many 64-channel input chunks up
    const char* epg2_ = getenv("CVB_VR_EPG2");
    const bool groups2 = L.block_n == 128 && L.epilogue == EPI_STORE && Cin <= 128 && epg2_ && epg2_[0] == '1';
    const bool pair = wide || (g_pair_clusters > 0 && L.block_n == 128 && L.epilogue == EPI_STORE && !(no_pair && no_pair[0] == '1') &&
                               Cin / 64 >= (min_cc ? atoi(min_cc) : (groups2 ? 1 : 2)));
    const int b_bytes = L.block_n * 128 / (pair ? 2 : 1);
    const int w_bytes = 9 * (Cin / 64) * b_bytes;
    // CVB_VR_EPG2=1 (A/B only): two epilogue groups for N = 128 with at most two input chunks.  Measured (profiles/README.md,
    // round 2): no gain -- down1.conv3 539 -> 526 us, and down1.conv0, which then has to run as a pair to fit the second
    // staging buffer, 280 -> 348 us: the epilogue is not what holds these layers at 69 % tensor activity.
    const bool two_groups = groups2;
    const int out_bytes = L.epilogue == EPI_OUTC ? 0 : (two_groups ? 2 : 1) * kOutBufBytes;
    const int budget = kVrMaxSmem - 1024 - 256 - kEpiConstBytes - out_bytes;
    // A/B: CVB_VR_PAIR_STREAM=1 streams the weights of a pair launch even when their halves would fit resident (resident
    // weights leave room for only three activation stages when Cin = 128)
    const char* pstream = getenv("CVB_VR_PAIR_STREAM");
    const bool ws = !wide && !(pair && pstream && pstream[0] == '1') && p.n_tiles == 1 && w_bytes + 3 * 18 * 1024 <= budget;
    if (L.epilogue == EPI_OUTC && !ws) return false;
    const int stage = 18 * 1024 + (ws ? 0 : 3 * b_bytes);
    int stages = (budget - (ws ? w_bytes : 0)) / stage;
    if (stages > 8) stages = 8;
    if (stages < 2) return false;
    p.vr_stages = stages;
    p.w_stationary = ws ? 1 : 0;
    p.smem_bytes = (ws ? w_bytes : 0) + stages * stage + out_bytes + 1024 + 256 + kEpiConstBytes;
    p.out_bufs = L.epilogue == EPI_OUTC ? 0 : 1;
    p.tn = 1; p.th = 16; p.tw = 8;
    p.tiles_w = Wo / 8;
    p.tiles_h = Ho / 16;
    L.variant = 1;
    L.pair = pair ? 1 : 0;
    L.vr_epg = two_groups ? 2 : 1;
    return true;
}

// Decide whether the row-streaming kernel applies (3x3 s1, Cout = 64, rows of a multiple of 128 pixels) and size it.
// CVB_NO_RS=1 disables it, CVB_RS_MODE=0|1 picks the activation staging (A/B measurements and tests); default 1.
bool conv_try_rs(ConvLaunch& L, int ksize, int stride, int Ho, int Wo, int Cin) {
    ConvParams& p = L.p;
    const char* off = getenv("CVB_NO_RS");
    if (off && off[0] == '1') return false;
    if (ksize != 3 || stride != 1 || L.block_n != 64 || p.n_tiles != 1) return false;
    if (L.epilogue != EPI_STORE && L.epilogue != EPI_OUTC) return false;
    // a streamed row of 128 pixels is either one 128-pixel segment of a wide image or the same row of eight 16x16 images
    const bool squares = Wo == 16 && Ho == 16;
    if (!squares && Wo % 128) return false;
    const int R = squares ? 16 : (Ho % 64 == 0 ? 64 : (Ho % 32 == 0 ? 32 : (Ho % 16 == 0 ? 16 : 0)));
    if (R == 0) return false;
    const char* m = getenv("CVB_RS_MODE");
    const int mode = squares ? 0 : (m ? atoi(m) : 1);   // the 130-pixel box of mode 1 needs contiguous pixels: wide images only
    if (mode < 0 || mode > 1) return false;
    const int stage = mode == 0 ? 49152 : 17408;
    const int w_bytes = 9 * (Cin / 64) * 64 * 128;
    const int fixed = 1024 + kRsBarBytes + kEpiConstBytes;
    int stages = (kVrMaxSmem - fixed - w_bytes) / stage;
    if (stages > 8) stages = 8;
    if (stages < 2) return false;
    p.vr_stages = stages;
    p.w_stationary = 1;
    p.rs_rows = R;
    p.rs_mode = mode;
    p.smem_bytes = w_bytes + stages * stage + fixed;
    p.out_bufs = 0;
    p.tn = squares ? 8 : 1; p.th = 1; p.tw = squares ? 16 : 128;
    p.tiles_w = Wo / p.tw;
    p.tiles_h = Ho / R;
    L.variant = 2;
    return true;
}

template <int BN, int EPI, bool WS>
static cudaError_t launch_vr(const ConvParams& p, int grid, cudaStream_t s, bool pdl) {
    return launch_k(conv3x3_vr_kernel<BN, EPI, WS>, grid, 256, p.smem_bytes, s, pdl, p);
}
template <int BN, bool WS, int EPG = 1>
static cudaError_t launch_vr_pair(const ConvParams& p, int clusters, cudaStream_t s, bool pdl) {
    return launch_kc(conv3x3_vr_kernel<BN, EPI_STORE, WS, 1, EPG>, 2 * clusters, 128 + 128 * EPG, p.smem_bytes, s, pdl, 2, p);
}

cudaError_t conv_launch(ConvLaunch& L, int n_images, int sm_count, cudaStream_t stream) {
    ConvParams& p = L.p;
    p.N = n_images;
    p.tiles_n = (n_images + p.tn - 1) / p.tn;
    p.idesc = umma_idesc_f16(128, L.block_n, 0);
    const long long total = 1LL * p.tiles_n * p.tiles_h * p.tiles_w * p.n_tiles;
    if (total <= 0) return cudaSuccess;
    const int grid = (int)(total < sm_count ? total : sm_count);
    const bool pdl = L.pdl != 0;
    if (L.epilogue == EPI_FUSED_CONVT) {
        if (L.variant != 0 || L.block_n != 128 || p.n_tiles != 1 || p.bias2 == nullptr) return cudaErrorInvalidValue;
        if (L.pair) {
            if (g_pair_clusters <= 0) return cudaErrorInvalidValue;
            const long long units = (total + 1) / 2;
            const int clusters = (int)(units < g_pair_clusters ? units : g_pair_clusters);
            static const bool deep_a = [] { const char* e = getenv("CVB_CONVT_DEEP_A"); return e && e[0] == '1'; }();   // A/B: 4 activation stages + 12 weight tiles
            if (deep_a) return launch_kc(conv_convt_pair_kernel<4, 12>, 2 * clusters, kFusedThreads, FusedPairCfg<4, 12>::kSmemBytes, stream, pdl, 2, p);
            return launch_kc(conv_convt_pair_kernel<3, 16>, 2 * clusters, kFusedThreads, FusedPairCfg<3, 16>::kSmemBytes, stream, pdl, 2, p);
        }
        return launch_k(conv_convt_kernel, grid, kFusedThreads, FusedCfg::kSmemBytes, stream, pdl, p);
    }
    if (L.variant == 2) {
        cudaError_t e;
        if (L.epilogue == EPI_OUTC) {
            if (p.rs_mode == 0) e = launch_k(conv3x3_rs_kernel<EPI_OUTC, 0, false>, grid, kRsThreads, p.smem_bytes, stream, pdl, p);
            else e = launch_k(conv3x3_rs_kernel<EPI_OUTC, 1, false>, grid, kRsThreads, p.smem_bytes, stream, pdl, p);
        } else if (p.pool_out != nullptr) {   // pooling form: no residual (the encoder convs have none)
            if (p.res != nullptr) return cudaErrorInvalidValue;
            if (p.rs_mode == 0) e = launch_k(conv3x3_rs_kernel<EPI_STORE, 0, true>, grid, kRsThreads, p.smem_bytes, stream, pdl, p);
            else e = launch_k(conv3x3_rs_kernel<EPI_STORE, 1, true>, grid, kRsThreads, p.smem_bytes, stream, pdl, p);
        } else {
            if (p.rs_mode == 0) e = launch_k(conv3x3_rs_kernel<EPI_STORE, 0, false>, grid, kRsThreads, p.smem_bytes, stream, pdl, p);
            else e = launch_k(conv3x3_rs_kernel<EPI_STORE, 1, false>, grid, kRsThreads, p.smem_bytes, stream, pdl, p);
        }
        return e;
    }
    if (L.variant == 1 && L.pair) {
        if (g_pair_clusters <= 0 || (L.block_n != 128 && L.block_n != 256) || L.epilogue != EPI_STORE) return cudaErrorInvalidValue;
        const long long units = ((1LL * p.tiles_n * p.tiles_h * p.tiles_w + 1) / 2) * p.n_tiles;
        const int clusters = (int)(units < g_pair_clusters ? units : g_pair_clusters);
        p.idesc = umma_idesc_f16(256, L.block_n, 0);
        if (L.block_n == 256) return p.w_stationary ? cudaErrorInvalidValue : launch_vr_pair<256, false>(p, clusters, stream, pdl);
        if (L.vr_epg == 2) return p.w_stationary ? launch_vr_pair<128, true, 2>(p, clusters, stream, pdl) : launch_vr_pair<128, false, 2>(p, clusters, stream, pdl);
        return p.w_stationary ? launch_vr_pair<128, true>(p, clusters, stream, pdl) : launch_vr_pair<128, false>(p, clusters, stream, pdl);
    }
    if (L.variant == 1) {
        const bool ws = p.w_stationary != 0;
        if (L.epilogue == EPI_OUTC) return launch_vr<64, EPI_OUTC, true>(p, grid, stream, pdl);
        if (L.block_n == 64) return ws ? launch_vr<64, EPI_STORE, true>(p, grid, stream, pdl) : launch_vr<64, EPI_STORE, false>(p, grid, stream, pdl);
        if (L.vr_epg == 2)
            return ws ? launch_k(conv3x3_vr_kernel<128, EPI_STORE, true, 0, 2>, grid, 384, p.smem_bytes, stream, pdl, p)
                      : launch_k(conv3x3_vr_kernel<128, EPI_STORE, false, 0, 2>, grid, 384, p.smem_bytes, stream, pdl, p);
        return ws ? launch_vr<128, EPI_STORE, true>(p, grid, stream, pdl) : launch_vr<128, EPI_STORE, false>(p, grid, stream, pdl);
    }
    if (L.pair) {
        if (g_pair_clusters <= 0) return cudaErrorInvalidValue;
        const long long m_tiles = 1LL * p.tiles_n * p.tiles_h * p.tiles_w;
        const long long units = ((m_tiles + 1) / 2) * p.n_tiles;
        const int clusters = (int)(units < g_pair_clusters ? units : g_pair_clusters);
        p.idesc = umma_idesc_f16(256, L.block_n, 0);
        switch (L.epilogue * 1000 + L.block_n) {
            case EPI_STORE * 1000 + 128: return launch_pair<128, EPI_STORE>(p, clusters, stream, pdl);
            case EPI_STORE * 1000 + 256: return launch_pair<256, EPI_STORE>(p, clusters, stream, pdl);
            case EPI_CONVT * 1000 + 128: return launch_pair<128, EPI_CONVT>(p, clusters, stream, pdl);
            case EPI_CONVT * 1000 + 256: return launch_pair<256, EPI_CONVT>(p, clusters, stream, pdl);
            default: return cudaErrorInvalidValue;
        }
    }
    if (L.epg == 2) {
        switch (L.epilogue * 1000 + L.block_n) {
            case EPI_STORE * 1000 + 128: return launch_one_epg2<128, EPI_STORE>(p, grid, stream, pdl);
            case EPI_STORE * 1000 + 256: return launch_one_epg2<256, EPI_STORE>(p, grid, stream, pdl);
            case EPI_CONVT * 1000 + 128: return launch_one_epg2<128, EPI_CONVT>(p, grid, stream, pdl);
            case EPI_CONVT * 1000 + 256: return launch_one_epg2<256, EPI_CONVT>(p, grid, stream, pdl);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (L.epilogue * 1000 + L.block_n) {
        case EPI_STORE * 1000 + 64: return launch_one<64, EPI_STORE>(p, grid, stream, pdl);
        case EPI_STORE * 1000 + 128: return launch_one<128, EPI_STORE>(p, grid, stream, pdl);
        case EPI_STORE * 1000 + 256: return launch_one<256, EPI_STORE>(p, grid, stream, pdl);
        case EPI_CONVT * 1000 + 64: return launch_one<64, EPI_CONVT>(p, grid, stream, pdl);
        case EPI_CONVT * 1000 + 128: return launch_one<128, EPI_CONVT>(p, grid, stream, pdl);
        case EPI_CONVT * 1000 + 256: return launch_one<256, EPI_CONVT>(p, grid, stream, pdl);
        case EPI_OUTC * 1000 + 64: return launch_one<64, EPI_OUTC>(p, grid, stream, pdl);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cvb
