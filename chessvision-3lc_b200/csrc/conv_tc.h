// Host-side description of one tcgen05 implicit-GEMM convolution launch (internal to the library).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

enum ConvEpilogue : int {
    EPI_STORE = 0,   // y = acc + bias (+ residual) (ReLU optional) -> fp16 NHWC with channel stride/offset
    EPI_CONVT = 1,   // 2x2 stride-2 transposed conv: column block q=(dy,dx) scattered to pixel (2h+dy, 2w+dx)
    EPI_OUTC = 2,    // ReLU(acc + bias) . w_out + b_out -> fp32 logit + u8 mask (BLOCK_N == 64 == Cout)
    EPI_FUSED_CONVT = 3,   // ReLU(acc + bias) stays on chip as the fp16 A operand of a second MMA: the ConvTranspose2d(k2, s2)
                           // that follows the layer, scattered into the concat buffer (conv_convt_kernel; Cout == 128)
};

// Kernel parameter block (passed __grid_constant__; the tensor maps must stay 64-byte aligned).
struct alignas(64) ConvParams {
    CUtensorMap a_map[4];  // activation views, 4-D {C, W, H, N}, box {64, tw, th, tn}, 128-byte swizzle
    CUtensorMap b_map;     // packed weights, 2-D {K_total, Cout_total}, box {64, BLOCK_N}, 128-byte swizzle
    CUtensorMap o_map[4];  // output views for the TMA tile store, box {64, tw, th, tn}; convT: one strided view per (dy,dx)
    CUtensorMap b2_map;    // EPI_FUSED_CONVT: packed transposed-conv weights, 2-D {Cin2 = 128, 4 * Cout2 = 256}, box {64, 128}
    const float* bias2;    // EPI_FUSED_CONVT: [Cout2] bias of the transposed convolution
    int out_bufs;          // 16 KB staging buffers for the store (0: EPI_OUTC, 1 or 2 otherwise)
    // K loop: taps x (Cin/64) chunks.  Tap t reads view tap_map[t] at spatial offset (tap_dy[t], tap_dx[t]).
    int taps;
    int c_chunks;
    int a_c_off;  // first input channel inside the activation buffer
    int8_t tap_map[12];
    int8_t tap_dy[12];
    int8_t tap_dx[12];
    // M tiling: one tile = tn images x th rows x tw columns = 128 output pixels
    int tn, th, tw;
    int tiles_w, tiles_h, tiles_n;
    int N, H, W;   // output extent covered by the M tiles (N = images in this launch)
    int n_tiles;   // Cout_total / BLOCK_N
    uint32_t idesc;
    int vr_stages;      // vertical-reuse / row-streaming variants: pipeline depth, weights resident in smem, dynamic smem size
    int w_stationary;
    int smem_bytes;
    int rs_rows;        // row-streaming variant: output rows per strip
    int rs_mode;        // ... activation staging: 0 = one box per (row, dx), 1 = one 130-pixel box per row
    // epilogue
    int relu;
    const float* bias;      // [Cout_total]
    __half* out;            // NHWC fp16
    int out_c_stride;       // channels per pixel of the output buffer
    int out_c_off;          // first output channel
    const __half* res;      // optional residual, NHWC fp16 (same pixel grid as out)
    __half* pool_out;       // row-streaming kernel only: also write MaxPool2d(2) of the output, NHWC fp16 [N,H/2,W/2,pool_c_stride]
    int pool_c_stride;
    int res_c_stride;
    int convt_cout;         // EPI_CONVT: Cout per (dy,dx) block
    const float* outc_w;    // EPI_OUTC: [64]
    float outc_b;
    float* logits;          // EPI_OUTC: [N,H,W] fp32
    uint8_t* mask;          // EPI_OUTC: [N,H,W] u8 {0,255}
    float thr;
};

struct ConvLaunch {
    ConvParams p;
    int block_n;   // 64, 128 or 256
    int epilogue;  // ConvEpilogue
    int variant;   // 0 generic kernel, 1 vertical-reuse 3x3 kernel, 2 row-streaming 3x3 kernel (Cout = 64)
    int n_max;     // images the activation / output views were built for
    int vr_epg;    // vertical-reuse kernel: epilogue groups (A/B switch)
    int epg;       // generic kernel: epilogue groups (2 for N >= 128 launches with a short K loop)
    int pair;      // generic kernel as CTA pairs (cta_group::2, M = 256 per MMA; each CTA loads half of the weight tile)
    int pdl;       // launch with programmatic stream serialization (resident weights are then requested before the previous
                   // kernel has finished: only for launches whose weights no kernel in the stream writes, i.e. inference)
};

// Resolve cuTensorMapEncodeTiled through the runtime (no link-time dependency on libcuda).  Returns 0 on success.
int tmap_init();
// 4-D NHWC activation view.  Strides are in elements of fp16 between consecutive w / h / n.
int tmap_act(CUtensorMap* m, const void* base, int C, int Wv, int Hv, int Nv, int64_t sW, int64_t sH, int64_t sN,
             int tw, int th, int tn);
// 2-D K-major weight matrix [rows = Cout_total][cols = K_total].
int tmap_weights(CUtensorMap* m, const void* base, int K_total, int rows, int block_n);

// Vertical-reuse 3x3 kernel: 4-D view with box {64, 8, 18, 1}; conv_try_vr switches a built launch over to it.
int tmap_act_vr(CUtensorMap* m, const void* base, int C, int Wv, int Hv, int Nv, int64_t sW, int64_t sH, int64_t sN);
bool conv_try_vr(ConvLaunch& L, int ksize, int stride, int Ho, int Wo, int Cin);
// Row-streaming 3x3 kernel (Cout = 64, W % 128 == 0): views with boxes {64, 128, 1, 1} and {64, 130, 1, 1}.
bool conv_try_rs(ConvLaunch& L, int ksize, int stride, int Ho, int Wo, int Cin);
// Generic kernel with BLOCK_N = 128 / 256: switch the launch to the CTA-pair form when the device runs clusters of two
// (conv_configure decides; CVB_NO_PAIR=1 turns it off).  Returns 0 or a tensor-map error.
int conv_try_pair(ConvLaunch& L, const __half* w, int K, int rows);
int conv_pair_clusters();

// Host-side construction of a launch (api.cu / train.cu).  Return 0, -5 (shape not supported) or a tensor-map error.
int conv_build(ConvLaunch& L, const __half* in, int Nmax, int Hin, int Win, int in_c_stride, int in_c_off, int Cin, const __half* w,
               const float* bias, int rows, int K, int ksize, int stride, int epilogue, bool use_vr);
int conv_build_k2s2(ConvLaunch& L, const __half* in, int Nmax, int Ho, int Wo, int in_c_stride, int in_c_off, int C, const __half* w,
                    const float* bias, int rows, int K);
int conv_set_store(ConvLaunch& L, __half* out, int out_c_stride, int out_c_off, int relu, const __half* res, int res_c_stride);
// EPI_FUSED_CONVT: the launch built for conv3x3(Cin -> 128) + bias + ReLU additionally applies ConvTranspose2d(128 -> cout2 = 64,
// k2, s2) (w2 packed [(dy*2+dx)*cout2 + co][128], bias2 [cout2]) and writes channels [out_c_off, out_c_off + cout2) of the
// double-resolution NHWC buffer `out`; the conv's own output never reaches global memory.
int conv_set_fused_convt(ConvLaunch& L, const __half* w2, const float* bias2, int cout2, __half* out, int out_c_stride, int out_c_off);

// Launch on `stream` for the first `n_images` images; grid sized to min(tiles, SM count).
cudaError_t conv_launch(ConvLaunch& L, int n_images, int sm_count, cudaStream_t stream);
// One-time: raise the dynamic shared-memory limit of every instantiation.
cudaError_t conv_configure();

}  // namespace cvb
