// CUDA-core kernels around the tensor-core convolutions: the tiny-K first layers of both networks (fused with the
// integer preprocessing), max-pooling, the logits->mask rule and the classifier tail (avg-pool + FC + softmax +
// argmax + rule 1 + FEN).  All of them are HBM/L2-bound byte work, written as coalesced, vectorised kernels.
#include "kernels.h"

#include <cuda_fp16.h>

namespace cvb {

// ----------------------------------------------------------------------------------------------------------------
// K0 + inc.double_conv.0:  cv2.resize(INTER_AREA 2x) -> /255 -> Conv3x3(3->64, pad 1) -> BN -> ReLU -> fp16 NHWC
// reference: core.py:212-216 (resize, /255, BGR kept) + unet_parts.py:16-18
// img  u8 [N, 2H, 2W, 3]     out fp16 [N, H, W, out_c_stride] channels [0,64)
// wf   fp32 [27][64] (k = (r*3+s)*3 + ci, BN scale folded)    bf fp32 [64]
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unet_stem(const uint8_t* __restrict__ img, const float* __restrict__ wf,
                                                   const float* __restrict__ bf, __half* __restrict__ out, int H, int W,
                                                   int out_c_stride) {
    __shared__ float s_in[18][18][3];
    __shared__ __align__(16) float s_w[27 * 64];
    __shared__ __align__(16) float s_b[64];
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * 16, y0 = blockIdx.y * 16;
    const int tid = threadIdx.y * 16 + threadIdx.x;
    for (int i = tid; i < 27 * 64; i += 256) s_w[i] = wf[i];
    if (tid < 64) s_b[tid] = bf[tid];
    const uint8_t* src = img + static_cast<size_t>(n) * (2 * H) * (2 * W) * 3;
    for (int i = tid; i < 18 * 18; i += 256) {
        const int ly = i / 18, lx = i % 18;
        const int y = y0 + ly - 1, x = x0 + lx - 1;
        float v[3] = {0.f, 0.f, 0.f};
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const uint8_t* p0 = src + (static_cast<size_t>(2 * y) * (2 * W) + 2 * x) * 3;
            const uint8_t* p1 = p0 + static_cast<size_t>(2 * W) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int s = p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2;  // INTER_AREA for an exact 2x reduction
                v[c] = __fdiv_rn(static_cast<float>(s >> 2), 255.0f);
            }
        }
        s_in[ly][lx][0] = v[0];
        s_in[ly][lx][1] = v[1];
        s_in[ly][lx][2] = v[2];
    }
    __syncthreads();
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = s_b[c];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int s = 0; s < 3; ++s) {
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float xin = s_in[threadIdx.y + r][threadIdx.x + s][ci];
                const float4* w4 = reinterpret_cast<const float4*>(&s_w[((r * 3 + s) * 3 + ci) * 64]);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float4 w = w4[c];
                    acc[4 * c + 0] = fmaf(xin, w.x, acc[4 * c + 0]);
                    acc[4 * c + 1] = fmaf(xin, w.y, acc[4 * c + 1]);
                    acc[4 * c + 2] = fmaf(xin, w.z, acc[4 * c + 2]);
                    acc[4 * c + 3] = fmaf(xin, w.w, acc[4 * c + 3]);
                }
            }
        }
    }
    const int y = y0 + threadIdx.y, x = x0 + threadIdx.x;
    uint4* dst = reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * H + y) * W + x) * out_c_stride);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            oh[j] = __floats2half2_rn(fmaxf(acc[8 * i + 2 * j], 0.f), fmaxf(acc[8 * i + 2 * j + 1], 0.f));
        dst[i] = o;
    }
}

cudaError_t launch_unet_stem(const uint8_t* img, const float* wf, const float* bf, __half* out, int N, int H, int W,
                             int out_c_stride, cudaStream_t s) {
    dim3 grid(W / 16, H / 16, N), block(16, 16);
    k_unet_stem<<<grid, block, 0, s>>>(img, wf, bf, out, H, W, out_c_stride);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// MaxPool2d(2) on fp16 NHWC (unet_parts.py:34).  in [N,H,W,in_c_stride] channels [0,C)  ->  out [N,H/2,W/2,C]
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_maxpool2(const __half* __restrict__ in, __half* __restrict__ out, int H, int W,
                                                  int C, int in_c_stride, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c8 = C / 8;
    const int cv = static_cast<int>(i % c8);
    long long pix = i / c8;
    const int Wo = W / 2, Ho = H / 2;
    const int x = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int y = static_cast<int>(pix % Ho);
    const long long n = pix / Ho;
    const __half* p = in + ((n * H + 2 * y) * W + 2 * x) * in_c_stride + cv * 8;
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(p + in_c_stride));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(p + static_cast<size_t>(W) * in_c_stride));
    const uint4 d = __ldg(reinterpret_cast<const uint4*>(p + static_cast<size_t>(W) * in_c_stride + in_c_stride));
    uint4 o;
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    const __half2* ch = reinterpret_cast<const __half2*>(&c);
    const __half2* dh = reinterpret_cast<const __half2*>(&d);
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __hmax2(__hmax2(ah[j], bh[j]), __hmax2(ch[j], dh[j]));
    *reinterpret_cast<uint4*>(out + ((n * Ho + y) * Wo + x) * C + cv * 8) = o;
}

cudaError_t launch_maxpool2(const __half* in, __half* out, int N, int H, int W, int C, int in_c_stride, cudaStream_t s) {
    const long long total = 1LL * N * (H / 2) * (W / 2) * (C / 8);
    if (total == 0) return cudaSuccess;
    k_maxpool2<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, H, W, C, in_c_stride, total);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// sigmoid(logit) > thr -> {0,255}   (core.py:273 + utils.py:101-112), for externally supplied logits
// ----------------------------------------------------------------------------------------------------------------
__global__ void k_mask_from_logits(const float* __restrict__ logits, uint8_t* __restrict__ mask, float thr, long long n4) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 v = __ldg(reinterpret_cast<const float4*>(logits) + i);
    uchar4 o;
    o.x = 1.0f / (1.0f + expf(-v.x)) > thr ? 255 : 0;
    o.y = 1.0f / (1.0f + expf(-v.y)) > thr ? 255 : 0;
    o.z = 1.0f / (1.0f + expf(-v.z)) > thr ? 255 : 0;
    o.w = 1.0f / (1.0f + expf(-v.w)) > thr ? 255 : 0;
    reinterpret_cast<uchar4*>(mask)[i] = o;
}

cudaError_t launch_mask_from_logits(const float* logits, uint8_t* mask, float thr, long long count, cudaStream_t s) {
    const long long n4 = count / 4;
    if (n4 == 0) return cudaSuccess;
    k_mask_from_logits<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, s>>>(logits, mask, thr, n4);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// ResNet-18 stem per square: /255 -> Conv7x7 s2 p3 (1->64) -> BN -> ReLU -> MaxPool3x3 s2 p1 -> fp16 NHWC [16,16,64]
// reference: core.py:232-237 (extract_squares, /255) + timm resnet18 conv1/bn1/act1/maxpool
// board u8 [N,512,512]; square index q = 8*row + col (core.py:420-439); out fp16 [N*64,16,16,64]
// wf fp32 [49][64] (BN scale folded), bf fp32 [64]
// ----------------------------------------------------------------------------------------------------------------
constexpr int kStemThreads = 512;
constexpr int kStemInStride = 72;
constexpr int kStemSmem = 70 * kStemInStride * 4 + 49 * 64 * 4 + 64 * 4 + 32 * 32 * 64 * 2;

__global__ void __launch_bounds__(kStemThreads, 1) k_resnet_stem(const uint8_t* __restrict__ board, const float* __restrict__ wf,
                                                                const float* __restrict__ bf, __half* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* s_in = reinterpret_cast<float*>(smem);                       // [70][72], s_in[r][c] = px(r-3, c-3)/255
    float* s_w = s_in + 70 * kStemInStride;                             // [49][64]
    float* s_b = s_w + 49 * 64;                                         // [64]
    __half* s_conv = reinterpret_cast<__half*>(s_b + 64);               // [32][32][64] post-ReLU
    const int sq = blockIdx.x;
    const int n = sq >> 6, q = sq & 63;
    const int tid = threadIdx.x;
    const uint8_t* src = board + (static_cast<size_t>(n) * 512 + (q >> 3) * 64) * 512 + (q & 7) * 64;
    for (int i = tid; i < 49 * 64; i += kStemThreads) s_w[i] = wf[i];
    if (tid < 64) s_b[tid] = bf[tid];
    for (int i = tid; i < 70 * kStemInStride; i += kStemThreads) {
        const int r = i / kStemInStride - 3, c = i % kStemInStride - 3;
        float v = 0.f;
        if (r >= 0 && r < 64 && c >= 0 && c < 64) v = __fdiv_rn(static_cast<float>(src[r * 512 + c]), 255.0f);
        s_in[i] = v;
    }
    __syncthreads();
    // task = (channel half, conv row, pixel pair); 1024 tasks over 512 threads
    for (int task = tid; task < 1024; task += kStemThreads) {
        const int xp = task & 15, y = (task >> 4) & 31, half = task >> 9;
        float a0[32], a1[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) a0[c] = a1[c] = s_b[half * 32 + c];
        for (int ky = 0; ky < 7; ++ky) {
            const float* row = s_in + (2 * y + ky) * kStemInStride + 4 * xp;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float i0 = row[kx], i1 = row[kx + 2];
                const float4* w4 = reinterpret_cast<const float4*>(s_w + (ky * 7 + kx) * 64 + half * 32);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 w = w4[c];
                    a0[4 * c + 0] = fmaf(i0, w.x, a0[4 * c + 0]);
                    a0[4 * c + 1] = fmaf(i0, w.y, a0[4 * c + 1]);
                    a0[4 * c + 2] = fmaf(i0, w.z, a0[4 * c + 2]);
                    a0[4 * c + 3] = fmaf(i0, w.w, a0[4 * c + 3]);
                    a1[4 * c + 0] = fmaf(i1, w.x, a1[4 * c + 0]);
                    a1[4 * c + 1] = fmaf(i1, w.y, a1[4 * c + 1]);
                    a1[4 * c + 2] = fmaf(i1, w.z, a1[4 * c + 2]);
                    a1[4 * c + 3] = fmaf(i1, w.w, a1[4 * c + 3]);
                }
            }
        }
        __half2* d0 = reinterpret_cast<__half2*>(s_conv + ((y * 32 + 2 * xp) * 64 + half * 32));
        __half2* d1 = d0 + 32;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            d0[c] = __floats2half2_rn(fmaxf(a0[2 * c], 0.f), fmaxf(a0[2 * c + 1], 0.f));
            d1[c] = __floats2half2_rn(fmaxf(a1[2 * c], 0.f), fmaxf(a1[2 * c + 1], 0.f));
        }
    }
    __syncthreads();
    // 3x3 stride-2 pad-1 max pool -> [16][16][64]; values are >= 0 so padding can be treated as 0
    __half2* dst = reinterpret_cast<__half2*>(out + static_cast<size_t>(sq) * 16 * 16 * 64);
    const __half2* cv = reinterpret_cast<const __half2*>(s_conv);
    for (int i = tid; i < 16 * 16 * 32; i += kStemThreads) {
        const int c2 = i & 31, px = (i >> 5) & 15, py = i >> 9;
        __half2 m = __floats2half2_rn(0.f, 0.f);
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            const int yy = 2 * py + dy;
            if (yy < 0) continue;
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                const int xx = 2 * px + dx;
                if (xx < 0) continue;
                m = __hmax2(m, cv[(yy * 32 + xx) * 32 + c2]);
            }
        }
        dst[i] = m;
    }
}

cudaError_t configure_resnet_stem() {
    return cudaFuncSetAttribute(k_resnet_stem, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmem);
}

cudaError_t launch_resnet_stem(const uint8_t* board, const float* wf, const float* bf, __half* out, int n_boards,
                               cudaStream_t s) {
    if (n_boards == 0) return cudaSuccess;
    k_resnet_stem<<<n_boards * 64, kStemThreads, kStemSmem, s>>>(board, wf, bf, out);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// Classifier tail, one CTA per board: global avg-pool (2x2) -> Linear 512->13 -> softmax -> argmax -> rule 1 -> FEN
// reference: timm resnet18 global_pool+fc; core.py:241-243 (softmax), :326-327 (argmax), :442-469 (rule 1),
//            :330-349 (python-chess board_fen).
// feat fp16 [N*64, 2, 2, 512];  fcw fp32 [13][512];  fcb fp32 [13]
// probs f32 [N,64,13]; labels/labels_valid u8 [N,64]; fen char [N,2,72] (0 = original, 1 = validated, NUL padded)
// ----------------------------------------------------------------------------------------------------------------
__constant__ char c_label_chars[13] = {'B', 'K', 'N', 'P', 'Q', 'R', 'b', 'k', 'n', 'p', 'q', 'r', 'f'};

__global__ void __launch_bounds__(256) k_head(const __half* __restrict__ feat, const float* __restrict__ fcw,
                                              const float* __restrict__ fcb, float* __restrict__ probs,
                                              uint8_t* __restrict__ labels, uint8_t* __restrict__ labels_valid,
                                              char* __restrict__ fen, int flip) {
    __shared__ float s_w[13 * 512];
    __shared__ uint8_t s_lab[2][64];
    const int n = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 13 * 512; i += 256) s_w[i] = fcw[i];
    __syncthreads();
    const int sq = tid >> 2, part = tid & 3;  // 4 threads per square, 128 channels each
    const __half* f = feat + (static_cast<size_t>(n) * 64 + sq) * 4 * 512 + part * 128;
    float acc[13];
#pragma unroll
    for (int k = 0; k < 13; ++k) acc[k] = 0.f;
    for (int c = 0; c < 128; c += 8) {
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = 0.f;
#pragma unroll
        for (int px = 0; px < 4; ++px) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(f + px * 512 + c));
            const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 t = __half22float2(h[j]);
                m[2 * j] += t.x;
                m[2 * j + 1] += t.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float mean = m[j] * 0.25f;
#pragma unroll
            for (int k = 0; k < 13; ++k) acc[k] = fmaf(mean, s_w[k * 512 + part * 128 + c + j], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 13; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
    }
    if (part == 0) {
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            acc[k] += __ldg(fcb + k);
            mx = fmaxf(mx, acc[k]);
        }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            acc[k] = expf(acc[k] - mx);
            sum += acc[k];
        }
        int best = 0, best_np = -1;
        float pb = -1.f, pnp = -1.f;
        float* dst = probs + (static_cast<size_t>(n) * 64 + sq) * 13;
#pragma unroll
        for (int k = 0; k < 13; ++k) {
            const float pr = acc[k] / sum;
            dst[k] = pr;
            if (pr > pb) { pb = pr; best = k; }                         // first maximum wins (np.argmax)
            if (k != 3 && k != 9 && pr >= pnp) { pnp = pr; best_np = k; }  // last maximum wins (reversed argsort)
        }
        const bool end_rank = sq < 8 || sq >= 56;  // ranks 8 and 1 in either orientation (constants.py:88-105)
        const int fixed = (end_rank && (best == 3 || best == 9)) ? best_np : best;
        s_lab[0][sq] = static_cast<uint8_t>(best);
        s_lab[1][sq] = static_cast<uint8_t>(fixed);
        labels[n * 64 + sq] = static_cast<uint8_t>(best);
        labels_valid[n * 64 + sq] = static_cast<uint8_t>(fixed);
    }
    __syncthreads();
    if (tid < 2) {
        char* o = fen + (static_cast<size_t>(n) * 2 + tid) * 72;
        int pos = 0;
        for (int r = 0; r < 8; ++r) {
            int empty = 0;
            for (int c = 0; c < 8; ++c) {
                const int j = r * 8 + c;                     // FEN order: rank 8 first, file a first
                const int lab = s_lab[tid][flip ? 63 - j : j];  // SQUARE_NAMES_FLIPPED is the reversed table
                if (lab == 12) {
                    ++empty;
                } else {
                    if (empty) o[pos++] = static_cast<char>('0' + empty);
                    empty = 0;
                    o[pos++] = c_label_chars[lab];
                }
            }
            if (empty) o[pos++] = static_cast<char>('0' + empty);
            if (r < 7) o[pos++] = '/';
        }
        while (pos < 72) o[pos++] = 0;
    }
}

cudaError_t launch_head(const __half* feat, const float* fcw, const float* fcb, float* probs, uint8_t* labels,
                        uint8_t* labels_valid, char* fen, int n_boards, int flip, cudaStream_t s) {
    if (n_boards == 0) return cudaSuccess;
    k_head<<<n_boards, 256, 0, s>>>(feat, fcw, fcb, probs, labels, labels_valid, fen, flip);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// cv2.resize(img, (w,h), INTER_AREA) for an exact 2x reduction (core.py:212): (a+b+c+d+2)>>2 per channel
// ----------------------------------------------------------------------------------------------------------------
__global__ void k_resize_area_half(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int h, int w, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int x = static_cast<int>(i % w);
    const long long t = i / w;
    const int y = static_cast<int>(t % h);
    const long long n = t / h;
    const uint8_t* p0 = img + ((n * (2 * h) + 2 * y) * (2LL * w) + 2 * x) * 3;
    const uint8_t* p1 = p0 + 2LL * w * 3;
    uint8_t* o = out + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
}

// ----------------------------------------------------------------------------------------------------------------
// cv2.resize(img, (dw,dh), INTER_AREA) for any reduction (core.py:212), bit-identical to OpenCV's C++ path
// (imgproc/resize.cpp: computeResizeAreaTab + ResizeArea_Invoker<uchar,float>, ResizeAreaFast_Invoker for integer scale
// factors).  The cell tables (source index, float32 weight per destination index) are built on the host with OpenCV's own
// double arithmetic; the kernel replays its float32 accumulation order with separate multiplies and adds (no FMA):
// per source row the horizontal products in table order, then sum = beta*buf for the first source row of a destination
// row and sum += beta*buf after it, round-half-even at the end.  One thread per destination pixel.
// ----------------------------------------------------------------------------------------------------------------
__global__ void k_resize_area(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int H, int W, int dh, int dw,
                              const int* __restrict__ xofs, const int* __restrict__ xsi, const float* __restrict__ xa,
                              const int* __restrict__ yofs, const int* __restrict__ ysi, const float* __restrict__ ya, int int_area,
                              long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int dx = static_cast<int>(i % dw);
    const long long t = i / dw;
    const int dy = static_cast<int>(t % dh);
    const long long n = t / dh;
    const uint8_t* src = img + n * static_cast<long long>(H) * W * 3;
    uint8_t* o = out + i * 3;
    if (int_area) {   // integer scale factors: whole cells, integer sums
        const int ix = W / dw, iy = H / dh;
        int sum[3] = {0, 0, 0};
        for (int r = 0; r < iy; ++r) {
            const uint8_t* p = src + (static_cast<long long>(dy) * iy + r) * W * 3 + static_cast<long long>(dx) * ix * 3;
            for (int c = 0; c < ix; ++c) {
                sum[0] += p[3 * c];
                sum[1] += p[3 * c + 1];
                sum[2] += p[3 * c + 2];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int v;
            if (int_area == 4) v = (sum[c] + 2) >> 2;
            else v = __float2int_rn(__fmul_rn(static_cast<float>(sum[c]), 1.0f / static_cast<float>(int_area)));
            o[c] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
        return;
    }
    float sum[3] = {0.f, 0.f, 0.f};
    const int y0 = yofs[dy], y1 = yofs[dy + 1], x0 = xofs[dx], x1 = xofs[dx + 1];
    for (int j = y0; j < y1; ++j) {
        const uint8_t* row = src + static_cast<long long>(ysi[j]) * W * 3;
        float buf[3] = {0.f, 0.f, 0.f};
        for (int k = x0; k < x1; ++k) {
            const uint8_t* p = row + xsi[k] * 3;
            const float a = xa[k];
            buf[0] = __fadd_rn(buf[0], __fmul_rn(static_cast<float>(p[0]), a));
            buf[1] = __fadd_rn(buf[1], __fmul_rn(static_cast<float>(p[1]), a));
            buf[2] = __fadd_rn(buf[2], __fmul_rn(static_cast<float>(p[2]), a));
        }
        const float b = ya[j];
#pragma unroll
        for (int c = 0; c < 3; ++c) sum[c] = j == y0 ? __fmul_rn(b, buf[c]) : __fadd_rn(sum[c], __fmul_rn(b, buf[c]));
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int v = __float2int_rn(sum[c]);
        o[c] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

cudaError_t launch_resize_area(const uint8_t* img, uint8_t* out, int N, int H, int W, int dh, int dw, const int* xofs, const int* xsi,
                               const float* xa, const int* yofs, const int* ysi, const float* ya, int int_area, cudaStream_t s) {
    const long long total = 1LL * N * dh * dw;
    if (total == 0) return cudaSuccess;
    k_resize_area<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(img, out, H, W, dh, dw, xofs, xsi, xa, yofs, ysi, ya, int_area, total);
    return cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
// cv2.resize(..., INTER_AREA) when at least one axis is ENLARGED (inputs smaller than 256 px, core.py:212 accepts them).
// OpenCV only has true area interpolation for reductions in both axes; otherwise it runs its fixed-point bilinear
// resizer (imgproc/resize.cpp, resizeGeneric_ with HResizeLinear / VResizeLinear<uchar, int, short>) with "area mode"
// coefficients built on the host (api.cu: linear_area_table): per destination column a source column and two 11-bit
// weights, the same per row, and
//     h(row, dx) = S[sx]*a0 + S[sx+1]*a1        (S[sx]*2048 for dx >= xmax, where sx+1 is outside the row)
//     out        = (((b0 * (h(r0) >> 4)) >> 16) + ((b1 * (h(r1) >> 4)) >> 16) + 2) >> 2
// in 32-bit integers.  One thread per destination pixel.
// ----------------------------------------------------------------------------------------------------------------
__global__ void k_resize_linear_area(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int H, int W, int dh, int dw,
                                     const int* __restrict__ xtab, const int* __restrict__ ytab, int xmax, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int dx = static_cast<int>(i % dw);
    const long long t = i / dw;
    const int dy = static_cast<int>(t % dh);
    const long long n = t / dh;
    const uint8_t* src = img + n * static_cast<long long>(H) * W * 3;
    const int sx = xtab[3 * dx], a0 = xtab[3 * dx + 1], a1 = xtab[3 * dx + 2];
    const int sy = ytab[3 * dy], b0 = ytab[3 * dy + 1], b1 = ytab[3 * dy + 2];
    const int r0 = min(max(sy, 0), H - 1), r1 = min(max(sy + 1, 0), H - 1);
    const uint8_t* p0 = src + (static_cast<long long>(r0) * W + sx) * 3;
    const uint8_t* p1 = src + (static_cast<long long>(r1) * W + sx) * 3;
    const bool edge = dx >= xmax;
    uint8_t* o = out + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = edge ? p0[c] * 2048 : p0[c] * a0 + p0[c + 3] * a1;
        const int h1 = edge ? p1[c] * 2048 : p1[c] * a0 + p1[c + 3] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

cudaError_t launch_resize_linear_area(const uint8_t* img, uint8_t* out, int N, int H, int W, int dh, int dw, const int* xtab,
                                      const int* ytab, int xmax, cudaStream_t s) {
    const long long total = 1LL * N * dh * dw;
    if (total == 0) return cudaSuccess;
    k_resize_linear_area<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(img, out, H, W, dh, dw, xtab, ytab, xmax, total);
    return cudaGetLastError();
}

// u8 [N,h,w,3] -> u8 [N,2h,2w,3] by pixel replication: the exact inverse of the 2x INTER_AREA reduction fused into the
// UNet stem ((4a+2)>>2 == a), so an image resized by k_resize_area enters the standard pipeline unchanged.
__global__ void k_double2x(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int h, int w, long long total) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // one thread per OUTPUT pixel
    if (i >= total) return;
    const int x = static_cast<int>(i % (2 * w));
    const long long t = i / (2 * w);
    const int y = static_cast<int>(t % (2 * h));
    const long long n = t / (2 * h);
    const uint8_t* p = in + ((n * h + (y >> 1)) * w + (x >> 1)) * 3;
    uint8_t* o = out + i * 3;
    o[0] = p[0];
    o[1] = p[1];
    o[2] = p[2];
}

cudaError_t launch_double2x(const uint8_t* in, uint8_t* out, int N, int h, int w, cudaStream_t s) {
    const long long total = 4LL * N * h * w;
    if (total == 0) return cudaSuccess;
    k_double2x<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(in, out, h, w, total);
    return cudaGetLastError();
}

cudaError_t launch_resize_area_half(const uint8_t* img, uint8_t* out, int N, int h, int w, cudaStream_t s) {
    const long long total = 1LL * N * h * w;
    if (total == 0) return cudaSuccess;
    k_resize_area_half<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(img, out, h, w, total);
    return cudaGetLastError();
}

}  // namespace cvb
