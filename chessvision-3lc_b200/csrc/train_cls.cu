// Piece-classifier training step on the device: the body of the reference's loop for one batch
//     output = model(data); loss = CrossEntropyLoss()(output, target); loss.backward(); optimizer.step()   (Adam, lr from StepLR)
// (scripts/train/train_classifier.py:63-88,218-221) for timm resnet18(num_classes=13, in_chans=1) in model.train() state
// (chessvision/utils.py:32-39), and the same forward in model.eval() state for the validation loop (:91-113).
//
// This is row n3 of SURVEY.md 8(f) -- a consumer either side of the hot path, not the hot path itself: everything is fp32 on the
// CUDA cores (the numbers of the reference's fp32 training run, no mixed precision to reason about), with the three
// convolution passes as one tiled implicit-GEMM kernel:
//   forward   z[(n,ho,wo)][co]      = sum_k x[n, ho*s-p+r, wo*s-p+q, ci] * W[co][(r,q,ci)]
//   dgrad     dx[(n,h,w)][ci]       = sum_k dz[n, (h+p-r)/s, (w+p-q)/s, co] * W[co][(r,q,ci)]        (k = (r,q,co))
//   wgrad     dW[co][(r,q,ci)]      = sum_j dz[j][co] * x[gather(j, r,q,ci)]                          (j = (n,ho,wo), split + reduce)
// Activations are NHWC fp32; weights live in one flat buffer as [co][r][q][ci] (converted from / to torch's [co][ci][r][q] at
// the boundary), with their gradients and the Adam moments in buffers of the same layout.
#include "ctx.h"

#include <math.h>
#include <string.h>

#include <string>
#include <vector>

namespace cvb {
namespace {

// ------------------------------------------------------------------------------------------------------------ implicit GEMM
struct GemmArgs {
    const float* a;     // fwd: x; dgrad: dz; wgrad: dz
    const float* b;     // fwd / dgrad: W; wgrad: x
    float* c;           // fwd: z; dgrad: dx; wgrad: partial sums [split][Cout][K]
    int B, Hin, Win, Cin, Hout, Wout, Cout, ks, stride, pad;
    int M, N, K;        // GEMM extents of this mode
    int accumulate;     // dgrad: add to dx instead of overwriting
    int k_per_split;    // wgrad: reduction rows per grid.z slice
};

enum { MODE_FWD = 0, MODE_DGRAD = 1, MODE_WGRAD = 2 };
constexpr int BM = 64, BN = 64, BK = 16;

template <int MODE>
__device__ __forceinline__ float load_a(const GemmArgs& g, int m, int k) {
    if (MODE == MODE_FWD) {
        const int wo = m % g.Wout, ho = (m / g.Wout) % g.Hout, n = m / (g.Wout * g.Hout);
        const int ci = k % g.Cin, rq = k / g.Cin, q = rq % g.ks, r = rq / g.ks;
        const int h = ho * g.stride - g.pad + r, w = wo * g.stride - g.pad + q;
        if (h < 0 || h >= g.Hin || w < 0 || w >= g.Win) return 0.f;
        return g.a[((static_cast<size_t>(n) * g.Hin + h) * g.Win + w) * g.Cin + ci];
    } else if (MODE == MODE_DGRAD) {
        const int w = m % g.Win, h = (m / g.Win) % g.Hin, n = m / (g.Win * g.Hin);
        const int co = k % g.Cout, rq = k / g.Cout, q = rq % g.ks, r = rq / g.ks;
        const int hs = h + g.pad - r, ws = w + g.pad - q;
        if (hs < 0 || ws < 0 || hs % g.stride || ws % g.stride) return 0.f;
        const int ho = hs / g.stride, wo = ws / g.stride;
        if (ho >= g.Hout || wo >= g.Wout) return 0.f;
        return g.a[((static_cast<size_t>(n) * g.Hout + ho) * g.Wout + wo) * g.Cout + co];
    } else {   // wgrad: A(m = co, k = j) = dz[j][co]
        return g.a[static_cast<size_t>(k) * g.Cout + m];
    }
}

template <int MODE>
__device__ __forceinline__ float load_b(const GemmArgs& g, int k, int n) {
    if (MODE == MODE_FWD) {
        return g.b[static_cast<size_t>(n) * g.K + k];   // W[co = n][k]
    } else if (MODE == MODE_DGRAD) {
        const int co = k % g.Cout, rq = k / g.Cout;
        return g.b[(static_cast<size_t>(co) * g.ks * g.ks + rq) * g.Cin + n];   // W[co][(r,q)][ci = n]
    } else {   // wgrad: B(k = j, n = (r,q,ci)) = x[gather]
        const int wo = k % g.Wout, ho = (k / g.Wout) % g.Hout, b = k / (g.Wout * g.Hout);
        const int ci = n % g.Cin, rq = n / g.Cin, q = rq % g.ks, r = rq / g.ks;
        const int h = ho * g.stride - g.pad + r, w = wo * g.stride - g.pad + q;
        if (h < 0 || h >= g.Hin || w < 0 || w >= g.Win) return 0.f;
        return g.b[((static_cast<size_t>(b) * g.Hin + h) * g.Win + w) * g.Cin + ci];
    }
}

// C[M x N] = A[M x K] . B[K x N] with the operands gathered as above; 64 x 64 tile, 16-deep K steps, 4 x 4 outputs per thread.
// FAST (every layer but the 7x7 / Cin = 1 stem): channel counts are multiples of 16 and the output extent is a power of two,
// so a K step lies inside one (r, q) tap -- the tap decode is uniform per K step, the pixel decode of a thread's four tile rows
// is done once (forward / dgrad) or with shifts (wgrad), and an element costs an address computation instead of five divisions.
template <int MODE, bool FAST>
__global__ void __launch_bounds__(256) k_cls_gemm(const GemmArgs g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int k_lo = 0, k_hi = g.K;
    if (MODE == MODE_WGRAD) {
        k_lo = blockIdx.z * g.k_per_split;
        k_hi = min(g.K, k_lo + g.k_per_split);
    }
    // per-thread decode of the four A rows (forward: output pixel; dgrad: input pixel) it loads in every K step
    int ph[4], pw[4];
    const float* pbase[4];
    bool pvalid[4];
    if (FAST && MODE != MODE_WGRAD) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int gm = m0 + ((tid + 256 * e) >> 4);
            pvalid[e] = gm < g.M;
            const int Wm = MODE == MODE_FWD ? g.Wout : g.Win, Hm = MODE == MODE_FWD ? g.Hout : g.Hin;
            const int w = gm % Wm, h = (gm / Wm) % Hm, n = gm / (Wm * Hm);
            if (MODE == MODE_FWD) {
                ph[e] = h * g.stride - g.pad;
                pw[e] = w * g.stride - g.pad;
                pbase[e] = g.a + static_cast<size_t>(n) * g.Hin * g.Win * g.Cin;
            } else {
                ph[e] = h + g.pad;
                pw[e] = w + g.pad;
                pbase[e] = g.a + static_cast<size_t>(n) * g.Hout * g.Wout * g.Cout;
            }
        }
    }
    float acc[4][4] = {};
    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
        if (FAST) {
            if (MODE == MODE_FWD) {
                const int rq = k0 / g.Cin, c0 = k0 - rq * g.Cin, r = rq / g.ks, q = rq - r * g.ks;
                const int ak = tid & 15;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int am = (tid + 256 * e) >> 4;
                    const int h = ph[e] + r, w = pw[e] + q;
                    const bool ok = pvalid[e] && h >= 0 && h < g.Hin && w >= 0 && w < g.Win;
                    As[ak][am] = ok ? pbase[e][(static_cast<size_t>(h) * g.Win + w) * g.Cin + c0 + ak] : 0.f;
                    const int bn = am;   // the same split of the tile index for W[co][k]
                    Bs[ak][bn] = n0 + bn < g.N ? g.b[static_cast<size_t>(n0 + bn) * g.K + k0 + ak] : 0.f;
                }
            } else if (MODE == MODE_DGRAD) {
                const int rq = k0 / g.Cout, c0 = k0 - rq * g.Cout, r = rq / g.ks, q = rq - r * g.ks;
                const int ak = tid & 15;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int am = (tid + 256 * e) >> 4;
                    const int hs = ph[e] - r, ws = pw[e] - q;
                    bool ok = pvalid[e] && hs >= 0 && ws >= 0;
                    int ho = hs, wo = ws;
                    if (g.stride == 2) {
                        ok = ok && !((hs | ws) & 1);
                        ho >>= 1;
                        wo >>= 1;
                    }
                    ok = ok && ho < g.Hout && wo < g.Wout;
                    As[ak][am] = ok ? pbase[e][(static_cast<size_t>(ho) * g.Wout + wo) * g.Cout + c0 + ak] : 0.f;
                    const int bn = tid & 63, bk = (tid >> 6) + 4 * e;   // W[co0 + bk][rq][ci = n0 + bn]
                    Bs[bk][bn] = n0 + bn < g.N ? g.b[(static_cast<size_t>(c0 + bk) * g.ks * g.ks + rq) * g.Cin + n0 + bn] : 0.f;
                }
            } else {   // wgrad: A(co, j) = dz[j][co]; B(j, (rq, ci)) = x[gather]; the 64 columns of this block lie in one tap
                const int rq = n0 / g.Cin, c0 = n0 - rq * g.Cin, r = rq / g.ks, q = rq - r * g.ks;
                const int lw = 31 - __clz(g.Wout), lh = 31 - __clz(g.Hout);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int am = tid & 63, ak = (tid >> 6) + 4 * e;
                    const int j = k0 + ak;
                    const bool in = j < k_hi;
                    As[ak][am] = (in && m0 + am < g.M) ? g.a[static_cast<size_t>(j) * g.Cout + m0 + am] : 0.f;
                    const int wo = j & (g.Wout - 1), ho = (j >> lw) & (g.Hout - 1), bi = j >> (lw + lh);
                    const int h = ho * g.stride - g.pad + r, w = wo * g.stride - g.pad + q;
                    const bool ok = in && h >= 0 && h < g.Hin && w >= 0 && w < g.Win && n0 + am < g.N;
                    Bs[ak][am] = ok ? g.b[((static_cast<size_t>(bi) * g.Hin + h) * g.Win + w) * g.Cin + c0 + am] : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int idx = tid + 256 * e;
                // A tile: consecutive threads along the dimension that is contiguous in memory for this mode
                int am, ak;
                if (MODE == MODE_WGRAD) { am = idx & (BM - 1); ak = idx >> 6; } else { ak = idx & (BK - 1); am = idx >> 4; }
                const int gm = m0 + am, gk = k0 + ak;
                As[ak][am] = (gm < g.M && gk < k_hi) ? load_a<MODE>(g, gm, gk) : 0.f;
                int bk, bn;
                if (MODE == MODE_FWD) { bk = idx & (BK - 1); bn = idx >> 4; } else { bn = idx & (BN - 1); bk = idx >> 6; }
                const int hk = k0 + bk, hn = n0 + bn;
                Bs[bk][bn] = (hk < k_hi && hn < g.N) ? load_b<MODE>(g, hk, hn) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* c = g.c + (MODE == MODE_WGRAD ? static_cast<size_t>(blockIdx.z) * g.M * g.N : 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float* dst = c + static_cast<size_t>(m) * g.N + n;
            *dst = (MODE == MODE_DGRAD && g.accumulate) ? *dst + acc[i][j] : acc[i][j];
        }
    }
}

// dW = sum over the wgrad splits (fixed order: deterministic)
__global__ void k_cls_reduce_splits(const float* __restrict__ part, float* __restrict__ out, size_t count, int splits) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[static_cast<size_t>(z) * count + i];
    out[i] = s;
}

// ------------------------------------------------------------------------------------------------------------ BatchNorm
// Column sums over the M rows of an [M][C] matrix: out[c] = sum_m f(m,c); two of them at once.  One block per 32 channels and
// row slice; the slices are combined by k_cls_bn_finish in a fixed order.
// stats:    s0 = sum z,            s1 = sum z^2
// backward: s0 = sum dy_eff,       s1 = sum dy_eff * xhat       (dy_eff = dy * (act > 0) when act != nullptr)
template <bool BWD>
__global__ void __launch_bounds__(256) k_cls_colsum(const float* __restrict__ z, const float* __restrict__ dy, const float* __restrict__ act,
                                                   const float* __restrict__ mean, const float* __restrict__ invstd, int M, int C,
                                                   float* __restrict__ part /* [slices][2][C] */) {
    __shared__ float sh[2][8][32];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
    const int slices = gridDim.y, rows = (M + slices - 1) / slices;
    const int m_lo = blockIdx.y * rows, m_hi = min(M, m_lo + rows);
    float s0 = 0.f, s1 = 0.f;
    if (c < C) {
        const float mu = BWD ? mean[c] : 0.f, is = BWD ? invstd[c] : 0.f;
        for (int m = m_lo + rl; m < m_hi; m += 8) {
            const size_t i = static_cast<size_t>(m) * C + c;
            if (BWD) {
                float d = dy[i];
                if (act != nullptr && !(act[i] > 0.f)) d = 0.f;
                s0 += d;
                s1 += d * (z[i] - mu) * is;
            } else {
                const float v = z[i];
                s0 += v;
                s1 += v * v;
            }
        }
    }
    sh[0][rl][threadIdx.x & 31] = s0;
    sh[1][rl][threadIdx.x & 31] = s1;
    __syncthreads();
    if (rl == 0 && c < C) {
        float t0 = 0.f, t1 = 0.f;
        for (int r = 0; r < 8; ++r) { t0 += sh[0][r][threadIdx.x]; t1 += sh[1][r][threadIdx.x]; }
        part[(static_cast<size_t>(blockIdx.y) * 2 + 0) * C + c] = t0;
        part[(static_cast<size_t>(blockIdx.y) * 2 + 1) * C + c] = t1;
    }
}

// training-mode statistics: mean, 1/sqrt(var + eps) (biased variance), running statistics with momentum (unbiased variance)
__global__ void k_cls_bn_finish_stats(const float* __restrict__ part, int slices, int M, int C, float eps, float momentum,
                                      float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ run_mean,
                                      float* __restrict__ run_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int s = 0; s < slices; ++s) { s0 += part[(static_cast<size_t>(s) * 2) * C + c]; s1 += part[(static_cast<size_t>(s) * 2 + 1) * C + c]; }
    const double mu = s0 / M;
    double var = s1 / M - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[c] = static_cast<float>(mu);
    invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * static_cast<float>(mu);
    const double unbiased = M > 1 ? var * M / (M - 1) : var;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * static_cast<float>(unbiased);
}

// eval mode: the running statistics take the place of the batch statistics
__global__ void k_cls_bn_eval_stats(const float* __restrict__ run_mean, const float* __restrict__ run_var, int C, float eps,
                                    float* __restrict__ mean, float* __restrict__ invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = run_mean[c];
    invstd[c] = rsqrtf(run_var[c] + eps);
}

// out = [relu]( gamma * (z - mean) * invstd + beta [+ residual] )
__global__ void k_cls_bn_apply(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ invstd,
                               const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ residual,
                               int relu, size_t count, int C, float* __restrict__ out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    const int c = static_cast<int>(i % C);
    float v = gamma[c] * (z[i] - mean[c]) * invstd[c] + beta[c];
    if (residual != nullptr) v += residual[i];
    out[i] = relu ? fmaxf(v, 0.f) : v;
}

// dgamma, dbeta from the two column sums; they also parameterise the data gradient below
__global__ void k_cls_bn_finish_bwd(const float* __restrict__ part, int slices, int C, float* __restrict__ sum_dy, float* __restrict__ sum_dyx,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double s0 = 0.0, s1 = 0.0;
    for (int s = 0; s < slices; ++s) { s0 += part[(static_cast<size_t>(s) * 2) * C + c]; s1 += part[(static_cast<size_t>(s) * 2 + 1) * C + c]; }
    sum_dy[c] = static_cast<float>(s0);
    sum_dyx[c] = static_cast<float>(s1);
    dbeta[c] = static_cast<float>(s0);
    dgamma[c] = static_cast<float>(s1);
}

// dz = gamma * invstd * (dy_eff - sum_dy / M - xhat * sum_dyx / M); optionally also passes dy_eff on (the shortcut's gradient)
__global__ void k_cls_bn_bwd_data(const float* __restrict__ z, const float* __restrict__ dy, const float* __restrict__ act,
                                  const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
                                  const float* __restrict__ sum_dy, const float* __restrict__ sum_dyx, size_t count, int C, int M,
                                  float* __restrict__ dz, float* __restrict__ dy_eff_out) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    const int c = static_cast<int>(i % C);
    float d = dy[i];
    if (act != nullptr && !(act[i] > 0.f)) d = 0.f;
    if (dy_eff_out != nullptr) dy_eff_out[i] = d;
    const float xhat = (z[i] - mean[c]) * invstd[c];
    const float inv_m = 1.f / static_cast<float>(M);
    dz[i] = gamma[c] * invstd[c] * (d - sum_dy[c] * inv_m - xhat * sum_dyx[c] * inv_m);
}

// a += b
__global__ void k_cls_add(float* __restrict__ a, const float* __restrict__ b, size_t count) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < count) a[i] += b[i];
}

// ------------------------------------------------------------------------------------------------------------ max-pool 3x3 s2 p1
__global__ void k_cls_maxpool_fwd(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ y, uint8_t* __restrict__ arg) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(B) * Ho * Wo * C) return;
    const int c = static_cast<int>(i % C);
    const int wo = static_cast<int>((i / C) % Wo), ho = static_cast<int>((i / (static_cast<size_t>(C) * Wo)) % Ho);
    const int n = static_cast<int>(i / (static_cast<size_t>(C) * Wo * Ho));
    float best = -INFINITY;
    int bi = 0;
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) {
            const int h = 2 * ho - 1 + r, w = 2 * wo - 1 + q;
            if (h < 0 || h >= H || w < 0 || w >= W) continue;
            const float v = x[((static_cast<size_t>(n) * H + h) * W + w) * C + c];
            if (v > best) { best = v; bi = r * 3 + q; }   // first maximum in window order, as ATen's kernel picks it
        }
    y[i] = best;
    arg[i] = static_cast<uint8_t>(bi);
}

// gather form of the backward pass: an input pixel collects the gradients of the (up to four) windows whose maximum it was
__global__ void k_cls_maxpool_bwd(const float* __restrict__ dy, const uint8_t* __restrict__ arg, int B, int H, int W, int C,
                                  float* __restrict__ dx) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<size_t>(B) * H * W * C) return;
    const int c = static_cast<int>(i % C);
    const int w = static_cast<int>((i / C) % W), h = static_cast<int>((i / (static_cast<size_t>(C) * W)) % H);
    const int n = static_cast<int>(i / (static_cast<size_t>(C) * W * H));
    float s = 0.f;
    for (int ho = h / 2; ho <= (h + 1) / 2; ++ho)
        for (int wo = w / 2; wo <= (w + 1) / 2; ++wo) {
            if (ho >= Ho || wo >= Wo) continue;
            const int r = h - (2 * ho - 1), q = w - (2 * wo - 1);
            if (r < 0 || r > 2 || q < 0 || q > 2) continue;
            const size_t j = ((static_cast<size_t>(n) * Ho + ho) * Wo + wo) * C + c;
            if (arg[j] == r * 3 + q) s += dy[j];
        }
    dx[i] = s;
}

// ------------------------------------------------------------------------------------------------------------ head + loss
// global average pool + fc + softmax cross entropy for one sample per block (C = 512 channels, HW pixels, 13 classes)
constexpr int kClasses = 13;
__global__ void __launch_bounds__(128) k_cls_head_fwd(const float* __restrict__ x, int HW, int C, const float* __restrict__ fw,
                                                     const float* __restrict__ fb, const int32_t* __restrict__ target,
                                                     float* __restrict__ pooled, float* __restrict__ logits, float* __restrict__ probs,
                                                     float* __restrict__ loss_each, int32_t* __restrict__ correct_each) {
    extern __shared__ float sp[];   // [C]
    __shared__ float sl[kClasses];
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += x[(static_cast<size_t>(n) * HW + p) * C + c];
        s /= static_cast<float>(HW);
        sp[c] = s;
        pooled[static_cast<size_t>(n) * C + c] = s;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < kClasses; j += 4) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += sp[c] * fw[static_cast<size_t>(j) * C + c];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sl[j] = s + fb[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mx = sl[0];
        int am = 0;
        for (int j = 1; j < kClasses; ++j)
            if (sl[j] > mx) { mx = sl[j]; am = j; }
        float den = 0.f;
        for (int j = 0; j < kClasses; ++j) den += expf(sl[j] - mx);
        const int t = target ? min(max(target[n], 0), kClasses - 1) : 0;   // (an out-of-range label must not index past the logits)
        for (int j = 0; j < kClasses; ++j) {
            logits[n * kClasses + j] = sl[j];
            probs[n * kClasses + j] = expf(sl[j] - mx) / den;
        }
        loss_each[n] = target ? -(sl[t] - mx - logf(den)) : 0.f;
        correct_each[n] = target ? (am == t ? 1 : 0) : 0;
    }
}

// mean loss and number of correct predictions of the batch (one block)
__global__ void k_cls_loss_reduce(const float* __restrict__ loss_each, const int32_t* __restrict__ correct_each, int B, float* __restrict__ loss,
                                  int32_t* __restrict__ correct) {
    if (threadIdx.x != 0) return;
    double s = 0.0;
    int c = 0;
    for (int n = 0; n < B; ++n) { s += loss_each[n]; c += correct_each[n]; }
    *loss = static_cast<float>(s / B);
    if (correct) *correct = c;
}

// dlogits = (softmax - onehot) / B; fc gradients; gradient of the pooled features spread over the HW pixels
__global__ void k_cls_head_bwd_logits(const float* __restrict__ probs, const int32_t* __restrict__ target, int B, float* __restrict__ dlogits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * kClasses) return;
    const int n = i / kClasses, j = i % kClasses;
    dlogits[i] = (probs[i] - (min(max(target[n], 0), kClasses - 1) == j ? 1.f : 0.f)) / static_cast<float>(B);
}
__global__ void k_cls_head_bwd_fc(const float* __restrict__ dlogits, const float* __restrict__ pooled, int B, int C, float* __restrict__ dfw,
                                  float* __restrict__ dfb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // (j, c)
    if (i < kClasses * C) {
        const int j = i / C, c = i % C;
        float s = 0.f;
        for (int n = 0; n < B; ++n) s += dlogits[n * kClasses + j] * pooled[static_cast<size_t>(n) * C + c];
        dfw[i] = s;
    }
    if (i < kClasses) {
        float s = 0.f;
        for (int n = 0; n < B; ++n) s += dlogits[n * kClasses + i];
        dfb[i] = s;
    }
}
__global__ void k_cls_head_bwd_x(const float* __restrict__ dlogits, const float* __restrict__ fw, int B, int HW, int C, float* __restrict__ dx) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;   // (n, p, c)
    if (i >= static_cast<size_t>(B) * HW * C) return;
    const int c = static_cast<int>(i % C), n = static_cast<int>(i / (static_cast<size_t>(C) * HW));
    float s = 0.f;
    for (int j = 0; j < kClasses; ++j) s += dlogits[n * kClasses + j] * fw[static_cast<size_t>(j) * C + c];
    dx[i] = s / static_cast<float>(HW);
}

// ------------------------------------------------------------------------------------------------------------ Adam (torch.optim.Adam)
__global__ void k_cls_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t count,
                           float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale, float bias1, float bias2_sqrt) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= count) return;
    float gi = g[i] * grad_scale;
    if (weight_decay != 0.f) gi += weight_decay * p[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias2_sqrt + eps;
    p[i] -= (lr / bias1) * (mi / denom);
}

}  // namespace
}  // namespace cvb

// ====================================================================================================================
// host side
// ====================================================================================================================
using namespace cvb;

namespace {

struct ClsConv {
    std::string name;
    int cin, cout, ks, stride, pad, hin, hout;
    size_t w_off;   // offset of W[co][r][q][ci] in the flat buffers
};
struct ClsBn {
    std::string name;
    int c;
    size_t g_off, b_off;         // gamma, beta in the flat buffers
    float *run_mean, *run_var;   // [c]
    float *mean, *invstd;        // batch (or running) statistics of the last forward
    float *sum_dy, *sum_dyx;     // backward column sums
    int64_t batches_tracked;
};
struct ClsBlock {
    int conv1, bn1, conv2, bn2, convd, bnd;   // indices; convd = -1 without downsample
    int h_in, h_out, c_in, c_out;
    float *z1, *a1, *z2, *zd, *sd, *out;      // conv outputs (pre-BN), activations
};

}  // namespace

struct cvb_cls_trainer {
    cvb_cls_train_config cfg;
    int B = 0;
    std::vector<ClsConv> convs;
    std::vector<ClsBn> bns;
    std::vector<ClsBlock> blocks;
    size_t n_params = 0;
    size_t fc_w_off = 0, fc_b_off = 0;
    float *P = nullptr, *G = nullptr, *M1 = nullptr, *M2 = nullptr;
    int64_t step = 0;
    // stem
    float *x0 = nullptr, *z0 = nullptr, *a0 = nullptr, *p0 = nullptr;
    uint8_t* arg0 = nullptr;
    // head
    float *pooled = nullptr, *logits = nullptr, *probs = nullptr, *loss_each = nullptr, *dlogits = nullptr;
    int32_t* correct_each = nullptr;
    float* d_loss = nullptr;
    int32_t* d_correct = nullptr;
    // scratch
    float *part = nullptr;        // column-sum slices [64][2][512]
    float *wpart = nullptr;       // wgrad splits
    size_t wpart_floats = 0;
    float *gA = nullptr, *gB = nullptr, *gC = nullptr;   // gradient ping-pong buffers (largest activation)
    std::vector<void*> allocs;
    bool have_forward = false;
};

namespace {

constexpr int kColSlices = 64;

template <class T>
int talloc(cvb_ctx* ctx, cvb_cls_trainer* t, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, -3, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    t->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
}

const cvb_tensor* find_t(const cvb_tensor* sd, int n, const std::string& name) {
    for (int i = 0; i < n; ++i)
        if (name == sd[i].name) return &sd[i];
    return nullptr;
}
int64_t numel_t(const cvb_tensor* t) {
    int64_t c = 1;
    for (int i = 0; i < t->ndim; ++i) c *= t->shape[i];
    return c;
}

inline unsigned blocks_for(size_t count, int threads = 256) { return static_cast<unsigned>((count + threads - 1) / threads); }

int add_conv(cvb_cls_trainer* t, const std::string& name, int cin, int cout, int ks, int stride, int hin) {
    ClsConv c;
    c.name = name; c.cin = cin; c.cout = cout; c.ks = ks; c.stride = stride; c.pad = ks / 2; c.hin = hin; c.hout = hin / stride;
    c.w_off = t->n_params;
    t->n_params += static_cast<size_t>(cout) * ks * ks * cin;
    t->convs.push_back(c);
    return static_cast<int>(t->convs.size()) - 1;
}
int add_bn(cvb_cls_trainer* t, const std::string& name, int c) {
    ClsBn b;
    b.name = name; b.c = c;
    b.g_off = t->n_params; t->n_params += c;
    b.b_off = t->n_params; t->n_params += c;
    b.run_mean = b.run_var = b.mean = b.invstd = b.sum_dy = b.sum_dyx = nullptr;
    b.batches_tracked = 0;
    t->bns.push_back(b);
    return static_cast<int>(t->bns.size()) - 1;
}

// ---- launches
cudaError_t conv_gemm(int mode, const ClsConv& c, int B, const float* a, const float* b, float* out, int accumulate, cvb_cls_trainer* t,
                      cudaStream_t s) {
    GemmArgs g;
    g.a = a; g.b = b; g.c = out;
    g.B = B; g.Hin = c.hin; g.Win = c.hin; g.Cin = c.cin; g.Hout = c.hout; g.Wout = c.hout; g.Cout = c.cout;
    g.ks = c.ks; g.stride = c.stride; g.pad = c.pad; g.accumulate = accumulate; g.k_per_split = 0;
    const int Kw = c.ks * c.ks * c.cin;
    // every layer but the stem: channels in multiples of 64, power-of-two output extent, stride 1 or 2
    const bool fast = c.cin % 64 == 0 && c.cout % 64 == 0 && (c.hout & (c.hout - 1)) == 0 && (c.stride == 1 || c.stride == 2);
    if (mode == MODE_FWD) {
        g.M = B * c.hout * c.hout; g.N = c.cout; g.K = Kw;
        dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
        if (fast) k_cls_gemm<MODE_FWD, true><<<grid, 256, 0, s>>>(g);
        else k_cls_gemm<MODE_FWD, false><<<grid, 256, 0, s>>>(g);
    } else if (mode == MODE_DGRAD) {
        g.M = B * c.hin * c.hin; g.N = c.cin; g.K = c.ks * c.ks * c.cout;
        dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
        if (fast) k_cls_gemm<MODE_DGRAD, true><<<grid, 256, 0, s>>>(g);
        else k_cls_gemm<MODE_DGRAD, false><<<grid, 256, 0, s>>>(g);
    } else {
        g.M = c.cout; g.N = Kw; g.K = B * c.hout * c.hout;
        // enough slices of the pixel dimension to fill the GPU, each a multiple of the K step
        const int tiles = ((g.N + BN - 1) / BN) * ((g.M + BM - 1) / BM);
        int splits = (4 * 148 + tiles - 1) / tiles;
        const int max_splits = static_cast<int>(t->wpart_floats / (static_cast<size_t>(g.M) * g.N));
        if (splits > max_splits) splits = max_splits;
        if (splits > (g.K + BK - 1) / BK) splits = (g.K + BK - 1) / BK;
        if (splits < 1) splits = 1;
        g.k_per_split = (((g.K + splits - 1) / splits) + BK - 1) / BK * BK;
        splits = (g.K + g.k_per_split - 1) / g.k_per_split;
        g.c = t->wpart;
        dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, splits);
        if (fast) k_cls_gemm<MODE_WGRAD, true><<<grid, 256, 0, s>>>(g);
        else k_cls_gemm<MODE_WGRAD, false><<<grid, 256, 0, s>>>(g);
        const size_t count = static_cast<size_t>(g.M) * g.N;
        k_cls_reduce_splits<<<blocks_for(count), 256, 0, s>>>(t->wpart, out, count, splits);
    }
    return cudaGetLastError();
}

// z -> batch / running statistics -> out = [relu](bn(z) [+ residual])
cudaError_t bn_forward(cvb_cls_trainer* t, ClsBn& b, const float* z, int M, int training, const float* residual, int relu, float* out,
                       cudaStream_t s) {
    const int C = b.c;
    if (training) {
        dim3 grid((C + 31) / 32, kColSlices);
        k_cls_colsum<false><<<grid, 256, 0, s>>>(z, nullptr, nullptr, nullptr, nullptr, M, C, t->part);
        k_cls_bn_finish_stats<<<(C + 127) / 128, 128, 0, s>>>(t->part, kColSlices, M, C, t->cfg.bn_eps, t->cfg.bn_momentum, b.mean, b.invstd,
                                                              b.run_mean, b.run_var);
        b.batches_tracked++;
    } else {
        k_cls_bn_eval_stats<<<(C + 127) / 128, 128, 0, s>>>(b.run_mean, b.run_var, C, t->cfg.bn_eps, b.mean, b.invstd);
    }
    const size_t count = static_cast<size_t>(M) * C;
    k_cls_bn_apply<<<blocks_for(count), 256, 0, s>>>(z, b.mean, b.invstd, t->P + b.g_off, t->P + b.b_off, residual, relu, count, C, out);
    return cudaGetLastError();
}

// dy (masked by act > 0 when act != nullptr) -> dgamma, dbeta, dz (and the masked dy itself when dy_eff_out != nullptr)
cudaError_t bn_backward(cvb_cls_trainer* t, ClsBn& b, const float* z, const float* dy, const float* act, int M, float* dz, float* dy_eff_out,
                        cudaStream_t s) {
    const int C = b.c;
    dim3 grid((C + 31) / 32, kColSlices);
    k_cls_colsum<true><<<grid, 256, 0, s>>>(z, dy, act, b.mean, b.invstd, M, C, t->part);
    k_cls_bn_finish_bwd<<<(C + 127) / 128, 128, 0, s>>>(t->part, kColSlices, C, b.sum_dy, b.sum_dyx, t->G + b.g_off, t->G + b.b_off);
    const size_t count = static_cast<size_t>(M) * C;
    k_cls_bn_bwd_data<<<blocks_for(count), 256, 0, s>>>(z, dy, act, b.mean, b.invstd, t->P + b.g_off, b.sum_dy, b.sum_dyx, count, C, M, dz,
                                                      dy_eff_out);
    return cudaGetLastError();
}

int cls_forward(cvb_ctx* ctx, cvb_cls_trainer* t, const float* x, const int32_t* target, int training, cudaStream_t s) {
    const int B = t->B;
    // input [B,1,64,64] NCHW == [B,64,64,1] NHWC
    CK(cudaMemcpyAsync(t->x0, x, static_cast<size_t>(B) * 4096 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CK(conv_gemm(MODE_FWD, t->convs[0], B, t->x0, t->P + t->convs[0].w_off, t->z0, 0, t, s));
    CK(bn_forward(t, t->bns[0], t->z0, B * 32 * 32, training, nullptr, 1, t->a0, s));
    k_cls_maxpool_fwd<<<blocks_for(static_cast<size_t>(B) * 16 * 16 * 64), 256, 0, s>>>(t->a0, B, 32, 32, 64, t->p0, t->arg0);
    CK(cudaGetLastError());
    const float* xin = t->p0;
    for (auto& k : t->blocks) {
        const int Mo = B * k.h_out * k.h_out;
        CK(conv_gemm(MODE_FWD, t->convs[k.conv1], B, xin, t->P + t->convs[k.conv1].w_off, k.z1, 0, t, s));
        CK(bn_forward(t, t->bns[k.bn1], k.z1, Mo, training, nullptr, 1, k.a1, s));
        CK(conv_gemm(MODE_FWD, t->convs[k.conv2], B, k.a1, t->P + t->convs[k.conv2].w_off, k.z2, 0, t, s));
        const float* shortcut = xin;
        if (k.convd >= 0) {
            CK(conv_gemm(MODE_FWD, t->convs[k.convd], B, xin, t->P + t->convs[k.convd].w_off, k.zd, 0, t, s));
            CK(bn_forward(t, t->bns[k.bnd], k.zd, Mo, training, nullptr, 0, k.sd, s));
            shortcut = k.sd;
        }
        CK(bn_forward(t, t->bns[k.bn2], k.z2, Mo, training, shortcut, 1, k.out, s));
        xin = k.out;
    }
    const ClsBlock& last = t->blocks.back();
    k_cls_head_fwd<<<B, 128, last.c_out * sizeof(float), s>>>(last.out, last.h_out * last.h_out, last.c_out, t->P + t->fc_w_off, t->P + t->fc_b_off,
                                                              target, t->pooled, t->logits, t->probs, t->loss_each, t->correct_each);
    CK(cudaGetLastError());
    k_cls_loss_reduce<<<1, 32, 0, s>>>(t->loss_each, t->correct_each, B, t->d_loss, t->d_correct);
    CK(cudaGetLastError());
    return 0;
}

int cls_backward(cvb_ctx* ctx, cvb_cls_trainer* t, const int32_t* target, cudaStream_t s) {
    const int B = t->B;
    const ClsBlock& last = t->blocks.back();
    const int C = last.c_out, HW = last.h_out * last.h_out;
    k_cls_head_bwd_logits<<<blocks_for(static_cast<size_t>(B) * kClasses), 256, 0, s>>>(t->probs, target, B, t->dlogits);
    k_cls_head_bwd_fc<<<blocks_for(static_cast<size_t>(kClasses) * C), 256, 0, s>>>(t->dlogits, t->pooled, B, C, t->G + t->fc_w_off, t->G + t->fc_b_off);
    float* dout = t->gA;   // gradient w.r.t. the current block's output
    k_cls_head_bwd_x<<<blocks_for(static_cast<size_t>(B) * HW * C), 256, 0, s>>>(t->dlogits, t->P + t->fc_w_off, B, HW, C, dout);
    CK(cudaGetLastError());
    for (int bi = static_cast<int>(t->blocks.size()) - 1; bi >= 0; --bi) {
        ClsBlock& k = t->blocks[bi];
        const float* xin = bi == 0 ? t->p0 : t->blocks[bi - 1].out;
        const int Mo = B * k.h_out * k.h_out;
        float* dz = t->gB;      // gradient w.r.t. a conv output
        float* dsum = t->gC;    // dout masked by the block's final ReLU = gradient of (bn2(z2) + shortcut)
        // out = relu(bn2(z2) + shortcut)
        CK(bn_backward(t, t->bns[k.bn2], k.z2, dout, k.out, Mo, dz, dsum, s));
        CK(conv_gemm(MODE_WGRAD, t->convs[k.conv2], B, dz, k.a1, t->G + t->convs[k.conv2].w_off, 0, t, s));
        float* da1 = dout;      // dout is dead from here on (dsum carries the shortcut's gradient): reuse it
        CK(conv_gemm(MODE_DGRAD, t->convs[k.conv2], B, dz, t->P + t->convs[k.conv2].w_off, da1, 0, t, s));
        // a1 = relu(bn1(z1))
        CK(bn_backward(t, t->bns[k.bn1], k.z1, da1, k.a1, Mo, dz, nullptr, s));
        CK(conv_gemm(MODE_WGRAD, t->convs[k.conv1], B, dz, xin, t->G + t->convs[k.conv1].w_off, 0, t, s));
        float* dxin = dout;     // da1 is dead: the block input's gradient goes where dout was
        CK(conv_gemm(MODE_DGRAD, t->convs[k.conv1], B, dz, t->P + t->convs[k.conv1].w_off, dxin, 0, t, s));
        if (k.convd >= 0) {
            CK(bn_backward(t, t->bns[k.bnd], k.zd, dsum, nullptr, Mo, dz, nullptr, s));
            CK(conv_gemm(MODE_WGRAD, t->convs[k.convd], B, dz, xin, t->G + t->convs[k.convd].w_off, 0, t, s));
            CK(conv_gemm(MODE_DGRAD, t->convs[k.convd], B, dz, t->P + t->convs[k.convd].w_off, dxin, 1, t, s));
        } else {
            k_cls_add<<<blocks_for(static_cast<size_t>(Mo) * k.c_out), 256, 0, s>>>(dxin, dsum, static_cast<size_t>(Mo) * k.c_out);
            CK(cudaGetLastError());
        }
    }
    // stem: p0 = maxpool(a0), a0 = relu(bn(z0)), z0 = conv1(x0); the input needs no gradient
    k_cls_maxpool_bwd<<<blocks_for(static_cast<size_t>(B) * 32 * 32 * 64), 256, 0, s>>>(t->gA, t->arg0, B, 32, 32, 64, t->gC);
    CK(cudaGetLastError());
    CK(bn_backward(t, t->bns[0], t->z0, t->gC, t->a0, B * 32 * 32, t->gB, nullptr, s));
    CK(conv_gemm(MODE_WGRAD, t->convs[0], B, t->gB, t->x0, t->G + t->convs[0].w_off, 0, t, s));
    return 0;
}

}  // namespace

void cvb_cls_trainer_free(cvb_cls_trainer* t) {
    if (!t) return;
    for (void* p : t->allocs) cudaFree(p);
    delete t;
}

extern "C" {

int cvb_cls_train_default_config(cvb_cls_train_config* cfg) {
    if (!cfg) return -1;
    cfg->batch = 64;
    cfg->beta1 = 0.9f;
    cfg->beta2 = 0.999f;
    cfg->eps = 1e-8f;
    cfg->weight_decay = 0.f;
    cfg->bn_momentum = 0.1f;
    cfg->bn_eps = 1e-5f;
    return 0;
}

int cvb_cls_train_create(cvb_ctx* ctx, const cvb_tensor* sd, int n, const cvb_cls_train_config* cfg_in) {
    if (!ctx || !sd) return -1;
    CVB_ON_DEVICE(ctx);
    if (ctx->cls_trainer) { cvb_cls_trainer_free(ctx->cls_trainer); ctx->cls_trainer = nullptr; }
    cvb_cls_trainer* t = new cvb_cls_trainer();
    if (cfg_in) t->cfg = *cfg_in; else cvb_cls_train_default_config(&t->cfg);
    if (t->cfg.batch <= 0 || t->cfg.batch > 4096) { delete t; return fail(ctx, -5, "classifier trainer: batch %d out of range", t->cfg.batch); }
    const int B = t->B = t->cfg.batch;
    auto bail = [&](int rc) { cvb_cls_trainer_free(t); return rc; };
    // ---- topology of timm resnet18(in_chans=1, num_classes=13) on 64x64 squares
    add_conv(t, "conv1", 1, 64, 7, 2, 64);
    add_bn(t, "bn1", 64);
    int cin = 64, h = 16;
    const int widths[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l)
        for (int b = 0; b < 2; ++b) {
            const std::string pre = "layer" + std::to_string(l + 1) + "." + std::to_string(b);
            const int cout = widths[l], stride = (l > 0 && b == 0) ? 2 : 1;
            ClsBlock k;
            k.h_in = h; k.h_out = h / stride; k.c_in = cin; k.c_out = cout;
            k.conv1 = add_conv(t, pre + ".conv1", cin, cout, 3, stride, h);
            k.bn1 = add_bn(t, pre + ".bn1", cout);
            k.conv2 = add_conv(t, pre + ".conv2", cout, cout, 3, 1, h / stride);
            k.bn2 = add_bn(t, pre + ".bn2", cout);
            k.convd = k.bnd = -1;
            if (stride != 1 || cin != cout) {
                k.convd = add_conv(t, pre + ".downsample.0", cin, cout, 1, stride, h);
                k.bnd = add_bn(t, pre + ".downsample.1", cout);
            }
            k.z1 = k.a1 = k.z2 = k.zd = k.sd = k.out = nullptr;
            t->blocks.push_back(k);
            cin = cout;
            h /= stride;
        }
    t->fc_w_off = t->n_params; t->n_params += static_cast<size_t>(kClasses) * 512;
    t->fc_b_off = t->n_params; t->n_params += kClasses;
    // ---- parameters from the state dict (conv weights [co][ci][r][q] -> [co][r][q][ci])
    std::vector<float> hp(t->n_params, 0.f);
    for (const ClsConv& c : t->convs) {
        const cvb_tensor* w = find_t(sd, n, c.name + ".weight");
        if (!w || numel_t(w) != 1LL * c.cout * c.cin * c.ks * c.ks) return bail(fail(ctx, -4, "classifier trainer: missing/bad '%s.weight'", c.name.c_str()));
        for (int co = 0; co < c.cout; ++co)
            for (int ci = 0; ci < c.cin; ++ci)
                for (int r = 0; r < c.ks; ++r)
                    for (int q = 0; q < c.ks; ++q)
                        hp[c.w_off + ((static_cast<size_t>(co) * c.ks + r) * c.ks + q) * c.cin + ci] =
                            w->data[((static_cast<size_t>(co) * c.cin + ci) * c.ks + r) * c.ks + q];
    }
    for (ClsBn& b : t->bns) {
        const cvb_tensor* g = find_t(sd, n, b.name + ".weight");
        const cvb_tensor* be = find_t(sd, n, b.name + ".bias");
        const cvb_tensor* rm = find_t(sd, n, b.name + ".running_mean");
        const cvb_tensor* rv = find_t(sd, n, b.name + ".running_var");
        if (!g || !be || !rm || !rv || numel_t(g) != b.c || numel_t(be) != b.c || numel_t(rm) != b.c || numel_t(rv) != b.c)
            return bail(fail(ctx, -4, "classifier trainer: missing/bad BatchNorm '%s'", b.name.c_str()));
        memcpy(&hp[b.g_off], g->data, b.c * sizeof(float));
        memcpy(&hp[b.b_off], be->data, b.c * sizeof(float));
        if (talloc(ctx, t, &b.run_mean, b.c) || talloc(ctx, t, &b.run_var, b.c) || talloc(ctx, t, &b.mean, b.c) || talloc(ctx, t, &b.invstd, b.c) ||
            talloc(ctx, t, &b.sum_dy, b.c) || talloc(ctx, t, &b.sum_dyx, b.c))
            return bail(-3);
        if (cudaMemcpy(b.run_mean, rm->data, b.c * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(b.run_var, rv->data, b.c * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(ctx, -2, "classifier trainer: upload failed"));
        const cvb_tensor* nb = find_t(sd, n, b.name + ".num_batches_tracked");
        b.batches_tracked = nb && numel_t(nb) == 1 ? static_cast<int64_t>(nb->data[0]) : 0;
    }
    {
        const cvb_tensor* w = find_t(sd, n, "fc.weight");
        const cvb_tensor* b = find_t(sd, n, "fc.bias");
        if (!w || !b || numel_t(w) != kClasses * 512 || numel_t(b) != kClasses) return bail(fail(ctx, -4, "classifier trainer: missing/bad 'fc'"));
        memcpy(&hp[t->fc_w_off], w->data, kClasses * 512 * sizeof(float));
        memcpy(&hp[t->fc_b_off], b->data, kClasses * sizeof(float));
    }
    if (talloc(ctx, t, &t->P, t->n_params) || talloc(ctx, t, &t->G, t->n_params) || talloc(ctx, t, &t->M1, t->n_params) || talloc(ctx, t, &t->M2, t->n_params))
        return bail(-3);
    if (cudaMemcpy(t->P, hp.data(), t->n_params * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemset(t->G, 0, t->n_params * sizeof(float)) != cudaSuccess || cudaMemset(t->M1, 0, t->n_params * sizeof(float)) != cudaSuccess ||
        cudaMemset(t->M2, 0, t->n_params * sizeof(float)) != cudaSuccess)
        return bail(fail(ctx, -2, "classifier trainer: upload failed"));
    // ---- activations
    const size_t sB = static_cast<size_t>(B);
    int rc = 0;
    rc |= talloc(ctx, t, &t->x0, sB * 4096);
    rc |= talloc(ctx, t, &t->z0, sB * 32 * 32 * 64);
    rc |= talloc(ctx, t, &t->a0, sB * 32 * 32 * 64);
    rc |= talloc(ctx, t, &t->p0, sB * 16 * 16 * 64);
    rc |= talloc(ctx, t, &t->arg0, sB * 16 * 16 * 64);
    for (ClsBlock& k : t->blocks) {
        const size_t no = sB * k.h_out * k.h_out * k.c_out;
        rc |= talloc(ctx, t, &k.z1, no);
        rc |= talloc(ctx, t, &k.a1, no);
        rc |= talloc(ctx, t, &k.z2, no);
        rc |= talloc(ctx, t, &k.out, no);
        if (k.convd >= 0) { rc |= talloc(ctx, t, &k.zd, no); rc |= talloc(ctx, t, &k.sd, no); }
    }
    rc |= talloc(ctx, t, &t->pooled, sB * 512);
    rc |= talloc(ctx, t, &t->logits, sB * kClasses);
    rc |= talloc(ctx, t, &t->probs, sB * kClasses);
    rc |= talloc(ctx, t, &t->dlogits, sB * kClasses);
    rc |= talloc(ctx, t, &t->loss_each, sB);
    rc |= talloc(ctx, t, &t->correct_each, sB);
    rc |= talloc(ctx, t, &t->d_loss, 1);
    rc |= talloc(ctx, t, &t->d_correct, 1);
    rc |= talloc(ctx, t, &t->part, static_cast<size_t>(kColSlices) * 2 * 512);
    t->wpart_floats = static_cast<size_t>(16) * 512 * 9 * 512;   // 16 splits of the largest weight tensor
    rc |= talloc(ctx, t, &t->wpart, t->wpart_floats);
    const size_t gmax = sB * 32 * 32 * 64;   // the largest activation (stem conv output)
    rc |= talloc(ctx, t, &t->gA, gmax);
    rc |= talloc(ctx, t, &t->gB, gmax);
    rc |= talloc(ctx, t, &t->gC, gmax);
    if (rc) return bail(-3);
    ctx->cls_trainer = t;
    return 0;
}

int cvb_cls_train_forward(cvb_ctx* ctx, const float* x, const int32_t* target, int training, float* loss, int32_t* correct, float* logits,
                          void* stream) {
    if (!ctx || !ctx->cls_trainer) return ctx ? fail(ctx, -7, "classifier trainer not created (call cvb_cls_train_create)") : -1;
    CVB_ON_DEVICE(ctx);
    cvb_cls_trainer* t = ctx->cls_trainer;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (cls_forward(ctx, t, x, target, training, s)) return -2;
    t->have_forward = training != 0;
    if (loss) CK(cudaMemcpyAsync(loss, t->d_loss, sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (correct) CK(cudaMemcpyAsync(correct, t->d_correct, sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (logits) CK(cudaMemcpyAsync(logits, t->logits, static_cast<size_t>(t->B) * kClasses * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

int cvb_cls_train_forward_backward(cvb_ctx* ctx, const float* x, const int32_t* target, float* loss, int32_t* correct, void* stream) {
    if (!target) return ctx ? fail(ctx, -1, "classifier trainer: targets required") : -1;
    int rc = cvb_cls_train_forward(ctx, x, target, 1, loss, correct, nullptr, stream);
    if (rc) return rc;
    CVB_ON_DEVICE(ctx);
    return cls_backward(ctx, ctx->cls_trainer, target, static_cast<cudaStream_t>(stream)) ? -2 : 0;
}

int cvb_cls_train_grads(cvb_ctx* ctx, float** grads, int64_t* count) {
    if (!ctx || !ctx->cls_trainer) return ctx ? fail(ctx, -7, "classifier trainer not created") : -1;
    if (grads) *grads = ctx->cls_trainer->G;
    if (count) *count = static_cast<int64_t>(ctx->cls_trainer->n_params);
    return 0;
}

int cvb_cls_train_optimizer_step(cvb_ctx* ctx, float lr, float grad_scale, void* stream) {
    if (!ctx || !ctx->cls_trainer) return ctx ? fail(ctx, -7, "classifier trainer not created") : -1;
    CVB_ON_DEVICE(ctx);
    cvb_cls_trainer* t = ctx->cls_trainer;
    t->step++;
    const double b1 = t->cfg.beta1, b2 = t->cfg.beta2;
    const float bias1 = static_cast<float>(1.0 - pow(b1, static_cast<double>(t->step)));
    const float bias2_sqrt = static_cast<float>(sqrt(1.0 - pow(b2, static_cast<double>(t->step))));
    k_cls_adam<<<blocks_for(t->n_params), 256, 0, static_cast<cudaStream_t>(stream)>>>(t->P, t->G, t->M1, t->M2, t->n_params, lr, t->cfg.beta1, t->cfg.beta2,
                                                                                      t->cfg.eps, t->cfg.weight_decay, grad_scale, bias1, bias2_sqrt);
    CK(cudaGetLastError());
    return 0;
}

int cvb_cls_train_step(cvb_ctx* ctx, const float* x, const int32_t* target, float lr, float* loss, int32_t* correct, void* stream) {
    int rc = cvb_cls_train_forward_backward(ctx, x, target, loss, correct, stream);
    if (rc) return rc;
    return cvb_cls_train_optimizer_step(ctx, lr, 1.f, stream);
}

int cvb_cls_train_export(cvb_ctx* ctx, int what, const cvb_tensor* out, int n) {
    if (!ctx || !ctx->cls_trainer || !out) return ctx ? fail(ctx, -7, "classifier trainer not created") : -1;
    CVB_ON_DEVICE(ctx);
    cvb_cls_trainer* t = ctx->cls_trainer;
    if (what < 0 || what > 3) return fail(ctx, -1, "classifier trainer export: what must be 0 (parameters), 1 (gradients), 2 / 3 (Adam moments)");
    CK(cudaDeviceSynchronize());
    const float* src = what == 0 ? t->P : (what == 1 ? t->G : (what == 2 ? t->M1 : t->M2));
    std::vector<float> h(t->n_params);
    CK(cudaMemcpy(h.data(), src, t->n_params * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const std::string name = out[i].name;
        float* dst = const_cast<float*>(out[i].data);
        bool done = false;
        for (const ClsConv& c : t->convs)
            if (name == c.name + ".weight") {
                if (numel_t(&out[i]) != 1LL * c.cout * c.cin * c.ks * c.ks) return fail(ctx, -4, "export: bad shape for '%s'", name.c_str());
                for (int co = 0; co < c.cout; ++co)
                    for (int ci = 0; ci < c.cin; ++ci)
                        for (int r = 0; r < c.ks; ++r)
                            for (int q = 0; q < c.ks; ++q)
                                dst[((static_cast<size_t>(co) * c.cin + ci) * c.ks + r) * c.ks + q] =
                                    h[c.w_off + ((static_cast<size_t>(co) * c.ks + r) * c.ks + q) * c.cin + ci];
                done = true;
            }
        for (const ClsBn& b : t->bns) {
            if (done) break;
            const bool is_g = name == b.name + ".weight", is_b = name == b.name + ".bias";
            const bool is_rm = name == b.name + ".running_mean", is_rv = name == b.name + ".running_var";
            const bool is_nb = name == b.name + ".num_batches_tracked";
            if (!(is_g || is_b || is_rm || is_rv || is_nb)) continue;
            if (is_nb) {
                if (numel_t(&out[i]) != 1) return fail(ctx, -4, "export: bad shape for '%s'", name.c_str());
                dst[0] = what == 0 ? static_cast<float>(b.batches_tracked) : 0.f;
            } else {
                if (numel_t(&out[i]) != b.c) return fail(ctx, -4, "export: bad shape for '%s'", name.c_str());
                if (is_g) memcpy(dst, &h[b.g_off], b.c * sizeof(float));
                else if (is_b) memcpy(dst, &h[b.b_off], b.c * sizeof(float));
                else if (what == 0) CK(cudaMemcpy(dst, is_rm ? b.run_mean : b.run_var, b.c * sizeof(float), cudaMemcpyDeviceToHost));
                else memset(dst, 0, b.c * sizeof(float));
            }
            done = true;
        }
        if (!done && name == "fc.weight") {
            if (numel_t(&out[i]) != kClasses * 512) return fail(ctx, -4, "export: bad shape for 'fc.weight'");
            memcpy(dst, &h[t->fc_w_off], kClasses * 512 * sizeof(float));
            done = true;
        }
        if (!done && name == "fc.bias") {
            if (numel_t(&out[i]) != kClasses) return fail(ctx, -4, "export: bad shape for 'fc.bias'");
            memcpy(dst, &h[t->fc_b_off], kClasses * sizeof(float));
            done = true;
        }
        if (!done) return fail(ctx, -4, "export: '%s' is not a tensor of the classifier", name.c_str());
    }
    return 0;
}

int64_t cvb_cls_train_steps(const cvb_ctx* ctx) { return ctx && ctx->cls_trainer ? ctx->cls_trainer->step : 0; }

}  // extern "C"
