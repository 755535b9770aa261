// CUDA-core kernels of the UNet training step (reference scripts/train/train_unet.py:293-323,
// chessvision/pytorch_unet/utils/dice_score.py:5-30, torch.optim.RMSprop, nn.BatchNorm2d in training mode).
// All of them are bandwidth-bound elementwise / reduction passes over fp16 NHWC activations; the contractions run on
// tensor cores (conv_tc.cu forward + data gradient, wgrad_tc.cu weight gradient).
//
// Loss scaling: activation gradients are fp16, so everything downstream of the loss carries the factor `S`
// (TrainKernelScale.s); reductions into the fp32 gradient buffer multiply by 1/S.
#include "train_kernels.h"

#include <math.h>

namespace cvb {

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ weight preparation
// fp32 master -> fp16 copy, same (packed) layout
__global__ void k_cast_f16(const float* __restrict__ src, __half* __restrict__ dst, long long n4) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n4) return;
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

// Data-gradient weights: out[ci][t'][co] = in[co*s_co + tap(t')*s_t + ci], tap(t') = flip ? T-1-t' : t'.
// grid (Cin/32, Cout/32, T), block (32, 8).
__global__ void k_transpose_w(const float* __restrict__ in, __half* __restrict__ out, int Cout, int Cin, int T, long long s_co,
                              long long s_t, int flip) {
    __shared__ float tile[32][33];
    const int t_out = blockIdx.z, t_in = flip ? T - 1 - t_out : t_out;
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8)
        tile[r][threadIdx.x] = in[(co0 + r) * s_co + t_in * s_t + ci0 + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8)
        out[(static_cast<long long>(ci0 + r) * T + t_out) * Cout + co0 + threadIdx.x] = __float2half_rn(tile[threadIdx.x][r]);
}

// ------------------------------------------------------------------------------------------------ first layer (Cin = 3)
// z0[n][h][w][co] = sum_{r,s,c} x[n][c][h+r-1][w+s-1] * w[co][(r*3+s)*3 + c]        x fp32 NCHW, z0 fp16 NHWC (64 ch)
// One thread per output pixel; weights [64][27] in shared memory (transposed to [27][64] for broadcast reads).
__global__ void __launch_bounds__(128) k_stem_fwd(const float* __restrict__ x, const float* __restrict__ w, __half* __restrict__ z,
                                                  int H, int W) {
    __shared__ float sw[27][64];
    for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i % 27][i / 27] = w[i];
    __syncthreads();
    const int n = blockIdx.z, h = blockIdx.y, wx = blockIdx.x * blockDim.x + threadIdx.x;
    if (wx >= W) return;
    float in[27];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int hh = h + r - 1, ww = wx + s - 1;
            const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                in[(r * 3 + s) * 3 + c] = ok ? __ldg(x + ((static_cast<size_t>(n) * 3 + c) * H + hh) * W + ww) : 0.f;
        }
    __half* dst = z + ((static_cast<size_t>(n) * H + h) * W + wx) * 64;
#pragma unroll 1
    for (int cb = 0; cb < 64; cb += 8) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 27; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(in[k], sw[k][cb + j], acc[j]);
        __half2 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
        *reinterpret_cast<uint4*>(dst + cb) = *reinterpret_cast<const uint4*>(o);
    }
}

// dW0[co][k] += (1/S) * sum_p dz[p][co] * x[p + tap(k)][c(k)].   One block per image row; 256 threads own the 1728
// outputs (o = tid + 256*j); dz row and the three input rows are staged in shared memory.
__global__ void __launch_bounds__(256) k_stem_wgrad(const float* __restrict__ x, const __half* __restrict__ dz, float* __restrict__ gw,
                                                    int H, int W, float inv_s) {
    extern __shared__ float sm[];
    float* s_x = sm;                     // [3 rows][3 ch][W + 2]
    float* s_dz = sm + 9 * (W + 2);      // [W][64]
    const int n = blockIdx.y, h = blockIdx.x;
    for (int i = threadIdx.x; i < 9 * (W + 2); i += 256) {
        const int col = i % (W + 2), rc = i / (W + 2), r = rc / 3, c = rc % 3;
        const int hh = h + r - 1, ww = col - 1;
        s_x[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(x + ((static_cast<size_t>(n) * 3 + c) * H + hh) * W + ww) : 0.f;
    }
    const __half2* drow = reinterpret_cast<const __half2*>(dz + (static_cast<size_t>(n) * H + h) * W * 64);
    for (int i = threadIdx.x; i < W * 32; i += 256) {
        const float2 f = __half22float2(drow[i]);
        s_dz[2 * i] = f.x;
        s_dz[2 * i + 1] = f.y;
    }
    __syncthreads();
    float acc[7] = {0, 0, 0, 0, 0, 0, 0};
    int co[7], xoff[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int o = threadIdx.x + 256 * j;          // o = k*64 + co  (k = tap*3 + c), valid while o < 1728
        const int k = o >> 6, tap = k / 3, c = k % 3, r = tap / 3, s = tap % 3;
        co[j] = o & 63;
        xoff[j] = o < 1728 ? (r * 3 + c) * (W + 2) + s : 0;
    }
    for (int p = 0; p < W; ++p) {
#pragma unroll
        for (int j = 0; j < 7; ++j) acc[j] = fmaf(s_dz[p * 64 + co[j]], s_x[xoff[j] + p], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const int o = threadIdx.x + 256 * j;
        if (o < 1728) atomicAdd(gw + (o & 63) * 27 + (o >> 6), acc[j] * inv_s);
    }
}

// ------------------------------------------------------------------------------------------------ batch-norm, forward
// Column reductions over a [rows][C] fp16 matrix (C a multiple of 8, at most 1024), shared by the forward statistics, the
// backward reduction and the bias gradients.  A thread owns 8 consecutive channels (one 16-byte load per row) and walks the
// rows of its block four at a time, so four independent loads are in flight per thread; `256 / (C / 8)` rows are covered
// per step by one block.  Partial sums meet in shared memory and leave as one double atomicAdd per channel and block.
// NV = accumulators per channel (2: sum + second moment, 1: sum only).
template <int NV>
__device__ __forceinline__ void col_reduce_tail(float (&acc)[8 * NV], double* __restrict__ out0, double* __restrict__ out1, float* red,
                                                int groups, int rpar, int g, int rl) {
    // red[rl][g][8 * NV]
#pragma unroll
    for (int j = 0; j < 8 * NV; ++j) red[(rl * groups + g) * (8 * NV) + j] = acc[j];
    __syncthreads();
    for (int idx = threadIdx.x; idx < groups * 8 * NV; idx += 256) {
        double t = 0.0;
        for (int k = 0; k < rpar; ++k) t += static_cast<double>(red[k * groups * 8 * NV + idx]);
        const int gg = idx / (8 * NV), j = idx - gg * (8 * NV);
        const int c = gg * 8 + (j % 8);
        atomicAdd((j < 8 ? out0 : out1) + c, t);
    }
}

// Per-channel sum and sum of squares of z [rows][C] (fp16, dense).  Each block covers `rows_per_block` rows.
__global__ void __launch_bounds__(256) k_bn_stats(const __half* __restrict__ z, double* __restrict__ sums, long long rows, int C,
                                                  int rows_per_block) {
    extern __shared__ float red[];
    const int groups = C >> 3, per = groups < 256 ? groups : 256, rpar = 256 / per;
    const int gl = threadIdx.x % per, rl = threadIdx.x / per;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    for (int gb = 0; gb < groups; gb += per) {
        const int g = gb + gl;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        const __half* col = z + 8 * g;
        for (long long r = r0 + rl; r < r1; r += 4LL * rpar) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long rr = r + static_cast<long long>(u) * rpar;
                v[u] = rr < r1 ? *reinterpret_cast<const uint4*>(col + rr * C) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                    acc[8 + 2 * j] = fmaf(f.x, f.x, acc[8 + 2 * j]);
                    acc[8 + 2 * j + 1] = fmaf(f.y, f.y, acc[8 + 2 * j + 1]);
                }
            }
        }
        col_reduce_tail<2>(acc, sums + 8 * gb, sums + C + 8 * gb, red, per, rpar, gl, rl);
        __syncthreads();
    }
}

// sums -> batch mean / biased variance -> scale = gamma*rstd, shift = beta - mean*scale; running statistics updated
// as nn.BatchNorm2d does (momentum 0.1, unbiased variance).
__global__ void k_bn_finalize(const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean, float* __restrict__ rstd,
                              float* __restrict__ run_mean, float* __restrict__ run_var, int C, double rows, float eps, float momentum) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / rows;
    double var = sums[C + c] / rows - m * m;
    if (var < 0.0) var = 0.0;
    const double rs = 1.0 / sqrt(var + static_cast<double>(eps));
    const double sc = static_cast<double>(gamma[c]) * rs;
    scale[c] = static_cast<float>(sc);
    shift[c] = static_cast<float>(static_cast<double>(beta[c]) - m * sc);
    mean[c] = static_cast<float>(m);
    rstd[c] = static_cast<float>(rs);
    const double unbiased = rows > 1.0 ? var * rows / (rows - 1.0) : var;
    run_mean[c] = static_cast<float>((1.0 - momentum) * run_mean[c] + momentum * m);
    run_var[c] = static_cast<float>((1.0 - momentum) * run_var[c] + momentum * unbiased);
}

// y = relu(z*scale + shift): z dense [rows][C] -> y at channel offset/stride.  One thread per 8 channels.
__global__ void __launch_bounds__(256) k_bn_apply_relu(const __half* __restrict__ z, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, __half* __restrict__ y, long long rows, int C,
                                                       int y_stride, int y_off) {
    const int groups = C >> 3;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= rows * groups) return;
    const long long r = i / groups;
    const int c = static_cast<int>(i - r * groups) << 3;
    const uint4 v = *reinterpret_cast<const uint4*>(z + r * C + c);
    const __half2* h = reinterpret_cast<const __half2*>(&v);
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(shift + c)), b1 = __ldg(reinterpret_cast<const float4*>(shift + c + 4));
    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    __half2 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        o[j] = __floats2half2_rn(fmaxf(fmaf(f.x, sc[2 * j], sh[2 * j]), 0.f), fmaxf(fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]), 0.f));
    }
    *reinterpret_cast<uint4*>(y + r * y_stride + y_off + c) = *reinterpret_cast<const uint4*>(o);
}

// ------------------------------------------------------------------------------------------------ batch-norm + ReLU, backward
// g = dy * [z*scale+shift > 0];  bsums[c] += g, bsums[C+c] += g * xhat,  xhat = (z - mean) * rstd
__global__ void __launch_bounds__(256) k_bn_bwd_reduce(const __half* __restrict__ dy, const __half* __restrict__ z,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       double* __restrict__ bsums, long long rows, int C, int rows_per_block) {
    extern __shared__ float red[];
    const int groups = C >> 3, per = groups < 256 ? groups : 256, rpar = 256 / per;
    const int gl = threadIdx.x % per, rl = threadIdx.x / per;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    for (int gb = 0; gb < groups; gb += per) {
        const int g = gb + gl, c0 = 8 * g;
        float sc[8], sh[8], mu[8], rs[8], acc[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = __ldg(scale + c0 + j);
            sh[j] = __ldg(shift + c0 + j);
            mu[j] = __ldg(mean + c0 + j);
            rs[j] = __ldg(rstd + c0 + j);
            acc[j] = acc[8 + j] = 0.f;
        }
        for (long long r = r0 + rl; r < r1; r += 2LL * rpar) {
            uint4 zv[2], dv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long rr = r + static_cast<long long>(u) * rpar;
                const bool in = rr < r1;
                zv[u] = in ? *reinterpret_cast<const uint4*>(z + rr * C + c0) : make_uint4(0, 0, 0, 0);
                dv[u] = in ? *reinterpret_cast<const uint4*>(dy + rr * C + c0) : make_uint4(0, 0, 0, 0);   // dy = 0: no contribution
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const __half2* zh = reinterpret_cast<const __half2*>(&zv[u]);
                const __half2* dh = reinterpret_cast<const __half2*>(&dv[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 zf = __half22float2(zh[j]), df = __half22float2(dh[j]);
                    const float g0 = fmaf(zf.x, sc[2 * j], sh[2 * j]) > 0.f ? df.x : 0.f;
                    const float g1 = fmaf(zf.y, sc[2 * j + 1], sh[2 * j + 1]) > 0.f ? df.y : 0.f;
                    acc[2 * j] += g0;
                    acc[2 * j + 1] += g1;
                    acc[8 + 2 * j] = fmaf(g0, (zf.x - mu[2 * j]) * rs[2 * j], acc[8 + 2 * j]);
                    acc[8 + 2 * j + 1] = fmaf(g1, (zf.y - mu[2 * j + 1]) * rs[2 * j + 1], acc[8 + 2 * j + 1]);
                }
            }
        }
        col_reduce_tail<2>(acc, bsums + 8 * gb, bsums + C + 8 * gb, red, per, rpar, gl, rl);
        __syncthreads();
    }
}

// dz = scale * (g - mean_rows(g) - xhat * mean_rows(g*xhat))  (fp16, dense);  block 0 also emits d_gamma, d_beta.
// Same thread layout as the reductions: a thread keeps the per-channel constants of its 8 channels in registers and walks
// the rows of its block.
__global__ void __launch_bounds__(256) k_bn_bwd_apply(const __half* __restrict__ dy, const __half* __restrict__ z,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      const double* __restrict__ bsums, __half* __restrict__ dz, float* __restrict__ g_gamma,
                                                      float* __restrict__ g_beta, long long rows, int C, float inv_s, int rows_per_block) {
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += 256) {
            g_beta[c] += static_cast<float>(bsums[c] * inv_s);
            g_gamma[c] += static_cast<float>(bsums[C + c] * inv_s);
        }
    }
    const int groups = C >> 3, per = groups < 256 ? groups : 256, rpar = 256 / per;
    const int gl = threadIdx.x % per, rl = threadIdx.x / per;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    const float inv_rows = 1.0f / static_cast<float>(rows);
    for (int gb = 0; gb < groups; gb += per) {
        const int c0 = 8 * (gb + gl);
        float sc[8], sh[8], mu[8], rs[8], mg[8], mgx[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = __ldg(scale + c0 + j);
            sh[j] = __ldg(shift + c0 + j);
            mu[j] = __ldg(mean + c0 + j);
            rs[j] = __ldg(rstd + c0 + j);
            mg[j] = static_cast<float>(bsums[c0 + j]) * inv_rows;
            mgx[j] = static_cast<float>(bsums[C + c0 + j]) * inv_rows;
        }
        for (long long r = r0 + rl; r < r1; r += 2LL * rpar) {
            uint4 zv[2], dv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long rr = r + static_cast<long long>(u) * rpar;
                const bool in = rr < r1;
                zv[u] = in ? *reinterpret_cast<const uint4*>(z + rr * C + c0) : make_uint4(0, 0, 0, 0);
                dv[u] = in ? *reinterpret_cast<const uint4*>(dy + rr * C + c0) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const long long rr = r + static_cast<long long>(u) * rpar;
                if (rr >= r1) break;
                const __half2* zh = reinterpret_cast<const __half2*>(&zv[u]);
                const __half2* dh = reinterpret_cast<const __half2*>(&dv[u]);
                __half2 oh[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 zf = __half22float2(zh[j]), df = __half22float2(dh[j]);
                    const float g0 = fmaf(zf.x, sc[2 * j], sh[2 * j]) > 0.f ? df.x : 0.f;
                    const float g1 = fmaf(zf.y, sc[2 * j + 1], sh[2 * j + 1]) > 0.f ? df.y : 0.f;
                    const float x0 = (zf.x - mu[2 * j]) * rs[2 * j], x1 = (zf.y - mu[2 * j + 1]) * rs[2 * j + 1];
                    oh[j] = __floats2half2_rn(sc[2 * j] * (g0 - mg[2 * j] - x0 * mgx[2 * j]), sc[2 * j + 1] * (g1 - mg[2 * j + 1] - x1 * mgx[2 * j + 1]));
                }
                *reinterpret_cast<uint4*>(dz + rr * C + c0) = *reinterpret_cast<const uint4*>(oh);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ max-pool backward + skip add
// dy[n][h][w][c] = dskip[n][h][w][c] + (argmax of the 2x2 window == (h,w) ? dpool[n][h/2][w/2][c] : 0)
// y, dskip: channels [0,C) of buffers with `y_stride` / `ds_stride` channels per pixel; dpool, dy dense.  The first
// maximum in scan order wins (PyTorch).  One thread per window and 8 channels.
__global__ void __launch_bounds__(256) k_pool_bwd_add(const __half* __restrict__ y, int y_stride, const __half* __restrict__ dskip,
                                                      int ds_stride, const __half* __restrict__ dpool, __half* __restrict__ dy, int N,
                                                      int H, int W, int C) {
    const int groups = C >> 3, Hp = H >> 1, Wp = W >> 1;
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<long long>(N) * Hp * Wp * groups) return;
    const int c = static_cast<int>(i % groups) << 3;
    long long t = i / groups;
    const int wp = static_cast<int>(t % Wp);
    t /= Wp;
    const int hp = static_cast<int>(t % Hp), n = static_cast<int>(t / Hp);
    const uint4 dpv = *reinterpret_cast<const uint4*>(dpool + ((static_cast<long long>(n) * Hp + hp) * Wp + wp) * C + c);
    const __half* dp = reinterpret_cast<const __half*>(&dpv);
    uint4 yv[4];
    const __half* yh[4];
    long long pix[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        pix[q] = (static_cast<long long>(n) * H + 2 * hp + (q >> 1)) * W + 2 * wp + (q & 1);
        yv[q] = *reinterpret_cast<const uint4*>(y + pix[q] * y_stride + c);
        yh[q] = reinterpret_cast<const __half*>(&yv[q]);
    }
    int best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float m = __half2float(yh[0][j]);
        int b = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            const float v = __half2float(yh[q][j]);
            if (v > m) { m = v; b = q; }
        }
        best[j] = b;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint4 sv = *reinterpret_cast<const uint4*>(dskip + pix[q] * ds_stride + c);
        const __half* sh = reinterpret_cast<const __half*>(&sv);
        __half o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = __float2half_rn(__half2float(sh[j]) + (best[j] == q ? __half2float(dp[j]) : 0.f));
        *reinterpret_cast<uint4*>(dy + pix[q] * C + c) = *reinterpret_cast<const uint4*>(o);
    }
}

// Same without a skip branch (not used by the UNet, kept for the unit test of the pooling rule): dy = scatter(dpool).
// ------------------------------------------------------------------------------------------------ column sums (bias gradients)
// out[c] += scale * sum_rows src[row*stride + off + c]
__global__ void __launch_bounds__(256) k_colsum(const __half* __restrict__ src, long long rows, int C, int stride, int off,
                                                float* __restrict__ out, float scale, int rows_per_block) {
    extern __shared__ float red[];
    const int groups = C >> 3, per = groups < 256 ? groups : 256, rpar = 256 / per;
    const int gl = threadIdx.x % per, rl = threadIdx.x / per;
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_block;
    const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
    for (int gb = 0; gb < groups; gb += per) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const __half* col = src + off + 8 * (gb + gl);
        for (long long r = r0 + rl; r < r1; r += 4LL * rpar) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long rr = r + static_cast<long long>(u) * rpar;
                v[u] = rr < r1 ? *reinterpret_cast<const uint4*>(col + rr * stride) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const __half2* h = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(h[j]);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(rl * per + gl) * 8 + j] = acc[j];
        __syncthreads();
        for (int idx = threadIdx.x; idx < per * 8; idx += 256) {
            float t = 0.f;
            for (int k = 0; k < rpar; ++k) t += red[k * per * 8 + idx];
            atomicAdd(out + 8 * gb + idx, t * scale);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ 1x1 head + loss
// logits[p] = sum_c y[p][c]*w[c] + b    (y fp16 [P][64] dense)
__global__ void __launch_bounds__(256) k_outc_fwd(const __half* __restrict__ y, const float* __restrict__ w, const float* __restrict__ b,
                                                  float* __restrict__ logits, long long P) {
    __shared__ float sw[64];
    if (threadIdx.x < 64) sw[threadIdx.x] = w[threadIdx.x];
    __syncthreads();
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (p >= P) return;
    float acc = b[0];
    const uint4* src = reinterpret_cast<const uint4*>(y + p * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint4 v = src[k];
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h[j]);
            acc = fmaf(f.x, sw[8 * k + 2 * j], acc);
            acc = fmaf(f.y, sw[8 * k + 2 * j + 1], acc);
        }
    }
    logits[p] = acc;
}

// Per sample n: lsums[4n+0] = sum sigmoid(x), [4n+1] = sum t, [4n+2] = sum sigmoid(x)*t;  lsums[4N] = sum BCE-with-logits.
// grid (blocks_per_sample, N)
__global__ void __launch_bounds__(256) k_loss_reduce(const float* __restrict__ logits, const float* __restrict__ target,
                                                     double* __restrict__ lsums, int HW, int N) {
    const int n = blockIdx.y;
    const float* x = logits + static_cast<size_t>(n) * HW;
    const float* t = target + static_cast<size_t>(n) * HW;
    float s = 0.f, st = 0.f, sx = 0.f, bce = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
        const float xv = x[i], tv = t[i];
        const float sg = 1.0f / (1.0f + expf(-xv));
        s += sg; st += tv; sx += sg * tv;
        bce += fmaxf(xv, 0.f) - xv * tv + log1pf(expf(-fabsf(xv)));
    }
    __shared__ float red[8][4];
    s = warp_sum(s); st = warp_sum(st); sx = warp_sum(sx); bce = warp_sum(bce);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp][0] = s; red[warp][1] = st; red[warp][2] = sx; red[warp][3] = bce; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0, d = 0;
        for (int k = 0; k < 8; ++k) { a += red[k][0]; b += red[k][1]; c += red[k][2]; d += red[k][3]; }
        atomicAdd(lsums + 4 * n, a);
        atomicAdd(lsums + 4 * n + 1, b);
        atomicAdd(lsums + 4 * n + 2, c);
        atomicAdd(lsums + 4 * N, d);
    }
}

// loss = BCE_mean + 1 - mean_n (2*I_n + eps) / (S_n + eps)              (train_unet.py:311-317, dice_score.py:5-30)
// dL/dx = (sig - t)/(N*HW) - (1/N) * (2*t*(S_n+eps) - (2*I_n+eps)) / (S_n+eps)^2 * sig*(1-sig)
// Emits dy[p][c] = S * dL/dx[p] * w[c] (fp16), accumulates d_w[c] += dL/dx[p]*y[p][c], d_b += dL/dx[p]; block (0,0)
// writes the loss value.  grid (blocks_per_sample, N), 256 threads, one pixel per thread per iteration.
__global__ void __launch_bounds__(256) k_loss_grad_outc(const float* __restrict__ logits, const float* __restrict__ target,
                                                        const double* __restrict__ lsums, const __half* __restrict__ y,
                                                        const float* __restrict__ w, __half* __restrict__ dy, float* __restrict__ g_w,
                                                        float* __restrict__ g_b, float* __restrict__ loss_out, int HW, int N, float S) {
    __shared__ float sw[64];
    __shared__ float s_gw[64];
    __shared__ float s_gb;
    if (threadIdx.x < 64) { sw[threadIdx.x] = w[threadIdx.x]; s_gw[threadIdx.x] = 0.f; }
    if (threadIdx.x == 0) s_gb = 0.f;
    const double eps = 1e-6;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        double dice = 0.0;
        for (int n = 0; n < N; ++n) {
            const double inter = 2.0 * lsums[4 * n + 2];
            double sets = lsums[4 * n] + lsums[4 * n + 1];
            if (sets == 0.0) sets = inter;
            dice += (inter + eps) / (sets + eps);
        }
        loss_out[0] = static_cast<float>(lsums[4 * N] / (static_cast<double>(N) * HW) + 1.0 - dice / N);
    }
    __syncthreads();
    const int n = blockIdx.y;
    const double Sn = lsums[4 * n] + lsums[4 * n + 1] + eps, In = 2.0 * lsums[4 * n + 2] + eps;
    const float a = static_cast<float>(2.0 / Sn / N), b = static_cast<float>(In / (Sn * Sn) / N);   // dDice/dsig / N = a*t - b
    const float inv_cnt = 1.0f / (static_cast<float>(N) * HW);
    float gw[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) gw[c] = 0.f;
    float gb = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < HW; i += gridDim.x * 256) {
        const size_t p = static_cast<size_t>(n) * HW + i;
        const float xv = logits[p], tv = target[p];
        const float sg = 1.0f / (1.0f + expf(-xv));
        const float g = (sg - tv) * inv_cnt - (a * tv - b) * sg * (1.0f - sg);
        gb += g;
        const float gs = g * S;
        const uint4* ysrc = reinterpret_cast<const uint4*>(y + p * 64);
        uint4* ddst = reinterpret_cast<uint4*>(dy + p * 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint4 v = ysrc[k];
            const __half2* h = reinterpret_cast<const __half2*>(&v);
            __half2 o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                gw[8 * k + 2 * j] = fmaf(g, f.x, gw[8 * k + 2 * j]);
                gw[8 * k + 2 * j + 1] = fmaf(g, f.y, gw[8 * k + 2 * j + 1]);
                o[j] = __floats2half2_rn(gs * sw[8 * k + 2 * j], gs * sw[8 * k + 2 * j + 1]);
            }
            ddst[k] = *reinterpret_cast<const uint4*>(o);
        }
    }
#pragma unroll
    for (int c = 0; c < 64; ++c) {
        const float v = warp_sum(gw[c]);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_gw[c], v);
    }
    gb = warp_sum(gb);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_gb, gb);
    __syncthreads();
    if (threadIdx.x < 64) atomicAdd(g_w + threadIdx.x, s_gw[threadIdx.x]);
    if (threadIdx.x == 0) atomicAdd(g_b, s_gb);
}

// ------------------------------------------------------------------------------------------------ optimizer
// norm[0] += sum (g*gscale)^2 ; norm[1] = 1 if any gradient is not finite
__global__ void __launch_bounds__(256) k_grad_sqnorm(const float* __restrict__ g, long long n, float gscale, double* __restrict__ norm) {
    double acc = 0.0;
    bool bad = false;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const float v = g[i] * gscale;
        bad |= !isfinite(v);
        acc += static_cast<double>(v) * v;
    }
    acc = warp_sum_d(acc);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    if (bad) norm[1] = 1.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(norm, t);
    }
}

// torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.RMSprop(lr, alpha, eps, weight_decay, momentum)
// (train_unet.py:236-242,321-323).  A step whose gradients are not finite is skipped (what GradScaler.step does).
__global__ void __launch_bounds__(256) k_rmsprop(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ sq,
                                                 float* __restrict__ buf, long long n, const double* __restrict__ norm, float gscale,
                                                 float max_norm, float lr, float alpha, float eps, float wd, float momentum) {
    if (norm[1] != 0.0) return;
    const float total = static_cast<float>(sqrt(norm[0]));
    float coef = max_norm / (total + 1e-6f);
    coef = coef > 1.0f ? 1.0f : coef;
    const float k = gscale * coef;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const float w = p[i];
        float gr = g[i] * k;
        gr = fmaf(wd, w, gr);
        const float s = alpha * sq[i] + (1.0f - alpha) * gr * gr;
        sq[i] = s;
        const float avg = sqrtf(s) + eps;
        float b = gr / avg;
        if (momentum > 0.f) {
            b = fmaf(momentum, buf[i], b);
            buf[i] = b;
        }
        p[i] = w - lr * b;
    }
}

inline unsigned int blocks_for(long long n, int per) { return static_cast<unsigned int>((n + per - 1) / per); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------- launchers
cudaError_t launch_cast_f16(const float* src, __half* dst, long long n, cudaStream_t s) {
    const long long n4 = n / 4;   // buffers are padded to multiples of 4
    if (n4 > 0) k_cast_f16<<<blocks_for(n4, 256), 256, 0, s>>>(src, dst, n4);
    return cudaGetLastError();
}
cudaError_t launch_transpose_w(const float* in, __half* out, int Cout, int Cin, int T, long long s_co, long long s_t, int flip,
                               cudaStream_t s) {
    dim3 grid(Cin / 32, Cout / 32, T), block(32, 8);
    k_transpose_w<<<grid, block, 0, s>>>(in, out, Cout, Cin, T, s_co, s_t, flip);
    return cudaGetLastError();
}
cudaError_t launch_stem_fwd(const float* x, const float* w, __half* z, int N, int H, int W, cudaStream_t s) {
    dim3 grid((W + 127) / 128, H, N);
    k_stem_fwd<<<grid, 128, 0, s>>>(x, w, z, H, W);
    return cudaGetLastError();
}
cudaError_t launch_stem_wgrad(const float* x, const __half* dz, float* gw, int N, int H, int W, float inv_s, cudaStream_t s) {
    dim3 grid(H, N);
    const size_t smem = (9 * (W + 2) + W * 64) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_stem_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_stem_wgrad<<<grid, 256, smem, s>>>(x, dz, gw, H, W, inv_s);
    return cudaGetLastError();
}
// rows per block of the column-reduction kernels: about four blocks per SM when the matrix is tall enough, never fewer rows
// than one block covers in a single step of its row loop
inline int reduce_rows_per_block(long long rows, int C, int unroll) {
    const int groups = C >> 3, per = groups < 256 ? groups : 256, rpar = 256 / per;
    long long rpb = (rows + 591) / 592;
    const long long step = static_cast<long long>(unroll) * rpar;
    rpb = ((rpb + step - 1) / step) * step;
    return static_cast<int>(rpb < step ? step : rpb);
}
cudaError_t launch_bn_stats(const __half* z, double* sums, long long rows, int C, cudaStream_t s) {
    const int rpb = reduce_rows_per_block(rows, C, 4);
    k_bn_stats<<<blocks_for(rows, rpb), 256, 256 * 16 * sizeof(float), s>>>(z, sums, rows, C, rpb);
    return cudaGetLastError();
}
cudaError_t launch_bn_finalize(const double* sums, const float* gamma, const float* beta, float* scale, float* shift, float* mean,
                               float* rstd, float* run_mean, float* run_var, int C, long long rows, float eps, float momentum,
                               cudaStream_t s) {
    k_bn_finalize<<<(C + 127) / 128, 128, 0, s>>>(sums, gamma, beta, scale, shift, mean, rstd, run_mean, run_var, C,
                                                  static_cast<double>(rows), eps, momentum);
    return cudaGetLastError();
}
cudaError_t launch_bn_apply_relu(const __half* z, const float* scale, const float* shift, __half* y, long long rows, int C, int y_stride,
                                 int y_off, cudaStream_t s) {
    k_bn_apply_relu<<<blocks_for(rows * (C / 8), 256), 256, 0, s>>>(z, scale, shift, y, rows, C, y_stride, y_off);
    return cudaGetLastError();
}
cudaError_t launch_bn_bwd(const __half* dy, const __half* z, const float* scale, const float* shift, const float* mean, const float* rstd,
                          double* bsums, __half* dz, float* g_gamma, float* g_beta, long long rows, int C, float inv_s, cudaStream_t s) {
    const int rpb = reduce_rows_per_block(rows, C, 2);
    k_bn_bwd_reduce<<<blocks_for(rows, rpb), 256, 256 * 16 * sizeof(float), s>>>(dy, z, scale, shift, mean, rstd, bsums, rows, C, rpb);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_bn_bwd_apply<<<blocks_for(rows, rpb), 256, 0, s>>>(dy, z, scale, shift, mean, rstd, bsums, dz, g_gamma, g_beta, rows, C, inv_s, rpb);
    return cudaGetLastError();
}
cudaError_t launch_pool_bwd_add(const __half* y, int y_stride, const __half* dskip, int ds_stride, const __half* dpool, __half* dy, int N,
                                int H, int W, int C, cudaStream_t s) {
    const long long work = static_cast<long long>(N) * (H / 2) * (W / 2) * (C / 8);
    k_pool_bwd_add<<<blocks_for(work, 256), 256, 0, s>>>(y, y_stride, dskip, ds_stride, dpool, dy, N, H, W, C);
    return cudaGetLastError();
}
cudaError_t launch_colsum(const __half* src, long long rows, int C, int stride, int off, float* out, float scale, cudaStream_t s) {
    const int rpb = reduce_rows_per_block(rows, C, 4);
    k_colsum<<<blocks_for(rows, rpb), 256, 256 * 8 * sizeof(float), s>>>(src, rows, C, stride, off, out, scale, rpb);
    return cudaGetLastError();
}
cudaError_t launch_outc_fwd(const __half* y, const float* w, const float* b, float* logits, long long P, cudaStream_t s) {
    k_outc_fwd<<<blocks_for(P, 256), 256, 0, s>>>(y, w, b, logits, P);
    return cudaGetLastError();
}
cudaError_t launch_loss(const float* logits, const float* target, double* lsums, const __half* y, const float* w, __half* dy, float* g_w,
                        float* g_b, float* loss_out, int N, int HW, float S, cudaStream_t s) {
    dim3 grid(64, N);
    k_loss_reduce<<<grid, 256, 0, s>>>(logits, target, lsums, HW, N);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_loss_grad_outc<<<grid, 256, 0, s>>>(logits, target, lsums, y, w, dy, g_w, g_b, loss_out, HW, N, S);
    return cudaGetLastError();
}
cudaError_t launch_optimizer(float* p, const float* g, float* sq, float* buf, long long n, double* norm, float gscale, float max_norm,
                             float lr, float alpha, float eps, float wd, float momentum, int sm_count, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(norm, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return e;
    k_grad_sqnorm<<<sm_count * 8, 256, 0, s>>>(g, n, gscale, norm);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_rmsprop<<<sm_count * 8, 256, 0, s>>>(p, g, sq, buf, n, norm, gscale, max_norm, lr, alpha, eps, wd, momentum);
    return cudaGetLastError();
}

}  // namespace cvb
