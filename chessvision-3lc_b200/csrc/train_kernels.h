// Launchers of the CUDA-core kernels of the UNet training step (internal to the library; see train_kernels.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

cudaError_t launch_cast_f16(const float* src, __half* dst, long long n, cudaStream_t s);
// out[ci][t'][co] (fp16) = in[co*s_co + tap*s_t + ci] (fp32), tap = flip ? T-1-t' : t'.  Cin, Cout multiples of 32.
cudaError_t launch_transpose_w(const float* in, __half* out, int Cout, int Cin, int T, long long s_co, long long s_t, int flip,
                               cudaStream_t s);
// inc.double_conv.0 (Cin = 3): x fp32 NCHW [N,3,H,W], w fp32 [64][27] (k = (r*3+s)*3 + c) -> z fp16 NHWC [N,H,W,64]
cudaError_t launch_stem_fwd(const float* x, const float* w, __half* z, int N, int H, int W, cudaStream_t s);
cudaError_t launch_stem_wgrad(const float* x, const __half* dz, float* gw, int N, int H, int W, float inv_s, cudaStream_t s);
// BatchNorm2d (training mode) over z fp16 [rows][C]; sums = double [2C], zeroed by the caller
cudaError_t launch_bn_stats(const __half* z, double* sums, long long rows, int C, cudaStream_t s);
cudaError_t launch_bn_finalize(const double* sums, const float* gamma, const float* beta, float* scale, float* shift, float* mean,
                               float* rstd, float* run_mean, float* run_var, int C, long long rows, float eps, float momentum,
                               cudaStream_t s);
cudaError_t launch_bn_apply_relu(const __half* z, const float* scale, const float* shift, __half* y, long long rows, int C, int y_stride,
                                 int y_off, cudaStream_t s);
// ReLU + BatchNorm backward: dy, z dense [rows][C] -> dz dense; g_gamma/g_beta += (1/S) * sums; bsums = double [2C], zeroed
cudaError_t launch_bn_bwd(const __half* dy, const __half* z, const float* scale, const float* shift, const float* mean, const float* rstd,
                          double* bsums, __half* dz, float* g_gamma, float* g_beta, long long rows, int C, float inv_s, cudaStream_t s);
// MaxPool2d(2) backward added to the skip-connection gradient (see kernel comment)
cudaError_t launch_pool_bwd_add(const __half* y, int y_stride, const __half* dskip, int ds_stride, const __half* dpool, __half* dy, int N,
                                int H, int W, int C, cudaStream_t s);
cudaError_t launch_colsum(const __half* src, long long rows, int C, int stride, int off, float* out, float scale, cudaStream_t s);
cudaError_t launch_outc_fwd(const __half* y, const float* w, const float* b, float* logits, long long P, cudaStream_t s);
// BCEWithLogits + Dice loss and its gradient through the 1x1 head; lsums = double [4N+1], zeroed by the caller
cudaError_t launch_loss(const float* logits, const float* target, double* lsums, const __half* y, const float* w, __half* dy, float* g_w,
                        float* g_b, float* loss_out, int N, int HW, float S, cudaStream_t s);
// clip_grad_norm_ + RMSprop over the flat parameter / gradient buffers; norm = double [2] scratch
cudaError_t launch_optimizer(float* p, const float* g, float* sq, float* buf, long long n, double* norm, float gscale, float max_norm,
                             float lr, float alpha, float eps, float wd, float momentum, int sm_count, cudaStream_t s);

}  // namespace cvb
