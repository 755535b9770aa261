// C ABI of the B200-native image->FEN path (include/chessvision_b200.h): context, weight folding/packing, the static
// launch plan of both networks, the per-chunk pipeline and the host<->device streaming entry point.
#include "../../include/chessvision_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cmath>
#include <algorithm>
#include <string>
#include <vector>

#include "conv_tc.h"
#include "kernels.h"

using namespace cvb;

#include "ctx.h"

namespace {

const cvb_tensor* find(const cvb_tensor* sd, int n, const std::string& name) {
    for (int i = 0; i < n; ++i)
        if (name == sd[i].name) return &sd[i];
    return nullptr;
}

int64_t numel(const cvb_tensor* t) {
    int64_t k = 1;
    for (int i = 0; i < t->ndim; ++i) k *= t->shape[i];
    return k;
}

struct Bn {
    std::vector<float> scale, shift;
};

// BatchNorm2d(eval, eps 1e-5): y = x*scale + shift
int load_bn(cvb_ctx* ctx, const cvb_tensor* sd, int n, const std::string& prefix, int C, Bn& bn) {
    const cvb_tensor* g = find(sd, n, prefix + ".weight");
    const cvb_tensor* b = find(sd, n, prefix + ".bias");
    const cvb_tensor* m = find(sd, n, prefix + ".running_mean");
    const cvb_tensor* v = find(sd, n, prefix + ".running_var");
    if (!g || !b || !m || !v) return fail(ctx, -4, "state_dict lacks BatchNorm tensors '%s.*'", prefix.c_str());
    if (numel(g) != C || numel(b) != C || numel(m) != C || numel(v) != C) return fail(ctx, -4, "bad BatchNorm shape at '%s'", prefix.c_str());
    bn.scale.resize(C);
    bn.shift.resize(C);
    for (int c = 0; c < C; ++c) {
        const double s = static_cast<double>(g->data[c]) / sqrt(static_cast<double>(v->data[c]) + 1e-5);
        bn.scale[c] = static_cast<float>(s);
        bn.shift[c] = static_cast<float>(static_cast<double>(b->data[c]) - static_cast<double>(m->data[c]) * s);
    }
    return 0;
}

int upload_conv(cvb_ctx* ctx, const std::vector<__half>& w, const std::vector<float>& bias, int rows, int K, ConvWeights& out) {
    if (dalloc(ctx, &out.w, w.size())) return -3;
    if (dalloc(ctx, &out.bias, bias.size())) return -3;
    CK(cudaMemcpy(out.w, w.data(), w.size() * sizeof(__half), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(out.bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
    out.rows = rows;
    out.K = K;
    return 0;
}

// Conv2d weight [Cout][Cin][k][k] (+BN) -> fp16 [Cout][(r*k+s)*Cin + ci], bias fp32 [Cout]
int pack_conv(cvb_ctx* ctx, const cvb_tensor* sd, int n, const std::string& conv, const std::string& bn_prefix, int Cout,
              int Cin, int k, ConvWeights& out) {
    const cvb_tensor* w = find(sd, n, conv + ".weight");
    if (!w) return fail(ctx, -4, "state_dict lacks '%s.weight'", conv.c_str());
    if (numel(w) != 1LL * Cout * Cin * k * k) return fail(ctx, -4, "bad shape for '%s.weight'", conv.c_str());
    Bn bn;
    if (load_bn(ctx, sd, n, bn_prefix, Cout, bn)) return -4;
    const int K = k * k * Cin;
    std::vector<__half> pw(static_cast<size_t>(Cout) * K);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int r = 0; r < k; ++r)
                for (int s = 0; s < k; ++s) {
                    const float v = w->data[((static_cast<size_t>(co) * Cin + ci) * k + r) * k + s] * bn.scale[co];
                    pw[static_cast<size_t>(co) * K + (r * k + s) * Cin + ci] = __float2half_rn(v);
                }
    return upload_conv(ctx, pw, bn.shift, Cout, K, out);
}

// ConvTranspose2d weight [Cin][Cout][2][2] + bias[Cout] -> fp16 [(dy*2+dx)*Cout + co][Cin], bias fp32 [4*Cout]
int pack_convt(cvb_ctx* ctx, const cvb_tensor* sd, int n, const std::string& name, int Cin, int Cout, ConvWeights& out) {
    const cvb_tensor* w = find(sd, n, name + ".weight");
    const cvb_tensor* b = find(sd, n, name + ".bias");
    if (!w || !b) return fail(ctx, -4, "state_dict lacks '%s.{weight,bias}'", name.c_str());
    if (numel(w) != 4LL * Cin * Cout || numel(b) != Cout) return fail(ctx, -4, "bad shape for '%s'", name.c_str());
    std::vector<__half> pw(static_cast<size_t>(4) * Cout * Cin);
    std::vector<float> pb(static_cast<size_t>(4) * Cout);
    for (int ci = 0; ci < Cin; ++ci)
        for (int co = 0; co < Cout; ++co)
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx)
                    pw[(static_cast<size_t>(dy * 2 + dx) * Cout + co) * Cin + ci] =
                        __float2half_rn(w->data[((static_cast<size_t>(ci) * Cout + co) * 2 + dy) * 2 + dx]);
    for (int q = 0; q < 4; ++q)
        for (int co = 0; co < Cout; ++co) pb[static_cast<size_t>(q) * Cout + co] = b->data[co];
    return upload_conv(ctx, pw, pb, 4 * Cout, Cin, out);
}

// tiny-K first layers stay fp32 on CUDA cores: [k*k*Cin][64] with BN folded
int pack_stem(cvb_ctx* ctx, const cvb_tensor* sd, int n, const std::string& conv, const std::string& bn_prefix, int Cin, int k,
              float** d_w, float** d_b) {
    const cvb_tensor* w = find(sd, n, conv + ".weight");
    if (!w || numel(w) != 64LL * Cin * k * k) return fail(ctx, -4, "missing/bad '%s.weight'", conv.c_str());
    Bn bn;
    if (load_bn(ctx, sd, n, bn_prefix, 64, bn)) return -4;
    std::vector<float> pw(static_cast<size_t>(k) * k * Cin * 64);
    for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int r = 0; r < k; ++r)
                for (int s = 0; s < k; ++s)
                    pw[(static_cast<size_t>(r * k + s) * Cin + ci) * 64 + co] =
                        w->data[((static_cast<size_t>(co) * Cin + ci) * k + r) * k + s] * bn.scale[co];
    if (dalloc(ctx, d_w, pw.size()) || dalloc(ctx, d_b, 64)) return -3;
    CK(cudaMemcpy(*d_w, pw.data(), pw.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(*d_b, bn.shift.data(), 64 * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

// The same tiny-K layers as one tcgen05 B tile: fp16 [64 cout][64 k] in the 128-byte-swizzled K-major layout the MMA
// descriptor expects (16-byte chunk j of row r at r*128 + ((j ^ (r & 7)) << 4)); k = ky*ky_stride + kx*kx_stride + ci,
// every other k is zero.  BN scale and the 256/255 factor (inputs enter as v/256, the reference divides by 255) folded.
int pack_stem_tc(cvb_ctx* ctx, const cvb_tensor* sd, int n, const std::string& conv, const std::string& bn_prefix, int Cin, int k,
                 int ky_stride, int kx_stride, int bias_k, void** d_w, int neg_k = -1) {
    const cvb_tensor* w = find(sd, n, conv + ".weight");
    if (!w || numel(w) != 64LL * Cin * k * k) return fail(ctx, -4, "missing/bad '%s.weight'", conv.c_str());
    Bn bn;
    if (load_bn(ctx, sd, n, bn_prefix, 64, bn)) return -4;
    // one or two swizzled [64 co][64 K] tiles: K index kk lives in tile kk / 64
    const int tiles = (bias_k + 1 >= 64 || neg_k >= 64) ? 2 : 1;
    std::vector<__half> img(static_cast<size_t>(tiles) * 64 * 64, __float2half_rn(0.f));
    auto at = [&](int co, int kk) -> __half& {
        const int t = kk >> 6, c = kk & 63;
        return img[static_cast<size_t>(t) * 4096 + (static_cast<size_t>(co) * 128 + (((c >> 3) ^ (co & 7)) << 4) + (c & 7) * 2) / 2];
    };
    for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int r = 0; r < k; ++r)
                for (int s = 0; s < k; ++s) {
                    const double v = static_cast<double>(w->data[((static_cast<size_t>(co) * Cin + ci) * k + r) * k + s]) *
                                     static_cast<double>(bn.scale[co]) * (256.0 / 255.0);
                    at(co, r * ky_stride + s * kx_stride + ci) = __float2half_rn(static_cast<float>(v));
                }
    // The folded BatchNorm shift rides along as two more K columns (the kernels keep 1.0 in columns bias_k, bias_k + 1 of every
    // im2col row): bias = hi + lo in fp16, exact to 2^-22 of its magnitude, accumulated in fp32 by the MMA itself.
    for (int co = 0; co < 64; ++co) {
        const float b = bn.shift[co];
        const __half hi = __float2half_rn(b);
        at(co, bias_k) = hi;
        at(co, bias_k + 1) = __float2half_rn(b - __half2float(hi));
        // K column neg_k: -30000 for every channel.  An im2col row with 1.0 there (a max-pool window position outside the conv
        // image) loses every maximum it takes part in (k_resnet_stem_tc).
        if (neg_k >= 0) at(co, neg_k) = __float2half_rn(-30000.f);
    }
    __half* d = nullptr;
    if (dalloc(ctx, &d, img.size())) return -3;
    CK(cudaMemcpy(d, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
    *d_w = d;
    return 0;
}

// Thin wrappers over conv_build / conv_set_store (conv_tc.cu) that turn their error codes into context messages.
int build_conv(cvb_ctx* ctx, ConvLaunch& L, const __half* in, int Nmax, int Hin, int Win, int in_c_stride, int in_c_off, int Cin,
               const ConvWeights& cw, int ksize, int stride, int epilogue) {
    const int rc = conv_build(L, in, Nmax, Hin, Win, in_c_stride, in_c_off, Cin, cw.w, cw.bias, cw.rows, cw.K, ksize, stride, epilogue,
                              ctx->use_vr);
    if (rc == -5) return fail(ctx, -5, "conv: Cin=%d Cout=%d K=%d ksize=%d on %dx%d not supported", Cin, cw.rows, cw.K, ksize, Hin, Win);
    if (rc) return fail(ctx, -6, "cuTensorMapEncodeTiled (conv operands) failed: %d", rc);
    return 0;
}

int set_store(cvb_ctx* ctx, ConvLaunch& L, __half* out, int out_c_stride, int out_c_off, int relu, const __half* res, int res_c_stride) {
    const int rc = conv_set_store(L, out, out_c_stride, out_c_off, relu, res, res_c_stride);
    if (rc) return fail(ctx, -6, "cuTensorMapEncodeTiled (output view) failed: %d", rc);
    return 0;
}

// ------------------------------------------------------------------------------------------------------ profiling
struct StageTimer {
    cvb_ctx* c;
    cudaStream_t s;
    bool on;
    StageTimer(cvb_ctx* ctx, int stage, cudaStream_t st) : c(ctx), s(st), on(false) {
        if (!c->profile || c->pev_used + 2 > static_cast<int>(c->pev.size())) return;
        on = true;
        c->pev_stage[c->pev_used] = stage;
        cudaEventRecord(c->pev[c->pev_used], s);
    }
    ~StageTimer() {
        if (!on) return;
        cudaEventRecord(c->pev[c->pev_used + 1], s);
        c->pev_used += 2;
    }
};

// ------------------------------------------------------------------------------------------------------ stages
int run_conv(cvb_ctx* ctx, ConvLaunch& L, int n, cudaStream_t s) {
    CK(conv_launch(L, n, ctx->sm_count, s));
    ctx->launches++;
    return 0;
}

// OpenCV's computeResizeAreaTab (imgproc/resize.cpp) for one axis, in its own double arithmetic: per destination index the
// source cells it covers (a partial cell on the left, whole cells, a partial cell on the right) with float32 weights.
void area_table(int ssize, int dsize, double scale, std::vector<int>& ofs, std::vector<int>& si, std::vector<float>& alpha) {
    ofs.assign(1, 0);
    si.clear();
    alpha.clear();
    for (int d = 0; d < dsize; ++d) {
        const double f1 = d * scale, f2 = f1 + scale;
        const double cell = std::min(scale, ssize - f1);
        int s1 = static_cast<int>(std::ceil(f1)), s2 = static_cast<int>(std::floor(f2));
        s2 = std::min(s2, ssize - 1);
        s1 = std::min(s1, s2);
        if (s1 - f1 > 1e-3) {
            si.push_back(s1 - 1);
            alpha.push_back(static_cast<float>((s1 - f1) / cell));
        }
        for (int sx = s1; sx < s2; ++sx) {
            si.push_back(sx);
            alpha.push_back(static_cast<float>(1.0 / cell));
        }
        if (f2 - s2 > 1e-3) {
            si.push_back(s2);
            alpha.push_back(static_cast<float>(std::min(std::min(f2 - s2, 1.), cell) / cell));
        }
        ofs.push_back(static_cast<int>(si.size()));
    }
}

// OpenCV's coefficient loop of the bilinear resizer in "area mode" (imgproc/resize.cpp, cv::resize with INTER_AREA when an
// axis is enlarged): (source index, round(2048*(1-f)), round(2048*f)) per destination index; returns xmax.
int linear_area_table(int ssize, int dsize, std::vector<int>& tab) {
    const double inv = static_cast<double>(dsize) / ssize, scale = 1.0 / inv;
    int xmax = dsize;
    tab.resize(3 * static_cast<size_t>(dsize));
    for (int d = 0; d < dsize; ++d) {
        int s = static_cast<int>(std::floor(d * scale));
        float f = static_cast<float>((d + 1) - (s + 1) * inv);
        f = f <= 0 ? 0.f : f - std::floor(f);
        if (s < 0) { f = 0.f; s = 0; }
        if (s + 1 >= ssize) {
            xmax = std::min(xmax, d);
            if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        }
        tab[3 * d] = s;
        tab[3 * d + 1] = static_cast<int>(std::lrintf((1.f - f) * 2048.f));   // saturate_cast<short>: round half to even
        tab[3 * d + 2] = static_cast<int>(std::lrintf(f * 2048.f));
    }
    return xmax;
}

// (Re)place a device copy of a host table; the previous copy, if any, is released.
template <class T>
int upload_table(T** dst, const std::vector<T>& v) {
    if (*dst) cudaFree(*dst);
    *dst = nullptr;
    void* q = nullptr;
    if (cudaMalloc(&q, v.size() * sizeof(T)) != cudaSuccess) return -3;
    if (cudaMemcpy(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(q);
        return -2;
    }
    *dst = static_cast<T*>(q);
    return 0;
}

// Tables + staging images for inputs of H x W != 512 x 512, (re)built when the size changes.
int prepare_size(cvb_ctx* ctx, int H, int W) {
    if (H < 1 || W < 1) return fail(ctx, -5, "input images must not be empty");
    if (ctx->gs_H == H && ctx->gs_W == W) return 0;
    CK(cudaDeviceSynchronize());   // a previous size's tables may still be in use
    if (!ctx->gs_small) {
        if (cudaMalloc(&ctx->gs_small, static_cast<size_t>(ctx->max_batch) * 256 * 256 * 3) != cudaSuccess ||
            cudaMalloc(&ctx->gs_big, static_cast<size_t>(ctx->max_batch) * 786432) != cudaSuccess)
            return fail(ctx, -3, "allocating the resize staging images failed");
    }
    ctx->gs_linear = H < 256 || W < 256;   // an enlarged axis: OpenCV switches BOTH axes to its bilinear emulation
    if (ctx->gs_linear) {
        std::vector<int> tx, ty;
        ctx->gs_xmax = linear_area_table(W, 256, tx);
        linear_area_table(H, 256, ty);
        if (upload_table(&ctx->gs_lx, tx) || upload_table(&ctx->gs_ly, ty)) return fail(ctx, -3, "allocating the resize tables failed");
        ctx->gs_H = H;
        ctx->gs_W = W;
        return 0;
    }
    const double fx = static_cast<double>(W) / 256, fy = static_cast<double>(H) / 256;
    const int ix = static_cast<int>(std::lrint(fx)), iy = static_cast<int>(std::lrint(fy));
    const bool fast = std::abs(fx - ix) < 2.220446049250313e-16 && std::abs(fy - iy) < 2.220446049250313e-16;
    std::vector<int> xo, xs, yo, ys;
    std::vector<float> xa, ya;
    area_table(W, 256, fx, xo, xs, xa);
    area_table(H, 256, fy, yo, ys, ya);
    if (upload_table(&ctx->gs_xofs, xo) || upload_table(&ctx->gs_xsi, xs) || upload_table(&ctx->gs_xa, xa) ||
        upload_table(&ctx->gs_yofs, yo) || upload_table(&ctx->gs_ysi, ys) || upload_table(&ctx->gs_ya, ya))
        return fail(ctx, -3, "allocating the INTER_AREA tables failed");
    ctx->gs_int_area = fast ? ix * iy : 0;
    ctx->gs_H = H;
    ctx->gs_W = W;
    return 0;
}

// cv2.resize(img, (256,256), INTER_AREA) (core.py:212) for n <= max_batch images of H x W into ctx->gs_small
int resize_to_256(cvb_ctx* ctx, const uint8_t* img, int n, int H, int W, uint8_t* out, cudaStream_t s) {
    if (prepare_size(ctx, H, W)) return -2;
    if (ctx->gs_linear) CK(launch_resize_linear_area(img, out, n, H, W, 256, 256, ctx->gs_lx, ctx->gs_ly, ctx->gs_xmax, s));
    else CK(launch_resize_area(img, out, n, H, W, 256, 256, ctx->gs_xofs, ctx->gs_xsi, ctx->gs_xa, ctx->gs_yofs, ctx->gs_ysi, ctx->gs_ya,
                               ctx->gs_int_area, s));
    ctx->launches++;
    return 0;
}

int unet_forward(cvb_ctx* ctx, const uint8_t* img, int n, float thr, float* logits, uint8_t* mask, cudaStream_t s);

// extract_board up to the logits for images of any size >= 256 x 256: the 512 x 512 case feeds the stem directly (it fuses the
// 2x reduction); other sizes are reduced by k_resize_area and replicated 2x, which the stem's reduction undoes exactly.
int unet_forward_hw(cvb_ctx* ctx, const uint8_t* img, int n, int H, int W, float thr, float* logits, uint8_t* mask, cudaStream_t s) {
    if (H == 512 && W == 512) return unet_forward(ctx, img, n, thr, logits, mask, s);
    {
        StageTimer t(ctx, 1, s);
        if (resize_to_256(ctx, img, n, H, W, ctx->gs_small, s)) return -2;
        CK(launch_double2x(ctx->gs_small, ctx->gs_big, n, 256, 256, s));
        ctx->launches++;
    }
    return unet_forward(ctx, ctx->gs_big, n, thr, logits, mask, s);
}

int unet_forward(cvb_ctx* ctx, const uint8_t* img, int n, float thr, float* logits, uint8_t* mask, cudaStream_t s) {
    if (!ctx->unet_loaded) return fail(ctx, -7, "UNet weights not loaded (call cvb_load_unet)");
    auto& P = ctx->unet_plan;
    auto aux = [&](cudaError_t e) { ctx->launches++; return e; };
    {
        StageTimer t(ctx, 1, s);
        if (ctx->stem_fp32) CK(aux(launch_unet_stem(img, ctx->stem_w, ctx->stem_b, ctx->t0, n, 256, 256, 64, s)));
        else CK(aux(launch_unet_stem_tc(img, ctx->stem_wsw, ctx->stem_b, &ctx->stem_omap, n, ctx->sm_count, s)));
    }
    { StageTimer t(ctx, 0, s); if (run_conv(ctx, P[0], n, s)) return -2; }                       // inc.3      t0 -> cat0[0:64)
    if (P[0].p.pool_out == nullptr) {   // the row-streaming kernel pools in its epilogue (conv_try_rs), the others do not
        StageTimer t(ctx, 1, s);
        CK(aux(launch_maxpool2(ctx->cat0, ctx->p1, n, 256, 256, 64, 128, s)));
    }
    { StageTimer t(ctx, 0, s); if (run_conv(ctx, P[1], n, s) || run_conv(ctx, P[2], n, s)) return -2; }   // down1
    if (P[2].p.pool_out == nullptr) { StageTimer t(ctx, 1, s); CK(aux(launch_maxpool2(ctx->cat1, ctx->p2, n, 128, 128, 128, 256, s))); }
    { StageTimer t(ctx, 0, s); if (run_conv(ctx, P[3], n, s) || run_conv(ctx, P[4], n, s)) return -2; }   // down2
    if (P[4].p.pool_out == nullptr) { StageTimer t(ctx, 1, s); CK(aux(launch_maxpool2(ctx->cat2, ctx->p3, n, 64, 64, 256, 512, s))); }
    { StageTimer t(ctx, 0, s); if (run_conv(ctx, P[5], n, s) || run_conv(ctx, P[6], n, s)) return -2; }   // down3
    if (P[6].p.pool_out == nullptr) { StageTimer t(ctx, 1, s); CK(aux(launch_maxpool2(ctx->cat3, ctx->p4, n, 32, 32, 512, 1024, s))); }
    {
        StageTimer t(ctx, 0, s);
        for (int i = 7; i <= 19; ++i) {                                                           // down4, up1..up4.conv0
            if (i == 18 && ctx->fuse_up4) continue;                                               // up4.up ran inside P[17]
            if (run_conv(ctx, P[i], n, s)) return -2;
        }
        ConvLaunch& last = P[20];                                                                 // up4.conv3 + outc
        last.p.logits = logits ? logits : ctx->ws_logits;
        last.p.mask = mask ? mask : ctx->ws_mask;
        last.p.thr = thr;
        if (run_conv(ctx, last, n, s)) return -2;
    }
    return 0;
}

int classify(cvb_ctx* ctx, const uint8_t* board, int n, int flip, float* probs, uint8_t* labels, uint8_t* labels_valid, char* fen,
             cudaStream_t s) {
    if (!ctx->resnet_loaded) return fail(ctx, -7, "classifier weights not loaded (call cvb_load_resnet18)");
    {
        StageTimer t(ctx, 4, s);
        if (ctx->stem_fp32) CK(launch_resnet_stem(board, ctx->rstem_w, ctx->rstem_b, ctx->rbuf[0], n, s));
        else CK(launch_resnet_stem_tc(board, ctx->rstem_wsw, ctx->rstem_b, ctx->rbuf[0], n, ctx->sm_count, s));
        ctx->launches++;
    }
    {
        StageTimer t(ctx, 5, s);
        for (auto& L : ctx->res_plan)
            if (run_conv(ctx, L, n * 64, s)) return -2;
    }
    {
        StageTimer t(ctx, 6, s);
        CK(launch_head(ctx->rbuf[10], ctx->fc_w, ctx->fc_b, probs ? probs : ctx->ws_probs, labels ? labels : ctx->ws_labels,
                       labels_valid ? labels_valid : ctx->ws_labels_valid, fen ? fen : ctx->ws_fen, n, flip, s));
        ctx->launches++;
    }
    return 0;
}

// One geometry group (n <= ctx->group boards): both networks run chunk by chunk (max_batch boards, their activation
// workspaces), the latency-bound mask->quad kernel and the warp run ONCE over the whole group so that many boards are
// resident per SM and the slow boards of a launch overlap with the rest.  `o` may hold NULL members (workspaces are
// used); `off` = index of the group's first board inside the caller's output arrays.  in_ready[c] (optional): event the
// network of chunk c has to wait for (host path: the chunk's host->device copy).
int pipeline_group_direct(cvb_ctx* ctx, const uint8_t* img, int n, float thr, int flip, const cvb_outputs& o, size_t off, cudaStream_t s,
                          const cudaEvent_t* in_ready, int H, int W);

constexpr size_t kMaxGraphs = 8;

// A group of at most one chunk is the latency case (ChessVision.process_image: one board, ~50 launches of a few
// microseconds each): its launch sequence is captured once into a CUDA graph -- programmatic-dependent-launch edges
// included -- and replayed with one cudaGraphLaunch as long as everything the launches bake in (pointers, count,
// threshold, orientation) is unchanged.  Larger groups are launch-bound nowhere and are enqueued directly.
int pipeline_group(cvb_ctx* ctx, const uint8_t* img, int n, float thr, int flip, const cvb_outputs& o, size_t off, cudaStream_t s,
                   const cudaEvent_t* in_ready, int H = 512, int W = 512) {
    if (!ctx->use_graph || ctx->profile || n > ctx->max_batch || n <= 0 || H != 512 || W != 512 || !ctx->unet_loaded || !ctx->resnet_loaded)
        return pipeline_group_direct(ctx, img, n, thr, flip, o, off, s, in_ready, H, W);
    if (in_ready) CK(cudaStreamWaitEvent(s, in_ready[0], 0));   // an event recorded outside the capture cannot be waited for inside it
    cvb_ctx::GraphEntry want;
    memset(want.key, 0, sizeof want.key);
    uint32_t thr_bits;
    memcpy(&thr_bits, &thr, 4);
    const void* ptrs[11] = {o.logits, o.mask, o.quad, o.found, o.status, o.board, o.probs, o.labels, o.labels_valid, o.fen, o.squares};
    want.key[0] = reinterpret_cast<uint64_t>(img);
    want.key[1] = (static_cast<uint64_t>(n) << 40) | (static_cast<uint64_t>(flip != 0) << 32) | thr_bits;
    want.key[2] = static_cast<uint64_t>(off);
    want.key[3] = reinterpret_cast<uint64_t>(s);
    for (int i = 0; i < 11; ++i) want.key[4 + i] = reinterpret_cast<uint64_t>(ptrs[i]);
    cvb_ctx::GraphEntry* hit = nullptr;
    for (auto& g : ctx->graphs)
        if (memcmp(g.key, want.key, sizeof want.key) == 0) hit = &g;
    if (!hit) {
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();   // e.g. the legacy default stream cannot be captured: enqueue directly
            return pipeline_group_direct(ctx, img, n, thr, flip, o, off, s, nullptr, H, W);
        }
        const int64_t l0 = ctx->launches;
        const int rc = pipeline_group_direct(ctx, img, n, thr, flip, o, off, s, nullptr, H, W);
        cudaGraph_t graph = nullptr;
        const cudaError_t ec = cudaStreamEndCapture(s, &graph);
        want.launches = ctx->launches - l0;
        ctx->launches = l0;
        if (rc || ec != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc) return rc;
            ctx->use_graph = false;   // capture is not possible here (e.g. a tool that forbids it): enqueue directly from now on
            return pipeline_group_direct(ctx, img, n, thr, flip, o, off, s, nullptr, H, W);
        }
        const cudaError_t ei = cudaGraphInstantiate(&want.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return fail(ctx, -2, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ei));
        if (ctx->graphs.size() >= kMaxGraphs) {   // evict the least recently used graph
            size_t lru = 0;
            for (size_t i = 1; i < ctx->graphs.size(); ++i)
                if (ctx->graphs[i].last_use < ctx->graphs[lru].last_use) lru = i;
            cudaGraphExecDestroy(ctx->graphs[lru].exec);
            ctx->graphs.erase(ctx->graphs.begin() + static_cast<long>(lru));
        }
        ctx->graphs.push_back(want);
        hit = &ctx->graphs.back();
    }
    hit->last_use = ++ctx->graph_clock;
    CK(cudaGraphLaunch(hit->exec, s));
    ctx->launches += hit->launches;
    ctx->graph_replays++;
    return 0;
}

int pipeline_group_direct(cvb_ctx* ctx, const uint8_t* img, int n, float thr, int flip, const cvb_outputs& o, size_t off, cudaStream_t s,
                          const cudaEvent_t* in_ready, int H, int W) {
    const int B = ctx->max_batch;
    const size_t img_bytes = static_cast<size_t>(H) * W * 3;
    float* logits = o.logits ? o.logits + off * 65536 : nullptr;
    uint8_t* mask = o.mask ? o.mask + off * 65536 : ctx->ws_mask;
    int32_t* quad = o.quad ? o.quad + off * 8 : ctx->ws_quad;
    uint8_t* found = o.found ? o.found + off : ctx->ws_found;
    int32_t* status = o.status ? o.status + off : ctx->ws_status;
    uint8_t* board = o.board ? o.board + off * 262144 : ctx->ws_board;
    for (int c = 0, c0 = 0; c0 < n; ++c, c0 += B) {
        const int nb = n - c0 < B ? n - c0 : B;
        if (in_ready) CK(cudaStreamWaitEvent(s, in_ready[c], 0));
        if (unet_forward_hw(ctx, img + static_cast<size_t>(c0) * img_bytes, nb, H, W, thr,
                            logits ? logits + static_cast<size_t>(c0) * 65536 : nullptr, mask + static_cast<size_t>(c0) * 65536, s))
            return -2;
    }
    {
        StageTimer t(ctx, 2, s);
        CK(launch_mask_to_quad(mask, quad, found, status, ctx->ws_ncont, ctx->ws_owner, ctx->ws_quad_big, n, ctx->quad_full_only, s));
        ctx->launches += ctx->quad_full_only ? 2 : 3;
    }
    {
        StageTimer t(ctx, 3, s);
        CK(launch_homography(quad, found, ctx->ws_minv, n, static_cast<float>(H) / 256.0f, 512, 512, s));   // _scale_quadrangle: H scales both axes
        CK(launch_warp_board(img, ctx->ws_minv, found, board, o.squares ? o.squares + off * 262144 : nullptr, n, H, W, s));
        ctx->launches += 2;
    }
    for (int c0 = 0; c0 < n; c0 += B) {
        const int nb = n - c0 < B ? n - c0 : B;
        const size_t q = off + c0;
        if (classify(ctx, board + static_cast<size_t>(c0) * 262144, nb, flip, o.probs ? o.probs + q * 832 : nullptr,
                     o.labels ? o.labels + q * 64 : nullptr, o.labels_valid ? o.labels_valid + q * 64 : nullptr,
                     o.fen ? o.fen + q * 144 : nullptr, s))
            return -2;
    }
    return 0;
}

}  // namespace

// ====================================================================================================================
// exported functions
// ====================================================================================================================
extern "C" {

int cvb_version(void) { return 100; }

const char* cvb_last_error(const cvb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int cvb_max_batch(const cvb_ctx* ctx) { return ctx ? ctx->max_batch : 0; }
int64_t cvb_launch_count(const cvb_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t cvb_graph_replays(const cvb_ctx* ctx) { return ctx ? ctx->graph_replays : 0; }

cvb_ctx* cvb_create(int device, int max_batch) {
    if (max_batch <= 0) return nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return nullptr;
    cvb_ctx* ctx = new cvb_ctx();
    ctx->device = device;
    ctx->max_batch = max_batch;
    auto bail = [&]() -> cvb_ctx* {
        fprintf(stderr, "cvb_create: %s\n", ctx->err.c_str());
        cvb_destroy(ctx);
        return nullptr;
    };
    DeviceGuard guard(device);
    if (!guard.ok) { ctx->err = "cudaSetDevice failed"; return bail(); }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { ctx->err = "cudaGetDeviceProperties failed"; return bail(); }
    if (prop.major != 10) {
        ctx->err = "this library contains sm_100a code only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return bail();
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->use_vr = getenv("CVB_NO_VR") == nullptr;
    if (tmap_init()) { ctx->err = "cuTensorMapEncodeTiled not available from the driver"; return bail(); }
    ctx->stem_fp32 = getenv("CVB_STEM_FP32") != nullptr;
    ctx->quad_full_only = getenv("CVB_QUAD_FULL") != nullptr;
    ctx->use_graph = getenv("CVB_NO_GRAPH") == nullptr;
    if (const char* g = getenv("CVB_GROUP_CHUNKS")) ctx->group_chunks = atoi(g) > 0 ? atoi(g) : 1;
    ctx->group = ctx->group_chunks * max_batch;
    if (conv_configure() != cudaSuccess || configure_resnet_stem() != cudaSuccess || configure_stems_tc() != cudaSuccess ||
        configure_quad() != cudaSuccess || configure_warp() != cudaSuccess) {
        ctx->err = std::string("kernel attribute setup failed: ") + cudaGetErrorString(cudaGetLastError());
        return bail();
    }
    const size_t B = max_batch, G = ctx->group;
    int rc = 0;
    // UNet activations (fp16 NHWC)
    rc |= dalloc(ctx, &ctx->cat0, B * 256 * 256 * 128);
    rc |= dalloc(ctx, &ctx->t0, B * 256 * 256 * 64);
    rc |= dalloc(ctx, &ctx->p1, B * 128 * 128 * 64);
    rc |= dalloc(ctx, &ctx->t1, B * 128 * 128 * 128);
    rc |= dalloc(ctx, &ctx->cat1, B * 128 * 128 * 256);
    rc |= dalloc(ctx, &ctx->p2, B * 64 * 64 * 128);
    rc |= dalloc(ctx, &ctx->t2, B * 64 * 64 * 256);
    rc |= dalloc(ctx, &ctx->cat2, B * 64 * 64 * 512);
    rc |= dalloc(ctx, &ctx->p3, B * 32 * 32 * 256);
    rc |= dalloc(ctx, &ctx->t3, B * 32 * 32 * 512);
    rc |= dalloc(ctx, &ctx->cat3, B * 32 * 32 * 1024);
    rc |= dalloc(ctx, &ctx->p4, B * 16 * 16 * 512);
    rc |= dalloc(ctx, &ctx->t4, B * 16 * 16 * 1024);
    rc |= dalloc(ctx, &ctx->x5, B * 16 * 16 * 1024);
    rc |= dalloc(ctx, &ctx->u1, B * 32 * 32 * 512);
    rc |= dalloc(ctx, &ctx->u2, B * 64 * 64 * 256);
    rc |= dalloc(ctx, &ctx->u3, B * 128 * 128 * 128);
    rc |= dalloc(ctx, &ctx->ws_logits, B * 65536);
    rc |= dalloc(ctx, &ctx->ws_mask, G * 65536);
    // geometry (per group)
    rc |= dalloc(ctx, &ctx->ws_quad, G * 8);
    rc |= dalloc(ctx, &ctx->ws_status, G);
    rc |= dalloc(ctx, &ctx->ws_ncont, G);
    rc |= dalloc(ctx, &ctx->ws_owner, G * (kQuadMaxBorders + 8));
    rc |= dalloc(ctx, &ctx->ws_found, G);
    rc |= dalloc(ctx, &ctx->ws_quad_big, quad_big_scratch_bytes());
    if (!rc && cudaMemset(ctx->ws_quad_big, 0, quad_big_scratch_bytes()) != cudaSuccess) rc = fail(ctx, -2, "clearing the mask->quad scratch failed");
    rc |= dalloc(ctx, &ctx->ws_minv, G * 9);
    rc |= dalloc(ctx, &ctx->ws_board, G * 262144);
    // classifier activations: 3 buffers per level; per square 16x16x64, 8x8x128, 4x4x256, 2x2x512 = 16384..2048 halfs
    for (int lvl = 0; lvl < 4 && !rc; ++lvl)
        for (int k = 0; k < 3 && !rc; ++k) rc |= dalloc(ctx, &ctx->rbuf[lvl * 3 + k], B * 64 * (16384 >> lvl));
    rc |= dalloc(ctx, &ctx->ws_probs, B * 64 * 13);
    rc |= dalloc(ctx, &ctx->ws_labels, B * 64);
    rc |= dalloc(ctx, &ctx->ws_labels_valid, B * 64);
    rc |= dalloc(ctx, &ctx->ws_fen, B * 144);
    if (rc) return bail();
    // host streaming resources
    bool ok = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
        ctx->ev_in[i].assign(ctx->group_chunks, nullptr);
        for (auto& e : ctx->ev_in[i]) ok = ok && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming) == cudaSuccess;
        memset(&ctx->slot_out[i], 0, sizeof(cvb_outputs));
        if (ok) ok = dalloc(ctx, &ctx->slot_img[i], G * 512 * 512 * 3) == 0;
        if (ok) ok = dalloc(ctx, &ctx->slot_out[i].quad, G * 8) == 0 && dalloc(ctx, &ctx->slot_out[i].found, G) == 0 &&
                     dalloc(ctx, &ctx->slot_out[i].status, G) == 0 && dalloc(ctx, &ctx->slot_out[i].probs, G * 832) == 0 &&
                     dalloc(ctx, &ctx->slot_out[i].labels, G * 64) == 0 && dalloc(ctx, &ctx->slot_out[i].labels_valid, G * 64) == 0 &&
                     dalloc(ctx, &ctx->slot_out[i].fen, G * 144) == 0 && dalloc(ctx, &ctx->slot_out[i].logits, G * 65536) == 0 &&
                     dalloc(ctx, &ctx->slot_out[i].mask, G * 65536) == 0 && dalloc(ctx, &ctx->slot_out[i].board, G * 262144) == 0 &&
                     dalloc(ctx, &ctx->slot_out[i].squares, G * 262144) == 0;
    }
    if (!ok) {
        if (ctx->err.empty()) ctx->err = "stream/event/slot setup failed";
        return bail();
    }
    return ctx;
}

void cvb_destroy(cvb_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaDeviceSynchronize();
    for (auto& g : ctx->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : ctx->allocs) cudaFree(p);
    for (cudaEvent_t e : ctx->pev) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        for (cudaEvent_t e : ctx->ev_in[i])
            if (e) cudaEventDestroy(e);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_free[i]) cudaEventDestroy(ctx->ev_free[i]);
    }
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    if (ctx->trainer) cvb_trainer_free(ctx->trainer);
    if (ctx->cls_trainer) cvb_cls_trainer_free(ctx->cls_trainer);
    if (ctx->jpeg) cvb_jpeg_free(ctx->jpeg);
    for (void* q : {static_cast<void*>(ctx->gs_xofs), static_cast<void*>(ctx->gs_xsi), static_cast<void*>(ctx->gs_xa), static_cast<void*>(ctx->gs_yofs),
                    static_cast<void*>(ctx->gs_ysi), static_cast<void*>(ctx->gs_ya), static_cast<void*>(ctx->gs_small), static_cast<void*>(ctx->gs_big),
                    static_cast<void*>(ctx->gs_lx), static_cast<void*>(ctx->gs_ly)})
        if (q) cudaFree(q);
    delete ctx;
}

int cvb_load_unet(cvb_ctx* ctx, const cvb_tensor* sd, int n) {
    if (!ctx || !sd) return -1;
    CVB_ON_DEVICE(ctx);
    if (ctx->unet_loaded) return fail(ctx, -8, "UNet weights already loaded");
    const int B = ctx->max_batch;
    static const int width[5] = {64, 128, 256, 512, 1024};
    if (pack_stem(ctx, sd, n, "inc.double_conv.0", "inc.double_conv.1", 3, 3, &ctx->stem_w, &ctx->stem_b)) return -4;
    if (pack_stem_tc(ctx, sd, n, "inc.double_conv.0", "inc.double_conv.1", 3, 3, 16, 4, 48, &ctx->stem_wsw)) return -4;
    if (tmap_act(&ctx->stem_omap, ctx->t0, 64, 256, 256, B, 64, 256 * 64, 65536LL * 64, 64, 2, 1))
        return fail(ctx, -6, "cuTensorMapEncodeTiled (stem output view) failed");
    ctx->unet_w.resize(21);
    auto& W = ctx->unet_w;
    int wi = 0;
    // 0: inc.3 | 1..8: down{1..4}.{0,3} | then per up block: up, conv0, conv3
    if (pack_conv(ctx, sd, n, "inc.double_conv.3", "inc.double_conv.4", 64, 64, 3, W[wi++])) return -4;
    for (int d = 1; d <= 4; ++d) {
        const std::string pre = "down" + std::to_string(d) + ".maxpool_conv.1.double_conv.";
        if (pack_conv(ctx, sd, n, pre + "0", pre + "1", width[d], width[d - 1], 3, W[wi++])) return -4;
        if (pack_conv(ctx, sd, n, pre + "3", pre + "4", width[d], width[d], 3, W[wi++])) return -4;
    }
    for (int u = 1; u <= 4; ++u) {
        const int cin = width[5 - u], cout = width[4 - u];
        const std::string pre = "up" + std::to_string(u);
        if (pack_convt(ctx, sd, n, pre + ".up", cin, cin / 2, W[wi++])) return -4;
        if (pack_conv(ctx, sd, n, pre + ".conv.double_conv.0", pre + ".conv.double_conv.1", cout, cin, 3, W[wi++])) return -4;
        if (pack_conv(ctx, sd, n, pre + ".conv.double_conv.3", pre + ".conv.double_conv.4", cout, cout, 3, W[wi++])) return -4;
    }
    const cvb_tensor* ow = find(sd, n, "outc.conv.weight");
    const cvb_tensor* ob = find(sd, n, "outc.conv.bias");
    if (!ow || !ob || numel(ow) != 64 || numel(ob) != 1) return fail(ctx, -4, "missing/bad outc.conv tensors");
    if (dalloc(ctx, &ctx->outc_w, 64)) return -3;
    CK(cudaMemcpy(ctx->outc_w, ow->data, 64 * sizeof(float), cudaMemcpyHostToDevice));
    ctx->outc_b = ob->data[0];

    // launch plan (see unet_forward for the order)
    auto& P = ctx->unet_plan;
    P.resize(21);
    int rc = 0;
    // encoder
    rc |= build_conv(ctx, P[0], ctx->t0, B, 256, 256, 64, 0, 64, W[0], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[0], ctx->cat0, 128, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[1], ctx->p1, B, 128, 128, 64, 0, 64, W[1], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[1], ctx->t1, 128, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[2], ctx->t1, B, 128, 128, 128, 0, 128, W[2], 3, 1, EPI_STORE);    rc |= set_store(ctx, P[2], ctx->cat1, 256, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[3], ctx->p2, B, 64, 64, 128, 0, 128, W[3], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[3], ctx->t2, 256, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[4], ctx->t2, B, 64, 64, 256, 0, 256, W[4], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[4], ctx->cat2, 512, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[5], ctx->p3, B, 32, 32, 256, 0, 256, W[5], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[5], ctx->t3, 512, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[6], ctx->t3, B, 32, 32, 512, 0, 512, W[6], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[6], ctx->cat3, 1024, 0, 1, nullptr, 0);
    if (getenv("CVB_NO_POOL_FUSE") == nullptr) {
        // MaxPool2d(2) (unet_parts.py:34) of each encoder level fused into the epilogue of the conv that produces it: the
        // row-streaming kernel pools over row pairs, the tile kernels inside a warp (tile rows 8 or 16 lanes apart)
        struct { int conv; __half* out; int c; } pools[4] = {{0, ctx->p1, 64}, {2, ctx->p2, 128}, {4, ctx->p3, 256}, {6, ctx->p4, 512}};
        for (auto& q : pools) {
            ConvLaunch& L = P[q.conv];
            if (L.variant == 2 || (L.p.tn == 1 && (L.p.tw == 8 || L.p.tw == 16) && L.p.th % 2 == 0)) {
                L.p.pool_out = q.out;
                L.p.pool_c_stride = q.c;
            }
        }
    }
    rc |= build_conv(ctx, P[7], ctx->p4, B, 16, 16, 512, 0, 512, W[7], 3, 1, EPI_STORE);      rc |= set_store(ctx, P[7], ctx->t4, 1024, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[8], ctx->t4, B, 16, 16, 1024, 0, 1024, W[8], 3, 1, EPI_STORE);    rc |= set_store(ctx, P[8], ctx->x5, 1024, 0, 1, nullptr, 0);
    // decoder: convT writes the upper channel half of the concat buffer (torch.cat([skip, up]), unet_parts.py:67)
    rc |= build_conv(ctx, P[9], ctx->x5, B, 16, 16, 1024, 0, 1024, W[9], 1, 1, EPI_CONVT);    rc |= set_store(ctx, P[9], ctx->cat3, 1024, 512, 0, nullptr, 0);  P[9].p.convt_cout = 512;
    rc |= build_conv(ctx, P[10], ctx->cat3, B, 32, 32, 1024, 0, 1024, W[10], 3, 1, EPI_STORE); rc |= set_store(ctx, P[10], ctx->t3, 512, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[11], ctx->t3, B, 32, 32, 512, 0, 512, W[11], 3, 1, EPI_STORE);    rc |= set_store(ctx, P[11], ctx->u1, 512, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[12], ctx->u1, B, 32, 32, 512, 0, 512, W[12], 1, 1, EPI_CONVT);    rc |= set_store(ctx, P[12], ctx->cat2, 512, 256, 0, nullptr, 0);  P[12].p.convt_cout = 256;
    rc |= build_conv(ctx, P[13], ctx->cat2, B, 64, 64, 512, 0, 512, W[13], 3, 1, EPI_STORE);  rc |= set_store(ctx, P[13], ctx->t2, 256, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[14], ctx->t2, B, 64, 64, 256, 0, 256, W[14], 3, 1, EPI_STORE);    rc |= set_store(ctx, P[14], ctx->u2, 256, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[15], ctx->u2, B, 64, 64, 256, 0, 256, W[15], 1, 1, EPI_CONVT);    rc |= set_store(ctx, P[15], ctx->cat1, 256, 128, 0, nullptr, 0);  P[15].p.convt_cout = 128;
    rc |= build_conv(ctx, P[16], ctx->cat1, B, 128, 128, 256, 0, 256, W[16], 3, 1, EPI_STORE); rc |= set_store(ctx, P[16], ctx->t1, 128, 0, 1, nullptr, 0);
    ctx->fuse_up4 = getenv("CVB_NO_CONVT_FUSE") == nullptr;
    if (ctx->fuse_up4) {
        // up3.conv.3 + up4.up in one kernel (conv_convt_kernel): the 128-channel tensor between them never reaches HBM
        rc |= build_conv(ctx, P[17], ctx->t1, B, 128, 128, 128, 0, 128, W[17], 3, 1, EPI_FUSED_CONVT);
        if (!rc && conv_set_fused_convt(P[17], W[18].w, W[18].bias, 64, ctx->cat0, 128, 64)) rc = fail(ctx, -6, "fused conv + transposed conv plan failed");
        memset(&P[18], 0, sizeof P[18]);
    } else {
        rc |= build_conv(ctx, P[17], ctx->t1, B, 128, 128, 128, 0, 128, W[17], 3, 1, EPI_STORE);  rc |= set_store(ctx, P[17], ctx->u3, 128, 0, 1, nullptr, 0);
        rc |= build_conv(ctx, P[18], ctx->u3, B, 128, 128, 128, 0, 128, W[18], 1, 1, EPI_CONVT);  rc |= set_store(ctx, P[18], ctx->cat0, 128, 64, 0, nullptr, 0);   P[18].p.convt_cout = 64;
    }
    rc |= build_conv(ctx, P[19], ctx->cat0, B, 256, 256, 128, 0, 128, W[19], 3, 1, EPI_STORE); rc |= set_store(ctx, P[19], ctx->t0, 64, 0, 1, nullptr, 0);
    rc |= build_conv(ctx, P[20], ctx->t0, B, 256, 256, 64, 0, 64, W[20], 3, 1, EPI_OUTC);     // + outc 1x1 + sigmoid/threshold
    P[20].p.relu = 1;
    P[20].p.out_bufs = 0;
    P[20].p.outc_w = ctx->outc_w;
    P[20].p.outc_b = ctx->outc_b;
    if (rc) return -5;
    ctx->unet_loaded = true;
    return 0;
}

int cvb_load_resnet18(cvb_ctx* ctx, const cvb_tensor* sd, int n) {
    if (!ctx || !sd) return -1;
    CVB_ON_DEVICE(ctx);
    if (ctx->resnet_loaded) return fail(ctx, -8, "classifier weights already loaded");
    const int S = ctx->max_batch * 64;
    if (pack_stem(ctx, sd, n, "conv1", "bn1", 1, 7, &ctx->rstem_w, &ctx->rstem_b)) return -4;
    if (pack_stem_tc(ctx, sd, n, "conv1", "bn1", 1, 7, 8, 1, 64, &ctx->rstem_wsw, 66)) return -4;
    const cvb_tensor* fw = find(sd, n, "fc.weight");
    const cvb_tensor* fb = find(sd, n, "fc.bias");
    if (!fw || !fb || numel(fw) != 13 * 512 || numel(fb) != 13) return fail(ctx, -4, "missing/bad fc tensors");
    if (dalloc(ctx, &ctx->fc_w, 13 * 512) || dalloc(ctx, &ctx->fc_b, 13)) return -3;
    CK(cudaMemcpy(ctx->fc_w, fw->data, 13 * 512 * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->fc_b, fb->data, 13 * sizeof(float), cudaMemcpyHostToDevice));

    static const int width[4] = {64, 128, 256, 512};
    ctx->res_w.reserve(32);
    ctx->res_plan.reserve(32);
    int rc = 0;
    auto add = [&](const std::string& conv, const std::string& bn, const __half* in, int Hin, int cin, int cout, int k, int stride,
                   __half* out, int relu, const __half* res) {
        ctx->res_w.emplace_back();
        if (pack_conv(ctx, sd, n, conv, bn, cout, cin, k, ctx->res_w.back())) { rc = -4; return; }
        ctx->res_plan.emplace_back();
        if (build_conv(ctx, ctx->res_plan.back(), in, S, Hin, Hin, cin, 0, cin, ctx->res_w.back(), k, stride, EPI_STORE)) { rc = -5; return; }
        if (set_store(ctx, ctx->res_plan.back(), out, cout, 0, relu, res, cout)) { rc = -6; return; }
    };
    // Buffers per level: a = rbuf[3l], b = rbuf[3l+1], c = rbuf[3l+2].  Level input arrives in `x`.
    __half* x = ctx->rbuf[0];  // stem output lives in level-0 buffer a
    int H = 16;
    for (int l = 0; l < 4 && !rc; ++l) {
        const std::string L = "layer" + std::to_string(l + 1);
        __half *a = ctx->rbuf[3 * l], *b = ctx->rbuf[3 * l + 1], *c = ctx->rbuf[3 * l + 2];
        const int cout = width[l];
        if (l == 0) {
            // x == a
            add(L + ".0.conv1", L + ".0.bn1", a, H, 64, 64, 3, 1, b, 1, nullptr);
            add(L + ".0.conv2", L + ".0.bn2", b, H, 64, 64, 3, 1, c, 1, a);
            add(L + ".1.conv1", L + ".1.bn1", c, H, 64, 64, 3, 1, b, 1, nullptr);
            add(L + ".1.conv2", L + ".1.bn2", b, H, 64, 64, 3, 1, a, 1, c);
            x = a;
        } else {
            const int cin = width[l - 1];
            add(L + ".0.conv1", L + ".0.bn1", x, H, cin, cout, 3, 2, a, 1, nullptr);
            add(L + ".0.downsample.0", L + ".0.downsample.1", x, H, cin, cout, 1, 2, b, 0, nullptr);
            H /= 2;
            add(L + ".0.conv2", L + ".0.bn2", a, H, cout, cout, 3, 1, c, 1, b);
            add(L + ".1.conv1", L + ".1.bn1", c, H, cout, cout, 3, 1, a, 1, nullptr);
            add(L + ".1.conv2", L + ".1.bn2", a, H, cout, cout, 3, 1, b, 1, c);
            x = b;
        }
    }
    if (rc) return rc;
    if (x != ctx->rbuf[10]) return fail(ctx, -9, "internal: unexpected classifier buffer rotation");  // the head reads rbuf[10]
    ctx->resnet_loaded = true;
    return 0;
}

int cvb_resize_area_half(cvb_ctx* ctx, const uint8_t* img, int N, int h, int w, uint8_t* out, void* stream) {
    if (!ctx || !img || !out || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    CK(launch_resize_area_half(img, out, N, h, w, static_cast<cudaStream_t>(stream)));
    ctx->launches++;
    return 0;
}

int cvb_unet_forward(cvb_ctx* ctx, const uint8_t* img, int N, float thr, float* logits, uint8_t* mask, void* stream) {
    if (!ctx || !img || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    for (int off = 0; off < N; off += ctx->max_batch) {
        const int n = N - off < ctx->max_batch ? N - off : ctx->max_batch;
        if (unet_forward(ctx, img + static_cast<size_t>(off) * 786432, n, thr, logits ? logits + static_cast<size_t>(off) * 65536 : nullptr,
                         mask ? mask + static_cast<size_t>(off) * 65536 : nullptr, static_cast<cudaStream_t>(stream)))
            return -2;
    }
    return 0;
}

int cvb_mask_from_logits(cvb_ctx* ctx, const float* logits, int N, float thr, uint8_t* mask, void* stream) {
    if (!ctx || !logits || !mask || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    CK(launch_mask_from_logits(logits, mask, thr, static_cast<long long>(N) * 65536, static_cast<cudaStream_t>(stream)));
    ctx->launches++;
    return 0;
}

int cvb_mask_to_quad(cvb_ctx* ctx, const uint8_t* mask, int N, int32_t* quad, uint8_t* found, int32_t* status, void* stream) {
    if (!ctx || !mask || !quad || !found || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    for (int off = 0; off < N; off += ctx->group) {
        const int n = N - off < ctx->group ? N - off : ctx->group;
        CK(launch_mask_to_quad(mask + static_cast<size_t>(off) * 65536, quad + off * 8, found + off,
                               status ? status + off : ctx->ws_status, ctx->ws_ncont, ctx->ws_owner, ctx->ws_quad_big, n, ctx->quad_full_only,
                               static_cast<cudaStream_t>(stream)));
        ctx->launches += ctx->quad_full_only ? 2 : 3;
    }
    return 0;
}

int cvb_warp_squares(cvb_ctx* ctx, const uint8_t* img, const int32_t* quad, const uint8_t* found, int N, int H, int W, uint8_t* board,
                     void* stream) {
    if (!ctx || !img || !quad || !found || !board || N < 0 || H <= 0 || W <= 0) return -1;
    CVB_ON_DEVICE(ctx);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    for (int off = 0; off < N; off += ctx->group) {
        const int n = N - off < ctx->group ? N - off : ctx->group;
        CK(launch_homography(quad + off * 8, found + off, ctx->ws_minv, n, static_cast<float>(H) / 256.0f, 512, 512, s));
        CK(launch_warp_board(img + static_cast<size_t>(off) * H * W * 3, ctx->ws_minv, found + off, board + static_cast<size_t>(off) * 262144, nullptr, n, H, W, s));
        ctx->launches += 2;
    }
    return 0;
}

int cvb_warp_perspective(cvb_ctx* ctx, const uint8_t* img, int H, int W, int C, const float* corners, int out_w, int out_h, uint8_t* out,
                         void* stream) {
    if (!ctx || !img || !corners || !out || H <= 0 || W <= 0 || (C != 1 && C != 3) || out_w <= 0 || out_h <= 0) return -1;
    CVB_ON_DEVICE(ctx);
    CK(launch_warp_perspective(img, H, W, C, corners, ctx->ws_minv, out, out_w, out_h, static_cast<cudaStream_t>(stream)));
    ctx->launches += 2;
    return 0;
}

int cvb_classify(cvb_ctx* ctx, const uint8_t* board, int N, int flip, float* probs, uint8_t* labels, uint8_t* labels_valid, char* fen,
                 void* stream) {
    if (!ctx || !board || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    for (int off = 0; off < N; off += ctx->max_batch) {
        const int n = N - off < ctx->max_batch ? N - off : ctx->max_batch;
        const size_t o = off;
        if (classify(ctx, board + o * 262144, n, flip, probs ? probs + o * 832 : nullptr, labels ? labels + o * 64 : nullptr,
                     labels_valid ? labels_valid + o * 64 : nullptr, fen ? fen + o * 144 : nullptr, static_cast<cudaStream_t>(stream)))
            return -2;
    }
    return 0;
}

int cvb_image_to_fen(cvb_ctx* ctx, const uint8_t* img, int N, float thr, int flip, const cvb_outputs* out, void* stream) {
    if (!ctx || !img || !out || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    for (int off = 0; off < N; off += ctx->group) {
        const int n = N - off < ctx->group ? N - off : ctx->group;
        if (pipeline_group(ctx, img + static_cast<size_t>(off) * 786432, n, thr, flip, *out, off, static_cast<cudaStream_t>(stream), nullptr))
            return -2;
    }
    return 0;
}

int cvb_image_to_fen_hw(cvb_ctx* ctx, const uint8_t* img, int N, int H, int W, float thr, int flip, const cvb_outputs* out, void* stream) {
    if (!ctx || !img || !out || N < 0 || H <= 0 || W <= 0) return -1;
    CVB_ON_DEVICE(ctx);
    if ((H != 512 || W != 512) && prepare_size(ctx, H, W)) return -5;
    for (int off = 0; off < N; off += ctx->group) {
        const int n = N - off < ctx->group ? N - off : ctx->group;
        if (pipeline_group(ctx, img + static_cast<size_t>(off) * H * W * 3, n, thr, flip, *out, off, static_cast<cudaStream_t>(stream), nullptr, H, W))
            return -2;
    }
    return 0;
}

int cvb_unet_forward_hw(cvb_ctx* ctx, const uint8_t* img, int N, int H, int W, float thr, float* logits, uint8_t* mask, void* stream) {
    if (!ctx || !img || N < 0 || H <= 0 || W <= 0) return -1;
    CVB_ON_DEVICE(ctx);
    if ((H != 512 || W != 512) && prepare_size(ctx, H, W)) return -5;
    const int B = ctx->max_batch;
    for (int off = 0; off < N; off += B) {
        const int n = N - off < B ? N - off : B;
        if (unet_forward_hw(ctx, img + static_cast<size_t>(off) * H * W * 3, n, H, W, thr, logits ? logits + static_cast<size_t>(off) * 65536 : nullptr,
                            mask ? mask + static_cast<size_t>(off) * 65536 : ctx->ws_mask, static_cast<cudaStream_t>(stream)))
            return -2;
    }
    return 0;
}

int cvb_resize_area(cvb_ctx* ctx, const uint8_t* img, int N, int H, int W, uint8_t* out, void* stream) {
    if (!ctx || !img || !out || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    return resize_to_256(ctx, img, N, H, W, out, static_cast<cudaStream_t>(stream)) ? -5 : 0;
}

// Copy-out of one group's results from its device slot to the caller's host arrays (stream s_out); records ev_free[sl].
static int copy_out_group(cvb_ctx* ctx, const cvb_outputs* oh, int sl, size_t o, int n) {
    const cvb_outputs& d = ctx->slot_out[sl];
    cudaStream_t so = ctx->s_out;
    CK(cudaStreamWaitEvent(so, ctx->ev_comp[sl], 0));
    if (oh->quad) CK(cudaMemcpyAsync(oh->quad + o * 8, d.quad, n * 8 * sizeof(int32_t), cudaMemcpyDeviceToHost, so));
    if (oh->found) CK(cudaMemcpyAsync(oh->found + o, d.found, n, cudaMemcpyDeviceToHost, so));
    if (oh->status) CK(cudaMemcpyAsync(oh->status + o, d.status, n * sizeof(int32_t), cudaMemcpyDeviceToHost, so));
    if (oh->probs) CK(cudaMemcpyAsync(oh->probs + o * 832, d.probs, static_cast<size_t>(n) * 832 * sizeof(float), cudaMemcpyDeviceToHost, so));
    if (oh->labels) CK(cudaMemcpyAsync(oh->labels + o * 64, d.labels, n * 64, cudaMemcpyDeviceToHost, so));
    if (oh->labels_valid) CK(cudaMemcpyAsync(oh->labels_valid + o * 64, d.labels_valid, n * 64, cudaMemcpyDeviceToHost, so));
    if (oh->fen) CK(cudaMemcpyAsync(oh->fen + o * 144, d.fen, n * 144, cudaMemcpyDeviceToHost, so));
    if (oh->logits) CK(cudaMemcpyAsync(oh->logits + o * 65536, d.logits, static_cast<size_t>(n) * 65536 * sizeof(float), cudaMemcpyDeviceToHost, so));
    if (oh->mask) CK(cudaMemcpyAsync(oh->mask + o * 65536, d.mask, static_cast<size_t>(n) * 65536, cudaMemcpyDeviceToHost, so));
    if (oh->board) CK(cudaMemcpyAsync(oh->board + o * 262144, d.board, static_cast<size_t>(n) * 262144, cudaMemcpyDeviceToHost, so));
    if (oh->squares) CK(cudaMemcpyAsync(oh->squares + o * 262144, d.squares, static_cast<size_t>(n) * 262144, cudaMemcpyDeviceToHost, so));
    CK(cudaEventRecord(ctx->ev_free[sl], so));
    return 0;
}

int cvb_image_to_fen_host_progress(cvb_ctx* ctx, const uint8_t* img_host, int N, float thr, int flip, const cvb_outputs* oh,
                                   volatile int32_t* boards_done) {
    if (!ctx || !img_host || !oh || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    if (boards_done) *boards_done = 0;
    const int B = ctx->max_batch, G = ctx->group;
    // Software pipeline over groups g (two device slots): copy-in(g) and compute(g) are enqueued BEFORE the copy-out of
    // group g-1, so a copy-out that blocks the calling thread (pageable destination) or the wait for it below never
    // leaves the GPU without queued work.
    int g = 0, prev_n = 0;
    size_t prev_o = 0;
    for (int off = 0; off < N; off += G, ++g) {
        const int n = N - off < G ? N - off : G;
        const int sl = g & 1;
        const size_t o = off;
        const cvb_outputs& d = ctx->slot_out[sl];
        // the slot is reusable once the copy-out of the group that used it two iterations ago has finished
        if (g >= 2) CK(cudaStreamWaitEvent(ctx->s_in, ctx->ev_free[sl], 0));
        for (int c = 0, c0 = 0; c0 < n; ++c, c0 += B) {
            const int nb = n - c0 < B ? n - c0 : B;
            CK(cudaMemcpyAsync(ctx->slot_img[sl] + static_cast<size_t>(c0) * 786432, img_host + (o + c0) * 786432,
                               static_cast<size_t>(nb) * 786432, cudaMemcpyHostToDevice, ctx->s_in));
            CK(cudaEventRecord(ctx->ev_in[sl][c], ctx->s_in));
        }
        if (g >= 2) CK(cudaStreamWaitEvent(ctx->s_comp, ctx->ev_free[sl], 0));
        cvb_outputs dev = d;
        if (!oh->logits) dev.logits = nullptr;
        if (!oh->mask) dev.mask = nullptr;
        if (!oh->board) dev.board = nullptr;
        if (!oh->squares) dev.squares = nullptr;
        if (pipeline_group(ctx, ctx->slot_img[sl], n, thr, flip, dev, 0, ctx->s_comp, ctx->ev_in[sl].data())) return -2;
        CK(cudaEventRecord(ctx->ev_comp[sl], ctx->s_comp));
        if (g >= 1) {
            if (copy_out_group(ctx, oh, sl ^ 1, prev_o, prev_n)) return -2;
            if (boards_done) {   // progress for a caller that consumes finished groups while the rest is still running
                CK(cudaEventSynchronize(ctx->ev_free[sl ^ 1]));
                *boards_done = static_cast<int32_t>(prev_o + prev_n);
            }
        }
        prev_o = o;
        prev_n = n;
    }
    if (g >= 1 && copy_out_group(ctx, oh, (g - 1) & 1, prev_o, prev_n)) return -2;
    CK(cudaStreamSynchronize(ctx->s_in));
    CK(cudaStreamSynchronize(ctx->s_comp));
    CK(cudaStreamSynchronize(ctx->s_out));
    if (boards_done) *boards_done = N;
    return 0;
}

int cvb_image_to_fen_host(cvb_ctx* ctx, const uint8_t* img_host, int N, float thr, int flip, const cvb_outputs* oh) {
    return cvb_image_to_fen_host_progress(ctx, img_host, N, thr, flip, oh, nullptr);
}

int cvb_conv2d_f16(cvb_ctx* ctx, const void* in, int N, int H, int W, int Cin, const void* w_packed, const float* bias, int Cout,
                   int ksize, int stride, int relu, const void* residual, void* out, void* stream) {
    if (!ctx || !in || !w_packed || !bias || !out) return -1;
    CVB_ON_DEVICE(ctx);
    if ((ksize != 1 && ksize != 3) || (stride != 1 && stride != 2)) return fail(ctx, -5, "conv2d: ksize/stride not supported");
    ConvWeights cw;
    cw.w = const_cast<__half*>(static_cast<const __half*>(w_packed));
    cw.bias = const_cast<float*>(bias);
    cw.rows = Cout;
    cw.K = ksize * ksize * Cin;
    ConvLaunch L;
    if (build_conv(ctx, L, static_cast<const __half*>(in), N, H, W, Cin, 0, Cin, cw, ksize, stride, EPI_STORE)) return -5;
    if (set_store(ctx, L, static_cast<__half*>(out), Cout, 0, relu, static_cast<const __half*>(residual), Cout)) return -6;
    return run_conv(ctx, L, N, static_cast<cudaStream_t>(stream));
}

int cvb_convt2x2_f16(cvb_ctx* ctx, const void* in, int N, int H, int W, int Cin, const void* w_packed, const float* bias, int Cout,
                     void* out, int out_c_stride, int out_c_off, void* stream) {
    if (!ctx || !in || !w_packed || !bias || !out) return -1;
    CVB_ON_DEVICE(ctx);
    ConvWeights cw;
    cw.w = const_cast<__half*>(static_cast<const __half*>(w_packed));
    cw.bias = const_cast<float*>(bias);
    cw.rows = 4 * Cout;
    cw.K = Cin;
    ConvLaunch L;
    if (build_conv(ctx, L, static_cast<const __half*>(in), N, H, W, Cin, 0, Cin, cw, 1, 1, EPI_CONVT)) return -5;
    if (set_store(ctx, L, static_cast<__half*>(out), out_c_stride, out_c_off, 0, nullptr, 0)) return -6;
    L.p.convt_cout = Cout;
    return run_conv(ctx, L, N, static_cast<cudaStream_t>(stream));
}

int cvb_conv3x3_convt2x2_f16(cvb_ctx* ctx, const void* in, int N, int H, int W, int Cin, const void* w_packed, const float* bias,
                             const void* w2_packed, const float* bias2, int Cout2, void* out, int out_c_stride, int out_c_off, void* stream) {
    if (!ctx || !in || !w_packed || !bias || !w2_packed || !bias2 || !out) return -1;
    CVB_ON_DEVICE(ctx);
    ConvWeights cw;
    cw.w = const_cast<__half*>(static_cast<const __half*>(w_packed));
    cw.bias = const_cast<float*>(bias);
    cw.rows = 128;
    cw.K = 9 * Cin;
    ConvLaunch L;
    if (build_conv(ctx, L, static_cast<const __half*>(in), N, H, W, Cin, 0, Cin, cw, 3, 1, EPI_FUSED_CONVT)) return -5;
    if (conv_set_fused_convt(L, static_cast<const __half*>(w2_packed), bias2, Cout2, static_cast<__half*>(out), out_c_stride, out_c_off))
        return fail(ctx, -5, "fused conv + transposed conv: only 128 -> 64 channels is supported");
    return run_conv(ctx, L, N, static_cast<cudaStream_t>(stream));
}

int cvb_unet_stem(cvb_ctx* ctx, const uint8_t* img, int N, void* out, void* stream) {
    if (!ctx || !img || !out || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    if (!ctx->unet_loaded) return fail(ctx, -7, "UNet weights not loaded (call cvb_load_unet)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (ctx->stem_fp32) {
        CK(launch_unet_stem(img, ctx->stem_w, ctx->stem_b, static_cast<__half*>(out), N, 256, 256, 64, s));
    } else {
        CUtensorMap om;
        if (tmap_act(&om, out, 64, 256, 256, N, 64, 256 * 64, 65536LL * 64, 64, 2, 1)) return fail(ctx, -6, "cuTensorMapEncodeTiled failed");
        CK(launch_unet_stem_tc(img, ctx->stem_wsw, ctx->stem_b, &om, N, ctx->sm_count, s));
    }
    ctx->launches++;
    return 0;
}

int cvb_resnet_stem(cvb_ctx* ctx, const uint8_t* board, int N, void* out, void* stream) {
    if (!ctx || !board || !out || N < 0) return -1;
    CVB_ON_DEVICE(ctx);
    if (!ctx->resnet_loaded) return fail(ctx, -7, "classifier weights not loaded (call cvb_load_resnet18)");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (ctx->stem_fp32) CK(launch_resnet_stem(board, ctx->rstem_w, ctx->rstem_b, static_cast<__half*>(out), N, s));
    else CK(launch_resnet_stem_tc(board, ctx->rstem_wsw, ctx->rstem_b, static_cast<__half*>(out), N, ctx->sm_count, s));
    ctx->launches++;
    return 0;
}

int cvb_profile(cvb_ctx* ctx, int enable) {
    if (!ctx) return -1;
    CVB_ON_DEVICE(ctx);
    if (enable && ctx->pev.empty()) {
        ctx->pev.resize(kMaxProfileEvents);
        ctx->pev_stage.resize(kMaxProfileEvents);
        for (auto& e : ctx->pev) CK(cudaEventCreate(&e));
    }
    ctx->profile = enable != 0;
    ctx->pev_used = 0;
    for (float& v : ctx->stage_ms) v = 0.f;
    return 0;
}

int cvb_profile_read(cvb_ctx* ctx, float* ms_out, int n_stages) {
    if (!ctx || !ms_out) return -1;
    CVB_ON_DEVICE(ctx);
    CK(cudaDeviceSynchronize());
    for (int i = 0; i + 1 < ctx->pev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->pev[i], ctx->pev[i + 1]) == cudaSuccess) ctx->stage_ms[ctx->pev_stage[i]] += ms;
    }
    ctx->pev_used = 0;
    for (int i = 0; i < n_stages && i < kStages; ++i) ms_out[i] = ctx->stage_ms[i];
    for (float& v : ctx->stage_ms) v = 0.f;
    return 0;
}

}  // extern "C"
