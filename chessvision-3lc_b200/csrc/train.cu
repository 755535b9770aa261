// UNet board-extractor training step on the B200 (BASELINE.json configs[4]).
//
// Reference semantics (scripts/train/train_unet.py:293-323 with autocast, i.e. fp16 tensor-core operands, fp32 master
// weights and fp32 accumulation): forward in BatchNorm *training* mode (batch statistics, running statistics updated),
// loss = BCEWithLogits(mean) + 1 - mean_n Dice_n (chessvision/pytorch_unet/utils/dice_score.py:5-30), backward,
// clip_grad_norm_(1.0), RMSprop(lr, weight_decay 1e-8, momentum 0.999) (train_unet.py:236-242).
//
// The step is split so that data-parallel training can all-reduce the flat fp32 gradient buffer between the two halves
// (cvb_train_forward_backward -> NCCL all-reduce of cvb_train_grads -> cvb_train_optimizer_step); the reference has no
// distributed code, the split mirrors torch DDP semantics (gradient average, local BatchNorm statistics).
//
// Master weights, gradients and optimizer state live in ONE flat fp32 buffer each, in the packed layout the tensor-core
// kernels read ([Cout][tap][Cin] for 3x3 convs, [(dy,dx)][Cout][Cin] for transposed convs), so that preparing the fp16
// operands is a single cast and the optimizer is a single elementwise kernel.
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "ctx.h"
#include "kernels.h"
#include "train_kernels.h"
#include "wgrad_tc.h"

namespace {

constexpr int kTileCapacity = 4096;

struct ConvL {   // Conv3x3 (no bias) + BatchNorm + ReLU
    std::string conv, bn;
    int Cin, Cout, H;
    size_t w_off, g_off, b_off;           // offsets into the flat parameter buffer (weights, gamma, beta)
    size_t bn_off;                        // offset into the per-channel BatchNorm workspaces
    const __half* x; int x_stride;        // input view (first Cin channels of a buffer with x_stride channels per pixel)
    __half* z;                            // conv output, dense [B,H,H,Cout]
    __half* y; int y_stride;              // activation output view
    ConvLaunch fwd, dgrad;
    bool has_dgrad;
    __half* dz;                           // which ping-pong buffer holds dz during backward
    __half* dy;                           // ... and dy
    __half* dx; int dx_stride;            // data-gradient destination
    WgLaunch wgrad;
};

struct ConvTL {   // ConvTranspose2d(k=2, s=2) + bias, writing the upper channel half of a concat buffer
    std::string name;
    int Cin, Cout, H;                     // H = input extent
    size_t w_off, b_off;
    const __half* x;                      // dense [B,H,H,Cin]
    __half* cat; const __half* dcat;      // [B,2H,2H,2*Cout], channels [Cout, 2*Cout)
    ConvLaunch fwd, dgrad;
    __half* dx;                           // dense [B,H,H,Cin]
    WgLaunch wgrad;
};

}  // namespace

struct cvb_trainer {
    cvb_train_config cfg;
    int B = 0;
    size_t n_params = 0;                  // padded to a multiple of 4
    float *P = nullptr, *G = nullptr, *SQ = nullptr, *MB = nullptr;   // master weights, gradients, RMSprop state
    __half *W16 = nullptr, *W16T = nullptr;                           // fp16 operands: forward / data-gradient layouts
    float* zero_bias = nullptr;
    // BatchNorm workspaces, pooled over layers (bn_off)
    size_t bn_channels = 0;
    float *scale = nullptr, *shift = nullptr, *mean = nullptr, *rstd = nullptr, *run_mean = nullptr, *run_var = nullptr;
    double* sums = nullptr;               // [2*bn_channels] forward sums | [2*bn_channels] backward sums | loss sums | norm
    size_t sums_count = 0;
    double *lsums = nullptr, *norm = nullptr;
    std::vector<ConvL> conv;
    std::vector<ConvTL> convt;
    size_t outc_w_off = 0, outc_b_off = 0;
    // activations
    __half *t0a, *cat0, *p1, *t1a, *cat1, *p2, *t2a, *cat2, *p3, *t3a, *cat3, *p4, *t4a, *x5;
    __half *t3b, *u1, *t2b, *u2, *t1b, *u3, *t0b, *t0c;
    __half *dcat0, *dcat1, *dcat2, *dcat3, *gA, *gB;
    float* logits = nullptr;
    WgTile* tiles = nullptr;
    int tiles_used = 0;
    long long steps = 0;
    // gradient buckets for the data-parallel all-reduce, in the order the backward pass completes them (the flat layout is
    // the forward order, so they are contiguous ranges taken from its end): [lo, hi) in parameters + the event recorded on
    // the compute stream when the last kernel writing into the range has been enqueued
    static constexpr int kBuckets = 4;
    size_t bucket_lo[kBuckets] = {0, 0, 0, 0}, bucket_hi[kBuckets] = {0, 0, 0, 0};
    cudaEvent_t bucket_ev[kBuckets] = {nullptr, nullptr, nullptr, nullptr};
};

namespace {

const cvb_tensor* find_t(const cvb_tensor* sd, int n, const std::string& name) {
    for (int i = 0; i < n; ++i)
        if (name == sd[i].name) return &sd[i];
    return nullptr;
}
int64_t numel_t(const cvb_tensor* t) {
    int64_t k = 1;
    for (int i = 0; i < t->ndim; ++i) k *= t->shape[i];
    return k;
}

void tap_shift(int tap, int8_t& dy, int8_t& dx) {
    dy = static_cast<int8_t>(tap / 3 - 1);
    dx = static_cast<int8_t>(tap % 3 - 1);
}

// weight-gradient plan of a 3x3 convolution: dz dense [B,H,H,Cout], x view with x_stride channels per pixel
int build_conv_wgrad(cvb_ctx* ctx, cvb_trainer* T, ConvL& L) {
    WgLaunch& W = L.wgrad;
    memset(&W, 0, sizeof W);
    const int B = T->B, H = L.H;
    int rc = wgrad_tmap(&W.p.maps[0], L.dz, L.Cout, H, H, B, L.Cout, static_cast<int64_t>(H) * L.Cout, static_cast<int64_t>(H) * H * L.Cout);
    rc |= wgrad_tmap(&W.p.maps[1], L.x, L.x_stride, H, H, B, L.x_stride, static_cast<int64_t>(H) * L.x_stride,
                     static_cast<int64_t>(H) * H * L.x_stride);
    for (int i = 2; i < 6; ++i) W.p.maps[i] = W.p.maps[0];
    if (rc) return fail(ctx, -6, "cuTensorMapEncodeTiled (wgrad views of %s) failed: %d", L.conv.c_str(), rc);
    std::vector<WgBlock> xb, zb;   // blocks of (tap, input-channel block) and of output-channel blocks
    for (int tap = 0; tap < 9; ++tap)
        for (int cb = 0; cb < L.Cin / 64; ++cb) {
            WgBlock b = {1, static_cast<int16_t>(cb * 64), 0, 0, 0, static_cast<int64_t>(tap) * L.Cin + cb * 64};
            tap_shift(tap, b.dy, b.dx);
            xb.push_back(b);
        }
    const long long K = 9LL * L.Cin;
    float* out = T->G + L.w_off;
    const float inv_s = 1.0f / T->cfg.loss_scale;
    if (T->tiles_used >= kTileCapacity) return fail(ctx, -5, "wgrad tile table full");
    WgTile* table = T->tiles + T->tiles_used;
    const int room = kTileCapacity - T->tiles_used;
    if (L.Cout >= 128) {   // rows = output channels, columns = (tap, input channel)
        for (int cb = 0; cb < L.Cout / 64; ++cb) zb.push_back({0, static_cast<int16_t>(cb * 64), 0, 0, 0, static_cast<int64_t>(cb) * 64 * K});
        const int count = static_cast<int>(xb.size());
        const int nb = count % 4 == 0 ? 4 : (count % 3 == 0 ? 3 : (count % 2 == 0 ? 2 : 1));
        rc = wgrad_build(W, zb, xb, nb, table, room, B, H, H, out, K, 1, inv_s, ctx->sm_count);
    } else {               // Cout = 64: rows = (tap, input channel), columns = output channels
        zb.push_back({0, 0, 0, 0, 0, 0});
        rc = wgrad_build(W, xb, zb, 1, table, room, B, H, H, out, 1, K, inv_s, ctx->sm_count);
    }
    if (rc) return fail(ctx, -5, "wgrad plan of %s failed: %d", L.conv.c_str(), rc);
    T->tiles_used += W.p.n_tiles;
    return 0;
}

int build_convt_wgrad(cvb_ctx* ctx, cvb_trainer* T, ConvTL& L) {
    WgLaunch& W = L.wgrad;
    memset(&W, 0, sizeof W);
    const int B = T->B, H = L.H, C2 = 2 * L.Cout, W2 = 2 * H;
    int rc = 0;
    for (int q = 0; q < 4 && !rc; ++q)   // parity views (dy,dx) of the output gradient
        rc = wgrad_tmap(&W.p.maps[q], L.dcat + (static_cast<int64_t>(q >> 1) * W2 + (q & 1)) * C2, C2, H, H, B, 2LL * C2, 2LL * W2 * C2,
                        static_cast<int64_t>(W2) * W2 * C2);
    rc |= wgrad_tmap(&W.p.maps[4], L.x, L.Cin, H, H, B, L.Cin, static_cast<int64_t>(H) * L.Cin, static_cast<int64_t>(H) * H * L.Cin);
    W.p.maps[5] = W.p.maps[4];
    if (rc) return fail(ctx, -6, "cuTensorMapEncodeTiled (wgrad views of %s) failed: %d", L.name.c_str(), rc);
    std::vector<WgBlock> U, V;
    for (int q = 0; q < 4; ++q)
        for (int cb = 0; cb < L.Cout / 64; ++cb)
            U.push_back({static_cast<int16_t>(q), static_cast<int16_t>(L.Cout + cb * 64), 0, 0, 0,
                         (static_cast<int64_t>(q) * L.Cout + cb * 64) * L.Cin});
    for (int cb = 0; cb < L.Cin / 64; ++cb) V.push_back({4, static_cast<int16_t>(cb * 64), 0, 0, 0, static_cast<int64_t>(cb) * 64});
    const int nb = L.Cin / 64 >= 4 ? 4 : L.Cin / 64;
    if (T->tiles_used >= kTileCapacity) return fail(ctx, -5, "wgrad tile table full");
    rc = wgrad_build(W, U, V, nb, T->tiles + T->tiles_used, kTileCapacity - T->tiles_used, B, H, H, T->G + L.w_off, L.Cin, 1,
                     1.0f / T->cfg.loss_scale, ctx->sm_count);
    if (rc) return fail(ctx, -5, "wgrad plan of %s failed: %d", L.name.c_str(), rc);
    T->tiles_used += W.p.n_tiles;
    return 0;
}

#define LAUNCH(call)                 \
    do {                             \
        CK(call);                    \
        ctx->launches++;             \
    } while (0)

int conv_forward(cvb_ctx* ctx, cvb_trainer* T, ConvL& L, const float* img, cudaStream_t s) {
    const long long rows = static_cast<long long>(T->B) * L.H * L.H;
    if (L.Cin == 3) LAUNCH(launch_stem_fwd(img, T->P + L.w_off, L.z, T->B, L.H, L.H, s));
    else LAUNCH((L.fwd.pdl = 0, conv_launch(L.fwd, T->B, ctx->sm_count, s)));
    LAUNCH(launch_bn_stats(L.z, T->sums + 2 * L.bn_off, rows, L.Cout, s));
    LAUNCH(launch_bn_finalize(T->sums + 2 * L.bn_off, T->P + L.g_off, T->P + L.b_off, T->scale + L.bn_off, T->shift + L.bn_off,
                              T->mean + L.bn_off, T->rstd + L.bn_off, T->run_mean + L.bn_off, T->run_var + L.bn_off, L.Cout, rows,
                              T->cfg.bn_eps, T->cfg.bn_momentum, s));
    LAUNCH(launch_bn_apply_relu(L.z, T->scale + L.bn_off, T->shift + L.bn_off, L.y, rows, L.Cout, L.y_stride, 0, s));
    return 0;
}

int conv_backward(cvb_ctx* ctx, cvb_trainer* T, ConvL& L, const float* img, cudaStream_t s) {
    const long long rows = static_cast<long long>(T->B) * L.H * L.H;
    const float inv_s = 1.0f / T->cfg.loss_scale;
    double* bs = T->sums + 2 * T->bn_channels + 2 * L.bn_off;
    LAUNCH(launch_bn_bwd(L.dy, L.z, T->scale + L.bn_off, T->shift + L.bn_off, T->mean + L.bn_off, T->rstd + L.bn_off, bs, L.dz,
                         T->G + L.g_off, T->G + L.b_off, rows, L.Cout, inv_s, s));
    ctx->launches++;   // two kernels
    if (L.Cin == 3) {
        LAUNCH(launch_stem_wgrad(img, L.dz, T->G + L.w_off, T->B, L.H, L.H, inv_s, s));
        return 0;
    }
    LAUNCH(wgrad_launch(L.wgrad, ctx->sm_count, s));
    if (L.has_dgrad) LAUNCH((L.dgrad.pdl = 0, conv_launch(L.dgrad, T->B, ctx->sm_count, s)));
    return 0;
}

int convt_backward(cvb_ctx* ctx, cvb_trainer* T, ConvTL& L, cudaStream_t s) {
    const long long rows = static_cast<long long>(T->B) * 4 * L.H * L.H;
    LAUNCH(launch_colsum(L.dcat, rows, L.Cout, 2 * L.Cout, L.Cout, T->G + L.b_off, 1.0f / T->cfg.loss_scale, s));
    LAUNCH(wgrad_launch(L.wgrad, ctx->sm_count, s));
    LAUNCH((L.dgrad.pdl = 0, conv_launch(L.dgrad, T->B, ctx->sm_count, s)));
    return 0;
}

int prepare_weights(cvb_ctx* ctx, cvb_trainer* T, cudaStream_t s) {
    LAUNCH(launch_cast_f16(T->P, T->W16, static_cast<long long>(T->n_params), s));
    for (auto& L : T->conv)
        if (L.has_dgrad)
            LAUNCH(launch_transpose_w(T->P + L.w_off, T->W16T + L.w_off, L.Cout, L.Cin, 9, 9LL * L.Cin, L.Cin, 1, s));
    for (auto& L : T->convt)
        LAUNCH(launch_transpose_w(T->P + L.w_off, T->W16T + L.w_off, 4 * L.Cout, L.Cin, 1, L.Cin, 0, 0, s));
    return 0;
}

template <class T>
int talloc(cvb_ctx* ctx, T** p, size_t count) {
    return dalloc(ctx, p, count);
}

}  // namespace

extern "C" {

int cvb_train_default_config(cvb_train_config* cfg) {
    if (!cfg) return -1;
    cfg->batch = 2;                 // scripts/bin/train_board_extractor.sh
    cfg->loss_scale = 4096.0f;
    cfg->momentum = 0.999f;         // train_unet.py:236-242
    cfg->alpha = 0.99f;             // torch.optim.RMSprop defaults
    cfg->eps = 1e-8f;
    cfg->weight_decay = 1e-8f;
    cfg->max_grad_norm = 1.0f;      // train_unet.py:321
    cfg->bn_momentum = 0.1f;        // nn.BatchNorm2d defaults
    cfg->bn_eps = 1e-5f;
    return 0;
}

int cvb_train_create(cvb_ctx* ctx, const cvb_tensor* sd, int n, const cvb_train_config* cfg_in) {
    if (!ctx || !sd) return -1;
    CVB_ON_DEVICE(ctx);
    if (ctx->trainer) return fail(ctx, -8, "a trainer already exists on this context");
    cvb_train_config cfg;
    cvb_train_default_config(&cfg);
    if (cfg_in) cfg = *cfg_in;
    if (cfg.batch < 1 || !(cfg.loss_scale > 0.f)) return fail(ctx, -1, "bad training configuration");
    if (wgrad_configure() != cudaSuccess) return fail(ctx, -2, "wgrad kernel attribute setup failed");
    cvb_trainer* T = new cvb_trainer();
    ctx->trainer = T;
    T->cfg = cfg;
    T->B = cfg.batch;
    const size_t B = cfg.batch;

    // ---- layer table and flat parameter layout
    static const int width[5] = {64, 128, 256, 512, 1024};
    size_t off = 0, bn_off = 0;
    auto add_conv = [&](const std::string& pre, int idx, int cin, int cout, int H) {
        ConvL L;
        L.conv = pre + std::to_string(idx);
        L.bn = pre + std::to_string(idx + 1);
        L.Cin = cin; L.Cout = cout; L.H = H;
        L.w_off = off; off += static_cast<size_t>(cout) * 9 * cin;
        off = (off + 63) & ~static_cast<size_t>(63);
        L.g_off = off; off += cout;
        L.b_off = off; off += cout;
        L.bn_off = bn_off; bn_off += cout;
        L.has_dgrad = cin != 3;
        T->conv.push_back(L);
    };
    add_conv("inc.double_conv.", 0, 3, 64, 256);
    add_conv("inc.double_conv.", 3, 64, 64, 256);
    for (int d = 1; d <= 4; ++d) {
        const std::string pre = "down" + std::to_string(d) + ".maxpool_conv.1.double_conv.";
        add_conv(pre, 0, width[d - 1], width[d], 256 >> d);
        add_conv(pre, 3, width[d], width[d], 256 >> d);
    }
    for (int u = 1; u <= 4; ++u) {
        const int cin = width[5 - u], cout = width[4 - u], H = 16 << u;
        ConvTL L;
        L.name = "up" + std::to_string(u) + ".up";
        L.Cin = cin; L.Cout = cin / 2; L.H = H / 2;
        L.w_off = off; off += static_cast<size_t>(4) * L.Cout * L.Cin;
        L.b_off = off; off += L.Cout;
        T->convt.push_back(L);
        const std::string pre = "up" + std::to_string(u) + ".conv.double_conv.";
        add_conv(pre, 0, cin, cout, H);
        add_conv(pre, 3, cout, cout, H);
    }
    T->outc_w_off = off; off += 64;
    T->outc_b_off = off; off += 1;
    T->n_params = (off + 3) & ~static_cast<size_t>(3);
    T->bn_channels = bn_off;
    // buckets: {up4, up3, up2 (+ outc)}, {up1}, {down4}, {down3 .. inc}: 3.9 M, 9.2 M, 14.2 M and 4.7 M parameters; the first
    // three are complete after 45 % / 55 % / 65 % of the backward pass, so their all-reduce hides behind the rest of it
    {
        const size_t cut[5] = {T->n_params, T->convt[1].w_off, T->convt[0].w_off, T->conv[8].w_off, 0};
        for (int b = 0; b < cvb_trainer::kBuckets; ++b) {
            T->bucket_lo[b] = cut[b + 1];
            T->bucket_hi[b] = cut[b];
            if (cudaEventCreateWithFlags(&T->bucket_ev[b], cudaEventDisableTiming) != cudaSuccess) return fail(ctx, -2, "event creation failed");
        }
    }

    // ---- device memory
    int rc = 0;
    rc |= talloc(ctx, &T->P, T->n_params);
    rc |= talloc(ctx, &T->G, T->n_params);
    rc |= talloc(ctx, &T->SQ, T->n_params);
    rc |= talloc(ctx, &T->MB, T->n_params);
    rc |= talloc(ctx, &T->W16, T->n_params);
    rc |= talloc(ctx, &T->W16T, T->n_params);
    rc |= talloc(ctx, &T->zero_bias, 4096);
    rc |= talloc(ctx, &T->scale, bn_off); rc |= talloc(ctx, &T->shift, bn_off); rc |= talloc(ctx, &T->mean, bn_off);
    rc |= talloc(ctx, &T->rstd, bn_off); rc |= talloc(ctx, &T->run_mean, bn_off); rc |= talloc(ctx, &T->run_var, bn_off);
    T->sums_count = 4 * bn_off + 4 * B + 1 + 2;
    rc |= talloc(ctx, &T->sums, T->sums_count);
    rc |= talloc(ctx, &T->tiles, kTileCapacity);
    const size_t px = B * 65536;
    rc |= talloc(ctx, &T->t0a, px * 64); rc |= talloc(ctx, &T->cat0, px * 128); rc |= talloc(ctx, &T->p1, px / 4 * 64);
    rc |= talloc(ctx, &T->t1a, px / 4 * 128); rc |= talloc(ctx, &T->cat1, px / 4 * 256); rc |= talloc(ctx, &T->p2, px / 16 * 128);
    rc |= talloc(ctx, &T->t2a, px / 16 * 256); rc |= talloc(ctx, &T->cat2, px / 16 * 512); rc |= talloc(ctx, &T->p3, px / 64 * 256);
    rc |= talloc(ctx, &T->t3a, px / 64 * 512); rc |= talloc(ctx, &T->cat3, px / 64 * 1024); rc |= talloc(ctx, &T->p4, px / 256 * 512);
    rc |= talloc(ctx, &T->t4a, px / 256 * 1024); rc |= talloc(ctx, &T->x5, px / 256 * 1024);
    rc |= talloc(ctx, &T->t3b, px / 64 * 512); rc |= talloc(ctx, &T->u1, px / 64 * 512);
    rc |= talloc(ctx, &T->t2b, px / 16 * 256); rc |= talloc(ctx, &T->u2, px / 16 * 256);
    rc |= talloc(ctx, &T->t1b, px / 4 * 128); rc |= talloc(ctx, &T->u3, px / 4 * 128);
    rc |= talloc(ctx, &T->t0b, px * 64); rc |= talloc(ctx, &T->t0c, px * 64);
    rc |= talloc(ctx, &T->dcat0, px * 128); rc |= talloc(ctx, &T->dcat1, px / 4 * 256); rc |= talloc(ctx, &T->dcat2, px / 16 * 512);
    rc |= talloc(ctx, &T->dcat3, px / 64 * 1024);
    rc |= talloc(ctx, &T->gA, px * 64); rc |= talloc(ctx, &T->gB, px * 64);
    rc |= talloc(ctx, &T->logits, px);
    // z buffers: one per conv layer
    for (auto& L : T->conv) rc |= talloc(ctx, &L.z, B * L.H * L.H * L.Cout);
    if (rc) return -3;
    T->lsums = T->sums + 4 * bn_off;
    T->norm = T->lsums + 4 * B + 1;
    CK(cudaMemset(T->zero_bias, 0, 4096 * sizeof(float)));
    CK(cudaMemset(T->SQ, 0, T->n_params * sizeof(float)));
    CK(cudaMemset(T->MB, 0, T->n_params * sizeof(float)));
    CK(cudaMemset(T->G, 0, T->n_params * sizeof(float)));

    // ---- parameters from the state dict (torch layouts -> packed layouts)
    std::vector<float> hp(T->n_params, 0.f), hrm(bn_off, 0.f), hrv(bn_off, 1.f);
    for (auto& L : T->conv) {
        const cvb_tensor* w = find_t(sd, n, L.conv + ".weight");
        const cvb_tensor* g = find_t(sd, n, L.bn + ".weight");
        const cvb_tensor* b = find_t(sd, n, L.bn + ".bias");
        if (!w || !g || !b || numel_t(w) != 9LL * L.Cin * L.Cout || numel_t(g) != L.Cout || numel_t(b) != L.Cout)
            return fail(ctx, -4, "state_dict lacks or mis-shapes '%s' / '%s'", L.conv.c_str(), L.bn.c_str());
        for (int co = 0; co < L.Cout; ++co)
            for (int ci = 0; ci < L.Cin; ++ci)
                for (int t = 0; t < 9; ++t)
                    hp[L.w_off + (static_cast<size_t>(co) * 9 + t) * L.Cin + ci] = w->data[(static_cast<size_t>(co) * L.Cin + ci) * 9 + t];
        memcpy(&hp[L.g_off], g->data, L.Cout * sizeof(float));
        memcpy(&hp[L.b_off], b->data, L.Cout * sizeof(float));
        const cvb_tensor* rm = find_t(sd, n, L.bn + ".running_mean");
        const cvb_tensor* rv = find_t(sd, n, L.bn + ".running_var");
        if (rm && numel_t(rm) == L.Cout) memcpy(&hrm[L.bn_off], rm->data, L.Cout * sizeof(float));
        if (rv && numel_t(rv) == L.Cout) memcpy(&hrv[L.bn_off], rv->data, L.Cout * sizeof(float));
    }
    for (auto& L : T->convt) {
        const cvb_tensor* w = find_t(sd, n, L.name + ".weight");
        const cvb_tensor* b = find_t(sd, n, L.name + ".bias");
        if (!w || !b || numel_t(w) != 4LL * L.Cin * L.Cout || numel_t(b) != L.Cout)
            return fail(ctx, -4, "state_dict lacks or mis-shapes '%s'", L.name.c_str());
        for (int ci = 0; ci < L.Cin; ++ci)
            for (int co = 0; co < L.Cout; ++co)
                for (int q = 0; q < 4; ++q)
                    hp[L.w_off + (static_cast<size_t>(q) * L.Cout + co) * L.Cin + ci] = w->data[(static_cast<size_t>(ci) * L.Cout + co) * 4 + q];
        memcpy(&hp[L.b_off], b->data, L.Cout * sizeof(float));
    }
    {
        const cvb_tensor* w = find_t(sd, n, "outc.conv.weight");
        const cvb_tensor* b = find_t(sd, n, "outc.conv.bias");
        if (!w || !b || numel_t(w) != 64 || numel_t(b) != 1) return fail(ctx, -4, "state_dict lacks or mis-shapes 'outc.conv'");
        memcpy(&hp[T->outc_w_off], w->data, 64 * sizeof(float));
        hp[T->outc_b_off] = b->data[0];
    }
    CK(cudaMemcpy(T->P, hp.data(), T->n_params * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(T->run_mean, hrm.data(), bn_off * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(T->run_var, hrv.data(), bn_off * sizeof(float), cudaMemcpyHostToDevice));

    // ---- data flow: forward views
    auto& C = T->conv;
    struct IO { const __half* x; int xs; __half* y; int ys; };
    const IO io[18] = {
        {nullptr, 0, T->t0a, 64},      {T->t0a, 64, T->cat0, 128},   {T->p1, 64, T->t1a, 128},    {T->t1a, 128, T->cat1, 256},
        {T->p2, 128, T->t2a, 256},     {T->t2a, 256, T->cat2, 512},  {T->p3, 256, T->t3a, 512},   {T->t3a, 512, T->cat3, 1024},
        {T->p4, 512, T->t4a, 1024},    {T->t4a, 1024, T->x5, 1024},  {T->cat3, 1024, T->t3b, 512}, {T->t3b, 512, T->u1, 512},
        {T->cat2, 512, T->t2b, 256},   {T->t2b, 256, T->u2, 256},    {T->cat1, 256, T->t1b, 128},  {T->t1b, 128, T->u3, 128},
        {T->cat0, 128, T->t0b, 64},    {T->t0b, 64, T->t0c, 64}};
    for (int i = 0; i < 18; ++i) { C[i].x = io[i].x; C[i].x_stride = io[i].xs; C[i].y = io[i].y; C[i].y_stride = io[i].ys; }
    const __half* tx[4] = {T->x5, T->u1, T->u2, T->u3};
    __half* tcat[4] = {T->cat3, T->cat2, T->cat1, T->cat0};
    __half* tdcat[4] = {T->dcat3, T->dcat2, T->dcat1, T->dcat0};
    for (int u = 0; u < 4; ++u) { T->convt[u].x = tx[u]; T->convt[u].cat = tcat[u]; T->convt[u].dcat = tdcat[u]; }

    // ---- backward data flow: simulate the ping-pong of (dy, dz) between gA and gB in execution order
    __half *cur = T->gA, *other = T->gB;
    auto bwd = [&](int i, __half* dx, int dx_stride) {   // dy in cur -> dz in other -> dx (default: back into cur)
        C[i].dy = cur; C[i].dz = other;
        C[i].dx = dx ? dx : cur; C[i].dx_stride = dx ? dx_stride : C[i].Cin;
    };
    auto swap = [&]() { __half* t = cur; cur = other; other = t; };
    bwd(17, nullptr, 0); bwd(16, T->dcat0, 128); T->convt[3].dx = cur;
    bwd(15, nullptr, 0); bwd(14, T->dcat1, 256); T->convt[2].dx = cur;
    bwd(13, nullptr, 0); bwd(12, T->dcat2, 512); T->convt[1].dx = cur;
    bwd(11, nullptr, 0); bwd(10, T->dcat3, 1024); T->convt[0].dx = cur;
    bwd(9, nullptr, 0); bwd(8, nullptr, 0); swap();   // pool backward writes `other`
    bwd(7, nullptr, 0); bwd(6, nullptr, 0); swap();
    bwd(5, nullptr, 0); bwd(4, nullptr, 0); swap();
    bwd(3, nullptr, 0); bwd(2, nullptr, 0); swap();
    bwd(1, nullptr, 0); bwd(0, nullptr, 0);

    // ---- launches
    for (int i = 1; i < 18; ++i) {
        ConvL& L = C[i];
        int r = conv_build(L.fwd, L.x, T->B, L.H, L.H, L.x_stride, 0, L.Cin, T->W16 + L.w_off, T->zero_bias, L.Cout, 9 * L.Cin, 3, 1,
                           EPI_STORE, ctx->use_vr);
        if (!r) r = conv_set_store(L.fwd, L.z, L.Cout, 0, 0, nullptr, 0);
        if (!r) r = conv_build(L.dgrad, L.dz, T->B, L.H, L.H, L.Cout, 0, L.Cout, T->W16T + L.w_off, T->zero_bias, L.Cin, 9 * L.Cout, 3, 1,
                               EPI_STORE, ctx->use_vr);
        if (!r) r = conv_set_store(L.dgrad, L.dx, L.dx_stride, 0, 0, nullptr, 0);
        if (r) return fail(ctx, -5, "launch plan of %s failed: %d", L.conv.c_str(), r);
        if (build_conv_wgrad(ctx, T, L)) return -5;
    }
    for (auto& L : T->convt) {
        int r = conv_build(L.fwd, L.x, T->B, L.H, L.H, L.Cin, 0, L.Cin, T->W16 + L.w_off, T->P + L.b_off, 4 * L.Cout, L.Cin, 1, 1, EPI_CONVT,
                           false);
        if (!r) r = conv_set_store(L.fwd, L.cat, 2 * L.Cout, L.Cout, 0, nullptr, 0);
        L.fwd.p.convt_cout = L.Cout;
        if (!r) r = conv_build_k2s2(L.dgrad, L.dcat, T->B, L.H, L.H, 2 * L.Cout, L.Cout, L.Cout, T->W16T + L.w_off, T->zero_bias, L.Cin,
                                    4 * L.Cout);
        if (!r) r = conv_set_store(L.dgrad, L.dx, L.Cin, 0, 0, nullptr, 0);
        if (r) return fail(ctx, -5, "launch plan of %s failed: %d", L.name.c_str(), r);
        if (build_convt_wgrad(ctx, T, L)) return -5;
    }
    return 0;
}

int cvb_train_forward_backward(cvb_ctx* ctx, const float* img, const float* mask, float* loss, void* stream) {
    if (!ctx || !img || !mask || !loss) return -1;
    cvb_trainer* T = ctx->trainer;
    if (!T) return fail(ctx, -7, "no trainer (call cvb_train_create)");
    CVB_ON_DEVICE(ctx);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto& C = T->conv;
    const int B = T->B;
    CK(cudaMemsetAsync(T->G, 0, T->n_params * sizeof(float), s));
    CK(cudaMemsetAsync(T->sums, 0, T->sums_count * sizeof(double), s));
    if (prepare_weights(ctx, T, s)) return -2;
    // ---- forward (unet_model.py:25-36 in training mode)
    if (conv_forward(ctx, T, C[0], img, s) || conv_forward(ctx, T, C[1], img, s)) return -2;
    __half* cats[4] = {T->cat0, T->cat1, T->cat2, T->cat3};
    __half* pools[4] = {T->p1, T->p2, T->p3, T->p4};
    for (int d = 0; d < 4; ++d) {
        const int H = 256 >> d, Cc = 64 << d;
        LAUNCH(launch_maxpool2(cats[d], pools[d], B, H, H, Cc, 2 * Cc, s));
        if (conv_forward(ctx, T, C[2 + 2 * d], img, s) || conv_forward(ctx, T, C[3 + 2 * d], img, s)) return -2;
    }
    for (int u = 0; u < 4; ++u) {
        LAUNCH((T->convt[u].fwd.pdl = 0, conv_launch(T->convt[u].fwd, B, ctx->sm_count, s)));
        if (conv_forward(ctx, T, C[10 + 2 * u], img, s) || conv_forward(ctx, T, C[11 + 2 * u], img, s)) return -2;
    }
    const long long P = static_cast<long long>(B) * 65536;
    LAUNCH(launch_outc_fwd(T->t0c, T->P + T->outc_w_off, T->P + T->outc_b_off, T->logits, P, s));
    // ---- loss + backward
    LAUNCH(launch_loss(T->logits, mask, T->lsums, T->t0c, T->P + T->outc_w_off, C[17].dy, T->G + T->outc_w_off, T->G + T->outc_b_off,
                       loss, B, 65536, T->cfg.loss_scale, s));
    ctx->launches++;
    for (int u = 3; u >= 0; --u) {
        if (conv_backward(ctx, T, C[11 + 2 * u], img, s) || conv_backward(ctx, T, C[10 + 2 * u], img, s)) return -2;
        if (convt_backward(ctx, T, T->convt[u], s)) return -2;
        if (u == 1) CK(cudaEventRecord(T->bucket_ev[0], s));   // up4, up3, up2 and the head are final
        if (u == 0) CK(cudaEventRecord(T->bucket_ev[1], s));   // up1
    }
    __half* dcats[4] = {T->dcat0, T->dcat1, T->dcat2, T->dcat3};
    for (int d = 3; d >= 0; --d) {
        if (conv_backward(ctx, T, C[3 + 2 * d], img, s) || conv_backward(ctx, T, C[2 + 2 * d], img, s)) return -2;
        if (d == 3) CK(cudaEventRecord(T->bucket_ev[2], s));   // down4
        const int H = 256 >> d, Cc = 64 << d;
        // gradient of the skip tensor x_{d+1}: pooled path (dx of the conv just processed) + concat path (lower half of dcat)
        LAUNCH(launch_pool_bwd_add(cats[d], 2 * Cc, dcats[d], 2 * Cc, C[2 + 2 * d].dx, C[1 + 2 * d].dy, B, H, H, Cc, s));
    }
    if (conv_backward(ctx, T, C[1], img, s) || conv_backward(ctx, T, C[0], img, s)) return -2;
    CK(cudaEventRecord(T->bucket_ev[3], s));                   // down3 .. inc
    return 0;
}

int cvb_train_buckets(cvb_ctx* ctx, int64_t* lo, int64_t* hi, int capacity) {
    if (!ctx || !ctx->trainer) return -1;
    const cvb_trainer* T = ctx->trainer;
    for (int b = 0; b < cvb_trainer::kBuckets && b < capacity; ++b) {
        if (lo) lo[b] = static_cast<int64_t>(T->bucket_lo[b]);
        if (hi) hi[b] = static_cast<int64_t>(T->bucket_hi[b]);
    }
    return cvb_trainer::kBuckets;
}

int cvb_train_bucket_wait(cvb_ctx* ctx, int bucket, void* stream) {
    if (!ctx || !ctx->trainer || bucket < 0 || bucket >= cvb_trainer::kBuckets) return -1;
    CVB_ON_DEVICE(ctx);
    CK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), ctx->trainer->bucket_ev[bucket], 0));
    return 0;
}

int cvb_train_grads(cvb_ctx* ctx, float** grads, int64_t* count) {
    if (!ctx || !ctx->trainer || !grads || !count) return -1;
    *grads = ctx->trainer->G;
    *count = static_cast<int64_t>(ctx->trainer->n_params);
    return 0;
}

int cvb_train_optimizer_step(cvb_ctx* ctx, float lr, float grad_scale, void* stream) {
    if (!ctx) return -1;
    cvb_trainer* T = ctx->trainer;
    if (!T) return fail(ctx, -7, "no trainer (call cvb_train_create)");
    CVB_ON_DEVICE(ctx);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CK(launch_optimizer(T->P, T->G, T->SQ, T->MB, static_cast<long long>(T->n_params), T->norm, grad_scale, T->cfg.max_grad_norm, lr,
                        T->cfg.alpha, T->cfg.eps, T->cfg.weight_decay, T->cfg.momentum, ctx->sm_count, s));
    ctx->launches += 2;
    T->steps++;
    return 0;
}

int cvb_train_step(cvb_ctx* ctx, const float* img, const float* mask, float lr, float* loss, void* stream) {
    const int rc = cvb_train_forward_backward(ctx, img, mask, loss, stream);
    if (rc) return rc;
    return cvb_train_optimizer_step(ctx, lr, 1.0f, stream);
}

// what: 0 parameters (+ BatchNorm running statistics), 1 gradients.  out[i].data must point to writable host memory of
// the tensor's torch shape; unknown names are reported as an error.
int cvb_train_export(cvb_ctx* ctx, int what, const cvb_tensor* out, int n) {
    if (!ctx || !out) return -1;
    cvb_trainer* T = ctx->trainer;
    if (!T) return fail(ctx, -7, "no trainer (call cvb_train_create)");
    CVB_ON_DEVICE(ctx);
    CK(cudaDeviceSynchronize());
    std::vector<float> hp(T->n_params), hrm(T->bn_channels), hrv(T->bn_channels);
    CK(cudaMemcpy(hp.data(), what == 0 ? T->P : T->G, T->n_params * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hrm.data(), T->run_mean, T->bn_channels * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hrv.data(), T->run_var, T->bn_channels * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        const std::string name = out[i].name;
        float* dst = const_cast<float*>(out[i].data);
        bool done = false;
        for (auto& L : T->conv) {
            if (name == L.conv + ".weight") {
                if (numel_t(&out[i]) != 9LL * L.Cin * L.Cout) return fail(ctx, -4, "bad shape for '%s'", name.c_str());
                for (int co = 0; co < L.Cout; ++co)
                    for (int ci = 0; ci < L.Cin; ++ci)
                        for (int t = 0; t < 9; ++t)
                            dst[(static_cast<size_t>(co) * L.Cin + ci) * 9 + t] = hp[L.w_off + (static_cast<size_t>(co) * 9 + t) * L.Cin + ci];
                done = true;
            } else if (name == L.bn + ".weight" || name == L.bn + ".bias" || name == L.bn + ".running_mean" || name == L.bn + ".running_var") {
                if (numel_t(&out[i]) != L.Cout) return fail(ctx, -4, "bad shape for '%s'", name.c_str());
                const float* src = name == L.bn + ".weight" ? &hp[L.g_off]
                                   : name == L.bn + ".bias" ? &hp[L.b_off]
                                   : name == L.bn + ".running_mean" ? &hrm[L.bn_off] : &hrv[L.bn_off];
                memcpy(dst, src, L.Cout * sizeof(float));
                done = true;
            }
            if (done) break;
        }
        for (auto& L : T->convt) {
            if (done) break;
            if (name == L.name + ".weight") {
                if (numel_t(&out[i]) != 4LL * L.Cin * L.Cout) return fail(ctx, -4, "bad shape for '%s'", name.c_str());
                for (int ci = 0; ci < L.Cin; ++ci)
                    for (int co = 0; co < L.Cout; ++co)
                        for (int q = 0; q < 4; ++q)
                            dst[(static_cast<size_t>(ci) * L.Cout + co) * 4 + q] = hp[L.w_off + (static_cast<size_t>(q) * L.Cout + co) * L.Cin + ci];
                done = true;
            } else if (name == L.name + ".bias") {
                if (numel_t(&out[i]) != L.Cout) return fail(ctx, -4, "bad shape for '%s'", name.c_str());
                memcpy(dst, &hp[L.b_off], L.Cout * sizeof(float));
                done = true;
            }
        }
        if (!done && name == "outc.conv.weight") { memcpy(dst, &hp[T->outc_w_off], 64 * sizeof(float)); done = true; }
        if (!done && name == "outc.conv.bias") { dst[0] = hp[T->outc_b_off]; done = true; }
        if (!done) return fail(ctx, -4, "'%s' is not a trainable tensor of the UNet", name.c_str());
    }
    return 0;
}

// Building block for parity tests: weight gradient of a 3x3 / pad 1 convolution through the tcgen05 kernel.
// dz fp16 [N,H,W,Cout] dense, x fp16 [N,H,W,Cin] dense -> dw fp32 [Cout][9][Cin] (overwritten), dw = scale * dz (*) x.
int cvb_wgrad3x3_f16(cvb_ctx* ctx, const void* dz, const void* x, int N, int H, int W, int Cout, int Cin, float scale, float* dw,
                     void* stream) {
    if (!ctx || !dz || !x || !dw) return -1;
    if (H != W || Cin % 64 || Cout % 64) return fail(ctx, -5, "cvb_wgrad3x3_f16: square images and channel multiples of 64 only");
    CVB_ON_DEVICE(ctx);
    if (wgrad_configure() != cudaSuccess) return fail(ctx, -2, "wgrad kernel attribute setup failed");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cvb_trainer tmp;
    tmp.B = N;
    tmp.cfg.loss_scale = 1.0f / scale;
    tmp.G = dw;
    WgTile* table = nullptr;
    CK(cudaMalloc(&table, kTileCapacity * sizeof(WgTile)));
    tmp.tiles = table;
    ConvL L;
    L.conv = "test";
    L.Cin = Cin; L.Cout = Cout; L.H = H;
    L.w_off = 0;
    L.dz = const_cast<__half*>(static_cast<const __half*>(dz));
    L.x = static_cast<const __half*>(x);
    L.x_stride = Cin;
    int rc = build_conv_wgrad(ctx, &tmp, L);
    if (!rc) {
        cudaMemsetAsync(dw, 0, sizeof(float) * 9 * Cin * Cout, s);
        cudaError_t e = wgrad_launch(L.wgrad, ctx->sm_count, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail(ctx, -2, "wgrad launch failed: %s", cudaGetErrorString(e));
        ctx->launches++;
    }
    cudaFree(table);
    return rc;
}

}  // extern "C"

void cvb_trainer_free(cvb_trainer* t) {
    if (!t) return;
    for (cudaEvent_t e : t->bucket_ev)
        if (e) cudaEventDestroy(e);
    delete t;
}
