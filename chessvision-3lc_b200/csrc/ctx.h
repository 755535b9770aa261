// Internal: the context object behind the C ABI (shared by api.cu and train.cu).
#pragma once
#include "../../include/chessvision_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "conv_tc.h"

struct cvb_trainer;   // train.cu
void cvb_trainer_free(cvb_trainer* t);
struct cvb_cls_trainer;   // train_cls.cu
void cvb_cls_trainer_free(cvb_cls_trainer* t);
struct cvb_jpeg_state;   // jpeg.cu: staging buffers of the JPEG decode front-end
void cvb_jpeg_free(cvb_jpeg_state* s);

using namespace cvb;

constexpr int kStages = 7;
constexpr int kMaxProfileEvents = 8192;

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct ConvWeights {
    __half* w = nullptr;   // [rows][K]
    float* bias = nullptr; // [rows]
    int rows = 0, K = 0;
};

struct cvb_ctx {
    int device = 0;
    int max_batch = 0;        // boards per network chunk (activation workspaces)
    int group_chunks = 8;     // chunks per geometry group: mask->quad / warp run once over up to group_chunks*max_batch boards
    int group = 0;            // = group_chunks * max_batch
    bool quad_full_only = false;   // CVB_QUAD_FULL=1: skip the compact mask->quad kernel (A/B measurements, tests)
    int sm_count = 148;
    bool use_vr = true;   // CVB_NO_VR=1 forces the generic conv kernel everywhere (A/B measurements)
    std::string err;
    int64_t launches = 0;
    std::vector<void*> allocs;

    // ---- UNet
    bool unet_loaded = false;
    float *stem_w = nullptr, *stem_b = nullptr;        // inc.double_conv.0 folded, fp32 [27][64], [64] (CVB_STEM_FP32 A/B path)
    void* stem_wsw = nullptr;                          // same layer as a swizzled fp16 [64][64] tcgen05 B tile
    CUtensorMap stem_omap;                             // TMA store view of t0 for the stem kernel
    bool stem_fp32 = false;                            // CVB_STEM_FP32=1: CUDA-core fp32 stems (A/B measurements only)
    float* outc_w = nullptr;                           // [64]
    float outc_b = 0.f;
    std::vector<ConvWeights> unet_w;                   // 17 conv3x3 + 4 convT, in plan order
    __half *cat0, *t0, *p1, *t1, *cat1, *p2, *t2, *cat2, *p3, *t3, *cat3, *p4, *t4, *x5, *u1, *u2, *u3;
    std::vector<ConvLaunch> unet_plan;                 // indices documented in build_unet_plan
    bool fuse_up4 = true;                              // up3.conv.3 + up4.up as one kernel (CVB_NO_CONVT_FUSE=1: two launches)
    float* ws_logits = nullptr;                        // [B,256,256]
    uint8_t* ws_mask = nullptr;                        // [B,256,256]

    // ---- inputs of other sizes than 512x512 (general INTER_AREA): cell tables for the current (H, W), staging images
    int gs_H = 0, gs_W = 0, gs_int_area = 0;
    int *gs_xofs = nullptr, *gs_xsi = nullptr, *gs_yofs = nullptr, *gs_ysi = nullptr;
    float *gs_xa = nullptr, *gs_ya = nullptr;
    int *gs_lx = nullptr, *gs_ly = nullptr;               // H < 256 or W < 256: tables of the fixed-point bilinear emulation
    int gs_xmax = 0;
    bool gs_linear = false;
    uint8_t *gs_small = nullptr, *gs_big = nullptr;      // [B,256,256,3] resized, [B,512,512,3] replicated (UNet stem input)

    // ---- geometry
    int32_t *ws_quad = nullptr, *ws_status = nullptr, *ws_ncont = nullptr, *ws_owner = nullptr;
    uint8_t* ws_found = nullptr;
    uint8_t* ws_quad_big = nullptr;                    // scratch slots + locks of the large-capacity mask->quad kernel
    double* ws_minv = nullptr;
    uint8_t* ws_board = nullptr;                       // [B,512,512]

    // ---- ResNet-18
    bool resnet_loaded = false;
    float *rstem_w = nullptr, *rstem_b = nullptr;      // conv1+bn1 folded fp32 [49][64], [64]
    void* rstem_wsw = nullptr;                         // conv1+bn1 as a swizzled fp16 [64][64] tcgen05 B tile
    float *fc_w = nullptr, *fc_b = nullptr;            // [13][512], [13]
    std::vector<ConvWeights> res_w;
    __half* rbuf[12] = {nullptr};                      // 3 per resolution level
    std::vector<ConvLaunch> res_plan;
    float* ws_probs = nullptr;
    uint8_t *ws_labels = nullptr, *ws_labels_valid = nullptr;
    char* ws_fen = nullptr;

    // ---- host streaming
    cudaStream_t s_in = nullptr, s_comp = nullptr, s_out = nullptr;
    std::vector<cudaEvent_t> ev_in[2];   // one per chunk of a group: the network of chunk c starts when its images have landed
    cudaEvent_t ev_comp[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    uint8_t* slot_img[2] = {nullptr, nullptr};
    // per-slot device outputs for the host path
    cvb_outputs slot_out[2];

    // ---- UNet training step (train.cu); owned, destroyed with the context
    cvb_trainer* trainer = nullptr;
    cvb_cls_trainer* cls_trainer = nullptr;            // piece-classifier training step (train_cls.cu); owned
    cvb_jpeg_state* jpeg = nullptr;

    // ---- CUDA graphs of single-chunk pipeline passes (the latency path: process_image on one board is ~50 launches)
    struct GraphEntry {
        uint64_t key[16];            // everything the captured launches bake in: pointers, board count, threshold, flip
        cudaGraphExec_t exec = nullptr;
        int64_t launches = 0;        // kernels one replay launches
        uint64_t last_use = 0;
    };
    std::vector<GraphEntry> graphs;
    bool use_graph = true;           // CVB_NO_GRAPH=1 turns it off (A/B measurements)
    uint64_t graph_clock = 0;
    int64_t graph_replays = 0;

    // ---- profiling
    bool profile = false;
    std::vector<cudaEvent_t> pev;
    std::vector<int> pev_stage;
    int pev_used = 0;
    float stage_ms[kStages] = {0};
};

inline int fail(cvb_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(ctx, -2, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Every exported function runs on the context's GPU and leaves the calling thread's current device as it found it
// (a host program may drive several contexts / GPUs, and PyTorch tracks the runtime's current device).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) return;
        if (cur == dev) { ok = true; return; }
        ok = cudaSetDevice(dev) == cudaSuccess;
        if (ok) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define CVB_ON_DEVICE(ctx)                         \
    DeviceGuard device_guard_((ctx)->device);      \
    if (!device_guard_.ok) return fail((ctx), -2, "cudaSetDevice(%d) failed", (ctx)->device)

template <class T>
inline int dalloc(cvb_ctx* ctx, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T));
    if (e != cudaSuccess) return fail(ctx, -3, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    ctx->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return 0;
}

