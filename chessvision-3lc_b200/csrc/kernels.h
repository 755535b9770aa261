// Launchers of the non-tensor-core kernels (internal to the library).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cvb {

// stem.cu
cudaError_t launch_unet_stem(const uint8_t* img, const float* wf, const float* bf, __half* out, int N, int H, int W,
                             int out_c_stride, cudaStream_t s);
cudaError_t launch_maxpool2(const __half* in, __half* out, int N, int H, int W, int C, int in_c_stride, cudaStream_t s);
cudaError_t launch_mask_from_logits(const float* logits, uint8_t* mask, float thr, long long count, cudaStream_t s);
cudaError_t launch_resize_area_half(const uint8_t* img, uint8_t* out, int N, int h, int w, cudaStream_t s);
// general INTER_AREA reduction u8 [N,H,W,3] -> u8 [N,dh,dw,3] from host-built cell tables (api.cu: area_tables); int_area != 0:
// both scale factors are integers and int_area = their product
cudaError_t launch_resize_area(const uint8_t* img, uint8_t* out, int N, int H, int W, int dh, int dw, const int* xofs, const int* xsi,
                               const float* xa, const int* yofs, const int* ysi, const float* ya, int int_area, cudaStream_t s);
// the enlarging case of INTER_AREA (OpenCV's fixed-point bilinear resizer with area-mode weights): xtab / ytab hold
// (source index, weight0, weight1) per destination column / row, xmax = first column whose right tap is outside the row
cudaError_t launch_resize_linear_area(const uint8_t* img, uint8_t* out, int N, int H, int W, int dh, int dw, const int* xtab,
                                      const int* ytab, int xmax, cudaStream_t s);
cudaError_t launch_double2x(const uint8_t* in, uint8_t* out, int N, int h, int w, cudaStream_t s);
cudaError_t configure_resnet_stem();
cudaError_t launch_resnet_stem(const uint8_t* board, const float* wf, const float* bf, __half* out, int n_boards,
                               cudaStream_t s);
cudaError_t launch_head(const __half* feat, const float* fcw, const float* fcb, float* probs, uint8_t* labels,
                        uint8_t* labels_valid, char* fen, int n_boards, int flip, cudaStream_t s);

// stem_tc.cu: the same two first layers on tcgen05 (wsw = swizzled fp16 [64][64] B tile, see pack_stem_tc in api.cu)
cudaError_t configure_stems_tc();
// omap: 4-D view {C, 256, 256, N} of the fp16 NHWC output with box {64, 64, 2, 1} (128-byte swizzle) for the tile store
cudaError_t launch_unet_stem_tc(const uint8_t* img, const void* wsw, const float* bias, const CUtensorMap* omap, int N, int sm_count,
                                cudaStream_t s);
cudaError_t launch_resnet_stem_tc(const uint8_t* board, const void* wsw, const float* bias, __half* out, int n_boards, int sm_count,
                                  cudaStream_t s);

// geometry.cu
constexpr int kQuadMaxPoints = 8192;     // border points of one contour held in shared memory
constexpr int kQuadMaxBorders = 32000;   // borders per mask (int16 labels)
constexpr int kQuadMaxVertices = 2048;   // vertices after the TC89_KCOS reduction
// status codes written per board by the quad kernel
enum QuadStatus : int { QUAD_NONE = 0, QUAD_FOUND = 1, QUAD_OVERFLOW = 2, QUAD_NEED_FULL = 3 /* internal: compact kernel -> full kernel */ };

cudaError_t configure_quad();
cudaError_t configure_warp();
// mask u8 [N,256,256]; quad int32 [N,4,2] (x,y in the 256x256 mask frame, after _rotate_quadrangle);
// found u8 [N]; status int32 [N]; owner_scratch int32 [N, kQuadMaxBorders]
// full_only: run only the full-state kernel (A/B measurements, tests); otherwise the compact kernel runs first and the
// full-state kernel re-runs the boards it flagged.
// A third launch re-runs masks that exceed those capacities with every array in global scratch (`big_scratch`,
// quad_big_scratch_bytes() bytes, zero-initialised once), so QUAD_OVERFLOW is only ever final for more than 4 x 65536
// border points in one contour -- which a 256 x 256 mask cannot produce.
constexpr int kBigSlots = 4;
size_t quad_big_scratch_bytes();
cudaError_t launch_mask_to_quad(const uint8_t* mask, int32_t* quad, uint8_t* found, int32_t* status, int32_t* n_contours,
                                int32_t* owner_scratch, uint8_t* big_scratch, int N, bool full_only, cudaStream_t s);
// quad -> inverse homography (double[9] per board), scale = H_img / 256 applied to both axes (core.py:414-417)
cudaError_t launch_homography(const int32_t* quad, const uint8_t* found, double* minv, int N, float scale, int out_w,
                              int out_h, cudaStream_t s);
// img u8 [N,H,W,3] + minv -> board u8 [N,512,512] (warpPerspective + BGR2GRAY + flip); zero where !found.
// squares (optional): the same pixels as u8 [N,64,64,64] in extract_squares order (core.py:420-439)
cudaError_t launch_warp_board(const uint8_t* img, const double* minv, const uint8_t* found, uint8_t* board, uint8_t* squares, int N,
                              int H, int W, cudaStream_t s);

// utils.extract_perspective (utils.py:115-132): one image u8 [H,W,C] (C = 1 or 3), corners f32 [4][2] (device), any out size;
// minv_scratch: 9 doubles of device scratch
cudaError_t launch_warp_perspective(const uint8_t* img, int H, int W, int C, const float* corners, double* minv_scratch, uint8_t* out, int out_w,
                                    int out_h, cudaStream_t s);

}  // namespace cvb
