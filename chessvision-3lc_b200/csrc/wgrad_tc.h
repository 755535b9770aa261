// Host-side description of one tcgen05 weight-gradient launch (internal to the library; UNet training step).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace cvb {

// One 64-channel operand block: 64 consecutive channels (from c0) of tensor view `map`, read at the pixel tile shifted
// by (dy, dx); `off` is the element offset this block contributes to the output address.  map < 0: block absent.
struct WgBlock {
    int16_t map;
    int16_t c0;
    int8_t dy, dx;
    int16_t pad;
    int64_t off;
};

// One accumulator tile: D[128 x 64*NB] = sum over pixels of  A[pixel][m]^T * B[pixel][n]
//   rows    m = 64*i + r  <->  block a[i], channel r          columns n = 64*j + c  <->  block b[j], channel c
//   out[a[i].off + b[j].off + r*row_stride + c*col_stride] += scale * D[m][n]          (fp32 atomics, split over pixels)
struct alignas(16) WgTile {
    WgBlock a[2];
    WgBlock b[4];
};

struct alignas(64) WgParams {
    CUtensorMap maps[6];     // 4-D NHWC views {C, W, H, N}, box {64, tw, th, 1} with tw*th = 64 pixels, 128-byte swizzle
    const WgTile* tiles;
    int n_tiles;
    int splits;              // pixel range of every tile is split `splits` ways (items = n_tiles * splits)
    int p_tiles;             // 64-pixel tiles in the whole pixel grid (N * H * W / 64)
    int tw, th, tiles_w, tiles_h;
    float* out;
    long long row_stride, col_stride;
    float scale;
    uint32_t idesc;
};

struct WgLaunch {
    WgParams p;
    int nb;                  // B blocks per tile (1..4): MMA N = 64*nb
};

// U = row blocks, V = column blocks (see WgTile).  Tiles are formed from pairs of U and groups of `nb` of V; the tile
// table is uploaded to `d_tiles` (capacity checked).  Pixel grid = N x H x W (H*W multiple of 64, W multiple of 16).
int wgrad_build(WgLaunch& L, const std::vector<WgBlock>& U, const std::vector<WgBlock>& V, int nb, WgTile* d_tiles, int tile_capacity,
                int N, int H, int W, float* out, long long row_stride, long long col_stride, float scale, int sm_count);
// 64-pixel box view of an NHWC fp16 tensor (strides in elements)
int wgrad_tmap(CUtensorMap* m, const void* base, int C, int Wv, int Hv, int Nv, int64_t sW, int64_t sH, int64_t sN);
cudaError_t wgrad_launch(const WgLaunch& L, int sm_count, cudaStream_t stream);
cudaError_t wgrad_configure();

}  // namespace cvb
