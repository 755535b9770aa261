// JPEG decode front-end of the image->FEN path.
//
// Replaces cv2.imread / cv2.imdecode(buf, IMREAD_COLOR) in front of ChessVision.process_image
// (scripts/eval/evaluate.py:147, app/computeroot/cv_endpoint.py:151-153), i.e. OpenCV's bundled libjpeg-turbo with its
// defaults (JDCT_ISLOW, fancy chroma upsampling, separate colour conversion, BGR output) — bit for bit.
//
// Split by what is sequential and what is not:
//   host   marker parsing + Huffman decoding (jdhuff.c): inherently serial per image, one image per host thread; the
//          result is the quantised coefficients, int16 [blocks][64] in natural order (1.5 * H * W * 2 bytes per image)
//   device k_jpeg_idct   dequantisation + 8x8 inverse DCT (jidctint.c jpeg_idct_islow, 32-bit integer arithmetic,
//                        CONST_BITS 13 / PASS1_BITS 2) + masked range limit -> Y, Cb, Cr planes
//          k_jpeg_color  h2v2 fancy upsampling of Cb/Cr (jdsample.c: 3/4 near + 1/4 far per axis, rounding 8 / 7
//                        alternating by column, edges replicated) + YCbCr->RGB with libjpeg's 16-bit fixed-point
//                        constants (jdcolor.c) -> u8 [N,H,W,3] BGR, the layout every other kernel of the path reads
// Both kernels are HBM-bound byte/integer work (3 B read + 3 B written per pixel in total).
// Supported subset (what the reference's data needs; anything else is error -6, never a silent fallback): baseline SOF0,
// 8-bit, 3 components, 4:2:0, dimensions multiples of 16, restart intervals, EXIF orientation absent or 1.
#include "ctx.h"

#include <string.h>

#include <stdlib.h>

#include <atomic>
#include <functional>
#include <thread>
#include <vector>

namespace cvb {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HuffTable {
    bool present = false;
    int32_t maxcode[18];   // largest code of each length (-1: none)
    int32_t valoff[17];    // symbol index of the first code of each length minus that code
    uint8_t symbols[256];
    uint16_t look[256];    // 8-bit lookahead: (length << 8) | symbol, 0 = longer than 8 bits
    // AC tables: 10-bit lookahead that resolves code AND value bits at once when both fit:
    // (coefficient << 16) | (run << 8) | total bits, 0 = take the general path
    int32_t fast_ac[1024];
};

struct JpegHeader {
    int h = 0, w = 0, restart = 0;
    size_t scan = 0;
    uint16_t qt[3][64];      // natural order
    HuffTable dc[4], ac[4];
    uint16_t qsrc[4][64];
    bool qpresent[4] = {false, false, false, false};
    int comp_q[3], comp_dc[3], comp_ac[3], comp_id[3];
};

// Returns false when the code-length counts do not describe a prefix code (over-subscribed: some code of length `len`
// would not fit in `len` bits — jdhuff.c's jpeg_make_d_derived_tbl rejects the same streams with JERR_BAD_HUFF_TABLE).
// Such counts come from corrupt or hostile files and must never index the lookahead table.
bool build_table(HuffTable& t, const uint8_t* counts, const uint8_t* symbols, int n) {
    t.present = false;
    if (n < 0 || n > 256) return false;
    memcpy(t.symbols, symbols, n);
    memset(t.look, 0, sizeof t.look);
    int code = 0, k = 0;
    for (int len = 1; len <= 16; ++len) {
        t.valoff[len] = k - code;
        if (code + counts[len - 1] > (1 << len)) return false;
        for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
            if (len <= 8) {
                const int lo = code << (8 - len);
                for (int f = 0; f < (1 << (8 - len)); ++f) t.look[lo + f] = static_cast<uint16_t>((len << 8) | symbols[k]);
            }
        }
        t.maxcode[len] = counts[len - 1] ? code - 1 : -1;
        code <<= 1;
    }
    t.maxcode[17] = 0x7fffffff;
    t.present = true;
    for (int i = 0; i < 1024; ++i) {
        t.fast_ac[i] = 0;
        const uint32_t e = t.look[i >> 2];
        if (!e) continue;
        const int len = e >> 8, sym = e & 255, run = sym >> 4, mag = sym & 15;
        if (mag == 0 || len + mag > 10) continue;
        int v = (i >> (10 - len - mag)) & ((1 << mag) - 1);
        if (v < (1 << (mag - 1))) v += 1 - (1 << mag);
        t.fast_ac[i] = (v * 65536) | (run << 8) | (len + mag);
    }
    return true;
}

int exif_orientation(const uint8_t* t, size_t n) {
    if (n < 8) return 0;
    const bool le = t[0] == 'I';
    auto u16 = [&](size_t o) { return le ? t[o] | (t[o + 1] << 8) : (t[o] << 8) | t[o + 1]; };
    auto u32 = [&](size_t o) { return le ? u16(o) | (u16(o + 2) << 16) : (u16(o) << 16) | u16(o + 2); };
    const size_t ifd = static_cast<size_t>(u32(4));
    if (ifd + 2 > n) return 0;
    const int cnt = u16(ifd);
    for (int k = 0; k < cnt; ++k) {
        const size_t e = ifd + 2 + 12 * static_cast<size_t>(k);
        if (e + 12 <= n && u16(e) == 0x0112) return u16(e + 8);
    }
    return 0;
}

// 0 ok, -6 unsupported / corrupt (msg says why)
int parse_header(const uint8_t* d, size_t n, JpegHeader& H, const char** msg) {
    auto bad = [&](const char* m) { *msg = m; return -6; };
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return bad("not a JPEG stream");
    size_t i = 2;
    bool sof = false;
    while (i + 4 <= n) {
        if (d[i] != 0xFF) return bad("corrupt JPEG: marker expected");
        const int m = d[i + 1];
        if (m == 0xFF) { ++i; continue; }
        const size_t len = (static_cast<size_t>(d[i + 2]) << 8) | d[i + 3];
        if (len < 2 || i + 2 + len > n) return bad("corrupt JPEG: segment exceeds the buffer");
        const uint8_t* s = d + i + 4;
        const size_t sl = len - 2;
        if (m == 0xDB) {
            for (size_t j = 0; j + 65 <= sl; j += 65) {
                if (s[j] >> 4) return bad("16-bit quantisation tables are not supported");
                const int tq = s[j] & 3;
                for (int k = 0; k < 64; ++k) H.qsrc[tq][kZigzag[k]] = s[j + 1 + k];
                H.qpresent[tq] = true;
            }
        } else if (m == 0xC4) {
            size_t j = 0;
            while (j + 17 <= sl) {
                const int tc = s[j] >> 4, th = s[j] & 3;
                int cnt = 0;
                for (int k = 0; k < 16; ++k) cnt += s[j + 1 + k];
                if (cnt > 256 || j + 17 + cnt > sl) return bad("corrupt JPEG: Huffman table");
                if (tc > 1) return bad("corrupt JPEG: Huffman table class");
                if (!build_table(tc ? H.ac[th] : H.dc[th], s + j + 1, s + j + 17, cnt)) return bad("corrupt JPEG: Huffman code lengths are over-subscribed");
                j += 17 + cnt;
            }
        } else if (m == 0xC0) {
            if (sl < 15 || s[0] != 8 || s[5] != 3) return bad("only 8-bit three-component baseline JPEGs are supported");
            H.h = (s[1] << 8) | s[2];
            H.w = (s[3] << 8) | s[4];
            for (int k = 0; k < 3; ++k) {
                H.comp_id[k] = s[6 + 3 * k];
                const int hv = s[7 + 3 * k];
                if (hv != (k == 0 ? 0x22 : 0x11)) return bad("only 4:2:0 chroma subsampling is supported");
                H.comp_q[k] = s[8 + 3 * k] & 3;
            }
            if (H.h % 16 || H.w % 16 || H.h <= 0 || H.w <= 0) return bad("image dimensions must be multiples of 16");
            sof = true;
        } else if ((m >= 0xC1 && m <= 0xCF) && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return bad("only baseline (SOF0) JPEGs are supported");
        } else if (m == 0xDD) {
            if (sl < 2) return bad("corrupt JPEG: restart interval segment");
            H.restart = (s[0] << 8) | s[1];
        } else if (m == 0xE1 && sl > 6 && memcmp(s, "Exif\0\0", 6) == 0) {
            const int o = exif_orientation(s + 6, sl - 6);
            if (o != 0 && o != 1) return bad("EXIF orientation other than 1 is not supported");
        } else if (m == 0xDA) {
            if (!sof || sl < 10 || s[0] != 3) return bad("corrupt JPEG: scan header");
            for (int k = 0; k < 3; ++k) {
                int c = -1;
                for (int q = 0; q < 3; ++q)
                    if (H.comp_id[q] == s[1 + 2 * k]) c = q;
                if (c != k) return bad("interleaved scan with components in frame order expected");
                H.comp_dc[c] = s[2 + 2 * k] >> 4;
                H.comp_ac[c] = s[2 + 2 * k] & 3;
                if (!H.dc[H.comp_dc[c] & 3].present || !H.ac[H.comp_ac[c]].present || !H.qpresent[H.comp_q[c]])
                    return bad("corrupt JPEG: missing table");
                memcpy(H.qt[c], H.qsrc[H.comp_q[c]], sizeof H.qt[c]);
            }
            H.scan = i + 2 + len;
            return 0;
        }
        i += 2 + len;
    }
    return bad("corrupt JPEG: no scan");
}

struct BitReader {
    const uint8_t* d;
    size_t n, p;
    uint64_t acc = 0;
    int bits = 0;
    // top up to more than 56 bits; afterwards a Huffman code (<= 16 bits) plus its value bits (<= 15) need no further check
    inline void refill() {
        if (bits <= 32 && p + 4 <= n) {   // four bytes at once when none of them is 0xFF
            const uint32_t w = (static_cast<uint32_t>(d[p]) << 24) | (static_cast<uint32_t>(d[p + 1]) << 16) | (static_cast<uint32_t>(d[p + 2]) << 8) | d[p + 3];
            if ((((w & 0x7F7F7F7Fu) + 0x01010101u) & w & 0x80808080u) == 0) {   // no byte equals 0xFF
                acc = (acc << 32) | w;
                bits += 32;
                p += 4;
                return;
            }
        }
        while (bits <= 56) {
            uint32_t b = 0;
            if (p < n) {
                b = d[p];
                if (b == 0xFF) {
                    const uint32_t nx = p + 1 < n ? d[p + 1] : 0xD9;
                    if (nx == 0) p += 2;
                    else b = 0;   // a marker: feed zeros (jdhuff.c does the same until the MCU ends)
                } else {
                    ++p;
                }
            }
            acc = (acc << 8) | b;
            bits += 8;
        }
    }
    inline uint32_t peek(int k) const { return static_cast<uint32_t>(acc >> (bits - k)) & ((1u << k) - 1); }
    inline void skip(int k) { bits -= k; }
    inline uint32_t get(int k) {
        const uint32_t v = peek(k);
        bits -= k;
        return v;
    }
    void restart() {
        acc = 0;
        bits = 0;
        while (p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7)) ++p;
        p += 2;
    }
};

// caller guarantees at least 32 buffered bits
inline int decode_symbol(BitReader& br, const HuffTable& t) {
    const uint32_t look = t.look[br.peek(8)];
    if (look) {
        br.skip(look >> 8);
        return look & 255;
    }
    int code = static_cast<int>(br.peek(9));
    for (int len = 9; len <= 16; ++len) {
        if (code <= t.maxcode[len]) {
            br.skip(len);
            return t.symbols[(t.valoff[len] + code) & 255];
        }
        code = static_cast<int>(br.peek(len + 1));
    }
    return -1;
}

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

// coef: int16 [Y blocks (H/8 x W/8) | Cb blocks | Cr blocks][64], zero-initialised by the caller.  0 ok / -6 corrupt.
int decode_scan(const uint8_t* d, size_t n, const JpegHeader& H, int16_t* coef) {
    const int mh = H.h / 16, mw = H.w / 16;
    int16_t* plane[3] = {coef, coef + static_cast<size_t>(4) * mh * mw * 64, coef + static_cast<size_t>(5) * mh * mw * 64};
    BitReader br{d, n, H.scan};
    int pred[3] = {0, 0, 0};
    for (int mcu = 0; mcu < mh * mw; ++mcu) {
        if (H.restart && mcu && mcu % H.restart == 0) {
            br.restart();
            pred[0] = pred[1] = pred[2] = 0;
        }
        const int my = mcu / mw, mx = mcu - my * mw;
        for (int b = 0; b < 6; ++b) {
            const int c = b < 4 ? 0 : b - 3;
            int16_t* blk = c == 0 ? plane[0] + (static_cast<size_t>(2 * my + (b >> 1)) * (2 * mw) + 2 * mx + (b & 1)) * 64
                                  : plane[c] + (static_cast<size_t>(my) * mw + mx) * 64;
            const HuffTable& dct = H.dc[H.comp_dc[c] & 3];
            const HuffTable& act = H.ac[H.comp_ac[c]];
            br.refill();
            int s = decode_symbol(br, dct);
            if (s < 0 || s > 15) return -6;
            if (s) pred[c] += extend(static_cast<int>(br.get(s)), s);
            blk[0] = static_cast<int16_t>(pred[c]);
            for (int k = 1; k < 64;) {
                if (br.bits < 32) br.refill();
                const int32_t f = act.fast_ac[br.peek(10)];
                if (f) {
                    k += (f >> 8) & 15;
                    if (k > 63) return -6;
                    br.skip(f & 255);
                    blk[kZigzag[k]] = static_cast<int16_t>(f >> 16);
                    ++k;
                    continue;
                }
                const int rs = decode_symbol(br, act);
                if (rs < 0) return -6;
                const int r = rs >> 4;
                s = rs & 15;
                if (s == 0) {
                    if (r != 15) break;
                    k += 16;
                    continue;
                }
                k += r;
                if (k > 63) return -6;
                blk[kZigzag[k]] = static_cast<int16_t>(extend(static_cast<int>(br.get(s)), s));
                ++k;
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- device
constexpr int F_0_298 = 2446, F_0_390 = 3196, F_0_541 = 4433, F_0_765 = 6270, F_0_899 = 7373, F_1_175 = 9633;
constexpr int F_1_501 = 12299, F_1_847 = 15137, F_1_961 = 16069, F_2_053 = 16819, F_2_562 = 20995, F_3_072 = 25172;

// one 1-D pass of jpeg_idct_islow on 8 values (32-bit integers like the C code); DESCALE by `shift`
__device__ __forceinline__ void idct8(const int (&v)[8], int shift, int (&o)[8]) {
    int z2 = v[2], z3 = v[6];
    int z1 = (z2 + z3) * F_0_541;
    int tmp2 = z1 + z3 * (-F_1_847);
    int tmp3 = z1 + z2 * F_0_765;
    z2 = v[0];
    z3 = v[4];
    int tmp0 = (z2 + z3) * 8192;
    int tmp1 = (z2 - z3) * 8192;
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = v[7];
    tmp1 = v[5];
    tmp2 = v[3];
    tmp3 = v[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * F_1_175;
    tmp0 *= F_0_298;
    tmp1 *= F_2_053;
    tmp2 *= F_3_072;
    tmp3 *= F_1_501;
    z1 *= -F_0_899;
    z2 *= -F_2_562;
    z3 = z3 * (-F_1_961) + z5;
    z4 = z4 * (-F_0_390) + z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    const int r = 1 << (shift - 1);
    o[0] = (tmp10 + tmp3 + r) >> shift;
    o[7] = (tmp10 - tmp3 + r) >> shift;
    o[1] = (tmp11 + tmp2 + r) >> shift;
    o[6] = (tmp11 - tmp2 + r) >> shift;
    o[2] = (tmp12 + tmp1 + r) >> shift;
    o[5] = (tmp12 - tmp1 + r) >> shift;
    o[3] = (tmp13 + tmp0 + r) >> shift;
    o[4] = (tmp13 - tmp0 + r) >> shift;
}

__device__ __forceinline__ uint32_t range_limit_masked(int x) {   // range_limit[x & RANGE_MASK], table centred on 128
    const int i = x & 1023;
    return i < 128 ? i + 128 : (i < 512 ? 255 : (i < 896 ? 0 : i - 896));
}

// 8 threads per 8x8 block (32 blocks per CTA of 256 threads): thread t transforms column t, then row t.
// coef int16 [N][blocks_per_image][64]; qt u16 [N][3][64]; planes u8 [N][H*W (Y) | H*W/4 (Cb) | H*W/4 (Cr)]
__global__ void __launch_bounds__(256) k_jpeg_idct(const int16_t* __restrict__ coef, const uint16_t* __restrict__ qt,
                                                   uint8_t* __restrict__ planes, int H, int W, long long total_blocks) {
    __shared__ int ws[32][8][9];
    const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
    const long long gb = static_cast<long long>(blockIdx.x) * 32 + lb;
    const int per_image = (H / 8) * (W / 8) * 3 / 2;
    const bool live = gb < total_blocks;
    int n = 0, b = 0, comp = 0, bw = W / 8;
    size_t plane_off = 0;
    if (live) {
        n = static_cast<int>(gb / per_image);
        b = static_cast<int>(gb - static_cast<long long>(n) * per_image);
        const int ny = (H / 8) * (W / 8);
        if (b >= ny) {
            comp = 1 + (b - ny) / (ny / 4);
            b = (b - ny) % (ny / 4);
            bw = W / 16;
            plane_off = static_cast<size_t>(H) * W + static_cast<size_t>(comp - 1) * (H / 2) * (W / 2);
        }
        const int16_t* src = coef + gb * 64;
        const uint16_t* q = qt + (static_cast<size_t>(n) * 3 + comp) * 64;
        int v[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = static_cast<int>(__ldg(src + r * 8 + t)) * static_cast<int>(__ldg(q + r * 8 + t));
        idct8(v, 13 - 2, o);   // pass 1: column t
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[lb][r][t] = o[r];
    }
    __syncwarp();
    if (live) {
        int v[8], o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = ws[lb][t][c];
        idct8(v, 13 + 2 + 3, o);   // pass 2: row t
        uint2 px;
        px.x = range_limit_masked(o[0]) | (range_limit_masked(o[1]) << 8) | (range_limit_masked(o[2]) << 16) | (range_limit_masked(o[3]) << 24);
        px.y = range_limit_masked(o[4]) | (range_limit_masked(o[5]) << 8) | (range_limit_masked(o[6]) << 16) | (range_limit_masked(o[7]) << 24);
        const int by = b / bw, bx = b - by * bw;
        uint8_t* dst = planes + static_cast<size_t>(n) * (static_cast<size_t>(H) * W * 3 / 2) + plane_off +
                       (static_cast<size_t>(by) * 8 + t) * (bw * 8) + bx * 8;
        *reinterpret_cast<uint2*>(dst) = px;
    }
}

// one thread per pair of horizontally adjacent output pixels (2c, 2c+1) of row y
__global__ void __launch_bounds__(256) k_jpeg_color(const uint8_t* __restrict__ planes, uint8_t* __restrict__ img, int H, int W,
                                                    long long total_pairs) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total_pairs) return;
    const int cw = W / 2, ch = H / 2;
    const int c = static_cast<int>(i % cw);
    const long long r = i / cw;
    const int y = static_cast<int>(r % H);
    const long long n = r / H;
    const uint8_t* base = planes + n * (static_cast<long long>(H) * W * 3 / 2);
    const uint8_t* Y = base + static_cast<long long>(y) * W + 2 * c;
    const int cy = y >> 1;
    int oy = (y & 1) ? cy + 1 : cy - 1;   // the further chroma row; edge rows are replicated (jdmainct.c context rows)
    oy = oy < 0 ? 0 : (oy >= ch ? ch - 1 : oy);
    const int cl = c > 0 ? c - 1 : 0, cr = c + 1 < cw ? c + 1 : cw - 1;
    int up[2][2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint8_t* P = base + static_cast<long long>(H) * W + static_cast<long long>(k) * ch * cw;
        const uint8_t* near = P + static_cast<long long>(cy) * cw;
        const uint8_t* far = P + static_cast<long long>(oy) * cw;
        const int cur = 3 * near[c] + far[c], last = 3 * near[cl] + far[cl], next = 3 * near[cr] + far[cr];
        up[k][0] = (3 * cur + last + 8) >> 4;
        up[k][1] = (3 * cur + next + 7) >> 4;
    }
    uint8_t out[6];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int yy = Y[k], xb = up[0][k] - 128, xr = up[1][k] - 128;
        int rr = yy + ((91881 * xr + 32768) >> 16);
        int bb = yy + ((116130 * xb + 32768) >> 16);
        int gg = yy + ((-22554 * xb + 32768 - 46802 * xr) >> 16);
        rr = rr < 0 ? 0 : (rr > 255 ? 255 : rr);
        gg = gg < 0 ? 0 : (gg > 255 ? 255 : gg);
        bb = bb < 0 ? 0 : (bb > 255 ? 255 : bb);
        out[3 * k] = static_cast<uint8_t>(bb);
        out[3 * k + 1] = static_cast<uint8_t>(gg);
        out[3 * k + 2] = static_cast<uint8_t>(rr);
    }
    uint16_t* dst = reinterpret_cast<uint16_t*>(img + (r * W + 2 * c) * 3);   // 6-byte aligned at an even pixel: 2-byte stores
    dst[0] = static_cast<uint16_t>(out[0] | (out[1] << 8));
    dst[1] = static_cast<uint16_t>(out[2] | (out[3] << 8));
    dst[2] = static_cast<uint16_t>(out[4] | (out[5] << 8));
}


// ---------------------------------------------------------------------------------------------------- entropy decoding on the device
// jdhuff.c's decode_mcu for a whole batch: one WARP per image.  The bit stream of an image is inherently sequential, so one
// lane walks it (the same state machine as decode_scan above: 64-bit bit buffer, 0xFF00 unstuffing, 8-bit lookahead tables,
// DC prediction, restart markers) with the six derived tables of the image in shared memory; the other lanes move data: they
// clear and write back the 128-byte coefficient block (one coalesced store per block instead of scattered 2-byte stores)
// and stage the tables.  Throughput comes from the number of images in flight (up to 64 warps per SM), which is why the host
// path stays in use for small batches.
struct DevTable {
    uint16_t look[256];
    int32_t maxcode[18];
    int32_t valoff[17];
    uint8_t symbols[256];
    uint8_t pad[4];
};
static_assert(sizeof(DevTable) == 912, "DevTable layout");
struct DevJpegMeta {
    uint32_t off, len;     // entropy-coded segment inside the packed byte buffer
    int32_t restart, pad;
};
constexpr int kHuffWarps = 8;
constexpr int kHuffSmemPerWarp = 6 * static_cast<int>(sizeof(DevTable)) + 128;

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct DevBits {
    const uint8_t* d;
    uint32_t n, p;
    uint64_t acc;
    int bits;
    __device__ __forceinline__ void refill() {
        while (bits <= 56) {
            uint32_t b = 0;
            if (p < n) {
                b = d[p];
                if (b == 0xFF) {
                    const uint32_t nx = p + 1 < n ? d[p + 1] : 0xD9u;
                    if (nx == 0) p += 2;
                    else b = 0;   // a marker: feed zeros
                } else {
                    ++p;
                }
            }
            acc = (acc << 8) | b;
            bits += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek(int k) const { return static_cast<uint32_t>(acc >> (bits - k)) & ((1u << k) - 1); }
    __device__ __forceinline__ uint32_t get(int k) {
        const uint32_t v = peek(k);
        bits -= k;
        return v;
    }
    __device__ void restart() {
        acc = 0;
        bits = 0;
        while (p + 1 < n && !(d[p] == 0xFF && d[p + 1] >= 0xD0 && d[p + 1] <= 0xD7)) ++p;
        p += 2;
    }
};

__device__ __forceinline__ int dev_decode_symbol(DevBits& br, const DevTable& t) {
    const uint32_t look = t.look[br.peek(8)];
    if (look) {
        br.bits -= static_cast<int>(look >> 8);
        return static_cast<int>(look & 255);
    }
    int code = static_cast<int>(br.peek(9));
    for (int len = 9; len <= 16; ++len) {
        if (code <= t.maxcode[len]) {
            br.bits -= len;
            return t.symbols[(t.valoff[len] + code) & 255];
        }
        code = static_cast<int>(br.peek(len + 1));
    }
    return -1;
}

__global__ void __launch_bounds__(kHuffWarps * 32) k_jpeg_huffman(const uint8_t* __restrict__ bytes, const DevJpegMeta* __restrict__ meta,
                                                                 const DevTable* __restrict__ tables, int16_t* __restrict__ coef,
                                                                 int32_t* __restrict__ err, int n_images, int H, int W) {
    extern __shared__ __align__(16) uint8_t hsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x * kHuffWarps + warp;
    if (img >= n_images) return;
    uint8_t* mine = hsm + warp * kHuffSmemPerWarp;
    DevTable* T = reinterpret_cast<DevTable*>(mine);                       // dc of components 0..2, then ac of components 0..2
    uint32_t* blk = reinterpret_cast<uint32_t*>(mine + 6 * sizeof(DevTable));   // one coefficient block, 64 x int16
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(tables + static_cast<size_t>(img) * 6);
        uint32_t* dst = reinterpret_cast<uint32_t*>(T);
        for (int i = lane; i < 6 * static_cast<int>(sizeof(DevTable)) / 4; i += 32) dst[i] = src[i];
    }
    blk[lane] = 0u;
    __syncwarp();
    const int mh = H / 16, mw = W / 16;
    const size_t per_image = static_cast<size_t>(H) * W * 3 / 2;
    int16_t* plane0 = coef + static_cast<size_t>(img) * per_image;
    int16_t* plane1 = plane0 + static_cast<size_t>(4) * mh * mw * 64;
    int16_t* plane2 = plane0 + static_cast<size_t>(5) * mh * mw * 64;
    DevBits br;
    br.d = bytes + meta[img].off;
    br.n = meta[img].len;
    br.p = 0;
    br.acc = 0;
    br.bits = 0;
    const int restart = meta[img].restart;
    int pred0 = 0, pred1 = 0, pred2 = 0;   // scalars: an array indexed by the component would live in local memory
    int bad = 0;
    for (int mcu = 0; mcu < mh * mw && !bad; ++mcu) {
        if (lane == 0 && restart && mcu && mcu % restart == 0) {
            br.restart();
            pred0 = pred1 = pred2 = 0;
        }
        const int my = mcu / mw, mx = mcu - my * mw;
        for (int b = 0; b < 6 && !bad; ++b) {
            const int c = b < 4 ? 0 : b - 3;
            if (lane == 0) {
                int16_t* out = reinterpret_cast<int16_t*>(blk);
                const DevTable& dct = T[c];
                const DevTable& act = T[3 + c];
                br.refill();
                int s = dev_decode_symbol(br, dct);
                if (s < 0 || s > 15) {
                    bad = 1;
                } else {
                    int diff = 0;
                    if (s) {
                        const int v = static_cast<int>(br.get(s));
                        diff = v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
                    }
                    int dcv;
                    if (c == 0) dcv = (pred0 += diff);
                    else if (c == 1) dcv = (pred1 += diff);
                    else dcv = (pred2 += diff);
                    out[0] = static_cast<int16_t>(dcv);
                    for (int k = 1; k < 64;) {
                        if (br.bits < 32) br.refill();
                        const int rs = dev_decode_symbol(br, act);
                        if (rs < 0) { bad = 1; break; }
                        const int r = rs >> 4;
                        s = rs & 15;
                        if (s == 0) {
                            if (r != 15) break;
                            k += 16;
                            continue;
                        }
                        k += r;
                        if (k > 63) { bad = 1; break; }
                        const int v = static_cast<int>(br.get(s));
                        out[c_zigzag[k]] = static_cast<int16_t>(v < (1 << (s - 1)) ? v - (1 << s) + 1 : v);
                        ++k;
                    }
                }
            }
            __syncwarp();
            bad = __shfl_sync(0xffffffffu, bad, 0);
            int16_t* dstb = c == 0 ? plane0 + (static_cast<size_t>(2 * my + (b >> 1)) * (2 * mw) + 2 * mx + (b & 1)) * 64
                                   : (c == 1 ? plane1 : plane2) + (static_cast<size_t>(my) * mw + mx) * 64;
            reinterpret_cast<uint32_t*>(dstb)[lane] = blk[lane];
            blk[lane] = 0u;
            __syncwarp();
        }
    }
    if (lane == 0) err[img] = bad;
}

}  // namespace
}  // namespace cvb

// Lazily sized staging for chunks of images; owned by the context.  Two pinned host buffers so that the host decodes
// chunk i+1 while chunk i is on its way to the device; the device side is ordered by the stream.
constexpr int kJpegChunk = 128;
struct cvb_jpeg_state {
    int16_t* h_coef[2] = {nullptr, nullptr};
    uint16_t* h_qt[2] = {nullptr, nullptr};
    int16_t* d_coef = nullptr;
    uint16_t* d_qt = nullptr;
    uint8_t* d_planes = nullptr;
    size_t cap_pixels = 0;   // images * H * W the coefficient / plane buffers were sized for
    cudaEvent_t done[2] = {nullptr, nullptr};
    // device entropy decoding: packed entropy-coded segments, per-image metadata, derived tables, error flags
    uint8_t* h_bytes = nullptr;  uint8_t* d_bytes = nullptr;  size_t cap_bytes = 0;
    void* h_meta = nullptr;      void* d_meta = nullptr;
    void* h_tables = nullptr;    void* d_tables = nullptr;
    int32_t* h_err = nullptr;    int32_t* d_err = nullptr;
    int cap_images = 0;
    bool huff_configured = false;
};

void cvb_jpeg_free(cvb_jpeg_state* s) {
    if (!s) return;
    for (int i = 0; i < 2; ++i) {
        if (s->h_coef[i]) cudaFreeHost(s->h_coef[i]);
        if (s->h_qt[i]) cudaFreeHost(s->h_qt[i]);
        if (s->done[i]) cudaEventDestroy(s->done[i]);
    }
    if (s->d_coef) cudaFree(s->d_coef);
    if (s->d_qt) cudaFree(s->d_qt);
    if (s->d_planes) cudaFree(s->d_planes);
    if (s->h_bytes) cudaFreeHost(s->h_bytes);
    if (s->d_bytes) cudaFree(s->d_bytes);
    if (s->h_meta) cudaFreeHost(s->h_meta);
    if (s->d_meta) cudaFree(s->d_meta);
    if (s->h_tables) cudaFreeHost(s->h_tables);
    if (s->d_tables) cudaFree(s->d_tables);
    if (s->h_err) cudaFreeHost(s->h_err);
    if (s->d_err) cudaFree(s->d_err);
    delete s;
}

extern "C" {

int cvb_jpeg_info(const uint8_t* data, int64_t nbytes, int32_t* h, int32_t* w) {
    if (!data || nbytes < 4) return -1;
    cvb::JpegHeader H;
    const char* msg = "";
    const int rc = cvb::parse_header(data, static_cast<size_t>(nbytes), H, &msg);
    if (rc) return rc;
    if (h) *h = H.h;
    if (w) *w = H.w;
    return 0;
}

int cvb_jpeg_coefficients(const uint8_t* data, int64_t nbytes, int16_t* coef, uint16_t* qt) {
    if (!data || !coef || nbytes < 4) return -1;
    cvb::JpegHeader H;
    const char* msg = "";
    int rc = cvb::parse_header(data, static_cast<size_t>(nbytes), H, &msg);
    if (rc) return rc;
    memset(coef, 0, static_cast<size_t>(H.h) * H.w * 3 / 2 * sizeof(int16_t));
    if (qt) memcpy(qt, H.qt, sizeof H.qt);
    return cvb::decode_scan(data, static_cast<size_t>(nbytes), H, coef);
}

// Batches of at least this many images are entropy-decoded on the device.  One warp per image walks ~120 k symbols at ~540
// clocks each (a single dependent instruction chain: ncu, profiles/README.md), i.e. ~50 ms for ANY batch that fits the GPU's
// 9,472 warp slots, against ~1.2 ms per image and host thread: the device wins from about a thousand images per call.
// CVB_JPEG_DEVICE_MIN overrides (1 = always, 0 = never).
static int jpeg_device_min() {
    const char* e = getenv("CVB_JPEG_DEVICE_MIN");
    if (!e) return 1024;
    const int v = atoi(e);
    return v <= 0 ? 0x7fffffff : v;
}

static void run_threads(int n, const std::function<void(int)>& fn) {
    std::atomic<int> next{0};
    auto work = [&]() {
        for (;;) {
            const int i = next.fetch_add(1);
            if (i >= n) return;
            fn(i);
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = static_cast<int>(hw ? hw : 4);
    if (nthreads > n) nthreads = n;
    if (nthreads > 32) nthreads = 32;
    std::vector<std::thread> pool;
    for (int k = 1; k < nthreads; ++k) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
}

// Device path of cvb_decode_jpeg: the host only parses the headers, derives the Huffman tables and packs the entropy-coded
// segments; k_jpeg_huffman produces the coefficients the two existing kernels consume.
static int decode_jpeg_device(cvb_ctx* ctx, const uint8_t* const* data, const int64_t* nbytes, int N, int H, int W, uint8_t* img, cudaStream_t s) {
    constexpr int kChunk = 4096;
    const size_t px = static_cast<size_t>(H) * W;
    const size_t coef_per_image = px * 3 / 2;
    if (!ctx->jpeg) ctx->jpeg = new cvb_jpeg_state();
    cvb_jpeg_state* J = ctx->jpeg;
    const int chunk = N < kChunk ? N : kChunk;
    if (J->cap_pixels < static_cast<size_t>(chunk) * px) {   // coefficient / plane buffers (shared with the host path)
        for (int i = 0; i < 2; ++i) {
            if (J->h_coef[i]) cudaFreeHost(J->h_coef[i]);
            J->h_coef[i] = nullptr;
        }
        if (J->d_coef) cudaFree(J->d_coef);
        if (J->d_planes) cudaFree(J->d_planes);
        J->d_coef = nullptr; J->d_planes = nullptr;
        J->cap_pixels = 0;
        CK(cudaMalloc(&J->d_coef, chunk * coef_per_image * sizeof(int16_t)));
        CK(cudaMalloc(&J->d_planes, chunk * coef_per_image));
        J->cap_pixels = static_cast<size_t>(chunk) * px;
    }
    if (J->cap_images < chunk) {
        if (J->h_meta) cudaFreeHost(J->h_meta);
        if (J->d_meta) cudaFree(J->d_meta);
        if (J->h_tables) cudaFreeHost(J->h_tables);
        if (J->d_tables) cudaFree(J->d_tables);
        if (J->h_err) cudaFreeHost(J->h_err);
        if (J->d_err) cudaFree(J->d_err);
        if (J->d_qt) cudaFree(J->d_qt);
        for (int i = 0; i < 2; ++i) {
            if (J->h_qt[i]) cudaFreeHost(J->h_qt[i]);
            J->h_qt[i] = nullptr;
        }
        J->h_meta = J->d_meta = J->h_tables = J->d_tables = nullptr;
        J->h_err = J->d_err = nullptr;
        J->d_qt = nullptr;
        J->cap_images = 0;
        const int cap = chunk > kJpegChunk ? chunk : kJpegChunk;
        CK(cudaMallocHost(&J->h_meta, cap * sizeof(cvb::DevJpegMeta)));
        CK(cudaMalloc(&J->d_meta, cap * sizeof(cvb::DevJpegMeta)));
        CK(cudaMallocHost(&J->h_tables, static_cast<size_t>(cap) * 6 * sizeof(cvb::DevTable)));
        CK(cudaMalloc(&J->d_tables, static_cast<size_t>(cap) * 6 * sizeof(cvb::DevTable)));
        CK(cudaMallocHost(&J->h_err, cap * sizeof(int32_t)));
        CK(cudaMalloc(&J->d_err, cap * sizeof(int32_t)));
        for (int i = 0; i < 2; ++i) CK(cudaMallocHost(&J->h_qt[i], static_cast<size_t>(cap) * 192 * sizeof(uint16_t)));
        CK(cudaMalloc(&J->d_qt, static_cast<size_t>(cap) * 192 * sizeof(uint16_t)));
        J->cap_images = cap;
    }
    if (!J->huff_configured) {   // per context: function attributes belong to the device the context runs on
        CK(cudaFuncSetAttribute(cvb::k_jpeg_huffman, cudaFuncAttributeMaxDynamicSharedMemorySize, cvb::kHuffWarps * cvb::kHuffSmemPerWarp));
        J->huff_configured = true;
    }
    std::vector<cvb::JpegHeader> hdr(chunk);
    std::vector<size_t> off(chunk + 1);
    for (int base = 0; base < N; base += chunk) {
        const int n = N - base < chunk ? N - base : chunk;
        CK(cudaStreamSynchronize(s));   // the previous chunk's copies have left the (single) pinned staging buffers
        std::atomic<int> err{0};
        const char* err_msg = "";
        run_threads(n, [&](int i) {
            const char* msg = "";
            int rc = cvb::parse_header(data[base + i], static_cast<size_t>(nbytes[base + i]), hdr[i], &msg);
            if (!rc && (hdr[i].h != H || hdr[i].w != W)) { rc = -6; msg = "image dimensions differ from the batch's H x W"; }
            if (rc) { err.store(rc); err_msg = msg; }
        });
        if (err.load()) return fail(ctx, err.load(), "cvb_decode_jpeg: %s", err_msg);
        off[0] = 0;
        for (int i = 0; i < n; ++i) {
            const size_t len = static_cast<size_t>(nbytes[base + i]) - hdr[i].scan;
            off[i + 1] = off[i] + ((len + 15) & ~static_cast<size_t>(15));
        }
        if (off[n] > 0xFFFFFFF0ull) return fail(ctx, -1, "cvb_decode_jpeg: a chunk of compressed data exceeds 4 GB");
        if (J->cap_bytes < off[n]) {
            if (J->h_bytes) cudaFreeHost(J->h_bytes);
            if (J->d_bytes) cudaFree(J->d_bytes);
            J->h_bytes = J->d_bytes = nullptr;
            J->cap_bytes = 0;
            const size_t want = off[n] + off[n] / 4 + 4096;
            CK(cudaMallocHost(&J->h_bytes, want));
            CK(cudaMalloc(&J->d_bytes, want));
            J->cap_bytes = want;
        }
        cvb::DevJpegMeta* meta = static_cast<cvb::DevJpegMeta*>(J->h_meta);
        cvb::DevTable* tabs = static_cast<cvb::DevTable*>(J->h_tables);
        uint16_t* h_qt = J->h_qt[0];
        run_threads(n, [&](int i) {
            const cvb::JpegHeader& hd = hdr[i];
            const size_t len = static_cast<size_t>(nbytes[base + i]) - hd.scan;
            memcpy(J->h_bytes + off[i], data[base + i] + hd.scan, len);
            meta[i].off = static_cast<uint32_t>(off[i]);
            meta[i].len = static_cast<uint32_t>(len);
            meta[i].restart = hd.restart;
            meta[i].pad = 0;
            memcpy(h_qt + static_cast<size_t>(i) * 192, hd.qt, sizeof hd.qt);
            for (int c = 0; c < 3; ++c)
                for (int cls = 0; cls < 2; ++cls) {
                    const cvb::HuffTable& t = cls ? hd.ac[hd.comp_ac[c]] : hd.dc[hd.comp_dc[c] & 3];
                    cvb::DevTable& d = tabs[static_cast<size_t>(i) * 6 + cls * 3 + c];
                    memcpy(d.look, t.look, sizeof d.look);
                    memcpy(d.maxcode, t.maxcode, sizeof d.maxcode);
                    memcpy(d.valoff, t.valoff, sizeof d.valoff);
                    memcpy(d.symbols, t.symbols, sizeof d.symbols);
                    memset(d.pad, 0, sizeof d.pad);
                }
        });
        CK(cudaMemcpyAsync(J->d_bytes, J->h_bytes, off[n], cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(J->d_meta, J->h_meta, n * sizeof(cvb::DevJpegMeta), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(J->d_tables, J->h_tables, static_cast<size_t>(n) * 6 * sizeof(cvb::DevTable), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(J->d_qt, h_qt, static_cast<size_t>(n) * 192 * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        cvb::k_jpeg_huffman<<<(n + cvb::kHuffWarps - 1) / cvb::kHuffWarps, cvb::kHuffWarps * 32, cvb::kHuffWarps * cvb::kHuffSmemPerWarp, s>>>(
            J->d_bytes, static_cast<const cvb::DevJpegMeta*>(J->d_meta), static_cast<const cvb::DevTable*>(J->d_tables), J->d_coef, J->d_err, n, H, W);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(J->h_err, J->d_err, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
        const long long blocks = static_cast<long long>(n) * (H / 8) * (W / 8) * 3 / 2;
        cvb::k_jpeg_idct<<<static_cast<unsigned>((blocks + 31) / 32), 256, 0, s>>>(J->d_coef, J->d_qt, J->d_planes, H, W, blocks);
        CK(cudaGetLastError());
        const long long pairs = static_cast<long long>(n) * H * (W / 2);
        cvb::k_jpeg_color<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, s>>>(J->d_planes, img + static_cast<size_t>(base) * px * 3, H, W, pairs);
        CK(cudaGetLastError());
        ctx->launches += 3;
        CK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i)
            if (J->h_err[i]) return fail(ctx, -6, "cvb_decode_jpeg: corrupt JPEG: entropy-coded data");
    }
    return 0;
}

int cvb_decode_jpeg(cvb_ctx* ctx, const uint8_t* const* data, const int64_t* nbytes, int N, int H, int W, uint8_t* img, void* stream) {
    if (!ctx || !data || !nbytes || !img || N < 0 || H <= 0 || W <= 0 || H % 16 || W % 16) return fail(ctx, -1, "cvb_decode_jpeg: bad argument");
    CVB_ON_DEVICE(ctx);
    if (N == 0) return 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (N >= jpeg_device_min()) return decode_jpeg_device(ctx, data, nbytes, N, H, W, img, s);
    const size_t px = static_cast<size_t>(H) * W;
    const size_t coef_per_image = px * 3 / 2;          // int16 values
    const int chunk = N < kJpegChunk ? N : kJpegChunk;
    if (!ctx->jpeg) ctx->jpeg = new cvb_jpeg_state();
    cvb_jpeg_state* J = ctx->jpeg;
    if (J->cap_pixels < static_cast<size_t>(chunk) * px || !J->h_coef[0]) {   // (the device path keeps no host coefficient buffers)
        for (int i = 0; i < 2; ++i) {
            if (J->h_coef[i]) cudaFreeHost(J->h_coef[i]);
            J->h_coef[i] = nullptr;
        }
        if (J->d_coef) cudaFree(J->d_coef);
        if (J->d_planes) cudaFree(J->d_planes);
        J->d_coef = nullptr; J->d_planes = nullptr;
        J->cap_pixels = 0;
        for (int i = 0; i < 2; ++i) {
            CK(cudaMallocHost(&J->h_coef[i], chunk * coef_per_image * sizeof(int16_t)));
            if (!J->h_qt[i]) CK(cudaMallocHost(&J->h_qt[i], static_cast<size_t>(kJpegChunk) * 192 * sizeof(uint16_t)));
            if (!J->done[i]) CK(cudaEventCreateWithFlags(&J->done[i], cudaEventDisableTiming));
        }
        CK(cudaMalloc(&J->d_coef, chunk * coef_per_image * sizeof(int16_t)));
        if (!J->d_qt) CK(cudaMalloc(&J->d_qt, static_cast<size_t>(kJpegChunk) * 192 * sizeof(uint16_t)));
        CK(cudaMalloc(&J->d_planes, chunk * coef_per_image));
        J->cap_pixels = static_cast<size_t>(chunk) * px;
    }
    const int cap = static_cast<int>(J->cap_pixels / px) < kJpegChunk ? static_cast<int>(J->cap_pixels / px) : kJpegChunk;
    int used[2] = {0, 0};
    for (int off = 0, it = 0; off < N; off += cap, ++it) {
        const int n = N - off < cap ? N - off : cap;
        const int hb = it & 1;
        if (used[hb]) CK(cudaEventSynchronize(J->done[hb]));   // the H2D copy issued two chunks ago must have left this buffer
        int16_t* h_coef = J->h_coef[hb];
        uint16_t* h_qt = J->h_qt[hb];
        std::atomic<int> next{0}, err{0};
        const char* err_msg = "";
        auto work = [&]() {
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n) return;
                cvb::JpegHeader hd;
                const char* msg = "";
                int rc = cvb::parse_header(data[off + i], static_cast<size_t>(nbytes[off + i]), hd, &msg);
                if (!rc && (hd.h != H || hd.w != W)) { rc = -6; msg = "image dimensions differ from the batch's H x W"; }
                if (!rc) {
                    memset(h_coef + i * coef_per_image, 0, coef_per_image * sizeof(int16_t));
                    memcpy(h_qt + static_cast<size_t>(i) * 192, hd.qt, sizeof hd.qt);
                    rc = cvb::decode_scan(data[off + i], static_cast<size_t>(nbytes[off + i]), hd, h_coef + i * coef_per_image);
                    if (rc) msg = "corrupt JPEG: entropy-coded data";
                }
                if (rc) { err.store(rc); err_msg = msg; }
            }
        };
        unsigned hw = std::thread::hardware_concurrency();
        int nthreads = static_cast<int>(hw ? hw : 4);
        if (nthreads > n) nthreads = n;
        if (nthreads > 32) nthreads = 32;
        std::vector<std::thread> pool;
        for (int k = 1; k < nthreads; ++k) pool.emplace_back(work);
        work();
        for (auto& th : pool) th.join();
        if (err.load()) return fail(ctx, err.load(), "cvb_decode_jpeg: %s", err_msg);
        CK(cudaMemcpyAsync(J->d_coef, h_coef, n * coef_per_image * sizeof(int16_t), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(J->d_qt, h_qt, static_cast<size_t>(n) * 192 * sizeof(uint16_t), cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(J->done[hb], s));
        used[hb] = 1;
        const long long blocks = static_cast<long long>(n) * (H / 8) * (W / 8) * 3 / 2;
        cvb::k_jpeg_idct<<<static_cast<unsigned>((blocks + 31) / 32), 256, 0, s>>>(J->d_coef, J->d_qt, J->d_planes, H, W, blocks);
        CK(cudaGetLastError());
        const long long pairs = static_cast<long long>(n) * H * (W / 2);
        cvb::k_jpeg_color<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, s>>>(J->d_planes, img + static_cast<size_t>(off) * px * 3, H, W, pairs);
        CK(cudaGetLastError());
        ctx->launches += 2;
    }
    for (int i = 0; i < 2; ++i)
        if (used[i]) CK(cudaEventSynchronize(J->done[i]));   // the pinned staging is ours, but the caller may free its streams' memory
    return 0;
}

}  // extern "C"
