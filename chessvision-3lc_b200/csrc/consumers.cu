// Per-board consumers of the image->FEN outputs, computed where the outputs already are.
//
// Replaces (reference, numpy / python-chess on the host, one board at a time):
//   n1  compute_model_topk_accuracy, compute_position_accuracy     scripts/eval/evaluate.py:37-52,109-140
//   n4  probability_distribution, mask_completeness,                scripts/process_new_raw/process_pipeline.py:357-467
//       quadrangle_regularity, probability_confidence
// All three are reductions over data the pipeline leaves in HBM (probs 3.3 KB, labels 128 B, logits 256 KB per board),
// HBM-bound: k_quality reads each board's 262,144 B once from HBM and three more times from L2.
#include "ctx.h"

#include <math.h>

namespace cvb {
namespace {

// ---------------------------------------------------------------------------------------------------------------------
// n1.  One block of 64 threads per board, thread <-> square (FEN order a8..h1).
//   rank(sq) = number of classes ranked ahead of the true class in np.argsort(p)[::-1] (evaluate.py:122-136): a class c
//   is ahead of t if p[c] > p[t], or p[c] == p[t] and c > t (stable ascending sort read from the end).
//   topk_hits[n][i] = #squares with rank <= i.   correct[n][0|1] = #squares whose original | validated label equals the
//   true label, with the flipped orientation read back to FEN order (SQUARE_NAMES_FLIPPED is the reversed table).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_eval_metrics(const float* __restrict__ probs, const uint8_t* __restrict__ labels,
                                                     const uint8_t* __restrict__ labels_valid, const uint8_t* __restrict__ truth,
                                                     int flip, int k, int32_t* __restrict__ topk_hits, int32_t* __restrict__ correct) {
    __shared__ int s_cnt[2][16];
    const int n = blockIdx.x, sq = threadIdx.x, warp = sq >> 5, lane = sq & 31;
    const int t = truth[n * 64 + sq];
    const float* p = probs + (static_cast<size_t>(n) * 64 + sq) * 13;
    float v[13];
#pragma unroll
    for (int c = 0; c < 13; ++c) v[c] = __ldg(p + c);
    int rank = 13;
    if (t < 13) {
        float pt = 0.f;
#pragma unroll
        for (int c = 0; c < 13; ++c) pt = c == t ? v[c] : pt;
        rank = 0;
#pragma unroll
        for (int c = 0; c < 13; ++c) rank += (v[c] > pt || (v[c] == pt && c > t)) ? 1 : 0;
    }
    for (int i = 0; i < k && i < 13; ++i) {
        const unsigned m = __ballot_sync(0xffffffffu, rank <= i);
        if (lane == 0) s_cnt[warp][i] = __popc(m);
    }
    const int src = flip ? 63 - sq : sq;
    const unsigned m0 = __ballot_sync(0xffffffffu, labels != nullptr && labels[n * 64 + src] == t);
    const unsigned m1 = __ballot_sync(0xffffffffu, labels_valid != nullptr && labels_valid[n * 64 + src] == t);
    if (lane == 0) {
        s_cnt[warp][13] = __popc(m0);
        s_cnt[warp][14] = __popc(m1);
    }
    __syncthreads();
    if (sq < k && sq < 13) topk_hits[n * k + sq] = s_cnt[0][sq] + s_cnt[1][sq];
    if (sq >= 13 && sq < 15 && correct != nullptr) correct[n * 2 + sq - 13] = s_cnt[0][sq] + s_cnt[1][sq];
}

// ---------------------------------------------------------------------------------------------------------------------
// n4.  One block of 1024 threads per board over its L float32 values (the reference passes the 256x256 logits array).
//   distribution: np.histogram(v, bins=10, range=(0,1)) in float32 arithmetic (numpy keeps the input dtype: index =
//                 int(v * 10), corrected against the float32 edges linspace(0,1,11)), then 1 - H(hist)/log2(10).
//   confidence:   mean(|x - 0.5|) * 2 over the k = int(L * 0.25) largest values: 4-pass radix select of the k-th largest
//                 key, then one sum; elements equal to the threshold contribute (k - #greater) times.
//   regularity:   side / angle spread of the quadrangle, float32 like numpy on the f32[4,1,2] corners.
// scores f64 [N,4] = {regularity, completeness (k_mask_completeness below; NaN unless L = 256*256), distribution, confidence}.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_key(float f) {   // order-preserving map float -> uint32
    const uint32_t b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

__constant__ float c_edges[11] = {0.0f, 0.1f, 0.2f, 0.3f, 0.4f, 0.5f, 0.6f, 0.7f, 0.8f, 0.9f, 1.0f};

__global__ void __launch_bounds__(1024) k_quality(const float* __restrict__ vals, const float* __restrict__ quad, const uint8_t* __restrict__ found,
                                                  int L, double* __restrict__ scores) {
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_bins[10];
    __shared__ uint32_t s_prefix, s_need;
    __shared__ double s_red[32];
    __shared__ unsigned int s_gt[32];
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* v = vals + static_cast<size_t>(n) * L;

    // ---- 10-bin histogram (warp-private counts in registers would need 10 ballots per value; shared atomics are enough)
    if (tid < 10) s_bins[tid] = 0;
    __syncthreads();
    {
        unsigned int local[10];
#pragma unroll
        for (int b = 0; b < 10; ++b) local[b] = 0;
        for (int i = tid; i < L; i += 1024) {
            const float x = __ldg(v + i);
            if (!(x >= 0.0f && x <= 1.0f)) continue;   // outside the range (or NaN): dropped
            int b = static_cast<int>(__fmul_rn(x, 10.0f));
            if (b == 10) b = 9;
            if (x < c_edges[b]) --b;
            if (b != 9 && x >= c_edges[b + 1]) ++b;
#pragma unroll
            for (int q = 0; q < 10; ++q) local[q] += (q == b) ? 1u : 0u;
        }
#pragma unroll
        for (int b = 0; b < 10; ++b) {
            unsigned int c = local[b];
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0 && c) atomicAdd(&s_bins[b], c);
        }
    }
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        for (int b = 0; b < 10; ++b) total += s_bins[b];
        double entropy = 0.0;
        for (int b = 0; b < 10; ++b) {
            const double h = static_cast<double>(s_bins[b]) / total;   // 0/0 -> NaN like numpy when nothing is in range
            entropy -= h * log2(h + 1e-10);
        }
        scores[n * 4 + 2] = 1.0 - entropy / 3.321928094887362;   // -log2(1/10)
    }

    // ---- k-th largest key by radix select, most significant byte first
    // np.sort(flat)[-k:] with k = int(L * 0.25): for L < 4 that is [-0:], i.e. the WHOLE array, not an empty slice
    int k = static_cast<int>(L * 0.25);
    if (k == 0) k = L;
    if (tid == 0) {
        s_prefix = 0;
        s_need = static_cast<uint32_t>(k);
    }
    for (int pass = 0; pass < 4 && k > 0; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const uint32_t himask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < L; i += 1024) {
            const uint32_t key = f32_key(__ldg(v + i));
            if ((key & himask) == prefix) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = s_need;
            int b = 255;
            for (; b > 0; --b) {
                if (s_hist[b] >= need) break;
                need -= s_hist[b];
            }
            s_prefix = prefix | (static_cast<uint32_t>(b) << shift);
            s_need = need;   // elements with this exact prefix still to take
        }
        __syncthreads();
    }
    if (k > 0) {
        const uint32_t thr = s_prefix;   // key of the k-th largest value
        double sum = 0.0;
        unsigned int gt = 0;
        for (int i = tid; i < L; i += 1024) {
            const float x = __ldg(v + i);
            if (f32_key(x) > thr) {
                sum += static_cast<double>(fabsf(__fsub_rn(x, 0.5f)));
                ++gt;
            }
        }
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, o);
            gt += __shfl_xor_sync(0xffffffffu, gt, o);
        }
        if (lane == 0) {
            s_red[warp] = sum;
            s_gt[warp] = gt;
        }
        __syncthreads();
        if (tid == 0) {
            double total = 0.0;
            unsigned int g = 0;
            for (int w = 0; w < 32; ++w) {
                total += s_red[w];
                g += s_gt[w];
            }
            const uint32_t tb = thr ^ ((thr >> 31) ? 0x80000000u : 0xFFFFFFFFu);   // inverse of f32_key
            const float tval = __uint_as_float(tb);
            total += static_cast<double>(k - static_cast<int>(g)) * static_cast<double>(fabsf(__fsub_rn(tval, 0.5f)));
            scores[n * 4 + 3] = total / k * 2.0;
        }
    } else if (tid == 0) {
        scores[n * 4 + 3] = nan("");   // L == 0: np.mean of an empty array
    }

    // ---- quadrangle regularity (float32 arithmetic, process_pipeline.py:430-456)
    if (tid == 32) {
        scores[n * 4 + 1] = nan("");
        double reg = 0.0;
        if (quad != nullptr && (found == nullptr || found[n])) {
            float qx[4], qy[4], side[4], ang[4];
            for (int i = 0; i < 4; ++i) {
                qx[i] = quad[(n * 4 + i) * 2];
                qy[i] = quad[(n * 4 + i) * 2 + 1];
            }
            for (int i = 0; i < 4; ++i) {
                const int j = (i + 1) & 3;
                const float dx = __fsub_rn(qx[i], qx[j]), dy = __fsub_rn(qy[i], qy[j]);
                side[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
            }
            for (int i = 0; i < 4; ++i) {
                const int a = (i + 3) & 3, b = (i + 1) & 3;
                const float v1x = __fsub_rn(qx[a], qx[i]), v1y = __fsub_rn(qy[a], qy[i]);
                const float v2x = __fsub_rn(qx[b], qx[i]), v2y = __fsub_rn(qy[b], qy[i]);
                const float dot = __fadd_rn(__fmul_rn(v1x, v2x), __fmul_rn(v1y, v2y));
                const float n1 = __fsqrt_rn(__fadd_rn(__fmul_rn(v1x, v1x), __fmul_rn(v1y, v1y)));
                const float n2 = __fsqrt_rn(__fadd_rn(__fmul_rn(v2x, v2x), __fmul_rn(v2y, v2y)));
                const float nn = __fmul_rn(n1, n2);
                ang[i] = nn > 0.f ? acosf(__fdiv_rn(dot, nn)) : 0.f;
            }
            // np.std / np.mean over a list of float32 scalars: float32 array, float32 result
            float ms = 0.f, ma = 0.f;
            for (int i = 0; i < 4; ++i) {
                ms += side[i];
                ma += ang[i];
            }
            ms *= 0.25f;
            ma *= 0.25f;
            float vs = 0.f, va = 0.f;
            for (int i = 0; i < 4; ++i) {
                vs += (side[i] - ms) * (side[i] - ms);
                va += (ang[i] - ma) * (ang[i] - ma);
            }
            const float sd_s = __fsqrt_rn(vs * 0.25f), sd_a = __fsqrt_rn(va * 0.25f);
            // numpy keeps float32 throughout (python scalars are weak under NEP 50); only the final float() widens
            const float side_var = ms > 0.f ? __fdiv_rn(sd_s, ms) : 1.0f;
            const float angle_var = __fdiv_rn(sd_a, 1.5707963267948966f);
            reg = static_cast<double>(__fsub_rn(1.0f, __fadd_rn(__fmul_rn(side_var, 0.5f), __fmul_rn(angle_var, 0.5f))));
        }
        scores[n * 4 + 0] = reg;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// n4, mask_completeness (process_pipeline.py:380-414) for 256x256 arrays: (#pixels > 0.5) / (#pixels of the filled largest
// external contour).  What cv2.findContours(RETR_EXTERNAL) + max(key=contourArea) + drawContours(thickness=-1) amount to
// (checked against live cv2 by the test suite): among the 8-connected components that touch the frame-connected background
// take the one whose outer border polygon has the largest area (Green's formula over the border-following chain; among
// equal areas the one found LAST in raster order, because cv2 lists external contours in reverse discovery order and
// max() keeps the first); its filled drawing is the component plus everything it encloses.
// One block of 256 threads per board, thread <-> image row; all pixel sets are bit planes of 258-bit rows (a one-pixel
// zero frame) in shared memory, flood fills run as row-parallel sweeps: vertical seeds from the rows above / below, then
// the whole horizontal run through the carry chain of an addition.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int MW = 9;                    // 32-bit words per plane row: bits 0..257, pixel x at bit x + 1
constexpr int MROWS = 258;

__device__ __forceinline__ void row_load(const uint32_t* pl, int y, uint32_t (&r)[MW]) {
#pragma unroll
    for (int i = 0; i < MW; ++i) r[i] = pl[y * MW + i];
}

// every bit of mask-run that contains a seed bit: smear up through the carry chain, then the same on the reversed row
__device__ __forceinline__ void run_fill(const uint32_t (&m)[MW], uint32_t (&s)[MW]) {
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < MW; ++i) {   // upwards: (m + s) ^ m covers seed .. top of its run (+ the bit above, masked off)
        const uint64_t sum = static_cast<uint64_t>(m[i]) + s[i] + carry;
        carry = static_cast<uint32_t>(sum >> 32);
        s[i] |= (static_cast<uint32_t>(sum) ^ m[i]) & m[i];
    }
    carry = 0;
#pragma unroll
    for (int i = MW - 1; i >= 0; --i) {   // downwards: bit-reversed words, most significant word first
        const uint32_t mr = __brev(m[i]), sr = __brev(s[i]);
        const uint64_t sum = static_cast<uint64_t>(mr) + sr + carry;
        carry = static_cast<uint32_t>(sum >> 32);
        s[i] |= __brev((static_cast<uint32_t>(sum) ^ mr) & mr);
    }
}

// Flood `cur` (seeded by the caller) through `mask` (INVERT: through its complement, i.e. the background, whose frame
// cells the caller has pre-seeded); CONN8: diagonal neighbours connect (foreground), else 4-connectivity (background).
// Thread <-> row y in 1..256; the frame rows 0 and 257 are only read.
template <bool CONN8, bool INVERT>
__device__ void flood(const uint32_t* mask, uint32_t* cur, int y) {
    for (;;) {
        uint32_t m[MW], c[MW], s[MW];
        row_load(mask, y, m);
        if (INVERT) {
#pragma unroll
            for (int i = 0; i < MW; ++i) m[i] = ~m[i] & (i == MW - 1 ? 0x3u : 0xFFFFFFFFu);
        }
        row_load(cur, y, c);
#pragma unroll
        for (int i = 0; i < MW; ++i) s[i] = c[i];
#pragma unroll
        for (int dy = -1; dy <= 1; dy += 2) {
            uint32_t n[MW];
            row_load(cur, y + dy, n);
#pragma unroll
            for (int i = 0; i < MW; ++i) {
                uint32_t v = n[i];
                if (CONN8) v |= (n[i] << 1) | (i ? n[i - 1] >> 31 : 0u) | (n[i] >> 1) | (i + 1 < MW ? n[i + 1] << 31 : 0u);
                s[i] |= v & m[i];
            }
        }
        run_fill(m, s);
        bool changed = false;
#pragma unroll
        for (int i = 0; i < MW; ++i) changed |= s[i] != c[i];
        __syncthreads();   // everybody has read its neighbours' rows
        if (changed) {
#pragma unroll
            for (int i = 0; i < MW; ++i) cur[y * MW + i] = s[i];
        }
        if (!__syncthreads_or(changed)) return;
    }
}

// all frame cells (row 0, row 257, bit 0 and bit 257 of every row) set, everything else clear
__device__ void seed_frame(uint32_t* pl, int tid) {
    for (int i = tid; i < MROWS * MW; i += 256) {
        const int r = i / MW, w = i - r * MW;
        const uint32_t full = w == MW - 1 ? 0x3u : 0xFFFFFFFFu;
        pl[i] = (r == 0 || r == MROWS - 1) ? full : (w == 0 ? 1u : (w == MW - 1 ? 2u : 0u));
    }
}

__device__ __forceinline__ int pl_get(const uint32_t* pl, int x, int y) { return (pl[y * MW + (x >> 5)] >> (x & 31)) & 1; }   // framed coordinates

// Outer border following from the raster-first pixel (x0, y0) of a component (framed coordinates), exactly
// the way cv2.findContours follows it (Suzuki-Abe with OpenCV's termination rule, cf. trace_border in geometry.cu); returns |sum of cross products| = 2 * contourArea.
__device__ long long border_area2(const uint32_t* F, int x0, int y0) {
    constexpr int DX[8] = {1, 1, 0, -1, -1, -1, 0, 1}, DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    int s = 4;
    const int s_end = 4;
    do {
        s = (s - 1) & 7;
    } while (!pl_get(F, x0 + DX[s], y0 + DY[s]) && s != s_end);
    if (s == s_end) return 0;   // isolated pixel
    const int x1 = x0 + DX[s], y1 = y0 + DY[s];
    int x3 = x0, y3 = y0;
    long long acc = 0;
    for (;;) {
        int x4, y4;
        for (;;) {
            ++s;
            x4 = x3 + DX[s & 7];
            y4 = y3 + DY[s & 7];
            if (pl_get(F, x4, y4)) break;
        }
        s &= 7;
        acc += static_cast<long long>(x3) * y4 - static_cast<long long>(y3) * x4;   // edge (x3,y3) -> (x4,y4)
        if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
        x3 = x4;
        y3 = y4;
        s = (s + 4) & 7;
    }
    return acc < 0 ? -acc : acc;
}

__global__ void __launch_bounds__(256) k_mask_completeness(const float* __restrict__ vals, double* __restrict__ scores) {
    // F foreground, W foreground not yet assigned to a component, O background connected to the frame,
    // C current component, B best component so far
    __shared__ uint32_t F[MROWS * MW], W[MROWS * MW], O[MROWS * MW], C[MROWS * MW], B[MROWS * MW];
    __shared__ int s_first, s_ext, s_new, s_orig[8], s_out[8];
    __shared__ long long s_best;
    const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, y = tid + 1;
    const float* v = vals + static_cast<size_t>(n) * 65536;
    for (int i = tid; i < MROWS * MW; i += 256) F[i] = B[i] = C[i] = 0u;
    __syncthreads();
    // pixel (x, r) > 0.5 -> bit x of word x/32 (unframed, staged in C); warp w covers rows 32w .. 32w+31
    for (int r = warp * 32; r < warp * 32 + 32; ++r)
        for (int k = 0; k < 8; ++k) {
            const unsigned bits = __ballot_sync(0xffffffffu, __ldg(v + r * 256 + k * 32 + lane) > 0.5f);
            if (lane == 0) C[(r + 1) * MW + k] = bits;
        }
    __syncthreads();
    {
        uint32_t u[MW];
        row_load(C, y, u);
#pragma unroll
        for (int i = 0; i < MW; ++i) {
            const uint32_t w = (i < 8 ? u[i] << 1 : 0u) | (i ? u[i - 1] >> 31 : 0u);   // shift by the frame column
            F[y * MW + i] = w;
            W[y * MW + i] = w;
        }
        if (tid < MW) W[tid] = W[(MROWS - 1) * MW + tid] = 0u;
    }
    seed_frame(O, tid);
    if (tid == 0) s_best = -1;
    __syncthreads();
    flood<false, true>(F, O, y);
    for (;;) {
        if (tid == 0) s_first = 0x7fffffff;
        __syncthreads();
        {   // raster-first foreground pixel not yet assigned to a component
            int firstx = -1;
#pragma unroll
            for (int i = MW - 1; i >= 0; --i) {
                const uint32_t w = W[y * MW + i];
                if (w) firstx = i * 32 + __ffs(w) - 1;
            }
            if (firstx >= 0) atomicMin(&s_first, y * 512 + firstx);
        }
        __syncthreads();
        const int first = s_first;
        if (first == 0x7fffffff) break;
        const int fx = first & 511, fy = first >> 9;
        for (int i = tid; i < MROWS * MW; i += 256) C[i] = 0u;
        __syncthreads();
        if (tid == 0) {
            C[fy * MW + (fx >> 5)] = 1u << (fx & 31);
            s_ext = 0;
            s_new = 0;
        }
        __syncthreads();
        flood<true, false>(F, C, y);   // C = the 8-connected component of that pixel
        {   // external: some pixel of C is 4-adjacent to the frame-connected background
            uint32_t c[MW], up[MW], dn[MW], o[MW];
            row_load(C, y, c);
            row_load(O, y - 1, up);
            row_load(O, y + 1, dn);
            row_load(O, y, o);
            bool ext = false;
#pragma unroll
            for (int i = 0; i < MW; ++i) {
                const uint32_t side = (o[i] << 1) | (i ? o[i - 1] >> 31 : 0u) | (o[i] >> 1) | (i + 1 < MW ? o[i + 1] << 31 : 0u);
                ext |= (c[i] & (up[i] | dn[i] | side)) != 0u;
                W[y * MW + i] &= ~c[i];
            }
            if (ext) s_ext = 1;
        }
        __syncthreads();
        if (s_ext && tid == 0) {
            const long long a2 = border_area2(F, fx, fy);
            if (a2 >= s_best) {   // ties: the later discovery wins (cv2 lists external contours in reverse discovery order, max() keeps the first)
                s_best = a2;
                s_new = 1;
            }
        }
        __syncthreads();
        if (s_new)
            for (int i = tid; i < MROWS * MW; i += 256) B[i] = C[i];
        __syncthreads();
    }
    // filled drawing of the chosen contour = everything NOT connected to the frame when only that component blocks
    const bool any = s_best >= 0;
    seed_frame(O, tid);
    __syncthreads();
    flood<false, true>(B, O, y);
    int orig = 0, outside = 0;
#pragma unroll
    for (int i = 0; i < MW; ++i) {
        orig += __popc(F[y * MW + i]);
        uint32_t o = O[y * MW + i];
        if (i == 0) o &= ~1u;            // frame column 0
        if (i == MW - 1) o &= 1u;        // bit 0 of the last word is pixel 255, bit 1 the frame column 257
        outside += __popc(o);
    }
    for (int off = 16; off > 0; off >>= 1) {
        orig += __shfl_xor_sync(0xffffffffu, orig, off);
        outside += __shfl_xor_sync(0xffffffffu, outside, off);
    }
    if (lane == 0) {
        s_orig[warp] = orig;
        s_out[warp] = outside;
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0, out = 0;
        for (int i = 0; i < 8; ++i) {
            o += s_orig[i];
            out += s_out[i];
        }
        const int filled = 65536 - out;
        scores[n * 4 + 1] = (!any || filled == 0) ? 0.0 : static_cast<double>(o) / static_cast<double>(filled);
    }
}

int set_dev(cvb_ctx* ctx) {
    CVB_ON_DEVICE(ctx);
    return 0;
}

}  // namespace
}  // namespace cvb

extern "C" {

int cvb_eval_metrics(cvb_ctx* ctx, const float* probs, const uint8_t* labels, const uint8_t* labels_valid, const uint8_t* true_labels,
                     int N, int flip, int k, int32_t* topk_hits, int32_t* correct, void* stream) {
    if (!ctx || !probs || !true_labels || !topk_hits || N < 0 || k < 1 || k > 13) return fail(ctx, -1, "cvb_eval_metrics: bad argument");
    if (cvb::set_dev(ctx)) return -2;
    if (N == 0) return 0;
    cvb::k_eval_metrics<<<N, 64, 0, static_cast<cudaStream_t>(stream)>>>(probs, labels, labels_valid, true_labels, flip, k, topk_hits, correct);
    CK(cudaGetLastError());
    ctx->launches++;
    return 0;
}

int cvb_quality_scores(cvb_ctx* ctx, const float* values, const float* quad, const uint8_t* found, int N, int L, double* scores,
                       void* stream) {
    if (!ctx || !values || !scores || N < 0 || L < 1) return fail(ctx, -1, "cvb_quality_scores: bad argument");
    if (cvb::set_dev(ctx)) return -2;
    if (N == 0) return 0;
    cvb::k_quality<<<N, 1024, 0, static_cast<cudaStream_t>(stream)>>>(values, quad, found, L, scores);
    CK(cudaGetLastError());
    ctx->launches++;
    if (L == 65536) {   // mask_completeness is defined on the 2-D array: 256 x 256, the shape the reference passes
        cvb::k_mask_completeness<<<N, 256, 0, static_cast<cudaStream_t>(stream)>>>(values, scores);
        CK(cudaGetLastError());
        ctx->launches++;
    }
    return 0;
}

}  // extern "C"
