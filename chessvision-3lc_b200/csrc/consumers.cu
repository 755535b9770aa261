// Per-board consumers of the image->FEN outputs, computed where the outputs already are.
//
// Replaces (reference, numpy / python-chess on the host, one board at a time):
//   n1  compute_model_topk_accuracy, compute_position_accuracy     scripts/eval/evaluate.py:37-52,109-140
//   n4  probability_distribution, probability_confidence,           scripts/process_new_raw/process_pipeline.py:357-378,
//       quadrangle_regularity                                        416-467
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
// scores f64 [N,4] = {regularity, NaN (mask_completeness: not computed here), distribution, confidence}.
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
    const int k = static_cast<int>(L * 0.25);
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
        scores[n * 4 + 3] = nan("");   // np.mean of an empty slice
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

int set_dev(cvb_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
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
    return 0;
}

}  // extern "C"
