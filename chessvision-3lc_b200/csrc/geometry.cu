// Integer / geometric half of the image->FEN path on the device (compiled with -fmad=false: every float64 result
// below must equal what OpenCV computes on the CPU, so no contraction is allowed; explicit fma() calls appear only
// where accuracy rather than OpenCV's order of roundings matters -- the float64 segment anchors of k_warp_board).
//
//   k_mask_to_quad   ChessVision._find_quadrangle (core.py:358-379): cv2.findContours(RETR_CCOMP, TC89_KCOS) +
//                    _filter_contours (core.py:382-404) + arcLength/approxPolyDP (core.py:372-377) +
//                    _rotate_quadrangle (core.py:407-411).  One warp per board, label image + contour in shared memory.
//   k_homography     _scale_quadrangle (core.py:414-417) + cv2.getPerspectiveTransform (utils.py:127-131) + the
//                    3x3 inversion cv2.warpPerspective performs.
//   k_warp_board     cv2.warpPerspective (utils.py:132) + cvtColor BGR2GRAY (core.py:299) + flip (core.py:300),
//                    written directly in the layout extract_squares (core.py:420-439) views as 64 squares.
#include "kernels.h"

#include <math.h>

#include <type_traits>

namespace cvb {

// =====================================================================================================================
// mask -> quad
// =====================================================================================================================
namespace {

constexpr int LW = 258;                 // padded label row
constexpr int kLabBytes = ((LW * LW * 2 + 15) / 16) * 16;
constexpr int kQuadSmem = kLabBytes + kQuadMaxPoints * (4 + 4 + 2 + 1) + 64;
constexpr unsigned FULL = 0xffffffffu;

// large-capacity fallback (k_mask_to_quad<true>): per scratch slot int32 labels + P, S, KS, dst (kBigPoints ints each) + the
// approxPolyDP stack (2 * kBigPoints) + the border-owner table + the chain codes
constexpr int kBigPoints = 4 * 65536;
constexpr int kBigBorders = 70000;
constexpr size_t kBigLabBytes = ((static_cast<size_t>(LW) * LW * 4 + 15) / 16) * 16;
constexpr size_t kBigSlotBytes = ((kBigLabBytes + static_cast<size_t>(kBigPoints) * 4 * 6 + (kBigBorders + 8) * 4 + kBigPoints + 255) / 256) * 256;

__constant__ int c_off16[16] = {1, -LW + 1, -LW, -LW - 1, -1, LW - 1, LW, LW + 1, 1, -LW + 1, -LW, -LW - 1, -1, LW - 1, LW, LW + 1};
__constant__ int c_kcos_t[15] = {1, 2, 3, 4, 3, 2, 1, 0, 1, 2, 3, 4, 3, 2, 1};

struct Contour {
    int n;          // border points (may exceed kQuadMaxPoints -> overflow)
    int minx, maxx, miny, maxy;
};

__device__ __forceinline__ int pack_pt(int idx) {
    const int y = idx / LW, x = idx - y * LW;
    return (x - 1) | ((y - 1) << 16);
}
__device__ __forceinline__ int px(int p) { return p & 0xffff; }
__device__ __forceinline__ int py(int p) { return p >> 16; }

// Suzuki-Abe border following from `start` (lane 0 only).  LabT: int16 labels in shared memory, int32 in the large-capacity
// kernel; `cap` = border points the P / CODE arrays hold (points beyond it are counted, not stored).
template <typename LabT>
__device__ void trace_border(LabT* lab, int start, int nbd, bool hole, int* P, uint8_t* CODE, Contour& c, int cap) {
    int s_end = hole ? 0 : 4;
    int s = s_end;
    int i1;
    do {
        s = (s - 1) & 7;
        i1 = start + c_off16[s];
    } while (lab[i1] == 0 && s != s_end);
    const int p0 = pack_pt(start);
    c.minx = c.maxx = px(p0);
    c.miny = c.maxy = py(p0);
    if (s == s_end) {  // isolated pixel
        lab[start] = static_cast<LabT>(-nbd);
        P[0] = p0;
        c.n = 1;
        return;
    }
    int i3 = start, n = 0;
    for (;;) {
        s_end = s;
        int i4;
        do {
            ++s;
            i4 = i3 + c_off16[s & 15];
        } while (lab[i4] == 0);
        s &= 7;
        if (static_cast<unsigned>(s - 1) < static_cast<unsigned>(s_end)) {
            lab[i3] = static_cast<LabT>(-nbd);
        } else if (lab[i3] == 1) {
            lab[i3] = static_cast<LabT>(nbd);
        }
        const int p = pack_pt(i3);
        if (n < cap) {
            P[n] = p;
            CODE[n] = static_cast<uint8_t>(s);
        }
        ++n;
        c.minx = min(c.minx, px(p));
        c.maxx = max(c.maxx, px(p));
        c.miny = min(c.miny, py(p));
        c.maxy = max(c.maxy, py(p));
        if (i4 == start && i3 == i1) break;
        i3 = i4;
        s = (s + 4) & 7;
    }
    c.n = n;
}

__device__ __forceinline__ int wrap(int i, int n) {
    if (i < 0) {
        i += n;
        if (i < 0) { i %= n; if (i < 0) i += n; }
    } else if (i >= n) {
        i -= n;
        if (i >= n) i %= n;
    }
    return i;
}

// TC89_KCOS pass 1 for point i (region of support + k-cosine); returns s, writes k.
__device__ int kcos_point(const int* P, int n, int i, int& k_out) {
    const int xi = px(P[i]), yi = py(P[i]);
    int d_num = 0, l = 0, k = 1;
    for (;; ++k) {
        const int p1 = P[wrap(i - k, n)], p2 = P[wrap(i + k, n)];
        const int dx = px(p2) - px(p1), dy = py(p2) - py(p1);
        const int lk = dx * dx + dy * dy;
        const int dk = (xi - px(p1)) * dy - (yi - py(p1)) * dx;
        const float t = static_cast<float>(static_cast<double>(d_num) * static_cast<double>(lk) -
                                           static_cast<double>(dk) * static_cast<double>(l));
        if (k > 1 && (l >= lk || (d_num > 0 && t <= 0.f) || (d_num < 0 && t >= 0.f))) break;
        d_num = dk;
        l = lk;
    }
    --k;
    k_out = k;
    int sv = 0;
    for (int j = k; j > 0; --j) {
        const int pa = P[wrap(i - j, n)], pb = P[wrap(i + j, n)];
        const int ax = px(pa) - xi, ay = py(pa) - yi, bx = px(pb) - xi, by = py(pb) - yi;
        if ((ax | ay) == 0 || (bx | by) == 0) break;
        const double num = static_cast<double>(ax * bx + ay * by);
        const double den = sqrt(static_cast<double>(ax * ax + ay * ay) * static_cast<double>(bx * bx + by * by));
        const float cs = static_cast<float>(num / den);
        const int sk = __float_as_int(static_cast<float>(static_cast<double>(cs) + 1.1));
        if (j < k && sk <= sv) break;
        sv = sk;
    }
    return sv;
}

struct PolyStats {
    double area;
    int bw, bh;
    double arclen;
};

__device__ void poly_stats(const int* R, int m, PolyStats& st) {
    long long a2 = 0;
    int minx = 1 << 20, maxx = -1, miny = 1 << 20, maxy = -1;
    double total = 0.0;
    float buf[16];
    int nb = 0;
    int prev = R[m - 1];
    for (int i = 0; i < m; ++i) {
        const int p = R[i];
        a2 += static_cast<long long>(px(prev)) * py(p) - static_cast<long long>(py(prev)) * px(p);
        minx = min(minx, px(p));
        maxx = max(maxx, px(p));
        miny = min(miny, py(p));
        maxy = max(maxy, py(p));
        const float dx = static_cast<float>(px(p)) - static_cast<float>(px(prev));
        const float dy = static_cast<float>(py(p)) - static_cast<float>(py(prev));
        buf[nb++] = __fsqrt_rn(dx * dx + dy * dy);
        if (nb == 16 || i == m - 1) {
            for (; nb > 0; --nb) total += static_cast<double>(buf[nb - 1]);
        }
        prev = p;
    }
    st.area = fabs(static_cast<double>(a2) * 0.5);
    st.bw = maxx - minx + 1;
    st.bh = maxy - miny + 1;
    st.arclen = m > 1 ? total : 0.0;
}

// cv::approxPolyDP(closed) on R[0..m) -> number of output vertices; the vertices end up in dst[0..count).
__device__ int approx_poly(const int* R, int m, double epsilon, int* dst, int* stack) {
    const double eps = epsilon * epsilon;
    int top = 0, count = 0;
    int pos = 0, rstart = 0;
    bool le_eps = false;
    int sp = 0;
    for (int it = 0; it < 3; ++it) {
        double max_dist = 0.0;
        pos = (pos + rstart) % m;
        sp = R[pos];
        pos = pos + 1 >= m ? 0 : pos + 1;
        for (int j = 1; j < m; ++j) {
            const int p = R[pos];
            pos = pos + 1 >= m ? 0 : pos + 1;
            const int dx = px(p) - px(sp), dy = py(p) - py(sp);
            const double dist = static_cast<double>(dx * dx + dy * dy);
            if (dist > max_dist) {
                max_dist = dist;
                rstart = j;
            }
        }
        le_eps = max_dist <= eps;
    }
    if (!le_eps) {
        const int s_start = pos % m;
        const int s_end = (rstart + s_start) % m;
        stack[2 * top] = s_end;  // right slice
        stack[2 * top + 1] = s_start;
        ++top;
        stack[2 * top] = s_start;  // slice, popped first
        stack[2 * top + 1] = s_end;
        ++top;
    } else {
        dst[count++] = sp;
    }
    while (top > 0) {
        --top;
        const int a = stack[2 * top], b = stack[2 * top + 1];
        const int ep = R[b];
        pos = a;
        sp = R[pos];
        pos = pos + 1 >= m ? 0 : pos + 1;
        bool le;
        int split = 0;
        if (pos != b) {
            const double dx = static_cast<double>(px(ep) - px(sp)), dy = static_cast<double>(py(ep) - py(sp));
            double max_dist = 0.0;
            while (pos != b) {
                const int p = R[pos];
                pos = pos + 1 >= m ? 0 : pos + 1;
                const double dist = fabs(static_cast<double>(py(p) - py(sp)) * dx - static_cast<double>(px(p) - px(sp)) * dy);
                if (dist > max_dist) {
                    max_dist = dist;
                    split = (pos + m - 1) % m;
                }
            }
            le = max_dist * max_dist <= eps * (dx * dx + dy * dy);
        } else {
            le = true;
        }
        if (le) {
            dst[count++] = sp;
        } else {
            stack[2 * top] = split;
            stack[2 * top + 1] = b;
            ++top;
            stack[2 * top] = a;
            stack[2 * top + 1] = split;
            ++top;
        }
    }
    // clean-up pass, in place
    int new_count = count;
    pos = count - 1;
    int start = dst[pos];
    pos = pos + 1 >= count ? 0 : pos + 1;
    int wpos = pos;
    int pt = dst[pos];
    pos = pos + 1 >= count ? 0 : pos + 1;
    for (int i = 0; i < count && new_count > 2; ++i) {
        const int end = dst[pos];
        pos = pos + 1 >= count ? 0 : pos + 1;
        const double dx = static_cast<double>(px(end) - px(start)), dy = static_cast<double>(py(end) - py(start));
        const double dist = fabs(static_cast<double>(px(pt) - px(start)) * dy - static_cast<double>(py(pt) - py(start)) * dx);
        const double sip = static_cast<double>((px(pt) - px(start)) * (px(end) - px(pt)) + (py(pt) - py(start)) * (py(end) - py(pt)));
        if (dist * dist <= 0.5 * eps * (dx * dx + dy * dy) && dx != 0.0 && dy != 0.0 && sip >= 0.0) {
            --new_count;
            dst[wpos] = start = end;
            wpos = wpos + 1 >= count ? 0 : wpos + 1;
            pt = dst[pos];
            pos = pos + 1 >= count ? 0 : pos + 1;
            ++i;
            continue;
        }
        dst[wpos] = start = pt;
        wpos = wpos + 1 >= count ? 0 : wpos + 1;
        pt = end;
    }
    return new_count;
}

// BIG = false: labels (int16), border points and the TC89_KCOS work arrays in shared memory, capacities kQuadMaxPoints /
// kQuadMaxBorders / kQuadMaxVertices; a mask that exceeds one of them gets status QUAD_OVERFLOW.
// BIG = true:  the same algorithm for exactly those masks with every array in global scratch (int32 labels, kBigPoints
// border points: four visits of every pixel, which border following cannot exceed), so that no mask a 256x256 image can
// hold is ever reported as "capacity exceeded".  A handful of scratch slots are shared by the (rare) boards that need them.
template <bool BIG>
__global__ void __launch_bounds__(32, 1) k_mask_to_quad(const uint8_t* __restrict__ mask, int32_t* __restrict__ quad,
                                                        uint8_t* __restrict__ found, int32_t* __restrict__ status,
                                                        int32_t* __restrict__ n_contours, int32_t* __restrict__ owner_scratch,
                                                        int only_flagged, uint8_t* __restrict__ big_scratch, int* __restrict__ big_locks) {
    using LabT = typename std::conditional<BIG, int32_t, int16_t>::type;
    using KsT = typename std::conditional<BIG, int32_t, uint16_t>::type;
    constexpr int kCapPoints = BIG ? kBigPoints : kQuadMaxPoints;
    constexpr int kCapBorders = BIG ? kBigBorders : kQuadMaxBorders;
    constexpr int kCapVertices = BIG ? kBigPoints : kQuadMaxVertices;
    extern __shared__ __align__(16) uint8_t smem[];
    const int b = blockIdx.x, lane = threadIdx.x;
    if (BIG ? status[b] != QUAD_OVERFLOW : (only_flagged && status[b] != QUAD_NEED_FULL)) return;   // resolved by an earlier kernel
    LabT* lab;
    int *P, *S, *owner, *poly_dst, *poly_stack;
    KsT* KS;
    uint8_t* CODE;
    Contour* cshare;
    int slot = -1;
    if constexpr (BIG) {
        // claim one of the scratch slots (blocks that hold one run to completion on their own, so spinning cannot deadlock)
        if (lane == 0) {
            while (slot < 0) {
                for (int k = 0; k < kBigSlots && slot < 0; ++k)
                    if (atomicCAS(big_locks + k, 0, 1) == 0) slot = k;
                if (slot < 0) __nanosleep(2000);
            }
            __threadfence();
        }
        slot = __shfl_sync(FULL, slot, 0);
        uint8_t* base = big_scratch + static_cast<size_t>(slot) * kBigSlotBytes;
        lab = reinterpret_cast<LabT*>(base);
        P = reinterpret_cast<int*>(base + kBigLabBytes);
        S = P + kBigPoints;
        KS = S + kBigPoints;
        poly_dst = KS + kBigPoints;
        poly_stack = poly_dst + kBigPoints;
        owner = poly_stack + 2 * kBigPoints;
        CODE = reinterpret_cast<uint8_t*>(owner + kBigBorders + 8);
        cshare = reinterpret_cast<Contour*>(smem);
    } else {
        lab = reinterpret_cast<LabT*>(smem);
        P = reinterpret_cast<int*>(smem + kLabBytes);
        S = P + kQuadMaxPoints;
        KS = reinterpret_cast<KsT*>(S + kQuadMaxPoints);
        CODE = reinterpret_cast<uint8_t*>(KS + kQuadMaxPoints);
        cshare = reinterpret_cast<Contour*>(CODE + kQuadMaxPoints);
        owner = owner_scratch + static_cast<size_t>(b) * (kQuadMaxBorders + 8);
        poly_dst = S;
        poly_stack = S + kQuadMaxVertices;
    }
    const uint8_t* m = mask + static_cast<size_t>(b) * 65536;

    // label image: 0 / 1 with a one-pixel zero frame
    for (int i = lane; i < LW; i += 32) {
        lab[i] = 0;
        lab[257 * LW + i] = 0;
        lab[i * LW] = 0;
        lab[i * LW + 257] = 0;
    }
    for (int i = lane; i < 65536 / 4; i += 32) {
        const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(m) + i);
        const int y = i >> 6, x = (i & 63) * 4;
        LabT* d = lab + (y + 1) * LW + x + 1;
        d[0] = (v & 0xffu) ? 1 : 0;
        d[1] = (v & 0xff00u) ? 1 : 0;
        d[2] = (v & 0xff0000u) ? 1 : 0;
        d[3] = (v & 0xff000000u) ? 1 : 0;
    }
    __syncwarp();

    int nbd = 1, ncont = 0;
    bool overflow = false;
    // running winners
    bool first_is4 = false;
    int first_q[4] = {0, 0, 0, 0};
    bool best_valid = false;
    int best_o = -1, best_hole = 0, best_d = -1;
    int best_q[4] = {0, 0, 0, 0};

    for (int y = 1; y <= 256; ++y) {
        const LabT* row = lab + y * LW;
        int x = 1;
        int lnbd = 0;
        while (x <= 257) {
            const int xi = x + lane;
            const bool in = xi <= 257;
            const int v = in ? row[xi] : 0;
            const int vp = in ? row[xi - 1] : 0;
            const bool cand = in && ((v != 0) != (vp != 0));
            const bool marked = in && vp != 0 && vp != 1;
            const unsigned cm = __ballot_sync(FULL, cand);
            const unsigned mm = __ballot_sync(FULL, marked);
            if (cm == 0) {
                if (mm) lnbd = abs(__shfl_sync(FULL, vp, 31 - __clz(mm)));
                x += 32;
                continue;
            }
            const int first = __ffs(cm) - 1;
            const unsigned mm2 = mm & (first == 31 ? FULL : ((2u << first) - 1u));
            if (mm2) lnbd = abs(__shfl_sync(FULL, vp, 31 - __clz(mm2)));
            const int p = __shfl_sync(FULL, v, first);
            const int prev = __shfl_sync(FULL, vp, first);
            const int xc = x + first;
            const bool outer = prev == 0 && p == 1;
            const bool hole = !outer && p == 0 && prev >= 1;
            if (outer || hole) {
                ++nbd;
                const int d = ncont++;
                if (nbd >= kCapBorders) {
                    overflow = true;
                    nbd = kCapBorders - 1;  // keep labels representable; result is flagged invalid anyway
                }
                if (lane == 0) {
                    Contour c;
                    trace_border<LabT>(lab, y * LW + (hole ? xc - 1 : xc), nbd, hole, P, CODE, c, kCapPoints);
                    *cshare = c;
                    const int o = hole ? owner[lnbd] : d;
                    owner[nbd] = o;
                    cshare[1].n = o;
                }
                __syncwarp();
                const Contour c = *cshare;
                const int o = cshare[1].n;
                __syncwarp();
                const bool big = (c.maxx - c.minx + 1) * (c.maxy - c.miny + 1) >= 22937;
                if (d == 0 || big) {
                    if (c.n > kCapPoints) {
                        overflow = true;
                    } else if (c.n > 1) {
                        const int n = c.n;
                        // pass 0: candidate points (direction changes)
                        for (int i = lane; i < n; i += 32) {
                            const int prev_code = CODE[i == 0 ? n - 1 : i - 1];
                            S[i] = c_kcos_t[static_cast<int>(CODE[i]) - prev_code + 7];
                            KS[i] = 0;
                        }
                        __syncwarp();
                        // pass 1: region of support and k-cosine, one candidate per lane
                        for (int i = lane; i < n; i += 32) {
                            if (S[i] != 0) {
                                int k;
                                const int sv = kcos_point(P, n, i, k);
                                S[i] = sv;
                                KS[i] = static_cast<KsT>(k);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) {
                            // pass 2: non-maximum suppression inside half the region of support
                            for (int i = 0; i < n; ++i) {
                                const int k2 = KS[i] >> 1;
                                if (KS[i] == 0) continue;
                                const int si = S[i];
                                bool keep = true;
                                for (int j = 1; j <= k2; ++j) {
                                    if (S[wrap(i - j, n)] > si || S[wrap(i + j, n)] > si) {
                                        keep = false;
                                        break;
                                    }
                                }
                                if (!keep) {
                                    S[i] = 0;
                                    KS[i] = 0;
                                }
                            }
                            // pass 3 + compaction into P[0..mv)
                            int mv = 0;
                            for (int i = 0; i < n; ++i) {
                                if (KS[i] == 0) continue;
                                if (KS[i] == 1 && (S[i] <= S[wrap(i - 1, n)] || S[i] <= S[wrap(i + 1, n)])) {
                                    S[i] = 0;
                                    continue;
                                }
                                P[mv++] = P[i];
                            }
                            int is4 = 0, ov = 0, pass = 0;
                            int q[4] = {0, 0, 0, 0};
                            if (mv > kCapVertices) {
                                ov = 1;
                            } else if (mv > 0) {
                                PolyStats st;
                                poly_stats(P, mv, st);
                                const double area = st.area / 65536.0;
                                const int lo = min(st.bw, st.bh), hi = max(st.bw, st.bh);
                                const double ratio = (lo == 0 || hi == 0) ? -1.0 : static_cast<double>(lo) / static_cast<double>(hi);
                                pass = !(area < 0.35 || area > 1.0) && !(ratio < 0.6);
                                if (d == 0 || pass) {
                                    int* dst = poly_dst;
                                    const int cnt = approx_poly(P, mv, 0.1 * st.arclen, dst, poly_stack);
                                    if (cnt == 4) {
                                        is4 = 1;
                                        for (int j = 0; j < 4; ++j) q[j] = dst[j];
                                    }
                                }
                            }
                            int* r = reinterpret_cast<int*>(cshare);
                            r[0] = is4;
                            r[1] = ov;
                            r[2] = pass;
                            r[3] = q[0];
                            r[4] = q[1];
                            r[5] = q[2];
                            r[6] = q[3];
                        }
                        __syncwarp();
                        const int* r = reinterpret_cast<const int*>(cshare);
                        const int is4 = r[0], ov = r[1], pass = r[2];
                        const int q0 = r[3], q1 = r[4], q2 = r[5], q3 = r[6];
                        __syncwarp();
                        if (ov) overflow = true;
                        if (d == 0 && is4) {
                            first_is4 = true;
                            first_q[0] = q0; first_q[1] = q1; first_q[2] = q2; first_q[3] = q3;
                        }
                        if (pass && is4) {
                            const int hl = hole ? 1 : 0;
                            const bool better = !best_valid || o > best_o ||
                                                (o == best_o && (hl < best_hole || (hl == best_hole && d > best_d)));
                            if (better) {
                                best_valid = true;
                                best_o = o; best_hole = hl; best_d = d;
                                best_q[0] = q0; best_q[1] = q1; best_q[2] = q2; best_q[3] = q3;
                            }
                        }
                    }
                }
            }
            x = xc + 1;
        }
    }

    if (lane == 0) {
        bool ok = false;
        int q[4] = {0, 0, 0, 0};
        if (ncont == 1) {
            ok = first_is4;
            for (int j = 0; j < 4; ++j) q[j] = first_q[j];
        } else if (ncont > 1) {
            ok = best_valid;
            for (int j = 0; j < 4; ++j) q[j] = best_q[j];
        }
        if (overflow) ok = false;
        if (ok && px(q[0]) < px(q[2])) {  // _rotate_quadrangle
            const int t = q[3];
            q[3] = q[2]; q[2] = q[1]; q[1] = q[0]; q[0] = t;
        }
        for (int j = 0; j < 4; ++j) {
            quad[(b * 4 + j) * 2 + 0] = ok ? px(q[j]) : 0;
            quad[(b * 4 + j) * 2 + 1] = ok ? py(q[j]) : 0;
        }
        found[b] = ok ? 1 : 0;
        status[b] = overflow ? QUAD_OVERFLOW : (ok ? QUAD_FOUND : QUAD_NONE);
        n_contours[b] = ncont;
        if constexpr (BIG) {
            __threadfence();
            atomicExch(big_locks + slot, 0);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------------
// Compact variant of the same algorithm: ~46 KB of shared memory per board instead of 223 KB, so four boards are resident
// per SM and the (latency-bound, single-warp) work of different boards overlaps.  State per pixel is three bit planes
// with a one-pixel zero frame (rows of 288 bits): F foreground (never changes: border following only ever relabels
// non-zero pixels), V visited, N "right-bound" (negative label).  Border identities are not kept, which is all that
// ownership of HOLE borders needs; a hole large enough to pass the area filter is therefore not decided here: the board
// is flagged QUAD_NEED_FULL (as is any capacity overflow) and k_mask_to_quad re-runs exactly those boards.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FW = 9;                       // 32-bit words per bit-plane row (258 used bits)
constexpr int FBITS = FW * 32;              // linear bit index = y * 288 + x
constexpr int kPlaneWords = 258 * FW;
constexpr int kFastPoints = 2048;           // border points of one contour
constexpr int kFastVertices = 512;          // vertices after TC89_KCOS
constexpr int kFastSmem = 3 * kPlaneWords * 4 + kFastPoints * (2 + 1 + 4 + 2) + 64;

__constant__ int c_boff16[16] = {1, -FBITS + 1, -FBITS, -FBITS - 1, -1, FBITS - 1, FBITS, FBITS + 1,
                                 1, -FBITS + 1, -FBITS, -FBITS - 1, -1, FBITS - 1, FBITS, FBITS + 1};

__device__ __forceinline__ int bit_get(const uint32_t* pl, int i) { return (pl[i >> 5] >> (i & 31)) & 1; }
__device__ __forceinline__ void bit_set(uint32_t* pl, int i) { pl[i >> 5] |= 1u << (i & 31); }
__device__ __forceinline__ int pack16(int i) {
    const int y = i / FBITS, x = i - y * FBITS;
    return (x - 1) | ((y - 1) << 8);
}
__device__ __forceinline__ int qx(int p) { return p & 255; }
__device__ __forceinline__ int qy(int p) { return p >> 8; }

// The eight neighbours of bit i of plane F as a mask, bit d = F[i + c_boff16[d]] (d: 0 E, 1 NE, 2 N, 3 NW, 4 W, 5 SW, 6 S,
// 7 SE): three funnel-shifted row reads instead of one dependent table lookup + shared-memory read per probed direction.
__device__ __forceinline__ uint32_t nbr_mask(const uint32_t* F, int i) {
    const int j = i - 1;
    const int b0 = j - FBITS, b1 = j, b2 = j + FBITS;
    const uint32_t t = __funnelshift_r(F[b0 >> 5], F[(b0 >> 5) + 1], b0 & 31) & 7u;    // bits: 0 = x-1, 1 = x, 2 = x+1
    const uint32_t m = __funnelshift_r(F[b1 >> 5], F[(b1 >> 5) + 1], b1 & 31) & 7u;
    const uint32_t u = __funnelshift_r(F[b2 >> 5], F[(b2 >> 5) + 1], b2 & 31) & 7u;
    return (m >> 2) | ((t >> 2) << 1) | (((t >> 1) & 1u) << 2) | ((t & 1u) << 3) | ((m & 1u) << 4) | ((u & 1u) << 5) | (((u >> 1) & 1u) << 6) |
           ((u >> 2) << 7);
}

// Suzuki-Abe border following on the bit planes (lane 0 only); same control flow as trace_border.  Per border point the next
// direction comes from one neighbourhood mask (rotate + find-first-set replaces the probing loop), the point's coordinates are
// carried along instead of being divided out of the bit index.
__device__ void trace_border_bits(const uint32_t* F, uint32_t* V, uint32_t* N, int start, bool hole, uint16_t* P, uint8_t* CODE, Contour& c) {
    int s_end = hole ? 0 : 4;
    int s = s_end;
    int i1;
    do {
        s = (s - 1) & 7;
        i1 = start + c_boff16[s];
    } while (!bit_get(F, i1) && s != s_end);
    const int p0 = pack16(start);
    c.minx = c.maxx = qx(p0);
    c.miny = c.maxy = qy(p0);
    if (s == s_end) {  // isolated pixel: labelled -nbd
        bit_set(V, start);
        bit_set(N, start);
        P[0] = static_cast<uint16_t>(p0);
        c.n = 1;
        return;
    }
    int i3 = start, n = 0;
    int x = qx(p0), y = qy(p0);                 // mask-frame coordinates of i3
    int minx = x, maxx = x, miny = y, maxy = y;
    for (;;) {
        s_end = s;
        // first foreground neighbour in directions s+1, s+2, ... (the point we came from is one, so the mask is never empty)
        const uint32_t mask = nbr_mask(F, i3);
        const uint32_t rot = ((mask | (mask << 8)) >> ((s + 1) & 7)) & 0xffu;
        s = (s + __ffs(rot)) & 7;
        const int dx = static_cast<int>((0x21000122u >> (4 * s)) & 0xfu) - 1;    // {1, 1, 0, -1, -1, -1, 0, 1}
        const int dy = static_cast<int>((0x22210001u >> (4 * s)) & 0xfu) - 1;    // {0, -1, -1, -1, 0, 1, 1, 1}
        const int i4 = i3 + dy * FBITS + dx;
        if (static_cast<unsigned>(s - 1) < static_cast<unsigned>(s_end)) {
            bit_set(V, i3);
            bit_set(N, i3);
        } else {
            bit_set(V, i3);   // "if (label == 1) label = nbd": a visited pixel keeps its sign
        }
        if (n < kFastPoints) {
            P[n] = static_cast<uint16_t>(x | (y << 8));
            CODE[n] = static_cast<uint8_t>(s);
        }
        ++n;
        minx = min(minx, x);
        maxx = max(maxx, x);
        miny = min(miny, y);
        maxy = max(maxy, y);
        if (i4 == start && i3 == i1) break;
        i3 = i4;
        x += dx;
        y += dy;
        s = (s + 4) & 7;
    }
    c.minx = minx; c.maxx = maxx; c.miny = miny; c.maxy = maxy;
    c.n = n;
}

// TC89_KCOS pass 1 on 16-bit packed points (same arithmetic as kcos_point).
__device__ int kcos_point16(const uint16_t* P, int n, int i, int& k_out) {
    const int xi = qx(P[i]), yi = qy(P[i]);
    int d_num = 0, l = 0, k = 1;
    for (;; ++k) {
        const int p1 = P[wrap(i - k, n)], p2 = P[wrap(i + k, n)];
        const int dx = qx(p2) - qx(p1), dy = qy(p2) - qy(p1);
        const int lk = dx * dx + dy * dy;
        const int dk = (xi - qx(p1)) * dy - (yi - qy(p1)) * dx;
        const float t = static_cast<float>(static_cast<double>(d_num) * static_cast<double>(lk) -
                                           static_cast<double>(dk) * static_cast<double>(l));
        if (k > 1 && (l >= lk || (d_num > 0 && t <= 0.f) || (d_num < 0 && t >= 0.f))) break;
        d_num = dk;
        l = lk;
    }
    --k;
    k_out = k;
    int sv = 0;
    for (int j = k; j > 0; --j) {
        const int pa = P[wrap(i - j, n)], pb = P[wrap(i + j, n)];
        const int ax = qx(pa) - xi, ay = qy(pa) - yi, bx = qx(pb) - xi, by = qy(pb) - yi;
        if ((ax | ay) == 0 || (bx | by) == 0) break;
        const double num = static_cast<double>(ax * bx + ay * by);
        const double den = sqrt(static_cast<double>(ax * ax + ay * ay) * static_cast<double>(bx * bx + by * by));
        const float cs = static_cast<float>(num / den);
        const int sk = __float_as_int(static_cast<float>(static_cast<double>(cs) + 1.1));
        if (j < k && sk <= sv) break;
        sv = sk;
    }
    return sv;
}

__global__ void __launch_bounds__(32, 4) k_mask_to_quad_fast(const uint8_t* __restrict__ mask, int32_t* __restrict__ quad,
                                                             uint8_t* __restrict__ found, int32_t* __restrict__ status,
                                                             int32_t* __restrict__ n_contours) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t* F = reinterpret_cast<uint32_t*>(smem);
    uint32_t* V = F + kPlaneWords;
    uint32_t* N = V + kPlaneWords;
    int* S = reinterpret_cast<int*>(N + kPlaneWords);
    uint16_t* P = reinterpret_cast<uint16_t*>(S + kFastPoints);
    uint16_t* KS = P + kFastPoints;
    uint8_t* CODE = reinterpret_cast<uint8_t*>(KS + kFastPoints);
    Contour* cshare = reinterpret_cast<Contour*>(CODE + kFastPoints);

    const int b = blockIdx.x, lane = threadIdx.x;
    const uint8_t* m = mask + static_cast<size_t>(b) * 65536;

    for (int i = lane; i < 2 * kPlaneWords; i += 32) V[i] = 0u;   // V and N are contiguous
    for (int i = lane; i < FW; i += 32) {
        F[i] = 0u;
        F[257 * FW + i] = 0u;
    }
    // foreground plane: 16 mask bytes per lane -> 16 bits, two lanes make one word-aligned half... assembled by shuffles
    for (int it0 = 0; it0 < 128; it0 += 8) {   // 512 pixels (two rows) per step; eight independent loads in flight
      uint4 vv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) vv[u] = __ldg(reinterpret_cast<const uint4*>(m) + (it0 + u) * 32 + lane);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int it = it0 + u;
        const uint4 v = vv[u];
        uint32_t bits = 0;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            bits |= (w[k] & 0xffu ? 1u : 0u) << (4 * k);
            bits |= (w[k] & 0xff00u ? 1u : 0u) << (4 * k + 1);
            bits |= (w[k] & 0xff0000u ? 1u : 0u) << (4 * k + 2);
            bits |= (w[k] & 0xff000000u ? 1u : 0u) << (4 * k + 3);
        }
        // lanes 0..15 hold row y = 2*it, lanes 16..31 row y + 1; lane j of a half holds pixels 16j..16j+15
        const uint32_t hi = __shfl_down_sync(FULL, bits, 1);
        const uint32_t word = bits | (hi << 16);                 // valid in even lanes: pixels 32k..32k+31, k = (lane&15)/2
        // shift the row by one bit (frame column 0): plane word k = (pix word k << 1) | (pix word k-1 >> 31)
        const uint32_t prev = __shfl_up_sync(FULL, word, 2);
        if ((lane & 1) == 0) {
            const int k = (lane & 15) >> 1, y = 2 * it + (lane >> 4) + 1;
            F[y * FW + k] = (word << 1) | (k ? prev >> 31 : 0u);
            if (k == 7) F[y * FW + 8] = word >> 31;              // pixel 255 -> bit 256 of the row, bit 257 stays 0
        }
      }
    }
    __syncwarp();

    int ncont = 0;
    bool need_full = false;
    bool first_is4 = false;
    int first_q[4] = {0, 0, 0, 0};
    bool best_valid = false;
    int best_d = -1;
    int best_q[4] = {0, 0, 0, 0};

    // Raster scan for border starting points = 0/1 transitions along a row.  They depend on F alone (border following only
    // changes V and N), so the warp finds them for three rows (27 plane words) at once and then walks the words that have
    // any, in raster order; the visited / right-bound tests still happen at the moment a transition is reached.
    const int r_lane = lane / FW, w_lane = lane - r_lane * FW;
    for (int y0 = 1; y0 <= 256 && !need_full; y0 += 3) {
        const bool act = lane < 3 * FW && y0 + r_lane <= 256;
        const uint32_t Fl = act ? F[(y0 + r_lane) * FW + w_lane] : 0u;
        uint32_t prevw = __shfl_up_sync(FULL, Fl, 1);
        if (w_lane == 0) prevw = 0u;
        const uint32_t Tl = act ? (Fl ^ ((Fl << 1) | (prevw >> 31))) : 0u;
        unsigned pending = __ballot_sync(FULL, Tl != 0u);
        while (pending && !need_full) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint32_t Fw = __shfl_sync(FULL, Fl, src);
            uint32_t T = __shfl_sync(FULL, Tl, src);
            const int y = y0 + src / FW, wi = src % FW;
            while (T && !need_full) {
                const int bpos = __ffs(T) - 1;
                T &= T - 1;
                const int idx = y * FBITS + wi * 32 + bpos;
                const bool fg = (Fw >> bpos) & 1u;
                // outer border: prev == 0 && p == 1 (unvisited).  hole border: p == 0 && prev >= 1 (not a right bound)
                const bool outer = fg && !bit_get(V, idx);
                const bool hole = !fg && !bit_get(N, idx - 1);
                if (!(outer || hole)) continue;
                const int d = ncont++;
                if (lane == 0) {
                    Contour c;
                    trace_border_bits(F, V, N, hole ? idx - 1 : idx, hole, P, CODE, c);
                    *cshare = c;
                }
                __syncwarp();
                const Contour c = *cshare;
                __syncwarp();
                const bool big = (c.maxx - c.minx + 1) * (c.maxy - c.miny + 1) >= 22937;
                if (!(d == 0 || big)) continue;
                if (hole || c.n > kFastPoints) {   // (a first contour is never a hole)
                    need_full = true;
                    break;
                }
                if (c.n <= 1) continue;
                const int n = c.n;
                for (int i = lane; i < n; i += 32) {
                    const int prev_code = CODE[i == 0 ? n - 1 : i - 1];
                    S[i] = c_kcos_t[static_cast<int>(CODE[i]) - prev_code + 7];
                    KS[i] = 0;
                }
                __syncwarp();
                for (int i = lane; i < n; i += 32) {
                    if (S[i] != 0) {
                        int k;
                        const int sv = kcos_point16(P, n, i, k);
                        S[i] = sv;
                        KS[i] = static_cast<uint16_t>(k);
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    for (int i = 0; i < n; ++i) {
                        const int k2 = KS[i] >> 1;
                        if (KS[i] == 0) continue;
                        const int si = S[i];
                        bool keep = true;
                        for (int j = 1; j <= k2; ++j) {
                            if (S[wrap(i - j, n)] > si || S[wrap(i + j, n)] > si) {
                                keep = false;
                                break;
                            }
                        }
                        if (!keep) {
                            S[i] = 0;
                            KS[i] = 0;
                        }
                    }
                    int mv = 0;
                    for (int i = 0; i < n; ++i) {
                        if (KS[i] == 0) continue;
                        if (KS[i] == 1 && (S[i] <= S[wrap(i - 1, n)] || S[i] <= S[wrap(i + 1, n)])) {
                            S[i] = 0;
                            continue;
                        }
                        P[mv++] = P[i];
                    }
                    int is4 = 0, ov = 0, pass = 0;
                    int q[4] = {0, 0, 0, 0};
                    if (mv > kFastVertices) {
                        ov = 1;
                    } else if (mv > 0) {
                        int* R = reinterpret_cast<int*>(KS);      // the k values are dead: reduced vertices as x | y << 16
                        for (int i = 0; i < mv; ++i) R[i] = qx(P[i]) | (qy(P[i]) << 16);
                        PolyStats st;
                        poly_stats(R, mv, st);
                        const double area = st.area / 65536.0;
                        const int lo = min(st.bw, st.bh), hi = max(st.bw, st.bh);
                        const double ratio = (lo == 0 || hi == 0) ? -1.0 : static_cast<double>(lo) / static_cast<double>(hi);
                        pass = !(area < 0.35 || area > 1.0) && !(ratio < 0.6);
                        if (d == 0 || pass) {
                            const int cnt = approx_poly(R, mv, 0.1 * st.arclen, S, S + kFastVertices);
                            if (cnt == 4) {
                                is4 = 1;
                                for (int j = 0; j < 4; ++j) q[j] = S[j];
                            }
                        }
                    }
                    int* r = reinterpret_cast<int*>(cshare);
                    r[0] = is4; r[1] = ov; r[2] = pass; r[3] = q[0]; r[4] = q[1]; r[5] = q[2]; r[6] = q[3];
                }
                __syncwarp();
                const int* r = reinterpret_cast<const int*>(cshare);
                const int is4 = r[0], ov = r[1], pass = r[2];
                const int q0 = r[3], q1 = r[4], q2 = r[5], q3 = r[6];
                __syncwarp();
                if (ov) {
                    need_full = true;
                    break;
                }
                if (d == 0 && is4) {
                    first_is4 = true;
                    first_q[0] = q0; first_q[1] = q1; first_q[2] = q2; first_q[3] = q3;
                }
                // all candidates here are outer borders: cv2 lists outer borders in reverse discovery order
                if (pass && is4 && d > best_d) {
                    best_valid = true;
                    best_d = d;
                    best_q[0] = q0; best_q[1] = q1; best_q[2] = q2; best_q[3] = q3;
                }
            }
        }
    }

    if (lane == 0) {
        if (need_full) {
            status[b] = QUAD_NEED_FULL;
            found[b] = 0;
            for (int j = 0; j < 8; ++j) quad[b * 8 + j] = 0;
            return;
        }
        bool ok = false;
        int q[4] = {0, 0, 0, 0};
        if (ncont == 1) {
            ok = first_is4;
            for (int j = 0; j < 4; ++j) q[j] = first_q[j];
        } else if (ncont > 1) {
            ok = best_valid;
            for (int j = 0; j < 4; ++j) q[j] = best_q[j];
        }
        if (ok && px(q[0]) < px(q[2])) {  // _rotate_quadrangle
            const int t = q[3];
            q[3] = q[2]; q[2] = q[1]; q[1] = q[0]; q[0] = t;
        }
        for (int j = 0; j < 4; ++j) {
            quad[(b * 4 + j) * 2 + 0] = ok ? px(q[j]) : 0;
            quad[(b * 4 + j) * 2 + 1] = ok ? py(q[j]) : 0;
        }
        found[b] = ok ? 1 : 0;
        status[b] = ok ? QUAD_FOUND : QUAD_NONE;
        n_contours[b] = ncont;
    }
}

}  // namespace

cudaError_t configure_quad() {
    cudaError_t e = cudaFuncSetAttribute(k_mask_to_quad_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_mask_to_quad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kQuadSmem);
}

size_t quad_big_scratch_bytes() { return kBigSlotBytes * kBigSlots + 256; }

cudaError_t launch_mask_to_quad(const uint8_t* mask, int32_t* quad, uint8_t* found, int32_t* status, int32_t* n_contours,
                                int32_t* owner_scratch, uint8_t* big_scratch, int N, bool full_only, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    int* locks = reinterpret_cast<int*>(big_scratch + kBigSlotBytes * kBigSlots);   // zero when no kernel is running
    cudaError_t e;
    if (full_only) {
        k_mask_to_quad<false><<<N, 32, kQuadSmem, s>>>(mask, quad, found, status, n_contours, owner_scratch, 0, nullptr, nullptr);
    } else {
        // compact kernel for every board, then the full-state kernel for the boards it flagged (early exit otherwise)
        k_mask_to_quad_fast<<<N, 32, kFastSmem, s>>>(mask, quad, found, status, n_contours);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        k_mask_to_quad<false><<<N, 32, kQuadSmem, s>>>(mask, quad, found, status, n_contours, owner_scratch, 1, nullptr, nullptr);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // ... and the large-capacity kernel for masks that exceed the shared-memory capacities (early exit otherwise)
    k_mask_to_quad<true><<<N, 32, 256, s>>>(mask, quad, found, status, n_contours, owner_scratch, 0, big_scratch, locks);
    return cudaGetLastError();
}

// =====================================================================================================================
// homography
// =====================================================================================================================
namespace {

// cv2.getPerspectiveTransform(corners, ((0,0),(w,0),(w,h),(0,h))) followed by the 3x3 inversion of cv2.warpPerspective:
// corners float32 [4][2] -> out double[9] (all zero when the system is singular).
__device__ void homography_inverse(const float (&cx)[4], const float (&cy)[4], int out_w, int out_h, double* __restrict__ out) {
    double A[8][8], B[8];
    const float dstx[4] = {0.f, static_cast<float>(out_w), static_cast<float>(out_w), 0.f};
    const float dsty[4] = {0.f, 0.f, static_cast<float>(out_h), static_cast<float>(out_h)};
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) A[i][j] = 0.0;
    for (int i = 0; i < 4; ++i) {
        const double sx = static_cast<double>(cx[i]), sy = static_cast<double>(cy[i]);
        A[i][0] = A[i + 4][3] = sx;
        A[i][1] = A[i + 4][4] = sy;
        A[i][2] = A[i + 4][5] = 1.0;
        // Point2f products: OpenCV forms them in float32 (exact for the power-of-two board size, not for a general out_size)
        A[i][6] = static_cast<double>(__fmul_rn(-cx[i], dstx[i]));
        A[i][7] = static_cast<double>(__fmul_rn(-cy[i], dstx[i]));
        A[i + 4][6] = static_cast<double>(__fmul_rn(-cx[i], dsty[i]));
        A[i + 4][7] = static_cast<double>(__fmul_rn(-cy[i], dsty[i]));
        B[i] = static_cast<double>(dstx[i]);
        B[i + 4] = static_cast<double>(dsty[i]);
    }
    bool singular = false;
    for (int i = 0; i < 8; ++i) {
        int k = i;
        for (int j = i + 1; j < 8; ++j)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (fabs(A[k][i]) < 2.220446049250313e-16 * 100) {
            singular = true;
            break;
        }
        if (k != i) {
            for (int j = i; j < 8; ++j) {
                const double t = A[i][j];
                A[i][j] = A[k][j];
                A[k][j] = t;
            }
            const double t = B[i];
            B[i] = B[k];
            B[k] = t;
        }
        const double d = -1.0 / A[i][i];
        for (int j = i + 1; j < 8; ++j) {
            const double alpha = A[j][i] * d;
            for (int kk = i + 1; kk < 8; ++kk) A[j][kk] += alpha * A[i][kk];
            B[j] += alpha * B[i];
        }
    }
    if (singular) {
        for (int i = 0; i < 9; ++i) out[i] = 0.0;
        return;
    }
    for (int i = 7; i >= 0; --i) {
        double s = B[i];
        for (int kk = i + 1; kk < 8; ++kk) s -= A[i][kk] * B[kk];
        B[i] = s / A[i][i];
    }
    const double a00 = B[0], a01 = B[1], a02 = B[2], a10 = B[3], a11 = B[4], a12 = B[5], a20 = B[6], a21 = B[7], a22 = 1.0;
    const double det = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
    const double d = 1.0 / det;
    out[0] = (a11 * a22 - a12 * a21) * d;
    out[1] = (a02 * a21 - a01 * a22) * d;
    out[2] = (a01 * a12 - a02 * a11) * d;
    out[3] = (a12 * a20 - a10 * a22) * d;
    out[4] = (a00 * a22 - a02 * a20) * d;
    out[5] = (a02 * a10 - a00 * a12) * d;
    out[6] = (a10 * a21 - a11 * a20) * d;
    out[7] = (a01 * a20 - a00 * a21) * d;
    out[8] = (a00 * a11 - a01 * a10) * d;
}

__global__ void k_homography(const int32_t* __restrict__ quad, const uint8_t* __restrict__ found, double* __restrict__ minv,
                             int N, float scale, int out_w, int out_h) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= N) return;
    double* out = minv + static_cast<size_t>(b) * 9;
    if (!found[b]) {
        for (int i = 0; i < 9; ++i) out[i] = 0.0;
        return;
    }
    float cx[4], cy[4];
    for (int i = 0; i < 4; ++i) {
        // np.array(approx * sf, dtype=float32): the product is formed in float64, then rounded to float32
        cx[i] = static_cast<float>(static_cast<double>(quad[(b * 4 + i) * 2 + 0]) * static_cast<double>(scale));
        cy[i] = static_cast<float>(static_cast<double>(quad[(b * 4 + i) * 2 + 1]) * static_cast<double>(scale));
    }
    homography_inverse(cx, cy, out_w, out_h, out);
}

// utils.extract_perspective (utils.py:115-132) for ONE image and caller-supplied float32 corners: getPerspectiveTransform +
// warpPerspective(INTER_LINEAR, BORDER_CONSTANT 0) to an arbitrary out_size, C = 1 or 3 channels, the literal arithmetic
// of cv::WarpPerspectiveInvoker per destination pixel (destination blocks bw wide: x1 counts from the block start).
__global__ void k_homography_f32(const float* __restrict__ corners, double* __restrict__ minv, int out_w, int out_h) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float cx[4], cy[4];
    for (int i = 0; i < 4; ++i) {
        cx[i] = corners[2 * i];
        cy[i] = corners[2 * i + 1];
    }
    homography_inverse(cx, cy, out_w, out_h, minv);
}

__global__ void k_warp_generic(const uint8_t* __restrict__ src, int H, int W, int C, const double* __restrict__ m, uint8_t* __restrict__ out,
                               int out_w, int out_h, int bw) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out_w) return;
    const int bx = (x / bw) * bw, x1 = x - bx;
    const double X0 = m[0] * bx + m[1] * y + m[2], Y0 = m[3] * bx + m[4] * y + m[5], W0 = m[6] * bx + m[7] * y + m[8];
    double w = W0 + m[6] * x1;
    w = w != 0.0 ? 32.0 / w : 0.0;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, (X0 + m[0] * x1) * w));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, (Y0 + m[3] * x1) * w));
    const int Xi = __double2int_rn(fX), Yi = __double2int_rn(fY);
    const int sx = max(-32768, min(32767, Xi >> 5)), sy = max(-32768, min(32767, Yi >> 5));
    const int ax = Xi & 31, ay = Yi & 31;
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    const bool y0ok = sy >= 0 && sy < H, y1ok = sy + 1 >= 0 && sy + 1 < H;
    const bool x0ok = sx >= 0 && sx < W, x1ok = sx + 1 >= 0 && sx + 1 < W;
    const uint8_t* p = src + (static_cast<long long>(sy) * W + sx) * C;
    const uint8_t* p2 = p + static_cast<long long>(W) * C;
    uint8_t* o = out + (static_cast<long long>(y) * out_w + x) * C;
    for (int c = 0; c < C; ++c) {
        int acc = 16384;
        if (y0ok && x0ok) acc += w00 * p[c];
        if (y0ok && x1ok) acc += w01 * p[C + c];
        if (y1ok && x0ok) acc += w10 * p2[c];
        if (y1ok && x1ok) acc += w11 * p2[C + c];
        o[c] = static_cast<uint8_t>(acc >> 15);
    }
}

// =====================================================================================================================
// warp + gray + flip.  cv2.warpPerspective evaluates, per destination pixel of a 64-wide block starting at bx,
//     W = W0 + M6*x1;  W = W ? 32/W : 0;  X = rint((X0 + M0*x1)*W);  Y = rint((Y0 + M3*x1)*W)         (float64, no FMA)
// and interpolates with 5-bit fractions, integer weights and (sum + 2^14) >> 15; BGR2GRAY and the mirror follow.
//
// One CTA per 64x64 destination tile (= one chess square of the board image).  The source footprint of the tile is
// staged once through shared memory (coalesced 48-byte groups -> one BGR0 word per pixel, zeros outside the image),
// the gather then runs on shared memory with packed 16x8-bit dot products.
//
// Coordinates.  A thread owns one 16-pixel segment of one tile row.  It anchors U = 32*X/W at the segment centre in
// float64 (once), and per pixel adds the float32 offset
//     U(c + d) - U(c) = d * (32*M0 - U(c)*M6) / W(c + d),        d = -8 .. 7,
// which is short (|offset| < 512 units of 1/32 px, checked per thread), so float32 carries it to 2^-13 of a unit: one
// FFMA for W, one MUFU reciprocal, one FMUL, and one FFMA per axis whose addend is the magic number 1.5 * 2^23 -- the
// integer lands in the mantissa and one IADD puts it on the anchor.  The error of that value is below 2.5 * 2^-13
// (profiles/probes/warp_fast_model.py: 0.5 anchor rounding + 0.5 FMA rounding + 5.3 * 2^-24 relative on the offset; measured
// maximum 2.0 with the reciprocal off by a whole ulp); a pixel whose value lies within 4 * 2^-13 of a rounding boundary
// (0.2 % of them) is marked and recomputed after the segment with the literal OpenCV arithmetic, so the output stays
// bit-identical; so are all pixels of threads whose offsets leave the magic range and of tiles whose footprint does not fit
// the staged patch.  The main loop has no branch.
//
// Footprint.  The offset is monotonic over the segment, so anchor + offsets at d = -8 and d = 7 bound the segment's source
// columns and rows; the CTA's footprint is the min / max of those over its 256 segments (redux.sync per warp, one
// shared-memory exchange).  No thread computes anything the others wait for.
// =====================================================================================================================
constexpr int kWpRows = 88;          // staged patch capacity (source rows)
constexpr int kWpCols16 = 7;         // ... and 16-pixel column groups
constexpr int kWpStride = 116;       // words per staged row (multiple of 4 for 128-bit stores)
constexpr int kWpSmem = kWpRows * kWpStride * 4;
constexpr int kWpFrac = 13;          // bits carried below 1/32 px
constexpr int kWpBand = 8;           // guard band in 2^-13 units (covers an error below 4)
constexpr int kWpMagicBits = 0x4B400000;   // 1.5 * 2^23 as float bits
__constant__ uint32_t kWpInv[8] = {0, 65536, 32768, 21846, 16384, 13108, 10923, 9363};   // ceil(2^16 / n)
constexpr int kWpFar = 1 << 20;      // footprint bound of a segment outside the float32 path: the tile does not fit

// The literal arithmetic for one destination pixel from global memory; returns the gray value.
__device__ __noinline__ uint32_t warp_px_exact(const uint8_t* __restrict__ src, int H, int W, int Xi, int Yi) {
    const int sx = max(-32768, min(32767, Xi >> 5)), sy = max(-32768, min(32767, Yi >> 5));
    const int ax = Xi & 31, ay = Yi & 31;
    const int w00 = (32 - ax) * (32 - ay) * 32, w01 = ax * (32 - ay) * 32, w10 = (32 - ax) * ay * 32, w11 = ax * ay * 32;
    int acc[3] = {16384, 16384, 16384};
    const bool y0ok = sy >= 0 && sy < H, y1ok = sy + 1 >= 0 && sy + 1 < H;
    const bool x0ok = sx >= 0 && sx < W, x1ok = sx + 1 >= 0 && sx + 1 < W;
    const uint8_t* p = src + (static_cast<long long>(sy) * W + sx) * 3;
    if (y0ok && x0ok) { acc[0] += w00 * p[0]; acc[1] += w00 * p[1]; acc[2] += w00 * p[2]; }
    if (y0ok && x1ok) { acc[0] += w01 * p[3]; acc[1] += w01 * p[4]; acc[2] += w01 * p[5]; }
    const uint8_t* p2 = p + static_cast<long long>(W) * 3;
    if (y1ok && x0ok) { acc[0] += w10 * p2[0]; acc[1] += w10 * p2[1]; acc[2] += w10 * p2[2]; }
    if (y1ok && x1ok) { acc[0] += w11 * p2[3]; acc[1] += w11 * p2[4]; acc[2] += w11 * p2[5]; }
    const int bl = acc[0] >> 15, gr = acc[1] >> 15, rd = acc[2] >> 15;
    return static_cast<uint32_t>((3735 * bl + 19235 * gr + 9798 * rd + 16384) >> 15);
}

// Four taps of the staged patch -> gray.  Xl, Yl: patch coordinates in 1/32 px (tap column Xl >> 5, fraction Xl & 31).
__device__ __forceinline__ uint32_t warp_gather(uint32_t patch_addr, int Xl, int Yl) {
    uint32_t p00, p01, p10, p11;
    uint32_t addr;   // row base + 4 * column as one multiply-add (the compiler's own form masks and adds separately)
    asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(addr) : "r"(Xl >> 5), "r"(patch_addr + static_cast<uint32_t>(Yl >> 5) * (kWpStride * 4)));
    asm("ld.shared.u32 %0, [%4];\n\tld.shared.u32 %1, [%4+4];\n\tld.shared.u32 %2, [%4+%5];\n\tld.shared.u32 %3, [%4+%6];"
        : "=r"(p00), "=r"(p01), "=r"(p10), "=r"(p11)
        : "r"(addr), "n"(kWpStride * 4), "n"(kWpStride * 4 + 4));
    const uint32_t ax = Xl & 31, ay = Yl & 31;
    const uint32_t u = ax * 0xffffu + 32u;             // (32-ax) | ax << 16
    const uint32_t wr1 = u * ay, wr0 = (u << 5) - wr1; // row weights, 16 bits each, sum 1024
    const uint32_t bg0 = __byte_perm(p00, p01, 0x5140), r0 = __byte_perm(p00, p01, 0x6262);
    const uint32_t bg1 = __byte_perm(p10, p11, 0x5140), r1 = __byte_perm(p10, p11, 0x6262);
    const uint32_t bl = __dp2a_lo(wr1, bg1, __dp2a_lo(wr0, bg0, 512u)) >> 10;
    const uint32_t gn = __dp2a_hi(wr1, bg1, __dp2a_hi(wr0, bg0, 512u)) >> 10;
    const uint32_t rd = __dp2a_lo(wr1, r1, __dp2a_lo(wr0, r0, 512u)) >> 10;
    return (3735u * bl + 19235u * gn + 9798u * rd + 16384u) >> 15;
}

__global__ void __launch_bounds__(256, 5) k_warp_board(const uint8_t* __restrict__ img, const double* __restrict__ minv,
                                                    const uint8_t* __restrict__ found, uint8_t* __restrict__ board,
                                                    uint8_t* __restrict__ squares, int H, int W) {
    extern __shared__ __align__(16) uint8_t wsm[];
    uint32_t* patch = reinterpret_cast<uint32_t*>(wsm);
    __shared__ int4 s_ext[8];
    const int b = blockIdx.y, t = threadIdx.x;
    const int bx = (blockIdx.x & 7) * 64, by = (blockIdx.x >> 3) * 64;
    const int row = t >> 2, seg = t & 3;                       // this thread's 16 destination pixels: tile row, segment
    // destination x -> 511 - x: the segment's 16 bytes, mirrored, are one aligned 16-byte store
    uint8_t* dst = board + (static_cast<size_t>(b) * 512 + by + row) * 512 + (448 - bx) + (48 - 16 * seg);
    // optional second copy in the layout extract_squares returns (core.py:420-439): u8 [64 squares][64][64], square = 8*row + col
    uint8_t* sq = squares ? squares + (static_cast<size_t>(b) * 64 + (blockIdx.x >> 3) * 8 + (7 - (blockIdx.x & 7))) * 4096 + row * 64 + (48 - 16 * seg)
                          : nullptr;
    if (!found[b]) {
        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        if (sq) *reinterpret_cast<uint4*>(sq) = make_uint4(0, 0, 0, 0);
        return;
    }
    const double* m = minv + static_cast<size_t>(b) * 9;
    const uint8_t* src = img + static_cast<size_t>(b) * H * W * 3;
    // The segment's anchor in float64.  Only its accuracy matters (2^-40 is ample), not OpenCV's order of roundings: fused
    // multiply-adds, and the quotient by two Newton steps on the hardware seed.
    const double m6 = m[6];
    const double xg = static_cast<double>(bx + 16 * seg + 8), yg = static_cast<double>(by + row);
    const double Xc = fma(m[0], xg, fma(m[1], yg, m[2])), Yc = fma(m[3], xg, fma(m[4], yg, m[5])), Wc = fma(m6, xg, fma(m[7], yg, m[8]));
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(Wc));
    q = fma(q, fma(-Wc, q, 1.0), q);
    q = fma(q, fma(-Wc, q, 1.0), q);
    const double Uc = Xc * q * 262144.0, Vc = Yc * q * 262144.0;                     // image coordinates, 2^-13 units of 1/32 px
    // offset slope 32 * (M0 - U_c * M6) in the same units
    const float Bx = static_cast<float>(fma(-Uc, m6, m[0] * 262144.0)), By = static_cast<float>(fma(-Vc, m6, m[3] * 262144.0));
    const float Wcf = static_cast<float>(Wc), m6f = static_cast<float>(m6);
    // The float32 path is valid for this thread when the offsets at both ends of the segment (the offset is monotonic in d)
    // stay inside the magic range; everything is written so that a NaN fails.
    float ra, rb;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ra) : "f"(fmaf(m6f, -8.0f, Wcf)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rb) : "f"(fmaf(m6f, 7.0f, Wcf)));
    const float oxa = Bx * (-8.0f * ra), oxb = Bx * (7.0f * rb), oya = By * (-8.0f * ra), oyb = By * (7.0f * rb);
    const float Uf = static_cast<float>(Uc), Vf = static_cast<float>(Vc);
    constexpr float kRange = 4000000.0f;   // < 2^22 = 4194304
    bool fast = fabsf(oxa) < kRange && fabsf(oxb) < kRange && fabsf(oya) < kRange && fabsf(oyb) < kRange;
    fast = fast && fabsf(Wcf) > 1e-30f && fabsf(Wcf) < 1e30f && fabsf(m6f * 8.0f) <= 0.25f * fabsf(Wcf);   // W within 25 % over the segment
    fast = fast && fabsf(Uf) < 5e8f && fabsf(Vf) < 5e8f;                                                    // |coordinate| < 2^29: integer arithmetic below is safe
    // source columns / rows the segment touches (floor of the extreme coordinates in pixels; the margins are added below)
    int4 ext = make_int4(-kWpFar, kWpFar, -kWpFar, kWpFar);   // outside the float32 path: a footprint no patch holds
    if (fast) {
        constexpr float kPx = 1.0f / 262144.0f;
        ext.x = __float2int_rd((Uf + fminf(fminf(oxa, oxb), 0.0f)) * kPx);
        ext.y = __float2int_rd((Uf + fmaxf(fmaxf(oxa, oxb), 0.0f)) * kPx);
        ext.z = __float2int_rd((Vf + fminf(fminf(oya, oyb), 0.0f)) * kPx);
        ext.w = __float2int_rd((Vf + fmaxf(fmaxf(oya, oyb), 0.0f)) * kPx);
    }
    ext.x = __reduce_min_sync(0xffffffffu, ext.x);
    ext.y = __reduce_max_sync(0xffffffffu, ext.y);
    ext.z = __reduce_min_sync(0xffffffffu, ext.z);
    ext.w = __reduce_max_sync(0xffffffffu, ext.w);
    if ((t & 31) == 0) s_ext[t >> 5] = ext;
    __syncthreads();
    ext = s_ext[t & 7];
    const int x_min = __reduce_min_sync(0xffffffffu, ext.x), x_max = __reduce_max_sync(0xffffffffu, ext.y);
    const int y_min = __reduce_min_sync(0xffffffffu, ext.z), y_max = __reduce_max_sync(0xffffffffu, ext.w);
    // rint can move a coordinate up by half a unit and the taps reach one further: columns x_min - 1 .. x_max + 2
    const int x_lo = (x_min - 1) & ~15, y_lo = y_min - 1;
    const int ncol16 = (x_max + 3 - x_lo + 15) >> 4;     // columns x_lo .. x_max + 2
    const int nrows = y_max + 3 - y_lo;                  // rows    y_lo .. y_max + 2
    const bool fits = (W & 15) == 0 && static_cast<unsigned>(ncol16 - 1) < kWpCols16 && static_cast<unsigned>(nrows - 1) < kWpRows;
    fast = fast && fits;
    // anchor in patch coordinates + rounding offset + band offset, minus the magic bits: ix = bits(fma) + Cx is
    // rint(U * 2^13) + 2^12 + band / 2.  A thread outside the float32 path gathers its 16 pixels at patch (0, 0) (discarded)
    // and recomputes all of them below.
    constexpr int kRound = (1 << (kWpFrac - 1)) + kWpBand / 2 - kWpMagicBits;
    const int Cx = (fast ? __double2int_rn(Uc - static_cast<double>(x_lo * 262144)) : 0) + kRound;
    const int Cy = (fast ? __double2int_rn(Vc - static_cast<double>(y_lo * 262144)) : 0) + kRound;
    const float fBx = fast ? Bx : 0.0f, fBy = fast ? By : 0.0f, fW = fast ? Wcf : 1.0f, fm6 = fast ? m6f : 0.0f;
    uint32_t redo = fast ? 0u : 0xffffu;   // pixels of the segment that need the literal arithmetic
    const uint32_t patch_addr = static_cast<uint32_t>(__cvta_generic_to_shared(patch));
    if (fits) {
        // 48-byte groups of the footprint: a thread's groups are requested together, so that one round of DRAM latency covers
        // the whole patch instead of one round per group.  Two groups per thread hold 512 (a footprint of up to ~1.2 source
        // pixels per destination pixel); a third round follows for larger ones.
        const int n_groups = nrows * ncol16;   // <= 88 * 7 = 616 <= 3 * 256
        const unsigned inv = kWpInv[ncol16];   // g / ncol16 == (g * inv) >> 16 for g < 9362
        auto stage = [&](auto n_const, int g0) {
            constexpr int kN = decltype(n_const)::value;
            uint4 ld[kN][3];
            int gr[kN], gc[kN];
#pragma unroll
            for (int u = 0; u < kN; ++u) {
                const int g = g0 + 256 * u;
                gr[u] = static_cast<int>((static_cast<unsigned>(g) * inv) >> 16);
                gc[u] = g - gr[u] * ncol16;
                const int gy = y_lo + gr[u], gx = x_lo + 16 * gc[u];
                const bool in = g < n_groups && gy >= 0 && gy < H && gx >= 0 && gx < W;
                const uint4* s4 = reinterpret_cast<const uint4*>(src + (static_cast<size_t>(in ? gy : 0) * W + (in ? gx : 0)) * 3);
                ld[u][0] = in ? __ldg(s4) : make_uint4(0, 0, 0, 0);
                ld[u][1] = in ? __ldg(s4 + 1) : make_uint4(0, 0, 0, 0);
                ld[u][2] = in ? __ldg(s4 + 2) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kN; ++u) {
                if (g0 + 256 * u >= n_groups) break;
                const uint32_t w[12] = {ld[u][0].x, ld[u][0].y, ld[u][0].z, ld[u][0].w, ld[u][1].x, ld[u][1].y, ld[u][1].z, ld[u][1].w,
                                        ld[u][2].x, ld[u][2].y, ld[u][2].z, ld[u][2].w};
                uint32_t o[16];   // pixel k = bytes 3k..3k+2 -> (B,G,R,x); the top byte is never read
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[4 * k + 0] = w[3 * k];
                    o[4 * k + 1] = __byte_perm(w[3 * k], w[3 * k + 1], 0x6543);
                    o[4 * k + 2] = __byte_perm(w[3 * k + 1], w[3 * k + 2], 0x5432);
                    o[4 * k + 3] = w[3 * k + 2] >> 8;
                }
                uint4* d4 = reinterpret_cast<uint4*>(patch + gr[u] * kWpStride + 16 * gc[u]);
#pragma unroll
                for (int k = 0; k < 4; ++k) d4[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
            }
        };
        stage(std::integral_constant<int, 2>{}, t);
        if (n_groups > 512) stage(std::integral_constant<int, 1>{}, t + 512);
    }
    __syncthreads();
    constexpr float kMagic = 12582912.0f;
    uint32_t out[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float d = static_cast<float>(4 * g + i - 8);
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(fm6, d, fW)));
            const float s = d * r;
            const int ix = __float_as_int(fmaf(fBx, s, kMagic)) + Cx, iy = __float_as_int(fmaf(fBy, s, kMagic)) + Cy;
            if ((ix & ((1 << kWpFrac) - 1)) < kWpBand || (iy & ((1 << kWpFrac) - 1)) < kWpBand) redo |= 1u << (4 * g + i);
            // gray (< 256) into byte 3 - i
            packed = __byte_perm(packed, warp_gather(patch_addr, ix >> kWpFrac, iy >> kWpFrac), i == 0 ? 0x4210 : i == 1 ? 0x3410 : i == 2 ? 0x3240 : 0x3214);
        }
        out[3 - g] = packed;
    }
    if (redo) {   // guard-band pixels (0.2 %), threads outside the float32 range, tiles that were not staged
        double X0, Y0, W0;   // row origin of the 64-wide block as cv::WarpPerspectiveInvoker forms it (separate roundings)
        {
            const int y = by + row;
            X0 = m[0] * bx + m[1] * y + m[2];
            Y0 = m[3] * bx + m[4] * y + m[5];
            W0 = m6 * bx + m[7] * y + m[8];
        }
        const unsigned lim_x = fits ? static_cast<unsigned>(ncol16 * 16 - 1) : 0u, lim_y = fits ? static_cast<unsigned>(nrows - 1) : 0u;
        do {
            const int k = __ffs(redo) - 1;
            redo &= redo - 1;
            const double xx = static_cast<double>(16 * seg + k);
            double w = W0 + m6 * xx;
            w = w != 0.0 ? 32.0 / w : 0.0;
            const double fX = fmax(-2147483648.0, fmin(2147483647.0, (X0 + m[0] * xx) * w));
            const double fY = fmax(-2147483648.0, fmin(2147483647.0, (Y0 + m[3] * xx) * w));
            const int Xi = __double2int_rn(fX), Yi = __double2int_rn(fY);
            const int lx = max(-32768, min(32767, Xi >> 5)) - x_lo, ly = max(-32768, min(32767, Yi >> 5)) - y_lo;
            // taps lx, lx + 1 and ly, ly + 1 inside the staged patch: the shared-memory gather; else global memory
            const uint32_t gray = static_cast<unsigned>(lx) < lim_x && static_cast<unsigned>(ly) < lim_y
                                      ? warp_gather(patch_addr, (lx << 5) | (Xi & 31), (ly << 5) | (Yi & 31))
                                      : warp_px_exact(src, H, W, Xi, Yi);
            const int sh = 8 * (3 - (k & 3));
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j == 3 - (k >> 2)) out[j] = (out[j] & ~(0xffu << sh)) | (gray << sh);
        } while (redo);
    }
    *reinterpret_cast<uint4*>(dst) = make_uint4(out[0], out[1], out[2], out[3]);
    if (sq) *reinterpret_cast<uint4*>(sq) = make_uint4(out[0], out[1], out[2], out[3]);
}

}  // namespace

cudaError_t launch_homography(const int32_t* quad, const uint8_t* found, double* minv, int N, float scale, int out_w,
                              int out_h, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    k_homography<<<(N + 63) / 64, 64, 0, s>>>(quad, found, minv, N, scale, out_w, out_h);
    return cudaGetLastError();
}

cudaError_t launch_warp_perspective(const uint8_t* img, int H, int W, int C, const float* corners, double* minv_scratch, uint8_t* out, int out_w,
                                    int out_h, cudaStream_t s) {
    if (out_w <= 0 || out_h <= 0) return cudaSuccess;
    k_homography_f32<<<1, 32, 0, s>>>(corners, minv_scratch, out_w, out_h);
    const int bh0 = out_h < 16 ? out_h : 16;              // cv::WarpPerspectiveInvoker: BLOCK_SZ = 32
    const int bw = (1024 / bh0) < out_w ? (1024 / bh0) : out_w;
    dim3 grid((out_w + 127) / 128, out_h);
    k_warp_generic<<<grid, 128, 0, s>>>(img, H, W, C, minv_scratch, out, out_w, out_h, bw);
    return cudaGetLastError();
}

cudaError_t configure_warp() {
    return cudaFuncSetAttribute(k_warp_board, cudaFuncAttributeMaxDynamicSharedMemorySize, kWpSmem);
}

cudaError_t launch_warp_board(const uint8_t* img, const double* minv, const uint8_t* found, uint8_t* board, uint8_t* squares, int N,
                              int H, int W, cudaStream_t s) {
    if (N == 0) return cudaSuccess;
    dim3 grid(64, N);
    k_warp_board<<<grid, 256, kWpSmem, s>>>(img, minv, found, board, squares, H, W);
    return cudaGetLastError();
}

}  // namespace cvb
