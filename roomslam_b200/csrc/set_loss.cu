// Batched Hungarian matching + set-prediction loss on the GPU.
// Replaces HungarianMatcher.forward (src/benchmark/train.py:21-61: per sample .item(), .cpu().numpy() and scipy's
// linear_sum_assignment -> B host round trips per step) and SetCriterion.forward (train.py:109-187).
//
// match kernel : one warp per sample.  cost[q][m] = -softmax(logits_q)[label_m] + 5 * ||box_q - gt_m||_1 over the
//   sample's VALID colliders, built in fp32 exactly as the torch expression rounds it, then solved in fp64 by the
//   shortest-augmenting-path algorithm for rectangular assignment (D. F. Crouse, IEEE TAES 52(4), 2016 -- the
//   algorithm behind scipy.optimize.linear_sum_assignment): rows = the smaller side, lanes own columns, the Dijkstra
//   frontier minimum is a warp shuffle reduction, duals and paths live in shared memory.  On inputs without exact
//   cost ties the optimum is unique, so the pairs equal scipy's.
// loss kernel  : one warp per sample; cross-entropy, L1 and 1 - GIoU of the matched pairs with analytic gradients
//   (normalised by the batch-wide number of pairs, which only depends on the valid masks), deterministic reductions.
#include <math.h>

#include "common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

constexpr int MAXQ = 128, MAXM = 64, MAXN = 128;      // queries, collider slots, max(Q, M)
constexpr int NCLS = 4;
// Class ids index the NCLS logits of a query: an id outside [0, NCLS) (upstream's CrossEntropyLoss would raise) is clamped
// so that it can never read or write out of bounds.
__device__ __forceinline__ int clamp_label(long long l) { return l < 0 ? 0 : (l >= NCLS ? NCLS - 1 : (int)l); }

__device__ __forceinline__ float class_prob(const float* l, int c) {
    const float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
    const float e0 = expf(l[0] - mx), e1 = expf(l[1] - mx), e2 = expf(l[2] - mx), e3 = expf(l[3] - mx);
    const float s = ((e0 + e1) + e2) + e3;
    const float e = c == 0 ? e0 : c == 1 ? e1 : c == 2 ? e2 : e3;
    return e / s;
}

__global__ void __launch_bounds__(32)
hungarian_match_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ pred_logits,
                       const float* __restrict__ gt_boxes, const long long* __restrict__ gt_labels,
                       const unsigned char* __restrict__ gt_valid, int Q, int M, int K, float w_class, float w_box,
                       int* __restrict__ match_pred, int* __restrict__ match_slot, int* __restrict__ match_rank,
                       int* __restrict__ n_match) {
    extern __shared__ __align__(16) unsigned char raw[];
    float* cost = reinterpret_cast<float*>(raw);                      // [Q][Mv]
    __shared__ double u[MAXN], v[MAXN], spc[MAXN];
    __shared__ int path[MAXN], row4col[MAXN], col4row[MAXN], slot[MAXM];
    __shared__ unsigned char SR[MAXN], SC[MAXN];
    const int b = blockIdx.x, lane = threadIdx.x;
    // compact the valid collider slots (order preserved, like boolean-mask indexing)
    int Mv = 0;
    for (int m0 = 0; m0 < M; m0 += 32) {
        const int m = m0 + lane;
        const bool ok = m < M && gt_valid[(long long)b * M + m];
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (ok) slot[Mv + __popc(bal & ((1u << lane) - 1))] = m;
        Mv += __popc(bal);
    }
    __syncwarp();
    int* mp = match_pred + (long long)b * K;
    int* ms = match_slot + (long long)b * K;
    int* mr = match_rank + (long long)b * K;
    for (int e = lane; e < K; e += 32) { mp[e] = -1; ms[e] = -1; mr[e] = -1; }
    if (Mv == 0) { if (lane == 0) n_match[b] = 0; return; }
    for (int e = lane; e < Q * Mv; e += 32) {
        const int q = e / Mv, m = slot[e % Mv];
        const float* pb = pred_boxes + ((long long)b * Q + q) * 6;
        const float* gb = gt_boxes + ((long long)b * M + m) * 6;
        float d = 0.0f;
#pragma unroll
        for (int k = 0; k < 6; ++k) d = __fadd_rn(d, fabsf(__fsub_rn(pb[k], gb[k])));
        const float p = class_prob(pred_logits + ((long long)b * Q + q) * NCLS, clamp_label(gt_labels[(long long)b * M + m]));
        cost[e] = __fadd_rn(__fmul_rn(w_class, -p), __fmul_rn(w_box, d));
    }
    const bool transposed = Mv < Q;                       // rows = the smaller side
    const int nr = transposed ? Mv : Q, nc = transposed ? Q : Mv;
    auto C = [&](int i, int j) -> double { return (double)(transposed ? cost[j * Mv + i] : cost[i * Mv + j]); };
    for (int e = lane; e < MAXN; e += 32) { u[e] = 0.0; v[e] = 0.0; row4col[e] = -1; col4row[e] = -1; }
    __syncwarp();
    bool failed = false;
    for (int cur = 0; cur < nr && !failed; ++cur) {
        for (int e = lane; e < MAXN; e += 32) { spc[e] = INFINITY; SC[e] = 0; SR[e] = 0; }
        __syncwarp();
        double min_val = 0.0;
        int i = cur, sink = -1;
        while (sink < 0) {
            if (lane == 0) SR[i] = 1;
            double best = INFINITY;
            int best_key = 0x7fffffff;                     // (assigned ? 1 : 0) << 16 | column : prefer free columns on ties
            const double ui = u[i];
            for (int j = lane; j < nc; j += 32) {
                if (SC[j]) continue;
                const double r = min_val + C(i, j) - ui - v[j];
                if (r < spc[j]) { spc[j] = r; path[j] = i; }
                const double s = spc[j];
                const int key = ((row4col[j] != -1) << 16) | j;
                if (s < best || (s == best && key < best_key)) { best = s; best_key = key; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int ok = __shfl_xor_sync(0xffffffffu, best_key, o);
                if (ob < best || (ob == best && ok < best_key)) { best = ob; best_key = ok; }
            }
            if (best_key == 0x7fffffff || !(best < INFINITY)) { failed = true; break; }   // only NaN / inf costs left (scipy raises)
            min_val = best;
            const int j = best_key & 0xffff;
            if (lane == 0) SC[j] = 1;
            __syncwarp();
            if (row4col[j] == -1) sink = j; else i = row4col[j];
        }
        if (failed) break;
        // dual update
        if (lane == 0) u[cur] += min_val;
        for (int r = lane; r < nr; r += 32)
            if (SR[r] && r != cur) u[r] += min_val - spc[col4row[r]];
        for (int j = lane; j < nc; j += 32)
            if (SC[j]) v[j] -= min_val - spc[j];
        __syncwarp();
        // augment along the alternating path
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int r = path[j];
                row4col[j] = r;
                const int prev = col4row[r];
                col4row[r] = j;
                j = prev;
                if (r == cur) break;
            }
        }
        __syncwarp();
    }
    if (failed) { if (lane == 0) n_match[b] = 0; return; }     // non-finite predictions: the sample contributes no pairs
    // emit the pairs ordered by query index (scipy returns them ordered by row of the Q x Mv cost matrix)
    if (lane == 0) {
        int n = 0;
        if (!transposed) {
            for (int q = 0; q < Q; ++q) { mp[n] = q; mr[n] = col4row[q]; ms[n] = slot[col4row[q]]; ++n; }
        } else {
            for (int q = 0; q < Q; ++q)
                if (row4col[q] != -1) { mp[n] = q; mr[n] = row4col[q]; ms[n] = slot[row4col[q]]; ++n; }
        }
        n_match[b] = n;
    }
}

__global__ void count_pairs_kernel(const int* __restrict__ n_match, int B, float* __restrict__ total) {
    __shared__ int red[32];
    int s = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) s += n_match[b];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < (blockDim.x + 31) / 32; ++w) t += red[w];
        total[0] = (float)t;
    }
}

// min / max with torch's tie rule for the gradient (equal operands share it)
__device__ __forceinline__ float share_lt(float a, float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.0f); }

__global__ void __launch_bounds__(32)
set_loss_pairs_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ pred_logits,
                      const float* __restrict__ gt_boxes, const long long* __restrict__ gt_labels, int Q, int M, int K,
                      const int* __restrict__ match_pred, const int* __restrict__ match_slot, const int* __restrict__ n_match,
                      const float* __restrict__ total, float* __restrict__ partial, float* __restrict__ g_logits,
                      float* __restrict__ g_l1, float* __restrict__ g_giou) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int n = n_match[b];
    const float N = total[0];
    const float inv_n = N > 0.0f ? 1.0f / N : 0.0f, inv_6n = N > 0.0f ? 1.0f / (6.0f * N) : 0.0f;
    float s_ce = 0.0f, s_l1 = 0.0f, s_gi = 0.0f;
    for (int e = lane; e < n; e += 32) {
        const int q = match_pred[(long long)b * K + e], m = match_slot[(long long)b * K + e];
        const float* l = pred_logits + ((long long)b * Q + q) * NCLS;
        const float* pb = pred_boxes + ((long long)b * Q + q) * 6;
        const float* gb = gt_boxes + ((long long)b * M + m) * 6;
        const int label = clamp_label(gt_labels[(long long)b * M + m]);
        {   // cross-entropy of the matched query against its collider's label
            const float mx = fmaxf(fmaxf(l[0], l[1]), fmaxf(l[2], l[3]));
            float ex[NCLS], s = 0.0f;
#pragma unroll
            for (int c = 0; c < NCLS; ++c) { ex[c] = expf(l[c] - mx); s += ex[c]; }
            s_ce += logf(s) - (l[label] - mx);
            float* g = g_logits + ((long long)b * Q + q) * NCLS;
#pragma unroll
            for (int c = 0; c < NCLS; ++c) g[c] = (ex[c] / s - (c == label ? 1.0f : 0.0f)) * inv_n;
        }
        float gl[6], gg[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float d = pb[k] - gb[k];
            s_l1 += fabsf(d);
            gl[k] = (d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f)) * inv_6n;
        }
        {   // 1 - GIoU with the analytic gradient w.r.t. the predicted box
            float isz[3], hsz[3], w_ilo[3], w_ihi[3], w_hlo[3], w_hhi[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float alo = pb[k] - pb[3 + k] / 2, ahi = pb[k] + pb[3 + k] / 2;
                const float blo = gb[k] - gb[3 + k] / 2, bhi = gb[k] + gb[3 + k] / 2;
                const float di = fminf(ahi, bhi) - fmaxf(alo, blo);
                const float dh = fmaxf(ahi, bhi) - fminf(alo, blo);
                isz[k] = fmaxf(di, 0.0f); hsz[k] = fmaxf(dh, 0.0f);
                const float pass_i = di >= 0.0f ? 1.0f : 0.0f, pass_h = dh >= 0.0f ? 1.0f : 0.0f;
                w_ihi[k] = pass_i * share_lt(ahi, bhi);            // d isz / d a_hi   (min picks the smaller)
                w_ilo[k] = -pass_i * share_lt(blo, alo);           // d isz / d a_lo   (max picks the larger)
                w_hhi[k] = pass_h * share_lt(bhi, ahi);            // d hsz / d a_hi
                w_hlo[k] = -pass_h * share_lt(alo, blo);           // d hsz / d a_lo
            }
            const float I = isz[0] * isz[1] * isz[2], Hh = hsz[0] * hsz[1] * hsz[2];
            const float Va = pb[3] * pb[4] * pb[5], Vb = gb[3] * gb[4] * gb[5];
            const float U = Va + Vb - I;
            const float iou = I / (U + 1e-6f);
            const float giou = iou - (Hh - U) / (Hh + 1e-6f);
            s_gi += 1.0f - giou;
            const float f_U = -I / ((U + 1e-6f) * (U + 1e-6f)) + 1.0f / (Hh + 1e-6f);
            const float f_I = 1.0f / (U + 1e-6f) - f_U;           // U = Va + Vb - I
            const float f_H = -(U + 1e-6f) / ((Hh + 1e-6f) * (Hh + 1e-6f));
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
                const float gi = f_I * isz[k1] * isz[k2], gh = f_H * hsz[k1] * hsz[k2];
                const float g_hi = gi * w_ihi[k] + gh * w_hhi[k], g_lo = gi * w_ilo[k] + gh * w_hlo[k];
                const float dc = g_hi + g_lo;
                const float ds = 0.5f * g_hi - 0.5f * g_lo + f_U * pb[3 + k1] * pb[3 + k2];
                gg[k] = -dc * inv_n;                               // loss = 1 - giou
                gg[3 + k] = -ds * inv_n;
            }
        }
        float* o1 = g_l1 + ((long long)b * Q + q) * 6;
        float* o2 = g_giou + ((long long)b * Q + q) * 6;
#pragma unroll
        for (int k = 0; k < 6; ++k) { o1[k] = gl[k]; o2[k] = gg[k]; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s_ce += __shfl_xor_sync(0xffffffffu, s_ce, o);
        s_l1 += __shfl_xor_sync(0xffffffffu, s_l1, o);
        s_gi += __shfl_xor_sync(0xffffffffu, s_gi, o);
    }
    if (lane == 0) { partial[b * 3] = s_ce; partial[b * 3 + 1] = s_l1; partial[b * 3 + 2] = s_gi; }
}

// losses[0..3] = class, l1, giou, weighted total; fixed summation order -> run-to-run identical
__global__ void set_loss_finalize_kernel(const float* __restrict__ partial, int B, const float* __restrict__ total, float w_class,
                                         float w_l1, float w_giou, float* __restrict__ losses) {
    __shared__ double red[3][8];
    double s[3] = {0.0, 0.0, 0.0};
    for (int b = threadIdx.x; b < B; b += 256)
        for (int k = 0; k < 3; ++k) s[k] += partial[b * 3 + k];
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t[3] = {0.0, 0.0, 0.0};
        for (int k = 0; k < 3; ++k)
            for (int w = 0; w < 8; ++w) t[k] += red[k][w];
        const float N = total[0];
        const float ce = N > 0.0f ? (float)(t[0] / N) : 0.0f;
        const float l1 = N > 0.0f ? (float)(t[1] / (6.0 * N)) : 0.0f;
        const float gi = N > 0.0f ? (float)(t[2] / N) : 0.0f;
        losses[0] = ce; losses[1] = l1; losses[2] = gi;
        losses[3] = w_class * ce + w_l1 * l1 + w_giou * gi;
    }
}

}  // namespace

extern "C" int rs_hungarian_match(const float* pred_boxes, const float* pred_logits, const float* gt_boxes,
                                  const int64_t* gt_labels, const unsigned char* gt_valid, int B, int Q, int M, float w_class,
                                  float w_box, int* match_pred, int* match_slot, int* match_rank, int* n_match, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(pred_boxes && pred_logits && gt_boxes && gt_labels && gt_valid && match_pred && match_slot && match_rank && n_match,
               "rs_hungarian_match: null pointer");
    RS_REQUIRE(Q >= 1 && Q <= MAXQ && M >= 1 && M <= MAXM, "rs_hungarian_match: need 1 <= Q <= %d queries and 1 <= M <= %d slots", MAXQ, MAXM);
    const size_t smem = (size_t)Q * M * sizeof(float);
    RS_CUDA_OK(cudaFuncSetAttribute(hungarian_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    hungarian_match_kernel<<<B, 32, smem, stream>>>(pred_boxes, pred_logits, gt_boxes, reinterpret_cast<const long long*>(gt_labels),
                                                    gt_valid, Q, M, Q < M ? Q : M, w_class, w_box, match_pred, match_slot, match_rank,
                                                    n_match);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_set_loss_f32(const float* pred_boxes, const float* pred_logits, const float* gt_boxes, const int64_t* gt_labels,
                               int B, int Q, int M, const int* match_pred, const int* match_slot, const int* n_match,
                               float w_class, float w_l1, float w_giou, float* workspace, float* losses, float* g_logits,
                               float* g_l1, float* g_giou, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) {                                            // empty batch: all losses are zero (train.py:146-147,165-166)
        RS_REQUIRE(losses, "rs_set_loss_f32: null pointer");
        RS_CUDA_OK(cudaMemsetAsync(losses, 0, 4 * sizeof(float), stream));
        return 0;
    }
    RS_REQUIRE(pred_boxes && pred_logits && gt_boxes && gt_labels && match_pred && match_slot && n_match && workspace && losses &&
                   g_logits && g_l1 && g_giou, "rs_set_loss_f32: null pointer");
    RS_REQUIRE(Q >= 1 && M >= 1 && B >= 0, "rs_set_loss_f32: bad sizes");
    RS_CUDA_OK(cudaMemsetAsync(g_logits, 0, sizeof(float) * (size_t)B * Q * NCLS, stream));
    RS_CUDA_OK(cudaMemsetAsync(g_l1, 0, sizeof(float) * (size_t)B * Q * 6, stream));
    RS_CUDA_OK(cudaMemsetAsync(g_giou, 0, sizeof(float) * (size_t)B * Q * 6, stream));
    float* total = workspace;              // [1]
    float* partial = workspace + 1;        // [B][3]
    count_pairs_kernel<<<1, 256, 0, stream>>>(n_match, B, total);
    rs::count_launch();
    if (B > 0) {
        set_loss_pairs_kernel<<<B, 32, 0, stream>>>(pred_boxes, pred_logits, gt_boxes, reinterpret_cast<const long long*>(gt_labels), Q, M,
                                                    Q < M ? Q : M, match_pred, match_slot, n_match, total, partial, g_logits, g_l1,
                                                    g_giou);
        rs::count_launch();
    }
    set_loss_finalize_kernel<<<1, 256, 0, stream>>>(partial, B, total, w_class, w_l1, w_giou, losses);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}
