// Batched evaluation on the GPU: matched-pair metrics (src/benchmark/train.py:234-328), per-class
// greedy NMS with the confidence filter (src/benchmark/inference.py:87-170) and true-positive flags for mAP
// (README.md:127-132).  The shipped code walks box pairs in Python with one .item() host sync per pair; here one warp
// owns one scene.  IoUs follow the reference's fp32 operation order (file compiled with -fmad=false), so threshold
// decisions agree; counts are integers, reductions run in a fixed order.
#include <math.h>

#include "common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

constexpr int MAXQ = 128, MAXM = 64, NCLS = 4;

__device__ __forceinline__ float iou_aabb(const float* a, const float* b) {      // train.py:277-292 = inference.py:60-84
    float inter = 1.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float lo = fmaxf(a[k] - a[3 + k] / 2, b[k] - b[3 + k] / 2);
        const float hi = fminf(a[k] + a[3 + k] / 2, b[k] + b[3 + k] / 2);
        inter = inter * fmaxf(hi - lo, 0.0f);
    }
    const float va = (a[3] * a[4]) * a[5], vb = (b[3] * b[4]) * b[5];
    return inter / (((va + vb) - inter) + 1e-6f);
}

// confidence = max softmax probability, label = first arg max (torch.max / argmax)
__device__ __forceinline__ void conf_label(const float* l, float& conf, int& label) {
    int best = 0;
#pragma unroll
    for (int c = 1; c < NCLS; ++c) if (l[c] > l[best]) best = c;
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < NCLS; ++c) s += expf(l[c] - l[best]);
    conf = 1.0f / s;
    label = best;
}

// per scene: [iou_sum, iou_cnt, tp, fp, fn, cls_correct, cls_total]
__global__ void __launch_bounds__(32)
eval_pairs_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ pred_logits, const float* __restrict__ gt_boxes,
                  const long long* __restrict__ gt_labels, const unsigned char* __restrict__ gt_valid, int Q, int M, int K,
                  const int* __restrict__ match_pred, const int* __restrict__ match_slot, const int* __restrict__ n_match,
                  float iou_thresh, double* __restrict__ partial) {
    const int b = blockIdx.x, lane = threadIdx.x;
    const int n = n_match[b];
    int nvalid = 0;
    for (int m = lane; m < M; m += 32) nvalid += gt_valid[(long long)b * M + m] ? 1 : 0;
    float iou_sum = 0.0f;
    int tp = 0, fp = 0, ok = 0;
    for (int e = lane; e < n; e += 32) {
        const int q = match_pred[(long long)b * K + e], m = match_slot[(long long)b * K + e];
        const float v = iou_aabb(pred_boxes + ((long long)b * Q + q) * 6, gt_boxes + ((long long)b * M + m) * 6);
        iou_sum += v;
        if (v >= iou_thresh) ++tp; else ++fp;
        float conf; int label;
        conf_label(pred_logits + ((long long)b * Q + q) * NCLS, conf, label);
        ok += (label == (int)gt_labels[(long long)b * M + m]) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        iou_sum += __shfl_xor_sync(0xffffffffu, iou_sum, o);
        tp += __shfl_xor_sync(0xffffffffu, tp, o);
        fp += __shfl_xor_sync(0xffffffffu, fp, o);
        ok += __shfl_xor_sync(0xffffffffu, ok, o);
        nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    }
    if (lane == 0) {
        double* p = partial + (long long)b * 7;
        p[0] = iou_sum; p[1] = n; p[2] = tp; p[3] = fp; p[4] = nvalid > n ? nvalid - n : 0; p[5] = ok; p[6] = n;
    }
}

// counts[7] += column sums of partial[B][7]  (single CTA, fixed order)
__global__ void eval_accumulate_kernel(const double* __restrict__ partial, int B, double* __restrict__ counts) {
    __shared__ double red[7][8];
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < B; b += 256)
        for (int k = 0; k < 7; ++k) s[k] += partial[(long long)b * 7 + k];
    for (int k = 0; k < 7; ++k) {
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        counts[threadIdx.x] += t;
    }
}

// rank of every eligible query inside its (class, descending confidence) group; ties broken by query index
__global__ void __launch_bounds__(32)
nms_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ pred_logits, int Q, float conf_thr, float nms_thr,
           int* __restrict__ keep_idx, int* __restrict__ n_keep, float* __restrict__ conf_out, int* __restrict__ label_out) {
    __shared__ float conf[MAXQ];
    __shared__ int label[MAXQ], sorted[MAXQ];
    __shared__ unsigned char alive[MAXQ];
    const int b = blockIdx.x, lane = threadIdx.x;
    const float* boxes = pred_boxes + (long long)b * Q * 6;
    for (int q = lane; q < Q; q += 32) {
        float c; int l;
        conf_label(pred_logits + ((long long)b * Q + q) * NCLS, c, l);
        conf[q] = c; label[q] = l;
        conf_out[(long long)b * Q + q] = c; label_out[(long long)b * Q + q] = l;
        keep_idx[(long long)b * Q + q] = -1;
    }
    __syncwarp();
    int emitted = 0;
    for (int cls = 0; cls < NCLS; ++cls) {
        // members of this class above the confidence threshold, sorted by descending confidence (rank by counting)
        int cnt = 0;
        for (int q0 = 0; q0 < Q; q0 += 32) {
            const int q = q0 + lane;
            const bool in = q < Q && label[q] == cls && conf[q] > conf_thr;
            if (in) {
                int rank = 0;
                for (int r = 0; r < Q; ++r)
                    if (label[r] == cls && conf[r] > conf_thr && (conf[r] > conf[q] || (conf[r] == conf[q] && r < q))) ++rank;
                sorted[rank] = q;
                alive[rank] = 1;
            }
            cnt += __popc(__ballot_sync(0xffffffffu, in));
        }
        __syncwarp();
        for (int i = 0; i < cnt; ++i) {
            if (!alive[i]) continue;                                  // uniform across the warp (shared memory)
            const int cur = sorted[i];
            if (lane == 0) keep_idx[(long long)b * Q + emitted] = cur;
            ++emitted;
            for (int j = i + 1 + lane; j < cnt; j += 32)
                if (alive[j] && !(iou_aabb(boxes + cur * 6, boxes + sorted[j] * 6) < nms_thr)) alive[j] = 0;
            __syncwarp();
        }
        __syncwarp();
    }
    if (lane == 0) n_keep[b] = emitted;
}

// mAP flags: predictions of a scene in descending confidence claim the best-IoU free collider of their class
__global__ void __launch_bounds__(32)
ap_flags_kernel(const float* __restrict__ pred_boxes, const float* __restrict__ pred_logits, const float* __restrict__ gt_boxes,
                const long long* __restrict__ gt_labels, const unsigned char* __restrict__ gt_valid, int Q, int M, float iou_thr,
                int* __restrict__ flags, float* __restrict__ conf_out, int* __restrict__ label_out, int* __restrict__ n_gt) {
    __shared__ float conf[MAXQ];
    __shared__ int label[MAXQ], sorted[MAXQ];
    __shared__ unsigned char claimed[MAXM];
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int q = lane; q < Q; q += 32) {
        float c; int l;
        conf_label(pred_logits + ((long long)b * Q + q) * NCLS, c, l);
        conf[q] = c; label[q] = l;
        conf_out[(long long)b * Q + q] = c; label_out[(long long)b * Q + q] = l;
    }
    for (int m = lane; m < M; m += 32) {
        claimed[m] = 0;
        const long long lab = gt_labels[(long long)b * M + m];           // a class id outside [0, NCLS) is counted nowhere
        if (gt_valid[(long long)b * M + m] && lab >= 0 && lab < NCLS) atomicAdd(&n_gt[(int)lab], 1);
    }
    __syncwarp();
    for (int q = lane; q < Q; q += 32) {
        int rank = 0;
        for (int r = 0; r < Q; ++r)
            if (conf[r] > conf[q] || (conf[r] == conf[q] && r < q)) ++rank;
        sorted[rank] = q;
    }
    __syncwarp();
    for (int i = 0; i < Q; ++i) {
        const int q = sorted[i];
        float best = -1.0f;
        int best_m = 0x7fffffff;
        for (int m = lane; m < M; m += 32) {
            if (!gt_valid[(long long)b * M + m] || claimed[m] || (int)gt_labels[(long long)b * M + m] != label[q]) continue;
            const float v = iou_aabb(pred_boxes + ((long long)b * Q + q) * 6, gt_boxes + ((long long)b * M + m) * 6);
            if (v > best) { best = v; best_m = m; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int om = __shfl_xor_sync(0xffffffffu, best_m, o);
            if (ob > best || (ob == best && om < best_m)) { best = ob; best_m = om; }
        }
        const bool hit = best_m != 0x7fffffff && best >= iou_thr;
        if (lane == 0) {
            flags[(long long)b * Q + q] = hit ? 1 : 0;
            if (hit) claimed[best_m] = 1;
        }
        __syncwarp();
    }
}


// ---- evaluation of the README GRU model (fixed object slots + validity head, README.md:93-132; BASELINE config 5) ----
// One thread per (trace, slot): 2-D axis-aligned IoU (orientation ignored, as in all shipped IoU code), predicted class,
// confidence = sigmoid(validity logit) * max class probability, true-positive flag (target slot valid, class right,
// IoU >= thr); block partials of [IoU sum over valid slots, valid slots, class hits, validity hits, TP, predicted valid].
__global__ void __launch_bounds__(256)
slot_eval_kernel(const float* __restrict__ cls, const float* __restrict__ pos, const float* __restrict__ size,
                 const float* __restrict__ vlogit, const long long* __restrict__ t_cls, const float* __restrict__ t_pos,
                 const float* __restrict__ t_size, const float* __restrict__ t_valid, long long total, int C, float iou_thr,
                 float* __restrict__ conf, int* __restrict__ label, int* __restrict__ flag, int* __restrict__ n_gt,
                 double* __restrict__ partial) {
    __shared__ double red[6][8];
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const float* l = cls + e * C;
        int best = 0;
        for (int c = 1; c < C; ++c) if (l[c] > l[best]) best = c;
        float den = 0.0f;
        for (int c = 0; c < C; ++c) den += expf(l[c] - l[best]);
        const float pv = 1.0f / (1.0f + expf(-vlogit[e]));
        float inter = 1.0f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float lo = fmaxf(pos[e * 2 + k] - size[e * 2 + k] / 2, t_pos[e * 2 + k] - t_size[e * 2 + k] / 2);
            const float hi = fminf(pos[e * 2 + k] + size[e * 2 + k] / 2, t_pos[e * 2 + k] + t_size[e * 2 + k] / 2);
            inter = inter * fmaxf(hi - lo, 0.0f);
        }
        const float uni = (size[e * 2] * size[e * 2 + 1] + t_size[e * 2] * t_size[e * 2 + 1]) - inter;
        const float iou = inter / fmaxf(uni, 1e-9f);
        const bool tv = t_valid[e] > 0.5f;
        const int tc = (int)t_cls[e];
        const bool hit = tv && best == tc && iou >= iou_thr;
        conf[e] = pv / den;
        label[e] = best;
        flag[e] = hit ? 1 : 0;
        if (tv) {
            if (tc >= 0 && tc < C) atomicAdd(&n_gt[tc], 1);              // out-of-range class ids never index the counters
            s[0] += iou; s[1] += 1.0; s[2] += (best == tc) ? 1.0 : 0.0;
        }
        s[3] += ((pv > 0.5f) == tv) ? 1.0 : 0.0;
        s[4] += hit ? 1.0 : 0.0;
        s[5] += (pv > 0.5f) ? 1.0 : 0.0;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = 0; k < 6; ++k) {
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if (lane == 0) red[k][warp] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        partial[(long long)blockIdx.x * 6 + threadIdx.x] = t;
    }
}

__global__ void slot_eval_accumulate_kernel(const double* __restrict__ partial, int n, double* __restrict__ counts) {
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int b = 0; b < n; ++b) t += partial[(long long)b * 6 + threadIdx.x];
        counts[threadIdx.x] += t;
    }
}

}  // namespace

extern "C" int rs_eval_pairs(const float* pred_boxes, const float* pred_logits, const float* gt_boxes, const int64_t* gt_labels,
                             const unsigned char* gt_valid, int B, int Q, int M, const int* match_pred, const int* match_slot,
                             const int* n_match, float iou_thresh, double* workspace, double* counts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(pred_boxes && pred_logits && gt_boxes && gt_labels && gt_valid && match_pred && match_slot && n_match && workspace &&
                   counts, "rs_eval_pairs: null pointer");
    eval_pairs_kernel<<<B, 32, 0, stream>>>(pred_boxes, pred_logits, gt_boxes, reinterpret_cast<const long long*>(gt_labels), gt_valid,
                                            Q, M, Q < M ? Q : M, match_pred, match_slot, n_match, iou_thresh, workspace);
    rs::count_launch();
    eval_accumulate_kernel<<<1, 256, 0, stream>>>(workspace, B, counts);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_nms_3d(const float* pred_boxes, const float* pred_logits, int B, int Q, float conf_thr, float nms_thr,
                         int* keep_idx, int* n_keep, float* conf, int* label, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(pred_boxes && pred_logits && keep_idx && n_keep && conf && label, "rs_nms_3d: null pointer");
    RS_REQUIRE(Q >= 1 && Q <= MAXQ, "rs_nms_3d: need 1 <= Q <= %d", MAXQ);
    nms_kernel<<<B, 32, 0, stream>>>(pred_boxes, pred_logits, Q, conf_thr, nms_thr, keep_idx, n_keep, conf, label);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_ap_flags(const float* pred_boxes, const float* pred_logits, const float* gt_boxes, const int64_t* gt_labels,
                           const unsigned char* gt_valid, int B, int Q, int M, float iou_thr, int* flags, float* conf, int* label,
                           int* n_gt, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(pred_boxes && pred_logits && gt_boxes && gt_labels && gt_valid && flags && conf && label && n_gt, "rs_ap_flags: null pointer");
    RS_REQUIRE(Q >= 1 && Q <= MAXQ && M >= 1 && M <= MAXM, "rs_ap_flags: need Q <= %d and M <= %d", MAXQ, MAXM);
    ap_flags_kernel<<<B, 32, 0, stream>>>(pred_boxes, pred_logits, gt_boxes, reinterpret_cast<const long long*>(gt_labels), gt_valid, Q,
                                          M, iou_thr, flags, conf, label, n_gt);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_slot_eval(const float* class_logits, const float* positions, const float* sizes, const float* validity_logits,
                            const int64_t* t_classes, const float* t_positions, const float* t_sizes, const float* t_valid,
                            int64_t n_slots, int C, float iou_thr, float* conf, int* label, int* flag, int* n_gt,
                            double* workspace, double* counts, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (n_slots == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(class_logits && positions && sizes && validity_logits && t_classes && t_positions && t_sizes && t_valid && conf &&
                   label && flag && n_gt && workspace && counts, "rs_slot_eval: null pointer");
    RS_REQUIRE(C >= 1 && C <= 64 && n_slots >= 0, "rs_slot_eval: bad sizes");
    long long blocks = (n_slots + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    slot_eval_kernel<<<(int)blocks, 256, 0, stream>>>(class_logits, positions, sizes, validity_logits,
                                                      reinterpret_cast<const long long*>(t_classes), t_positions, t_sizes, t_valid,
                                                      n_slots, C, iou_thr, conf, label, flag, n_gt, workspace);
    rs::count_launch();
    slot_eval_accumulate_kernel<<<1, 32, 0, stream>>>(workspace, (int)blocks, counts);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}
