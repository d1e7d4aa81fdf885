// On-GPU trace preprocessing: the 11-D kinematic features the shipped pipeline computes in
// numpy per item and per epoch: src/benchmark/dataloader.py:410-457 (_process_traces) = src/benchmark/inference.py:24-57
// (process_traces), followed by the pad-to-batch-max collate of dataloader.py:510-559.
//
//   a[i]    = (x, y, z, t - t_0)                    fp32
//   diff[i] = a[i] - a[i-1]   (diff[0] = 0)         np.diff(..., prepend=first row)
//   dt[i]   = max(diff_t[i], 1e-3f)                 np.clip(diffs[:, 3], 1e-3, None)
//   vel[i]  = diff_xyz[i] / dt[i]                   IEEE fp32 division
//   acc[i]  = vel[i] - vel[i-1]   (acc[0] = 0)
//   speed[i]= sqrt((vx^2 + vy^2) + vz^2)            np.linalg.norm(vel, axis=1): products, left-to-right adds, sqrt
//   row     = [x, y, z, t, vx, vy, vz, ax, ay, az, speed]
//   traces longer than max_len keep the rows idx[j] = (int)(j * ((N-1)/(max_len-1))) in float64, idx[last] = N-1
//   (np.linspace(0, N-1, max_len, dtype=int)); an empty trace yields one zero row.
// Every operation is an explicit round-to-nearest intrinsic and the file is compiled with -fmad=false: the result is
// bit-identical to the numpy reference.  One thread per output row (a row needs the three source points i-2 .. i,
// served by L1); each warp stages its 32 rows in shared memory and writes them back as contiguous 16-byte stores
// (warp-level synchronisation only).
#include "common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

struct P4 { float x, y, z, t; };

__device__ __forceinline__ P4 load_norm(const float4* pts, long long i, float t0) {
    const float4 v = __ldg(pts + i);
    P4 p; p.x = v.x; p.y = v.y; p.z = v.z; p.t = __fsub_rn(v.w, t0);
    return p;
}

__device__ __forceinline__ void velocity(const P4& a, const P4& b, bool first, float* v) {   // vel at the point a (b = previous)
    const float dx = first ? 0.0f : __fsub_rn(a.x, b.x), dy = first ? 0.0f : __fsub_rn(a.y, b.y);
    const float dz = first ? 0.0f : __fsub_rn(a.z, b.z), dtr = first ? 0.0f : __fsub_rn(a.t, b.t);
    const float dt = fmaxf(dtr, 1e-3f);
    v[0] = __fdiv_rn(dx, dt); v[1] = __fdiv_rn(dy, dt); v[2] = __fdiv_rn(dz, dt);
}

__global__ void __launch_bounds__(256)
trace_features_kernel(const float4* __restrict__ pts, const long long* __restrict__ offsets, int B, int max_len, int out_len,
                      float* __restrict__ feats, unsigned char* __restrict__ mask, long long* __restrict__ lengths,
                      int* __restrict__ unsorted_flag, int vec_ok) {
    __shared__ __align__(16) float rows_all[256 * 11];           // per warp: 32 output rows (1408 B), stored coalesced below
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* rows = rows_all + warp * 32 * 11;
    const long long total = (long long)B * out_len;
    const long long groups = (total + 31) / 32;                  // one warp per group of 32 consecutive output rows
    for (long long gidx = (long long)blockIdx.x * 8 + warp; gidx < groups; gidx += (long long)gridDim.x * 8) {
        const long long e = gidx * 32 + lane;
        if (e < total) {
            const int b = (int)(e / out_len), j = (int)(e % out_len);
            const long long o0 = offsets[b];
            const long long N = offsets[b + 1] - o0;
            const long long L = N == 0 ? 1 : (N > max_len ? max_len : N);
            if (j == 0) lengths[b] = L;
            float row[11];
#pragma unroll
            for (int k = 0; k < 11; ++k) row[k] = 0.0f;
            const bool valid = j < L;
            if (valid && N > 0) {
                long long i = j;
                if (N > max_len) {
                    const double step = __ddiv_rn((double)(N - 1), (double)(max_len - 1));
                    i = (j == max_len - 1) ? (N - 1) : (long long)__dmul_rn((double)j, step);
                    // time order must hold over ALL source points, not only the kept rows (the reference argsorts before it
                    // differences and down-samples, inference.py:38-39): this thread checks the points skipped since row j - 1
                    const long long ip = j == 0 ? 0 : (long long)__dmul_rn((double)(j - 1), step);
                    for (long long k = ip + 1; k < i; ++k)
                        if (__ldg(pts + o0 + k).w < __ldg(pts + o0 + k - 1).w) *unsorted_flag = 1;
                }
                const float t0 = __ldg(pts + o0).w;
                const P4 a = load_norm(pts, o0 + i, t0);
                float v[3] = {0.f, 0.f, 0.f}, vp[3] = {0.f, 0.f, 0.f};
                if (i >= 1) {
                    const P4 p1 = load_norm(pts, o0 + i - 1, t0);
                    velocity(a, p1, false, v);
                    if (__ldg(pts + o0 + i).w < __ldg(pts + o0 + i - 1).w) *unsorted_flag = 1;
                    if (i >= 2) {
                        const P4 p2 = load_norm(pts, o0 + i - 2, t0);
                        velocity(p1, p2, false, vp);
                    }                                        // i == 1: vel[0] = 0 / dt = 0
                }
                row[0] = a.x; row[1] = a.y; row[2] = a.z; row[3] = a.t;
                row[4] = v[0]; row[5] = v[1]; row[6] = v[2];
                row[7] = (i >= 1) ? __fsub_rn(v[0], vp[0]) : 0.0f;
                row[8] = (i >= 1) ? __fsub_rn(v[1], vp[1]) : 0.0f;
                row[9] = (i >= 1) ? __fsub_rn(v[2], vp[2]) : 0.0f;
                row[10] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
            }
#pragma unroll
            for (int k = 0; k < 11; ++k) rows[lane * 11 + k] = row[k];           // stride 11 words: conflict-free
            mask[e] = valid ? 1 : 0;
        }
        __syncwarp();
        const long long nrows = (total - gidx * 32 < 32) ? (total - gidx * 32) : 32;
        const int nfl = (int)nrows * 11;
        float* dst = feats + gidx * 32 * 11;                     // 32 * 44 B per group: 16-byte aligned when feats is
        if (vec_ok) {
            for (int q = lane; q < nfl / 4; q += 32)
                reinterpret_cast<float4*>(dst)[q] = reinterpret_cast<const float4*>(rows)[q];
            for (int q = (nfl & ~3) + lane; q < nfl; q += 32) dst[q] = rows[q];
        } else {
            for (int q = lane; q < nfl; q += 32) dst[q] = rows[q];
        }
        __syncwarp();
    }
}

// ---- uniform-rate resampling + windowing for the GRU input (decision D14: README.md:145 "10 Hz", windows of seq_len) ----
// One thread per output sample: t_i of numpy's arange(t_first, t_last, 1/hz) (element 0 = start, 1 = start + step,
// i >= 2: start + i * ((start + step) - start), as numpy fills it), binary search of the bracketing source points, and
// np.interp's  slope * (t - t_j) + f_j  in fp64 without FMA contraction, for the floor-plane coordinates (x, z).
__global__ void __launch_bounds__(256)
resample_windows_kernel(const double* __restrict__ pts, const long long* __restrict__ offsets, const long long* __restrict__ win_trace,
                        const long long* __restrict__ win_start, long long total, int seq_len, double step,
                        float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long w = e / seq_len;
        const long long i = win_start[w] + e % seq_len;
        const long long b = win_trace[w];
        const double* tr = pts + offsets[b] * 4;
        const long long n = offsets[b + 1] - offsets[b];
        const double start = tr[3];
        const double next = __dadd_rn(start, step);
        const double delta = __dsub_rn(next, start);
        const double t = i == 0 ? start : (i == 1 ? next : __dadd_rn(start, __dmul_rn((double)i, delta)));
        // largest j with t_j <= t (t lies inside [t_0, t_{n-1}) by construction)
        long long lo = 0, hi = n - 1;
        while (hi - lo > 1) {
            const long long mid = (lo + hi) >> 1;
            if (tr[mid * 4 + 3] <= t) lo = mid; else hi = mid;
        }
        float v[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int col = c == 0 ? 0 : 2;                        // x and z: the floor plane (y is height)
            const double f0 = tr[lo * 4 + col], x0 = tr[lo * 4 + 3];
            double r = f0;
            if (t != x0 && lo + 1 < n) {
                const double slope = __ddiv_rn(__dsub_rn(tr[(lo + 1) * 4 + col], f0), __dsub_rn(tr[(lo + 1) * 4 + 3], x0));
                r = __dadd_rn(__dmul_rn(slope, __dsub_rn(t, x0)), f0);
            }
            v[c] = (float)r;
        }
        out[e * 2] = v[0];
        out[e * 2 + 1] = v[1];
    }
}

}  // namespace

extern "C" int rs_trace_features(const float* pts, const int64_t* offsets, int B, int max_len, int out_len, float* feats,
                                 unsigned char* mask, int64_t* lengths, int* unsorted_flag, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) {                                            // empty batch: empty tensors carry null pointers
        if (unsorted_flag) RS_CUDA_OK(cudaMemsetAsync(unsorted_flag, 0, sizeof(int), stream));
        return 0;
    }
    RS_REQUIRE(offsets && feats && mask && lengths && unsorted_flag, "rs_trace_features: null pointer");
    RS_REQUIRE(B >= 0 && max_len >= 2 && out_len >= 1, "rs_trace_features: need max_len >= 2 and out_len >= 1");
    RS_REQUIRE((reinterpret_cast<uintptr_t>(pts) & 15) == 0, "rs_trace_features: points must be 16-byte aligned (x, y, z, t rows)");
    RS_CUDA_OK(cudaMemsetAsync(unsorted_flag, 0, sizeof(int), stream));
    const long long total = (long long)B * out_len;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;                       // 8 resident CTAs per SM, chunk-strided
    const int vec_ok = (reinterpret_cast<uintptr_t>(feats) & 15) == 0;
    trace_features_kernel<<<(int)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(pts),
                                                           reinterpret_cast<const long long*>(offsets), B, max_len, out_len, feats,
                                                           mask, reinterpret_cast<long long*>(lengths), unsorted_flag, vec_ok);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_resample_windows_f64(const double* pts, const int64_t* offsets, const int64_t* win_trace, const int64_t* win_start,
                                       int64_t n_windows, int seq_len, double step, float* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (n_windows == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(pts && offsets && win_trace && win_start && out && seq_len >= 1 && step > 0.0, "rs_resample_windows_f64: bad arguments");
    const long long total = (long long)n_windows * seq_len;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    resample_windows_kernel<<<(int)blocks, 256, 0, stream>>>(pts, reinterpret_cast<const long long*>(offsets),
                                                             reinterpret_cast<const long long*>(win_trace),
                                                             reinterpret_cast<const long long*>(win_start), total, seq_len, step, out);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}
