// fp32 GRU recurrence, forward and backward-through-time (fp32 parity mode,
// tolerance 1e-4 against torch.nn.GRU on the CPU).  Generic in hidden size H (multiple of 32, <= 512).
//
// Gate equations (torch/nn/modules/rnn.py:1221-1224, weight rows ordered r|z|n):
//   r = sigma(Px_r + Wh_r h + b_hr)   z = sigma(Px_z + Wh_z h + b_hz)
//   n = tanh(Px_n + r * (Wh_n h + b_hn))   h' = (1 - z) * n + z * h        Px = W_ih x + b_ih
//
// One CTA owns a tile of Bt = S*R traces of ONE direction for all T steps (persistent over time).
// blockDim = (H, S): thread (i, y) owns hidden unit i of traces y*R .. y*R+R-1.  W_hh (3H x H fp32) stays
// resident in shared memory for the whole sequence when it fits (H <= 128), otherwise it streams from L2.
// The hidden state is double-buffered in shared memory: one __syncthreads() per time step.
// Layer 0 (input size <= 4) fuses the input projection (K = 2 is not a GEMM); deeper layers read the
// time-parallel projection P computed by a GEMM.
//
// Sequence buffers are addressed as element(b, t, c) = p[((b*trace_rows + row0 + t) * ld) + c] so that the
// padded activation layout (one zero row before and after every trace, used by the shifted weight-gradient
// GEMM) and plain (B, T, C) user tensors go through the same code.
#include "common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

struct Seq {
    float* p;
    long long ld, trace_rows, row0;
    __device__ __forceinline__ float* at(long long b, long long t) const { return p + ((b * trace_rows + row0 + t) * ld); }
};

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int R, bool WSMEM, int MAXT>
__global__ void __launch_bounds__(MAXT)
gru_fwd_f32_kernel(Seq x, int I, const float* __restrict__ w_ih, const float* __restrict__ b_ih, Seq P,
                   const float* __restrict__ w_hh_t, const float* __restrict__ b_hh, Seq out,
                   float* __restrict__ h_n, float* __restrict__ gates, const int* __restrict__ lengths, int B, int T, int H) {
    extern __shared__ __align__(16) float smem_f[];
    const int i = threadIdx.x, y = threadIdx.y, S = blockDim.y, Bt = S * R;
    const int dir = blockIdx.y;
    const int H3 = 3 * H;
    float* Ws = smem_f;
    float* hs = smem_f + (WSMEM ? H * H3 : 0);   // [2][Bt][H]
    const float* Wg = w_hh_t + (size_t)dir * H * H3;
    if (WSMEM) {
        for (int e = threadIdx.y * blockDim.x + threadIdx.x; e < H * H3; e += blockDim.x * blockDim.y) Ws[e] = Wg[e];
    }
    for (int e = threadIdx.y * blockDim.x + threadIdx.x; e < 2 * Bt * H; e += blockDim.x * blockDim.y) hs[e] = 0.0f;

    float wi[3][4], bi[3], bh[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
        bh[g] = b_hh[dir * H3 + g * H + i];
        bi[g] = b_ih[dir * H3 + g * H + i];
#pragma unroll
        for (int c = 0; c < 4; ++c) wi[g][c] = (P.p == nullptr && c < I) ? w_ih[((size_t)dir * H3 + g * H + i) * I + c] : 0.0f;
    }
    __syncthreads();

    const long long b0 = (long long)blockIdx.x * Bt + y * R;
    int len[R];                                    // valid steps of each trace (packed-sequence semantics); T when no lengths
#pragma unroll
    for (int rr = 0; rr < R; ++rr) len[rr] = (lengths && b0 + rr < B) ? lengths[b0 + rr] : T;
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        float acc[R][3];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr][0] = acc[rr][1] = acc[rr][2] = 0.0f;
        const float* hcur = hs + (size_t)cur * Bt * H + (size_t)y * R * H;
#pragma unroll 2
        for (int k = 0; k < H; k += 4) {
            float w[4][3];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int g = 0; g < 3; ++g)
                    w[kk][g] = WSMEM ? Ws[(k + kk) * H3 + g * H + i] : __ldg(&Wg[(size_t)(k + kk) * H3 + g * H + i]);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const float4 hv = *reinterpret_cast<const float4*>(hcur + rr * H + k);
                const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int g = 0; g < 3; ++g) acc[rr][g] = fmaf(w[kk][g], h4[kk], acc[rr][g]);
            }
        }
        float* hnxt = hs + (size_t)(cur ^ 1) * Bt * H + (size_t)y * R * H;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const long long b = b0 + rr;
            if (b < B) {
                float px[3];
                if (P.p != nullptr) {
                    const float* pp = P.at(b, t) + dir * H3 + i;
#pragma unroll
                    for (int g = 0; g < 3; ++g) px[g] = pp[g * H];
                } else {
                    const float* xp = x.at(b, t);
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        float s = bi[g];
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (c < I) s = fmaf(wi[g][c], xp[c], s);
                        px[g] = s;
                    }
                }
                const bool active = t < len[rr];   // past the end of a shorter trace: state frozen, output 0, z = 1 saved
                const float r = sigmoid_acc(px[0] + acc[rr][0] + bh[0]);
                const float z = active ? sigmoid_acc(px[1] + acc[rr][1] + bh[1]) : 1.0f;
                const float hn = acc[rr][2] + bh[2];
                const float n = tanhf(px[2] + r * hn);
                const float hold = hcur[rr * H + i];
                const float hnew = active ? (1.0f - z) * n + z * hold : hold;
                hnxt[rr * H + i] = hnew;
                out.at(b, t)[dir * H + i] = active ? hnew : 0.0f;
                if (gates) {
                    float* gp = gates + ((((size_t)dir * B + b) * T + t) * 4) * H + i;
                    gp[0] = r; gp[H] = z; gp[2 * H] = n; gp[3 * H] = hn;
                }
                if (step == T - 1) h_n[((size_t)dir * B + b) * H + i] = hnew;
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

// Backward through time.  Writes, for every (b, t):
//   dGx[., dir*3H + {r,z,n}] = gradient w.r.t. the input-side pre-activations  (-> dW_ih, db_ih, dX)
//   dGh[., dir*3H + {r,z,n}] = gradient w.r.t. the hidden-side pre-activations (-> dW_hh, db_hh)
// and carries dh_{t-1} = z * dh + dGh . W_hh in registers / shared memory.
template <int R, bool WSMEM, int MAXT>
__global__ void __launch_bounds__(MAXT)
gru_bwd_f32_kernel(Seq d_out, const float* __restrict__ d_h_n, const float* __restrict__ gates, Seq out,
                   const float* __restrict__ w_hh, Seq dGx, Seq dGh, const int* __restrict__ lengths, int B, int T, int H) {
    extern __shared__ __align__(16) float smem_f[];
    const int i = threadIdx.x, y = threadIdx.y, S = blockDim.y, Bt = S * R;
    const int dir = blockIdx.y;
    const int H3 = 3 * H;
    float* Ws = smem_f;                              // [3H][H] row-major (original layout)
    float* dgs = smem_f + (WSMEM ? H * H3 : 0);      // [Bt][3H]
    const float* Wg = w_hh + (size_t)dir * H3 * H;
    if (WSMEM) {
        for (int e = threadIdx.y * blockDim.x + threadIdx.x; e < H * H3; e += blockDim.x * blockDim.y) Ws[e] = Wg[e];
    }
    const long long b0 = (long long)blockIdx.x * Bt + y * R;
    float dh[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
        const long long b = b0 + rr;
        dh[rr] = (d_h_n && b < B) ? d_h_n[((size_t)dir * B + b) * H + i] : 0.0f;
    }
    __syncthreads();

    for (int step = T - 1; step >= 0; --step) {
        const int t = dir ? (T - 1 - step) : step;           // forward visited t at position `step`
        const int t_prev = dir ? (t + 1) : (t - 1);          // where h_{prev} of this step was produced
        float direct[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const long long b = b0 + rr;
            float g_r = 0.0f, g_z = 0.0f, g_n = 0.0f, g_hn = 0.0f;
            direct[rr] = 0.0f;
            if (b < B) {
                float d = dh[rr];
                if (d_out.p && (!lengths || t < lengths[b])) d += d_out.at(b, t)[dir * H + i];   // padded outputs carry no gradient
                const float* gp = gates + ((((size_t)dir * B + b) * T + t) * 4) * H + i;
                const float r = gp[0], z = gp[H], n = gp[2 * H], hn = gp[3 * H];
                const float hprev = (step == 0) ? 0.0f : out.at(b, t_prev)[dir * H + i];
                const float dn = d * (1.0f - z);
                const float dz = d * (hprev - n);
                g_n = dn * (1.0f - n * n);
                g_z = dz * z * (1.0f - z);
                g_hn = g_n * r;
                g_r = g_n * hn * r * (1.0f - r);
                direct[rr] = d * z;
                float* gx = dGx.at(b, t) + dir * H3 + i;
                gx[0] = g_r; gx[H] = g_z; gx[2 * H] = g_n;
                float* gh = dGh.at(b, t) + dir * H3 + i;
                gh[0] = g_r; gh[H] = g_z; gh[2 * H] = g_hn;
            }
            float* ds = dgs + (size_t)(y * R + rr) * H3;
            ds[i] = g_r; ds[H + i] = g_z; ds[2 * H + i] = g_hn;
        }
        __syncthreads();
        float acc[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0f;
        const float* dsy = dgs + (size_t)y * R * H3;
#pragma unroll 2
        for (int j = 0; j < H3; j += 4) {
            float w[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) w[jj] = WSMEM ? Ws[(j + jj) * H + i] : __ldg(&Wg[(size_t)(j + jj) * H + i]);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const float4 gv = *reinterpret_cast<const float4*>(dsy + rr * H3 + j);
                acc[rr] = fmaf(w[0], gv.x, acc[rr]);
                acc[rr] = fmaf(w[1], gv.y, acc[rr]);
                acc[rr] = fmaf(w[2], gv.z, acc[rr]);
                acc[rr] = fmaf(w[3], gv.w, acc[rr]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) dh[rr] = direct[rr] + acc[rr];
        __syncthreads();
    }
}

// out[b, t, c] = a[b, t, c] * m[b, t, c] over sequence buffers (dropout mask between layers, both directions)
__global__ void seq_mul_kernel(Seq a, Seq m, Seq o, int B, int T, int C) {
    const long long n = (long long)B * T * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int c = e % C;
        const long long bt = e / C;
        const int t = bt % T;
        const long long b = bt / T;
        o.at(b, t)[c] = a.at(b, t)[c] * (m.p ? m.at(b, t)[c] : 1.0f);
    }
}

Seq mk(const float* p, int64_t ld, int64_t trace_rows, int64_t row0) {
    Seq s;
    s.p = const_cast<float*>(p);
    s.ld = ld; s.trace_rows = trace_rows; s.row0 = row0;
    return s;
}

template <typename K>
int set_smem(K kern, size_t bytes) {
    RS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

}  // namespace

extern "C" int rs_gru_fwd_f32(const float* x, int64_t x_ld, int64_t x_rows, int64_t x_row0, int I, const float* w_ih,
                              const float* b_ih, const float* P, int64_t p_ld, int64_t p_rows, int64_t p_row0,
                              const float* w_hh_t, const float* b_hh, float* out, int64_t o_ld, int64_t o_rows,
                              int64_t o_row0, float* h_n, float* gates, const int* lengths, int B, int T, int H, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(H >= 32 && H % 32 == 0 && H <= 512, "rs_gru_fwd_f32: hidden size %d must be a multiple of 32 in [32,512]", H);
    RS_REQUIRE(P || (x && I >= 1 && I <= 4 && w_ih), "rs_gru_fwd_f32: layers with input size > 4 need the projection P");
    RS_REQUIRE(w_hh_t && b_hh && b_ih && out && h_n && B >= 0 && T >= 0, "rs_gru_fwd_f32: bad arguments");
    if (T == 0) {
        RS_CUDA_OK(cudaMemsetAsync(h_n, 0, sizeof(float) * 2 * (size_t)B * H, stream));
        return 0;
    }
    const bool wsmem = (size_t)H * 3 * H * 4 <= 200 * 1024;
    const int S = (H <= 128) ? 2 : 1024 / H >= 2 ? 2 : 1;
    const bool big = B > 148;                     // few traces: one per thread row, more CTAs, lower step latency
    const int R = big ? 4 : 1;
    const int Bt = S * R;
    const size_t smem = ((wsmem ? (size_t)H * 3 * H : 0) + 2 * (size_t)Bt * H) * sizeof(float);
    dim3 grid((B + Bt - 1) / Bt, 2), block(H, S);
    Seq sx = mk(x, x_ld, x_rows, x_row0), sp = mk(P, p_ld, p_rows, p_row0), so = mk(out, o_ld, o_rows, o_row0);
#define RS_LAUNCH_FWD(RR, WS, MT)                                                                                \
    do {                                                                                                         \
        if (set_smem(gru_fwd_f32_kernel<RR, WS, MT>, smem)) return 2;                                            \
        gru_fwd_f32_kernel<RR, WS, MT><<<grid, block, smem, stream>>>(sx, I, w_ih, b_ih, sp, w_hh_t, b_hh, so,   \
                                                                      h_n, gates, lengths, B, T, H);             \
    } while (0)
#define RS_PICK_FWD(RR)                                                                                          \
    do {                                                                                                         \
        if (wsmem) RS_LAUNCH_FWD(RR, true, 256);                                                                 \
        else if (H * S <= 512) RS_LAUNCH_FWD(RR, false, 512);                                                    \
        else RS_LAUNCH_FWD(RR, false, 1024);                                                                     \
    } while (0)
    if (big) RS_PICK_FWD(4); else RS_PICK_FWD(1);
    rs::count_launch();
#undef RS_PICK_FWD
#undef RS_LAUNCH_FWD
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_gru_bwd_f32(const float* d_out, int64_t do_ld, int64_t do_rows, int64_t do_row0, const float* d_h_n,
                              const float* gates, const float* out, int64_t o_ld, int64_t o_rows, int64_t o_row0,
                              const float* w_hh, float* dGx, float* dGh, int64_t g_ld, int64_t g_rows, int64_t g_row0,
                              const int* lengths, int B, int T, int H, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0 || T == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(H >= 32 && H % 32 == 0 && H <= 512, "rs_gru_bwd_f32: hidden size %d must be a multiple of 32 in [32,512]", H);
    RS_REQUIRE(gates && out && w_hh && dGx && dGh && B >= 0 && T >= 0, "rs_gru_bwd_f32: bad arguments");
    const bool wsmem = (size_t)H * 3 * H * 4 <= 200 * 1024;
    const int S = (H <= 128) ? 2 : 1024 / H >= 2 ? 2 : 1;
    const bool big = B > 148;
    const int R = big ? 4 : 1;
    const int Bt = S * R;
    const size_t smem = ((wsmem ? (size_t)H * 3 * H : 0) + (size_t)Bt * 3 * H) * sizeof(float);
    dim3 grid((B + Bt - 1) / Bt, 2), block(H, S);
    Seq sdo = mk(d_out, do_ld, do_rows, do_row0), so = mk(out, o_ld, o_rows, o_row0);
    Seq sgx = mk(dGx, g_ld, g_rows, g_row0), sgh = mk(dGh, g_ld, g_rows, g_row0);
#define RS_LAUNCH_BWD(RR, WS, MT)                                                                               \
    do {                                                                                                        \
        if (set_smem(gru_bwd_f32_kernel<RR, WS, MT>, smem)) return 2;                                           \
        gru_bwd_f32_kernel<RR, WS, MT><<<grid, block, smem, stream>>>(sdo, d_h_n, gates, so, w_hh, sgx, sgh,    \
                                                                      lengths, B, T, H);                        \
    } while (0)
#define RS_PICK_BWD(RR)                                                                                         \
    do {                                                                                                        \
        if (wsmem) RS_LAUNCH_BWD(RR, true, 256);                                                                \
        else if (H * S <= 512) RS_LAUNCH_BWD(RR, false, 512);                                                   \
        else RS_LAUNCH_BWD(RR, false, 1024);                                                                    \
    } while (0)
    if (big) RS_PICK_BWD(4); else RS_PICK_BWD(1);
    rs::count_launch();
#undef RS_PICK_BWD
#undef RS_LAUNCH_BWD
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_seq_mul_f32(const float* a, int64_t a_ld, int64_t a_rows, int64_t a_row0, const float* m, int64_t m_ld,
                              int64_t m_rows, int64_t m_row0, float* o, int64_t o_ld, int64_t o_rows, int64_t o_row0,
                              int B, int T, int C, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    RS_REQUIRE(a && o, "rs_seq_mul_f32: bad arguments");
    const long long n = (long long)B * T * C;
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    seq_mul_kernel<<<(int)blocks, 256, 0, stream>>>(mk(a, a_ld, a_rows, a_row0), mk(m, m_ld, m_rows, m_row0),
                                                    mk(o, o_ld, o_rows, o_row0), B, T, C);
                                                    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}
