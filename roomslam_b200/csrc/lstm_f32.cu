// fp32 LSTM recurrence, forward and backward-through-time, for the model the upstream repository trains
// (src/benchmark/model.py:16-23 nn.LSTM(d_model, d_model/2, 2 layers, bidirectional)).
// Generic in the hidden size H (multiple of 32, <= 256).  Tolerance 1e-4 against torch.nn.LSTM on the CPU.
//
// Gate equations (torch.nn.LSTM, weight rows ordered i | f | g | o):
//   i = sigma(P_i + Wh_i h)  f = sigma(P_f + Wh_f h)  g = tanh(P_g + Wh_g h)  o = sigma(P_o + Wh_o h)
//   c' = f * c + i * g        h' = o * tanh(c')                  P = W_ih x + b_ih + b_hh   (time-parallel GEMM)
//
// One CTA owns Bt = S*R traces of ONE direction for all T steps.  blockDim = (H, S): thread (u, y) owns hidden unit u
// of traces y*R .. y*R+R-1 and keeps their cell state in registers.  W_hh^T (H x 4H fp32) stays resident in shared
// memory when it fits (H <= 64: 64 KB), otherwise it streams from L2.  h is double-buffered in shared memory:
// one __syncthreads() per step.  Sequence buffers use the (rows, row0, ld) addressing of roomslam_b200.h.
#include "common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

struct SeqF {
    float* p;
    long long ld, trace_rows, row0;
    __device__ __forceinline__ float* at(long long b, long long t) const { return p + ((b * trace_rows + row0 + t) * ld); }
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int R, bool WSMEM>
__global__ void __launch_bounds__(256)
lstm_fwd_f32_kernel(SeqF P, const float* __restrict__ w_hh_t, SeqF out, float* __restrict__ saved, int B, int T, int H) {
    extern __shared__ __align__(16) float lsm[];
    const int u = threadIdx.x, y = threadIdx.y, S = blockDim.y, Bt = S * R;
    const int dir = blockIdx.y, H4 = 4 * H;
    float* Ws = lsm;                                   // [H][4H]
    float* hs = lsm + (WSMEM ? H * H4 : 0);            // [2][Bt][H]
    const float* Wg = w_hh_t + (size_t)dir * H * H4;
    const int tid = y * blockDim.x + u, nthr = blockDim.x * blockDim.y;
    if (WSMEM) for (int e = tid; e < H * H4; e += nthr) Ws[e] = Wg[e];
    for (int e = tid; e < 2 * Bt * H; e += nthr) hs[e] = 0.0f;
    __syncthreads();

    const long long b0 = (long long)blockIdx.x * Bt + y * R;
    float c[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) c[rr] = 0.0f;
    int cur = 0;
    for (int step = 0; step < T; ++step) {
        const int t = dir ? (T - 1 - step) : step;
        float acc[R][4];
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[rr][g] = 0.0f;
        // issue this step's projection loads before the matvec so their latency hides behind it
        float px[R][4];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const long long b = b0 + rr;
            if (b < B) {
                const float* pp = P.at(b, t) + dir * H4 + u;
#pragma unroll
                for (int g = 0; g < 4; ++g) px[rr][g] = __ldg(pp + g * H);
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) px[rr][g] = 0.0f;
            }
        }
        const float* hcur = hs + (size_t)cur * Bt * H + (size_t)y * R * H;
#pragma unroll 2
        for (int k = 0; k < H; k += 4) {
            float w[4][4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    w[kk][g] = WSMEM ? Ws[(k + kk) * H4 + g * H + u] : __ldg(&Wg[(size_t)(k + kk) * H4 + g * H + u]);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const float4 hv = *reinterpret_cast<const float4*>(hcur + rr * H + k);
                const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int g = 0; g < 4; ++g) acc[rr][g] = fmaf(w[kk][g], h4[kk], acc[rr][g]);
            }
        }
        float* hnxt = hs + (size_t)(cur ^ 1) * Bt * H + (size_t)y * R * H;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const long long b = b0 + rr;
            if (b < B) {
                const float gi = sigm(px[rr][0] + acc[rr][0]);
                const float gf = sigm(px[rr][1] + acc[rr][1]);
                const float gg = tanhf(px[rr][2] + acc[rr][2]);
                const float go = sigm(px[rr][3] + acc[rr][3]);
                c[rr] = gf * c[rr] + gi * gg;
                const float hnew = go * tanhf(c[rr]);
                hnxt[rr * H + u] = hnew;
                out.at(b, t)[dir * H + u] = hnew;
                if (saved) {
                    float* sp = saved + ((((size_t)dir * B + b) * T + t) * 5) * H + u;
                    sp[0] = gi; sp[H] = gf; sp[2 * H] = gg; sp[3 * H] = go; sp[4 * H] = c[rr];
                }
            }
        }
        __syncthreads();
        cur ^= 1;
    }
}

// Backward through time: dG[b, t, dir*4H + {i,f,g,o}] = gradient w.r.t. the gate pre-activations (the same tensor
// feeds dW_ih, dW_hh, the biases and dX); the carries dh_{t-1} = dG . W_hh and dc_{t-1} = dc * f stay on chip.
template <int R, bool WSMEM>
__global__ void __launch_bounds__(256)
lstm_bwd_f32_kernel(SeqF d_out, const float* __restrict__ saved, const float* __restrict__ w_hh, SeqF dG, int B, int T, int H) {
    extern __shared__ __align__(16) float lsm[];
    const int u = threadIdx.x, y = threadIdx.y, S = blockDim.y, Bt = S * R;
    const int dir = blockIdx.y, H4 = 4 * H;
    float* Ws = lsm;                                   // [4H][H] (the original row-major layout)
    float* dgs = lsm + (WSMEM ? H * H4 : 0);           // [Bt][4H]
    const float* Wg = w_hh + (size_t)dir * H4 * H;
    const int tid = y * blockDim.x + u, nthr = blockDim.x * blockDim.y;
    if (WSMEM) for (int e = tid; e < H * H4; e += nthr) Ws[e] = Wg[e];
    const long long b0 = (long long)blockIdx.x * Bt + y * R;
    float dh[R], dc[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) { dh[rr] = 0.0f; dc[rr] = 0.0f; }
    __syncthreads();

    for (int step = T - 1; step >= 0; --step) {
        const int t = dir ? (T - 1 - step) : step;
        const int t_prev = dir ? (t + 1) : (t - 1);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const long long b = b0 + rr;
            float g_i = 0.0f, g_f = 0.0f, g_g = 0.0f, g_o = 0.0f;
            if (b < B) {
                const float d = dh[rr] + d_out.at(b, t)[dir * H + u];
                const float* sp = saved + ((((size_t)dir * B + b) * T + t) * 5) * H + u;
                const float gi = sp[0], gf = sp[H], gg = sp[2 * H], go = sp[3 * H], cc = sp[4 * H];
                const float cprev = (step == 0) ? 0.0f : saved[((((size_t)dir * B + b) * T + t_prev) * 5 + 4) * H + u];
                const float tc = tanhf(cc);
                const float dcell = dc[rr] + d * go * (1.0f - tc * tc);
                g_o = d * tc * go * (1.0f - go);
                g_i = dcell * gg * gi * (1.0f - gi);
                g_f = dcell * cprev * gf * (1.0f - gf);
                g_g = dcell * gi * (1.0f - gg * gg);
                dc[rr] = dcell * gf;
                float* gp = dG.at(b, t) + dir * H4 + u;
                gp[0] = g_i; gp[H] = g_f; gp[2 * H] = g_g; gp[3 * H] = g_o;
            }
            float* ds = dgs + (size_t)(y * R + rr) * H4;
            ds[u] = g_i; ds[H + u] = g_f; ds[2 * H + u] = g_g; ds[3 * H + u] = g_o;
        }
        __syncthreads();
        float acc[R];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) acc[rr] = 0.0f;
        const float* dsy = dgs + (size_t)y * R * H4;
#pragma unroll 2
        for (int j = 0; j < H4; j += 4) {
            float w[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) w[jj] = WSMEM ? Ws[(j + jj) * H + u] : __ldg(&Wg[(size_t)(j + jj) * H + u]);
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                const float4 gv = *reinterpret_cast<const float4*>(dsy + rr * H4 + j);
                acc[rr] = fmaf(w[0], gv.x, acc[rr]);
                acc[rr] = fmaf(w[1], gv.y, acc[rr]);
                acc[rr] = fmaf(w[2], gv.z, acc[rr]);
                acc[rr] = fmaf(w[3], gv.w, acc[rr]);
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr) dh[rr] = acc[rr];
        __syncthreads();
    }
}

SeqF mkf(const float* p, int64_t ld, int64_t trace_rows, int64_t row0) {
    SeqF s;
    s.p = const_cast<float*>(p);
    s.ld = ld; s.trace_rows = trace_rows; s.row0 = row0;
    return s;
}

struct Shape { int S, R; bool wsmem; };
Shape pick(int B, int H) {
    Shape s;
    s.wsmem = (size_t)H * 4 * H * 4 <= 96 * 1024;          // H <= 64 keeps two CTAs per SM with W resident
    s.S = (H <= 64) ? 4 : (H <= 128 ? 2 : 1);
    s.R = (B > 148 * 2) ? 4 : 1;                          // few traces: one per thread row, more CTAs, lower step latency
    return s;
}

}  // namespace

extern "C" int rs_lstm_fwd_f32(const float* P, int64_t p_ld, int64_t p_rows, int64_t p_row0, const float* w_hh_t,
                               float* out, int64_t o_ld, int64_t o_rows, int64_t o_row0, float* saved, int B, int T, int H,
                               void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    RS_REQUIRE(H >= 32 && H % 32 == 0 && H <= 256, "rs_lstm_fwd_f32: hidden size %d must be a multiple of 32 in [32,256]", H);
    RS_REQUIRE(P && w_hh_t && out && B >= 0 && T >= 0, "rs_lstm_fwd_f32: bad arguments");
    if (B == 0 || T == 0) return 0;
    const Shape s = pick(B, H);
    const int Bt = s.S * s.R;
    const size_t smem = ((s.wsmem ? (size_t)H * 4 * H : 0) + 2 * (size_t)Bt * H) * sizeof(float);
    dim3 grid((B + Bt - 1) / Bt, 2), block(H, s.S);
    SeqF sp = mkf(P, p_ld, p_rows, p_row0), so = mkf(out, o_ld, o_rows, o_row0);
#define RS_LSTM_FWD(RR, WS)                                                                                            \
    do {                                                                                                               \
        RS_CUDA_OK(cudaFuncSetAttribute(lstm_fwd_f32_kernel<RR, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        lstm_fwd_f32_kernel<RR, WS><<<grid, block, smem, stream>>>(sp, w_hh_t, so, saved, B, T, H);                    \
    } while (0)
    if (s.R == 4) { if (s.wsmem) RS_LSTM_FWD(4, true); else RS_LSTM_FWD(4, false); }
    else          { if (s.wsmem) RS_LSTM_FWD(1, true); else RS_LSTM_FWD(1, false); }
#undef RS_LSTM_FWD
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_lstm_bwd_f32(const float* d_out, int64_t do_ld, int64_t do_rows, int64_t do_row0, const float* saved,
                               const float* w_hh, float* dG, int64_t g_ld, int64_t g_rows, int64_t g_row0, int B, int T,
                               int H, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    RS_REQUIRE(H >= 32 && H % 32 == 0 && H <= 256, "rs_lstm_bwd_f32: hidden size %d must be a multiple of 32 in [32,256]", H);
    RS_REQUIRE(d_out && saved && w_hh && dG && B >= 0 && T >= 0, "rs_lstm_bwd_f32: bad arguments");
    if (B == 0 || T == 0) return 0;
    const Shape s = pick(B, H);
    const int Bt = s.S * s.R;
    const size_t smem = ((s.wsmem ? (size_t)H * 4 * H : 0) + (size_t)Bt * 4 * H) * sizeof(float);
    dim3 grid((B + Bt - 1) / Bt, 2), block(H, s.S);
    SeqF sdo = mkf(d_out, do_ld, do_rows, do_row0), sg = mkf(dG, g_ld, g_rows, g_row0);
#define RS_LSTM_BWD(RR, WS)                                                                                            \
    do {                                                                                                               \
        RS_CUDA_OK(cudaFuncSetAttribute(lstm_bwd_f32_kernel<RR, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        lstm_bwd_f32_kernel<RR, WS><<<grid, block, smem, stream>>>(sdo, saved, w_hh, sg, B, T, H);                     \
    } while (0)
    if (s.R == 4) { if (s.wsmem) RS_LSTM_BWD(4, true); else RS_LSTM_BWD(4, false); }
    else          { if (s.wsmem) RS_LSTM_BWD(1, true); else RS_LSTM_BWD(1, false); }
#undef RS_LSTM_BWD
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}
