// Persistent tensor-core GRU recurrence for the bf16 mode (H = 128): forward and backward through time
// (SURVEY.md 8(a) rows a3, a4).  One CTA owns 128 traces of ONE direction for all T steps.
//
// Forward, per step:   G[128 x 384] = h_{t-1}[128 x 128] . W_hh^T   on tcgen05 (bf16 operands, fp32 accumulator in TMEM)
//   * W_hh (3H x H bf16, 96 KB) is loaded ONCE into shared memory and stays resident for the whole sequence;
//   * h_{t-1} lives in shared memory as the A operand (no-swizzle core-matrix layout = the tile-major block layout);
//   * 8 epilogue warps (256 threads; thread = trace row x 64 hidden units) pull the accumulator out of TMEM
//     (tcgen05.ld), add the input-side pre-activation (layer 0: fused K=2 projection of the raw (x, y) sample;
//     deeper layers: the time-parallel projection P), apply r/z/n with tanh.approx (sigma(a) = 0.5 tanh(a/2) + 0.5),
//     blend h_t = n + z (h_{t-1} - n), and write h_t back into the A operand in place, to `out` and (training) the
//     gates r, z, n, W_hn h + b_hn for the backward pass -- all with 16-byte accesses that are contiguous across
//     the 32 rows of a warp (512 B per warp instruction) thanks to the tile-major layout;
//   * one mbarrier hands h_t to the MMA-issuing warp, tcgen05.commit hands the accumulator back.
// Backward, per step (reverse time):  dh_{t-1} = z (.) dh_t + dGh_t[128 x 384] . W_hh  on tcgen05, with W_hh^T resident
//   in shared memory, dGh_t written by the epilogue warps as the A operand, the z (.) dh carry kept in fp32 registers.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "rec_common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

using namespace rs;

constexpr int H = 128;
constexpr int CHUNK = 2048;                 // bytes of one 16-byte chunk column over 128 rows
constexpr int W_BYTES = 3 * H * H * 2;      // 96 KB
constexpr int A_FWD_BYTES = (H + 16) * 128 * 2;   // 36 KB  (h tile + one K=16 step for the layer-0 input columns)
constexpr int WX_BYTES = 2 * 384 * 16;       // the two extra 16-byte chunk columns of the layer-0 weight image
constexpr int A_BWD_BYTES = 3 * H * 128 * 2;  // 96 KB (dGh tile)
constexpr int H32_BYTES = H * 128 * 4;      // 64 KB  (fp32 hidden state, forward)
constexpr int NUM_THREADS = 320;            // warp 0: MMA issuer, warp 1: spare, warps 2..9: epilogue
constexpr int EPI_THREADS = 256;

struct FwdParams {
    const float* x; int I;                  // layer 0: raw input (B, T, I <= 4); its projection rides on the MMA (see below)
    const uint8_t* P; long long p_block_bytes;   // deeper layers: tile-major projection, C = 6H, bias folded in
    const uint8_t* Whh;                     // [2][16 or 18][384][8] bf16 (B operand image; 18 chunks with the input rows)
    const float* b_hn;                      // [2][H]
    uint8_t* out; long long out_block_bytes;     // tile-major, C = 2H
    uint8_t* gates;                         // [tiles][T][2][64][128][8] fp16 (private to fwd/bwd) or NULL
    float* h_n;                             // [2][B][H]
    const int* lengths;                     // [B] valid steps per trace (packed-sequence semantics) or NULL
    int B, T;
    int pf_dist;                            // L2 prefetch distance in steps (0 = off)
};

template <bool kVarLen>
__global__ void __launch_bounds__(NUM_THREADS, 1) rec_fwd_bf16_kernel(const FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;                               // [16 chunks][384 rows][16 B]
    uint8_t* a_s = smem + W_BYTES + WX_BYTES;          // [18 chunks][128 rows][16 B]  h_{t-1} | input columns
    uint8_t* h32_s = a_s + A_FWD_BYTES;                // [32 chunks of 4 floats][128 rows][16 B]  fp32 master copy of h
    float* bhn_s = reinterpret_cast<float*>(h32_s + H32_BYTES);   // [H]
    uint64_t* bars = reinterpret_cast<uint64_t*>(bhn_s + H);
    uint64_t* w_full = bars;
    uint64_t* h_ready = bars + 1;       // epilogue -> MMA (8 arrivals, one per epilogue warp)
    uint64_t* acc_full = bars + 2;      // MMA -> epilogue (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, dir = blockIdx.y;
    const int T = p.T;

    if (threadIdx.x == 0) {
        rs::mbar_init(w_full, 1);
        rs::mbar_init(h_ready, 8);
        rs::mbar_init(acc_full, 1);
        rs::fence_mbar_init();
    }
    if (warp == 0) rs::tmem_alloc<512>(tmem_slot);
    for (int i = threadIdx.x; i < H; i += NUM_THREADS) bhn_s[i] = p.b_hn[dir * H + i];
    for (int i = threadIdx.x; i < (A_FWD_BYTES + H32_BYTES) / 16; i += NUM_THREADS) reinterpret_cast<uint4*>(a_s)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const bool fused_x = (p.x != nullptr);
    if (fused_x && threadIdx.x < 128) {     // input columns of the first step
        const long long b = (long long)tile * 128 + threadIdx.x;
        *reinterpret_cast<uint4*>(a_s + 16 * CHUNK + threadIdx.x * 16) =
            pack_x(b < p.B ? p.x + (b * T + (dir ? T - 1 : 0)) * p.I : nullptr, p.I);
    }
    rs::fence_proxy_async();            // h_0 = 0 and the input columns must be visible to the tensor core (async proxy)
    rs::tc_fence_before();
    __syncthreads();
    rs::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // pad rows t' = 0 and T + 1 of this tile: the time-shifted weight-gradient GEMM reads them as h = 0
    for (int i = threadIdx.x; i < 2 * 16 * 128; i += NUM_THREADS) {
        const int r = i & 127, c = (i >> 7) & 15, pad = i >> 11;
        stg16(p.out + ((long long)tile * (T + 2) + (pad ? T + 1 : 0)) * p.out_block_bytes + (long long)(dir * 16 + c) * CHUNK + r * 16,
              make_uint4(0, 0, 0, 0));
    }

    if (warp == 0) {
        // ===================== W loader + MMA issuer =====================
        const uint32_t w_bytes = fused_x ? (W_BYTES + WX_BYTES) : W_BYTES;
        if (lane == 0) {
            rs::mbar_expect_tx(w_full, w_bytes);
            for (int i = 0; i < 6; ++i)
                rs::bulk_load(w_s + i * (w_bytes / 6), p.Whh + (long long)dir * w_bytes + i * (w_bytes / 6), w_bytes / 6, w_full);
        }
        rs::mbar_wait(w_full, 0);
        constexpr uint32_t idesc256 = rs::umma_idesc_bf16(128, 256, 0, 0);
        constexpr uint32_t idesc128 = rs::umma_idesc_bf16(128, 128, 0, 0);
        const uint32_t a_addr = rs::smem_u32(a_s), w_addr = rs::smem_u32(w_s);
        for (int step = 0; step < T; ++step) {
            if (step > 0) {                       // h_0 = 0 (and the first input columns) are already in place for step 0
                rs::mbar_wait(h_ready, (step - 1) & 1);
                rs::tc_fence_after();
            }
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t da = rs::umma_desc_noswz(a_addr + k * 2 * CHUNK, CHUNK, 128);
                    const uint64_t db0 = rs::umma_desc_noswz(w_addr + k * 2 * (384 * 16), 384 * 16, 128);
                    const uint64_t db1 = rs::umma_desc_noswz(w_addr + k * 2 * (384 * 16) + 256 * 16, 384 * 16, 128);
                    rs::tc_mma_bf16(tmem_base, da, db0, idesc256, k != 0);          // r | z  -> columns [0, 256)
                    rs::tc_mma_bf16(tmem_base + 256, da, db1, idesc128, k != 0);    // W_hn h -> columns [256, 384)
                }
                if (fused_x) {
                    // layer 0: one more K=16 step whose A columns are (x_hi, x_lo, x_hi) per input and (1, 1), against
                    // (w_hi, w_hi, w_lo) and (b_hi, b_lo): W_ih x + b accurate to ~2^-16 although the operands are bf16.
                    // r and z simply accumulate it; the n gate keeps it apart (r multiplies only the hidden part).
                    const uint64_t da = rs::umma_desc_noswz(a_addr + 16 * CHUNK, CHUNK, 128);
                    const uint64_t db0 = rs::umma_desc_noswz(w_addr + 16 * (384 * 16), 384 * 16, 128);
                    const uint64_t db1 = rs::umma_desc_noswz(w_addr + 16 * (384 * 16) + 256 * 16, 384 * 16, 128);
                    rs::tc_mma_bf16(tmem_base, da, db0, idesc256, 1u);
                    rs::tc_mma_bf16(tmem_base + 384, da, db1, idesc128, 0u);        // W_in x + b_in -> columns [384, 512)
                }
                rs::tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== L2 prefetcher: pulls the projection block of step+2 towards the SM =====================
        if (lane == 0 && p.P && p.pf_dist > 0) {
            for (int step = 0; step < T; ++step) {
                const int t = dir ? (T - 1 - step) : step;
                const long long blk = (long long)tile * (T + 2) + t + 1;
                rs::l2_prefetch(p.P + blk * p.p_block_bytes + (long long)(dir * 48) * CHUNK, 48 * CHUNK);
                if (step >= p.pf_dist) {               // stay pf_dist steps ahead of the epilogue warps
                    rs::mbar_wait(h_ready, (step - p.pf_dist) & 1);
                }
            }
        }
    } else if (warp >= 2) {
        // ===================== epilogue: gates, blend, stores =====================
        const int ew = warp - 2;
        const int q = warp & 3;                        // TMEM lane quadrant of this warp
        const int half = ew >> 2;                      // which 64 hidden units
        const int row = q * 32 + lane;
        const long long b = (long long)tile * 128 + row;
        const bool live = b < p.B;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        uint8_t* a_row = a_s + row * 16;
        uint8_t* h32_row = h32_s + row * 16;
        const float* xrow = p.x ? p.x + b * T * p.I : nullptr;
        const int len = (kVarLen && live) ? p.lengths[b] : T;      // steps past the end of a shorter trace: h frozen, out = 0

        for (int step = 0; step < T; ++step) {
            const int t = dir ? (T - 1 - step) : step;
            const bool active = !kVarLen || t < len;              // compile-time true without lengths: the fast path is unchanged
            const long long blk = (long long)tile * (T + 2) + t + 1;
            const uint8_t* pblk = p.P ? p.P + blk * p.p_block_bytes + (long long)(dir * 48) * CHUNK + row * 16 : nullptr;
            uint8_t* oblk = p.out + blk * p.out_block_bytes + (long long)(dir * 16) * CHUNK + row * 16;
            uint8_t* gblk = p.gates ? p.gates + (((long long)tile * T + t) * 2 + dir) * (64LL * CHUNK) + row * 16 : nullptr;
            // layer 0: the input columns of the NEXT step (one thread per row writes them before handing h_t over)
            uint4 xnext = make_uint4(0, 0, 0, 0);
            const bool write_x = fused_x && half == 0 && step + 1 < T;
            if (write_x) xnext = pack_x(live ? xrow + (long long)(dir ? t - 1 : t + 1) * p.I : nullptr, p.I);
            // input-side pre-activations: 8 units (one 16-byte chunk per gate) per group, fetched one group ahead;
            // the first group of a step is requested before waiting for the tensor core
            uint4 pv[3];
            auto load_p = [&](int grp) {      // grp = 0..7: units half*64 + grp*8 ..
#pragma unroll
                for (int g = 0; g < 3; ++g) pv[g] = ldg16(pblk + (long long)(g * 16 + half * 8 + grp) * CHUNK);
            };
            if (pblk) load_p(0);
            rs::mbar_wait(acc_full, step & 1);
            rs::tc_fence_after();
            uint32_t ar[8], az[8], an[8], ax[8];
            rs::tmem_ld_32x32b_x8(taddr + half * 64, ar);
            rs::tmem_ld_32x32b_x8(taddr + 128 + half * 64, az);
            rs::tmem_ld_32x32b_x8(taddr + 256 + half * 64, an);
            if (fused_x) rs::tmem_ld_32x32b_x8(taddr + 384 + half * 64, ax);
#pragma unroll
            for (int grp = 0; grp < 8; ++grp) {
                const int u0 = half * 64 + grp * 8;
                float pr[8], pz[8], pn[8], ho[8];
                if (pblk) {
                    unpack8(pv[0], pr); unpack8(pv[1], pz); unpack8(pv[2], pn);
                    if (grp < 7) load_p(grp + 1);
                }
                {   // fp32 h_{t-1}: the blend must not re-round the state every step
                    const float4 v0 = *reinterpret_cast<const float4*>(h32_row + (u0 / 4) * CHUNK);
                    const float4 v1 = *reinterpret_cast<const float4*>(h32_row + (u0 / 4 + 1) * CHUNK);
                    ho[0] = v0.x; ho[1] = v0.y; ho[2] = v0.z; ho[3] = v0.w; ho[4] = v1.x; ho[5] = v1.y; ho[6] = v1.z; ho[7] = v1.w;
                }
                rs::tmem_ld_wait();
                float gr_[8], gz_[8], gn_[8];        // this group's accumulator values; the registers are then reloaded
#pragma unroll
                for (int j = 0; j < 8; ++j) { gr_[j] = __uint_as_float(ar[j]); gz_[j] = __uint_as_float(az[j]); gn_[j] = __uint_as_float(an[j]); }
                if (!pblk) {                        // layer 0: the tensor core already added W_ih x + b to r and z
#pragma unroll
                    for (int j = 0; j < 8; ++j) { pr[j] = 0.0f; pz[j] = 0.0f; pn[j] = __uint_as_float(ax[j]); }
                }
                if (grp < 7) {                      // TMEM loads of the next group fly while this group is computed
                    rs::tmem_ld_32x32b_x8(taddr + u0 + 8, ar);
                    rs::tmem_ld_32x32b_x8(taddr + 128 + u0 + 8, az);
                    rs::tmem_ld_32x32b_x8(taddr + 256 + u0 + 8, an);
                    if (fused_x) rs::tmem_ld_32x32b_x8(taddr + 384 + u0 + 8, ax);
                }
                float hv[8], rv[8], zv[8], nv[8], hnv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // the host folds the 1/2 of sigma(a) = 1/2 tanh(a/2) + 1/2 into the r and z rows of W_hh, W_ih and the biases
                    const float r = fmaf(0.5f, tanh_fast(gr_[j] + pr[j]), 0.5f);
                    const float z = active ? fmaf(0.5f, tanh_fast(gz_[j] + pz[j]), 0.5f) : 1.0f;   // z = 1 saved: BPTT passes dh through
                    const float hn = gn_[j] + bhn_s[u0 + j];
                    const float n = tanh_fast(fmaf(r, hn, pn[j]));
                    hv[j] = active ? fmaf(z, ho[j] - n, n) : ho[j];
                    rv[j] = r; zv[j] = z; nv[j] = n; hnv[j] = hn;
                }
                *reinterpret_cast<float4*>(h32_row + (u0 / 4) * CHUNK) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                *reinterpret_cast<float4*>(h32_row + (u0 / 4 + 1) * CHUNK) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                const uint4 o0 = pack8(hv);
                *reinterpret_cast<uint4*>(a_row + (u0 / 8) * CHUNK) = o0;      // next step's A operand, in place
                stg16(oblk + (long long)(u0 / 8) * CHUNK, active ? o0 : make_uint4(0, 0, 0, 0));
                if (gblk) {
                    stg16(gblk + (long long)(0 * 16 + u0 / 8) * CHUNK, pack8h(rv));
                    stg16(gblk + (long long)(1 * 16 + u0 / 8) * CHUNK, pack8h(zv));
                    stg16(gblk + (long long)(2 * 16 + u0 / 8) * CHUNK, pack8h(nv));
                    stg16(gblk + (long long)(3 * 16 + u0 / 8) * CHUNK, pack8h(hnv));
                }
                if (step == T - 1 && live) {
                    float* hn_out = p.h_n + ((long long)dir * p.B + b) * H + u0;
                    *reinterpret_cast<float4*>(hn_out) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    *reinterpret_cast<float4*>(hn_out + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                }
            }
            if (write_x) *reinterpret_cast<uint4*>(a_row + 16 * CHUNK) = xnext;
            rs::fence_proxy_async();        // h_t written with ordinary stores -> visible to tcgen05.mma
            rs::tc_fence_before();          // our TMEM reads are done before the next MMA overwrites the accumulator
            __syncwarp();
            if (lane == 0) rs::mbar_arrive(h_ready);
        }
    }
    rs::tc_fence_before();
    __syncthreads();
    if (warp == 0) rs::tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------------
struct BwdParams {
    const uint8_t* d_out; long long dout_block_bytes;    // tile-major C = 2H gradient w.r.t. this layer's output, or NULL
    const float* d_h_n;                                  // [2][B][H] or NULL
    const uint8_t* gates;                                // [tiles][T][2][64][128][8]
    const uint8_t* out; long long out_block_bytes;       // this layer's h (tile-major, C = 2H, zero pad rows)
    const uint8_t* WhhT;                                 // [2][48][128][8] bf16: rows = h index, K = (r | z | hn) gate rows
    uint8_t* dG; long long dg_block_bytes;               // tile-major C = 8H: [dir][r | z | n | hn][H]
    const int* lengths;                                  // [B] or NULL
    int B, T;
    int pf_dist;
};

template <bool kVarLen>
__global__ void __launch_bounds__(NUM_THREADS, 1) rec_bwd_bf16_kernel(const BwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_s = smem;                               // [48 chunks][128 rows][16 B]
    uint8_t* a_s = smem + W_BYTES;                     // [48 chunks][128 rows][16 B]  dGh_t
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_s + A_BWD_BYTES);
    uint64_t* w_full = bars;
    uint64_t* a_ready = bars + 1;
    uint64_t* acc_full = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, dir = blockIdx.y;
    const int T = p.T;

    if (threadIdx.x == 0) {
        rs::mbar_init(w_full, 1);
        rs::mbar_init(a_ready, 8);
        rs::mbar_init(acc_full, 1);
        rs::fence_mbar_init();
    }
    if (warp == 0) rs::tmem_alloc<128>(tmem_slot);
    rs::tc_fence_before();
    __syncthreads();
    rs::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    for (int i = threadIdx.x; i < 2 * 64 * 128; i += NUM_THREADS) {      // pad rows of dG
        const int r = i & 127, c = (i >> 7) & 63, pad = i >> 13;
        stg16(p.dG + ((long long)tile * (T + 2) + (pad ? T + 1 : 0)) * p.dg_block_bytes + (long long)(dir * 64 + c) * CHUNK + r * 16,
              make_uint4(0, 0, 0, 0));
    }

    if (warp == 0) {
        if (lane == 0) {
            rs::mbar_expect_tx(w_full, W_BYTES);
            for (int i = 0; i < 6; ++i)
                rs::bulk_load(w_s + i * (W_BYTES / 6), p.WhhT + (long long)dir * W_BYTES + i * (W_BYTES / 6), W_BYTES / 6, w_full);
        }
        rs::mbar_wait(w_full, 0);
        constexpr uint32_t idesc = rs::umma_idesc_bf16(128, 128, 0, 0);
        const uint32_t a_addr = rs::smem_u32(a_s), w_addr = rs::smem_u32(w_s);
        for (int s = 0; s < T - 1; ++s) {              // the result of the last reverse step (dh before t = first) is unused
            rs::mbar_wait(a_ready, s & 1);
            rs::tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < 24; ++k) {
                    const uint64_t da = rs::umma_desc_noswz(a_addr + k * 2 * CHUNK, CHUNK, 128);
                    const uint64_t db = rs::umma_desc_noswz(w_addr + k * 2 * CHUNK, CHUNK, 128);
                    rs::tc_mma_bf16(tmem_base, da, db, idesc, 1u);   // accumulates ONTO the z (.) dh carry stored in TMEM
                }
                rs::tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // L2 prefetcher: saved gates, h_{prev} and d_out of the step two ahead (contiguous tile-major ranges)
        if (lane == 0 && p.pf_dist > 0) {
            for (int s = 0; s < T; ++s) {
                const int fstep = T - 1 - s;
                const int t = dir ? (T - 1 - fstep) : fstep;
                const int t_prev = dir ? t + 1 : t - 1;
                const long long blk = (long long)tile * (T + 2) + t + 1;
                const long long blk_prev = (long long)tile * (T + 2) + t_prev + 1;
                rs::l2_prefetch(p.gates + (((long long)tile * T + t) * 2 + dir) * (64LL * CHUNK), 64 * CHUNK);
                rs::l2_prefetch(p.out + blk_prev * p.out_block_bytes + (long long)(dir * 16) * CHUNK, 16 * CHUNK);
                if (p.d_out) rs::l2_prefetch(p.d_out + blk * p.dout_block_bytes + (long long)(dir * 16) * CHUNK, 16 * CHUNK);
                if (s >= p.pf_dist) rs::mbar_wait(a_ready, (s - p.pf_dist) & 1);
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;
        const int half = ew >> 2;
        const int row = q * 32 + lane;
        const long long b = (long long)tile * 128 + row;
        const bool live = b < p.B;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * 64;
        uint8_t* a_row = a_s + row * 16;
        const int len = (kVarLen && live) ? p.lengths[b] : T;
        // the carry z (.) dh lives in the TMEM accumulator; it starts as d_h_n
#pragma unroll
        for (int sc = 0; sc < 4; ++sc) {
            uint32_t init[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) init[j] = 0u;
            if (p.d_h_n && live) {
                const float* src = p.d_h_n + ((long long)dir * p.B + b) * H + half * 64 + sc * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + j));
                    init[j] = __float_as_uint(v.x); init[j + 1] = __float_as_uint(v.y);
                    init[j + 2] = __float_as_uint(v.z); init[j + 3] = __float_as_uint(v.w);
                }
            }
            rs::tmem_st_32x32b_x16(taddr + sc * 16, init);
        }
        rs::tmem_st_wait();

        // raw 16-byte pieces of one sub-chunk (16 units): r, z, n, hn (fp16), h_prev, d_out (bf16), two chunks each
        uint4 raw[12];
        auto load_raw = [&](int s, int sc) {
            const int fstep = T - 1 - s;
            const int t = dir ? (T - 1 - fstep) : fstep;
            const int t_prev = dir ? t + 1 : t - 1;
            const long long blk = (long long)tile * (T + 2) + t + 1;
            const long long blk_prev = (long long)tile * (T + 2) + t_prev + 1;
            const uint8_t* gblk = p.gates + (((long long)tile * T + t) * 2 + dir) * (64LL * CHUNK) + row * 16;
            const uint8_t* hblk = p.out + blk_prev * p.out_block_bytes + (long long)(dir * 16) * CHUNK + row * 16;
            const int c0 = half * 8 + sc * 2;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                raw[2 * g] = ldg16(gblk + (long long)(g * 16 + c0) * CHUNK);
                raw[2 * g + 1] = ldg16(gblk + (long long)(g * 16 + c0 + 1) * CHUNK);
            }
            raw[8] = ldg16(hblk + (long long)c0 * CHUNK);
            raw[9] = ldg16(hblk + (long long)(c0 + 1) * CHUNK);
            if (p.d_out) {
                const uint8_t* doblk = p.d_out + blk * p.dout_block_bytes + (long long)(dir * 16) * CHUNK + row * 16;
                raw[10] = ldg16(doblk + (long long)c0 * CHUNK);
                raw[11] = ldg16(doblk + (long long)(c0 + 1) * CHUNK);
            } else {
                raw[10] = make_uint4(0, 0, 0, 0);
                raw[11] = make_uint4(0, 0, 0, 0);
            }
        };
        load_raw(0, 0);
        for (int s = 0; s < T; ++s) {                   // s-th reverse step = forward position T-1-s
            const int fstep = T - 1 - s;
            const int t = dir ? (T - 1 - fstep) : fstep;
            const long long blk = (long long)tile * (T + 2) + t + 1;
            const bool active = !kVarLen || t < len;        // padded outputs carry no gradient (their saved z is 1)
            uint8_t* dgblk = p.dG + blk * p.dg_block_bytes + (long long)(dir * 64) * CHUNK + row * 16;
            if (s > 0) {
                rs::mbar_wait(acc_full, (s - 1) & 1);
                rs::tc_fence_after();
            }
#pragma unroll
            for (int sc = 0; sc < 4; ++sc) {
                const int c0 = half * 8 + sc * 2;
                uint32_t acc[16];
                rs::tmem_ld_32x32b_x16(taddr + sc * 16, acc);
                uint4 cur[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) cur[i] = raw[i];
                if (sc < 3) load_raw(s, sc + 1);                    // next sub-chunk of this step
                else if (s + 1 < T) load_raw(s + 1, 0);             // first sub-chunk of the next step (before its MMA wait)
                rs::tmem_ld_wait();
                uint32_t carry[16];
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {                    // two groups of 8 units keep the live set small
                    float r[8], z[8], n[8], hn[8], hp[8], dout[8];
                    unpack8h(cur[0 + hf], r); unpack8h(cur[2 + hf], z); unpack8h(cur[4 + hf], n); unpack8h(cur[6 + hf], hn);
                    unpack8(cur[8 + hf], hp); unpack8(cur[10 + hf], dout);
                    float gr[8], gz[8], gn[8], ghn[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dh = __uint_as_float(acc[hf * 8 + j]) + (active ? dout[j] : 0.0f);
                        const float dn = dh * (1.0f - z[j]);
                        const float dz = dh * (hp[j] - n[j]);
                        gn[j] = dn * (1.0f - n[j] * n[j]);
                        gz[j] = dz * z[j] * (1.0f - z[j]);
                        ghn[j] = gn[j] * r[j];
                        gr[j] = gn[j] * hn[j] * r[j] * (1.0f - r[j]);
                        carry[hf * 8 + j] = __float_as_uint(dh * z[j]);
                    }
                    const uint4 vr = pack8(gr), vz = pack8(gz), vn = pack8(gn), vh = pack8(ghn);
                    // A operand of the dh matvec: K order r | z | hn
                    *reinterpret_cast<uint4*>(a_row + (0 * 16 + c0 + hf) * CHUNK) = vr;
                    *reinterpret_cast<uint4*>(a_row + (1 * 16 + c0 + hf) * CHUNK) = vz;
                    *reinterpret_cast<uint4*>(a_row + (2 * 16 + c0 + hf) * CHUNK) = vh;
                    stg16(dgblk + (long long)(0 * 16 + c0 + hf) * CHUNK, vr);
                    stg16(dgblk + (long long)(1 * 16 + c0 + hf) * CHUNK, vz);
                    stg16(dgblk + (long long)(2 * 16 + c0 + hf) * CHUNK, vn);
                    stg16(dgblk + (long long)(3 * 16 + c0 + hf) * CHUNK, vh);
                }
                rs::tmem_st_32x32b_x16(taddr + sc * 16, carry);     // the next MMA accumulates dGh . W_hh onto it
            }
            rs::tmem_st_wait();
            rs::fence_proxy_async();
            rs::tc_fence_before();
            __syncwarp();
            if (lane == 0) rs::mbar_arrive(a_ready);
        }
    }
    rs::tc_fence_before();
    __syncthreads();
    if (warp == 0) rs::tmem_dealloc<128>(tmem_base);
}

// Tuning knobs: L2 bulk-prefetch distance in steps (0 = off).  Measured at B = 8192 (tools/step_probe.py): the forward
// kernel is fastest one step ahead (6.8 ms vs 7.4 ms off); the backward kernel is fastest WITHOUT prefetch (7.3 ms vs
// 9.7 ms at distance 3: it already runs at ~86 % of the HBM peak, extra prefetches only get evicted and re-read).
// x (B, T, I <= 16) fp32 -> tile-major bf16 with 16 columns (the rest zero), pad rows and pad traces zero:
// the B operand that carries the layer-0 input into the fused weight-gradient pass.
__global__ void pack_x_tm_kernel(const float* __restrict__ x, int B, int T, int I, uint4* __restrict__ out, long long n_rows) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_rows; e += (long long)gridDim.x * blockDim.x) {
        const int row = e & 127;                       // e = ((tile * (T+2) + t') * 128 + row)
        const long long bt = e >> 7;
        const int tp = bt % (T + 2);
        const long long b = (bt / (T + 2)) * 128 + row;
        float c[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = 0.0f;
        if (b < B && tp >= 1 && tp <= T) {
            const float* xp = x + (b * T + (tp - 1)) * I;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < I) c[i] = __ldg(xp + i);
        }
        const long long blk = bt * 256;                // 2 chunks x 128 rows of uint4 per block
        out[blk + row] = pack8(c);
        out[blk + 128 + row] = pack8(c + 8);
    }
}

int pf_dist_env(const char* name, int dflt) {
    const char* e = getenv(name);
    int v = e ? atoi(e) : dflt;
    if (v < 0) v = 0;
    if (v > 64) v = 64;
    return v;
}

}  // namespace

extern "C" int rs_rec_fwd_bf16(const float* x, int I, const void* P, int64_t p_cols, const void* Whh,
                               const float* b_hn, void* out, void* gates, float* h_n, const int* lengths,
                               const void* drop_bits, const float* drop_scale, void* out_drop, int split, int B, int T,
                               void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE((x != nullptr) != (P != nullptr), "rs_rec_fwd_bf16: exactly one of x (layer 0) and P (deeper layers) must be given");
    RS_REQUIRE(!x || (I >= 1 && I <= 2), "rs_rec_fwd_bf16: the MMA-fused input projection takes 1 or 2 input columns");
    RS_REQUIRE(!P || p_cols == 6 * H, "rs_rec_fwd_bf16: P must have 6H = %d columns", 6 * H);
    RS_REQUIRE(Whh && b_hn && out && h_n && B >= 0 && T >= 0, "rs_rec_fwd_bf16: bad arguments");
    if (T == 0) {
        RS_CUDA_OK(cudaMemsetAsync(h_n, 0, sizeof(float) * 2 * (size_t)B * H, stream));
        return 0;
    }
    RS_REQUIRE((drop_bits != nullptr) == (out_drop != nullptr) && (!drop_bits || drop_scale),
               "rs_rec_fwd_bf16: drop_bits, drop_scale and out_drop go together");
    if (const int nt = rs::rec_pair_nt(B, drop_bits != nullptr || split, false))
        return rs::rec_fwd_pair(x, I, P, Whh, b_hn, out, gates, h_n, lengths, drop_bits, drop_scale, out_drop, split, B, T, nt,
                                pf_dist_env("RS_PF_DIST_FWD", 1), stream);
    FwdParams p = {};
    p.x = x; p.I = I;
    p.P = static_cast<const uint8_t*>(P); p.p_block_bytes = 6LL * H * 256;
    p.Whh = static_cast<const uint8_t*>(Whh); p.b_hn = b_hn;
    p.out = static_cast<uint8_t*>(out); p.out_block_bytes = 2LL * H * 256;
    p.gates = static_cast<uint8_t*>(gates); p.h_n = h_n; p.lengths = lengths; p.B = B; p.T = T;
    p.pf_dist = pf_dist_env("RS_PF_DIST_FWD", 1);
    const int smem = W_BYTES + WX_BYTES + A_FWD_BYTES + H32_BYTES + H * 4 + 64;
    if (lengths) {
        RS_CUDA_OK(cudaFuncSetAttribute(rec_fwd_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rec_fwd_bf16_kernel<true><<<dim3((B + 127) / 128, 2), NUM_THREADS, smem, stream>>>(p);
    } else {
        RS_CUDA_OK(cudaFuncSetAttribute(rec_fwd_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rec_fwd_bf16_kernel<false><<<dim3((B + 127) / 128, 2), NUM_THREADS, smem, stream>>>(p);
    }
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_rec_bwd_bf16(const void* d_out, const float* d_h_n, const void* gates, const void* out, const void* WhhT,
                               void* dG, const int* lengths, const void* drop_bits, const float* drop_scale, int split, int B,
                               int T, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    if (B == 0 || T == 0) return 0;        // nothing to do: empty tensors carry null pointers
    RS_REQUIRE(gates && out && WhhT && dG && B >= 0 && T >= 0, "rs_rec_bwd_bf16: bad arguments");
    RS_REQUIRE(!drop_bits || drop_scale, "rs_rec_bwd_bf16: drop_bits needs drop_scale");
    if (const int nt = rs::rec_pair_nt(B, drop_bits != nullptr || split, true))
        // L2 prefetch (RS_PF_DIST_BWD steps ahead) is off by default: at 8192 traces the kernel runs against the HBM roof and
        // prefetched lines are evicted before use (round 1), at 1024 traces the step is bound by its compute / sync chain, not
        // by load latency (7.28 ms per training step without, 7.35 ms with distance 2)
        return rs::rec_bwd_pair(d_out, d_h_n, gates, out, WhhT, dG, lengths, drop_bits, drop_scale, split, B, T, nt,
                                pf_dist_env("RS_PF_DIST_BWD", 0), stream);
    BwdParams p = {};
    p.d_out = static_cast<const uint8_t*>(d_out); p.dout_block_bytes = 2LL * H * 256;
    p.d_h_n = d_h_n; p.gates = static_cast<const uint8_t*>(gates);
    p.out = static_cast<const uint8_t*>(out); p.out_block_bytes = 2LL * H * 256;
    p.WhhT = static_cast<const uint8_t*>(WhhT);
    p.dG = static_cast<uint8_t*>(dG); p.dg_block_bytes = 8LL * H * 256;
    p.lengths = lengths;
    p.B = B; p.T = T;
    p.pf_dist = pf_dist_env("RS_PF_DIST_BWD", 0);
    const int smem = W_BYTES + A_BWD_BYTES + 64;
    if (lengths) {
        RS_CUDA_OK(cudaFuncSetAttribute(rec_bwd_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rec_bwd_bf16_kernel<true><<<dim3((B + 127) / 128, 2), NUM_THREADS, smem, stream>>>(p);
    } else {
        RS_CUDA_OK(cudaFuncSetAttribute(rec_bwd_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        rec_bwd_bf16_kernel<false><<<dim3((B + 127) / 128, 2), NUM_THREADS, smem, stream>>>(p);
    }
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int rs_pack_x_tm(const float* x, int B, int T, int I, void* out, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    RS_REQUIRE(x && out && B >= 0 && T >= 0 && I >= 1 && I <= 16, "rs_pack_x_tm: bad arguments");
    const long long n_rows = (long long)((B + 127) / 128) * (T + 2) * 128;
    if (n_rows == 0) return 0;
    long long blocks = (n_rows + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_x_tm_kernel<<<(int)blocks, 256, 0, stream>>>(x, B, T, I, static_cast<uint4*>(out), n_rows);
    rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}
