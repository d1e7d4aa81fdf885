// Device helpers shared by the bf16 recurrence kernels (rec_pair.cu: H = 128, rec_wide.cu: H = 256; a CTA pair per tile on
// tcgen05 cta_group::2): fast tanh, bf16 / fp16 pack and unpack of 16-byte pieces, the layer-0 input columns.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace rs {

__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
    return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    float2 a = bf2_to_f2(v.x), b = bf2_to_f2(v.y), c = bf2_to_f2(v.z), d = bf2_to_f2(v.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    return make_uint4(f2_to_bf2(f[0], f[1]), f2_to_bf2(f[2], f[3]), f2_to_bf2(f[4], f[5]), f2_to_bf2(f[6], f[7]));
}
// saved gates are private to the two recurrence kernels: fp16 (r, z, n live in [-1, 1], where fp16 is 8x finer than bf16)
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8h(const float* f) {
    return make_uint4(f2_to_h2(f[0], f[1]), f2_to_h2(f[2], f[3]), f2_to_h2(f[4], f[5]), f2_to_h2(f[6], f[7]));
}
__device__ __forceinline__ void unpack8h(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 ldg16(const uint8_t* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(uint8_t* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// Layer-0 input columns of one trace row: per input c the triple (hi, lo, hi) with hi = bf16(x_c), lo = bf16(x_c - hi),
// then (1, 1); at most 2 inputs fit the 8 columns of one 16-byte chunk.
__device__ __forceinline__ uint4 pack_x(const float* xp, int I) {
    float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (xp) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (i < I) {
                const float v = __ldg(xp + i);
                const float hi = __bfloat162float(__float2bfloat16_rn(v));
                c[3 * i] = hi; c[3 * i + 1] = v - hi; c[3 * i + 2] = hi;
            }
        }
    }
    c[6] = 1.0f; c[7] = 1.0f;
    return pack8(c);
}



#ifdef __CUDACC__
// rec_pair.cu: the CTA-pair kernels behind rs_rec_fwd_bf16 / rs_rec_bwd_bf16 (same operands and results as rec_bf16.cu)
int rec_fwd_nt(int B);                      // tiles in flight per pair of the forward kernel (1 or 2)
int rec_fwd_pair(const float* x, int I, const void* P, const void* X, const void* Wih, const void* Whh, const float* b_hn, void* out,
                 void* gates, float* h_n, const int* lengths, const void* drop_bits, const float* drop_scale, void* out_drop, int split,
                 int B, int T, int nt, int pf_dist, cudaStream_t stream);
int rec_bwd_pair(const void* d_out, const float* d_h_n, const void* gates, const void* out, const void* WhhT, const void* Whh,
                 int whh_chunks, const float* b_hn, void* dG, const int* lengths, const void* drop_bits, const float* drop_scale,
                 int split, int B, int T, int pf_dist, cudaStream_t stream);
#endif

}  // namespace rs
