// Probe for tcgen05 cta_group::2 (a CTA pair on one TPC): checks the TMEM accumulator layout of the M = 128 ("2x2") and
// M = 256 ("4x1") pair MMAs, the multicast commit and the remote mbarrier arrive that rec_pair.cu relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I include -o tools/pair_probe tools/pair_probe.cu
// Prints "pair_probe M=128 ok" / "M=256 ok" or the first mismatches.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../roomslam_b200/csrc/common.cuh"

namespace rs {
void set_error(const char*, ...) {}
}

constexpr int K = 32;          // two K = 16 steps
constexpr int N = 256;

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(rs::smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(rs::smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

__host__ __device__ inline float a_val(int row, int k) { return (float)((row * 3 + k) % 7 - 3); }
__host__ __device__ inline float b_val(int n, int k) { return (float)((n * 5 + k * 2) % 5 - 2); }

// M_TOTAL = 128: 64 rows per CTA; 256: 128 rows per CTA.  out[cta][lane 128][col 256] raw TMEM dump.
template <int M_TOTAL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(160, 1) probe_kernel(float* out) {
    constexpr int ROWS = M_TOTAL / 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_s = smem;                                   // [K/8 chunks][ROWS][16 B]
    uint8_t* b_s = smem + (K / 8) * ROWS * 16;             // [K/8 chunks][N/2 rows][16 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + (K / 8) * (N / 2) * 16);
    uint64_t* a_ready = bars;          // leader only: both CTAs' operands are in place (2 arrivals)
    uint64_t* acc_full = bars + 1;     // each CTA: multicast commit
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const uint32_t rank = cluster_rank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        rs::mbar_init(a_ready, 2);
        rs::mbar_init(acc_full, 1);
        rs::fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(rs::smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // operands: A rows [rank*ROWS, +ROWS), B rows n in [rank*N/2, +N/2)
    for (int i = threadIdx.x; i < (K / 8) * ROWS; i += blockDim.x) {
        const int c = i / ROWS, r = i % ROWS;
        __nv_bfloat16 v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(a_val(rank * ROWS + r, c * 8 + j));
        *reinterpret_cast<uint4*>(a_s + (c * ROWS + r) * 16) = *reinterpret_cast<uint4*>(v);
    }
    for (int i = threadIdx.x; i < (K / 8) * (N / 2); i += blockDim.x) {
        const int c = i / (N / 2), r = i % (N / 2);
        __nv_bfloat16 v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16(b_val(rank * (N / 2) + r, c * 8 + j));
        *reinterpret_cast<uint4*>(b_s + (c * (N / 2) + r) * 16) = *reinterpret_cast<uint4*>(v);
    }
    rs::fence_proxy_async();
    rs::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // barriers initialised and TMEM allocated in both CTAs
    rs::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (threadIdx.x == 0) {             // operands ready: tell the leader (remote arrive from the peer, local from the leader)
        mbar_arrive_remote(mapa(rs::smem_u32(a_ready), 0));
    }
    if (warp == 4 && rank == 0) {
        while (!mbar_try_wait_cluster(a_ready, 0)) {
        }
        rs::tc_fence_after();
        if (lane == 0) {
            constexpr uint32_t idesc = rs::umma_idesc_bf16(M_TOTAL, N, 0, 0);
            for (int k = 0; k < K / 16; ++k) {
                const uint64_t da = rs::umma_desc_noswz(rs::smem_u32(a_s) + k * 2 * ROWS * 16, ROWS * 16, 128);
                const uint64_t db = rs::umma_desc_noswz(rs::smem_u32(b_s) + k * 2 * (N / 2) * 16, (N / 2) * 16, 128);
                tc_mma_bf16_pair(tmem_base, da, db, idesc, k != 0);
            }
            tc_commit_pair(acc_full);
        }
        __syncwarp();
    }
    if (warp < 4) {
        rs::mbar_wait(acc_full, 0);
        rs::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            rs::tmem_ld_32x32b_x32(taddr + c0, v);
            rs::tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[((size_t)rank * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    rs::tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // nobody frees TMEM / exits while the pair may still touch the peer
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

static float c_ref(int row, int n) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += a_val(row, k) * b_val(n, k);
    return s;
}

template <int M_TOTAL>
int run() {
    constexpr int ROWS = M_TOTAL / 2;
    float* d;
    cudaMalloc(&d, sizeof(float) * 2 * 128 * N);
    cudaMemset(d, 0xff, sizeof(float) * 2 * 128 * N);
    const int smem = (K / 8) * ROWS * 16 + (K / 8) * (N / 2) * 16 + 64;
    cudaFuncSetAttribute(probe_kernel<M_TOTAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<M_TOTAL><<<2, 160, smem>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair_probe M=%d: CUDA error %s\n", M_TOTAL, cudaGetErrorString(e)); return 1; }
    std::vector<float> h(2 * 128 * N);
    cudaMemcpy(h.data(), d, sizeof(float) * h.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int cta = 0; cta < 2; ++cta)
        for (int lane = 0; lane < 128; ++lane)
            for (int col = 0; col < (M_TOTAL == 128 ? N / 2 : N); ++col) {
                int row, n;
                if (M_TOTAL == 128) { row = cta * 64 + lane % 64; n = (lane / 64) * (N / 2) + col; }   // "2x2" atom
                else { row = cta * 128 + lane; n = col; }                                              // "4x1" atom
                const float got = h[((size_t)cta * 128 + lane) * N + col], want = c_ref(row, n);
                if (got != want && bad++ < 8) printf("  M=%d cta %d lane %d col %d: got %g want %g\n", M_TOTAL, cta, lane, col, got, want);
            }
    printf("pair_probe M=%d %s (%d mismatches)\n", M_TOTAL, bad ? "FAILED" : "ok", bad);
    cudaFree(d);
    return bad != 0;
}

int main() {
    int rc = run<128>();
    rc |= run<256>();
    return rc;
}
