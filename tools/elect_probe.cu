// scratch: what ptxas makes of a single-thread tcgen05.mma sequence under `if (lane == 0)` versus `if (elect_one())`.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -cubin -o /tmp/elect_probe.cubin tools/elect_probe.cu
//   cuobjdump -sass /tmp/elect_probe.cubin | grep -c "BRA.U.ANY"      # k_lane0: one ELECT loop per MMA; k_elect: none
// ("single-thread instruction issue belongs under elect.sync")
#include <cstdint>
#include "../roomslam_b200/csrc/common.cuh"
using namespace rs;

__global__ void k_lane0(uint32_t tmem, uint32_t a, uint32_t b, uint64_t* bar) {
    const int lane = threadIdx.x & 31;
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            tc_mma_bf16_pair(tmem, umma_desc_noswz(a + k * 2048, 1024, 128), umma_desc_noswz(b + k * 2048, 1024, 128), idesc, 1u);
        tc_commit_pair(bar);
    }
}

__global__ void k_elect(uint32_t tmem, uint32_t a, uint32_t b, uint64_t* bar) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
    if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            tc_mma_bf16_pair(tmem, umma_desc_noswz(a + k * 2048, 1024, 128), umma_desc_noswz(b + k * 2048, 1024, 128), idesc, 1u);
        tc_commit_pair(bar);
    }
}
