// Occupancy-heatmap + stationary-time binning (upstream spec README.md:15,163-164).
//
// One streaming pass over float2 points (8 B/point, HBM-bound).  Rules D10/D11 are reproduced bit for bit:
//   fx = fl(fl(x - x_min) / res); binned iff 0 <= fx < Gx (same for y); cell = floor(fy)*Gx + floor(fx)
//   stationary(t>=1) iff fl(fl(dx*dx) + fl(dy*dy)) < thr2, counted at the cell of p_t.
// This translation unit is compiled with -fmad=false and uses explicit *_rn intrinsics: no FMA is formed.
//
// Fast kernel (heatmap_tma_kernel): persistent CTAs, one per SM.
//   * each CTA owns a privatised histogram in shared memory: one 32-bit word per cell packing
//     occupancy (low 16 bits) and stationary (high 16 bits) so that ONE shared atomic serves both grids;
//     a field that crosses 0x8000 is folded into the global int32 grid by the thread that saw the crossing
//     (increments are capped at kChunkPts per atomic, so a field can never carry into its neighbour);
//   * each warp walks tiles of 32 traces (lane <-> trace, so the lanes of a warp hit unrelated cells and
//     the previous point of a trace lives in a register); the tile is streamed through a 2-stage
//     TMA pipeline (cp.async.bulk.tensor, box 32 traces x kChunkPts points, hardware swizzle so that
//     the per-lane row walk is bank-conflict free with 128-bit shared loads);
//   * consecutive points of a trace that fall in the same cell are merged in registers (run-length)
//     before the atomic: a paused person costs one atomic per chunk instead of one per sample.
//   * floor(fl(s/res)) is obtained without a division: q = s * fl(1/res) is within 3 ulp of the exact
//     quotient, so whenever q is farther than eps from every integer both floors agree; the rare
//     near-integer point (and only it) re-does the IEEE division.
// Generic kernel (heatmap_generic_kernel): thread per point with global atomics; used for odd seq_len,
// unaligned pointers or grids too large for shared memory.
#include "common.cuh"
#include "../../include/roomslam_b200.h"

namespace {

struct BinParams {
    float x_min, y_min, res, inv_res, near_eps, thr2;
    int gx, gy;
};

constexpr int kTileRows = 32;
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: (q + kMagic) - kMagic == rint(q) for |q| < 2^22
constexpr int kMagicBits = 0x4B400000;

// Exact (slow) cell: the IEEE division of rule D10.  Out of line: it runs for about one point in 10^4.
__device__ __noinline__ int cell_exact(float x, float y, BinParams P) {
    const float fx = __fdiv_rn(__fsub_rn(x, P.x_min), P.res), fy = __fdiv_rn(__fsub_rn(y, P.y_min), P.res);
    const bool ok = (fx >= 0.0f) && (fx < static_cast<float>(P.gx)) && (fy >= 0.0f) && (fy < static_cast<float>(P.gy));
    return ok ? static_cast<int>(floorf(fy)) * P.gx + static_cast<int>(floorf(fx)) : -1;
}

// Fast cell: q = s * fl(1/res); t = RD(q + 1.5*2^23) holds floor(q) in its mantissa; d = q - floor(q).
// `near` is set when d is within eps of 0 or 1, i.e. when floor(q) might differ from floor(fl(s/res)).
// Returns the cell index; `ok` says whether the point is inside the grid (NaN / Inf / huge values fail it).
__device__ __forceinline__ int cell_fast(float x, float y, const BinParams& P, bool& ok, bool& near) {
    const float qx = __fmul_rn(__fsub_rn(x, P.x_min), P.inv_res), qy = __fmul_rn(__fsub_rn(y, P.y_min), P.inv_res);
    const float tx = __fadd_rd(qx, kMagic), ty = __fadd_rd(qy, kMagic);
    const float dx = __fsub_rn(qx, __fsub_rn(tx, kMagic)), dy = __fsub_rn(qy, __fsub_rn(ty, kMagic));
    const int ix = __float_as_int(tx) - kMagicBits, iy = __float_as_int(ty) - kMagicBits;
    near = (fabsf(__fsub_rn(dx, 0.5f)) >= P.near_eps) || (fabsf(__fsub_rn(dy, 0.5f)) >= P.near_eps);  // near_eps = 0.5 - eps
    ok = (static_cast<unsigned>(ix) < static_cast<unsigned>(P.gx)) && (static_cast<unsigned>(iy) < static_cast<unsigned>(P.gy));
    return iy * P.gx + ix;
}

__device__ __forceinline__ int cell_of(float x, float y, const BinParams& P) {
    bool ok, near;
    const int c = cell_fast(x, y, P, ok, near);
    return near ? cell_exact(x, y, P) : (ok ? c : -1);
}

__device__ __forceinline__ bool is_stationary(float x, float y, float px, float py, float thr2) {
    const float dx = __fsub_rn(x, px), dy = __fsub_rn(y, py);
    const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    return d2 < thr2;
}

// Rare: a 16-bit field of the packed shared counter reached 0x8000; move 0x8000 of it to the global grid.
// which: bit 0 = occupancy field crossed, bit 16 = stationary field crossed.
__device__ __noinline__ void hist_fold(uint32_t* hist, int cell, uint32_t which, int* g_occ, int* g_stat,
                                       unsigned int* folds) {
    if (which & 0x1u) {
        atomicSub(&hist[cell], 0x00008000u);
        atomicAdd(&g_occ[cell], 0x8000);
        atomicAdd(folds, 1u);
    }
    if (which & 0x10000u) {
        atomicSub(&hist[cell], 0x80000000u);
        atomicAdd(&g_stat[cell], 0x8000);
    }
}

// One shared atomic serves both grids (occupancy in the low half, stationary in the high half).  Returns a word
// with bit 0 / bit 16 set when this very increment made the low / high field reach 0x8000.
__device__ __forceinline__ uint32_t hist_add(uint32_t hist_u32, int cell, uint32_t inc) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(hist_u32 + 4u * static_cast<uint32_t>(cell)), "r"(inc) : "memory");
    return (((old + inc) & ~old) & 0x80008000u) >> 15;
}

template <int kWarps, int kChunkPts, int kStages>
__global__ void __launch_bounds__(kWarps * 32, 1)
heatmap_tma_kernel(const __grid_constant__ CUtensorMap tmap, const BinParams P, const long long n_traces,
                   const int seq_len, const int n_tiles, const int hist_bytes, int* __restrict__ g_occ,
                   int* __restrict__ g_stat, unsigned long long* __restrict__ g_dropped) {
    constexpr int kRowBytes = kChunkPts * 8;           // 64 (SWIZZLE_64B) or 128 (SWIZZLE_128B)
    constexpr int kBufBytes = kTileRows * kRowBytes;
    constexpr int kVecs = kRowBytes / 16;              // float4 per row
    constexpr int kSwzMask = kVecs - 1;
    constexpr int kSwzShift = (kRowBytes == 64) ? 1 : 0;  // 64B: chunk ^= (row>>1)&3 ; 128B: chunk ^= row&7

    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem);
    uint8_t* stage = smem + hist_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + kWarps * kStages * kBufBytes);
    __shared__ unsigned int s_folds;
    __shared__ unsigned long long s_points, s_binned;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cells = P.gx * P.gy;
    for (int i = threadIdx.x; i < hist_bytes / 4; i += blockDim.x) hist[i] = 0u;
    if (threadIdx.x == 0) {
        s_folds = 0u;
        s_points = 0ull;
        s_binned = 0ull;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) rs::mbar_init(&bars[kStages * warp + s], 1);
        rs::fence_mbar_init();
    }
    if (threadIdx.x == 0) rs::prefetch_tmap(&tmap);
    __syncthreads();

    uint8_t* buf0 = stage + warp * kStages * kBufBytes;
    uint64_t* bar = &bars[kStages * warp];
    const uint64_t policy = rs::policy_evict_first();
    const int n_chunks = (seq_len + kChunkPts - 1) / kChunkPts;
    const int swz = ((lane >> kSwzShift) & kSwzMask);
    uint32_t phase_bits = 0;     // bit s = parity to wait for on stage s
    uint32_t hist_u32;           // shared-window address of the histogram, kept opaque so it stays in a register
    asm volatile("mov.u32 %0, %1;" : "=r"(hist_u32) : "r"(rs::smem_u32(hist)));
    unsigned long long my_points = 0;

    for (int tile = blockIdx.x * kWarps + warp; tile < n_tiles; tile += gridDim.x * kWarps) {
        const int row0 = tile * kTileRows;
        const long long rows_left = n_traces - row0;
        const bool active = lane < rows_left;
        my_points += static_cast<unsigned long long>(rows_left < kTileRows ? rows_left : kTileRows) * seq_len;
        if (rs::elect_one()) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                if (s < n_chunks) {
                    rs::mbar_expect_tx(&bar[s], kBufBytes);
                    rs::tma_load_2d_hint(buf0 + s * kBufBytes, &tmap, &bar[s], s * kChunkPts * 2, row0, policy);
                }
            }
        }
        float px = 0.0f, py = 0.0f;
        for (int c = 0; c < n_chunks; ++c) {
            const int s = (kStages == 1) ? 0 : (c % kStages);
            rs::mbar_wait(&bar[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
            const uint8_t* row = buf0 + s * kBufBytes + lane * kRowBytes;
            float4 v[kVecs];
#pragma unroll
            for (int j = 0; j < kVecs; ++j) v[j] = *reinterpret_cast<const float4*>(row + ((j ^ swz) << 4));

            // 1) cells and stationary flags of the kChunkPts points (independent work: ILP)
            int cell[kChunkPts];
            bool ok[kChunkPts];
            uint32_t inc[kChunkPts];
            bool any_near = false;
            const int t0 = c * kChunkPts;
#pragma unroll
            for (int j = 0; j < kVecs; ++j) {
                const float x0 = v[j].x, y0 = v[j].y, x1 = v[j].z, y1 = v[j].w;
                bool near0, near1;
                cell[2 * j] = cell_fast(x0, y0, P, ok[2 * j], near0);
                cell[2 * j + 1] = cell_fast(x1, y1, P, ok[2 * j + 1], near1);
                any_near |= near0 | near1;
                const bool st0 = is_stationary(x0, y0, px, py, P.thr2) && (t0 + 2 * j > 0);
                const bool st1 = is_stationary(x1, y1, x0, y0, P.thr2);
                inc[2 * j] = st0 ? 0x10001u : 1u;
                inc[2 * j + 1] = st1 ? 0x10001u : 1u;
                px = x1;
                py = y1;
            }
            // Every float of the row has been consumed by the arithmetic above (the asm pins that order), so the
            // shared loads have completed in every lane after the __syncwarp: only now may TMA refill the stage.
#pragma unroll
            for (int i = 0; i < kChunkPts; ++i) asm volatile("" ::"r"(cell[i]), "r"(inc[i]) : "memory");
            __syncwarp();
            if (c + kStages < n_chunks && rs::elect_one()) {       // elect.sync: ptxas issues the TMA without a per-lane loop
                rs::mbar_expect_tx(&bar[s], kBufBytes);
                rs::tma_load_2d_hint(buf0 + s * kBufBytes, &tmap, &bar[s], (c + kStages) * kChunkPts * 2, row0, policy);
            }
            if (any_near) {  // rare: redo every point of this lane's chunk with the exact division
#pragma unroll
                for (int j = 0; j < kVecs; ++j) {
                    cell[2 * j] = cell_exact(v[j].x, v[j].y, P);
                    cell[2 * j + 1] = cell_exact(v[j].z, v[j].w, P);
                    ok[2 * j] = cell[2 * j] >= 0;
                    ok[2 * j + 1] = cell[2 * j + 1] >= 0;
                }
            }
            const int n_valid = active ? (seq_len - t0) : 0;
            if (n_valid < kChunkPts) {   // ragged tail / inactive lane (TMA zero-fills what lies outside the tensor)
#pragma unroll
                for (int i = 0; i < kChunkPts; ++i) ok[i] = ok[i] && (i < n_valid);
            }
            // 2) one shared atomic per point, branch-free: points outside the grid go to a scratch word past the
            //    last cell.  (The lanes of a warp are different traces: no same-address serialisation; ATOMS
            //    wavefronts are far below the LSU limit, issue slots are what is scarce.)
            uint32_t crossed = 0;   // bit (7-i) / (23-i): point i made the occupancy / stationary field reach 0x8000
#pragma unroll
            for (int i = 0; i < kChunkPts; ++i) {
                cell[i] = ok[i] ? cell[i] : cells;
                crossed = (crossed << 1) + hist_add(hist_u32, cell[i], inc[i]);
            }
            if (crossed) {          // rare
#pragma unroll
                for (int i = 0; i < kChunkPts; ++i) {
                    const uint32_t which = (crossed >> (kChunkPts - 1 - i)) & 0x10001u;
                    if (which && cell[i] != cells) hist_fold(hist, cell[i], which, g_occ, g_stat, &s_folds);
                }
            }
        }
    }

    if (lane == 0 && my_points) atomicAdd(&s_points, my_points);
    __syncthreads();
    // flush the private histogram; dropped points = points walked - points binned (no per-point counting)
    unsigned int binned = 0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) {
        const uint32_t w = hist[i];
        binned += (w & 0xFFFFu);
        if (w & 0xFFFFu) atomicAdd(&g_occ[i], static_cast<int>(w & 0xFFFFu));
        if (w >> 16) atomicAdd(&g_stat[i], static_cast<int>(w >> 16));
    }
    binned = __reduce_add_sync(0xffffffffu, binned);
    if (lane == 0 && binned) atomicAdd(&s_binned, static_cast<unsigned long long>(binned));
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long total_binned = s_binned + 0x8000ull * s_folds;
        if (s_points != total_binned) atomicAdd(g_dropped, s_points - total_binned);
    }
}

__global__ void __launch_bounds__(256)
heatmap_generic_kernel(const float* __restrict__ pts, const long long n_points, const int seq_len, const BinParams P,
                       int* __restrict__ g_occ, int* __restrict__ g_stat, unsigned long long* __restrict__ g_dropped) {
    unsigned int dropped = 0;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_points; i += stride) {
        const float x = pts[2 * i], y = pts[2 * i + 1];
        const int cl = cell_of(x, y, P);
        if (cl < 0) {
            ++dropped;
        } else {
            atomicAdd(&g_occ[cl], 1);
            if ((i % seq_len) != 0 && is_stationary(x, y, pts[2 * i - 2], pts[2 * i - 1], P.thr2))
                atomicAdd(&g_stat[cl], 1);
        }
    }
    dropped = __reduce_add_sync(0xffffffffu, dropped);
    if ((threadIdx.x & 31) == 0 && dropped) atomicAdd(g_dropped, static_cast<unsigned long long>(dropped));
}

int g_num_sms = 0;

template <int kWarps, int kChunkPts, int kStages>
int launch_tma(const float* points, long long n_traces, int seq_len, const BinParams& P, int* occ, int* stat,
               unsigned long long* dropped, cudaStream_t stream) {
    constexpr int kRowBytes = kChunkPts * 8;
    const int cells = P.gx * P.gy;
    const int hist_bytes = (((cells + 1) * 4 + 1023) / 1024) * 1024;  // +1: scratch word for out-of-grid points
    const int smem = hist_bytes + kWarps * kStages * kTileRows * kRowBytes + kWarps * kStages * 8;
    CUtensorMap tmap;
    if (rs::make_tmap_2d(&tmap, points, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 2ull * seq_len, n_traces,
                         8ull * seq_len, kChunkPts * 2, kTileRows,
                         kRowBytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B))
        return 2;
    auto kern = heatmap_tma_kernel<kWarps, kChunkPts, kStages>;
    RS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const long long n_tiles = (n_traces + kTileRows - 1) / kTileRows;
    RS_REQUIRE(n_tiles < (1ll << 26), "rs_heatmap_bin: too many traces for one launch (%lld)", n_traces);
    int grid = static_cast<int>((n_tiles + kWarps - 1) / kWarps);
    if (grid > g_num_sms) grid = g_num_sms;
    kern<<<grid, kWarps * 32, smem, stream>>>(tmap, P, n_traces, seq_len, static_cast<int>(n_tiles), hist_bytes, occ,
                                              stat, dropped);
                                              rs::count_launch();
    RS_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" int rs_heatmap_bin_variant(const float* points, int64_t n_traces, int64_t seq_len, float x_min, float y_min,
                                      float res, int gx, int gy, float thr2, int32_t* occ, int32_t* stat,
                                      unsigned long long* n_dropped, int accumulate, int variant, void* stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rs::check_device_sm100()) return 3;
    RS_REQUIRE(n_traces >= 0 && seq_len >= 0 && seq_len < (1 << 30), "rs_heatmap_bin: bad shape (%lld, %lld)",
               (long long)n_traces, (long long)seq_len);
    RS_REQUIRE((points || n_traces * seq_len == 0) && occ && stat && n_dropped, "rs_heatmap_bin: null pointer argument");
    RS_REQUIRE(gx > 0 && gy > 0 && (long long)gx * gy < (1ll << 30) && gx < (1 << 21) && gy < (1 << 21),
               "rs_heatmap_bin: bad grid %d x %d", gx, gy);
    RS_REQUIRE(res > 0.0f, "rs_heatmap_bin: resolution must be positive");
    if (g_num_sms == 0) {
        int dev = 0;
        RS_CUDA_OK(cudaGetDevice(&dev));
        RS_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long cells = (long long)gx * gy;
    if (!accumulate) {
        RS_CUDA_OK(cudaMemsetAsync(occ, 0, cells * sizeof(int32_t), stream));
        RS_CUDA_OK(cudaMemsetAsync(stat, 0, cells * sizeof(int32_t), stream));
        RS_CUDA_OK(cudaMemsetAsync(n_dropped, 0, sizeof(unsigned long long), stream));
    }
    if (n_traces == 0 || seq_len == 0) return 0;

    BinParams P;
    P.x_min = x_min;
    P.y_min = y_min;
    P.res = res;
    P.inv_res = 1.0f / res;
    const int gmax = gx > gy ? gx : gy;
    P.near_eps = 0.5f - (2.0f * gmax + 4.0f) * 2.384185791015625e-07f;  // 0.5 - eps, eps = (2G+4) * 2^-22 >= 3 ulp of q
    P.thr2 = thr2;
    P.gx = gx;
    P.gy = gy;

    const bool tma_ok = (seq_len % 2 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15) == 0) && cells <= 40960 &&
                        gmax <= 4096;
    if (variant == 0) variant = tma_ok ? 5 : 3;
    RS_REQUIRE(variant == 3 || tma_ok, "rs_heatmap_bin: variant %d needs even seq_len, 16-byte aligned points and "
               "a grid of at most 40960 cells", variant);
    switch (variant) {
        case 1:
            return launch_tma<16, 8, 2>(points, n_traces, (int)seq_len, P, occ, stat, n_dropped, stream);
        case 2:
            return launch_tma<8, 16, 2>(points, n_traces, (int)seq_len, P, occ, stat, n_dropped, stream);
        case 4:
            return launch_tma<32, 8, 1>(points, n_traces, (int)seq_len, P, occ, stat, n_dropped, stream);
        case 5:
            return launch_tma<24, 8, 1>(points, n_traces, (int)seq_len, P, occ, stat, n_dropped, stream);
        case 3: {
            const long long n_points = n_traces * seq_len;
            long long blocks = (n_points + 255) / 256;
            if (blocks > 148ll * 16) blocks = 148ll * 16;
            heatmap_generic_kernel<<<(int)blocks, 256, 0, stream>>>(points, n_points, (int)seq_len, P, occ, stat,
                                                                    n_dropped);
                                                                    rs::count_launch();
            RS_CUDA_OK(cudaGetLastError());
            return 0;
        }
        default:
            RS_REQUIRE(false, "rs_heatmap_bin: unknown variant %d", variant);
    }
    return 0;
}

extern "C" int rs_heatmap_bin(const float* points, int64_t n_traces, int64_t seq_len, float x_min, float y_min,
                              float res, int gx, int gy, float thr2, int32_t* occ, int32_t* stat,
                              unsigned long long* n_dropped, int accumulate, void* stream) {
    return rs_heatmap_bin_variant(points, n_traces, seq_len, x_min, y_min, res, gx, gy, thr2, occ, stat, n_dropped,
                                  accumulate, 0, stream);
}

// Host-buffer entry: what a caller holding numpy/CPU tensors uses.  Streams the traces through two device
// staging buffers so that the host->device copy of chunk k+1 overlaps the binning of chunk k, then copies
// the two grids and the dropped count back.  Blocking.
extern "C" int rs_heatmap_bin_host(const float* host_points, int64_t n_traces, int64_t seq_len, float x_min,
                                   float y_min, float res, int gx, int gy, float thr2, int32_t* host_occ,
                                   int32_t* host_stat, unsigned long long* host_dropped) {
    if (rs::check_device_sm100()) return 3;
    RS_REQUIRE(n_traces >= 0 && seq_len >= 0, "rs_heatmap_bin_host: bad shape");
    RS_REQUIRE((host_points || n_traces * seq_len == 0) && host_occ && host_stat && host_dropped,
               "rs_heatmap_bin_host: null pointer argument");
    static cudaStream_t streams[2] = {nullptr, nullptr};
    static float* stage[2] = {nullptr, nullptr};
    static int32_t* d_grids = nullptr;
    static long long d_grid_cells = 0;
    static unsigned long long* d_dropped = nullptr;
    constexpr long long kStageBytes = 256ll << 20;
    if (!streams[0]) {
        for (int i = 0; i < 2; ++i) {
            RS_CUDA_OK(cudaStreamCreateWithFlags(&streams[i], cudaStreamNonBlocking));
            RS_CUDA_OK(cudaMalloc(&stage[i], kStageBytes));
        }
        RS_CUDA_OK(cudaMalloc(&d_dropped, sizeof(unsigned long long)));
    }
    const long long cells = (long long)gx * gy;
    if (cells > d_grid_cells) {
        if (d_grids) RS_CUDA_OK(cudaFree(d_grids));
        RS_CUDA_OK(cudaMalloc(&d_grids, 2 * cells * sizeof(int32_t)));
        d_grid_cells = cells;
    }
    int32_t* d_occ = d_grids;
    int32_t* d_stat = d_grids + cells;
    RS_CUDA_OK(cudaMemsetAsync(d_grids, 0, 2 * cells * sizeof(int32_t), streams[0]));
    RS_CUDA_OK(cudaMemsetAsync(d_dropped, 0, sizeof(unsigned long long), streams[0]));
    RS_CUDA_OK(cudaStreamSynchronize(streams[0]));
    const long long trace_bytes = seq_len * 8;
    long long per_chunk = trace_bytes > 0 ? kStageBytes / trace_bytes : 0;
    per_chunk = (per_chunk / 32) * 32;
    RS_REQUIRE(n_traces == 0 || seq_len == 0 || per_chunk > 0, "rs_heatmap_bin_host: seq_len too large for staging");
    int k = 0;
    for (long long t0 = 0; t0 < n_traces && seq_len > 0; t0 += per_chunk, ++k) {
        const long long n = (n_traces - t0 < per_chunk) ? (n_traces - t0) : per_chunk;
        cudaStream_t s = streams[k & 1];
        RS_CUDA_OK(cudaMemcpyAsync(stage[k & 1], host_points + t0 * seq_len * 2, n * trace_bytes,
                                   cudaMemcpyHostToDevice, s));
        int rc = rs_heatmap_bin_variant(stage[k & 1], n, seq_len, x_min, y_min, res, gx, gy, thr2, d_occ, d_stat,
                                        d_dropped, 1, 0, s);
        if (rc) return rc;
    }
    RS_CUDA_OK(cudaStreamSynchronize(streams[0]));
    RS_CUDA_OK(cudaStreamSynchronize(streams[1]));
    RS_CUDA_OK(cudaMemcpy(host_occ, d_occ, cells * sizeof(int32_t), cudaMemcpyDeviceToHost));
    RS_CUDA_OK(cudaMemcpy(host_stat, d_stat, cells * sizeof(int32_t), cudaMemcpyDeviceToHost));
    RS_CUDA_OK(cudaMemcpy(host_dropped, d_dropped, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return 0;
}
