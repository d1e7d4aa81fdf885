/* roomslam_b200 C ABI: the drop-in boundary for Room-SLAM's data-parallel hot path on B200 (sm_100a).
 *
 * Upstream (Ex10si0n/room-slam) is pure Python and has no FFI; the seam it exposes is
 *   model(traces) -> dict of tensors            src/benchmark/model.py:150-153, built by build_model :406-443
 *   criterion(outputs, targets) -> dict          src/benchmark/train.py:109-135
 * and the README-specified GRU model / heatmap baseline (README.md:110-125, :15, :163-164) that this library
 * implements.  A Python caller binds these symbols with ctypes (roomslam_b200/_lib.py; INTEGRATION.md shows the
 * stub a maintainer would add upstream).  Conventions for every entry point:
 *   - plain pointers and sizes only; pointers are DEVICE pointers unless the name ends in _host;
 *   - tensors are dense, row-major, with the shapes given per function; nothing is copied or re-laid out silently;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous on it
 *     unless documented as blocking;
 *   - return 0 on success; non-zero on error (1 bad argument, 2 CUDA error, 3 no sm_100 device) with the text
 *     available from rs_last_error().  There is no CPU fallback anywhere.
 */
#ifndef ROOMSLAM_B200_H
#define ROOMSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ------------------------------------------------------------------------------------------- */
const char* rs_last_error(void);       /* thread-local message of the last failing call */
int rs_abi_version(void);              /* bumped on any signature change */
int rs_device_ok(void);                /* 1 when the current CUDA device is sm_100 (B200), else 0 */

/* ---- occupancy heatmap + stationary time (replaces README.md:15,163-164 `src/models/baseline.py`) ------- */
/* points: float32 [n_traces, seq_len, 2].  occ, stat: int32 [gy, gx].  n_dropped: one uint64.
 * Binning rules: SURVEY.md 8(a) D10/D11 (bit-exact with oracle/baseline_ref.py).  accumulate != 0 adds to
 * the existing contents of occ/stat/n_dropped instead of zeroing them first (used when sharding by trace). */
int rs_heatmap_bin(const float* points, int64_t n_traces, int64_t seq_len, float x_min, float y_min, float res,
                   int gx, int gy, float thr2, int32_t* occ, int32_t* stat, unsigned long long* n_dropped,
                   int accumulate, void* stream);
/* Same, forcing a kernel variant (0 auto, 1 TMA 16 warps x 8-point chunks, 2 TMA 8 warps x 16-point chunks,
 * 3 generic global-atomic kernel).  For tests and tuning. */
int rs_heatmap_bin_variant(const float* points, int64_t n_traces, int64_t seq_len, float x_min, float y_min,
                           float res, int gx, int gy, float thr2, int32_t* occ, int32_t* stat,
                           unsigned long long* n_dropped, int accumulate, int variant, void* stream);
/* HOST pointers in and out; blocking; overlaps the host->device copy with binning. */
int rs_heatmap_bin_host(const float* host_points, int64_t n_traces, int64_t seq_len, float x_min, float y_min,
                        float res, int gx, int gy, float thr2, int32_t* host_occ, int32_t* host_stat,
                        unsigned long long* host_dropped);

#ifdef __cplusplus
}
#endif
#endif /* ROOMSLAM_B200_H */
