/* CPU oracle (plain C restatement) of the occupancy-heatmap / stationary-time binning rules.
 * TEST INFRASTRUCTURE ONLY: linked/loaded only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  PARITY UNPINNED (see oracle/baseline_ref.py header: upstream
 * README.md:15,34,163-164 names the component, ships no code for it).  Rules D10-D12 of SURVEY.md 8(a):
 *   fx = (x - x_min)/res (fp32, RN), binned iff 0 <= fx < Gx && 0 <= fy < Gy, cell = floor(fy)*Gx + floor(fx)
 *   stationary(t>=1) iff fl(fl(dx*dx) + fl(dy*dy)) < thr2, counted at the cell of p_t when p_t is binned.
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fPIC -shared  (contraction off: no FMA may be formed).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int cell_of(float x, float y, float x_min, float y_min, float res, int gx, int gy)
{
    volatile float sx = x - x_min, sy = y - y_min; /* volatile: keep the sub and the div separate roundings */
    float fx = sx / res, fy = sy / res;
    if (!(fx >= 0.0f && fx < (float)gx && fy >= 0.0f && fy < (float)gy)) return -1;
    return (int)floorf(fy) * gx + (int)floorf(fx);
}

/* points: n_traces x seq_len x 2 floats.  occ/stat: gy*gx int32 (overwritten).  returns dropped points.
 * n_threads <= 0: use all OpenMP threads. */
int64_t rs_ref_heatmap_bin(const float *points, int64_t n_traces, int64_t seq_len, float x_min, float y_min,
                           float res, int gx, int gy, float thr2, int32_t *occ, int32_t *stat, int n_threads)
{
    const int64_t cells = (int64_t)gx * gy;
    int64_t dropped = 0;
    memset(occ, 0, sizeof(int32_t) * cells);
    memset(stat, 0, sizeof(int32_t) * cells);
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel reduction(+ : dropped)
    {
        int32_t *po = (int32_t *)calloc(cells, sizeof(int32_t));
        int32_t *ps = (int32_t *)calloc(cells, sizeof(int32_t));
#pragma omp for schedule(static)
        for (int64_t b = 0; b < n_traces; ++b) {
            const float *p = points + b * seq_len * 2;
            float px = 0.f, py = 0.f;
            for (int64_t t = 0; t < seq_len; ++t) {
                float x = p[2 * t], y = p[2 * t + 1];
                int c = cell_of(x, y, x_min, y_min, res, gx, gy);
                if (c < 0) {
                    dropped++;
                } else {
                    po[c]++;
                    if (t > 0) {
                        volatile float dx = x - px, dy = y - py;
                        volatile float a = dx * dx, bq = dy * dy;
                        float d2 = a + bq;
                        if (d2 < thr2) ps[c]++;
                    }
                }
                px = x;
                py = y;
            }
        }
#pragma omp critical
        {
            for (int64_t i = 0; i < cells; ++i) {
                occ[i] += po[i];
                stat[i] += ps[i];
            }
        }
        free(po);
        free(ps);
    }
    return dropped;
}
