// Host-side planners of the dataset ops: the random draws of the reference's Dataset.__getitem__
// (/root/reference/src/pretrain/cola_training.py:56-80, src/pretrain/mae_training.py:57-109) evaluated over a whole
// batch of items from ONE stream of uniform numbers, in the reference's order.  The uniforms come from Python's
// Mersenne Twister (datasets.py transplants the generator state into numpy's MT19937 and back), so crop starts,
// masked rows and gains are bit-identical with the per-item Python code; this file only replaces the per-frame
// Python loop of random_mask (src/util.py:35-46) by native code.  No device work here.
#include <math.h>
#include <stdint.h>

#include "api_common.h"

extern "C" {

// Python's int(x): truncation toward zero
static inline int64_t py_int(double x) { return (int64_t)x; }

int64_t hmfe_cola_draws(const double* u, int64_t n_u, const int64_t* rows, int64_t n_items, int max_len, int windowing,
                        int augment, double rate_start, double rate_seq, uint8_t* mask, int64_t* mask_off,
                        int64_t* win_start, int64_t* start1, int64_t* start2, float* gain1, float* gain2) {
    if (!u || !rows || n_items < 0 || max_len <= 0 || !mask_off || !win_start || !start1 || !start2 || !gain1 || !gain2 ||
        (augment && !mask)) {
        hmfe::set_error("hmfe_cola_draws: bad argument");
        return HMFE_ERR_INVALID;
    }
    int64_t k = 0, moff = 0;
    for (int64_t it = 0; it < n_items; ++it) {
        int64_t T = rows[it];
        win_start[it] = 0;
        if (windowing && T > 3 * (int64_t)max_len) {  // random_crop(x, 3 * max_len): mae_training.py:66-67
            if (k + 1 > n_u) return -100;
            win_start[it] = py_int(u[k++] * (double)(T - 3 * (int64_t)max_len));
            T = 3 * (int64_t)max_len;
        }
        mask_off[it] = moff;
        if (augment) {  // random_mask: r1 < rate_start or (prev and r2 < rate_seq), r2 drawn only when needed
            if (k + 2 * T > n_u) return -100;
            bool prev = false;
            for (int64_t r = 0; r < T; ++r) {
                bool m = u[k++] < rate_start;
                if (!m && prev) m = u[k++] < rate_seq;
                mask[moff + r] = m ? 1 : 0;
                prev = m;
            }
            moff += T;
        }
        if (k + 4 > n_u) return -100;
        start1[it] = py_int(u[k++] * (double)(T - max_len));  // random_crop: int(random.random() * (T - crop))
        start2[it] = py_int(u[k++] * (double)(T - max_len));
        gain1[it] = gain2[it] = 1.0f;
        if (augment) {  // random_multiply: 0.9 + random.random() / 5.0, applied in float32
            gain1[it] = (float)(0.9 + u[k++] / 5.0);
            gain2[it] = (float)(0.9 + u[k++] / 5.0);
        }
    }
    return k;  // uniforms consumed
}

}  // extern "C"
