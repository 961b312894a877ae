// Host-side constant tables (float64 math, rounded once to float32).
//
// The formulas are the published definitions the reference's libraries use:
//   * periodic Hann window + Slaney mel scale / Slaney-normalised triangles
//     (librosa 0.10.1 filters.mel, called from /root/reference/src/util.py:484-492)
//   * Kaldi HTK-mel triangles and symmetric Hann window
//     (torchaudio.compliance.kaldi.get_mel_banks / _feature_window_function,
//      called from /root/reference/src/util.py:845-856)
// plus the layouts the kernels consume (per-lane twiddle planes, banded mel slots).
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace hmfe {

static const double kPi = 3.14159265358979323846;

// ------------------------------------------------------------------ windows
// 0.5 * periodic Hann(n).  The 0.5 folds the 1/4 of the two-real-frames-in-one-complex-FFT
// separation |X|^2 = |Z[k] +- conj Z[N-k]|^2 / 4 into the window (exact: power of two).
inline std::vector<float> half_hann_periodic(int n) {
    std::vector<float> w(n);
    for (int k = 0; k < n; ++k) {
        const double h = 0.5 - 0.5 * cos(2.0 * kPi * k / n);
        w[k] = 0.5f * (float)h;
    }
    return w;
}
// torch.hann_window(n, periodic=False) scaled by 0.5 (same folding as above)
inline std::vector<float> half_hann_symmetric(int n) {
    std::vector<float> w(n);
    for (int k = 0; k < n; ++k) {
        const double h = 0.5 - 0.5 * cos(2.0 * kPi * k / (n - 1));
        w[k] = 0.5f * (float)h;
    }
    return w;
}

// ------------------------------------------------------------------ Slaney mel (librosa)
inline double hz_to_mel_slaney(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
inline double mel_to_hz_slaney(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0;
    const double min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// dense [n_mels][n_fft/2+1], float32, Slaney norm
inline std::vector<float> mel_filterbank_slaney(int sr, int n_fft, int n_mels, double fmin, double fmax) {
    const int n_bins = n_fft / 2 + 1;
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel_slaney(fmin), m_hi = hz_to_mel_slaney(fmax);
    for (int i = 0; i < n_mels + 2; ++i) {
        // numpy.linspace: start + i*step, last point exact
        const double step = (m_hi - m_lo) / (n_mels + 1);
        const double m = (i == n_mels + 1) ? m_hi : m_lo + i * step;
        mel_f[i] = mel_to_hz_slaney(m);
    }
    std::vector<float> w((size_t)n_mels * n_bins, 0.0f);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int k = 0; k < n_bins; ++k) {
            const double f = (double)k * sr / n_fft;
            const double lower = -(mel_f[i] - f) / fd0;
            const double upper = (mel_f[i + 2] - f) / fd1;
            const float tri = (float)std::max(0.0, std::min(lower, upper));  // float32 store
            w[(size_t)i * n_bins + k] = (float)((double)tri * enorm);         // float32 *= float64
        }
    }
    return w;
}

// ------------------------------------------------------------------ Kaldi HTK mel (torchaudio)
inline double mel_htk(double f) { return 1127.0 * log(1.0 + f / 700.0); }

// dense [n_mels][padded/2 + 1]; last column (Nyquist) is zero as in kaldi.fbank
inline std::vector<float> mel_banks_kaldi(int n_mels, int padded, double sr, double low, double high) {
    const int n_fft_bins = padded / 2, n_bins = n_fft_bins + 1;
    const double nyq = 0.5 * sr;
    if (high <= 0.0) high += nyq;
    const double bin_w = sr / padded;
    const double mlo = mel_htk(low), mhi = mel_htk(high);
    const double delta = (mhi - mlo) / (n_mels + 1);
    std::vector<float> w((size_t)n_mels * n_bins, 0.0f);
    for (int i = 0; i < n_mels; ++i) {
        const double left = mlo + i * delta, center = mlo + (i + 1.0) * delta, right = mlo + (i + 2.0) * delta;
        for (int k = 0; k < n_fft_bins; ++k) {
            const double m = mel_htk(bin_w * k);
            const double up = (m - left) / (center - left), down = (right - m) / (right - center);
            w[(size_t)i * n_bins + k] = (float)std::max(0.0, std::min(up, down));
        }
    }
    return w;
}

// ------------------------------------------------------------------ banded mel slots
// The kernels evaluate the sparse mel projection with one mel row per lane per "slot".
// Rows are sorted by support length so the 32 rows of a slot have similar trip counts;
// every (slot, lane) gets a window [start, start+trip) inside [0, n_bins) that covers the
// row's support, zero weights elsewhere.  Weight plane layout: w[(wbase[s] + i)*32 + lane].
struct BandedMel {
    int n_mels = 0, n_bins = 0, n_slots = 0;
    std::vector<int> trip, wbase;   // per slot
    std::vector<int> start, row;    // per (slot, lane); row = -1 for padding lanes
    std::vector<float> w;
    int total_trip = 0;
};

// `group`: lanes whose shared-memory reads are served in one phase (8 for 16-byte tile elements,
// 16 for 8-byte ones).  Lane l gets a window start congruent to l modulo `group`, which makes the
// per-iteration tile reads bank-conflict free; rows are placed on the lane that needs the
// smallest downward shift of their window.  Trip counts are rounded up to a multiple of 8 (no
// remainder loop) and every window must end at or before `cap` tile rows.
// Minimum-cost assignment of n rows to n columns (Hungarian algorithm with potentials, O(n^3)); cost[i][j] >= 0.
// Returns col_of_row.
inline std::vector<int> assign_min_cost(const std::vector<std::vector<int>>& cost) {
    const int n = (int)cost.size();
    const int INF = 1 << 29;
    std::vector<int> u(n + 1, 0), v(n + 1, 0), p(n + 1, 0), way(n + 1, 0);
    for (int i = 1; i <= n; ++i) {
        p[0] = i;
        int j0 = 0;
        std::vector<int> minv(n + 1, INF);
        std::vector<char> used(n + 1, 0);
        do {
            used[j0] = 1;
            const int i0 = p[j0];
            int delta = INF, j1 = 0;
            for (int j = 1; j <= n; ++j)
                if (!used[j]) {
                    const int cur = cost[i0 - 1][j - 1] - u[i0] - v[j];
                    if (cur < minv[j]) {
                        minv[j] = cur;
                        way[j] = j0;
                    }
                    if (minv[j] < delta) {
                        delta = minv[j];
                        j1 = j;
                    }
                }
            for (int j = 0; j <= n; ++j)
                if (used[j]) {
                    u[p[j]] += delta;
                    v[j] -= delta;
                } else {
                    minv[j] -= delta;
                }
            j0 = j1;
        } while (p[j0] != 0);
        do {
            const int j1 = way[j0];
            p[j0] = p[j1];
            j0 = j1;
        } while (j0);
    }
    std::vector<int> col(n, -1);
    for (int j = 1; j <= n; ++j)
        if (p[j] > 0) col[p[j] - 1] = j - 1;
    return col;
}

// `prefer` (a multiple of `group`, 0 = off): after the trip counts are fixed by the modulo-`group` placement, the rows
// of every slot are re-assigned to lanes so that as many windows as possible start congruent to their lane modulo
// `prefer` WITHOUT lengthening the slot's trip count (minimum-cost assignment: 0 for a modulo-`prefer` fit, 1 for a
// modulo-`group` fit).  For 8-byte tile elements (a shared-memory wavefront serves 16 lanes) group = 8 keeps the
// windows short and prefer = 16 removes most of the two-way conflicts between lanes l and l + 8 that it would cost.
inline BandedMel build_banded(const std::vector<float>& dense, int n_mels, int n_bins, int group = 8, int cap = 576,
                              int prefer = 0) {
    BandedMel b;
    b.n_mels = n_mels;
    b.n_bins = n_bins;
    std::vector<int> first(n_mels, 0), len(n_mels, 0), order(n_mels);
    for (int m = 0; m < n_mels; ++m) {
        int lo = -1, hi = -1;
        for (int k = 0; k < n_bins; ++k)
            if (dense[(size_t)m * n_bins + k] != 0.0f) {
                if (lo < 0) lo = k;
                hi = k;
            }
        first[m] = lo < 0 ? 0 : lo;
        len[m] = lo < 0 ? 0 : hi - lo + 1;
        order[m] = m;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return len[a] > len[c]; });
    b.n_slots = (n_mels + 31) / 32;
    b.trip.assign(b.n_slots, 8);
    b.wbase.assign(b.n_slots, 0);
    b.start.assign((size_t)b.n_slots * 32, 0);
    b.row.assign((size_t)b.n_slots * 32, -1);
    std::vector<int> shift((size_t)b.n_slots * 32, 0);
    for (int s = 0; s < b.n_slots; ++s) {
        bool used[32] = {false};
        int need = 1;
        for (int j = 0; j < 32; ++j) {
            const int idx = s * 32 + j;
            if (idx >= n_mels) break;
            const int m = order[idx];
            int best = -1, best_shift = 1 << 30;
            for (int l = 0; l < 32; ++l) {
                if (used[l]) continue;
                const int sh = ((first[m] - l) % group + group) % group;  // start = first - sh == l (mod group)
                if (first[m] - sh < 0) continue;
                if (sh < best_shift) {
                    best_shift = sh;
                    best = l;
                }
            }
            if (best < 0) {  // only possible for rows starting below `group`: take any free lane unshifted
                for (int l = 0; l < 32 && best < 0; ++l)
                    if (!used[l]) best = l;
                best_shift = 0;
            }
            used[best] = true;
            b.row[(size_t)s * 32 + best] = m;
            b.start[(size_t)s * 32 + best] = first[m] - best_shift;
            need = std::max(need, len[m] + best_shift);
        }
        b.trip[s] = (need + 7) / 8 * 8;
        b.wbase[s] = b.total_trip;
        b.total_trip += b.trip[s];
        if (prefer > group && prefer % group == 0) {
            const int T = b.trip[s], BIG = 1000;
            std::vector<int> rows;
            for (int j = 0; j < 32 && s * 32 + j < n_mels; ++j) rows.push_back(order[s * 32 + j]);
            std::vector<std::vector<int>> cost(32, std::vector<int>(32, 0));  // rows beyond rows.size() are dummies
            auto shift_of = [&](int m, int l, int mod) { return ((first[m] - l) % mod + mod) % mod; };
            auto fits = [&](int m, int sh) { return first[m] - sh >= 0 && len[m] + sh <= T; };
            for (size_t j = 0; j < rows.size(); ++j)
                for (int l = 0; l < 32; ++l) {
                    const int m = rows[j];
                    cost[j][l] = fits(m, shift_of(m, l, prefer)) ? 0 : (fits(m, shift_of(m, l, group)) ? 1 : BIG);
                }
            const std::vector<int> lane_of = assign_min_cost(cost);
            int total = 0;
            for (size_t j = 0; j < rows.size(); ++j) total += cost[j][lane_of[j]];
            if (total < BIG) {  // otherwise keep the greedy placement (rows that start below `group`)
                for (int l = 0; l < 32; ++l) {
                    b.row[(size_t)s * 32 + l] = -1;
                    b.start[(size_t)s * 32 + l] = 0;
                }
                for (size_t j = 0; j < rows.size(); ++j) {
                    const int m = rows[j], l = lane_of[j];
                    const int sh = cost[j][l] == 0 ? shift_of(m, l, prefer) : shift_of(m, l, group);
                    b.row[(size_t)s * 32 + l] = m;
                    b.start[(size_t)s * 32 + l] = first[m] - sh;
                }
            }
        }
    }
    b.w.assign((size_t)b.total_trip * 32, 0.0f);
    for (int s = 0; s < b.n_slots; ++s)
        for (int l = 0; l < 32; ++l) {
            const int idx = s * 32 + l;
            const int m = b.row[idx];
            if (m < 0) {
                b.start[idx] = l % group;  // idle lane: harmless conflict-free reads, zero weights
                continue;
            }
            int st = b.start[idx];
            while (st + b.trip[s] > cap && st >= group) st -= group;  // keep the window inside the tile
            b.start[idx] = st;
            for (int i = 0; i < b.trip[s]; ++i) {
                const int k = st + i;
                if (k < n_bins) b.w[((size_t)b.wbase[s] + i) * 32 + l] = dense[(size_t)m * n_bins + k];
            }
        }
    return b;
}

// true iff expanding the banded tables reproduces the dense matrix exactly (plan creation
// refuses to continue otherwise: a wrong table would silently change the features)
inline bool verify_banded(const BandedMel& b, const std::vector<float>& dense, int cap) {
    std::vector<float> re((size_t)b.n_mels * b.n_bins, 0.0f);
    std::vector<char> seen(b.n_mels, 0);
    for (int s = 0; s < b.n_slots; ++s) {
        if (b.trip[s] % 8) return false;
        for (int l = 0; l < 32; ++l) {
            const int idx = s * 32 + l, m = b.row[idx], st = b.start[idx];
            if (st < 0 || st + b.trip[s] > cap) return false;
            for (int i = 0; i < b.trip[s]; ++i) {
                const float w = b.w[((size_t)b.wbase[s] + i) * 32 + l];
                if (m < 0) {
                    if (w != 0.0f) return false;
                    continue;
                }
                const int k = st + i;
                if (k < b.n_bins)
                    re[(size_t)m * b.n_bins + k] = w;
                else if (w != 0.0f)
                    return false;
            }
            if (m >= 0) seen[m] = 1;
        }
    }
    for (int m = 0; m < b.n_mels; ++m)
        if (!seen[m]) return false;
    return re == dense;
}

// ------------------------------------------------------------------ inter-pass twiddles
// N = 32 * n2 point transform split as n = n1 + 32*n2, k = n2_count*k1 + k2:
// plane[k2][lane] = exp(-2*pi*i*lane*k2/N) stored as (cos, -sin).
inline std::vector<float> twiddle_plane(int n, int n_k2) {
    std::vector<float> t((size_t)n_k2 * 32 * 2);
    for (int k2 = 0; k2 < n_k2; ++k2)
        for (int l = 0; l < 32; ++l) {
            const double a = 2.0 * kPi * ((double)l * k2) / n;
            t[((size_t)k2 * 32 + l) * 2 + 0] = (float)cos(a);
            t[((size_t)k2 * 32 + l) * 2 + 1] = (float)(-sin(a));
        }
    return t;
}

// paired variant: plane[(p*32 + lane)*4 + {0,1,2,3}] = (cos_p, cos_{p+16}, -sin_p, -sin_{p+16}), k2 = p / p+16
inline std::vector<float> twiddle_plane_paired(int n) {
    std::vector<float> t((size_t)16 * 32 * 4);
    for (int p = 0; p < 16; ++p)
        for (int l = 0; l < 32; ++l) {
            const double a0 = 2.0 * kPi * ((double)l * p) / n, a1 = 2.0 * kPi * ((double)l * (p + 16)) / n;
            float* q = &t[((size_t)p * 32 + l) * 4];
            q[0] = (float)cos(a0);
            q[1] = (float)cos(a1);
            q[2] = (float)(-sin(a0));
            q[3] = (float)(-sin(a1));
        }
    return t;
}

}  // namespace hmfe
