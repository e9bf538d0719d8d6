// Device helpers shared by the eMBB step kernels (fp64, reference operation order).
#pragma once
#include <cuda_runtime.h>

#include "philox.cuh"
#include "ranslice_state.cuh"

namespace rs {

// ProportionalFair / UE EWMA constants: b = 1/window, a = 1 - b (schedulers.py:16-17, slice_ran.py:28-29)
constexpr double PF_B = 1.0 / 50;
constexpr double PF_A = 1 - PF_B;
constexpr double SLOT_LEN = 1e-3;

// MI sigmoid parameters by modulation (channel_models.py:268-270)
static __device__ __constant__ double c_MI_X0[3] = {-0.25040431, 5.12440916, 9.16962738};
static __device__ __constant__ double c_MI_K[3] = {0.31591749, 0.25423209, 0.22298101};

// node_b.py:71-74 contiguous windows in slice order; out-of-contract actions are clamped and flagged
__device__ __forceinline__ void slice_window(const StepParams &p, int env, int s, int &i_prb, int &n_prbs,
                                             uint32_t &flags) {
    const int32_t *a = p.action + (size_t)env * p.S;
    int off = 0, mine = 0;
    for (int j = 0; j <= s; ++j) {
        int v = a[j];
        if (v < 0) { v = 0; flags |= 4u; }
        if (off + v > p.n_prbs) { v = p.n_prbs - off; flags |= 4u; }
        if (j == s) mine = v; else off += v;
    }
    i_prb = off; n_prbs = mine;
}

// window of a unit precomputed by window_kernel (embb_fast.cu)
__device__ __forceinline__ void unpack_window(uint32_t w, int &i_prb, int &n_prbs) { i_prb = (int)(w & 0xFFFFu); n_prbs = (int)(w >> 16); }

__device__ __forceinline__ int exp_slots_ms(PhiloxStream &r, double scale) {   // np.rint(exp / slot_length)
    return __double2int_rn(r.exponential(scale) / SLOT_LEN);
}
__device__ __forceinline__ int exp_slots(PhiloxStream &r, double scale) {      // np.rint(exp)
    return __double2int_rn(r.exponential(scale));
}

// channel_models.py:44-60,74
__device__ __forceinline__ double find_y(double x1, double y1, double x2, double y2, double x) {
    const double m = (y2 - y1) / (x2 - x1);
    const double b = -m * x1 + y1;
    return m * x + b;
}
__device__ __forceinline__ bool in_cell(double x, double y) {
    return (y > find_y(0, 0.5, 0.25, 0, x)) && (y > find_y(0.75, 0, 1, 0.5, x)) &&
           (y < find_y(0, 0.5, 0.25, 1, x)) && (y < find_y(0.75, 1, 1, .5, x));
}
// macro_cell (channel_models.py:80-97) with location (:62-68) and antenna_pattern (:76-78)
static __device__ __noinline__ double draw_nominal_sinr(PhiloxStream &r, double A, double B) {
    double x, y;
    do { x = r.u01(); y = r.u01(); } while (!in_cell(x, y));
    const double logf = r.normal(0.0, 10.0);
    const double x_t = x - 0.5 / 2;
    const double d = sqrt(x_t * x_t + y * y);
    const double theta = acos(x_t / d) * (180.0 / 3.141592653589793) - 60;
    const double R = fmax(d * 2, 0.1);
    const double ap = 12 * ((theta / 65) * (theta / 65));
    const double G = 15 + -1 * fmin(ap, 20.0);
    double L = A + B * log10(R);
    const double fspl = 20 * log10(2.0) + 92.45 + 2.6 * 10 * log10(R);
    L = fmax(L, fspl);
    const double rx = 30 - fmax(L + logf - G, 70.0);
    return rx - (-110) - 9;
}

// SINRSelectiveFading.get_snr index walk (channel_models.py:171-191).  The re-draw at a trace end happens once per
// ~10 000 TTIs of a UE: it is kept out of line so that its two Philox evaluations (~200 instructions) do not sit in
// the middle of the per-TTI instruction stream (the hot loop is instruction-fetch sensitive).
static __device__ __noinline__ void walk_trace_redraw(PhiloxStream &r, int &index, int &step) {
    for (;;) {
        if (index >= N_SAMPLES || index < 0) {
            index = (int)r.integers(N_SAMPLES);
            step = r.integers(2) ? 1 : -1;
        }
        if (index != N_SAMPLES - 1) break;       // column 10000 is all-NaN
        index += step;
    }
}
__device__ __forceinline__ void walk_trace(PhiloxStream &r, int &index, int &step) {
    index += step;
    if (index >= N_SAMPLES - 1 || index < 0) walk_trace_redraw(r, index, step);
}

// VbrSource.step (traffic_generators.py:70-99) on a UE record held in registers; returns this slot's bits.
// The order of the burst list never matters (bits = 1000 per non-ending burst; draws only on arrival), so
// slots are not compacted: togo[j] == 0 marks a free slot, and a length drawn as 0 (a burst that never
// ends) is stored as -1, which behaves identically (never hits 0; saturates at -32768).
__device__ __forceinline__ int vbr_source_step(UeRec &r, PhiloxStream &r_vbr, uint32_t &flags) {
    int bits = 0, nb = 0;
#pragma unroll
    for (int j = 0; j < MAX_BURSTS; ++j) {
        int tg = r.togo[j];
        if (tg != 0) {
            tg = max(tg - 1, -32768);
            if (tg != 0) { bits += 1000; ++nb; }              // == 0: burst ends, contributes nothing this slot
            r.togo[j] = (int16_t)tg;
        }
    }
    int vn = r.vnext - 1;
    if (vn == 0) {
        int len = exp_slots(r_vbr, 500.0);
        if (len == 0) len = -1;
        bool placed = false;
#pragma unroll
        for (int j = 0; j < MAX_BURSTS; ++j)
            if (!placed && r.togo[j] == 0) { r.togo[j] = (int16_t)len; placed = true; }
        if (placed) ++nb; else flags |= 2u;
        vn = exp_slots(r_vbr, 1000.0);
    }
    r.vnext = vn;
    r.nb = nb;
    return bits;
}

// 128-bit moves of a UE record between HBM and registers
__device__ __forceinline__ void load_rec(const UeRec *g, UeRec &r) {
    const int4 *s = reinterpret_cast<const int4 *>(g);
    int4 *d = reinterpret_cast<int4 *>(&r);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}
__device__ __forceinline__ void store_rec(UeRec *g, const UeRec &r) {
    const int4 *s = reinterpret_cast<const int4 *>(&r);
    int4 *d = reinterpret_cast<int4 *>(g);
    d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3];
}

__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

// MCSCodeset.response (channel_models.py:297-313) over the sub-band col[row0 .. row0+n) (rows wrap at 100)
__device__ __forceinline__ double response_fp64(const Tables &tb, int mcs, const double *col, int row0, int n,
                                                double nominal) {
    double s;
    if (n > 1) {
        const int m = tb.mod[mcs];
        const double x0 = c_MI_X0[m], k = c_MI_K[m];
        double sum = 0.0;
        int row = row0;
        for (int j = 0; j < n; ++j) {
            const double snr = col[row] + nominal;
            sum += 1.0 / (1.0 + exp(-k * (snr - x0)));
            row = (row + 1 == TRACE_ROWS) ? 0 : row + 1;
        }
        const double avg = sum / (double)n;
        s = -(1.0 / k) * log(1.0 / avg - 1.0) + x0;     // inv_sigmoid, channel_models.py:39-41
    } else {
        s = col[row0] + nominal;
    }
    return sigmoid_d(tb.A * (s - tb.snr_ref[mcs]) - tb.B);   // estimate_rx_prob, :281-286
}

// End of observation period for one eMBB unit: normalised state (slice_ran.py:321-325), SLA predicate
// (:307-319), L1 label (slice_l1.py:160-171); acc = the 10 raw accumulators in state-variable order.
__device__ __forceinline__ void finish_embb_unit(const StepParams &p, const EmbbState &st, int env, int s, int u,
                                                 const double (&acc)[10], uint32_t flags) {
    float *obs = p.obs + (size_t)env * p.V + s * 10;
#pragma unroll
    for (int j = 0; j < 10; ++j) {
        obs[j] = (float)(acc[j] / p.norm_embb[j]);
        st.acc[(size_t)u * 10 + j] = acc[j];
    }
    const double sps = (double)p.slots;
    const bool cbr_ok = acc[1] / p.obs_time > 10e6 || acc[2] / sps > 20.0 || acc[3] / sps < 10e4;
    const bool vbr_ok = acc[6] / p.obs_time > 15e6 || acc[7] / sps > 30.0 || acc[8] / sps < 15e4;
    const int viol = !(cbr_ok && vbr_ok);
    p.violations[(size_t)env * p.S + s] = viol;
    p.labels[(size_t)env * p.S + s] = viol ? -1 : 1;
    if (flags) atomicOr(p.flags_acc + env, flags);
}

}  // namespace rs
