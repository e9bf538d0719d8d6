// Guarded fast-math building blocks shared by the default eMBB kernels (embb_smem.cu, embb_fast.cu).
#pragma once
#include "embb_device.cuh"

namespace rs {

// RS_EXP: bit mask of timing experiments (results become wrong; never set in a release build)
//   1: window-mean trace loads replaced by a constant   2: MI-loop trace loads replaced by a constant
//   4: skip the MI loop                                   8: skip the contended PF iterations
#ifndef RS_EXP
#define RS_EXP 0
#endif
#if RS_EXP & 1
#define LDQ_B(p) make_int4(1 << 22, 2 << 22, 3 << 22, 4 << 22)
#elif defined(RS_LD_CG)
#define LDQ_B(p) __ldcg(p)
#else
#define LDQ_B(p) __ldg(p)
#endif
#if RS_EXP & 2
#define LDQ_D(p) make_int4(1 << 22, 2 << 22, 3 << 22, 4 << 22)
#elif defined(RS_LD_CG)
#define LDQ_D(p) __ldcg(p)
#else
#define LDQ_D(p) __ldg(p)
#endif
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void atomic_max_float(float *addr, float v) {   // v >= 0
    atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
}

__device__ __forceinline__ void prefetch_l1(const void *ptr) { asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr)); }

// Fading traces on the fast paths: int32 fixed point with 22 fractional bits.  |v| < 127 dB (checked when the tables
// are built), so a value is < 2^29 and the four values of an aligned quad sum without overflow in 32 bits; the
// representation error of a window mean is <= 2^-23 = 1.2e-7 dB, 17x inside the rounding guard below.
constexpr int FIX_BITS = 22;
constexpr double FIX_ONE = 4194304.0;
constexpr float FIX_SCALE = 1.0f / 4194304.0f;
constexpr double SNR_ROUND_GUARD = 2e-6;                     // |frac(mean) - 0.5| below this: round(np.mean(snr)) is re-done in fp64
constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int QUADS_PER_COL = TRACE_ROWS / 4;                // 25: quads never straddle the row wrap

// (b * bits) / slot_length, exactly rounded: q = RN(y * 1000), r = y - q * 1e-3 (exact, FMA),
// q' = RN(q + r * 1000).  rs_selftest checks q' == y / 1e-3 for every bits in the domain.
__device__ __forceinline__ double b_bits_over_slot(int bits) {
    const double y = __dmul_rn(PF_B, (double)bits);
    const double q = __dmul_rn(y, 1000.0);
    const double r = __fma_rn(-q, SLOT_LEN, y);
    return __fma_rn(r, 1000.0, q);
}

// Exact integer sum of the window [row0, row0 + n) of one trace column; aligned quads, end quads masked.
__device__ __forceinline__ int quad_total(const int4 v) { return (v.x + v.y) + (v.z + v.w); }   // < 2^31, see FIX_BITS
__device__ __forceinline__ int wrap_quad(int q) {             // q < 3 * QUADS_PER_COL
    q -= q >= QUADS_PER_COL ? QUADS_PER_COL : 0;
    q -= q >= QUADS_PER_COL ? QUADS_PER_COL : 0;
    return q;
}
#ifdef RS_WSUM_V1
template <int WIDE = 0>
__device__ __forceinline__ long long window_sum_fix(const int32_t *col, int row0, int n) {
    const int4 *col4 = reinterpret_cast<const int4 *>(col);
    const int lo = row0, hi = row0 + n;                      // absolute rows, may run past 100 (wrap)
    int q = lo >> 2;
    const int q_last = (hi - 1) >> 2;
    long long sum = 0;
    {   // first quad (masked below lo and at/after hi)
        const int qq = q >= QUADS_PER_COL ? q - QUADS_PER_COL : q;
        const int4 v = LDQ_B(col4 + qq);
        const int b = q << 2;
        int s = (b + 0 >= lo && b + 0 < hi) ? v.x : 0;
        s += (b + 1 >= lo && b + 1 < hi) ? v.y : 0;
        s += (b + 2 >= lo && b + 2 < hi) ? v.z : 0;
        s += (b + 3 >= lo && b + 3 < hi) ? v.w : 0;
        sum = s;
        ++q;
    }
    int qq = q;
    while (qq >= QUADS_PER_COL) qq -= QUADS_PER_COL;
    if (WIDE) {                                              // latency variant (small batches): 16 independent loads in flight
#pragma unroll 1
        for (; q + 16 <= q_last; q += 16) {
            int4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { int i = qq + j; if (i >= QUADS_PER_COL) i -= QUADS_PER_COL; v[j] = LDQ_B(col4 + i); }
            long long part = 0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) part += (long long)quad_total(v[j]) + (long long)quad_total(v[j + 1]);
            sum += part;
            qq += 16;
            if (qq >= QUADS_PER_COL) qq -= QUADS_PER_COL;
        }
    }
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
    for (; q + 4 <= q_last; q += 4) {                        // interior quads, 4 independent loads in flight
        int i0 = qq, i1 = qq + 1, i2 = qq + 2, i3 = qq + 3;
        if (i1 >= QUADS_PER_COL) i1 -= QUADS_PER_COL;
        if (i2 >= QUADS_PER_COL) i2 -= QUADS_PER_COL;
        if (i3 >= QUADS_PER_COL) i3 -= QUADS_PER_COL;
        const int4 v0 = LDQ_B(col4 + i0), v1 = LDQ_B(col4 + i1), v2 = LDQ_B(col4 + i2), v3 = LDQ_B(col4 + i3);
        sum += ((long long)quad_total(v0) + (long long)quad_total(v1)) + ((long long)quad_total(v2) + (long long)quad_total(v3));
        qq += 4;
        if (qq >= QUADS_PER_COL) qq -= QUADS_PER_COL;
    }
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
    for (; q < q_last; ++q) {                                // remaining interior quads: no masks
        const int4 v = LDQ_B(col4 + qq);
        sum += (long long)quad_total(v);
        qq = (qq + 1 == QUADS_PER_COL) ? 0 : qq + 1;
    }
    if (q == q_last) {                                       // last quad (masked at/after hi)
        const int4 v = LDQ_B(col4 + qq);
        const int b = q << 2;
        int s = (b + 0 < hi) ? v.x : 0;
        s += (b + 1 < hi) ? v.y : 0;
        s += (b + 2 < hi) ? v.z : 0;
        s += (b + 3 < hi) ? v.w : 0;
        sum += (long long)s;
    }
    return sum;
}
#else
// Every wait on a load below is an L2 round trip (each lane reads its own column), and those waits are a third of the
// kernel's stall samples; so the first quad, the last quad and the 1..3 interior quads that do not fill a batch of four
// are all issued together, and only the full batches wait once each: a 33-PRB window costs 2 round trips instead of 6.
template <int WIDE = 0>
__device__ __forceinline__ long long window_sum_fix(const int32_t *col, int row0, int n) {
    const int4 *col4 = reinterpret_cast<const int4 *>(col);
    const int lo = row0, hi = row0 + n;                      // absolute rows, may run past 100 (wrap)
    const int qf = lo >> 2, ql = (hi - 1) >> 2;              // first / last quad (qf <= 24: no wrap)
    const int n_int = ql - qf - 1;                           // interior quads (-1: the window sits in one quad)
    const int rem = n_int > 0 ? (n_int & 3) : 0;             // taken from the END of the interior: quads ql-rem .. ql-1
    const int4 zero = make_int4(0, 0, 0, 0);
    const int4 vf = LDQ_B(col4 + qf);
    int4 vl = zero, r0 = zero, r1 = zero, r2 = zero;
    if (ql != qf) vl = LDQ_B(col4 + wrap_quad(ql));
    if (rem > 0) r0 = LDQ_B(col4 + wrap_quad(ql - 1));
    if (rem > 1) r1 = LDQ_B(col4 + wrap_quad(ql - 2));
    if (rem > 2) r2 = LDQ_B(col4 + wrap_quad(ql - 3));
    long long sum;
    {   // first quad (masked below lo and at/after hi), last quad (masked at/after hi; all zero when ql == qf)
        const int bf = qf << 2, bl = ql << 2;
        int s = (bf + 0 >= lo && bf + 0 < hi) ? vf.x : 0;
        s += (bf + 1 >= lo && bf + 1 < hi) ? vf.y : 0;
        s += (bf + 2 >= lo && bf + 2 < hi) ? vf.z : 0;
        s += (bf + 3 >= lo && bf + 3 < hi) ? vf.w : 0;
        int t = (bl + 0 < hi) ? vl.x : 0;
        t += (bl + 1 < hi) ? vl.y : 0;
        t += (bl + 2 < hi) ? vl.z : 0;
        t += (bl + 3 < hi) ? vl.w : 0;
        sum = ((long long)s + (long long)t) + (((long long)quad_total(r0) + (long long)quad_total(r1)) + (long long)quad_total(r2));
    }
    int qq = qf + 1;                                         // wrapped index of the next interior quad
    qq -= qq >= QUADS_PER_COL ? QUADS_PER_COL : 0;
    int batches = n_int > 0 ? (n_int >> 2) : 0;
    if (WIDE) {                                              // latency variant (small batches): 16 independent loads in flight
#pragma unroll 1
        for (; batches >= 4; batches -= 4) {
            int4 v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { int i = qq + j; if (i >= QUADS_PER_COL) i -= QUADS_PER_COL; v[j] = LDQ_B(col4 + i); }
            long long part = 0;
#pragma unroll
            for (int j = 0; j < 16; j += 2) part += (long long)quad_total(v[j]) + (long long)quad_total(v[j + 1]);
            sum += part;
            qq += 16;
            if (qq >= QUADS_PER_COL) qq -= QUADS_PER_COL;
        }
    }
#pragma unroll 1
    for (; batches > 0; --batches) {                         // interior quads, 4 independent loads in flight
        int i1 = qq + 1, i2 = qq + 2, i3 = qq + 3;
        if (i1 >= QUADS_PER_COL) i1 -= QUADS_PER_COL;
        if (i2 >= QUADS_PER_COL) i2 -= QUADS_PER_COL;
        if (i3 >= QUADS_PER_COL) i3 -= QUADS_PER_COL;
        const int4 v0 = LDQ_B(col4 + qq), v1 = LDQ_B(col4 + i1), v2 = LDQ_B(col4 + i2), v3 = LDQ_B(col4 + i3);
        sum += ((long long)quad_total(v0) + (long long)quad_total(v1)) + ((long long)quad_total(v2) + (long long)quad_total(v3));
        qq += 4;
        if (qq >= QUADS_PER_COL) qq -= QUADS_PER_COL;
    }
    return sum;
}
#endif

// Window sum from the per-column prefix table (Tables::trace_pre): rows [row0, row0 + n) with row0 < 100, n <= 200, rows
// wrapping at 100 (channel_models.py:144-148).  E(x) = (x / 100) T + pre[x % 100] with T = pre[100] is the prefix over the
// wrapped rows, so the sum is E(row0 + n) - E(row0): two loads, plus T when the window wraps.  int32 arithmetic is
// modular; the true sum fits (checked when the table is built), so the difference is exact.
__device__ __forceinline__ int window_sum_prefix(const int32_t *pre, int row0, int n) {
    const int hi = row0 + n;                                 // <= 299
    const int wraps = (hi >= TRACE_ROWS) + (hi >= 2 * TRACE_ROWS);
    const int lo_v = __ldg(pre + row0);
    const int hi_v = __ldg(pre + (hi - wraps * TRACE_ROWS));
    int sum = hi_v - lo_v;
    if (wraps) sum += wraps * __ldg(pre + TRACE_ROWS);
    return sum;
}

// x / n for a small positive integer count n, exactly rounded: q = RN(x * RN(1/n)), r = x - q n (exact, FMA),
// q' = RN(q + r * RN(1/n)) (Markstein; rs_selftest samples the identity against IEEE division).
__device__ __forceinline__ double div_count(double x, int n) {
    if (n <= 1) return x;
    const double dn = (double)n, rcp = __drcp_rn(dn);
    const double q = __dmul_rn(x, rcp);
    return __fma_rn(__fma_rn(-q, dn, x), rcp, q);
}

// exact fp64 window mean (same operation order as embb_step.cu); rare
static __device__ __noinline__ double window_mean_fp64(const double *col, int row0, int n, double nominal) {
    double sum = 0.0;
    int row = row0;
    for (int j = 0; j < n; ++j) {
        sum += col[row] + nominal;
        row = (row + 1 == TRACE_ROWS) ? 0 : row + 1;
    }
    return sum / (double)n;
}

// exact reception probability (reference fp64 path); rare
static __device__ __noinline__ double response_exact(const Tables &tb, int mcs, size_t col_off, int row0, int n, double nominal) {
    return response_fp64(tb, mcs, tb.trace + col_off, row0, n, nominal);
}


// State a rare-event call may change; copied in/out around the call so that the hot loop keeps it in registers.
struct RanCtx { uint32_t c_ran, c_chan, c_vbr, next_dep, flags; int n_ues, cbr_next, vbr_next; };

}  // namespace rs
