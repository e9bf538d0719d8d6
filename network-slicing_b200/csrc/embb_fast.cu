// K1 general variant ("sorted unit-per-thread, guarded fast math", state in global memory) + the sort pre-pass shared
// with the default shared-memory kernel (embb_smem.cu).
//
// Same algorithm and results as embb_step.cu (bit-exact: tests compare both against the oracle), but
// organised for lane utilisation and instruction count on B200.  One thread still owns one
// (env, eMBB slice) unit for the whole observation period; what changes is who shares its warp and
// how each decision is evaluated:
//
//  * window_kernel / scan_kernel / scatter_kernel sort the units of a step (counting sort) by
//    (live UEs, n_prbs, PF contention class), descending.  Every inner loop of the step has a trip
//    count given by one of those three numbers (window mean: UEs x PRBs, MI: PRBs, PF: contended RB
//    chunks), so after the sort the 32 lanes of a warp run the same loops the same number of times.
//    The unsorted variant measured 6.2 of 32 lanes active per instruction.
//  * every phase of a TTI starts with __syncwarp: lanes that drifted apart on a rare branch (arrival,
//    fp64 re-evaluation) are NOT re-merged by the hardware on their own (measured: lanes of one warp
//    working on different TTIs).
//  * the window mean behind round(np.mean(snr)) (slice_ran.py:43-45) is an exact int64 sum over a
//    2^-22 fixed-point copy of the traces, read as aligned 128-bit quads with the two end quads
//    masked; only when the mean lies within SNR_ROUND_GUARD (2e-6) of a rounding boundary is it recomputed in fp64.
//  * the MI-effective-SNR reception probability (channel_models.py:297-313) is evaluated in fp32
//    (ex2/rcp/lg2 SFU ops) together with a bound eps on |p32 - p64|; the Bernoulli decision
//    u < p (slice_l1.py:223) is taken from p32 unless |u - p32| <= eps, in which case p is
//    recomputed in fp64 exactly like the reference.  Saturated sub-bands (mean MI within 1e-4 of 0
//    or 1) give p == 1.0 / p < 2^-53 in fp64 and need no evaluation at all.  The per-PRB MI loop is
//    flattened over (served UE, quad) so that its trip count is ~n_prbs/4 for every lane.
//  * the PF loop (schedulers.py:47-63) keeps an fp32 copy of the metric rate/th and takes the
//    argmax from it when the runner-up is more than 1e-6 (relative) behind; otherwise the candidates
//    are compared with the exact fp64 quotient.  (b*bits)/slot_length is an exactly rounded
//    division by a constant (two FMAs, Markstein; exhaustively checked over the domain by
//    rs_selftest).  Closed form once a single UE is backlogged, early exit when every queue is
//    drained (the remaining PRBs go to UE 0 with 0 bits, schedulers.py:52 argmax of all-zero).
#include "embb_device.cuh"
#include "embb_fastmath.cuh"
#include "embb_ran.cuh"

namespace rs {

constexpr uint32_t KEY_BINS = SORT_BINS;

#ifndef RS_FAST_MIN_BLOCKS
#define RS_FAST_MIN_BLOCKS 4      // resident 128-thread blocks per SM the register budget is sized for
#endif

// Contention class: predicted contended RB-loop iterations per TTI of this step, from the previous
// step's count scaled to the new PRB allocation (hint = iterations << 8 | previous n_prbs); log2 buckets.
// Only a scheduling hint: it decides which units share a warp, never what they compute.
__device__ __forceinline__ uint32_t contention_class(uint32_t hint, int n_now, int slots) {
    const uint32_t iters_prev = hint >> 8, n_prev = max(hint & 0xFFu, 1u);
    const uint32_t pred = (uint32_t)(((unsigned long long)iters_prev * (unsigned)n_now) / ((unsigned long long)n_prev * (unsigned)slots));
    return pred == 0 ? 0u : min(31u - (uint32_t)__clz(pred) + 1u, 7u);
}
constexpr uint32_t HEAVY_BIT = 1u << 15;   // units that own a pair of lanes sort ahead of everything else
__device__ __forceinline__ uint32_t sort_key(int n_prbs, uint32_t hint, int n_ues, int slots, bool heavy) {
    // Key order measured on B200 (65 536 envs, M env-steps/s): (n_prbs, class, UEs) 12.43 | (n_prbs, UEs, class) 12.85 |
    // (n_prbs/4, UEs, class) 13.04 | (class, n_prbs, UEs) 12.32 | (UEs, class, n_prbs) 13.26 | (UEs, n_prbs, class) 13.41;
    // a second hint (UEs served per TTI in the previous step) on top of the last one: 13.35 (no gain)
    return (heavy ? HEAVY_BIT : 0u) | ((uint32_t)min(n_ues, 15) << 11) | ((uint32_t)n_prbs << 3) | contention_class(hint, n_prbs, slots);
}

// units routed to the warp-per-unit kernel by the default variant (heavy_min_ues < 2^30 marks the shared-memory route)
__device__ __forceinline__ bool is_heavy(const EmbbState &st, int u, int heavy_min_ues, int n_ues, int slots) {
#ifdef RS_HEAVY_RAW
    return st.heavy_thr > 0 && heavy_min_ues < (1 << 30) && (int)(st.hint[u] >> 8) >= st.heavy_thr;
#else
    // Predicted contended chunks of THIS step: last step's count scaled to the new PRB allocation (a saturated slice contends
    // for every chunk it is given); a slice that had no PRBs at all has queued a whole period of traffic and counts as
    // saturated.  Measured against the raw previous count: 4096 envs 1.95 -> 1.60 ms/step, env step under the KBRL policy at
    // 16 384 envs 3.96 -> 3.13 ms (profiles/r02f_heavy_prediction.txt).
    if (st.heavy_thr <= 0 || heavy_min_ues >= (1 << 30)) return false;
    const uint32_t hint = st.hint[u], n_prev = hint & 0xFFu, n_now = st.win[u] >> 16;
    if (n_prev == 0) return n_ues >= 2 && (unsigned)(slots >> 1) * n_now >= (unsigned)st.heavy_thr;
    return (unsigned long long)(hint >> 8) * n_now >= (unsigned long long)st.heavy_thr * n_prev;
#endif
}

// ---------------------------------------------------------------------------------------------
// Pre-pass 1: PRB windows of all eMBB units of a step (node_b.py:71-74) + histogram of sort keys.
// Units whose live-UE count fits ONE lane's slots of the shared-memory kernel (embb_smem.cu) are counting-
// sorted into the front list of perm[]; units that need TWO lanes' slots (heavy_min_ues <= n_ues <=
// max_front_ues) are placed before them as (unit, -1) pairs -- the odd lane only lends its shared memory --;
// the rest is appended from the back (list L) for the general kernel below.
__global__ void __launch_bounds__(256) window_kernel(const __grid_constant__ StepParams p,
                                                     const __grid_constant__ EmbbState st, const int max_front_ues,
                                                     const int heavy_min_ues) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.N) return;
    const int32_t *a = p.action + (size_t)env * p.S;
    int off = 0;
    uint32_t flags = 0;
    for (int s = 0; s < p.n_embb; ++s) {
        int v = a[s];
        if (v < 0) { v = 0; flags |= 4u; }
        if (off + v > p.n_prbs) { v = p.n_prbs - off; flags |= 4u; }
        const int u = env * p.n_embb + s;
        st.win[u] = (uint32_t)off | ((uint32_t)v << 16);
        st.cur_prbs[u] = v;
        const int n_ues = st.hdr[u].n_ues;
        bool to_warp = is_heavy(st, u, heavy_min_ues, n_ues, p.slots);          // long PF loop last step: warp-per-unit kernel
        if (to_warp) {
            const int pos = atomicAdd(&st.wlist[st.U], 1);
            if (pos < st.heavy_cap) st.wlist[pos] = u;
            else { atomicSub(&st.wlist[st.U], 1); st.hint[u] = 0xFFu; to_warp = false; }   // list full: a normal lane after all (hint = no chunks at 255 PRBs, so that scatter_kernel agrees)
        }
        if (to_warp) {
        } else if (n_ues < heavy_min_ues) {
            atomicAdd(&st.hist[sort_key(v, st.hint[u], n_ues, p.slots, false)], 1u);
        } else if (n_ues <= max_front_ues) {
            atomicAdd(&st.hist[sort_key(v, st.hint[u], n_ues, p.slots, true)], 2u);      // (unit, pad) pair
        } else {
            st.perm[st.perm_len - 1 - (int)atomicAdd(&st.hist[2 * KEY_BINS + 1], 1u)] = u;  // list L
        }
        off += v;
    }
    if (flags) atomicOr(p.flags_acc + env, flags);
}

// Pre-pass 2: exclusive prefix of the histogram in DESCENDING key order (long units first), as two small kernels over
// SCAN_BLOCKS blocks (a single 1024-thread block took 73 us, 1.4 % of a step): block-local scan of 1024 bins + block
// totals, then every block adds the totals of the blocks ahead of it.  r = KEY_BINS - 1 - bin is the scan index.
constexpr int SCAN_THREADS = KEY_BINS / SCAN_BLOCKS / 4;       // 4 bins per thread, one aligned 16-byte load
static_assert(SCAN_THREADS == 256 && KEY_BINS % (SCAN_BLOCKS * 4) == 0, "scan geometry");
__global__ void __launch_bounds__(SCAN_THREADS) scan_local_kernel(const __grid_constant__ EmbbState st) {
    __shared__ uint32_t s_warp[SCAN_THREADS / 32];
    const int r0 = (blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    const uint4 v = *reinterpret_cast<const uint4 *>(st.hist + (KEY_BINS - 4 - r0));   // bins r0+3 .. r0 (x is the highest r)
    const uint32_t c0 = v.w, c1 = v.z, c2 = v.y, c3 = v.x;
    const uint32_t sum = c0 + c1 + c2 + c3;
    uint32_t inc = sum;                                          // inclusive scan of the thread sums: warp shuffles, then warp totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if ((threadIdx.x & 31) >= d) inc += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t before = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += s_warp[w];
    const uint32_t e = before + inc - sum;                       // exclusive prefix of this thread inside the block
    uint4 o;
    o.w = e; o.z = e + c0; o.y = e + c0 + c1; o.x = e + c0 + c1 + c2;
    *reinterpret_cast<uint4 *>(st.hist + KEY_BINS + (KEY_BINS - 4 - r0)) = o;
    if (threadIdx.x == SCAN_THREADS - 1) st.hist[2 * KEY_BINS + 4 + blockIdx.x] = before + inc;   // block total
}
__global__ void __launch_bounds__(SCAN_THREADS) scan_offset_kernel(const __grid_constant__ EmbbState st) {
    __shared__ uint32_t s_off;
    if (threadIdx.x < 32) {                                      // totals of the blocks ahead (SCAN_BLOCKS = 64: two per lane)
        uint32_t t = 0;
        for (int b = threadIdx.x; b < (int)blockIdx.x; b += 32) t += st.hist[2 * KEY_BINS + 4 + b];
#pragma unroll
        for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (threadIdx.x == 0) s_off = t;
    }
    __syncthreads();
    const uint32_t off = s_off;
    const int r0 = (blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
    uint4 *o = reinterpret_cast<uint4 *>(st.hist + KEY_BINS + (KEY_BINS - 4 - r0));
    uint4 v = *o;
    v.x += off; v.y += off; v.z += off; v.w += off;
    *o = v;
    if (r0 == KEY_BINS - (int)HEAVY_BIT) st.hist[2 * KEY_BINS + 3] = v.w >> 1;   // bin HEAVY_BIT - 1: entries ahead of the first single-lane bin = 2 x pairs
    if (blockIdx.x == SCAN_BLOCKS - 1 && threadIdx.x == SCAN_THREADS - 1)       // entries of the front list
        { st.hist[2 * KEY_BINS + 0] = off + st.hist[2 * KEY_BINS + 4 + blockIdx.x]; st.hist[2 * KEY_BINS + 2] = st.hist[2 * KEY_BINS + 1]; }   // + list L as routed by the pre-pass
}

// Pre-pass 3: scatter.
__global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ StepParams p,
                                                      const __grid_constant__ EmbbState st, const int max_front_ues,
                                                      const int heavy_min_ues) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= st.U) return;
    const int n_ues = st.hdr[u].n_ues;
    if (n_ues > max_front_ues || is_heavy(st, u, heavy_min_ues, n_ues, p.slots)) return;
    const bool heavy = n_ues >= heavy_min_ues;
    const uint32_t key = sort_key((int)(st.win[u] >> 16), st.hint[u], n_ues, p.slots, heavy);
    const uint32_t pos = atomicAdd(&st.hist[KEY_BINS + key], heavy ? 2u : 1u);
    st.perm[pos << st.dil] = u;                                // diluted list: the lanes in between stay -1 (launch_embb_sort)
    if (heavy) st.perm[(pos << st.dil) + 1] = -1;                          // the odd lane only lends its shared-memory slots
}

template <int K>
__global__ void __launch_bounds__(128, RS_FAST_MIN_BLOCKS) embb_step_fast(const __grid_constant__ StepParams p,
                                                         const __grid_constant__ EmbbState st,
                                                         const __grid_constant__ Tables tb, const int back_list) {
    // small lookup tables: constant-bank reads with divergent indices serialise, shared memory does not
    __shared__ int16_t s_rate[256];
    __shared__ int8_t s_mcs[256];
    __shared__ float s_ref[26];
    __shared__ int8_t s_mod[26];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { s_rate[i] = tb.lut_rate[i]; s_mcs[i] = tb.lut_mcs[i]; }
    if (threadIdx.x < 26) { s_ref[threadIdx.x] = (float)tb.snr_ref[threadIdx.x]; s_mod[threadIdx.x] = tb.mod[threadIdx.x]; }
    __syncthreads();

    const int tix = blockIdx.x * blockDim.x + threadIdx.x;
    const int count = (int)st.hist[2 * KEY_BINS + back_list];           // front (sorted) list or back list L
    const unsigned warp_mask = __ballot_sync(0xffffffffu, tix < count);
    if (tix >= count) return;
    const int u = st.perm[back_list ? st.perm_len - 1 - tix : tix];    // (variant 2 has no pair entries: heavy_min = inf)
    const int env = u / p.n_embb, s = u - env * p.n_embb;
    int i_prb, n_prbs;
    unpack_window(st.win[u], i_prb, n_prbs);
    const int row_base = i_prb % TRACE_ROWS;
    uint32_t flags = 0;

    UnitHdr hdr = st.hdr[u];
    UeRec *ue = st.ue + (size_t)u * st.K;
    const uint64_t seed = p.seed0;
    const uint32_t genv = p.env0 + (uint32_t)env;   // global env id: Philox counter word 3
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c_ran = hdr.ctr[0];
    PhiloxStream r_chan{k0, k1, (uint32_t)s, STREAM_CHAN, hdr.ctr[1], genv}, r_rx{k0, k1, (uint32_t)s, STREAM_L1RX, hdr.ctr[2], genv},
        r_vbr{k0, k1, (uint32_t)s, STREAM_VBR, hdr.ctr[3], genv};

    int n_ues = hdr.n_ues, cbr_next = hdr.cbr_next, vbr_next = hdr.vbr_next;
    uint32_t clock = hdr.clock, next_dep = DEP_NEVER;
    for (int k = 0; k < n_ues; ++k) next_dep = min(next_dep, ue[k].dep_at);
    int a_traffic[2] = {0, 0}, a_th[2] = {0, 0}, a_prb[2] = {0, 0};     // slice_ran.py:270-273 reset_info
    double a_queue[2] = {0.0, 0.0}, a_snr[2] = {0.0, 0.0};
    unsigned trace_elems = 0, slow_snr = 0, slow_rx = 0, pf_iters = 0;
    const float Af = (float)tb.A, Bf = (float)tb.B;
#ifdef RS_WSUM_QUADS
    const double inv_n = n_prbs > 0 ? 1.0 / ((double)n_prbs * FIX_ONE) : 0.0;
#else
    const double inv_n = n_prbs > 0 ? tb.pre_inv / (double)n_prbs : 0.0;
#endif

    // per-UE scratch of one TTI (local memory, only the first n_ues entries are touched)
    long long qq[K];
    double th[K];
    float metf[K], nomf[K], mavg[K];
    int coloff[K], bits[K], pe_l[K];
    int16_t rate[K], new_bits[K];
    int8_t mcs[K];
    uint8_t rbs[K];

    for (int t = 1; t <= p.slots; ++t) {          // slot_counter == t (zeroed by reset_info each step)
        __syncwarp(warp_mask);
        ++clock;
        // ================= slice_ran.slot(): arrivals / departures only on event slots
        if (cbr_next == 0 || vbr_next == 0 || clock == next_dep) {
            RanCtx c{c_ran, r_chan.n, r_vbr.n, next_dep, flags, n_ues, cbr_next, vbr_next};
            ran_events(p, st.K, ue, k0, k1, genv, (uint32_t)s, t, clock, a_prb[0], a_th[0], c);
            c_ran = c.c_ran; r_chan.n = c.c_chan; r_vbr.n = c.c_vbr; next_dep = c.next_dep; flags = c.flags;
            n_ues = c.n_ues; cbr_next = c.cbr_next; vbr_next = c.vbr_next;
        } else { cbr_next -= 1; vbr_next -= 1; }

        // ================= per-UE traffic + SNR estimate (slice_l1.py:200-213); fills the PF scratch
        __syncwarp(warp_mask);
        long long queued = 0;
        int n_backlog = 0;
        uint32_t types = 0;
        for (int k = 0; k < n_ues; ++k) {
            UeRec r;
            load_rec(ue + k, r);
            if (k + 1 < n_ues) prefetch_l1(ue + k + 1);          // next record while this UE's trace window is summed
            int nb_bits;
            if ((r.meta & 1u) == 0) nb_bits = 500;               // CbrSource: 500000 b/s * 1e-3 every slot
            else { nb_bits = vbr_source_step(r, r_vbr, flags); types |= 1u << k; }
            new_bits[k] = (int16_t)nb_bits;
            r.queue += nb_bits;
            queued += r.queue;
            if (n_prbs > 0) {
                int index = (int)(r.meta >> 4), step = (r.meta & 8u) ? 1 : -1;
                const int fading = (int)((r.meta >> 1) & 3u);
                walk_trace(r_chan, index, step);               // channel_models.py:171-191
                r.meta = pack_meta((int)(r.meta & 1u), fading, step, index);
                const int col_off = (fading * N_SAMPLES + index) * TRACE_ROWS;
                coloff[k] = col_off;
                trace_elems += (unsigned)n_prbs;
#ifdef RS_WSUM_QUADS
                const long long isum = window_sum_fix(tb.trace_fix + col_off, row_base, n_prbs);
                double mean = (double)isum * inv_n + r.nominal;  // |mean - reference mean| < 2^-25 + few ulp
                const double guard = SNR_ROUND_GUARD;
#else
                const int isum = window_sum_prefix(tb.trace_pre + (fading * N_SAMPLES + index) * PRE_STRIDE, row_base, n_prbs);
                double mean = (double)isum * inv_n + r.nominal;  // |mean - reference mean| <= 2^-(pre_bits + 1) + few ulp
                const double guard = tb.pre_guard;
#endif
                const double fr = mean - floor(mean);
                const bool near = fabs(fr - 0.5) < guard;        // within the guard of a rounding boundary
                if (near || p.debug_check) {
                    const double exact = window_mean_fp64(tb.trace + col_off, row_base, n_prbs, r.nominal);
                    if (p.debug_check) atomic_max_float(st.dbg + 1, (float)(fabs(exact - mean) / guard));
                    if (near) { mean = exact; ++slow_snr; }
                }
                const int e_snr = __double2int_rn(mean);         // round(np.mean(snr)), slice_ran.py:43-45
                r.pe = (r.pe & 0xFFFF) | (e_snr << 16);
            }
            // write back what changed: quad 0 (meta, vnext), quad 2 (queue, pe, nb), quad 3 (bursts, VBR only)
            {
                int4 *g = reinterpret_cast<int4 *>(ue + k);
                const int4 *l = reinterpret_cast<const int4 *>(&r);
                g[0] = l[0]; g[2] = l[2];
                if (r.meta & 1u) g[3] = l[3];
            }
            // PF scratch (schedulers.py:37-45)
            const int e = min(max(r.pe >> 16, -128), 127) + 128;
            mcs[k] = s_mcs[e];
            rate[k] = s_rate[e];
            th[k] = r.th > 1.0 ? r.th : 1.0;
            qq[k] = r.queue;
            nomf[k] = (float)r.nominal;
            bits[k] = r.bits;                                    // stale values, kept if the slice is not scheduled
            pe_l[k] = r.pe;
            n_backlog += r.queue > 0;
        }
        // ================= scheduling + reception (slice_l1.py:215-224)
        __syncwarp(warp_mask);
        const bool scheduled = queued > 0 && n_prbs > 0;
        const unsigned sched_mask = __ballot_sync(warp_mask, scheduled);
        if (scheduled) {
            for (int k = 0; k < n_ues; ++k) { rbs[k] = 0; bits[k] = 0; }
            // ---- ProportionalFair.allocate RB loop (schedulers.py:47-63)
            int r = 0;
            if (n_backlog > 1)
                for (int k = 0; k < n_ues; ++k)
                    metf[k] = qq[k] > 0 ? (float)rate[k] * rcp_approx((float)th[k]) : 0.0f;
            // phase 1: contended chunks (>= 2 backlogged UEs); one uniform loop body for all lanes still in it
            while (n_backlog >= 2 && r < n_prbs) {
                ++pf_iters;
                if (RS_EXP & 8) break;
                const int c = min(n_prbs - r, 2);
                // argmax of rate * (queue > 0) / th, first maximum (np.argmax): fp32 copy, exact when close
                int idx = 0;
                float best = metf[0], second = -1.0f;
                for (int k = 1; k < n_ues; ++k) {
                    const float m = metf[k];
                    if (m > best) { second = best; best = m; idx = k; }
                    else second = fmaxf(second, m);
                }
                if (second >= best * (1.0f - 1e-6f)) {                      // too close for fp32: exact quotients
                    const float lim = best * (1.0f - 1e-6f);
                    double best64 = -1.0;
                    for (int k = 0; k < n_ues; ++k)
                        if (metf[k] >= lim && qq[k] > 0) {
                            const double m64 = (double)rate[k] / th[k];
                            if (m64 > best64) { best64 = m64; idx = k; }
                        }
                }
                rbs[idx] += c;
                const int cap = c * rate[idx];
                const int tx = (long long)cap < qq[idx] ? cap : (int)qq[idx];
                qq[idx] -= tx;
                bits[idx] += tx;
                th[idx] = __dadd_rn(__dmul_rn(PF_A, th[idx]), b_bits_over_slot(bits[idx]));
                if (qq[idx] > 0) metf[idx] = (float)rate[idx] * rcp_approx((float)th[idx]);
                else { metf[idx] = 0.0f; --n_backlog; }
                r += 2;
            }
            __syncwarp(sched_mask);
            // phase 2: a single backlogged UE takes chunks until it is drained or the PRBs run out (closed form)
            if (r < n_prbs && n_backlog == 1) {
                int j = 0;
                while (qq[j] <= 0) ++j;
                const int left = n_prbs - r, full = left >> 1;
                const int cap2 = 2 * rate[j];
                if (qq[j] <= (long long)full * cap2) {                      // drained within the 2-PRB chunks
                    const int q32 = (int)qq[j];
                    const int need = (q32 + cap2 - 1) / cap2;
                    rbs[j] += 2 * need; bits[j] += q32; qq[j] = 0; r += 2 * need;
                } else {
                    int tx = full * cap2;
                    rbs[j] += 2 * full; qq[j] -= tx; bits[j] += tx;
                    if (left & 1) {                                         // last, single-PRB chunk
                        tx = (int)min((long long)rate[j], qq[j]);
                        rbs[j] += 1; qq[j] -= tx; bits[j] += tx;
                    }
                    r = n_prbs;
                }
            }
            // phase 3: every queue drained -> all metrics 0 -> argmax 0 with 0 bits for the remaining PRBs
            if (r < n_prbs) rbs[0] += n_prbs - r;
            // ---- MI sums of the served sub-bands, flattened over (UE, quad): ~n_prbs/4 iterations per lane
            __syncwarp(sched_mask);
            {
                int k = -1, left = 0, q = 0, lo = 0, hi = 0, o = row_base;
                float c0 = 0.f, c1 = 0.f, nf = 0.f;
                double msum = 0.0;
                const int4 *col4 = nullptr;
                for (; !(RS_EXP & 4);) {
                    if (left == 0) {
                        if (k >= 0) mavg[k] = (float)(msum / (double)rbs[k]);
                        do { ++k; if (k < n_ues) { lo = o; o += rbs[k]; } } while (k < n_ues && rbs[k] < 2);
                        if (k >= n_ues) break;
                        hi = lo + rbs[k];
                        q = lo >> 2;
                        left = ((hi - 1) >> 2) - q + 1;
                        const int m = s_mod[mcs[k]];
                        const float kf = (float)c_MI_K[m], x0f = (float)c_MI_X0[m];
                        c1 = -kf * LOG2E_F; c0 = kf * x0f * LOG2E_F; nf = nomf[k];
                        col4 = reinterpret_cast<const int4 *>(tb.trace_fix + coloff[k]);
                        msum = 0.0;
                    }
                    int qq4 = q;
                    while (qq4 >= QUADS_PER_COL) qq4 -= QUADS_PER_COL;
                    const int4 v = LDQ_D(col4 + qq4);
                    const int b = q << 2;
                    float part = 0.f;
                    {
                        const float e0 = ex2_approx(__fmaf_rn(__fmaf_rn((float)v.x, FIX_SCALE, nf), c1, c0));
                        const float e1 = ex2_approx(__fmaf_rn(__fmaf_rn((float)v.y, FIX_SCALE, nf), c1, c0));
                        const float e2 = ex2_approx(__fmaf_rn(__fmaf_rn((float)v.z, FIX_SCALE, nf), c1, c0));
                        const float e3 = ex2_approx(__fmaf_rn(__fmaf_rn((float)v.w, FIX_SCALE, nf), c1, c0));
                        const float m0 = rcp_approx(1.0f + e0), m1 = rcp_approx(1.0f + e1);
                        const float m2 = rcp_approx(1.0f + e2), m3 = rcp_approx(1.0f + e3);
                        part += (b + 0 >= lo && b + 0 < hi) ? m0 : 0.f;
                        part += (b + 1 >= lo && b + 1 < hi) ? m1 : 0.f;
                        part += (b + 2 >= lo && b + 2 < hi) ? m2 : 0.f;
                        part += (b + 3 >= lo && b + 3 < hi) ? m3 : 0.f;
                    }
                    msum += (double)part;
                    ++q; --left;
                }
            }
            // ---- per-UE reception (schedulers.py:66-76, slice_l1.py:219-224) + transmission_step (slice_ran.py:51-55)
            __syncwarp(sched_mask);
            int o = 0;
            for (int k = 0; k < n_ues; ++k) {
                const int prbs = rbs[k];
                int b = bits[k];
                UeRec *g = ue + k;
                if (prbs) {
                    const int row0 = (row_base + o) % TRACE_ROWS;
                    const double u01 = r_rx.u01();
                    trace_elems += (unsigned)prbs;
                    bool received = false, need_exact = false;
                    float dbg_p32 = -1.f, dbg_eps = 0.f;
                    if (prbs == 1) need_exact = true;                        // single RB: no MI averaging, one fp64 sigmoid
                    else {
                        const float m = mavg[k];
                        if (m >= 1.0f - 1e-4f) received = true;              // p == 1.0 exactly in fp64
                        else if (m <= 1e-4f) received = false;               // p < 2^-53
                        else {
                            const int md = s_mod[mcs[k]];
                            const float kf = (float)c_MI_K[md], x0f = (float)c_MI_X0[md];
                            const float rr = rcp_approx(m) - 1.0f;
                            const float seff = x0f - __logf(rr) / kf;       // inv_sigmoid, channel_models.py:39-41
                            const float L = Af * (seff - s_ref[mcs[k]]) - Bf;
                            const float p32 = rcp_approx(1.0f + __expf(-L));
                            const float epsL = 2e-5f / (kf * m * (1.0f - m)) + 4e-5f;   // 8 dm / (k m (1-m)), dm <= 2.5e-6
                            const float eps = 1.1f * p32 * (1.0f - p32) * epsL + 5e-7f;
                            const double d = u01 - (double)p32;
                            received = d < 0.0;
                            need_exact = fabs(d) <= (double)eps;
                            dbg_p32 = p32; dbg_eps = eps;
                        }
                    }
                    if (need_exact || p.debug_check) {
                        const double pr = response_exact(tb, mcs[k], (size_t)coloff[k], row0, prbs, g->nominal);
                        const bool exact = u01 < pr;
                        if (p.debug_check && !need_exact) {
                            if (dbg_p32 >= 0.f) atomic_max_float(st.dbg + 0, (float)(fabs(pr - (double)dbg_p32) / (double)dbg_eps));
                            if (exact != received) atomicAdd(reinterpret_cast<unsigned *>(st.dbg + 2), 1u);
                        }
                        if (need_exact) { received = exact; slow_rx += prbs > 1; }
                    }
                    if (!received) b = 0;
                } else b = 0;
                o += prbs;
                const long long q = qq[k] + bits[k] - b;         // ue.queue - received bits (qq already had bits[k] taken out)
                g->queue = q;
                g->th = __dadd_rn(__dmul_rn(PF_A, g->th), b_bits_over_slot(b));
                g->bits = b;
                pe_l[k] = (pe_l[k] & 0xFFFF0000) | prbs;
                g->pe = pe_l[k];
                bits[k] = b;
                qq[k] = q;
            }
        }
        // ================= update_info (slice_ran.py:278-305), from the per-TTI scratch
        __syncwarp(warp_mask);
        {
            long long q[2] = {0, 0};
            int sn[2] = {0, 0}, n[2] = {0, 0};
            for (int k = 0; k < n_ues; ++k) {
                const int ty = (int)((types >> k) & 1u);
                a_traffic[ty] += new_bits[k];
                a_th[ty] += bits[k];
                a_prb[ty] += pe_l[k] & 0xFFFF;
                q[ty] += qq[k];
                sn[ty] += pe_l[k] >> 16;
                n[ty] += 1;
            }
#pragma unroll
            for (int ty = 0; ty < 2; ++ty) {
                const double nn = (double)max(n[ty], 1);
                a_queue[ty] += (double)q[ty] / nn;
                a_snr[ty] += (double)sn[ty] / nn;
            }
        }
    }

    __syncwarp(warp_mask);
    // ---- persist slice scalars
    hdr.n_ues = n_ues; hdr.cbr_next = cbr_next; hdr.vbr_next = vbr_next; hdr.clock = clock;
    hdr.ctr[0] = c_ran; hdr.ctr[1] = r_chan.n; hdr.ctr[2] = r_rx.n; hdr.ctr[3] = r_vbr.n;
    st.hdr[u] = hdr;
    st.hint[u] = (pf_iters << 8) | (uint32_t)n_prbs;

    // ---- end of observation period: state, SLA label (slice_ran.py:307-325, slice_l1.py:160-171)
    const double acc[10] = {(double)a_traffic[0], (double)a_th[0], (double)a_prb[0], a_queue[0], a_snr[0],
                            (double)a_traffic[1], (double)a_th[1], (double)a_prb[1], a_queue[1], a_snr[1]};
    finish_embb_unit(p, st, env, s, u, acc, flags);
    if (trace_elems) atomicAdd(p.trace_elems, (unsigned long long)trace_elems);
    if (slow_snr) atomicAdd(p.slow_paths + 0, (unsigned long long)slow_snr);
    if (slow_rx) atomicAdd(p.slow_paths + 1, (unsigned long long)slow_rx);
}

// NodeB.reset for the eMBB units (slice_l1.py:145-148, slice_ran.py:182-190): UEs and timers cleared,
// Philox counters and the unit clock keep running (the reference never reseeds on reset).
__global__ void __launch_bounds__(256) embb_reset_kernel(const __grid_constant__ EmbbState st) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= st.U) return;
    UnitHdr h = st.hdr[u];
    h.n_ues = 0; h.cbr_next = 0; h.vbr_next = 0;
    st.hdr[u] = h;
    st.hint[u] = 0;
    for (int j = 0; j < 10; ++j) st.acc[(size_t)u * 10 + j] = 0.0;
}
void launch_embb_reset(const EmbbState &st, cudaStream_t stream) {
    embb_reset_kernel<<<(st.U + 255) / 256, 256, 0, stream>>>(st);
}

void launch_embb_sort(const StepParams &p, const EmbbState &st, int max_front_ues, int heavy_min_ues, cudaStream_t stream) {
    cudaMemsetAsync(st.hist, 0, (2 * KEY_BINS + 4 + SCAN_BLOCKS) * sizeof(uint32_t), stream);
    if (st.heavy_thr > 0) cudaMemsetAsync(st.wlist + st.U, 0, sizeof(int32_t), stream);
    if (st.dil) cudaMemsetAsync(st.perm, 0xFF, (size_t)st.perm_len * sizeof(int32_t), stream);   // idle lanes of the diluted list
    window_kernel<<<(p.N + 255) / 256, 256, 0, stream>>>(p, st, max_front_ues, heavy_min_ues);
    scan_local_kernel<<<SCAN_BLOCKS, SCAN_THREADS, 0, stream>>>(st);
    scan_offset_kernel<<<SCAN_BLOCKS, SCAN_THREADS, 0, stream>>>(st);
    scatter_kernel<<<(st.U + 255) / 256, 256, 0, stream>>>(p, st, max_front_ues, heavy_min_ues);
}

// general kernel over the front list (back_list = 0) or over list L (back_list = 1)
void launch_embb_general(const StepParams &p, const EmbbState &st, const Tables &tb, int back_list, cudaStream_t stream) {
    const int threads = 128, blocks = (st.U + threads - 1) / threads;
    if (st.K <= 16) embb_step_fast<16><<<blocks, threads, 0, stream>>>(p, st, tb, back_list);
    else embb_step_fast<32><<<blocks, threads, 0, stream>>>(p, st, tb, back_list);
}

// variant 2: every unit through the general kernel, sorted
int launch_embb_fast(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream) {
    launch_embb_sort(p, st, 1 << 30, 1 << 30, stream);
    launch_embb_general(p, st, tb, 0, stream);
    return 5;   // kernels launched
}

}  // namespace rs
