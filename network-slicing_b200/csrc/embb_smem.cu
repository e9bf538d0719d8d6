// K1 default variant: the unit's UE table stays in SHARED MEMORY for the whole observation period.
//
// Same algorithm, same guarded fast math and same sorted thread->unit mapping as embb_fast.cu
// (which remains the general kernel).  What changes is where the state lives during the 50 TTIs:
// the general kernel re-reads and re-writes every 64-byte UE record in global memory each TTI and
// keeps its per-TTI PF scratch in local memory; both go through L1, which is far too small for
// 512 threads x (records + scratch + fading-trace lines) and is write-through (measured: 10.7 GB of
// L2 writes per launch, ~2000 cycles average wait per L1-missing load).  Here each thread loads its
// unit's records once, keeps them as conflict-free SoA words smem[(slot, word)][thread], runs the 50
// TTIs out of shared memory, and writes the records back once.  L1 then only serves the trace lines.
//
// Capacity: KS = 8 slots of 13 hot words per lane (52 KB per 128-thread block, 4 blocks per SM).  The cold part of
// a UE (departure time, VBR countdowns) is only touched on events -- VBR traffic is evaluated lazily: between
// events a source contributes 1000 bits per active burst -- and lives in a per-step global scratch copy, so an
// aborted unit leaves the records untouched.  A unit with
// up to 6 live UEs at the start of the step owns one lane; a unit with 7..14 owns a PAIR of lanes (the odd
// lane idles and lends its 8 slots: slot k >= 8 lives in the neighbour's column), placed at the head of the
// sorted list by the pre-pass.  Only units with more than 14 UEs go to the general kernel (list L).  A unit
// that outgrows its slots during the step (or whose queue no longer fits 31 bits) ABORTS before writing
// anything and is appended to list L, i.e. it is replayed from its untouched state by the general
// kernel.  Results are therefore independent of the routing.
#include "embb_device.cuh"
#include "embb_fastmath.cuh"

namespace rs {

constexpr int SM_THREADS = 128;
constexpr int SM_KS = 8;          // UE slots per thread held in shared memory
constexpr int SM_WORDS = 13;      // 32-bit words per slot
constexpr int SM_MAX_START_UES = SM_KS - 2;            // single-lane units
constexpr int SM_MAX_START_UES_PAIR = 2 * SM_KS - 2;   // units that own a pair of lanes
constexpr int QUEUE_LIMIT = 1 << 30;

// word offsets inside a slot (64-bit fields first so that they are 8-byte aligned in their own planes)
struct SmemView {
    double *th;        // [KS][T] ue.th
    double *thpf;      // [KS][T] scheduler's working copy of max(th, 1)   (schedulers.py:40)
    double *nominal;   // [KS][T]
    int *queue;        // [KS][T] ue.queue (bits, < 2^30 or the unit aborts)
    uint32_t *meta;    // [KS][T]
    uint32_t *vev;     // [KS][T] VBR: slots until this source's next event (burst end / burst arrival), VEV_NONE if none
    int *bits;         // [KS][T] ue.bits; doubles as the scheduler's ue_bits while a TTI is scheduled
    int *pe;           // [KS][T] ue.prbs (bits 0-7) | active bursts (bits 8-11) | ue.e_snr << 16; prbs doubles as ue_rbs
    uint32_t *rm;      // [KS][T] rate (low 16) | mcs << 16 of this TTI
    float *metf;       // [KS][T] fp32 PF metric rate / th of a backlogged UE, 0 otherwise (schedulers.py:52)
};

__device__ __forceinline__ SmemView carve_smem(unsigned char *base) {
    SmemView v;
    constexpr int P = SM_KS * SM_THREADS;      // elements per plane
    v.th = reinterpret_cast<double *>(base);
    v.thpf = v.th + P;
    v.nominal = v.thpf + P;
    v.queue = reinterpret_cast<int *>(v.nominal + P);
    v.meta = reinterpret_cast<uint32_t *>(v.queue + P);
    v.vev = v.meta + P;
    v.bits = reinterpret_cast<int *>(v.vev + P);
    v.pe = v.bits + P;
    v.rm = reinterpret_cast<uint32_t *>(v.pe + P);
    v.metf = reinterpret_cast<float *>(v.rm + P);
    return v;
}
static_assert(SM_WORDS == 3 * 2 + 7, "slot word budget");

// element of slot k in a [KS][T] plane: slots 8.. of a pair-owning unit live in the neighbouring (idle) lane's column
#define SIX(k) ((((k) & (SM_KS - 1)) * SM_THREADS) + tid + ((k) >> 3))
static_assert(SM_KS == 8, "SIX assumes 8 slots per lane");

constexpr uint32_t VEV_NONE = 0xFFFFu;

// next VBR event of a source: the smallest positive countdown among its bursts and its next arrival
__device__ __forceinline__ uint32_t next_vbr_event(const ColdRec &c) {
    int ev = 0x7FFFFFFF;
#pragma unroll
    for (int j = 0; j < MAX_BURSTS; ++j) if (c.togo[j] > 0) ev = min(ev, (int)c.togo[j]);
    if (c.vnext > 0) ev = min(ev, c.vnext);
    return ev == 0x7FFFFFFF ? VEV_NONE : (uint32_t)ev;
}
__device__ __forceinline__ int active_bursts(const ColdRec &c) {
    int nb = 0;
#pragma unroll
    for (int j = 0; j < MAX_BURSTS; ++j) nb += c.togo[j] != 0;
    return nb;
}
__device__ __forceinline__ void load_cold(const ColdRec *g, ColdRec &c) {
    const int4 *s = reinterpret_cast<const int4 *>(g);
    int4 *d = reinterpret_cast<int4 *>(&c);
    d[0] = s[0]; d[1] = s[1];
}
__device__ __forceinline__ void store_cold(ColdRec *g, const ColdRec &c) {
    const int4 *s = reinterpret_cast<const int4 *>(&c);
    int4 *d = reinterpret_cast<int4 *>(g);
    d[0] = s[0]; d[1] = s[1];
}

__device__ __forceinline__ void smem_move_slot(const SmemView &v, ColdRec *cold, int tid, int from, int to) {
    v.th[SIX(to)] = v.th[SIX(from)]; v.nominal[SIX(to)] = v.nominal[SIX(from)];
    v.queue[SIX(to)] = v.queue[SIX(from)]; v.meta[SIX(to)] = v.meta[SIX(from)]; v.vev[SIX(to)] = v.vev[SIX(from)];
    v.bits[SIX(to)] = v.bits[SIX(from)]; v.pe[SIX(to)] = v.pe[SIX(from)];
    ColdRec c;
    load_cold(cold + from, c);
    store_cold(cold + to, c);
}

// Rare RAN events of a slot on the shared-memory table; same order as ran_events in embb_fast.cu
// (slice_ran.py:263-268, slice_l1.py:196-198).  Sets c.flags bit 31 when the unit must abort.
__device__ __noinline__ void ran_events_smem(const StepParams &p, const SmemView &v, ColdRec *cold, int tid, uint32_t k0, uint32_t k1, uint32_t genv,
                                             uint32_t s, int t, uint32_t clock, int a_prb0, int a_th0, int slots_cap, RanCtx &c) {
    struct { PhiloxStream ran, chan, vbr; } rng{{k0, k1, s, STREAM_RAN, c.c_ran, genv}, {k0, k1, s, STREAM_CHAN, c.c_chan, genv},
                                                {k0, k1, s, STREAM_VBR, c.c_vbr, genv}};
    int n_ues = c.n_ues, cbr_next = c.cbr_next, vbr_next = c.vbr_next;
    uint32_t next_dep = c.next_dep, flags = c.flags;
    int arr_type[2], arr_rem[2], arr_vnext[2], n_arr = 0;
    if (cbr_next == 0) {                                                          // slice_ran.py:205-227
        cbr_next = exp_slots_ms(rng.ran, 1.0 / (2.0 / 60.0));
        const double cbr_prb = (double)a_prb0 / (double)t;                        // cbr_cac, :195-203
        const double cbr_th = (double)a_th0 / ((double)t * 1e-3);
        if (!(cbr_prb >= 20.0 || cbr_th >= 10e6)) {
            arr_type[n_arr] = 0; arr_vnext[n_arr] = 0;
            arr_rem[n_arr++] = exp_slots_ms(rng.ran, 30.0);
        }
    } else cbr_next -= 1;
    if (vbr_next == 0) {                                                          // :229-249
        arr_type[n_arr] = 1;
        arr_vnext[n_arr] = exp_slots(rng.vbr, (1.0 / 1) / 1e-3);                  // VbrSource.__init__, traffic_generators.py:65-66
        arr_rem[n_arr++] = exp_slots_ms(rng.ran, 30.0);
        vbr_next = exp_slots_ms(rng.ran, 1.0 / (5.0 / 60.0));
    } else vbr_next -= 1;
    if (clock == next_dep) {                                                      // departures, :251-261 (order kept)
        int w = 0;
        uint32_t nd = DEP_NEVER;
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
        for (int k = 0; k < n_ues; ++k) {
            const uint32_t d = cold[k].dep_at;
            if (d != clock) {
                if (w != k) smem_move_slot(v, cold, tid, k, w);
                nd = min(nd, d);
                ++w;
            }
        }
        n_ues = w;
        next_dep = nd;
    }
    for (int a = 0; a < n_arr; ++a) {                                             // slice_l1.py:183-186
        const int rem = arr_rem[a] - 1;                          // this slot's departures() already ticked it
        if (rem == 0) { flags |= 8u; continue; }
        if (n_ues >= slots_cap) { flags |= 0x80000000u; break; } // out of slots: replay in the general kernel
        const int fading = (int)rng.chan.integers(3);                             // channel_models.py:163-169
        const int index = (int)rng.chan.integers(N_SAMPLES);
        const int step = rng.chan.integers(2) ? 1 : -1;
        const int k = n_ues;
        v.nominal[SIX(k)] = draw_nominal_sinr(rng.chan, p.prop_A, p.prop_B);
        v.meta[SIX(k)] = pack_meta(arr_type[a], fading, step, index);
        const uint32_t dep_at = arr_rem[a] == 0 ? DEP_NEVER : clock + (uint32_t)rem;
        ColdRec c;
#pragma unroll
        for (int j = 0; j < MAX_BURSTS; ++j) c.togo[j] = 0;
        c.dep_at = dep_at; c.vnext = arr_vnext[a];
        c.sync = clock - 1u;                                     // its first traffic step happens in this very slot
        c.pad = 0u;
        store_cold(cold + k, c);
        v.vev[SIX(k)] = arr_vnext[a] > 0 ? (uint32_t)arr_vnext[a] : VEV_NONE;
        v.bits[SIX(k)] = 0; v.th[SIX(k)] = 0.0; v.queue[SIX(k)] = 0; v.pe[SIX(k)] = 0;
        next_dep = min(next_dep, dep_at);
        ++n_ues;
    }
    c.c_ran = rng.ran.n; c.c_chan = rng.chan.n; c.c_vbr = rng.vbr.n;
    c.n_ues = n_ues; c.cbr_next = cbr_next; c.vbr_next = vbr_next; c.next_dep = next_dep; c.flags = flags;
}

// A VBR source event (VbrSource.step, traffic_generators.py:70-99, evaluated lazily): bring the source's
// countdowns up to `clock`, retire the bursts that end now (they contribute nothing this slot), then the burst
// arrival if it is due (the new burst contributes from the next slot).  nb_now = bursts contributing THIS slot,
// nb = bursts alive afterwards.  Returns the slots until the next event.
__device__ __noinline__ uint32_t vbr_event(ColdRec *g, uint32_t clock, PhiloxStream &r_vbr, uint32_t &flags, int &nb_now, int &nb) {
    ColdRec c;
    load_cold(g, c);
    const int elapsed = (int)(clock - c.sync);
    c.sync = clock;
    int alive = 0;
#pragma unroll
    for (int j = 0; j < MAX_BURSTS; ++j) {
        int tg = c.togo[j];
        if (tg > 0) { tg -= elapsed; c.togo[j] = (int16_t)tg; }   // a burst whose length was drawn as 0 stays at -1: it never ends
        alive += tg != 0;
    }
    nb_now = alive;
    if (c.vnext > 0) {
        c.vnext -= elapsed;
        if (c.vnext == 0) {                                      // traffic_generators.py:92-97
            int len = exp_slots(r_vbr, 500.0);
            if (len == 0) len = -1;
            bool placed = false;
#pragma unroll
            for (int j = 0; j < MAX_BURSTS; ++j)
                if (!placed && c.togo[j] == 0) { c.togo[j] = (int16_t)len; placed = true; }
            if (placed) ++alive; else flags |= 2u;
            c.vnext = exp_slots(r_vbr, 1000.0);                  // a draw of 0 never fires again
        }
    }
    nb = alive;
    store_cold(g, c);
    return next_vbr_event(c);
}

// PF argmax among the candidates within 1e-6 (relative) of the fp32 maximum: exact fp64 quotients, first maximum
// (np.argmax, schedulers.py:52).  Taken by 0.6 % of the chunks; out of line to keep the RB loop compact.
__device__ __noinline__ int pf_exact_argmax(const SmemView &v, int tid, int n_ues, float best) {
    const float lim = best * (1.0f - 1e-6f);
    double best64 = -1.0;
    int idx = 0;
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
    for (int k = 0; k < n_ues; ++k) {
        const float m = v.metf[SIX(k)];
        if (m >= lim && m > 0.0f) {
            const double m64 = (double)(v.rm[SIX(k)] & 0xFFFFu) / v.thpf[SIX(k)];
            if (m64 > best64) { best64 = m64; idx = k; }
        }
    }
    return idx;
}

#ifdef RS_STATS
// Workload statistics of the PF loop (experiment builds only; read with rs_debug_stats)
__device__ unsigned long long g_stats[64];
#define STAT(i, v) atomicAdd(&g_stats[i], (unsigned long long)(v))
__device__ __forceinline__ int log2_bucket(int x) { return x <= 0 ? 0 : min(1 + (31 - __clz(x)), 7); }
#else
#define STAT(i, v)
#endif

#ifndef RS_SM_BLOCKS
#define RS_SM_BLOCKS 4
#endif
#ifndef RS_WIDE_BLOCKS
#define RS_WIDE_BLOCKS 2
#endif
// WIDE = 1: latency variant for batches that cannot fill the GPU (more registers per thread, deeper load batches)
template <int WIDE>
__global__ void __launch_bounds__(SM_THREADS, WIDE ? RS_WIDE_BLOCKS : RS_SM_BLOCKS) embb_step_smem(const __grid_constant__ StepParams p,
                                                                const __grid_constant__ EmbbState st,
                                                                const __grid_constant__ Tables tb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int16_t s_rate[256];
    __shared__ int8_t s_mcs[256];
    __shared__ float s_ref[26];
    __shared__ int8_t s_mod[26];
    __shared__ float s_mi[3][4];                                 // per modulation: k, x0, c1 = -k log2(e), c0 = k x0 log2(e)  (fp32 MI fast path)
    __shared__ float s_inv[2 * TRACE_ROWS + 1];                  // 1 / n for the MI mean (fp32 fast path, inside the eps budget)
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += SM_THREADS) { s_rate[i] = tb.lut_rate[i]; s_mcs[i] = tb.lut_mcs[i]; }
    if (tid < 26) { s_ref[tid] = (float)tb.snr_ref[tid]; s_mod[tid] = tb.mod[tid]; }
    if (tid < 3) {
        const float kf = (float)c_MI_K[tid], x0f = (float)c_MI_X0[tid];
        s_mi[tid][0] = kf; s_mi[tid][1] = x0f; s_mi[tid][2] = -kf * LOG2E_F; s_mi[tid][3] = kf * x0f * LOG2E_F;
    }
    for (int i = tid; i <= 2 * TRACE_ROWS; i += SM_THREADS) s_inv[i] = i ? __frcp_rn((float)i) : 0.f;
    __syncthreads();
    const SmemView v = carve_smem(smem_raw);

    const int tix = blockIdx.x * SM_THREADS + tid;
    const int count = (int)st.hist[2 * SORT_BINS + 0] << st.dil;
    const unsigned warp_mask = __ballot_sync(0xffffffffu, tix < count);
    if (tix >= count) return;
    const int u_raw = st.perm[tix];
    const bool pad = u_raw < 0;                                  // idle lane lending its slots to the unit on its left
    const int slots_cap = min(st.K, tix < (2 * (int)st.hist[2 * SORT_BINS + 3]) << st.dil ? st.route[3] : st.route[1]);
    const int u = pad ? 0 : u_raw;
    const int env = u / p.n_embb, s = u - env * p.n_embb;
    int i_prb, n_prbs;
    unpack_window(st.win[u], i_prb, n_prbs);
    const int row_base = i_prb % TRACE_ROWS;
    uint32_t flags = 0;

    UnitHdr hdr = st.hdr[u];
    UeRec *ue = st.ue + (size_t)u * st.K;
    ColdRec *cold = st.cold + (size_t)u * st.K;                  // per-step scratch: committed only if the unit does not abort
    const uint64_t seed = p.seed0;
    const uint32_t genv = p.env0 + (uint32_t)env;   // global env id: Philox counter word 3
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c_ran = hdr.ctr[0];
    PhiloxStream r_chan{k0, k1, (uint32_t)s, STREAM_CHAN, hdr.ctr[1], genv}, r_rx{k0, k1, (uint32_t)s, STREAM_L1RX, hdr.ctr[2], genv},
        r_vbr{k0, k1, (uint32_t)s, STREAM_VBR, hdr.ctr[3], genv};

    int n_ues = pad ? 0 : hdr.n_ues, cbr_next = hdr.cbr_next, vbr_next = hdr.vbr_next;
    uint32_t clock = hdr.clock, next_dep = DEP_NEVER;
    bool dead = pad;                                             // aborted (replayed by the general kernel) or idle pad lane

    // ---- gather the unit's records into shared memory (once per step)
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
    for (int k = 0; k < n_ues; ++k) {
        UeRec r;
        load_rec(ue + k, r);
        v.th[SIX(k)] = r.th; v.nominal[SIX(k)] = r.nominal;
        v.queue[SIX(k)] = (int)r.queue; v.meta[SIX(k)] = r.meta; v.bits[SIX(k)] = r.bits;
        ColdRec c;
#pragma unroll
        for (int j = 0; j < MAX_BURSTS; ++j) c.togo[j] = r.togo[j];
        c.dep_at = r.dep_at; c.vnext = r.vnext; c.sync = clock; c.pad = 0u;
        store_cold(cold + k, c);
        v.vev[SIX(k)] = (r.meta & 1u) ? next_vbr_event(c) : VEV_NONE;
        v.pe[SIX(k)] = (r.pe & (int)0xFFFF00FF) | (active_bursts(c) << 8);
        next_dep = min(next_dep, r.dep_at);
        dead |= r.queue >= QUEUE_LIMIT;
    }
    if (dead) n_ues = 0;

    // slice_ran.py:270-273 reset_info.  Per-type accumulators are kept as (both types, VBR only) scalar pairs: an array
    // indexed by the UE type would live in local memory (measured: 21 % of the stall samples of this kernel).
    int a_traffic_all = 0, a_traffic_v = 0, a_th_all = 0, a_th_v = 0, a_prb_all = 0, a_prb_v = 0;
    double a_queue_c = 0.0, a_queue_v = 0.0, a_snr_c = 0.0, a_snr_v = 0.0;
    unsigned trace_elems = 0, slow_snr = 0, slow_rx = 0, pf_iters = 0;
    const float Af = (float)tb.A, Bf = (float)tb.B;
#ifdef RS_WSUM_QUADS
    const double inv_n = n_prbs > 0 ? 1.0 / ((double)n_prbs * FIX_ONE) : 0.0;
#else
    const double inv_n = n_prbs > 0 ? tb.pre_inv / (double)n_prbs : 0.0;
#endif

    for (int t = 1; t <= p.slots; ++t) {          // slot_counter == t (zeroed by reset_info each step)
        __syncwarp(warp_mask);
        ++clock;
        // ================= slice_ran.slot(): arrivals / departures only on event slots
        if (!dead) {
            if (cbr_next == 0 || vbr_next == 0 || clock == next_dep) {
                RanCtx c{c_ran, r_chan.n, r_vbr.n, next_dep, flags, n_ues, cbr_next, vbr_next};
                ran_events_smem(p, v, cold, tid, k0, k1, genv, (uint32_t)s, t, clock, a_prb_all - a_prb_v, a_th_all - a_th_v, slots_cap, c);
                c_ran = c.c_ran; r_chan.n = c.c_chan; r_vbr.n = c.c_vbr; next_dep = c.next_dep; flags = c.flags;
                n_ues = c.n_ues; cbr_next = c.cbr_next; vbr_next = c.vbr_next;
                if (flags & 0x80000000u) { dead = true; n_ues = 0; }
            } else { cbr_next -= 1; vbr_next -= 1; }
        }

        // ================= per-UE traffic + SNR estimate (slice_l1.py:200-213)
        __syncwarp(warp_mask);
        int n_backlog = 0;
        int sn_all = 0, sn_v = 0, cnt_all = 0, cnt_v = 0;
        long long qsum_all = 0, qsum_v = 0;
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
        for (int k = 0; k < n_ues; ++k) {
            uint32_t meta = v.meta[SIX(k)];
            const int ty = (int)(meta & 1u);
            int pe = v.pe[SIX(k)];
            int nb_bits = 500;                                   // CbrSource: 500000 b/s * 1e-3 every slot
            if (ty) {                                            // VbrSource: 1000 bits per active burst; countdowns only on events
                int nb_now = (pe >> 8) & 0xF;
                uint32_t ev = v.vev[SIX(k)];
                if (ev != VEV_NONE) {
                    if (--ev == 0u) {
                        int nb = nb_now;
                        ev = vbr_event(cold + k, clock, r_vbr, flags, nb_now, nb);
                        pe = (pe & (int)0xFFFFF0FF) | (nb << 8);
                    }
                    v.vev[SIX(k)] = ev;
                }
                nb_bits = 1000 * nb_now;
            }
            a_traffic_all += nb_bits; a_traffic_v += ty ? nb_bits : 0;
            const int queue = v.queue[SIX(k)] + nb_bits;
            v.queue[SIX(k)] = queue;
            if (queue >= QUEUE_LIMIT) dead = true;
            if (n_prbs > 0) {
                int index = (int)(meta >> 4), step = (meta & 8u) ? 1 : -1;
                const int fading = (int)((meta >> 1) & 3u);
                walk_trace(r_chan, index, step);                 // channel_models.py:171-191
                meta = pack_meta(ty, fading, step, index);
                v.meta[SIX(k)] = meta;
                const int col_off = (fading * N_SAMPLES + index) * TRACE_ROWS;
                trace_elems += (unsigned)n_prbs;
                const double nominal = v.nominal[SIX(k)];
#ifdef RS_WSUM_QUADS
                const long long isum = window_sum_fix<WIDE>(tb.trace_fix + col_off, row_base, n_prbs);
                double mean = (double)isum * inv_n + nominal;    // |mean - reference mean| < 2^-25 + few ulp
                const double guard = SNR_ROUND_GUARD;
#else
                const int isum = window_sum_prefix(tb.trace_pre + (fading * N_SAMPLES + index) * PRE_STRIDE, row_base, n_prbs);
                double mean = (double)isum * inv_n + nominal;    // |mean - reference mean| <= 2^-(pre_bits + 1) + few ulp
                const double guard = tb.pre_guard;
#endif
                const double fr = mean - floor(mean);
                const bool near = fabs(fr - 0.5) < guard;        // within the guard of a rounding boundary
                if (near || p.debug_check) {
                    const double exact = window_mean_fp64(tb.trace + col_off, row_base, n_prbs, nominal);
                    if (p.debug_check) atomic_max_float(st.dbg + 1, (float)(fabs(exact - mean) / guard));
                    if (near) { mean = exact; ++slow_snr; }
                }
                const int e_snr = __double2int_rn(mean);         // round(np.mean(snr)), slice_ran.py:43-45
                pe = (pe & 0xFFFF) | (e_snr << 16);
            }
            v.pe[SIX(k)] = pe;
            // scheduler inputs (schedulers.py:37-45) + the update_info terms that are already final
            const int e = min(max(pe >> 16, -128), 127) + 128;
            v.rm[SIX(k)] = (uint32_t)(uint16_t)s_rate[e] | ((uint32_t)(uint8_t)s_mcs[e] << 16);
            const double th = v.th[SIX(k)];
            const double thp = th > 1.0 ? th : 1.0;
            v.thpf[SIX(k)] = thp;
            v.metf[SIX(k)] = queue > 0 ? (float)s_rate[e] * rcp_approx((float)thp) : 0.0f;
            n_backlog += queue > 0;
            sn_all += pe >> 16; sn_v += ty ? (pe >> 16) : 0;
            cnt_all += 1; cnt_v += ty;
        }
        if (dead) n_ues = 0;
        // ================= scheduling + reception (slice_l1.py:215-224)
        __syncwarp(warp_mask);
        const bool scheduled = n_backlog > 0 && n_prbs > 0;      // queued_data > 0 <=> some queue > 0
        const unsigned sched_mask = __ballot_sync(warp_mask, scheduled);
#ifdef RS_STATS
        if (!pad && !dead) { STAT(0, 1); STAT(1, scheduled); STAT(8 + min(n_backlog, 7), 1); STAT(16 + min(n_ues, 15), 1); }
        int st_iters = 0, st_prev = -1, st_same = 0;
#endif
        if (scheduled) {
            {   // ue_bits = 0, ue_rbs = 0 (stepped slot address, see the scan below)
                int *bp = v.bits + tid, *pp = v.pe + tid;
                const int n_own = min(n_ues, SM_KS);
#pragma unroll 1
                for (int k = 0; k < n_own; ++k) { *bp = 0; *pp &= (int)0xFFFFFF00; bp += SM_THREADS; pp += SM_THREADS; }
#pragma unroll 1
                for (int k = SM_KS; k < n_ues; ++k) { v.bits[SIX(k)] = 0; v.pe[SIX(k)] &= (int)0xFFFFFF00; }
            }
            // ---- ProportionalFair.allocate RB loop (schedulers.py:47-63); queue left = queue - ue_bits
            int r = 0;
            // phase 1: contended chunks (>= 2 backlogged UEs); one uniform loop body for all lanes still in it
            while (n_backlog >= 2 && r < n_prbs) {
                ++pf_iters;
                if (RS_EXP & 8) break;
                const int c = min(n_prbs - r, 2);
                // argmax of rate * (queue > 0) / th, first maximum (np.argmax): fp32 metric, exact when close
                // (the slot address is stepped instead of recomputed: SIX(k) is 6 of the 15 instructions of a scan step;
                //  slots 8.. of a pair-owning unit sit in the neighbouring lane's column)
                int idx = 0;
                const float *mp = v.metf + tid;
                float best = *mp, second = -1.0f;
                const int n_own = min(n_ues, SM_KS);
#pragma unroll 1
                for (int k = 1; k < n_own; ++k) {
                    mp += SM_THREADS;
                    const float m = *mp;
                    if (m > best) { second = best; best = m; idx = k; }
                    else second = fmaxf(second, m);
                }
                if (n_ues > SM_KS) {
                    mp = v.metf + tid + 1 - SM_THREADS;
#pragma unroll 1
                    for (int k = SM_KS; k < n_ues; ++k) {
                        mp += SM_THREADS;
                        const float m = *mp;
                        if (m > best) { second = best; best = m; idx = k; }
                        else second = fmaxf(second, m);
                    }
                }
#ifdef RS_STATS
                if (second >= best * (1.0f - 1e-6f)) STAT(2, 1);
#endif
                if (second >= best * (1.0f - 1e-6f)) idx = pf_exact_argmax(v, tid, n_ues, best);   // too close for fp32
                const int rate = (int)(v.rm[SIX(idx)] & 0xFFFFu);
                const int left_q = v.queue[SIX(idx)] - v.bits[SIX(idx)];
                const int tx = min(c * rate, left_q);
                const int nbits = v.bits[SIX(idx)] + tx;
                v.bits[SIX(idx)] = nbits;
                v.pe[SIX(idx)] += c;
                const double thn = __dadd_rn(__dmul_rn(PF_A, v.thpf[SIX(idx)]), b_bits_over_slot(nbits));
                v.thpf[SIX(idx)] = thn;
                if (left_q - tx <= 0) { v.metf[SIX(idx)] = 0.0f; --n_backlog; }
                else v.metf[SIX(idx)] = (float)rate * rcp_approx((float)thn);
                r += 2;
#ifdef RS_STATS
                ++st_iters; st_same += idx == st_prev; st_prev = idx;
                STAT(3, n_ues);
                if (left_q - tx <= 0) STAT(4, 1);
#endif
            }
#ifdef RS_STATS
            STAT(5, st_iters); STAT(6, st_same); STAT(32 + log2_bucket(st_iters), 1); STAT(40 + log2_bucket(st_iters), st_iters);
            STAT(7, (r < n_prbs && n_backlog == 1));
#endif
            __syncwarp(sched_mask);
            // phase 2: a single backlogged UE takes chunks until it is drained or the PRBs run out (closed form)
            if (r < n_prbs && n_backlog == 1) {
                int j = 0;
                while (v.queue[SIX(j)] - v.bits[SIX(j)] <= 0) ++j;
                const int left = n_prbs - r, full = left >> 1;
                const int rate = (int)(v.rm[SIX(j)] & 0xFFFFu), cap2 = 2 * rate;
                const int q32 = v.queue[SIX(j)] - v.bits[SIX(j)];
                if ((long long)q32 <= (long long)full * cap2) {             // drained within the 2-PRB chunks
                    const int need = (q32 + cap2 - 1) / cap2;
                    v.pe[SIX(j)] += 2 * need; v.bits[SIX(j)] += q32; r += 2 * need;
                } else {
                    int tx = full * cap2;
                    if (left & 1) tx += min(rate, q32 - tx);                // last, single-PRB chunk
                    v.pe[SIX(j)] += left; v.bits[SIX(j)] += tx;
                    r = n_prbs;
                }
            }
            // phase 3: every queue drained -> all metrics 0 -> argmax 0 with 0 bits for the remaining PRBs
            if (r < n_prbs) v.pe[SIX(0)] += n_prbs - r;
            // ---- MI sums of the served sub-bands, flattened over (UE, quad): ~n_prbs/4 iterations per lane
            __syncwarp(sched_mask);
            {
                int k = -1, left = 0, q = 0, lo = 0, hi = 0, o = row_base, rbs_k = 0;
                float c0 = 0.f, c1 = 0.f, nf = 0.f;
                double msum = 0.0;
                const int4 *col4 = nullptr;
                for (; !(RS_EXP & 4);) {
                    if (left == 0) {
                        if (k >= 0) v.metf[SIX(k)] = (float)msum * s_inv[rbs_k];    // mean MI (the PF metric is dead by now)
                        do { ++k; if (k < n_ues) { rbs_k = v.pe[SIX(k)] & 0xFF; lo = o; o += rbs_k; } } while (k < n_ues && rbs_k < 2);
                        if (k >= n_ues) break;
                        hi = lo + rbs_k;
                        q = lo >> 2;
                        left = ((hi - 1) >> 2) - q + 1;
                        const uint32_t meta = v.meta[SIX(k)];
                        const int m = s_mod[v.rm[SIX(k)] >> 16];
                        c1 = s_mi[m][2]; c0 = s_mi[m][3]; nf = (float)v.nominal[SIX(k)];
                        col4 = reinterpret_cast<const int4 *>(tb.trace_fix + ((int)((meta >> 1) & 3u) * N_SAMPLES + (int)(meta >> 4)) * TRACE_ROWS);
                        msum = 0.0;
                    }
#ifndef RS_MI_V1
                    // two quads per iteration (independent load -> ex2 -> rcp chains); a second quad past the UE's last one
                    // is masked out by the `< hi` tests below (its rows are >= hi) and its load wraps inside the column
                    const int qa = wrap_quad(q), qb = qa + 1 == QUADS_PER_COL ? 0 : qa + 1;
                    const int4 x = LDQ_D(col4 + qa), y = LDQ_D(col4 + qb);
                    const int b = q << 2;
                    const float e0 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.x, FIX_SCALE, nf), c1, c0));
                    const float e1 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.y, FIX_SCALE, nf), c1, c0));
                    const float e2 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.z, FIX_SCALE, nf), c1, c0));
                    const float e3 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.w, FIX_SCALE, nf), c1, c0));
                    const float e4 = ex2_approx(__fmaf_rn(__fmaf_rn((float)y.x, FIX_SCALE, nf), c1, c0));
                    const float e5 = ex2_approx(__fmaf_rn(__fmaf_rn((float)y.y, FIX_SCALE, nf), c1, c0));
                    const float e6 = ex2_approx(__fmaf_rn(__fmaf_rn((float)y.z, FIX_SCALE, nf), c1, c0));
                    const float e7 = ex2_approx(__fmaf_rn(__fmaf_rn((float)y.w, FIX_SCALE, nf), c1, c0));
                    float part = (b + 0 >= lo && b + 0 < hi) ? rcp_approx(1.0f + e0) : 0.f;
                    part += (b + 1 >= lo && b + 1 < hi) ? rcp_approx(1.0f + e1) : 0.f;
                    part += (b + 2 >= lo && b + 2 < hi) ? rcp_approx(1.0f + e2) : 0.f;
                    part += (b + 3 >= lo && b + 3 < hi) ? rcp_approx(1.0f + e3) : 0.f;
                    msum += (double)part;
                    float part2 = (b + 4 < hi) ? rcp_approx(1.0f + e4) : 0.f;
                    part2 += (b + 5 < hi) ? rcp_approx(1.0f + e5) : 0.f;
                    part2 += (b + 6 < hi) ? rcp_approx(1.0f + e6) : 0.f;
                    part2 += (b + 7 < hi) ? rcp_approx(1.0f + e7) : 0.f;
                    msum += (double)part2;
                    q += 2; left = max(left - 2, 0);
#else
                    int qq4 = q;
                    while (qq4 >= QUADS_PER_COL) qq4 -= QUADS_PER_COL;
                    const int4 x = LDQ_D(col4 + qq4);
                    const int b = q << 2;
                    const float e0 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.x, FIX_SCALE, nf), c1, c0));
                    const float e1 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.y, FIX_SCALE, nf), c1, c0));
                    const float e2 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.z, FIX_SCALE, nf), c1, c0));
                    const float e3 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.w, FIX_SCALE, nf), c1, c0));
                    float part = (b + 0 >= lo && b + 0 < hi) ? rcp_approx(1.0f + e0) : 0.f;
                    part += (b + 1 >= lo && b + 1 < hi) ? rcp_approx(1.0f + e1) : 0.f;
                    part += (b + 2 >= lo && b + 2 < hi) ? rcp_approx(1.0f + e2) : 0.f;
                    part += (b + 3 >= lo && b + 3 < hi) ? rcp_approx(1.0f + e3) : 0.f;
                    msum += (double)part;
                    ++q; --left;
#endif
                }
            }
            // ---- per-UE reception (schedulers.py:66-76, slice_l1.py:219-224) + transmission_step (slice_ran.py:51-55)
            __syncwarp(sched_mask);
            int o = 0;
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
            for (int k = 0; k < n_ues; ++k) {
                {
                    const int pe = v.pe[SIX(k)];
                    const int prbs = pe & 0xFF;
                    int b = v.bits[SIX(k)];
                    const int ty = (int)(v.meta[SIX(k)] & 1u);
                    if (prbs) {
                        const int mcs = (int)(v.rm[SIX(k)] >> 16);
                        const double u01 = r_rx.u01();
                        trace_elems += (unsigned)prbs;
                        bool received = false, need_exact = false;
                        float dbg_p32 = -1.f, dbg_eps = 0.f;
                        if (prbs == 1) need_exact = true;                    // single RB: no MI averaging, one fp64 sigmoid
                        else {
                            const float m = v.metf[SIX(k)];
                            if (m >= 1.0f - 1e-4f) received = true;          // p == 1.0 exactly in fp64
                            else if (m <= 1e-4f) received = false;           // p < 2^-53
                            else {
                                const int md = s_mod[mcs];
                                const float kf = s_mi[md][0], x0f = s_mi[md][1];
                                const float rr = rcp_approx(m) - 1.0f;
                                const float seff = x0f - __logf(rr) / kf;   // inv_sigmoid, channel_models.py:39-41
                                const float L = Af * (seff - s_ref[mcs]) - Bf;
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
                            const uint32_t meta = v.meta[SIX(k)];
                            const size_t col_off = (size_t)((int)((meta >> 1) & 3u) * N_SAMPLES + (int)(meta >> 4)) * TRACE_ROWS;
                            const double pr = response_exact(tb, mcs, col_off, (row_base + o) % TRACE_ROWS, prbs, v.nominal[SIX(k)]);
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
                    const int queue = v.queue[SIX(k)] - b;       // max(queue - bits, 0): bits never exceed the queue
                    v.queue[SIX(k)] = queue;
                    v.th[SIX(k)] = __dadd_rn(__dmul_rn(PF_A, v.th[SIX(k)]), b_bits_over_slot(b));
                    v.bits[SIX(k)] = b;
                    a_th_all += b; a_prb_all += prbs; qsum_all += queue;     // update_info terms (slice_ran.py:278-305)
                    if (ty) { a_th_v += b; a_prb_v += prbs; qsum_v += queue; }
                }
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < n_ues; ++k) {                    // nothing touched: stale bits / prbs accumulate
                const int ty = (int)(v.meta[SIX(k)] & 1u);
                const int b = v.bits[SIX(k)], prbs = v.pe[SIX(k)] & 0xFF, queue = v.queue[SIX(k)];
                a_th_all += b; a_prb_all += prbs; qsum_all += queue;
                if (ty) { a_th_v += b; a_prb_v += prbs; qsum_v += queue; }
            }
        }
        // ================= update_info means (slice_ran.py:290-291, 304-305)
        a_queue_c += div_count((double)(qsum_all - qsum_v), cnt_all - cnt_v);
        a_snr_c += div_count((double)(sn_all - sn_v), cnt_all - cnt_v);
        a_queue_v += div_count((double)qsum_v, cnt_v);
        a_snr_v += div_count((double)sn_v, cnt_v);
    }

    __syncwarp(warp_mask);
    if (pad) return;
    if (dead) {                                                  // replay this unit in the general kernel (list L)
        st.perm[st.perm_len - 1 - (int)atomicAdd(&st.hist[2 * SORT_BINS + 1], 1u)] = u;
        return;
    }
    // ---- scatter the records back (once per step) and persist the slice scalars
#ifndef RS_NO_UNROLL2
#pragma unroll 1
#endif
    for (int k = 0; k < n_ues; ++k) {
        ColdRec c;
        load_cold(cold + k, c);
        const int elapsed = (int)(clock - c.sync);               // bring the lazy VBR countdowns up to the end of the step
        UeRec r;
        int nb = 0;
#pragma unroll
        for (int j = 0; j < MAX_BURSTS; ++j) {
            int tg = c.togo[j];
            if (tg > 0) tg -= elapsed;
            r.togo[j] = (int16_t)tg;
            nb += tg != 0;
        }
        r.vnext = c.vnext > 0 ? c.vnext - elapsed : c.vnext;
        r.meta = v.meta[SIX(k)]; r.dep_at = c.dep_at; r.bits = v.bits[SIX(k)];
        r.nominal = v.nominal[SIX(k)]; r.th = v.th[SIX(k)]; r.queue = v.queue[SIX(k)];
        r.pe = v.pe[SIX(k)] & (int)0xFFFF00FF; r.nb = nb;
        store_rec(ue + k, r);
    }
    hdr.n_ues = n_ues; hdr.cbr_next = cbr_next; hdr.vbr_next = vbr_next; hdr.clock = clock;
    hdr.ctr[0] = c_ran; hdr.ctr[1] = r_chan.n; hdr.ctr[2] = r_rx.n; hdr.ctr[3] = r_vbr.n;
    st.hdr[u] = hdr;
    st.hint[u] = (pf_iters << 8) | (uint32_t)n_prbs;

    // ---- end of observation period: state, SLA label (slice_ran.py:307-325, slice_l1.py:160-171)
    const double acc[10] = {(double)(a_traffic_all - a_traffic_v), (double)(a_th_all - a_th_v), (double)(a_prb_all - a_prb_v), a_queue_c, a_snr_c,
                            (double)a_traffic_v, (double)a_th_v, (double)a_prb_v, a_queue_v, a_snr_v};
    finish_embb_unit(p, st, env, s, u, acc, flags);
    if (trace_elems) atomicAdd(p.trace_elems, (unsigned long long)trace_elems);
    if (slow_snr) atomicAdd(p.slow_paths + 0, (unsigned long long)slow_snr);
    if (slow_rx) atomicAdd(p.slow_paths + 1, (unsigned long long)slow_rx);
}

#ifdef RS_STATS
extern "C" int rs_debug_stats(unsigned long long *out64, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out64, g_stats, sizeof(g_stats)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(g_stats, z, sizeof(z)); }
    return 0;
}
#endif

void launch_embb_sort(const StepParams &p, const EmbbState &st, int max_front_ues, int heavy_min_ues, cudaStream_t stream);
void launch_embb_general(const StepParams &p, const EmbbState &st, const Tables &tb, int back_list, cudaStream_t stream);

// default variant: shared-memory kernel over the sorted front list, general kernel over list L
void launch_embb_warp_heavy(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream);

int launch_embb_smem(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream, cudaEvent_t *prof, const HeavyFork *fork) {
    static bool configured = false;
    constexpr int smem_bytes = SM_THREADS * SM_KS * SM_WORDS * 4;
    static_assert(RS_SM_BLOCKS * (smem_bytes + 2048 + 1024) <= 227 * 1024, "RS_SM_BLOCKS blocks per SM must fit");
    if (!configured) {
        cudaFuncSetAttribute(embb_step_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        cudaFuncSetAttribute(embb_step_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
#ifdef RS_CARVEOUT
        cudaFuncSetAttribute(embb_step_smem<0>, cudaFuncAttributePreferredSharedMemoryCarveout, RS_CARVEOUT);   // experiment: percent of the 228 KB
#endif
        configured = true;
    }
    launch_embb_sort(p, st, st.route[2], st.route[0] + 1, stream);
    int launched = 6;
    if (st.heavy_thr > 0 && fork) {                               // heavy list: warp-per-unit kernel, concurrent on the side stream
        cudaEventRecord(fork->fork, stream);
        cudaStreamWaitEvent(fork->stream, fork->fork, 0);
        launch_embb_warp_heavy(p, st, tb, fork->stream);
        cudaEventRecord(fork->join, fork->stream);
        ++launched;
    }
    const int blocks = (st.perm_len + SM_THREADS - 1) / SM_THREADS;   // worst case: every unit owns a pair of lanes
    if (prof) cudaEventRecord(prof[0], stream);               // profiling: events around the dominant kernel alone
    if (st.wide) embb_step_smem<1><<<blocks, SM_THREADS, smem_bytes, stream>>>(p, st, tb);
    else embb_step_smem<0><<<blocks, SM_THREADS, smem_bytes, stream>>>(p, st, tb);
    if (prof) cudaEventRecord(prof[1], stream);
    if (st.heavy_thr > 0 && !fork) {                              // serialised (per-kernel profiling): after the dominant kernel's event
        launch_embb_warp_heavy(p, st, tb, stream);                // pair, so that it is timed with the other eMBB kernels (embb_rest)
        ++launched;
    }
    launch_embb_general(p, st, tb, 1, stream);
    if (st.heavy_thr > 0 && fork) cudaStreamWaitEvent(stream, fork->join, 0);
    return launched;   // kernels launched
}

}  // namespace rs
