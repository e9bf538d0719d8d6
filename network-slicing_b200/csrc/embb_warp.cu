// K1 warp-per-unit variant: one WARP steps one (env, eMBB slice) unit through its 50 TTIs.
//
// The default kernel (embb_smem.cu) gives every unit one lane; a step then lasts as long as its heaviest lane's serial
// work (a slice with ~190 PRBs and a deep backlog runs ~95 PF chunks and a 50-quad MI loop per TTI), which is what a
// batch too small to fill the GPU waits for: 4096 envs took 2.1 ms per step with 94 % of the lanes idle.  Here the
// lanes of a warp share ONE unit:
//   * lane k owns UE k (<= 16 live UEs): traffic source, trace walk, window mean (prefix table), MCS lookup, reception,
//     transmission_step all run in parallel over the UEs; the UE record stays in the lane's registers for the whole step;
//   * the PF argmax over the UEs of a chunk (schedulers.py:52) is two warp REDUX instructions on the fp32 metric bits
//     (+ the exact fp64 comparison when the runner-up is within 1e-6, as in the other kernels); long RB loops hand out
//     several chunks per warp-wide step (every lane speculates its own UE's next chunks; see the RB loop);
//   * the per-PRB mutual-information sum of a served UE (channel_models.py:304-310) is spread over all 32 lanes, a quad
//     of PRBs each, and tree-reduced;
//   * the end-of-slot accumulators (slice_ran.py:278-305) are integer warp reductions.
// Draw order is kept exactly: a Philox stream is (key, counter), so lane k takes the counter the sequential code would
// have reached -- base + (draws of the lanes before it) -- from a ballot; variable-length draws (trace re-draw at a trace
// end) and RAN events (arrivals / departures, ~0.1 per unit-step) are serialised: for an event slot the lanes park
// their records in a shared-memory table, lane 0 runs the same ran_events() as the general kernel on it, and the lanes
// reload.  Results are bit-identical to the other variants (tests/test_gpu_parity.py runs all of them).
//
// A warp executes ~3x the warp-instructions of a lane doing the same unit, so this variant only pays while the batch
// leaves the GPU underfilled; rs_create routes batches of up to RS_WARP_AUTO_UNITS units to it (measured crossover).
#include "embb_device.cuh"
#include "embb_fastmath.cuh"
#include "embb_ran.cuh"

namespace rs {

// Units (warps) per block.  ONE: a block's registers and shared memory go back to the SM the moment its unit is done, instead
// of waiting for the slowest of 8 units.  Measured (profiles/r02f_warp_block_geometry.txt), 8 / 4 / 2 / 1 warps per block at 16
// warps per SM: 2048 envs 1.168 / 1.124 / 1.102 / 1.051 ms/step, 4096 envs 2.043 / 1.926 / 1.862 / 1.751, multiplexed L1 at
// 16 384 envs 9.50 / 8.36 / 7.88 / 7.20.
#ifndef RS_WP_WARPS
#define RS_WP_WARPS 1
#endif
constexpr int WP_WARPS = RS_WP_WARPS;
constexpr int WP_K = 16;           // UE slots per unit (== the UE cap of the other variants)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double shfl_d(double v, int src) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_sync(FULL, lo, src); hi = __shfl_sync(FULL, hi, src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    int lo = __double2loint(v), hi = __double2hiint(v);
    lo = __shfl_xor_sync(FULL, lo, m); hi = __shfl_xor_sync(FULL, hi, m);
    return __hiloint2double(hi, lo);
}
// sum of a non-negative 64-bit quantity over the warp (per-lane values < 2^47): two 32-bit REDUX on 24-bit halves
__device__ __forceinline__ long long reduce_add_ll(long long v) {
    const unsigned lo = (unsigned)(v & 0xFFFFFF), hi = (unsigned)(v >> 24);
    return ((long long)__reduce_add_sync(FULL, hi) << 24) + (long long)__reduce_add_sync(FULL, lo);
}

struct WarpCtx { uint32_t c_ran, c_chan, c_vbr, next_dep, flags; int n_ues, cbr_next, vbr_next; };

// ---- multiplexed L1 (create_env(L1_level=False), scenario_creator.py:168-177): the R eMBB RAN slices of an env share ONE L1
// scheduler, so one warp steps the whole env's eMBB part: lane k owns UE k of ANY RAN slice (at most 32; the UE's RAN slice
// rides in bits 28-30 of its meta word), lane r < R also owns RAN slice r's arrival countdowns, RAN Philox counter and
// accumulators.  Arrivals / departures are per RAN slice (each with its own CAC); scheduling, reception and the channel
// are the L1's.
struct MuxPark { int cbr_next, vbr_next; uint32_t c_ran; int a_prb0, a_th0; };           // RAN slice r as lane 0 sees it on an event slot
struct MuxCtx { uint32_t c_chan, c_vbr, flags, next_dep; int n_ues; };
static __device__ __noinline__ void mux_ran_events(const StepParams &p, const int K, UeRec *ue, MuxPark *mx, const int R, uint32_t k0,
                                                   uint32_t k1, uint32_t genv, int t, uint32_t clock, MuxCtx &c) {
    PhiloxStream r_chan{k0, k1, 0u, STREAM_CHAN, c.c_chan, genv}, r_vbr{k0, k1, 0u, STREAM_VBR, c.c_vbr, genv};
    int n_ues = c.n_ues;
    uint32_t flags = c.flags;
    for (int r = 0; r < R; ++r) {                                // for slice_ran in slices_ran: slot(), extract_users, add_users (slice_l1.py:195-198)
        PhiloxStream r_ran{k0, k1, (uint32_t)r, STREAM_RAN, mx[r].c_ran, genv};
        int arr_type[2], arr_rem[2], arr_vnext[2], n_arr = 0;
        if (mx[r].cbr_next == 0) {                                                   // slice_ran.py:205-227
            mx[r].cbr_next = exp_slots_ms(r_ran, 1.0 / (2.0 / 60.0));
            const double cbr_prb = (double)mx[r].a_prb0 / (double)t;                 // cbr_cac of THIS RAN slice, :195-203
            const double cbr_th = (double)mx[r].a_th0 / ((double)t * 1e-3);
            if (!(cbr_prb >= 20.0 || cbr_th >= 10e6)) {
                arr_type[n_arr] = 0; arr_vnext[n_arr] = 0;
                arr_rem[n_arr++] = exp_slots_ms(r_ran, 30.0);
            }
        } else mx[r].cbr_next -= 1;
        if (mx[r].vbr_next == 0) {                                                   // :229-249
            arr_type[n_arr] = 1;
            arr_vnext[n_arr] = exp_slots(r_vbr, (1.0 / 1) / 1e-3);                   // VbrSource.__init__: the L1's VBR stream
            arr_rem[n_arr++] = exp_slots_ms(r_ran, 30.0);
            mx[r].vbr_next = exp_slots_ms(r_ran, 1.0 / (5.0 / 60.0));
        } else mx[r].vbr_next -= 1;
        mx[r].c_ran = r_ran.n;
        {   // departures of this RAN slice (slice_ran.py:251-261), order of the rest kept (slice_l1.py:188-191)
            int w = 0;
            for (int k = 0; k < n_ues; ++k) {
                const bool mine = (int)(ue[k].meta >> MUX_RAN_SHIFT) == r;
                if (!(mine && ue[k].dep_at == clock)) {
                    if (w != k) { UeRec tmp; load_rec(ue + k, tmp); store_rec(ue + w, tmp); }
                    ++w;
                }
            }
            n_ues = w;
        }
        for (int a = 0; a < n_arr; ++a) {                                            // add_users -> insert_user
            const int rem = arr_rem[a] - 1;
            if (rem == 0) { flags |= 8u; continue; }
            if (n_ues >= K) { flags |= 1u; continue; }
            UeRec rec;
            const int fading = (int)r_chan.integers(3);
            const int index = (int)r_chan.integers(N_SAMPLES);
            const int step = r_chan.integers(2) ? 1 : -1;
            rec.nominal = draw_nominal_sinr(r_chan, p.prop_A, p.prop_B);
            rec.meta = pack_meta(arr_type[a], fading, step, index) | ((uint32_t)r << MUX_RAN_SHIFT);
            rec.dep_at = arr_rem[a] == 0 ? DEP_NEVER : clock + (uint32_t)rem; rec.vnext = arr_vnext[a];
            rec.bits = 0; rec.th = 0.0; rec.queue = 0; rec.pe = 0; rec.nb = 0;
#pragma unroll
            for (int j = 0; j < MAX_BURSTS; ++j) rec.togo[j] = 0;
            store_rec(ue + n_ues, rec);
            ++n_ues;
        }
    }
    uint32_t nd = DEP_NEVER;
    for (int k = 0; k < n_ues; ++k) nd = min(nd, ue[k].dep_at);
    c.c_chan = r_chan.n; c.c_vbr = r_vbr.n; c.flags = flags; c.next_dep = nd; c.n_ues = n_ues;
}

// ---- TMA staging of the fading-trace columns.  After the trace walk of a TTI the column of every live UE is
// known, but its per-PRB values are only read after the PF loop (MI sums of the served sub-bands): each live lane issues one
// cp.async.bulk.tensor (2-D tensor map over the trace table, box = one whole column of 100 rows = 400 bytes) into its slot of
// the warp's shared-memory buffer, completion is signalled on the warp's mbarrier (expect_tx = live UEs x 400 bytes), and the
// warp waits on the barrier's phase right before the MI loop.  SASS: UTMALDG (cuobjdump -sass; profiles/r02e_tma_sass.txt).
constexpr int TMA_SLOT = 512;                                    // bytes per UE slot (400 used; tensor loads need 128-byte aligned destinations)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_column(void *dst, const void *tmap, int row, int col, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(row), "r"(col), "r"(smem_u32(bar)) : "memory");
}

// Resident warps per SM the register budget is sized for.  Whole batches (many waves of units: throughput) run the 96-register
// instantiation -- 18 warps per SM with the staging buffers: 2048 envs 1.051 -> 0.998 ms/step, 4096 envs 1.751 -> 1.603,
// multiplexed L1 7.20 -> 6.67 (80 registers: 1.047 / 1.653 / 6.57) --; the heavy list of the default route (one wave of the
// longest units: latency) keeps 128 registers (env step under the KBRL policy 3.13 vs 3.23 ms, 4096 envs 1.56 vs 1.63).
#ifndef RS_WARP_BATCH_WARPS
#define RS_WARP_BATCH_WARPS 20
#endif
#ifndef RS_WARP_HEAVY_WARPS
#define RS_WARP_HEAVY_WARPS 16
#endif
// batched PF step (see the RB loop below): events speculated per UE and step, bisection rounds of the budget cut, and the
// loop lengths it is tried for (full chunks left, backlogged UEs)
#ifndef RS_PF_BATCH
#define RS_PF_BATCH 5             // bit 0: in the multiplexed L1, bit 1: in the per-slice kernel (whole batches), bit 2: (heavy list).
#endif                            // Measured (profiles/r02f_pf_batch.txt): multiplexed L1 (12 UEs on ~130 PRBs: ~65 chunks per TTI) 2.45 -> 3.54 M
                                  // env-steps/s at 16 384 envs; heavy list: 4096 envs 1.565 -> 1.50 ms/step, env step under KBRL 3.10 -> 2.80 ms;
                                  // whole per-slice batches (2.8 UEs per slice, short loops) lose 9-15 % to the attempts that find nothing to
                                  // hand out: off there
#ifndef RS_PF_SPEC
#define RS_PF_SPEC 8              // (6: 3.23 M, 8: 3.54 M, 12: 2.88 M)
#endif
#ifndef RS_PF_BATCH_NMIN
#define RS_PF_BATCH_NMIN 6        // (full chunks left, backlogged UEs) >= (12, 3): 3.29 M, (6, 2): 3.54 M, (4, 2): 3.50 M, (2, 2): 3.44 M,
#endif                            // (24, 4): 3.08 M, (16, 5): 2.94 M
#ifndef RS_PF_BATCH_BMIN
#define RS_PF_BATCH_BMIN 2
#endif
constexpr int PF_SPEC = RS_PF_SPEC, PF_BISECT = 16, PF_BATCH_NMIN = RS_PF_BATCH_NMIN, PF_BATCH_BMIN = RS_PF_BATCH_BMIN;

template <bool MUX, int MIN_WARPS>
__global__ void __launch_bounds__(WP_WARPS * 32, MIN_WARPS / WP_WARPS) embb_step_warp(const __grid_constant__ StepParams p,
                                                               const __grid_constant__ EmbbState st,
                                                               const __grid_constant__ Tables tb, const int heavy_list) {
    __shared__ __align__(128) LutBlock s_lut;                    // MCS / rate LUT, snr_ref, modulation, MI constants, 1/n: ONE bulk copy
    constexpr int TK = MUX ? 32 : WP_K;                          // UE slots per unit (a multiplexed L1 holds the UEs of all its RAN slices)
    __shared__ __align__(16) UeRec s_tbl[WP_WARPS][TK];          // RAN-event scratch (8 / 16 KB)
    __shared__ WarpCtx s_ctx[WP_WARPS];
    constexpr bool BATCH = MUX ? (RS_PF_BATCH & 1) != 0 : (RS_PF_BATCH & (MIN_WARPS == RS_WARP_HEAVY_WARPS ? 4 : 2)) != 0;
    __shared__ double s_spec[WP_WARPS][BATCH ? PF_SPEC : 1][BATCH ? 32 : 1];   // speculated working throughputs of the batched PF step (2 KB per warp)
    __shared__ MuxPark s_mux[MUX ? WP_WARPS : 1][MAX_SLICES];
    __shared__ int s_racc[MUX ? WP_WARPS : 1][MAX_SLICES][2][5]; // per (RAN slice, UE type) sums of one TTI: traffic, bits, PRBs, e_snr, UEs
    __shared__ unsigned long long s_rq[MUX ? WP_WARPS : 1][MAX_SLICES][2];   // ... and queues
    extern __shared__ __align__(128) unsigned char s_cols[];     // [WP_WARPS][WP_K][TMA_SLOT] staged trace columns
    __shared__ __align__(8) uint64_t s_bar[WP_WARPS + 1];        // one mbarrier per warp (columns) + one for the lookup tables
    const int tid = threadIdx.x;
    if (tid == 0) {                                              // mcs_codeset tables: cp.async.bulk global -> shared, completion on an mbarrier
        mbar_init(&s_bar[WP_WARPS], 1);
        mbar_expect_tx(&s_bar[WP_WARPS], (unsigned)sizeof(LutBlock));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(&s_lut)), "l"(tb.lut), "r"((unsigned)sizeof(LutBlock)), "r"(smem_u32(&s_bar[WP_WARPS])) : "memory");
    }
    __syncthreads();                                             // (the barrier is initialised before anybody polls it)
    mbar_wait(&s_bar[WP_WARPS], 0);
    const int16_t *s_rate = s_lut.rate;
    const int8_t *s_mcs = s_lut.mcs, *s_mod = s_lut.mod;
    const float *s_ref = s_lut.ref, *s_inv = s_lut.inv;
    const float (*s_mi)[4] = s_lut.mi;
    const bool use_tma = !MUX && tb.tmap_ok != 0;                // fading-trace columns staged by TMA (else read through L1 / L2)

    const int lane = tid & 31, w = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned char *my_cols = s_cols + (size_t)w * WP_K * TMA_SLOT;
    unsigned tma_phase = 0;
    if (lane == 0) mbar_init(&s_bar[w], 1);
    __syncwarp();
    const int ix = blockIdx.x * WP_WARPS + w;
    // units: the heavy list of the default route (units whose PF loop was long in the previous step, embb_fast.cu
    // window_kernel), or every unit, heaviest first (variant 3: sorted front list without pair entries)
    if (ix >= (MUX ? st.U : (int)(heavy_list ? st.wlist[st.U] : st.hist[2 * SORT_BINS + 0]))) return;            // whole warp
    const int u = MUX ? ix : (heavy_list ? st.wlist[ix] : st.perm[ix]);
    const int env = MUX ? u : u / p.n_embb, s = MUX ? 0 : u - env * p.n_embb;
    const int R = MUX ? st.R : 1;
    int i_prb, n_prbs;
    uint32_t flags = 0;
    if (MUX) {                                                   // no sort pre-pass for the multiplexed L1: window from the action
        uint32_t wf = 0;
        slice_window(p, env, 0, i_prb, n_prbs, wf);
        if (lane == 0) { flags = wf; st.cur_prbs[u] = n_prbs; }
    } else unpack_window(st.win[u], i_prb, n_prbs);
    const int row_base = i_prb % TRACE_ROWS;
    MuxRan *mux = MUX ? st.mux + (size_t)u * R : nullptr;

    const UnitHdr hdr = st.hdr[u];
    UeRec *ue = st.ue + (size_t)u * st.K;
    UeRec *tbl = s_tbl[w];
    const uint64_t seed = p.seed0;
    const uint32_t genv = p.env0 + (uint32_t)env;                // global env id: Philox counter word 3
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c_ran = hdr.ctr[0], c_chan = hdr.ctr[1], c_rx = hdr.ctr[2], c_vbr = hdr.ctr[3];
    int n_ues = hdr.n_ues, cbr_next = hdr.cbr_next, vbr_next = hdr.vbr_next;
    uint32_t clock = hdr.clock;
    if (MUX) {                                                   // lane r < R: countdowns and RAN counter of RAN slice r
        cbr_next = vbr_next = 1 << 30; c_ran = 0;
        if (lane < R) { cbr_next = mux[lane].cbr_next; vbr_next = mux[lane].vbr_next; c_ran = mux[lane].c_ran; }
        for (int i = lane; i < MAX_SLICES * 2 * 5; i += 32) (&s_racc[w][0][0][0])[i] = 0;
        if (lane < MAX_SLICES * 2) (&s_rq[w][0][0])[lane] = 0ull;
        __syncwarp();
    }
    int m_tr[2] = {0, 0}, m_th[2] = {0, 0}, m_pr[2] = {0, 0};    // multiplexed L1, lane r < R: accumulators of RAN slice r by UE type
    double m_q[2] = {0.0, 0.0}, m_s[2] = {0.0, 0.0};

    UeRec r;                                                     // lane k < n_ues: UE k
    {
        int4 *z = reinterpret_cast<int4 *>(&r);
        z[0] = z[1] = z[2] = z[3] = make_int4(0, 0, 0, 0);
    }
    if (lane < n_ues) load_rec(ue + lane, r);
    uint32_t next_dep = __reduce_min_sync(FULL, lane < n_ues ? r.dep_at : DEP_NEVER);

    // slice_ran.py:270-273 reset_info: (both types, VBR only) pairs, warp-uniform
    int a_traffic_all = 0, a_traffic_v = 0, a_th_all = 0, a_th_v = 0, a_prb_all = 0, a_prb_v = 0;
    double a_queue_c = 0.0, a_queue_v = 0.0, a_snr_c = 0.0, a_snr_v = 0.0;
    unsigned trace_elems = 0, slow_snr = 0, slow_rx = 0, pf_iters = 0;      // per lane, reduced at the end
    unsigned pf_batched = 0;                                     // chunks handed out by batched PF steps (uniform; rs_get_diag)
    const float Af = (float)tb.A, Bf = (float)tb.B;
    const double inv_n = n_prbs > 0 ? tb.pre_inv / (double)n_prbs : 0.0;

    for (int t = 1; t <= p.slots; ++t) {          // slot_counter == t (zeroed by reset_info each step)
        ++clock;
        // ================= slice_ran.slot(): arrivals / departures only on event slots (serial: lane 0 on the parked table)
        if (MUX) {
            const bool any = __ballot_sync(FULL, lane < R && (cbr_next == 0 || vbr_next == 0)) != 0u || clock == next_dep;
            if (any) {
                if (lane < n_ues) store_rec(tbl + lane, r);
                if (lane < R) s_mux[w][lane] = MuxPark{cbr_next, vbr_next, c_ran, m_pr[0], m_th[0]};
                __syncwarp();
                if (lane == 0) {
                    MuxCtx c{c_chan, c_vbr, flags, next_dep, n_ues};
                    mux_ran_events(p, min(st.K, TK), tbl, s_mux[w], R, k0, k1, genv, t, clock, c);
                    s_ctx[w] = WarpCtx{0u, c.c_chan, c.c_vbr, c.next_dep, c.flags, c.n_ues, 0, 0};
                }
                __syncwarp();
                const WarpCtx c = s_ctx[w];
                c_chan = c.c_chan; c_vbr = c.c_vbr; next_dep = c.next_dep;
                if (lane == 0) flags = c.flags;
                n_ues = c.n_ues;
                if (lane < R) { cbr_next = s_mux[w][lane].cbr_next; vbr_next = s_mux[w][lane].vbr_next; c_ran = s_mux[w][lane].c_ran; }
                if (lane < n_ues) load_rec(tbl + lane, r);
                __syncwarp();
            } else if (lane < R) { cbr_next -= 1; vbr_next -= 1; }
        } else if (cbr_next == 0 || vbr_next == 0 || clock == next_dep) {
            if (lane < n_ues) store_rec(tbl + lane, r);
            __syncwarp();
            if (lane == 0) {
                RanCtx c{c_ran, c_chan, c_vbr, next_dep, flags, n_ues, cbr_next, vbr_next};
                ran_events(p, min(st.K, WP_K), tbl, k0, k1, genv, (uint32_t)s, t, clock, a_prb_all - a_prb_v, a_th_all - a_th_v, c);
                s_ctx[w] = WarpCtx{c.c_ran, c.c_chan, c.c_vbr, c.next_dep, c.flags, c.n_ues, c.cbr_next, c.vbr_next};
            }
            __syncwarp();
            const WarpCtx c = s_ctx[w];
            c_ran = c.c_ran; c_chan = c.c_chan; c_vbr = c.c_vbr; next_dep = c.next_dep;
            if (lane == 0) flags = c.flags;
            n_ues = c.n_ues; cbr_next = c.cbr_next; vbr_next = c.vbr_next;
            if (lane < n_ues) load_rec(tbl + lane, r);
            __syncwarp();
        } else { cbr_next -= 1; vbr_next -= 1; }

        // ================= per-UE traffic + SNR estimate (slice_l1.py:200-213), lane = UE
        const bool live = lane < n_ues;
        const int ty = (int)(r.meta & 1u);
        int nb_bits = 0;
        {
            const unsigned ev = __ballot_sync(FULL, live && ty && r.vnext - 1 == 0);      // VBR burst arrivals: two draws each, in UE order
            if (live) {
                if (ty) {
                    PhiloxStream rv{k0, k1, (uint32_t)s, STREAM_VBR, c_vbr + 2u * (uint32_t)__popc(ev & lt), genv};
                    nb_bits = vbr_source_step(r, rv, flags);
                } else nb_bits = 500;                            // CbrSource: 500000 b/s * 1e-3 every slot
                r.queue += nb_bits;
            }
            c_vbr += 2u * (uint32_t)__popc(ev);
        }
        int col_off = 0;
        if (n_prbs > 0) {                                        // uniform
            int index = (int)((r.meta >> 4) & MUX_INDEX_MASK), step = (r.meta & 8u) ? 1 : -1;   // (bits 28-30: RAN slice of a multiplexed L1, else 0)
            const int fading = (int)((r.meta >> 1) & 3u);
            index += step;                                       // channel_models.py:171-191
            unsigned redraw = __ballot_sync(FULL, live && (index >= N_SAMPLES - 1 || index < 0));
            while (redraw) {                                     // a trace end: variable number of draws, in UE order
                const int src = __ffs(redraw) - 1;
                redraw &= redraw - 1;
                uint32_t cn = c_chan;
                if (lane == src) {
                    PhiloxStream rc{k0, k1, (uint32_t)s, STREAM_CHAN, c_chan, genv};
                    walk_trace_redraw(rc, index, step);
                    cn = rc.n;
                }
                c_chan = __shfl_sync(FULL, cn, src);
            }
            if (live) {
                r.meta = pack_meta(ty, fading, step, index) | (r.meta & (7u << MUX_RAN_SHIFT));
                col_off = (fading * N_SAMPLES + index) * TRACE_ROWS;
                const int isum = window_sum_prefix(tb.trace_pre + (fading * N_SAMPLES + index) * PRE_STRIDE, row_base, n_prbs);
                trace_elems += (unsigned)n_prbs;
                double mean = (double)isum * inv_n + r.nominal;  // |mean - reference mean| <= 2^-(pre_bits + 1) + few ulp
                const double fr = mean - floor(mean);
                if (fabs(fr - 0.5) < tb.pre_guard) {             // within the guard of a rounding boundary: exact fp64 mean
                    mean = window_mean_fp64(tb.trace + col_off, row_base, n_prbs, r.nominal);
                    ++slow_snr;
                }
                const int e_snr = __double2int_rn(mean);         // round(np.mean(snr)), slice_ran.py:43-45
                r.pe = (r.pe & 0xFFFF) | (e_snr << 16);
            }
            if (use_tma && n_ues > 0) {                          // stage this TTI's columns; waited for right before the MI loop
                if (lane == 0) mbar_expect_tx(&s_bar[w], (unsigned)n_ues * TRACE_ROWS * 4u);
                __syncwarp();
                if (live) tma_load_column(my_cols + lane * TMA_SLOT, tb.tmap_fix, 0, col_off / TRACE_ROWS, &s_bar[w]);
            }
        }
        // scheduler inputs (schedulers.py:37-45)
        const int e = min(max(r.pe >> 16, -128), 127) + 128;
        const int rate = s_rate[e], mcs = s_mcs[e];
        double thpf = r.th > 1.0 ? r.th : 1.0;
        float metf = live && r.queue > 0 ? (float)rate * rcp_approx((float)thpf) : 0.0f;
        // update_info terms that are already final (slice_ran.py:278-305)
        int sn_all = 0, sn_v = 0, cnt_v = 0;
        const int cnt_all = n_ues;
        if (!MUX) {
            a_traffic_all += __reduce_add_sync(FULL, nb_bits); a_traffic_v += __reduce_add_sync(FULL, ty ? nb_bits : 0);
            sn_all = __reduce_add_sync(FULL, live ? (r.pe >> 16) : 0); sn_v = __reduce_add_sync(FULL, live && ty ? (r.pe >> 16) : 0);
            cnt_v = __popc(__ballot_sync(FULL, live && ty));
        }
        int n_backlog = __popc(__ballot_sync(FULL, live && r.queue > 0));

        // ================= scheduling + reception (slice_l1.py:215-224)
        int tti_bits = 0, tti_prbs = 0;                          // this UE's received bits and PRBs of the TTI (stale when unscheduled)
        if (n_backlog > 0 && n_prbs > 0) {                       // queued_data > 0 <=> some queue > 0 (uniform)
            int bits_l = 0, rbs_l = 0;                           // ue_bits, ue_rbs of this lane's UE
            int rr = 0;
            // ---- ProportionalFair.allocate RB loop (schedulers.py:47-63), phase 1: contended chunks
            while (n_backlog >= 2 && rr < n_prbs) {
                if (RS_EXP & 8) break;
                // ---- several chunks per warp-wide step (long loops only).  Event (k, j) = "UE k takes its j-th chunk from here"; its key is
                // the minimum of k's metric before each of its chunks 1..j.  The loop below hands the chunks out in the order of decreasing
                // keys (a UE's metric only changes when it is served), and a UE's state depends on HOW MANY chunks it took, not on the
                // interleaving -- so every event whose key is clearly above (a) every key any UE can still produce beyond the PF_SPEC
                // events it speculates here and (b) every event left out, is applied at once.  "Clearly": fp32 keys carry 2.4e-7, the cut
                // keeps a band of 1e-6 empty on both sides; ties, near-ties and the last (possibly 1-PRB) chunk go through the exact
                // step below.  tools/pf_stats_oracle.py: 16 chunks per step instead of 1.6 in the multiplexed L1, 6.8 instead of 2.4 else.
                if (BATCH && ((n_prbs - rr) >> 1) >= PF_BATCH_NMIN && n_backlog >= PF_BATCH_BMIN) {
                    const int nrem = (n_prbs - rr) >> 1;                 // full 2-PRB chunks left
                    const long long cap2 = 2ll * rate;
                    float key[PF_SPEC];
                    double t = thpf;
                    int bb = bits_l;
                    long long q = live ? r.queue - bits_l : 0ll;
                    float kmin = metf;
#pragma unroll
                    for (int j = 0; j < PF_SPEC; ++j) {
                        key[j] = q > 0 ? kmin : -1.0f;
                        if (q > 0) {
                            const int tx = (int)min(cap2, q);
                            q -= tx; bb += tx;
                            t = __dadd_rn(__dmul_rn(PF_A, t), b_bits_over_slot(bb));
                            kmin = fminf(kmin, q > 0 ? (float)rate * rcp_approx((float)t) : 0.0f);
                        }
                        s_spec[w][j][lane] = t;                          // working throughput after j + 1 chunks
                    }
                    // (a) the largest key still to come from beyond a horizon
                    const float T = __uint_as_float(__reduce_max_sync(FULL, __float_as_uint(q > 0 ? kmin : 0.0f)));
                    constexpr float UP = 1.0f + 1e-6f, DN = 1.0f - 1e-6f;
                    auto count = [&](float C) {                          // events clearly above C | events within the band around C
                        const float hi = C * UP, lo = C * DN;
                        int in = 0, near = 0;
#pragma unroll
                        for (int j = 0; j < PF_SPEC; ++j) { in += key[j] > hi; near += key[j] > lo; }
                        return in | ((near - in) << 16);
                    };
                    float C = T * UP;
                    int mine = count(C);
                    int tot = __reduce_add_sync(FULL, mine);
                    if ((tot & 0xFFFF) > nrem) {                         // more than the PRBs allow: raise the cut (bisection on the bit patterns)
                        // (the search starts at most a factor 256 below the best key: 16 rounds resolve 1e-4 of it; a cut that ends up
                        //  too high only leaves more chunks to the next step)
                        const unsigned top_b = __reduce_max_sync(FULL, __float_as_uint(fmaxf(key[0], 0.0f)));
                        unsigned lo_b = max(__float_as_uint(C), top_b > (8u << 23) ? top_b - (8u << 23) : 0u), hi_b = top_b + 1u;
#pragma unroll 1
                        for (int it = 0; it < PF_BISECT && hi_b - lo_b > 1u; ++it) {
                            const unsigned mid = lo_b + ((hi_b - lo_b) >> 1);
                            const int cnt = __reduce_add_sync(FULL, count(__uint_as_float(mid)) & 0xFFFF);
                            if (cnt > nrem) lo_b = mid; else hi_b = mid;
                        }
                        C = __uint_as_float(hi_b);
                        mine = count(C);
                        tot = __reduce_add_sync(FULL, mine);
                    }
                    if (tot >> 16) {                                     // events inside the band: move the cut above them, once
                        C *= 1.0f + 3e-6f;
                        mine = count(C);
                        tot = __reduce_add_sync(FULL, mine);
                    }
                    if ((tot >> 16) == 0 && (tot & 0xFFFF) > 0) {
                        const int cnt = mine & 0xFFFF;                   // chunks of this lane's UE
                        int drained = 0;
                        if (cnt > 0) {
                            const long long left_q = r.queue - bits_l;
                            const long long tx = min((long long)cnt * cap2, left_q);
                            bits_l += (int)tx;
                            rbs_l += 2 * cnt;
                            thpf = s_spec[w][cnt - 1][lane];
                            drained = left_q - tx <= 0;
                            metf = drained ? 0.0f : (float)rate * rcp_approx((float)thpf);
                        }
                        n_backlog -= __popc(__ballot_sync(FULL, drained));
                        rr += 2 * (tot & 0xFFFF);
                        pf_iters += (unsigned)(tot & 0xFFFF);
                        pf_batched += (unsigned)(tot & 0xFFFF);
                        continue;
                    }
                }
                ++pf_iters;
                const int c = min(n_prbs - rr, 2);
                // argmax of rate * (queue > 0) / th, first maximum: metrics are >= 0, so their bit patterns order like the values
                const unsigned mb = live ? __float_as_uint(metf) : 0u;
                const unsigned best_b = __reduce_max_sync(FULL, mb);
                int idx = __ffs(__ballot_sync(FULL, live && mb == best_b)) - 1;
                const unsigned second_b = __reduce_max_sync(FULL, live && lane != idx ? mb : 0u);
                const float best = __uint_as_float(best_b);
                if (__uint_as_float(second_b) >= best * (1.0f - 1e-6f)) {       // too close for fp32: exact fp64 quotients, first maximum
                    const float lim = best * (1.0f - 1e-6f);
                    const bool cand = live && metf >= lim && metf > 0.0f;
                    const unsigned long long m64 = cand ? (unsigned long long)__double_as_longlong((double)rate / thpf) : 0ull;   // > 0 for a candidate
                    const unsigned hi = __reduce_max_sync(FULL, (unsigned)(m64 >> 32));
                    const unsigned lo = __reduce_max_sync(FULL, (unsigned)(m64 >> 32) == hi ? (unsigned)m64 : 0u);
                    const unsigned win = __ballot_sync(FULL, cand && m64 == (((unsigned long long)hi << 32) | lo));
                    idx = win ? __ffs(win) - 1 : 0;
                }
                // The winner keeps taking chunks on its own while the next warp-wide argmax would pick it again without the
                // exact comparison, i.e. while the runner-up stays more than 1e-6 (relative) behind its NEW metric -- the
                // same fp32 test as above, so the sequence of winners is unchanged; 59 % of the chunks repeat the previous UE.
                int drained = 0, rr_new = rr + 2;
                if (lane == idx) {
                    const float second = __uint_as_float(second_b);
                    int cc = c;
                    for (;;) {
                        const long long left_q = r.queue - bits_l;
                        const int tx = (int)min((long long)(cc * rate), left_q);
                        bits_l += tx;
                        rbs_l += cc;
                        thpf = __dadd_rn(__dmul_rn(PF_A, thpf), b_bits_over_slot(bits_l));
                        drained = left_q - tx <= 0;
                        metf = drained ? 0.0f : (float)rate * rcp_approx((float)thpf);
                        if (drained || rr_new >= n_prbs || !(second < metf * (1.0f - 1e-6f))) break;
                        cc = min(n_prbs - rr_new, 2);
                        rr_new += 2;
                        ++pf_iters;
                    }
                }
                n_backlog -= __shfl_sync(FULL, drained, idx);
                rr = __shfl_sync(FULL, rr_new, idx);
                pf_iters = __shfl_sync(FULL, pf_iters, idx);
            }
            // phase 2: a single backlogged UE takes chunks until it is drained or the PRBs run out (closed form)
            if (rr < n_prbs && n_backlog == 1) {
                const int j = __ffs(__ballot_sync(FULL, live && r.queue - bits_l > 0)) - 1;
                int rr_new = rr;
                if (lane == j) {
                    const int left = n_prbs - rr, full = left >> 1;
                    const int cap2 = 2 * rate;
                    const long long q = r.queue - bits_l;
                    if (q <= (long long)full * cap2) {                          // drained within the 2-PRB chunks
                        const int need = (int)((q + cap2 - 1) / cap2);
                        rbs_l += 2 * need; bits_l += (int)q; rr_new = rr + 2 * need;
                    } else {
                        long long tx = (long long)full * cap2;
                        if (left & 1) tx += min((long long)rate, q - tx);       // last, single-PRB chunk
                        rbs_l += left; bits_l += (int)tx;
                        rr_new = n_prbs;
                    }
                }
                rr = __shfl_sync(FULL, rr_new, j);
            }
            // phase 3: every queue drained -> all metrics 0 -> argmax 0 with 0 bits for the remaining PRBs
            if (rr < n_prbs && lane == 0) rbs_l += n_prbs - rr;

            // ---- sub-band offsets (PRBs are handed out in UE order, schedulers.py:66-73) and MI sums, one served UE at a time
            int o = rbs_l;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(FULL, o, d); if (lane >= d) o += v; }
            o -= rbs_l;                                          // exclusive prefix
            float mavg = 0.f;
            if (use_tma && n_ues > 0) { mbar_wait(&s_bar[w], tma_phase); tma_phase ^= 1u; }
            unsigned multi = (RS_EXP & 4) ? 0u : __ballot_sync(FULL, live && rbs_l >= 2);
            while (multi) {
                const int k = __ffs(multi) - 1;
                multi &= multi - 1;
                const int rbs_k = __shfl_sync(FULL, rbs_l, k), lo = row_base + __shfl_sync(FULL, o, k), hi = lo + rbs_k;
                const int coff_k = __shfl_sync(FULL, col_off, k);
                const int m = s_mod[__shfl_sync(FULL, mcs, k)];
                const float c1 = s_mi[m][2], c0 = s_mi[m][3], nf = __shfl_sync(FULL, (float)r.nominal, k);
                const int4 *col4 = use_tma ? reinterpret_cast<const int4 *>(my_cols + k * TMA_SLOT)
                                           : reinterpret_cast<const int4 *>(tb.trace_fix + coff_k);
                const int q0 = lo >> 2, nq = ((hi - 1) >> 2) - q0 + 1;
                double msum = 0.0;
                for (int qi = lane; qi < nq; qi += 32) {
                    const int q = q0 + qi, b = q << 2;
                    const int4 x = col4[wrap_quad(q)];
                    const float e0 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.x, FIX_SCALE, nf), c1, c0));
                    const float e1 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.y, FIX_SCALE, nf), c1, c0));
                    const float e2 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.z, FIX_SCALE, nf), c1, c0));
                    const float e3 = ex2_approx(__fmaf_rn(__fmaf_rn((float)x.w, FIX_SCALE, nf), c1, c0));
                    float part = (b + 0 >= lo && b + 0 < hi) ? rcp_approx(1.0f + e0) : 0.f;
                    part += (b + 1 >= lo && b + 1 < hi) ? rcp_approx(1.0f + e1) : 0.f;
                    part += (b + 2 >= lo && b + 2 < hi) ? rcp_approx(1.0f + e2) : 0.f;
                    part += (b + 3 >= lo && b + 3 < hi) ? rcp_approx(1.0f + e3) : 0.f;
                    msum += (double)part;
                }
#pragma unroll
                for (int d = 16; d; d >>= 1) msum += shfl_xor_d(msum, d);
                if (lane == k) mavg = (float)msum * s_inv[rbs_k];            // mean MI of this UE's sub-band
            }
            // ---- per-UE reception (schedulers.py:66-76, slice_l1.py:219-224) + transmission_step (slice_ran.py:51-55)
            const unsigned served = __ballot_sync(FULL, live && rbs_l > 0);
            int b = 0;
            if (live && rbs_l > 0) {
                PhiloxStream rx{k0, k1, (uint32_t)s, STREAM_L1RX, c_rx + (uint32_t)__popc(served & lt), genv};
                const double u01 = rx.u01();
                trace_elems += (unsigned)rbs_l;
                bool received = false, need_exact = false;
                if (rbs_l == 1) need_exact = true;                           // single RB: no MI averaging, one fp64 sigmoid
                else {
                    if (mavg >= 1.0f - 1e-4f) received = true;               // p == 1.0 exactly in fp64
                    else if (mavg <= 1e-4f) received = false;                // p < 2^-53
                    else {
                        const int md = s_mod[mcs];
                        const float kf = s_mi[md][0], x0f = s_mi[md][1];
                        const float rrm = rcp_approx(mavg) - 1.0f;
                        const float seff = x0f - __logf(rrm) / kf;          // inv_sigmoid, channel_models.py:39-41
                        const float L = Af * (seff - s_ref[mcs]) - Bf;
                        const float p32 = rcp_approx(1.0f + __expf(-L));
                        const float epsL = 2e-5f / (kf * mavg * (1.0f - mavg)) + 4e-5f;   // 8 dm / (k m (1-m)), dm <= 2.5e-6
                        const float eps = 1.1f * p32 * (1.0f - p32) * epsL + 5e-7f;
                        const double d = u01 - (double)p32;
                        received = d < 0.0;
                        need_exact = fabs(d) <= (double)eps;
                    }
                }
                if (need_exact) {
                    const double pr = response_exact(tb, mcs, (size_t)col_off, (row_base + o) % TRACE_ROWS, rbs_l, r.nominal);
                    received = u01 < pr;
                    slow_rx += rbs_l > 1;
                }
                b = received ? bits_l : 0;
            }
            c_rx += (uint32_t)__popc(served);
            if (live) {
                r.queue -= b;                                    // max(queue - bits, 0): bits never exceed the queue
                r.th = __dadd_rn(__dmul_rn(PF_A, r.th), b_bits_over_slot(b));
                r.bits = b;
                r.pe = (r.pe & (int)0xFFFF0000) | rbs_l;
            }
            tti_bits = b; tti_prbs = rbs_l;
        } else {                                                 // nothing touched: stale bits / prbs accumulate
            if (use_tma && n_prbs > 0 && n_ues > 0) { mbar_wait(&s_bar[w], tma_phase); tma_phase ^= 1u; }   // (columns staged but not needed)
            tti_bits = live ? r.bits : 0; tti_prbs = live ? (r.pe & 0xFFFF) : 0;
        }
        // ================= update_info (slice_ran.py:278-305): sums over the UEs, means per UE type
        if (!MUX) {
            a_th_all += __reduce_add_sync(FULL, tti_bits); a_th_v += __reduce_add_sync(FULL, ty ? tti_bits : 0);
            a_prb_all += __reduce_add_sync(FULL, tti_prbs); a_prb_v += __reduce_add_sync(FULL, ty ? tti_prbs : 0);
            const long long qsum_all = reduce_add_ll(live ? r.queue : 0), qsum_v = reduce_add_ll(live && ty ? r.queue : 0);
            a_queue_c += div_count((double)(qsum_all - qsum_v), cnt_all - cnt_v);
            a_snr_c += div_count((double)(sn_all - sn_v), cnt_all - cnt_v);
            a_queue_v += div_count((double)qsum_v, cnt_v);
            a_snr_v += div_count((double)sn_v, cnt_v);
        } else {                                                 // every RAN slice over its own UEs: integer atomics (exact in any order), lane r collects
            if (live) {
                const int ran = (int)(r.meta >> MUX_RAN_SHIFT);
                int *A = s_racc[w][ran][ty];
                atomicAdd(&A[0], nb_bits); atomicAdd(&A[1], tti_bits); atomicAdd(&A[2], tti_prbs); atomicAdd(&A[3], r.pe >> 16); atomicAdd(&A[4], 1);
                atomicAdd(&s_rq[w][ran][ty], (unsigned long long)r.queue);
            }
            __syncwarp();
            if (lane < R) {
#pragma unroll
                for (int y = 0; y < 2; ++y) {
                    int *A = s_racc[w][lane][y];
                    m_tr[y] += A[0]; m_th[y] += A[1]; m_pr[y] += A[2];
                    const double nn = (double)max(A[4], 1);
                    m_q[y] += (double)(long long)s_rq[w][lane][y] / nn;
                    m_s[y] += (double)A[3] / nn;
                    A[0] = A[1] = A[2] = A[3] = A[4] = 0; s_rq[w][lane][y] = 0ull;
                }
            }
            __syncwarp();
        }
    }

    // ---- write the records back (once per step) and persist the slice scalars
    if (lane < n_ues) store_rec(ue + lane, r);
    flags = __reduce_or_sync(FULL, flags);
    trace_elems = __reduce_add_sync(FULL, trace_elems); slow_snr = __reduce_add_sync(FULL, slow_snr); slow_rx = __reduce_add_sync(FULL, slow_rx);
    if (MUX) {                                                   // state of every RAN slice (slice_ran.py:307-325); the L1 adds their violations up (slice_l1.py:160-171)
        int viol = 0;
        if (lane < R) {
            mux[lane].cbr_next = cbr_next; mux[lane].vbr_next = vbr_next; mux[lane].c_ran = c_ran;
            const double acc[10] = {(double)m_tr[0], (double)m_th[0], (double)m_pr[0], m_q[0], m_s[0],
                                    (double)m_tr[1], (double)m_th[1], (double)m_pr[1], m_q[1], m_s[1]};
            float *obs = p.obs + (size_t)env * p.V + lane * 10;
#pragma unroll
            for (int j = 0; j < 10; ++j) { obs[j] = (float)(acc[j] / p.norm_embb[j]); st.acc[((size_t)u * R + lane) * 10 + j] = acc[j]; }
            const double sps = (double)p.slots;
            const bool cbr_ok = acc[1] / p.obs_time > 10e6 || acc[2] / sps > 20.0 || acc[3] / sps < 10e4;
            const bool vbr_ok = acc[6] / p.obs_time > 15e6 || acc[7] / sps > 30.0 || acc[8] / sps < 15e4;
            viol = !(cbr_ok && vbr_ok);
        }
        const int l1_viol = __reduce_add_sync(FULL, viol);
        if (lane != 0) return;
        UnitHdr hm = hdr;
        hm.n_ues = n_ues; hm.clock = clock; hm.ctr[1] = c_chan; hm.ctr[2] = c_rx; hm.ctr[3] = c_vbr;
        st.hdr[u] = hm;
        p.violations[(size_t)env * p.S] = l1_viol;
        p.labels[(size_t)env * p.S] = l1_viol ? -1 : 1;
        if (flags) atomicOr(p.flags_acc + env, flags);
        if (trace_elems) atomicAdd(p.trace_elems, (unsigned long long)trace_elems);
        if (slow_snr) atomicAdd(p.slow_paths + 0, (unsigned long long)slow_snr);
        if (slow_rx) atomicAdd(p.slow_paths + 1, (unsigned long long)slow_rx);
        if (pf_batched) atomicAdd(p.slow_paths + 2, (unsigned long long)pf_batched);
        return;
    }
    if (lane != 0) return;
    UnitHdr h2 = hdr;
    h2.n_ues = n_ues; h2.cbr_next = cbr_next; h2.vbr_next = vbr_next; h2.clock = clock;
    h2.ctr[0] = c_ran; h2.ctr[1] = c_chan; h2.ctr[2] = c_rx; h2.ctr[3] = c_vbr;
    st.hdr[u] = h2;
    st.hint[u] = (pf_iters << 8) | (uint32_t)n_prbs;
    // ---- end of observation period: state, SLA label (slice_ran.py:307-325, slice_l1.py:160-171)
    const double acc[10] = {(double)(a_traffic_all - a_traffic_v), (double)(a_th_all - a_th_v), (double)(a_prb_all - a_prb_v), a_queue_c, a_snr_c,
                            (double)a_traffic_v, (double)a_th_v, (double)a_prb_v, a_queue_v, a_snr_v};
    finish_embb_unit(p, st, env, s, u, acc, flags);
    if (trace_elems) atomicAdd(p.trace_elems, (unsigned long long)trace_elems);
    if (slow_snr) atomicAdd(p.slow_paths + 0, (unsigned long long)slow_snr);
    if (slow_rx) atomicAdd(p.slow_paths + 1, (unsigned long long)slow_rx);
    if (pf_batched) atomicAdd(p.slow_paths + 2, (unsigned long long)pf_batched);
}

void launch_embb_sort(const StepParams &p, const EmbbState &st, int max_front_ues, int heavy_min_ues, cudaStream_t stream);

static size_t warp_dyn_smem() {
    static bool configured = false;
    constexpr size_t bytes = (size_t)WP_WARPS * WP_K * TMA_SLOT;
    if (!configured) {
        cudaFuncSetAttribute(embb_step_warp<false, RS_WARP_BATCH_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        cudaFuncSetAttribute(embb_step_warp<false, RS_WARP_HEAVY_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        configured = true;
    }
    return bytes;
}

// variant 3 (and the automatic route for small batches): every unit through the warp-per-unit kernel, heaviest first
int launch_embb_warp(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream, cudaEvent_t *prof) {
    launch_embb_sort(p, st, 1 << 30, 1 << 30, stream);
    if (prof) cudaEventRecord(prof[0], stream);
    embb_step_warp<false, RS_WARP_BATCH_WARPS><<<(st.U + WP_WARPS - 1) / WP_WARPS, WP_WARPS * 32, warp_dyn_smem(), stream>>>(p, st, tb, 0);
    if (prof) cudaEventRecord(prof[1], stream);
    return 5;   // kernels launched
}
// the heavy list of the default route (at most heavy_cap units), concurrent with the shared-memory kernel on another stream
void launch_embb_warp_heavy(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream) {
    embb_step_warp<false, RS_WARP_HEAVY_WARPS><<<(st.heavy_cap + WP_WARPS - 1) / WP_WARPS, WP_WARPS * 32, warp_dyn_smem(), stream>>>(p, st, tb, 1);
}

// multiplexed L1 (L1_level=False): one warp per env
void launch_embb_mux_warp(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream) {
    embb_step_warp<true, RS_WARP_BATCH_WARPS><<<(st.U + WP_WARPS - 1) / WP_WARPS, WP_WARPS * 32, 0, stream>>>(p, st, tb, 0);
}

}  // namespace rs
