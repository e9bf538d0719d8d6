// K1 variant 1 ("unit-per-thread, all fp64"): one thread advances one (env, eMBB slice) unit through
// the whole observation period (slots_per_step TTIs) in a single launch, units in natural order.
// Every decision is evaluated in fp64 in the reference's operation order (compiled with
// --fmad=false).  This is the in-product correctness anchor: the default variant (embb_fast.cu)
// must produce identical results.
//
// Reference path restated here (file:line in the reference tree):
//   SliceL1eMBB.slot                slice_l1.py:193-228
//   SliceRANeMBB.slot/arrivals/...  slice_ran.py:195-305
//   VbrSource / CbrSource           traffic_generators.py:56-99
//   SINRSelectiveFading             channel_models.py:163-194
//   macro_cell / generate_xy        channel_models.py:62-97
//   ProportionalFair.allocate       schedulers.py:21-76
//   MCSCodeset.response             channel_models.py:297-313
#include "embb_device.cuh"
#include "embb_fastmath.cuh"

namespace rs {

template <int K>
__global__ void __launch_bounds__(128) embb_step_unit_thread(const __grid_constant__ StepParams p,
                                                             const __grid_constant__ EmbbState st,
                                                             const __grid_constant__ Tables tb) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= st.U) return;
    const int env = u / p.n_embb, s = u - env * p.n_embb;

    // ---- NodeB.step prologue: PRB window of this slice (node_b.py:71-74), clamped
    uint32_t flags = 0;
    int i_prb, n_prbs;
    slice_window(p, env, s, i_prb, n_prbs, flags);
    st.cur_prbs[u] = n_prbs;

    UnitHdr hdr = st.hdr[u];
    UeRec *ue = st.ue + (size_t)u * st.K;
    const uint64_t seed = p.seed0;
    const uint32_t genv = p.env0 + (uint32_t)env;   // global env id: Philox counter word 3
    PhiloxStream r_ran{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_RAN, hdr.ctr[0], genv};
    PhiloxStream r_chan{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_CHAN, hdr.ctr[1], genv};
    PhiloxStream r_rx{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_L1RX, hdr.ctr[2], genv};
    PhiloxStream r_vbr{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_VBR, hdr.ctr[3], genv};

    int n_ues = hdr.n_ues, cbr_next = hdr.cbr_next, vbr_next = hdr.vbr_next;
    uint32_t clock = hdr.clock;
    int a_traffic[2] = {0, 0}, a_th[2] = {0, 0}, a_prb[2] = {0, 0};     // slice_ran.py:270-273 reset_info
    double a_queue[2] = {0.0, 0.0}, a_snr[2] = {0.0, 0.0};
    unsigned long long trace_elems = 0;

    for (int t = 1; t <= p.slots; ++t) {          // slot_counter == t (zeroed by reset_info each step)
        ++clock;
        // ================= slice_ran.slot(): arrivals (slice_ran.py:205-249)
        int arr_type[2], arr_rem[2], arr_vnext[2], n_arr = 0;
        if (cbr_next == 0) {
            cbr_next = exp_slots_ms(r_ran, 1.0 / (2.0 / 60.0));
            const double cbr_prb = (double)a_prb[0] / (double)t;                 // cbr_cac, :195-203
            const double cbr_th = (double)a_th[0] / ((double)t * 1e-3);
            if (!(cbr_prb >= 20.0 || cbr_th >= 10e6)) {
                arr_type[n_arr] = 0; arr_vnext[n_arr] = 0;
                arr_rem[n_arr++] = exp_slots_ms(r_ran, 30.0);
            }
        } else cbr_next -= 1;
        if (vbr_next == 0) {
            arr_type[n_arr] = 1;
            arr_vnext[n_arr] = exp_slots(r_vbr, (1.0 / 1) / 1e-3);              // VbrSource.__init__, traffic_generators.py:65-66
            arr_rem[n_arr++] = exp_slots_ms(r_ran, 30.0);
            vbr_next = exp_slots_ms(r_ran, 1.0 / (5.0 / 60.0));
        } else vbr_next -= 1;
        // ================= departures (slice_ran.py:251-261) + order-preserving compaction (slice_l1.py:188-191)
        {
            int w = 0;
            for (int k = 0; k < n_ues; ++k) {
                if (ue[k].dep_at != clock) {                     // remaining_time hits 0 exactly at dep_at
                    if (w != k) { UeRec tmp; load_rec(ue + k, tmp); store_rec(ue + w, tmp); }
                    ++w;
                }
            }
            n_ues = w;
        }
        // ================= add_users (slice_l1.py:183-186) -> insert_user (channel_models.py:163-169)
        for (int a = 0; a < n_arr; ++a) {
            const int rem = arr_rem[a] - 1;                      // this slot's departures() already ticked it
            if (rem == 0) { flags |= 8u; continue; }
            if (n_ues >= st.K) { flags |= 1u; continue; }
            UeRec r;
            const int fading = (int)r_chan.integers(3);
            const int index = (int)r_chan.integers(N_SAMPLES);
            const int step = r_chan.integers(2) ? 1 : -1;
            r.nominal = draw_nominal_sinr(r_chan, p.prop_A, p.prop_B);
            r.meta = pack_meta(arr_type[a], fading, step, index);
            r.dep_at = arr_rem[a] == 0 ? DEP_NEVER : clock + (uint32_t)rem; r.vnext = arr_vnext[a]; r.bits = 0; r.th = 0.0; r.queue = 0; r.pe = 0; r.nb = 0;
#pragma unroll
            for (int j = 0; j < MAX_BURSTS; ++j) r.togo[j] = 0;
            store_rec(ue + n_ues, r);
            ++n_ues;
        }
        // ================= per-UE traffic + SNR estimate (slice_l1.py:200-213)
        long long queued = 0;
        int new_bits[K];
        for (int k = 0; k < n_ues; ++k) {
            UeRec r;
            load_rec(ue + k, r);
            int nb_bits;
            if ((r.meta & 1u) == 0) nb_bits = 500;               // CbrSource: 500000 b/s * 1e-3 every slot
            else nb_bits = vbr_source_step(r, r_vbr, flags);
            new_bits[k] = nb_bits;
            r.queue += nb_bits;
            queued += r.queue;
            if (n_prbs > 0) {
                int index = (int)(r.meta >> 4), step = (r.meta & 8u) ? 1 : -1;
                const int fading = (int)((r.meta >> 1) & 3u);
                walk_trace(r_chan, index, step);                 // channel_models.py:171-191
                r.meta = pack_meta((int)(r.meta & 1u), fading, step, index);
                const double *col = tb.trace + ((size_t)fading * N_SAMPLES + index) * TRACE_ROWS;
                double sum = 0.0;
                int row = i_prb % TRACE_ROWS;
                for (int j = 0; j < n_prbs; ++j) {
                    sum += col[row] + r.nominal;
                    row = (row + 1 == TRACE_ROWS) ? 0 : row + 1;
                }
                trace_elems += (unsigned)n_prbs;
                const int e_snr = __double2int_rn(sum / (double)n_prbs);         // round(np.mean(snr)), slice_ran.py:43-45
                r.pe = (r.pe & 0xFFFF) | (e_snr << 16);
            }
            store_rec(ue + k, r);
        }
        // ================= scheduling + reception (slice_l1.py:215-224)
        if (queued > 0 && n_prbs > 0) {
            int rbs[K], mcs[K], rate[K];
            long long bits[K], qq[K];
            double th[K];
            for (int k = 0; k < n_ues; ++k) {                    // schedulers.py:37-45
                const double uth = ue[k].th;
                th[k] = uth > 1.0 ? uth : 1.0;
                qq[k] = ue[k].queue;
                const int e = min(max(ue[k].pe >> 16, -128), 127) + 128;
                mcs[k] = tb.lut_mcs[e];
                rate[k] = tb.lut_rate[e];
                rbs[k] = 0; bits[k] = 0;
            }
            for (int r = 0; r < n_prbs; r += 2) {                // schedulers.py:47-63
                const int c = min(n_prbs - r, 2);
                int idx = 0;
                double best = -1.0;
                for (int k = 0; k < n_ues; ++k) {                // np.argmax -> first maximum
                    const double m = (double)(qq[k] > 0 ? rate[k] : 0) / th[k];
                    if (m > best) { best = m; idx = k; }
                }
                rbs[idx] += c;
                const long long cap = (long long)c * rate[idx];
                const long long tx = cap < qq[idx] ? cap : qq[idx];
                qq[idx] -= tx;
                bits[idx] += tx;
                th[idx] = PF_A * th[idx] + PF_B * (double)bits[idx] / SLOT_LEN;
            }
            int o = 0;
            for (int k = 0; k < n_ues; ++k) {                    // schedulers.py:66-76 + slice_l1.py:219-224
                const int prbs = rbs[k];
                long long b = bits[k];
                if (prbs) {
                    const uint32_t meta = ue[k].meta;
                    const double *col = tb.trace + ((size_t)((meta >> 1) & 3u) * N_SAMPLES + (meta >> 4)) * TRACE_ROWS;
                    const double pr = response_fp64(tb, mcs[k], col, (i_prb + o) % TRACE_ROWS, prbs, ue[k].nominal);
                    trace_elems += (unsigned)prbs;
                    const bool received = r_rx.u01() < pr;
                    if (!received) b = 0;
                } else b = 0;
                o += prbs;
                const long long q = ue[k].queue - b;             // UE.transmission_step, slice_ran.py:51-55
                ue[k].queue = q > 0 ? q : 0;
                ue[k].th = PF_A * ue[k].th + PF_B * (double)b / SLOT_LEN;
                ue[k].bits = (int)b;
                ue[k].pe = (ue[k].pe & 0xFFFF0000) | prbs;
            }
        }
        // ================= update_info (slice_ran.py:278-305)
        {
            long long q[2] = {0, 0};
            int sn[2] = {0, 0}, n[2] = {0, 0};
            for (int k = 0; k < n_ues; ++k) {
                const int ty = (int)(ue[k].meta & 1u);
                const int pe = ue[k].pe;
                a_traffic[ty] += new_bits[k];
                a_th[ty] += ue[k].bits;
                a_prb[ty] += pe & 0xFFFF;
                q[ty] += ue[k].queue;
                sn[ty] += pe >> 16;
                n[ty] += 1;
            }
            for (int ty = 0; ty < 2; ++ty) {
                const double nn = (double)max(n[ty], 1);
                a_queue[ty] += (double)q[ty] / nn;
                a_snr[ty] += (double)sn[ty] / nn;
            }
        }
    }

    // ---- persist slice scalars
    hdr.n_ues = n_ues; hdr.cbr_next = cbr_next; hdr.vbr_next = vbr_next; hdr.clock = clock;
    hdr.ctr[0] = r_ran.n; hdr.ctr[1] = r_chan.n; hdr.ctr[2] = r_rx.n; hdr.ctr[3] = r_vbr.n;
    st.hdr[u] = hdr;

    // ---- end of observation period: state, SLA label (slice_ran.py:307-325, slice_l1.py:160-171)
    const double acc[10] = {(double)a_traffic[0], (double)a_th[0], (double)a_prb[0], a_queue[0], a_snr[0],
                            (double)a_traffic[1], (double)a_th[1], (double)a_prb[1], a_queue[1], a_snr[1]};
    finish_embb_unit(p, st, env, s, u, acc, flags);
    if (trace_elems) atomicAdd(p.trace_elems, trace_elems);
}

void launch_embb_unit_thread(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream) {
    const int threads = 128, blocks = (st.U + threads - 1) / threads;
    // template K only sizes the per-thread scratch arrays; the cap itself is st.K (<= 32)
    if (st.K <= 16) embb_step_unit_thread<16><<<blocks, threads, 0, stream>>>(p, st, tb);
    else embb_step_unit_thread<32><<<blocks, threads, 0, stream>>>(p, st, tb);
}

// ---------------------------------------------------------------------------------------------------------------------
// Multiplexed L1 (create_env(L1_level=False), scenario_creator.py:168-177): ONE SliceL1eMBB per env holds all R = n_embb
// SliceRANeMBB slices; their UEs share one list, one PF scheduler and one PRB window (the single eMBB action entry).
// Same all-fp64 arithmetic as embb_step_unit_thread above; what changes is the slice_l1.slot() prologue -- a loop over
// the RAN slices (slice_l1.py:195-198), each with its own arrival countdowns, CAC accumulators and RAN Philox stream
// (slice index r), while the channel / reception / VbrSource draws come from the streams of the L1 (slice index 0) --
// and update_info / compute_reward, which run per RAN slice over its own UEs (slice_ran.py:278-325) before the L1 adds
// the violations up (slice_l1.py:160-171).  One thread per env; correctness-first (the reference's drivers never use
// this mode): pinned by tests/golden/B_mux*.npz through oracle/ranslice_oracle.c (l1_mux).
template <int K>
__global__ void __launch_bounds__(128) embb_step_mux_thread(const __grid_constant__ StepParams p,
                                                            const __grid_constant__ EmbbState st,
                                                            const __grid_constant__ Tables tb) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= st.U) return;
    const int env = u, s = 0, R = st.R;                        // one multiplexed eMBB L1 per env, L1 index 0

    uint32_t flags = 0;
    int i_prb, n_prbs;
    slice_window(p, env, s, i_prb, n_prbs, flags);
    st.cur_prbs[u] = n_prbs;

    UnitHdr hdr = st.hdr[u];
    UeRec *ue = st.ue + (size_t)u * st.K;
    MuxRan *mux = st.mux + (size_t)u * R;
    const uint64_t seed = p.seed0;
    const uint32_t genv = p.env0 + (uint32_t)env;   // global env id: Philox counter word 3
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    PhiloxStream r_chan{k0, k1, (uint32_t)s, STREAM_CHAN, hdr.ctr[1], genv};
    PhiloxStream r_rx{k0, k1, (uint32_t)s, STREAM_L1RX, hdr.ctr[2], genv};
    PhiloxStream r_vbr{k0, k1, (uint32_t)s, STREAM_VBR, hdr.ctr[3], genv};

    int n_ues = hdr.n_ues;
    uint32_t clock = hdr.clock;
    int cbr_next[MAX_SLICES], vbr_next[MAX_SLICES];
    uint32_t c_ran[MAX_SLICES];
    int a_traffic[MAX_SLICES][2], a_th[MAX_SLICES][2], a_prb[MAX_SLICES][2];   // slice_ran.py:270-273 reset_info, per RAN slice
    double a_queue[MAX_SLICES][2], a_snr[MAX_SLICES][2];
    for (int r = 0; r < R; ++r) {
        cbr_next[r] = mux[r].cbr_next; vbr_next[r] = mux[r].vbr_next; c_ran[r] = mux[r].c_ran;
        for (int ty = 0; ty < 2; ++ty) { a_traffic[r][ty] = a_th[r][ty] = a_prb[r][ty] = 0; a_queue[r][ty] = a_snr[r][ty] = 0.0; }
    }
    unsigned long long trace_elems = 0;

    for (int t = 1; t <= p.slots; ++t) {
        ++clock;
        // ================= for slice_ran in slices_ran: slot(), extract_users, add_users (slice_l1.py:195-198)
        for (int r = 0; r < R; ++r) {
            PhiloxStream r_ran{k0, k1, (uint32_t)r, STREAM_RAN, c_ran[r], genv};
            int arr_type[2], arr_rem[2], arr_vnext[2], n_arr = 0;
            if (cbr_next[r] == 0) {                                              // slice_ran.py:205-227
                cbr_next[r] = exp_slots_ms(r_ran, 1.0 / (2.0 / 60.0));
                const double cbr_prb = (double)a_prb[r][0] / (double)t;          // cbr_cac of THIS RAN slice, :195-203
                const double cbr_th = (double)a_th[r][0] / ((double)t * 1e-3);
                if (!(cbr_prb >= 20.0 || cbr_th >= 10e6)) {
                    arr_type[n_arr] = 0; arr_vnext[n_arr] = 0;
                    arr_rem[n_arr++] = exp_slots_ms(r_ran, 30.0);
                }
            } else cbr_next[r] -= 1;
            if (vbr_next[r] == 0) {                                              // :229-249
                arr_type[n_arr] = 1;
                arr_vnext[n_arr] = exp_slots(r_vbr, (1.0 / 1) / 1e-3);           // VbrSource.__init__: the L1's VBR stream
                arr_rem[n_arr++] = exp_slots_ms(r_ran, 30.0);
                vbr_next[r] = exp_slots_ms(r_ran, 1.0 / (5.0 / 60.0));
            } else vbr_next[r] -= 1;
            c_ran[r] = r_ran.n;
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
            for (int a = 0; a < n_arr; ++a) {                                    // add_users -> insert_user
                const int rem = arr_rem[a] - 1;
                if (rem == 0) { flags |= 8u; continue; }
                if (n_ues >= st.K) { flags |= 1u; continue; }
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
        // ================= per-UE traffic + SNR estimate over ALL UEs of the L1 (slice_l1.py:200-213)
        long long queued = 0;
        int new_bits[K];
        for (int k = 0; k < n_ues; ++k) {
            UeRec rec;
            load_rec(ue + k, rec);
            int nb_bits;
            if ((rec.meta & 1u) == 0) nb_bits = 500;
            else nb_bits = vbr_source_step(rec, r_vbr, flags);
            new_bits[k] = nb_bits;
            rec.queue += nb_bits;
            queued += rec.queue;
            if (n_prbs > 0) {
                const uint32_t ran_bits = rec.meta & (7u << MUX_RAN_SHIFT);
                int index = (int)((rec.meta >> 4) & MUX_INDEX_MASK), step = (rec.meta & 8u) ? 1 : -1;
                const int fading = (int)((rec.meta >> 1) & 3u);
                walk_trace(r_chan, index, step);
                rec.meta = pack_meta((int)(rec.meta & 1u), fading, step, index) | ran_bits;
                // window mean from the prefix table (two or three loads instead of n_prbs: a multiplexed L1 holds most of the
                // band and a dozen UEs); inside the rounding guard the reference's fp64 mean, as in the other kernels
                const int isum = window_sum_prefix(tb.trace_pre + (fading * N_SAMPLES + index) * PRE_STRIDE, i_prb % TRACE_ROWS, n_prbs);
                double mean = (double)isum * (tb.pre_inv / (double)n_prbs) + rec.nominal;
                if (fabs(mean - floor(mean) - 0.5) < tb.pre_guard)
                    mean = window_mean_fp64(tb.trace + ((size_t)fading * N_SAMPLES + index) * TRACE_ROWS, i_prb % TRACE_ROWS, n_prbs, rec.nominal);
                trace_elems += (unsigned)n_prbs;
                const int e_snr = __double2int_rn(mean);
                rec.pe = (rec.pe & 0xFFFF) | (e_snr << 16);
            }
            store_rec(ue + k, rec);
        }
        // ================= ONE scheduler over all UEs + reception from the L1's stream (slice_l1.py:215-224)
        if (queued > 0 && n_prbs > 0) {
            int rbs[K], mcs[K], rate[K];
            long long bits[K], qq[K];
            double th[K];
            for (int k = 0; k < n_ues; ++k) {
                const double uth = ue[k].th;
                th[k] = uth > 1.0 ? uth : 1.0;
                qq[k] = ue[k].queue;
                const int e = min(max(ue[k].pe >> 16, -128), 127) + 128;
                mcs[k] = tb.lut_mcs[e];
                rate[k] = tb.lut_rate[e];
                rbs[k] = 0; bits[k] = 0;
            }
            for (int rb = 0; rb < n_prbs; rb += 2) {
                const int c = min(n_prbs - rb, 2);
                int idx = 0;
                double best = -1.0;
                for (int k = 0; k < n_ues; ++k) {
                    const double m = (double)(qq[k] > 0 ? rate[k] : 0) / th[k];
                    if (m > best) { best = m; idx = k; }
                }
                rbs[idx] += c;
                const long long cap = (long long)c * rate[idx];
                const long long tx = cap < qq[idx] ? cap : qq[idx];
                qq[idx] -= tx;
                bits[idx] += tx;
                th[idx] = PF_A * th[idx] + PF_B * (double)bits[idx] / SLOT_LEN;
            }
            int o = 0;
            for (int k = 0; k < n_ues; ++k) {
                const int prbs = rbs[k];
                long long b = bits[k];
                if (prbs) {
                    const uint32_t meta = ue[k].meta;
                    const double *col = tb.trace + ((size_t)((meta >> 1) & 3u) * N_SAMPLES + ((meta >> 4) & MUX_INDEX_MASK)) * TRACE_ROWS;
                    const double pr = response_fp64(tb, mcs[k], col, (i_prb + o) % TRACE_ROWS, prbs, ue[k].nominal);
                    trace_elems += (unsigned)prbs;
                    const bool received = r_rx.u01() < pr;
                    if (!received) b = 0;
                } else b = 0;
                o += prbs;
                const long long q = ue[k].queue - b;
                ue[k].queue = q > 0 ? q : 0;
                ue[k].th = PF_A * ue[k].th + PF_B * (double)b / SLOT_LEN;
                ue[k].bits = (int)b;
                ue[k].pe = (ue[k].pe & 0xFFFF0000) | prbs;
            }
        }
        // ================= update_info of every RAN slice over its own UEs (slice_ran.py:278-305)
        for (int r = 0; r < R; ++r) {
            long long q[2] = {0, 0};
            int sn[2] = {0, 0}, n[2] = {0, 0};
            for (int k = 0; k < n_ues; ++k) {
                const uint32_t meta = ue[k].meta;
                if ((int)(meta >> MUX_RAN_SHIFT) != r) continue;
                const int ty = (int)(meta & 1u);
                const int pe = ue[k].pe;
                a_traffic[r][ty] += new_bits[k];
                a_th[r][ty] += ue[k].bits;
                a_prb[r][ty] += pe & 0xFFFF;
                q[ty] += ue[k].queue;
                sn[ty] += pe >> 16;
                n[ty] += 1;
            }
            for (int ty = 0; ty < 2; ++ty) {
                const double nn = (double)max(n[ty], 1);
                a_queue[r][ty] += (double)q[ty] / nn;
                a_snr[r][ty] += (double)sn[ty] / nn;
            }
        }
    }

    hdr.n_ues = n_ues; hdr.clock = clock;
    hdr.ctr[1] = r_chan.n; hdr.ctr[2] = r_rx.n; hdr.ctr[3] = r_vbr.n;
    st.hdr[u] = hdr;
    // ---- state of every RAN slice (slice_ran.py:307-325), violations added up by the L1 (slice_l1.py:160-171)
    int l1_viol = 0;
    const double sps = (double)p.slots;
    for (int r = 0; r < R; ++r) {
        mux[r].cbr_next = cbr_next[r]; mux[r].vbr_next = vbr_next[r]; mux[r].c_ran = c_ran[r];
        const double acc[10] = {(double)a_traffic[r][0], (double)a_th[r][0], (double)a_prb[r][0], a_queue[r][0], a_snr[r][0],
                                (double)a_traffic[r][1], (double)a_th[r][1], (double)a_prb[r][1], a_queue[r][1], a_snr[r][1]};
        float *obs = p.obs + (size_t)env * p.V + r * 10;
        for (int j = 0; j < 10; ++j) {
            obs[j] = (float)(acc[j] / p.norm_embb[j]);
            st.acc[((size_t)u * R + r) * 10 + j] = acc[j];
        }
        const bool cbr_ok = acc[1] / p.obs_time > 10e6 || acc[2] / sps > 20.0 || acc[3] / sps < 10e4;
        const bool vbr_ok = acc[6] / p.obs_time > 15e6 || acc[7] / sps > 30.0 || acc[8] / sps < 15e4;
        l1_viol += !(cbr_ok && vbr_ok);
    }
    p.violations[(size_t)env * p.S + s] = l1_viol;
    p.labels[(size_t)env * p.S + s] = l1_viol ? -1 : 1;
    if (flags) atomicOr(p.flags_acc + env, flags);
    if (trace_elems) atomicAdd(p.trace_elems, trace_elems);
}

void launch_embb_mux(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream) {
    const int threads = 128, blocks = (st.U + threads - 1) / threads;
    embb_step_mux_thread<32><<<blocks, threads, 0, stream>>>(p, st, tb);
}

// NodeB.reset of the multiplexed RAN slices (slice_ran.py:182-190): countdowns cleared, RAN counters keep running
__global__ void __launch_bounds__(256) embb_mux_reset_kernel(const __grid_constant__ EmbbState st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.U * st.R) return;
    st.mux[i].cbr_next = 0; st.mux[i].vbr_next = 0;
    for (int j = 0; j < 10; ++j) st.acc[(size_t)i * 10 + j] = 0.0;
}
void launch_embb_mux_reset(const EmbbState &st, cudaStream_t stream) {
    embb_mux_reset_kernel<<<(st.U * st.R + 255) / 256, 256, 0, stream>>>(st);
}

}  // namespace rs
