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
    const uint64_t seed = p.seed0 + (uint64_t)env;
    PhiloxStream r_ran{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_RAN, hdr.ctr[0]};
    PhiloxStream r_chan{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_CHAN, hdr.ctr[1]};
    PhiloxStream r_rx{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_L1RX, hdr.ctr[2]};
    PhiloxStream r_vbr{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)s, STREAM_VBR, hdr.ctr[3]};

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

}  // namespace rs
