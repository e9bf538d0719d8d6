// RAN events of one slot on a unit's UE table (shared by the general kernel, embb_fast.cu, and the warp-per-unit kernel,
// embb_warp.cu; `ue` may point to global or shared memory).
#pragma once
#include "embb_device.cuh"
#include "embb_fastmath.cuh"

namespace rs {

// Rare RAN events of a slot, exactly in the reference's order (slice_ran.py:263-268, slice_l1.py:196-198):
// cbr_arrivals (+CAC), vbr_arrivals, departures, extract_users, add_users -> insert_user.
static __device__ __noinline__ void ran_events(const StepParams &p, const int K, UeRec *ue, uint32_t k0, uint32_t k1, uint32_t genv,
                                        uint32_t s, int t, uint32_t clock, int a_prb0, int a_th0, RanCtx &c) {
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
        for (int k = 0; k < n_ues; ++k) {
            const uint32_t d = ue[k].dep_at;
            if (d != clock) {
                if (w != k) { UeRec tmp; load_rec(ue + k, tmp); store_rec(ue + w, tmp); }
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
        if (n_ues >= K) { flags |= 1u; continue; }
        UeRec r;
        const int fading = (int)rng.chan.integers(3);                             // channel_models.py:163-169
        const int index = (int)rng.chan.integers(N_SAMPLES);
        const int step = rng.chan.integers(2) ? 1 : -1;
        r.nominal = draw_nominal_sinr(rng.chan, p.prop_A, p.prop_B);
        r.meta = pack_meta(arr_type[a], fading, step, index);
        r.dep_at = arr_rem[a] == 0 ? DEP_NEVER : clock + (uint32_t)rem;
        r.vnext = arr_vnext[a]; r.bits = 0; r.th = 0.0; r.queue = 0; r.pe = 0; r.nb = 0;
#pragma unroll
        for (int j = 0; j < MAX_BURSTS; ++j) r.togo[j] = 0;
        store_rec(ue + n_ues, r);
        next_dep = min(next_dep, r.dep_at);
        ++n_ues;
    }
    c.c_ran = rng.ran.n; c.c_chan = rng.chan.n; c.c_vbr = rng.vbr.n;
    c.n_ues = n_ues; c.cbr_next = cbr_next; c.vbr_next = vbr_next; c.next_dep = next_dep; c.flags = flags;
}

}  // namespace rs
