// Internal device-state layout of libranslice_b200 (not part of the C ABI).
//
// A "unit" is one (env, eMBB slice) pair: u = env * n_embb + s.  SURVEY hard-part 1 shows units
// are independent given their PRB window, so they are the grain of parallelism.  All per-UE
// fields are SoA, UE-slot-major: field[k * U + u], so that neighbouring units (neighbouring
// lanes) touch neighbouring addresses (coalesced 128 B per warp per field).
#pragma once
#include <cstdint>

namespace rs {

constexpr int N_SAMPLES = 10001;   // channel_models.py:165 (time columns incl. the NaN one)
constexpr int TRACE_ROWS = 100;    // PRB rows of the trace files; rows >= 100 wrap (channel_models.py:144-148)
constexpr int N_MTC_DEV = 1000;    // scenario_creator.py:87
constexpr int MAX_SLICES = 8;

// meta word of a UE: bit0 type (0 CBR / 1 VBR), bits1-2 fading trace, bit3 step (+1 -> 1), bits4.. index
__host__ __device__ inline uint32_t pack_meta(int type, int fading, int step, int index) {
    return (uint32_t)type | ((uint32_t)fading << 1) | ((step > 0 ? 1u : 0u) << 3) | ((uint32_t)index << 4);
}

struct EmbbState {
    int U, K, MB;          // units, UE slots per unit, burst slots per UE
    int32_t *n_ues;        // [U]
    int32_t *cbr_next;     // [U] slice_ran.py:185 cbr_steps_next_arrival
    int32_t *vbr_next;     // [U]
    uint32_t *ctr;         // [4][U] Philox draw counters: RAN, CHAN, L1RX, VBR
    uint32_t *meta;        // [K][U]
    int32_t *rem;          // [K][U] remaining holding time (slice_ran.py:222)
    double *nominal;       // [K][U] nominal SINR dB (channel_models.py:167)
    long long *queue;      // [K][U] ue.queue (bits)
    double *th;            // [K][U] ue.th EWMA throughput
    int32_t *bits;         // [K][U] ue.bits of the last scheduled TTI (stale when unscheduled)
    int32_t *pe;           // [K][U] ue.prbs (low 16) | ue.e_snr (high 16, signed)
    int32_t *vnext;        // [K][U] VbrSource.steps_to_next_arrival
    int32_t *nb;           // [K][U] active bursts
    int32_t *togo;         // [K][MB][U] VbrSource.steps_to_go
    double *acc;           // [U][10] raw accumulators of the last step (info['l1_info'])
    int32_t *cur_prbs;     // [U] PRBs in force (after clamping)
};

struct MmtcState {
    int U, Q;              // units (env * n_mmtc + m), backlog cap
    uint32_t *next_abs;    // [N_MTC_DEV][U] absolute slot of the next message arrival
    uint8_t *period_ix;    // [N_MTC_DEV][U] index into period_set
    uint8_t *rep_ix;       // [N_MTC_DEV][U] index into repetition_set
    int32_t *q_rep;        // [Q][U] remaining repetitions, FIFO order (slice_l1.py:30-33)
    uint32_t *q_t0;        // [Q][U] arrival slot
    int32_t *q_n;          // [U]
    uint32_t *time;        // [U] SliceL1mMTC.time
    uint32_t *ctr;         // [U] Philox counter (STREAM_MTC)
    double *acc;           // [U][3] devices, avg_rep, delay
    int32_t *cur_prbs;     // [U]
};

struct Tables {
    const double *trace;   // [3][N_SAMPLES][TRACE_ROWS] fp64, time-major
    const float *trace32;  // same, fp32 (fast path; decisions near a boundary are redone in fp64)
    int8_t lut_mcs[256];   // e_snr + 128 -> mcs        (MCSCodeset.mcs_rate_vs_error, channel_models.py:288-295)
    int16_t lut_rate[256]; // e_snr + 128 -> int(158 * rate*order)  (schedulers.py:45)
    double snr_ref[26];    // mcs -> snr_ref
    int8_t mod[26];        // mcs -> modulation class
    double A, B;           // MCSCodeset.compute_factors(0.1)
};

struct StepParams {
    int N, S, n_embb, n_mmtc, n_prbs, slots, V;
    double penalty, prop_A, prop_B;
    double norm_embb[10], norm_mmtc[3];
    double obs_time;               // slots_per_step * slot_length (slice_ran.py:165)
    uint64_t seed0;                // base_seed + first_env_id
    const int32_t *action;         // [N][S]
    float *obs;                    // [N][V]
    float *reward;                 // [N]
    int32_t *labels, *violations;  // [N][S]
    uint32_t *flags;               // [N]
    uint32_t *flags_acc;           // [N] scratch OR-ed by the slice kernels
    unsigned long long *trace_elems; // [1] algorithmic trace elements touched (B_trace counter)
};

}  // namespace rs
