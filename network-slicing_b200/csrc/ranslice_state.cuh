// Internal device-state layout of libranslice_b200 (not part of the C ABI).
//
// A "unit" is one (env, eMBB slice) pair: u = env * n_embb + s.  SURVEY hard-part 1 shows units
// are independent given their PRB window, so they are the grain of parallelism (one thread each).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rs {

constexpr int N_SAMPLES = 10001;   // channel_models.py:165 (time columns incl. the NaN one)
constexpr int TRACE_ROWS = 100;    // PRB rows of the trace files; rows >= 100 wrap (channel_models.py:144-148)
constexpr int N_MTC_DEV = 1000;    // scenario_creator.py:87
constexpr int MAX_SLICES = 8;
constexpr int PRE_STRIDE = 104;   // int32 words per column of the prefix table (101 used; 416 bytes = 13 sectors)

// meta word of a UE: bit0 type (0 CBR / 1 VBR), bits1-2 fading trace, bit3 step (+1 -> 1), bits4.. index
__host__ __device__ inline uint32_t pack_meta(int type, int fading, int step, int index) {
    return (uint32_t)type | ((uint32_t)fading << 1) | ((step > 0 ? 1u : 0u) << 3) | ((uint32_t)index << 4);
}

// One live UE = one 64-byte record (4 x 16 B quads, so a thread moves it with 128-bit loads/stores).
// Units are processed in PRB-sorted order (see embb_fast.cu), i.e. gathered, so the record is AoS:
// everything a UE needs in one TTI sits in two 32 B sectors.
constexpr int MAX_BURSTS = 8;      // VbrSource active bursts kept per UE (P(>8) ~ 1e-9 per UE-sample, flagged)
struct __align__(16) UeRec {
    uint32_t meta;         // type / fading / step / trace index (pack_meta)
    uint32_t dep_at;       // unit clock value at which the UE departs (slice_ran.py:222,251-261 as an absolute time; DEP_NEVER = holding drawn as 0)
    int32_t vnext;         // VbrSource.steps_to_next_arrival (traffic_generators.py:66)
    int32_t bits;          // ue.bits of the last scheduled TTI (stale when unscheduled)
    double nominal;        // nominal SINR dB                 (channel_models.py:167)
    double th;             // ue.th EWMA throughput           (slice_ran.py:55)
    long long queue;       // ue.queue, bits
    int32_t pe;            // ue.prbs (low 16) | ue.e_snr (high 16, signed)
    int32_t nb;            // active bursts
    int16_t togo[MAX_BURSTS]; // VbrSource.steps_to_go; draws are <= 18369 (53-bit uniform), saturating at -32768
};
static_assert(sizeof(UeRec) == 64, "UeRec must be 64 bytes");

// Cold part of a UE as the shared-memory kernel keeps it during a step (global scratch, embb_smem.cu)
struct __align__(16) ColdRec {
    int16_t togo[MAX_BURSTS]; // burst countdowns as of slot `sync`
    uint32_t dep_at;
    int32_t vnext;            // countdown to the next burst arrival as of slot `sync` (<= 0: never)
    uint32_t sync;            // unit clock the countdowns refer to
    uint32_t pad;
};
static_assert(sizeof(ColdRec) == 32, "ColdRec must be 32 bytes");

constexpr uint32_t DEP_NEVER = 0xFFFFFFFFu;
struct __align__(16) UnitHdr {
    int32_t n_ues;
    int32_t cbr_next;      // slice_ran.py:185 cbr_steps_next_arrival
    int32_t vbr_next;
    uint32_t clock;        // slots simulated by this unit since rs_create (wraps after 2^32 slots = 8.6e7 steps)
    uint32_t ctr[4];       // Philox draw counters: RAN, CHAN, L1RX, VBR
};
static_assert(sizeof(UnitHdr) == 32, "UnitHdr must be 32 bytes");

// RAN slice of a multiplexed L1 (create_env(L1_level=False), scenario_creator.py:168-177): arrival countdowns and the RAN
// Philox counter of SliceRANeMBB r inside the L1 unit; the UE's RAN slice id rides in bits 28-30 of UeRec::meta.
struct __align__(16) MuxRan { int32_t cbr_next, vbr_next; uint32_t c_ran, pad; };
constexpr int MUX_RAN_SHIFT = 28;
constexpr uint32_t MUX_INDEX_MASK = 0x3FFFu;               // trace index field of meta (bits 4-17) when the RAN id is present

struct EmbbState {
    int U, K, MB;          // units, UE records per unit, burst slots per UE (== MAX_BURSTS)
    int R;                 // RAN slices per unit: 1, or n_embb when the L1 multiplexes them (then U = envs, acc is [U][R][10])
    MuxRan *mux;           // [U][R] (multiplexed mode only)
    int route[4];          // routing limits of the shared-memory kernel: units starting a step with <= route[0] UEs own one lane and
                           // may grow to route[1] slots, up to route[2] UEs a pair of lanes / route[3] slots, beyond: list L (general
                           // kernel); outgrowing the slots aborts and replays.  Defaults 6 / 8 / 14 / 16; tests shrink them (rs_set_route_limits)
    int heavy_thr, heavy_cap; // default route: units whose PF loop ran >= heavy_thr contended chunks in the previous step (hint) go to the
                           // warp-per-unit kernel (at most heavy_cap of them per step), concurrently with the shared-memory kernel: a step
                           // lasts as long as its slowest lane, and those lanes are these units.  0: off
    int32_t *wlist;        // [U + 1] the heavy list of this step; wlist[U] = its length (scratch)
    int wide;              // 1: launch the latency variant of the shared-memory kernel (small batches)
    int dil, perm_len;     // lane dilution (log2) of the shared-memory kernel's front list and the length of perm[]: when a batch
                           // cannot fill the GPU, every 2^dil-th lane carries a unit and the rest idle -- fewer divergent units
                           // per warp shorten the slowest warp, which is what a step of a small batch waits for
    UnitHdr *hdr;          // [U]
    UeRec *ue;             // [U][K], live UEs first, in arrival order (order decides PF ties and RNG draw order)
    double *acc;           // [U][10] raw accumulators of the last step (info['l1_info'])
    int32_t *cur_prbs;     // [U] PRBs in force (after clamping)
    // per-step scheduling scratch (not part of the checkpoint semantics, rebuilt every step)
    uint32_t *win;         // [U] i_prb (low 16) | n_prbs (high 16) of this step
    int32_t *perm;         // [perm_len = 2U << dil] front: unit ids sorted by descending (live UEs, n_prbs, contention class); list L grows down from the end
    uint32_t *hist;        // [2 * SORT_BINS + 4 + SCAN_BLOCKS] histogram, offsets / scatter cursors, {front count, list-L count (direct + aborted), list-L count before the slice kernel (direct), pair entries}, block totals of the scan
    uint32_t *hint;        // [U] contended PF-loop iterations of the previous step << 8 | its n_prbs (sort hint only; never affects results)
    ColdRec *cold;         // [U][K] per-step scratch of the shared-memory kernel
    float *dbg;            // [8] guard-band validation maxima (debug_check runs only)
};

constexpr int SCAN_BLOCKS = 64;    // blocks of the two-kernel prefix scan over the sort bins (1024 bins each)
constexpr int SORT_BINS = 65536;   // key = pair-of-lanes bit << 15 | min(live UEs, 15) << 11 | n_prbs << 3 | contention class (3 bits)

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
    // per-step scratch (rebuilt every step, not part of the checkpoint): arrivals found by the scan kernel
    uint32_t *arr_n;       // [U] arrivals of this period (may exceed MTC_MAX_ARR: the excess is flagged and dropped)
    uint32_t *arr;         // [MTC_MAX_ARR][U] slot-in-period << 16 | device index, unordered
};
constexpr int MTC_MAX_ARR = 96;    // arrivals buffered per unit per step (mean 8.2, Poisson-like)

// The small lookup tables of the eMBB kernels packed for ONE bulk copy into shared memory (cp.async.bulk, embb_warp.cu):
// MCS / rate LUT over the integer e_snr (MCSCodeset.mcs_rate_vs_error), snr_ref and modulation per MCS, the MI sigmoid
// constants per modulation in fp32 (k, x0, -k log2 e, k x0 log2 e) and 1/n for the MI mean.
struct alignas(16) LutBlock {
    int16_t rate[256];
    int8_t mcs[256];
    float ref[32];
    int8_t mod[32];
    float mi[3][4];
    float inv[2 * TRACE_ROWS + 4];
};
static_assert(sizeof(LutBlock) % 16 == 0, "bulk copies move multiples of 16 bytes");

// stream / events on which the default route runs its heavy list next to the shared-memory kernel
struct HeavyFork { cudaStream_t stream; cudaEvent_t fork, join; };

struct Tables {
    const double *trace;   // [3][N_SAMPLES][TRACE_ROWS] fp64, time-major
    const int32_t *trace_fix; // same, fixed point round(v * 2^22): per-PRB values of the MI fast path
    const int32_t *trace_pre; // [3][N_SAMPLES][PRE_STRIDE] per-column PREFIX sums over the rows of round(v * 2^pre_bits): pre[r] = sum of
                              // rows < r, pre[100] = the whole column; a PRB-window sum is two or three loads (embb_fastmath.cuh)
    int pre_bits;             // fractional bits of trace_pre (largest that keeps every window sum of up to 300 rows inside int32; 19 for the shipped traces)
    double pre_inv, pre_guard; // 2^-pre_bits; rounding guard of the window mean: 2.5 x the representation error 2^-(pre_bits + 1)
    int8_t lut_mcs[256];   // e_snr + 128 -> mcs        (MCSCodeset.mcs_rate_vs_error, channel_models.py:288-295)
    int16_t lut_rate[256]; // e_snr + 128 -> int(158 * rate*order)  (schedulers.py:45)
    double snr_ref[26];    // mcs -> snr_ref
    int8_t mod[26];        // mcs -> modulation class
    double A, B;           // MCSCodeset.compute_factors(0.1)
    // CUtensorMap (cuTensorMapEncodeTiled) over trace_fix seen as [3 * N_SAMPLES columns][100 rows] int32, box = one whole column
    // (400 bytes): the warp-per-unit kernel stages the column of every live UE into shared memory with cp.async.bulk.tensor
    // (TMA) while the PF loop runs (embb_warp.cu, RS_WARP_TMA).  All zero when the driver entry point is unavailable.
    alignas(64) unsigned char tmap_fix[128];
    int tmap_ok;
    const LutBlock *lut;   // device copy of the packed lookup tables
};

struct StepParams {
    int N, S, n_embb, n_mmtc, n_prbs, slots, V;   // S = action entries = L1 slices; n_embb / n_mmtc = RAN slices (V = 10 n_embb + 3 n_mmtc)
    int n_l1e;                     // eMBB L1 slices: n_embb, or 1 when they are multiplexed (L1_level=False)
    int n_l1m;                     // mMTC L1 slices: n_mmtc, or 1 when they are multiplexed (one queue and one action entry for all of them)
    double penalty, prop_A, prop_B;
    double norm_embb[10], norm_mmtc[3];
    double obs_time;               // slots_per_step * slot_length (slice_ran.py:165)
    uint64_t seed0;                // base_seed: the Philox key of every env of the batch
    uint32_t env0;                 // first_env_id: global id of local env 0 (Philox counter word 3 = env0 + env)
    const int32_t *action;         // [N][S]
    float *obs;                    // [N][V]
    float *reward;                 // [N]
    int32_t *labels, *violations;  // [N][S]
    uint32_t *flags;               // [N]
    uint32_t *flags_acc;           // [N] scratch OR-ed by the slice kernels
    unsigned long long *trace_elems; // [1] algorithmic trace elements touched (B_trace counter)
    unsigned long long *slow_paths;  // [3] fp64 re-evaluations taken: [0] e_snr rounding guard, [1] reception guard; [2] PRB chunks handed out by batched PF steps
    int debug_check;                 // tests only: evaluate fp64 next to every fast decision and record error/guard ratios
};

}  // namespace rs
