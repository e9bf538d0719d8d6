/*
 * ranslice_b200.h -- C ABI of the B200-native batched RAN-slicing env (libranslice_b200.so).
 *
 * Drop-in boundary for the reference's gym env step path.  The reference has no FFI (it is
 * pure Python); each entry point below names the reference interface it replaces, N
 * independent environments at a time.  Binding shown in INTEGRATION.md (ctypes).
 *
 * Conventions: plain C, opaque handle, caller-owned buffers, return 0 on success and a negative
 * RS_E_* code on failure (rs_last_error() gives the message); no exceptions cross the boundary;
 * one handle is used from one host thread at a time.  Arrays are row-major:
 *   action [N][S] int32   PRBs per slice (eMBB slices first, then mMTC; scenario_creator.py:158-166)
 *   obs    [N][V] float32 V = 10*n_embb + 3*n_mmtc            (node_b.py:40-44)
 *   reward [N]    float32                                     (ran_slice.py:45-54)
 *   labels [N][S] int32   +1 / -1                             (slice_l1.py:160-171)
 *   violations [N][S] int32 0 / 1                             (slice_ran.py:307-319, 145-148)
 *   flags  [N]    uint32  RS_FLAG_* bits (out-of-contract events; the reference prints or crashes)
 */
#ifndef RANSLICE_B200_H
#define RANSLICE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_ABI_VERSION 2   /* 2: Philox env id in the counter, kb_config (pool, tie-break stream), rs_get_profile */

enum { RS_OK = 0, RS_E_ARG = -1, RS_E_CUDA = -2, RS_E_NOMEM = -3, RS_E_STATE = -4 };

enum {
    RS_FLAG_UE_CAP = 1u,        /* an arriving UE was dropped: max_ues live UEs in the slice            */
    RS_FLAG_BURST_CAP = 2u,     /* a VBR burst was dropped: max_bursts active bursts in the UE          */
    RS_FLAG_ACTION_CLAMP = 4u,  /* action < 0 or sum(action) > n_prbs: clamped (node_b.py:71-74 has no check) */
    RS_FLAG_SAME_SLOT_DEP = 8u, /* holding time rounded to 1 slot: UE dropped (reference: KeyError, channel_models.py:194) */
    RS_FLAG_MTC_QUEUE_CAP = 16u /* mMTC backlog exceeded mtc_queue_cap: arrival dropped                 */
};

/* scenario_creator.create_env(rng, n, slots_per_step, propagation_type, L1_level=True, penalty)
 * (scenario_creator.py:100-183) flattened into a POD.  Traffic / SLA / normalisation constants
 * (scenario_creator.py:55-96,115-134) are fixed inside the library exactly as in the reference. */
typedef struct rs_config {
    int32_t abi_version;    /* RS_ABI_VERSION */
    int32_t device;         /* CUDA device ordinal */
    int32_t n_envs;         /* N environments stepped in lockstep on this device */
    int32_t n_prbs, n_embb, n_mmtc;   /* scenario_creator.py:26-50 */
    int32_t slots_per_step; /* 50 */
    int32_t max_ues;        /* live-UE cap per eMBB slice (<= 32); 0 -> default 16 */
    int32_t max_bursts;     /* active-burst cap per VBR UE (<= 16); 0 -> default 8 */
    int32_t mtc_queue_cap;  /* mMTC backlog cap per slice; 0 -> default 128 */
    int32_t kernel_variant; /* 0 = default: the shared-memory kernel (one lane per unit), or the warp-per-unit kernel while the batch
                             *     leaves the GPU underfilled; 1 all-fp64 anchor kernel; 2 general kernel; 3 warp-per-unit kernel;
                             *     4 shared-memory kernel whatever the batch size.  Results do not depend on it. */
    int32_t l1_mux;        /* 0: every slice has its own L1 (create_env(L1_level=True), the default);
                            * 1: L1_level=False (scenario_creator.py:168-177): the n_embb eMBB RAN slices share ONE L1 scheduler and ONE
                            *    action entry, and so do the n_mmtc mMTC RAN slices (one queue); action / labels / violations are then
                            *    [N][(n_embb > 0) + (n_mmtc > 0)], violations count the
                            *    RAN slices in breach per L1, obs keeps its 10 n_embb + 3 n_mmtc columns.  Correctness-first
                            *    kernel (all fp64, one thread per env, at most 32 UEs per env). */
    double penalty;         /* ran_slice.py:19 */
    double prop_A, prop_B;  /* channel_models.py:117-124 */
    uint64_t base_seed;     /* Philox KEY of every env of the batch; the global env id first_env_id + e is counter word 3, so
                             * batches with adjacent seeds (the reference's default_rng(seed=i) replications) share no streams */
    uint64_t first_env_id;  /* global id of local env 0 (multi-GPU sharding, results invariant to the split); ids < 2^32 */
} rs_config;

/* Host pointers; copied to the device by rs_create. */
typedef struct rs_tables {
    const double *trace;    /* [3][10001][100] time-major fading traces (channel_models.py:141-150), col 10000 = NaN */
    const double *mcs_rate; /* [26] datasets/mcs_codeset.csv 'rate'  */
    const double *mcs_snr;  /* [26] 'snr'   */
    const int32_t *mcs_order; /* [26] 'order' */
    const int32_t *mcs_mod;   /* [26] 0 qpsk, 1 16qam, 2 64qam */
} rs_tables;

typedef struct rs_handle rs_handle;

/* create_env(...) -> env.   Allocates SoA state for N envs in HBM, uploads tables. */
int rs_create(const rs_config *cfg, const rs_tables *tables, rs_handle **out);
int rs_destroy(rs_handle *h);

/* RanSlice.reset() (ran_slice.py:30-36 -> node_b.py:17-22).  obs (host, may be NULL) <- zeros. */
int rs_reset(rs_handle *h, float *obs);

/* RanSlice.step(action) (ran_slice.py:38-54) for all N envs; HOST buffers (any of the outputs may
 * be NULL).  Copies action H2D, runs the step kernels, copies results D2H, synchronises. */
int rs_step(rs_handle *h, const int32_t *action, float *obs, float *reward, int32_t *labels,
            int32_t *violations, uint32_t *flags);

/* Pipelined variant of rs_step for throughput loops (no reference equivalent: the reference is synchronous).
 * rs_step_async enqueues H2D(action) + kernels on the handle's compute stream and the D2H of the results on its
 * copy stream, and returns at once with a ticket (0 or 1, alternating); rs_wait blocks until the results of that
 * ticket are in the caller's host buffers.  Two steps may be in flight: the copies of step i overlap the kernels of
 * step i+1 (device outputs are double-buffered).  The action and output buffers of a ticket must stay valid and
 * untouched until rs_wait(ticket) returns; pinned memory is needed for the overlap.  Results are identical to
 * rs_step.  Do not mix with rs_step / rs_step_device while a ticket is outstanding. */
int rs_step_async(rs_handle *h, const int32_t *action, float *obs, float *reward, int32_t *labels,
                  int32_t *violations, uint32_t *flags, int32_t *ticket);
int rs_wait(rs_handle *h, int32_t ticket);

/* Same with DEVICE buffers, asynchronous on `stream` (a cudaStream_t; NULL = legacy default). */
int rs_step_device(rs_handle *h, const int32_t *d_action, float *d_obs, float *d_reward,
                   int32_t *d_labels, int32_t *d_violations, uint32_t *d_flags, void *stream);

/* (with l1_mux the accumulator rows are the RAN slices, L1-major: [n_embb + n_mmtc][10], and n_prbs has one entry per L1)
 * info['l1_info'] of one env (node_b.py:46-49): raw accumulators of the last step, [S][10]
 * (eMBB order scenario_creator.py:80-82; mMTC: devices, avg_rep, delay, then zeros) and the
 * PRBs in force per slice [S]. */
int rs_get_info(rs_handle *h, int32_t env, double *acc, int32_t *n_prbs);

/* live UEs per eMBB slice [N][n_embb] (host buffer); population diagnostics */
int rs_get_n_ues(rs_handle *h, int32_t *n_ues);

/* checkpoint / restore of the whole device state (state_dict-style dump) */
int rs_state_size(rs_handle *h, size_t *bytes);
int rs_get_state(rs_handle *h, void *blob, size_t bytes);
int rs_set_state(rs_handle *h, const void *blob, size_t bytes);

/* counters: kernels launched by this handle so far; algorithmic fading-trace elements touched in
 * the last step summed over envs (B_trace of SURVEY 8d, 0 if the variant does not count) */
int rs_get_counters(rs_handle *h, uint64_t *kernel_launches, uint64_t *trace_elems_last_step);

/* profiling: when enabled, rs_step_device runs its kernels on ONE stream and brackets them with CUDA events recorded on
 * the launching stream; rs_get_profile synchronises and returns the summed durations (ms) of the profiled steps since the
 * last call in ms6[6]: [0] sort pre-pass (window / scan / scatter), [1] the dominant eMBB slice kernel ALONE
 * (embb_step_smem; bench.py's roofline.achieved denominator; 0 for the other variants), [2] the eMBB kernels after it
 * (general kernel over list L; the whole eMBB part for the other variants), [3] mMTC scan, [4] mMTC FIFO kernel,
 * [5] reward kernel; *steps = profiled steps. */
int rs_set_profiling(rs_handle *h, int32_t enable);
int rs_get_profile(rs_handle *h, double *ms6, uint64_t *steps);

/* guard-band validation (tests): with debug_check on, the default eMBB kernel evaluates the exact fp64
 * expression next to every fast-path decision.  rs_get_diag: out[0] max |p64 - p32| / eps over
 * reception decisions, out[1] max |mean64 - mean_fast| / guard over SNR estimates (guard = 2.5 x the fixed-point representation error), out[2] decisions
 * that would have differed (must be 0), out[3] / out[4] fp64 re-evaluations taken in the last step
 * (SNR rounding guard / reception guard); with n >= 6, out[5] = PRB chunks of the last step that the multiplexed-L1 kernel
 * handed out several at a time (its batched ProportionalFair step; the others went through the chunk-by-chunk argmax). */
int rs_set_debug_check(rs_handle *h, int32_t enable);
int rs_get_diag(rs_handle *h, double *out, int32_t n);

/* routing of the default eMBB kernel (tests; results never depend on it): a unit that starts a step with at most
 * single_start_max live UEs is stepped by one lane with single_slots UE slots in shared memory, up to pair_start_max by
 * a pair of lanes with pair_slots slots, larger ones by the general kernel; a unit that outgrows its slots during the
 * step aborts untouched and is replayed by the general kernel.  Defaults 6 / 8 / 14 / 16 (the maxima for slots).
 * rs_get_routes: units of the LAST step by route: [0] one lane, [1] pair of lanes, [2] general kernel directly,
 * [3] aborted and replayed, [4] warp-per-unit kernel (the whole batch when it is small; at lane dilution 2 the units whose PF
 * loop was long in the previous step). */
int rs_set_route_limits(rs_handle *h, int32_t single_start_max, int32_t single_slots, int32_t pair_start_max, int32_t pair_slots);
/* Heavy list of the lane-per-unit route (results never depend on it): units whose PF loop ran at least
 * contended_chunks_per_step contended chunks in the previous step are stepped by the warp-per-unit kernel, concurrently (at most
 * max_units of them, 0 = default).  A workload knob: with uniformly random allocations it only pays at the smallest batches
 * (where it is on by default, 600); a controller that allocates just enough PRBs (KBRL) leaves many slices saturated, and
 * 1000 takes 24 % off the env step at 16 384 envs.  0 switches it off.  Ignored by the other routes. */
int rs_set_heavy_threshold(rs_handle *h, int32_t contended_chunks_per_step, int32_t max_units);
int rs_get_routes(rs_handle *h, uint64_t *out5);

/* host-only self test of the exact-arithmetic identities the default kernel relies on */
int rs_selftest(void);

int rs_n_variables(const rs_handle *h);
/* the kernel variant that actually steps this handle's eMBB slices (kernel_variant 0 resolved: 3 or 4; 5 = multiplexed L1 kernel) */
int rs_active_variant(const rs_handle *h);
const char *rs_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
