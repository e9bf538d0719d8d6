/*
 * kbrl_b200.h -- C ABI of kernel #2 (KBRL inner loop) in libranslice_b200.so.
 *
 * Batched replacement for the per-learner calls KBRL_Control makes into Projectron / GaussianKernel
 * (reference: kbrl_control.py:41-114 -> algorithms/projectron.py:32-60 -> algorithms/kernel.py:8-28).
 * One learner per (env, slice): L = n_envs * n_slices learners, each with its own GROWING dictionary
 * (SVvariable, algorithms/projectron.py:3-21: landmarks, coefficients, K^-1) resident in HBM, fp64 (with the
 * reference's float32 stage while a dictionary holds a single landmark, kernel.py:15-16).  Dictionaries grow in rows
 * of 32 landmarks taken from one device pool (K^-1 packed symmetric in 32 x 32 tiles, ~D^2/2 doubles per learner);
 * nothing is pre-allocated per learner, so a few large dictionaries next to many small ones cost what they hold.
 *
 * x of a learner = [its slice's state variables (float32 -> float64), l1_prbs / n_prbs]
 * (kbrl_control.py:56,88).  Same conventions as ranslice_b200.h (0 / negative error code, rs_last_error()).
 */
#ifndef KBRL_B200_H
#define KBRL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    KB_FLAG_DICT_CAP = 1u, /* the dictionary holds dict_cap landmarks: a sample that should have been added was dropped */
    KB_FLAG_POOL = 2u      /* the device pool is exhausted: a sample that should have been added was dropped */
};

/* create_kbrl_agent(rng, n, accuracy_range) (scenario_creator.py:197-238): one
 * Learner(Projectron(GaussianKernel(SVvariable(), gamma), eta)) per slice, here for n_envs envs. */
typedef struct kb_config {
    int32_t abi_version;   /* RS_ABI_VERSION */
    int32_t device;
    int32_t n_envs, n_slices, n_prbs;
    int32_t n_variables;   /* V: row length of the state arrays */
    int32_t dict_cap;      /* most landmarks ONE learner may hold (<= 2048, rounded up to 32; 0 -> 1024).  Only bounds the
                            * per-learner row table and the shared-memory staging: memory is taken from the pool as a
                            * dictionary grows.  The reference is unbounded (KB_FLAG_DICT_CAP when hit). */
    int32_t pool_mb;       /* size of the dictionary pool in MiB; 0 -> min(what L full dictionaries need, 70 % of the free memory) */
    double gamma, eta;     /* scenario_creator.py:218 (gamma = 1), projectron.py:25 (eta = 0.1) */
    uint64_t tie_seed;     /* GaussianKernel.predict draws np.random.choice([-1, 1]) when f == 0 (kernel.py:26-27): here from the */
    uint64_t first_env_id; /* Philox stream (key tie_seed, counter (n, STREAM_KBRL, slice, first_env_id + env)), ranslice_b200/philox.py */
    int32_t algorithm;     /* 0: Projectron (projectron.py:23-60, what create_kbrl_agent builds); 1: ProjectronPlus (:66-107) */
    int32_t reserved;
} kb_config;

typedef struct kb_handle kb_handle;

/* dims[s] = len(x) of slice s (11 eMBB / 4 mMTC), offsets[s] = first state variable of slice s (Learner.indexes) */
int kb_create(const kb_config *cfg, const int32_t *dims, const int32_t *offsets, kb_handle **out);
int kb_destroy(kb_handle *h);
int kb_reset(kb_handle *h);                      /* empty dictionaries */

/* Projectron part of KBRL_Control.update_control (kbrl_control.py:88-89,103-112), all learners:
 *   y_pred[l] = predict([state_l, action_l / n_prbs])            (0 while the dictionary is empty)
 *   then, with y = labels[l]: for a in (action_l .. n_prbs) if y == +1 else (0 .. action_l):
 *        predict([state_l, a / n_prbs]); update(., y)
 * state [N][V] float32, action / labels / y_pred [N][S] int32.  DEVICE pointers, async on stream. */
int kb_update_device(kb_handle *h, const float *d_state, const int32_t *d_action, const int32_t *d_labels,
                     int32_t *d_y_pred, void *stream);
/* scan of KBRL_Control.select_action (kbrl_control.py:54-61): first_pos[l] = smallest l1_prbs in 0..n_prbs
 * with predict([state_l, l1_prbs / n_prbs]) == +1, or -1 if none.  DEVICE pointers. */
int kb_predict_device(kb_handle *h, const float *d_state, int32_t *d_first_pos, void *stream);

/* ---- device-resident KBRL_Control (kbrl_control.py:23-114): E-learner accuracies, security factors, margins, adjusted
 * flag and the current action live in HBM next to the dictionaries, so that the control loop
 *     new_state, _, _, info = system.step(action); update_control(state, action, info['SLA_labels']);
 *     action, adjusted = select_action(new_state)                                   (kbrl_control.py:126-134)
 * runs without a host round trip (rs_step_device -> kb_control_update_device -> kb_control_select_device on one stream).
 * kb_control_init: KBRL_Control.__init__ (:28-39); initial_action / security_factor [N][S] HOST arrays
 * (create_kbrl_agent draws them, scenario_creator.py:222-223), accuracies start at (lo + hi) / 2. */
int kb_control_init(kb_handle *h, const int32_t *initial_action, const int32_t *security_factor, double alfa,
                    double accuracy_lo, double accuracy_hi);
/* update_control(state, action, reward = SLA labels) (:80-114) for all learners: prediction at the taken action,
 * E-learner accuracy update, security factor (unless the last select_action adjusted that env), sample augmentation.
 * hits [N][S] = (label == prediction), may be NULL.  DEVICE pointers, async on stream. */
int kb_control_update_device(kb_handle *h, const float *d_state, const int32_t *d_action, const int32_t *d_labels,
                             int32_t *d_hits, void *stream);
/* select_action(state) (:41-73) incl. adjust_action (:75-78): action [N][S] and adjusted [N] out (either may be NULL;
 * both are also kept in the handle).  DEVICE pointers, async on stream. */
int kb_control_select_device(kb_handle *h, const float *d_state, int32_t *d_action, int32_t *d_adjusted, void *stream);
/* controller state to HOST buffers (any may be NULL): action / security_factors / margins [N][S], adjusted [N],
 * accuracies [N][S][n_prbs] */
int kb_control_get(kb_handle *h, int32_t *action, int32_t *security_factors, int32_t *margins, int32_t *adjusted,
                   double *accuracies);

/* HOST-buffer convenience wrappers (copy in, run, copy out, synchronise) */
int kb_update(kb_handle *h, const float *state, const int32_t *action, const int32_t *labels, int32_t *y_pred);
int kb_predict(kb_handle *h, const float *state, int32_t *first_pos);

/* dictionary sizes [N][S] and KB_FLAG_* per learner (host buffers, either may be NULL) */
int kb_get_sizes(kb_handle *h, int32_t *sizes, uint32_t *flags);
/* pool statistics (any may be NULL): bytes handed out / pool size, the largest dictionary, tie-break draws taken so far */
int kb_get_pool(kb_handle *h, uint64_t *used_bytes, uint64_t *total_bytes, int32_t *max_dictionary, uint64_t *tie_breaks);
/* dictionary of one learner, packed: landmarks [D][dims[s]], coeff [D], kinv [D][D]; returns D in *D_out */
int kb_get_learner(kb_handle *h, int32_t learner, double *landmarks, double *coeff, double *kinv, int32_t *D_out);
/* checkpoint / restore of the learners (dictionaries: only the part of the pool in use) and, when kb_control_init has been
 * called, of the controller state; the blob belongs to handles of the same shape (n_envs, n_slices, n_prbs, dict_cap).
 * kb_state_size synchronises and reports the size the NEXT kb_get_state needs (it grows with the dictionaries). */
int kb_state_size(kb_handle *h, size_t *bytes);
int kb_get_state(kb_handle *h, void *blob, size_t bytes);
int kb_set_state(kb_handle *h, const void *blob, size_t bytes);
/* Validation switch: with on != 0 every kernel evaluation f(x) is done in fp64 in the reference's operation order;
 * by default f is first summed in guarded fp32 and only re-evaluated in fp64 when its sign is not certain
 * (csrc/kbrl.cu eval_f_guarded).  Decisions are identical either way (tests/test_gpu_kbrl.py). */
int kb_set_exact(kb_handle *h, int on);
/* kernels launched so far; predict-mistakes (dictionary updates) applied in the last kb_update */
int kb_get_counters(kb_handle *h, uint64_t *kernel_launches, uint64_t *updates_last_call);

#ifdef __cplusplus
}
#endif
#endif
