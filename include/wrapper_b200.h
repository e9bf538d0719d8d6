/*
 * wrapper_b200.h -- C ABI of the batched agent-side wrappers in libranslice_b200.so.
 *
 * Device-side equivalents of what the reference's gym wrappers do around env.step() for ONE env
 * (reference: wrapper.py ReportWrapper.step :71-117, DQNWrapper :132-154), here for N envs per call so that a
 * model-free agent can drive the batched env without per-env Python:
 *   - the action mapping of a continuous agent (simplex weights -> PRBs per slice), wrapper.py:77-82;
 *   - the discrete action table of DQNWrapper, wrapper.py:141-154;
 *   - the observation normalisation clip(obs, -0.5, 1.5) - 0.5, wrapper.py:88-90;
 *   - the history buffers violation / reward / resources, wrapper.py:101-106 (written to .npz by the host side
 *     with the reference's keys, wrapper.py:120-123).
 * Stateless: plain DEVICE pointers and sizes, asynchronous on the given stream.  Same error conventions as
 * ranslice_b200.h (0 / negative code, rs_last_error()).
 */
#ifndef WRAPPER_B200_H
#define WRAPPER_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* action [N][S+1] float32 (is_f64 == 0) or float64 (is_f64 != 0), as an agent with action_space
 * Box(0, 1, (S+1,)) emits it  ->  prbs [N][S] int32:
 *   a = |action|; t = a.sum() (numpy's summation order; t == 0 -> 1); prbs[i] = floor(n_prbs * a[i] / t), i < S
 * evaluated in the array's own dtype like numpy does (wrapper.py:77-82). */
int rs_wrap_action_device(const void *d_action, int32_t is_f64, int32_t n_envs, int32_t n_slices, int32_t n_prbs,
                          int32_t *d_prbs, void *stream);

/* DQNWrapper.step (wrapper.py:152-154): prbs[e][:] = table[index[e]][:]; table [n_actions][S] int32 is
 * DQNWrapper.actions (built by the host side exactly like wrapper.py:141-149).  An index outside
 * [0, n_actions) yields all-zero PRBs and sets bit 0 of d_flags[e] (d_flags may be NULL). */
int rs_wrap_dqn_action_device(const int32_t *d_index, const int32_t *d_table, int32_t n_envs, int32_t n_slices,
                              int32_t n_actions, int32_t *d_prbs, uint32_t *d_flags, void *stream);

/* obs_out[i] = clip(obs[i], -0.5, 1.5) - 0.5 in float32 (wrapper.py:88-90); n = N * V elements; in place allowed. */
int rs_wrap_obs_device(const float *d_obs, float *d_obs_out, int64_t n, void *stream);

/* History row of one step (wrapper.py:101-106), all envs: violation[step][e] = sum_s violations[e][s],
 * reward[step][e] = reward[e], resources[step][e] = sum_s prbs[e][s].  Histories are [steps][N] (step-major so that
 * a step writes one contiguous row); int16 like the reference's np.int16 buffers, reward float64. */
int rs_wrap_record_device(const int32_t *d_violations, const float *d_reward, const int32_t *d_prbs, int32_t n_envs,
                          int32_t n_slices, int64_t step, int16_t *d_violation_hist, double *d_reward_hist,
                          int16_t *d_resources_hist, void *stream);

#ifdef __cplusplus
}
#endif
#endif
