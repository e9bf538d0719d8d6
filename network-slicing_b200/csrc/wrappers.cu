// Batched agent-side wrappers (C ABI: include/wrapper_b200.h): the arithmetic ReportWrapper / DQNWrapper do around
// env.step() (reference wrapper.py:71-154), for N envs per launch.  All four are a few bytes per env: one thread per
// env (or per element), coalesced, launched on the caller's stream between the agent and rs_step_device.
#include <cuda_runtime.h>

#include <string>

#include "../../include/ranslice_b200.h"
#include "../../include/wrapper_b200.h"

extern "C" void rs_set_error(const char *msg);

namespace rw {

constexpr int MAX_S = 8;

// np.sum of a short 1-D array: sequential below 8 elements, otherwise 8 running sums combined pairwise and a
// sequential tail (numpy's pairwise_sum for n < 128).  n <= MAX_S + 1 = 9 here.
template <typename T>
__device__ __forceinline__ T numpy_sum(const T *a, int n) {
    if (n < 8) {
        T r = (T)0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    T r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    for (int i = 8; i < n; ++i) r += a[i];
    return r;
}

// wrapper.py:77-82
template <typename T>
__global__ void __launch_bounds__(256) map_action_kernel(const T *__restrict__ action, int N, int S, int n_prbs,
                                                         int32_t *__restrict__ prbs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    T a[MAX_S + 1];
    for (int i = 0; i <= S; ++i) { const T v = action[(size_t)e * (S + 1) + i]; a[i] = v < (T)0 ? -v : v; }   // abs(action)
    T t = numpy_sum(a, S + 1);
    if (t == (T)0) t = (T)1;
    for (int i = 0; i < S; ++i) {
        const T q = ((T)n_prbs * a[i]) / t;                   // n_prbs * action[i] / t_action, left to right, in dtype T
        prbs[(size_t)e * S + i] = (int32_t)floor(q);
    }
}

// wrapper.py:152-154
__global__ void __launch_bounds__(256) dqn_action_kernel(const int32_t *__restrict__ index, const int32_t *__restrict__ table,
                                                         int N, int S, int A, int32_t *__restrict__ prbs, uint32_t *flags) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    const int ix = index[e];
    const bool ok = ix >= 0 && ix < A;
    for (int i = 0; i < S; ++i) prbs[(size_t)e * S + i] = ok ? table[(size_t)ix * S + i] : 0;
    if (!ok && flags) atomicOr(flags + e, 1u);
}

// wrapper.py:88-90
__global__ void __launch_bounds__(256) obs_kernel(const float *__restrict__ obs, float *__restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = obs[i];
    v = fminf(fmaxf(v, -0.5f), 1.5f);                         // np.clip (NaN passes through, like numpy)
    if (obs[i] != obs[i]) v = obs[i];
    out[i] = v - 0.5f;
}

// wrapper.py:101-106
__global__ void __launch_bounds__(256) record_kernel(const int32_t *__restrict__ violations, const float *__restrict__ reward,
                                                     const int32_t *__restrict__ prbs, int N, int S, long long step,
                                                     int16_t *vh, double *rh, int16_t *ah) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    int v = 0, a = 0;
    for (int i = 0; i < S; ++i) { v += violations[(size_t)e * S + i]; a += prbs[(size_t)e * S + i]; }
    const size_t o = (size_t)step * N + e;
    if (vh) vh[o] = (int16_t)v;
    if (rh) rh[o] = (double)reward[e];
    if (ah) ah[o] = (int16_t)a;
}

int fail(int code, const char *m) { rs_set_error(m); return code; }
int check_launch() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { rs_set_error((std::string("wrapper kernel launch: ") + cudaGetErrorString(e)).c_str()); return RS_E_CUDA; }
    return RS_OK;
}

}  // namespace rw

extern "C" {

int rs_wrap_action_device(const void *d_action, int32_t is_f64, int32_t n_envs, int32_t n_slices, int32_t n_prbs,
                          int32_t *d_prbs, void *stream) {
    if (!d_action || !d_prbs) return rw::fail(RS_E_ARG, "null argument");
    if (n_envs <= 0 || n_slices <= 0 || n_slices > rw::MAX_S || n_prbs <= 0) return rw::fail(RS_E_ARG, "bad n_envs / n_slices / n_prbs");
    const int blocks = (n_envs + 255) / 256;
    if (is_f64) rw::map_action_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>((const double *)d_action, n_envs, n_slices, n_prbs, d_prbs);
    else rw::map_action_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float *)d_action, n_envs, n_slices, n_prbs, d_prbs);
    return rw::check_launch();
}

int rs_wrap_dqn_action_device(const int32_t *d_index, const int32_t *d_table, int32_t n_envs, int32_t n_slices,
                              int32_t n_actions, int32_t *d_prbs, uint32_t *d_flags, void *stream) {
    if (!d_index || !d_table || !d_prbs) return rw::fail(RS_E_ARG, "null argument");
    if (n_envs <= 0 || n_slices <= 0 || n_slices > rw::MAX_S || n_actions <= 0) return rw::fail(RS_E_ARG, "bad n_envs / n_slices / n_actions");
    rw::dqn_action_kernel<<<(n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_index, d_table, n_envs, n_slices, n_actions, d_prbs, d_flags);
    return rw::check_launch();
}

int rs_wrap_obs_device(const float *d_obs, float *d_obs_out, int64_t n, void *stream) {
    if (!d_obs || !d_obs_out) return rw::fail(RS_E_ARG, "null argument");
    if (n <= 0) return rw::fail(RS_E_ARG, "bad element count");
    rw::obs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_obs, d_obs_out, (long long)n);
    return rw::check_launch();
}

int rs_wrap_record_device(const int32_t *d_violations, const float *d_reward, const int32_t *d_prbs, int32_t n_envs,
                          int32_t n_slices, int64_t step, int16_t *d_violation_hist, double *d_reward_hist,
                          int16_t *d_resources_hist, void *stream) {
    if (!d_violations || !d_reward || !d_prbs) return rw::fail(RS_E_ARG, "null argument");
    if (n_envs <= 0 || n_slices <= 0 || n_slices > rw::MAX_S || step < 0) return rw::fail(RS_E_ARG, "bad n_envs / n_slices / step");
    rw::record_kernel<<<(n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_violations, d_reward, d_prbs, n_envs, n_slices,
                                                                             (long long)step, d_violation_hist, d_reward_hist, d_resources_hist);
    return rw::check_launch();
}

}  // extern "C"
