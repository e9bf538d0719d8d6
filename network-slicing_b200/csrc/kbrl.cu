// K3: KBRL inner loop -- Gaussian-kernel evaluation and Projectron dictionary projection / update,
// batched over L = n_envs * n_slices independent learners (C ABI: include/kbrl_b200.h).
//
//   GaussianKernel.k_eval / k / predict   algorithms/kernel.py:8-28
//   Projectron.predict / update           algorithms/projectron.py:32-60
//   callers restated: the select_action scan (kbrl_control.py:54-61) and the sample-augmentation loop of
//   update_control (kbrl_control.py:103-112)
//
// One thread GROUP per learner (KB_GROUP = 32 threads = one warp by default, 4 groups per 128-thread block): a
// learner-step is a chain of short phases (stage the dictionary, evaluate <= 201 candidates, find the first mistake,
// project / extend) separated by barriers, so what matters is how many learners an SM keeps in flight and how cheap
// the barriers are -- with one 256-thread CTA per learner the kernels spent 9.8 barrier-stall cycles per issued
// instruction (profiles/r01f_kb_update_16384_full.txt).  Both callers evaluate f(a) = sum_j coeff_j exp(-gamma ||l_j - [s, a/n]||^2) for
// up to n_prbs + 1 candidate allocations a of one state s.  numpy's pairwise sum over the 11 (4)
// squared differences adds the action coordinate LAST, so the distance is separable bit-exactly:
// base_j (state part, once per landmark) + (l_j,last - a/n)^2.  Threads own candidates and walk the
// landmarks (broadcast from shared memory) in index order; a dictionary update (rare: ~0.3 per
// learner-step) is a block-cooperative K^-1 k mat-vec (columns are coalesced, K^-1 is exactly
// symmetric) and, when the sample is not well approximated, a rank-1 extension of K^-1.
// The Gram / K^-1 products stay warp-reduced fp64: a D x D mat-vec 1.4 times per env-step is far below
// any tensor-core crossover.
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <vector>

#include "../../include/kbrl_b200.h"
#include "../../include/ranslice_b200.h"

namespace kb {

// Threads per learner.  Measured in the config-3 loop (16 384 envs, D mean 27, cap 128; tools/gpu_kb_sweep.sh):
//   update_kernel   32: 8.9 ms   64: 5.4 ms   128: 3.9 ms   256: 3.4 ms   -- work-bound per learner (K^-1 mat-vec and rank-1
//                   extension over D^2 elements, <= 201 candidates x D terms per round): wide groups win; routing
//                   learners with D <= 16/32/64 to a warp each and the rest to a CTA each: 3.9-4.0 vs 3.8 ms (no gain)
//   predict_kernel  32: 0.44 ms  128: 0.54 ms  256: 0.62 ms               -- one short pass: more learners in flight win
#ifndef KB_GROUP_UPDATE
#define KB_GROUP_UPDATE 256
#endif
#ifndef KB_GROUP_PREDICT
#define KB_GROUP_PREDICT 32
#endif
constexpr size_t GROUP_SMEM_PER_LANDMARK = 5 * sizeof(double) + sizeof(float4);   // base, coeff, last coord, k, d*, fp32 copy
template <int G> struct Cfg {
    static_assert(G == 32 || G == 64 || G == 128 || G == 256, "threads per learner: 32, 64, 128 or 256");
    static constexpr int THREADS = G < 128 ? 128 : G;        // threads per block
    static constexpr int GROUPS = THREADS / G;               // learners per block
};
// barrier among the G threads that work on one learner
template <int G> __device__ __forceinline__ void gsync(int group) {
    if (G == 32) __syncwarp();
    else if (Cfg<G>::GROUPS == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(G) : "memory");
}

constexpr int MAX_CAND = 256;      // n_prbs + 1 <= 201
constexpr int MAX_DIM = 16;

struct State {
    int L, S, V, n_prbs, cap;
    int exact_only;    // kb_set_exact: evaluate every f in fp64 like the reference (validation of the guarded fast path)
    double gamma, eta;
    int dims[8], offs[8];
    int *D;            // [L]
    double *lm;        // [L][cap][MAX_DIM]
    double *coeff;     // [L][cap]
    double *kinv;      // [L][cap][cap] (exactly symmetric)
    uint32_t *flags;   // [L]
    unsigned long long *updates;   // [1]
};

// Device-resident KBRL_Control state (kbrl_control.py:28-39); acc == nullptr while kb_control_init has not been called
struct Control {
    double *acc;       // [L][n_prbs] E-learner accuracies
    int32_t *sec;      // [L] security_factors
    int32_t *margins;  // [L]
    int32_t *adjusted; // [N] self.adjusted of the last select_action
    int32_t *action;   // [L] self.action
    int32_t *first;    // [L] scratch: first allocation predicted +1 (select_action scan)
    double alfa, acc_lo;
};

// state part of ((l - x)**2).sum() in numpy's pairwise order (action coordinate excluded; it is added last)
__device__ __forceinline__ double base_dist(const double *l, const double *x, int ns) {
    if (ns < 8) {                                   // n = ns + 1 < 8 or exactly the sequential tail below
        double r = 0.0;
        for (int i = 0; i < ns; ++i) { const double t = l[i] - x[i]; r += t * t; }
        return r;
    }
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double t = l[i] - x[i]; a[i] = t * t; }
    double r = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    for (int i = 8; i < ns; ++i) { const double t = l[i] - x[i]; r += t * t; }
    return r;
}

// shared-memory slice of one group
__device__ __forceinline__ void carve(unsigned char *raw, int cap, int group, double *&base, double *&cf, double *&ll, double *&kf,
                                      double *&ds, float4 *&fast) {
    double *p = reinterpret_cast<double *>(raw + (size_t)group * cap * GROUP_SMEM_PER_LANDMARK);
    base = p; cf = base + cap; ll = cf + cap; kf = ll + cap; ds = kf + cap;
    fast = reinterpret_cast<float4 *>(ds + cap);             // [cap] fp32 copy of the staged dictionary (guarded fast path)
}
// per-group scalars
struct GroupVars { double xs[MAX_DIM]; double delta, fa0; unsigned long long minb; int first, D, sf, pad; };

// f(a) for candidate a with the dictionary staged in shared memory (kernel.py:13-25 incl. the D == 1 float32 stage)
__device__ __forceinline__ double eval_f(int D, double gamma, const double *base, const double *cf, const double *ll, double xa) {
    if (D == 0) return 0.0;
    if (D == 1) {
        const double t = ll[0] - xa;
        const float k = (float)exp(-gamma * (base[0] + t * t));
        return (double)(k * (float)cf[0]);
    }
    double f = 0.0;
    for (int j = 0; j < D; ++j) {
        const double t = ll[j] - xa;
        f += exp(-gamma * (base[j] + t * t)) * cf[j];
    }
    return f;
}

// Guarded fp32 evaluation of f(a).  Only the SIGN of f is ever used (kernel.py:25, projectron.py:40), so f is first
// summed in fp32 with ex2.approx together with G = sum |coeff_j| k_j; the fp32 result differs from the reference's
// fp64 sum by less than (1e-5 + 1.2e-7 D) G  (|arg| <= 126 before a term flushes to zero: argument error <= 5e-7 +
// 1.2e-7 |arg|, ex2.approx 2^-22, D sequential fp32 additions); the guard (4e-5 + 3e-7 D) G leaves a factor 2.5-4.  Inside that band -- a candidate sitting on the decision boundary -- f is re-evaluated
// exactly like the reference (eval_f).  Dictionaries of 0 or 1 landmarks (the reference's float32 stage) go straight
// to eval_f.
constexpr float LOG2E = 1.4426950408889634f;
#ifdef KB_CHECK
__device__ unsigned long long g_kb_dbg[4];   // [0] accepted fast results, [1] sign mismatches among them, [2] max |f32-f64|/G * 1e9, [3] guard hits
#endif
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ double eval_f_guarded(int D, double gamma, const double *base, const double *cf, const double *ll,
                                                 const float4 *fast, bool fast_ok, double xa, unsigned *slow) {
#ifdef KB_NO_FAST
    return eval_f(D, gamma, base, cf, ll, xa);
#endif
    if (D <= 1 || !fast_ok) return eval_f(D, gamma, base, cf, ll, xa);
    const float xf = (float)xa, g2 = (float)gamma * LOG2E;
    float f = 0.f, G = 0.f;
    for (int j = 0; j < D; ++j) {
        const float4 e = fast[j];                            // x: gamma log2(e) base_j, y: coeff_j, z: l_j[last]
        const float t = e.z - xf;
        const float term = e.y * ex2f(-fmaf(t * t, g2, e.x));
        f += term;
        G += fabsf(term);
    }
#ifdef KB_CHECK
    {
        const double fx = eval_f(D, gamma, base, cf, ll, xa);
        const bool acc = fabsf(f) > (4e-5f + 3e-7f * (float)D) * G;
        if (acc) {
            atomicAdd(&g_kb_dbg[0], 1ull);
            if ((fx > 0.0) != (f > 0.f)) atomicAdd(&g_kb_dbg[1], 1ull);
            if (G > 0.f) atomicMax(&g_kb_dbg[2], (unsigned long long)(fabs((double)f - fx) / (double)G * 1e9));
        } else atomicAdd(&g_kb_dbg[3], 1ull);
    }
#endif
    if (fabsf(f) > (4e-5f + 3e-7f * (float)D) * G) return (double)f;
    if (slow) ++*slow;
    return eval_f(D, gamma, base, cf, ll, xa);
}

// Stages the dictionary of learner l for state xs: exact fp64 (base, coeff, last coordinate) and the fp32 copy of the
// fast path.  The fp32 exponents are taken RELATIVE to the closest landmark (min_j base_j): sign(f) does not change
// when every term is scaled by exp(gamma min_j base_j), and without the shift a state far from all landmarks puts
// every term below the fp32 normal range (measured: flushed terms flipped 140 of 1.1e9 decisions).
// Group-wide; ends with the dictionary visible to all threads of the group (callers need no further barrier for it).
// Returns whether the fp32 fast path may be used for this (dictionary, state).
template <int GROUP>
__device__ bool stage_dictionary(const State &kb, int l, int D, int d, GroupVars &g, int group, int gt, double *base, double *cf,
                                 double *ll, float4 *fast) {
    const double *lm = kb.lm + (size_t)l * kb.cap * MAX_DIM;
    const double g2 = kb.gamma * 1.4426950408889634;
    if (gt == 0) g.minb = ~0ull;
    gsync<GROUP>(group);
    for (int j = gt; j < D; j += GROUP) {
        const double b = base_dist(lm + (size_t)j * MAX_DIM, g.xs, d - 1);
        base[j] = b; cf[j] = kb.coeff[(size_t)l * kb.cap + j]; ll[j] = lm[(size_t)j * MAX_DIM + d - 1];
        atomicMin(&g.minb, (unsigned long long)__double_as_longlong(b));       // b >= 0: the bit pattern orders like the value
    }
    gsync<GROUP>(group);
    const double minb = D > 0 ? __longlong_as_double((long long)g.minb) : 0.0;
    for (int j = gt; j < D; j += GROUP)
        fast[j] = make_float4((float)(g2 * (base[j] - minb)), (float)cf[j], (float)ll[j], 0.f);
    gsync<GROUP>(group);
    // Past exp(-600) the reference's own fp64 terms run into the denormal range / underflow to 0 (f == 0 is a decision
    // of its own): such far-away states are evaluated exactly.  Below it, every term that fp32 flushes after the shift
    // (< 2^-126 of the scale) is equally negligible in fp64.
    return kb.gamma * minb <= 600.0 && !kb.exact_only;
}

// ---------------------------------------------------------------------------------------------
template <int GROUP>
__global__ void __launch_bounds__(Cfg<GROUP>::THREADS) predict_kernel(const State kb, const float *__restrict__ state, int32_t *first_pos) {
    extern __shared__ __align__(16) unsigned char raw[];
    constexpr int GROUPS = Cfg<GROUP>::GROUPS;
    __shared__ GroupVars gv[GROUPS];
    const int group = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
    const int l = blockIdx.x * GROUPS + group;
    if (l >= kb.L) return;                                    // whole group
    double *base, *cf, *ll, *kf, *ds;
    float4 *fast;
    carve(raw, kb.cap, group, base, cf, ll, kf, ds, fast);
    GroupVars &g = gv[group];
    const int env = l / kb.S, s = l - env * kb.S, d = kb.dims[s];
    if (gt < d - 1) g.xs[gt] = (double)state[(size_t)env * kb.V + kb.offs[s] + gt];
    if (gt == 0) g.first = 1 << 30;
    gsync<GROUP>(group);
    const int D = kb.D[l];
    const bool fast_ok = stage_dictionary<GROUP>(kb, l, D, d, g, group, gt, base, cf, ll, fast);
    int first = 1 << 30;
    for (int a = gt; a <= kb.n_prbs; a += GROUP) {
        const double f = eval_f_guarded(D, kb.gamma, base, cf, ll, fast, fast_ok, (double)a / (double)kb.n_prbs, nullptr);
        if (f > 0.0) first = min(first, a);                   // prediction == +1 (kernel.py:25)
    }
    if (first != (1 << 30)) atomicMin(&g.first, first);
    gsync<GROUP>(group);
    if (gt == 0) first_pos[l] = g.first == (1 << 30) ? -1 : g.first;
}

// ---------------------------------------------------------------------------------------------
template <int GROUP>
__global__ void __launch_bounds__(Cfg<GROUP>::THREADS, GROUP == 256 ? 4 : 1) update_kernel(const State kb, const float *__restrict__ state,
                                                         const int32_t *__restrict__ action,
                                                         const int32_t *__restrict__ labels, int32_t *y_pred,
                                                         const Control ctl, int32_t *hits) {
    extern __shared__ __align__(16) unsigned char raw[];
    constexpr int GROUPS = Cfg<GROUP>::GROUPS;
    __shared__ GroupVars gv[GROUPS];
    const int group = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
    const int l = blockIdx.x * GROUPS + group;
    if (l >= kb.L) return;                                    // whole group
    double *base, *cf, *ll, *kf, *ds;
    float4 *fast;
    carve(raw, kb.cap, group, base, cf, ll, kf, ds, fast);
    GroupVars &g = gv[group];
    double *xs = g.xs;
    const int env = l / kb.S, s = l - env * kb.S, d = kb.dims[s], n = kb.n_prbs, cap = kb.cap;
    if (gt < d - 1) xs[gt] = (double)state[(size_t)env * kb.V + kb.offs[s] + gt];
    const int y = labels[l];
    const int a0 = min(max(action[l], 0), n);
    const int lo = y == 1 ? a0 : 0, hi = y == 1 ? n : a0;     // kbrl_control.py:103-112
    double *lm = kb.lm + (size_t)l * cap * MAX_DIM;
    double *coeff = kb.coeff + (size_t)l * cap;
    double *kinv = kb.kinv + (size_t)l * cap * cap;
    int D = kb.D[l];
    int cur = lo;
    bool first_round = true;
    unsigned n_updates = 0;
    gsync<GROUP>(group);
    // The dictionary is staged from HBM once; a round's update is mirrored into the staged copy (the state part of
    // a new landmark equals the current state exactly, so its base distance is 0 and becomes the new minimum).
    bool fast_ok = stage_dictionary<GROUP>(kb, l, D, d, g, group, gt, base, cf, ll, fast);
    const double g2 = kb.gamma * 1.4426950408889634;
    // E-learner inputs, loaded early so that their latency hides behind the first evaluation round
    const int ctl_margin = ctl.acc ? max(0, ctl.margins[l]) : 0;
    const int ctl_adjusted = ctl.acc ? ctl.adjusted[env] : 0;
    double acc_v = 0.0;
    if (ctl.acc && gt < n) acc_v = ctl.acc[(size_t)l * n + gt];
    while (cur <= hi) {
        if (gt == 0) g.first = 1 << 30;
        gsync<GROUP>(group);
        int first = 1 << 30;
        for (int a = cur + gt; a <= hi; a += GROUP) {
            const double f = eval_f_guarded(D, kb.gamma, base, cf, ll, fast, fast_ok, (double)a / (double)n, nullptr);
            if (first_round && a == a0) g.fa0 = f;
            if (f * (double)y <= 0.0) first = min(first, a);   // Projectron.update acts only on mistakes (projectron.py:40)
        }
        if (first != (1 << 30)) atomicMin(&g.first, first);
        gsync<GROUP>(group);
        if (first_round) {                                     // the predict of kbrl_control.py:89 (before any update)
            const int yp = D == 0 ? 0 : (g.fa0 > 0.0 ? 1 : -1);
            if (gt == 0 && y_pred) y_pred[l] = yp;
            first_round = false;
            if (ctl.acc) {                                     // E-learner part of update_control (kbrl_control.py:90-101)
                const bool hit = y == yp;
                const int margin = ctl_margin;
                const double om = 1.0 - ctl.alfa;
                double *acc = ctl.acc + (size_t)l * n;
                if (gt == 0) g.sf = 1 << 30;
                gsync<GROUP>(group);
                for (int i = gt; i < n; i += GROUP) {
                    double v = i == gt ? acc_v : acc[i];
                    if (yp == 1) {
                        if (!hit) { if (i <= margin) { v = om * v; acc[i] = v; } }            // same or less margin: same mistake
                        else if (i >= margin) { v = om * v + ctl.alfa; acc[i] = v; }          // same or more margin: same success
                    }
                    if (v > ctl.acc_lo) atomicMin(&g.sf, i);   // np.argmax(accuracies > accuracy_range[0]): first True, 0 if none
                }
                gsync<GROUP>(group);
                if (gt == 0) {
                    if (!ctl_adjusted) ctl.sec[l] = g.sf == (1 << 30) ? 0 : g.sf;
                    if (hits) hits[l] = hit ? 1 : 0;
                }
            }
        }
        const int astar = g.first;
        if (astar == (1 << 30)) break;
        // ---- Projectron.update at x = [s, astar / n] (projectron.py:41-60)
        const double xa = (double)astar / (double)n;
        ++n_updates;
        if (D <= 1) {                                          // float32 stage: Kinv, K_f, coeff are float32 arrays of length 1
            if (gt == 0) {
                float kfv = 0.f, ki = 0.f;
                if (D == 1) { const double t = ll[0] - xa; kfv = (float)exp(-kb.gamma * (base[0] + t * t)); ki = (float)kinv[0]; }
                const float dsv = ki * kfv;
                const double dot = (double)(float)(dsv * kfv);
                double delta = 1.0 - dot;
                if (delta < 0.0) delta = 0.0;
                if (delta <= kb.eta) coeff[0] = (double)(float)((float)coeff[0] + (float)y * dsv);
                else {
                    for (int i = 0; i < d - 1; ++i) lm[(size_t)D * MAX_DIM + i] = xs[i];
                    lm[(size_t)D * MAX_DIM + d - 1] = xa;
                    coeff[D] = (double)y;
                    if (D == 0) kinv[0] = (double)(float)(1.0 / 1.0);
                    else {
                        const double dse[2] = {(double)dsv, -1.0};
                        kinv[1] = 0.0; kinv[cap] = 0.0; kinv[cap + 1] = 0.0;
                        for (int i = 0; i < 2; ++i)
                            for (int j = 0; j < 2; ++j) kinv[(size_t)i * cap + j] += dse[i] * dse[j] / delta;
                    }
                    g.D = D + 1;
                }
                if (delta <= kb.eta) g.D = D;
            }
            gsync<GROUP>(group);
            D = g.D;
            fast_ok = stage_dictionary<GROUP>(kb, l, D, d, g, group, gt, base, cf, ll, fast);   // float32 stage: tiny, re-staged
        } else {
            for (int j = gt; j < D; j += GROUP) { const double t = ll[j] - xa; kf[j] = exp(-kb.gamma * (base[j] + t * t)); }
            gsync<GROUP>(group);
            for (int i = gt; i < D; i += GROUP) {      // d* = K^-1 k; column i == row i (symmetric), coalesced
                double acc = 0.0;
                for (int j = 0; j < D; ++j) acc += kinv[(size_t)j * cap + i] * kf[j];
                ds[i] = acc;
            }
            gsync<GROUP>(group);
            if (gt == 0) {                                   // delta = max(Kii - d* . k, 0), index order
                double dot = 0.0;
                for (int i = 0; i < D; ++i) dot += ds[i] * kf[i];
                double delta = 1.0 - dot;
                g.delta = delta < 0.0 ? 0.0 : delta;
            }
            gsync<GROUP>(group);
            const double delta = g.delta;
            if (delta <= kb.eta) {                                    // sv.update(y * d_star)
                for (int i = gt; i < D; i += GROUP) {
                    const double c = coeff[i] + (double)y * ds[i];
                    coeff[i] = c; cf[i] = c; fast[i].y = (float)c;
                }
            } else if (D < cap) {                                     // sv.extend / insert + rank-1 extension of K^-1
                if (gt == 0) {
                    for (int i = 0; i < d - 1; ++i) lm[(size_t)D * MAX_DIM + i] = xs[i];
                    lm[(size_t)D * MAX_DIM + d - 1] = xa;
                    coeff[D] = (double)y;
                    ds[D] = -1.0;
                    base[D] = 0.0; cf[D] = (double)y; ll[D] = xa;          // staged copy of the new landmark
                }
                for (int i = gt; i <= D; i += GROUP) { kinv[(size_t)i * cap + D] = 0.0; kinv[(size_t)D * cap + i] = 0.0; }
                gsync<GROUP>(group);
                const int m = D + 1;
                for (int idx = gt; idx < m * m; idx += GROUP) {
                    const int i = idx / m, j = idx - i * m;
                    kinv[(size_t)i * cap + j] += ds[i] * ds[j] / delta;
                }
                D = m;
                for (int j = gt; j < D; j += GROUP)                      // fp32 copy relative to the new minimum (0)
                    fast[j] = make_float4((float)(g2 * base[j]), (float)cf[j], (float)ll[j], 0.f);
                fast_ok = !kb.exact_only;
            } else if (gt == 0) kb.flags[l] |= KB_FLAG_DICT_CAP;
            gsync<GROUP>(group);
        }
        cur = astar + 1;
    }
    if (gt == 0) {
        kb.D[l] = D;
        if (n_updates) atomicAdd(kb.updates, (unsigned long long)n_updates);
    }
}

// Tail of KBRL_Control.select_action (kbrl_control.py:57-73) + adjust_action (:75-78); one thread per env.
__global__ void select_kernel(const State kb, const Control ctl, int32_t *action_out, int32_t *adjusted_out) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env * kb.S >= kb.L) return;
    const int n = kb.n_prbs, S = kb.S;
    int a[8], m[8], assigned = 0;
    for (int s = 0; s < S; ++s) {
        const int l = env * S + s, first = ctl.first[l];
        if (first >= 0) { a[s] = min(n, first + ctl.sec[l]); m[s] = a[s] - first; }
        else { a[s] = n; m[s] = 0; }                            // the scan ran out: l1_prbs = n_prbs, margin 0
        assigned += a[s];
    }
    const int adj = assigned > n;
    for (int s = 0; s < S; ++s) {
        const int l = env * S + s;
        int act = a[s], mar = m[s];
        if (adj) {
            const double p = (double)a[s] / (double)assigned;
            act = (int)floor((double)n * p);
            mar -= a[s] - act;
        }
        ctl.action[l] = act; ctl.margins[l] = mar;
        if (action_out) action_out[l] = act;
    }
    ctl.adjusted[env] = adj;
    if (adjusted_out) adjusted_out[env] = adj;
}

}  // namespace kb

// ================================================================================================= C ABI
struct kb_handle {
    kb_config cfg;
    kb::State st;
    kb::Control ctl;
    size_t smem_update, smem_predict;
    cudaStream_t stream;
    float *d_state;
    int32_t *d_action, *d_labels, *d_out;
    uint64_t launches;
};

// error text is shared with ranslice_cabi.cu through rs_set_error
extern "C" void rs_set_error(const char *msg);
namespace {
constexpr int UG = kb::Cfg<KB_GROUP_UPDATE>::GROUPS, PG = kb::Cfg<KB_GROUP_PREDICT>::GROUPS;
int kfail(int code, const std::string &m) { rs_set_error(m.c_str()); return code; }
}
#define KCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return kfail(RS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

int kb_create(const kb_config *cfg, const int32_t *dims, const int32_t *offsets, kb_handle **out) {
    if (!cfg || !dims || !offsets || !out) return kfail(RS_E_ARG, "null argument");
    if (cfg->abi_version != RS_ABI_VERSION) return kfail(RS_E_ARG, "abi_version mismatch");
    if (cfg->n_envs <= 0 || cfg->n_slices <= 0 || cfg->n_slices > 8 || cfg->n_prbs <= 0 || cfg->n_prbs + 1 > kb::MAX_CAND)
        return kfail(RS_E_ARG, "bad n_envs / n_slices / n_prbs");
    const int cap = cfg->dict_cap ? cfg->dict_cap : 256;
    if (cap < 2 || cap > 1024) return kfail(RS_E_ARG, "dict_cap must be in [2, 1024]");
    for (int s = 0; s < cfg->n_slices; ++s) {
        // separable distance needs the action coordinate to be added last by numpy's pairwise sum: len(x) != 8, < 16
        if (dims[s] < 2 || dims[s] == 8 || dims[s] >= kb::MAX_DIM) return kfail(RS_E_ARG, "unsupported x dimension (need 2..15, != 8)");
        if (offsets[s] < 0 || offsets[s] + dims[s] - 1 > cfg->n_variables) return kfail(RS_E_ARG, "state slice out of range");
    }
    int ndev = 0;
    KCU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return kfail(RS_E_ARG, "bad device ordinal");
    KCU(cudaSetDevice(cfg->device));
    kb_handle *h = new kb_handle();
    h->cfg = *cfg; h->cfg.dict_cap = cap; h->launches = 0;
    h->ctl = kb::Control{};
    kb::State &st = h->st;
    st.L = cfg->n_envs * cfg->n_slices; st.S = cfg->n_slices; st.V = cfg->n_variables; st.n_prbs = cfg->n_prbs; st.cap = cap;
    st.gamma = cfg->gamma; st.eta = cfg->eta; st.exact_only = 0;
    for (int s = 0; s < 8; ++s) { st.dims[s] = s < cfg->n_slices ? dims[s] : 0; st.offs[s] = s < cfg->n_slices ? offsets[s] : 0; }
    const size_t L = (size_t)st.L;
    KCU(cudaMalloc(&st.D, L * sizeof(int)));
    KCU(cudaMalloc(&st.lm, L * cap * kb::MAX_DIM * sizeof(double)));
    KCU(cudaMalloc(&st.coeff, L * cap * sizeof(double)));
    KCU(cudaMalloc(&st.kinv, L * cap * cap * sizeof(double)));
    KCU(cudaMalloc(&st.flags, L * sizeof(uint32_t)));
    KCU(cudaMalloc(&st.updates, sizeof(unsigned long long)));
    KCU(cudaMalloc(&h->d_state, (size_t)cfg->n_envs * cfg->n_variables * sizeof(float)));
    KCU(cudaMalloc(&h->d_action, L * sizeof(int32_t)));
    KCU(cudaMalloc(&h->d_labels, L * sizeof(int32_t)));
    KCU(cudaMalloc(&h->d_out, L * sizeof(int32_t)));
    KCU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->smem_update = (size_t)kb::Cfg<KB_GROUP_UPDATE>::GROUPS * cap * kb::GROUP_SMEM_PER_LANDMARK;
    h->smem_predict = (size_t)kb::Cfg<KB_GROUP_PREDICT>::GROUPS * cap * kb::GROUP_SMEM_PER_LANDMARK;
    KCU(cudaFuncSetAttribute(kb::update_kernel<KB_GROUP_UPDATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_update));
    KCU(cudaFuncSetAttribute(kb::predict_kernel<KB_GROUP_PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_predict));
    *out = h;
    return kb_reset(h);
}

int kb_reset(kb_handle *h) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaMemset(h->st.D, 0, (size_t)h->st.L * sizeof(int)));
    KCU(cudaMemset(h->st.flags, 0, (size_t)h->st.L * sizeof(uint32_t)));
    KCU(cudaMemset(h->st.updates, 0, sizeof(unsigned long long)));
    return RS_OK;
}

int kb_destroy(kb_handle *h) {
    if (!h) return RS_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(h->st.D); cudaFree(h->st.lm); cudaFree(h->st.coeff); cudaFree(h->st.kinv); cudaFree(h->st.flags);
    cudaFree(h->ctl.acc); cudaFree(h->ctl.sec); cudaFree(h->ctl.margins); cudaFree(h->ctl.adjusted); cudaFree(h->ctl.action);
    cudaFree(h->ctl.first);
    cudaFree(h->st.updates); cudaFree(h->d_state); cudaFree(h->d_action); cudaFree(h->d_labels); cudaFree(h->d_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return RS_OK;
}

int kb_update_device(kb_handle *h, const float *d_state, const int32_t *d_action, const int32_t *d_labels,
                     int32_t *d_y_pred, void *stream) {
    if (!h || !d_state || !d_action || !d_labels || !d_y_pred) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    KCU(cudaMemsetAsync(h->st.updates, 0, sizeof(unsigned long long), st));
    kb::update_kernel<KB_GROUP_UPDATE><<<(h->st.L + UG - 1) / UG, kb::Cfg<KB_GROUP_UPDATE>::THREADS, h->smem_update, st>>>(h->st, d_state, d_action, d_labels, d_y_pred, kb::Control{}, nullptr);
    h->launches += 1;
    KCU(cudaGetLastError());
    return RS_OK;
}

int kb_control_init(kb_handle *h, const int32_t *initial_action, const int32_t *security_factor, double alfa,
                    double accuracy_lo, double accuracy_hi) {
    if (!h || !initial_action || !security_factor) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    const size_t L = (size_t)h->st.L, n = (size_t)h->st.n_prbs;
    kb::Control &c = h->ctl;
    if (!c.acc) {
        KCU(cudaMalloc(&c.acc, L * n * sizeof(double)));
        KCU(cudaMalloc(&c.sec, L * sizeof(int32_t)));
        KCU(cudaMalloc(&c.margins, L * sizeof(int32_t)));
        KCU(cudaMalloc(&c.adjusted, (size_t)h->cfg.n_envs * sizeof(int32_t)));
        KCU(cudaMalloc(&c.action, L * sizeof(int32_t)));
        KCU(cudaMalloc(&c.first, L * sizeof(int32_t)));
    }
    c.alfa = alfa; c.acc_lo = accuracy_lo;
    std::vector<double> acc(L * n, (accuracy_lo + accuracy_hi) / 2);               // kbrl_control.py:38-39
    KCU(cudaMemcpy(c.acc, acc.data(), acc.size() * sizeof(double), cudaMemcpyHostToDevice));
    KCU(cudaMemcpy(c.sec, security_factor, L * sizeof(int32_t), cudaMemcpyHostToDevice));
    KCU(cudaMemcpy(c.action, initial_action, L * sizeof(int32_t), cudaMemcpyHostToDevice));
    KCU(cudaMemset(c.margins, 0, L * sizeof(int32_t)));
    KCU(cudaMemset(c.adjusted, 0, (size_t)h->cfg.n_envs * sizeof(int32_t)));
    return RS_OK;
}

int kb_control_update_device(kb_handle *h, const float *d_state, const int32_t *d_action, const int32_t *d_labels,
                             int32_t *d_hits, void *stream) {
    if (!h || !d_state || !d_action || !d_labels) return kfail(RS_E_ARG, "null argument");
    if (!h->ctl.acc) return kfail(RS_E_ARG, "kb_control_init has not been called");
    KCU(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    KCU(cudaMemsetAsync(h->st.updates, 0, sizeof(unsigned long long), st));
    kb::update_kernel<KB_GROUP_UPDATE><<<(h->st.L + UG - 1) / UG, kb::Cfg<KB_GROUP_UPDATE>::THREADS, h->smem_update, st>>>(h->st, d_state, d_action, d_labels, nullptr, h->ctl, d_hits);
    h->launches += 1;
    KCU(cudaGetLastError());
    return RS_OK;
}

int kb_control_select_device(kb_handle *h, const float *d_state, int32_t *d_action, int32_t *d_adjusted, void *stream) {
    if (!h || !d_state) return kfail(RS_E_ARG, "null argument");
    if (!h->ctl.acc) return kfail(RS_E_ARG, "kb_control_init has not been called");
    KCU(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    kb::predict_kernel<KB_GROUP_PREDICT><<<(h->st.L + PG - 1) / PG, kb::Cfg<KB_GROUP_PREDICT>::THREADS, h->smem_predict, st>>>(h->st, d_state, h->ctl.first);
    kb::select_kernel<<<(h->cfg.n_envs + 127) / 128, 128, 0, st>>>(h->st, h->ctl, d_action, d_adjusted);
    h->launches += 2;
    KCU(cudaGetLastError());
    return RS_OK;
}

int kb_control_get(kb_handle *h, int32_t *action, int32_t *security_factors, int32_t *margins, int32_t *adjusted,
                   double *accuracies) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    if (!h->ctl.acc) return kfail(RS_E_ARG, "kb_control_init has not been called");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    const size_t L = (size_t)h->st.L;
    if (action) KCU(cudaMemcpy(action, h->ctl.action, L * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (security_factors) KCU(cudaMemcpy(security_factors, h->ctl.sec, L * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (margins) KCU(cudaMemcpy(margins, h->ctl.margins, L * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (adjusted) KCU(cudaMemcpy(adjusted, h->ctl.adjusted, (size_t)h->cfg.n_envs * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (accuracies) KCU(cudaMemcpy(accuracies, h->ctl.acc, L * h->st.n_prbs * sizeof(double), cudaMemcpyDeviceToHost));
    return RS_OK;
}

int kb_predict_device(kb_handle *h, const float *d_state, int32_t *d_first_pos, void *stream) {
    if (!h || !d_state || !d_first_pos) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    kb::predict_kernel<KB_GROUP_PREDICT><<<(h->st.L + PG - 1) / PG, kb::Cfg<KB_GROUP_PREDICT>::THREADS, h->smem_predict, (cudaStream_t)stream>>>(h->st, d_state, d_first_pos);
    h->launches += 1;
    KCU(cudaGetLastError());
    return RS_OK;
}

int kb_update(kb_handle *h, const float *state, const int32_t *action, const int32_t *labels, int32_t *y_pred) {
    if (!h || !state || !action || !labels || !y_pred) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    const size_t L = (size_t)h->st.L;
    KCU(cudaMemcpyAsync(h->d_state, state, (size_t)h->cfg.n_envs * h->cfg.n_variables * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    KCU(cudaMemcpyAsync(h->d_action, action, L * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    KCU(cudaMemcpyAsync(h->d_labels, labels, L * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int rc = kb_update_device(h, h->d_state, h->d_action, h->d_labels, h->d_out, h->stream);
    if (rc) return rc;
    KCU(cudaMemcpyAsync(y_pred, h->d_out, L * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    KCU(cudaStreamSynchronize(h->stream));
    return RS_OK;
}

int kb_predict(kb_handle *h, const float *state, int32_t *first_pos) {
    if (!h || !state || !first_pos) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    const size_t L = (size_t)h->st.L;
    KCU(cudaMemcpyAsync(h->d_state, state, (size_t)h->cfg.n_envs * h->cfg.n_variables * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    int rc = kb_predict_device(h, h->d_state, h->d_out, h->stream);
    if (rc) return rc;
    KCU(cudaMemcpyAsync(first_pos, h->d_out, L * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    KCU(cudaStreamSynchronize(h->stream));
    return RS_OK;
}

int kb_get_sizes(kb_handle *h, int32_t *sizes, uint32_t *flags) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    if (sizes) KCU(cudaMemcpy(sizes, h->st.D, (size_t)h->st.L * sizeof(int), cudaMemcpyDeviceToHost));
    if (flags) KCU(cudaMemcpy(flags, h->st.flags, (size_t)h->st.L * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return RS_OK;
}

int kb_get_learner(kb_handle *h, int32_t l, double *landmarks, double *coeff, double *kinv, int32_t *D_out) {
    if (!h || l < 0 || l >= h->st.L) return kfail(RS_E_ARG, "bad handle / learner");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    int D = 0;
    KCU(cudaMemcpy(&D, h->st.D + l, sizeof(int), cudaMemcpyDeviceToHost));
    const int cap = h->st.cap, d = h->st.dims[l % h->st.S];
    if (landmarks && D) {
        std::vector<double> tmp((size_t)D * kb::MAX_DIM);
        KCU(cudaMemcpy(tmp.data(), h->st.lm + (size_t)l * cap * kb::MAX_DIM, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < D; ++i) for (int j = 0; j < d; ++j) landmarks[(size_t)i * d + j] = tmp[(size_t)i * kb::MAX_DIM + j];
    }
    if (coeff && D) KCU(cudaMemcpy(coeff, h->st.coeff + (size_t)l * cap, (size_t)D * sizeof(double), cudaMemcpyDeviceToHost));
    if (kinv && D)
        KCU(cudaMemcpy2D(kinv, (size_t)D * sizeof(double), h->st.kinv + (size_t)l * cap * cap, (size_t)cap * sizeof(double),
                         (size_t)D * sizeof(double), D, cudaMemcpyDeviceToHost));
    if (D_out) *D_out = D;
    return RS_OK;
}

#ifdef KB_CHECK
int kb_debug_counters(unsigned long long *out4) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out4, kb::g_kb_dbg, 4 * sizeof(unsigned long long)) == cudaSuccess ? 0 : -1;
}
#endif

int kb_set_exact(kb_handle *h, int on) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    h->st.exact_only = on ? 1 : 0;
    return RS_OK;
}

int kb_get_counters(kb_handle *h, uint64_t *kernel_launches, uint64_t *updates_last_call) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    KCU(cudaSetDevice(h->cfg.device));
    if (kernel_launches) *kernel_launches = h->launches;
    if (updates_last_call) {
        KCU(cudaDeviceSynchronize());
        unsigned long long v = 0;
        KCU(cudaMemcpy(&v, h->st.updates, sizeof v, cudaMemcpyDeviceToHost));
        *updates_last_call = v;
    }
    return RS_OK;
}

}  // extern "C"
