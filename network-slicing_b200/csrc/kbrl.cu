// K3: KBRL inner loop -- Gaussian-kernel evaluation and Projectron dictionary projection / update,
// batched over L = n_envs * n_slices independent learners (C ABI: include/kbrl_b200.h).
//
//   GaussianKernel.k_eval / k / predict   algorithms/kernel.py:8-28   (incl. the random tie-break at f == 0, :26-27)
//   SVvariable.extend / update / insert   algorithms/projectron.py:3-21  (unbounded growing arrays)
//   Projectron.predict / update           algorithms/projectron.py:32-60
//   ProjectronPlus.update                 algorithms/projectron.py:66-107 (kb_config.algorithm = 1)
//   callers restated: the select_action scan (kbrl_control.py:54-61) and the sample-augmentation loop of
//   update_control (kbrl_control.py:103-112)
//
// Both callers evaluate f(a) = sum_j coeff_j exp(-gamma ||l_j - [s, a/n]||^2) for up to n_prbs + 1 candidate
// allocations a of one state s.  numpy's pairwise sum over the 11 (4) squared differences adds the action coordinate
// LAST, so the distance is separable bit-exactly: base_j (state part, once per landmark) + (l_j,last - a/n)^2.
// Threads own candidates and walk the landmarks (broadcast from shared memory) in index order; a dictionary update
// is a block-cooperative K^-1 k mat-vec and, when the sample is not well approximated, a rank-1 extension of K^-1.
//
// DICTIONARY STORAGE (round 2).  The reference's dictionaries are unbounded (np.append / np.vstack); its own runs
// reach several hundred landmarks within 3000 steps and > 1000 in the stored 50 400-step experiments.  A dense
// K^-1[cap][cap] per learner cannot hold that (cap 1024 x 81 920 learners = 687 GB).  Each learner therefore GROWS in
// rows of 32 landmarks taken from one device pool by a bump allocator (an atomicAdd inside the update kernel; nothing
// is ever copied or freed until kb_reset):
//
//     tile row I of a learner = [ landmarks 32 x 16 | coeff 32 | K^-1 tiles (I,0) (I,1) ... (I,I), 32 x 32 each ]
//
// K^-1 is exactly symmetric, so only the lower block triangle is stored (diagonal tiles hold both triangles):
// D(D+1)/2 doubles rounded up to tiles instead of cap^2 -- 0.37 MB at D = 289 instead of 8 MB at cap 1024 -- and the
// mat-vec / rank-1 extension read and write half the bytes.  Element (i, j), i >= j: row[i >> 5] + 544 + (j >> 5) 1024
// + (i & 31) 32 + (j & 31).  d* = K^-1 k: thread i walks j in index order (the order of the oracle): row i of the
// tiles left of the diagonal (a 256-byte run per tile and lane, sector-efficient through L1), then column i of the
// diagonal tile and of the tiles below it (coalesced across the lanes of a warp).
//
// The Gram / K^-1 products stay on CUDA cores in fp64: the only matrix-shaped work is a D x D mat-VEC (0.25 flop per
// byte, HBM-bound from D ~ 100 on) and f(a) is one exp per (landmark, candidate) pair on a separable distance -- there
// is no contraction for tensor cores to take (measured: tools/kb_crossover.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/kbrl_b200.h"
#include "../../include/ranslice_b200.h"
#include "philox.cuh"

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
#ifndef KB_GROUP_PREDICT_BIG
#define KB_GROUP_PREDICT_BIG 128
#endif
#ifndef KB_GROUP_UPDATE_BIG
#define KB_GROUP_UPDATE_BIG 1024     // large dictionaries: the K^-1 mat-vec / extension stream megabytes, one CTA per SM pulls them with every lane
#endif
constexpr size_t GROUP_SMEM_PER_LANDMARK = 5 * sizeof(double) + sizeof(float4);   // base, coeff, last coord, k, d*, fp32 copy
template <int G> struct Cfg {
    static_assert(G == 32 || G == 64 || G == 128 || G == 256 || G == 512 || G == 1024, "threads per learner: 32 .. 1024");
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
constexpr int TILE = 32;                               // landmarks per tile row
constexpr int TILE_ELEMS = TILE * TILE;                // doubles per K^-1 tile
constexpr int ROW_HDR = TILE * MAX_DIM + TILE;         // doubles ahead of the tiles of a row: landmarks [32][16], coeff [32]
constexpr int MAX_CAP = 2048;                          // landmarks per learner the staging in shared memory can hold
constexpr int SMALL_CAP = 256;                         // staging capacity of the first (small-dictionary) update launch
constexpr int SMALL_CAP_PREDICT = 128;                 // dictionaries up to this size are scanned by one warp each
constexpr int MID_CAP_PREDICT = 512;                   // ... up to this size by a 128-thread group with a 20 KB staging area
__host__ __device__ inline unsigned long long row_doubles(int I) { return (unsigned long long)ROW_HDR + (unsigned long long)(I + 1) * TILE_ELEMS; }

constexpr int PEND_FRESH = -1, PEND_DONE = -2;         // State::pend: >= 0 = resume the augmentation loop at this allocation

struct State {
    int L, S, V, n_prbs, cap, tmax;
    int exact_only;    // kb_set_exact: evaluate every f in fp64 like the reference (validation of the guarded fast path)
    int plus;          // ProjectronPlus.update (projectron.py:66-107) instead of Projectron.update
    double gamma, eta;
    int dims[8], offs[8];
    int *D;                          // [L]
    unsigned long long *rows;        // [L][tmax] pool offset (in doubles) of tile row I; rows (D + 31) / 32 .. are not allocated
    double *pool;
    unsigned long long pool_doubles;
    unsigned long long *cursor;      // [1] bump allocator
    uint32_t *flags;                 // [L]
    unsigned long long *updates;     // [1]
    int *pend;                       // [L] hand-over between the two update launches of a step (small / large dictionaries)
    int *big_list;                   // [L] learners handed to the large-dictionary update launch of this step ...
    int *big_ctl;                    // [3] ... {entries from the front (largest dictionaries: taken first), work cursor of that (persistent)
                                     //      launch, entries from the back}
    int *pred_list;                  // [2][L] learners of the select_action scan with a mid-size / large dictionary ...
    int *pred_ctl;                   // [4] ... {mid count, mid cursor, large count, large cursor}
    uint32_t *tie_ctr;               // [L] draws taken from the learner's tie-break stream
    int *max_d;                      // [1] largest dictionary seen
    uint32_t k0, k1, env0;           // tie-break stream: Philox key = seed, counter = (n, STREAM_KBRL, slice, env0 + env)
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

// shared-memory slice of one group: tile-row pointers of its learner, then the staged dictionary
constexpr size_t PREDICT_SMEM_PER_LANDMARK = 3 * sizeof(double) + sizeof(float4);  // the scan needs no k / d* planes
template <bool UPDATE>
__device__ __forceinline__ void carve(unsigned char *raw, int capS, int tmax, int group, double **&rowp, double *&base, double *&cf,
                                      double *&ll, double *&kf, double *&ds, float4 *&fast) {
    const size_t rowp_bytes = ((size_t)tmax * sizeof(double *) + 15) & ~size_t(15);     // the float4 plane behind it needs 16-byte alignment
    unsigned char *p0 = raw + (size_t)group * (rowp_bytes + (size_t)capS * (UPDATE ? GROUP_SMEM_PER_LANDMARK : PREDICT_SMEM_PER_LANDMARK));
    rowp = reinterpret_cast<double **>(p0);
    double *p = reinterpret_cast<double *>(p0 + rowp_bytes);
    base = p; cf = base + capS; ll = cf + capS;
    kf = UPDATE ? ll + capS : nullptr; ds = UPDATE ? kf + capS : nullptr;
    fast = reinterpret_cast<float4 *>(UPDATE ? ds + capS : ll + capS);   // [capS] fp32 copy of the staged dictionary (guarded fast path)
}
// per-group scalars
struct GroupVars { double xs[MAX_DIM]; double delta, fa0; unsigned long long minb; int first, D, sf, ok; };

__device__ __forceinline__ double *lm_ptr(double *const *rowp, int j) { return rowp[j >> 5] + (j & 31) * MAX_DIM; }
__device__ __forceinline__ double *cf_ptr(double *const *rowp, int j) { return rowp[j >> 5] + TILE * MAX_DIM + (j & 31); }
__device__ __forceinline__ double *tile_ptr(double *const *rowp, int I, int J) { return rowp[I] + ROW_HDR + (size_t)J * TILE_ELEMS; }

// np.random.choice([-1, 1]) of GaussianKernel.predict (kernel.py:26-27) from the learner's Philox stream (philox.py: choice =
// seq[integers(len(seq))], one counter tick)
__device__ __forceinline__ int tie_draw(const State &kb, int env, int s, uint32_t n) {
    rs::PhiloxStream r{kb.k0, kb.k1, (uint32_t)s, rs::STREAM_KBRL, n, kb.env0 + (uint32_t)env};
    return r.integers(2u) ? 1 : -1;
}

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

// Guarded fp32 evaluation of f(a).  Only the SIGN of f is ever used by Projectron (kernel.py:25, projectron.py:40), so f
// is first summed in fp32 with ex2.approx together with G = sum |coeff_j| k_j; the fp32 result differs from the
// reference's fp64 sum by less than (1e-5 + 1.2e-7 D) G  (|arg| <= 126 before a term flushes to zero: argument error
// <= 5e-7 + 1.2e-7 |arg|, ex2.approx 2^-22, D sequential fp32 additions); the guard (4e-5 + 3e-7 D) G leaves a factor
// 2.5-4.  Inside that band -- a candidate sitting on the decision boundary, or f == 0 exactly (the tie-break) -- f is
// re-evaluated exactly like the reference (eval_f).  Dictionaries of 0 or 1 landmarks (the reference's float32 stage)
// go straight to eval_f.
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

// The same evaluation shared by P adjacent lanes (wide groups: a 1024-thread group has five threads per candidate): lane
// `part` sums the landmarks j = part (mod P) in fp32, the partial (f, G) are added across the P lanes, and the guard test
// is made on the totals (its bound covers any summation order of the fp32 terms).  Inside the band -- and for dictionaries
// of 0 or 1 landmarks -- lane 0 of the P evaluates exactly, in the reference's order, and the others take its value.
template <int P>
__device__ __forceinline__ double eval_f_guarded_split(int D, double gamma, const double *base, const double *cf, const double *ll,
                                                       const float4 *fast, bool fast_ok, double xa, int part) {
    float f = 0.f, G = 0.f;
    const bool try_fast = D > 1 && fast_ok;
    if (try_fast) {
        const float xf = (float)xa, g2 = (float)gamma * LOG2E;
        for (int j = part; j < D; j += P) {
            const float4 e = fast[j];
            const float t = e.z - xf;
            const float term = e.y * ex2f(-fmaf(t * t, g2, e.x));
            f += term;
            G += fabsf(term);
        }
    }
#pragma unroll
    for (int d = 1; d < P; d <<= 1) { f += __shfl_xor_sync(0xffffffffu, f, d); G += __shfl_xor_sync(0xffffffffu, G, d); }
    double r = (double)f;
    const bool exact = !try_fast || !(fabsf(f) > (4e-5f + 3e-7f * (float)D) * G);
    if (exact && part == 0) r = eval_f(D, gamma, base, cf, ll, xa);
    int lo = __double2loint(r), hi = __double2hiint(r);          // (unconditional: every lane of the warp takes part in the shuffle)
    lo = __shfl_sync(0xffffffffu, lo, 0, P); hi = __shfl_sync(0xffffffffu, hi, 0, P);
    return __hiloint2double(hi, lo);
}

// Stages the dictionary of a learner for state xs: exact fp64 (base, coeff, last coordinate) and the fp32 copy of the
// fast path.  The fp32 exponents are taken RELATIVE to the closest landmark (min_j base_j): sign(f) does not change
// when every term is scaled by exp(gamma min_j base_j), and without the shift a state far from all landmarks puts
// every term below the fp32 normal range (measured: flushed terms flipped 140 of 1.1e9 decisions).
// Group-wide; ends with the dictionary visible to all threads of the group (callers need no further barrier for it).
// Returns whether the fp32 fast path may be used for this (dictionary, state).
template <int GROUP>
__device__ bool stage_dictionary(const State &kb, double *const *rowp, int D, int d, GroupVars &g, int group, int gt, double *base,
                                 double *cf, double *ll, float4 *fast) {
    const double g2 = kb.gamma * 1.4426950408889634;
    if (gt == 0) g.minb = ~0ull;
    gsync<GROUP>(group);
    for (int j = gt; j < D; j += GROUP) {
        const double *lm = lm_ptr(rowp, j);
        const double b = base_dist(lm, g.xs, d - 1);
        base[j] = b; cf[j] = *cf_ptr(rowp, j); ll[j] = lm[d - 1];
        atomicMin(&g.minb, (unsigned long long)__double_as_longlong(b));       // b >= 0: the bit pattern orders like the value
    }
    gsync<GROUP>(group);
    const double minb = D > 0 ? __longlong_as_double((long long)g.minb) : 0.0;
    for (int j = gt; j < D; j += GROUP)
        fast[j] = make_float4((float)(g2 * (base[j] - minb)), (float)cf[j], (float)ll[j], 0.f);
    gsync<GROUP>(group);
    // Past exp(-600) the reference's own fp64 terms run into the denormal range / underflow to 0 (f == 0 is a decision
    // of its own: the tie-break): such far-away states are evaluated exactly.  Below it, every term that fp32 flushes
    // after the shift (< 2^-126 of the scale) is equally negligible in fp64.  ProjectronPlus needs the VALUE of f.
    return kb.gamma * minb <= 600.0 && !kb.exact_only && !kb.plus;
}

// tile-row pointers of learner l into shared memory (rows 0 .. (D + 31) / 32 - 1 exist)
template <int GROUP>
__device__ __forceinline__ void load_rows(const State &kb, int l, int D, double **rowp, int gt) {
    const int T = (D + TILE - 1) >> 5;
    for (int I = gt; I < T; I += GROUP) rowp[I] = kb.pool + kb.rows[(size_t)l * kb.tmax + I];
}

// ---------------------------------------------------------------------------------------------
// select_action scan (kbrl_control.py:54-61) of one learner: first allocation whose prediction is +1.  A candidate with
// f == 0 exactly draws its prediction from the tie-break stream, in scan order.
template <int GROUP>
__device__ void predict_learner(const State &kb, const int capS, const int l, const int D, const float *__restrict__ state,
                                int32_t *first_pos, unsigned char *raw, GroupVars &g, const int group, const int gt) {
    double **rowp, *base, *cf, *ll, *kf, *ds;
    float4 *fast;
    carve<false>(raw, capS, kb.tmax, group, rowp, base, cf, ll, kf, ds, fast);
    const int env = l / kb.S, s = l - env * kb.S, d = kb.dims[s];
    if (gt < d - 1) g.xs[gt] = (double)state[(size_t)env * kb.V + kb.offs[s] + gt];
    if (gt == 0) g.first = 1 << 30;
    load_rows<GROUP>(kb, l, D, rowp, gt);
    gsync<GROUP>(group);
    const bool fast_ok = stage_dictionary<GROUP>(kb, rowp, D, d, g, group, gt, base, cf, ll, fast);
    // The scan wants the FIRST allocation predicted +1, so the candidates are taken in ascending passes of GROUP and the scan
    // stops after the pass that found one (a dictionary of D landmarks costs D ex2 per candidate; f is mostly increasing in
    // the allocation, so the hit usually comes early).  key: a << 1 | (f > 0); ties (f == 0, D > 0) enter with bit 0 clear.
    int first = 1 << 30;
    constexpr int P = GROUP >= 128 ? 4 : 1;                   // wide groups (larger dictionaries): four threads per candidate, shorter passes
    for (int a0p = 0; a0p <= kb.n_prbs; a0p += GROUP / P) {
        const int a = a0p + gt / P;
        if (P > 1) {                                          // (whole warps stay in the shuffles of the split evaluation)
            const double f = eval_f_guarded_split<P>(a <= kb.n_prbs ? D : 0, kb.gamma, base, cf, ll, fast, fast_ok, (double)a / (double)kb.n_prbs, gt % P);
            if (a <= kb.n_prbs && gt % P == 0) {
                if (f > 0.0) first = (a << 1) | 1;
                else if (f == 0.0 && D > 0) first = a << 1;
            }
        } else if (a <= kb.n_prbs) {
            const double f = eval_f_guarded(D, kb.gamma, base, cf, ll, fast, fast_ok, (double)a / (double)kb.n_prbs, nullptr);
            if (f > 0.0) first = (a << 1) | 1;                // prediction == +1 (kernel.py:25)
            else if (f == 0.0 && D > 0) first = a << 1;
        }
        if (GROUP == 32) {                                    // warp: no shared memory needed
            first = (int)__reduce_min_sync(0xffffffffu, (unsigned)first);
            if (first != (1 << 30)) break;
        } else {
            if (first != (1 << 30)) atomicMin(&g.first, first);
            gsync<GROUP>(group);
            const bool found = g.first != (1 << 30);
            gsync<GROUP>(group);                              // (everybody has read the flag before the next pass may change it)
            if (found) break;
        }
    }
    if (GROUP == 32 && gt == 0) g.first = first;
    gsync<GROUP>(group);
    if (gt == 0) {
        int key = g.first, res = -1;
        if (key != (1 << 30)) {
            if (key & 1) res = key >> 1;
            else {                                            // a tie before the first +1: replay the scan from there in order
                uint32_t tn = kb.tie_ctr[l];
                for (int a = key >> 1; a <= kb.n_prbs; ++a) {
                    const double f = eval_f(D, kb.gamma, base, cf, ll, (double)a / (double)kb.n_prbs);
                    if (f > 0.0 || (f == 0.0 && tie_draw(kb, env, s, tn++) == 1)) { res = a; break; }
                }
                kb.tie_ctr[l] = tn;
            }
        }
        first_pos[l] = res;
    }
}

// Three launches per scan, by dictionary size: one warp per learner up to SMALL_CAP_PREDICT landmarks (most learners for
// the first ~1000 steps); the others are listed by that kernel and scanned by two persistent launches with wider groups
// and larger staging areas (mid: up to MID_CAP_PREDICT, large: up to dict_cap).
template <int GROUP>
__global__ void __launch_bounds__(Cfg<GROUP>::THREADS) predict_kernel(const State kb, const int capS, const float *__restrict__ state,
                                                                      int32_t *first_pos) {
    extern __shared__ __align__(16) unsigned char raw[];
    constexpr int GROUPS = Cfg<GROUP>::GROUPS;
    __shared__ GroupVars gv[GROUPS];
    const int group = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
    const int l = blockIdx.x * GROUPS + group;
    if (l >= kb.L) return;                                    // whole group
    const int D = kb.D[l];
    if (D > capS) {                                           // listed for the mid / large launch
        if (gt == 0) {
            const int which = D > MID_CAP_PREDICT;
            kb.pred_list[(size_t)which * kb.L + atomicAdd(&kb.pred_ctl[2 * which], 1)] = l;
        }
        return;
    }
    predict_learner<GROUP>(kb, capS, l, D, state, first_pos, raw, gv[group], group, gt);
}
template <int GROUP>
__global__ void __launch_bounds__(GROUP) predict_list_kernel(const State kb, const int capS, const int which,
                                                             const float *__restrict__ state, int32_t *first_pos) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ GroupVars g;
    __shared__ int s_idx;
    const int *list = kb.pred_list + (size_t)which * kb.L;
    int *ctl = kb.pred_ctl + 2 * which;
    const int count = ctl[0];
    for (;;) {
        if (threadIdx.x == 0) s_idx = atomicAdd(&ctl[1], 1);
        __syncthreads();
        const int idx = s_idx;
        __syncthreads();
        if (idx >= count) return;
        const int l = list[idx];
        predict_learner<GROUP>(kb, capS, l, kb.D[l], state, first_pos, raw, g, 0, threadIdx.x);
        __syncthreads();
    }
}

// d* = K^-1 k for the staged k (kf): thread i sums row i of the symmetric matrix in index order -- row i of the tiles
// left of the diagonal, then column i of the diagonal tile and of the tiles below it (K^-1[i][j] == K^-1[j][i]).
template <int GROUP>
__device__ __forceinline__ void kinv_matvec(double *const *rowp, int D, const double *kf, double *ds, int gt) {
    for (int i = gt; i < D; i += GROUP) {
        const int I = i >> 5, r = i & 31;
        double acc = 0.0;
        const double *rowI = rowp[I] + ROW_HDR + r * TILE;
        for (int J = 0; J < I; ++J) {
            const double2 *p = reinterpret_cast<const double2 *>(rowI + (size_t)J * TILE_ELEMS);
            const double *k = kf + J * TILE;
#pragma unroll 8
            for (int c = 0; c < TILE / 2; ++c) {
                const double2 v = p[c];
                acc += v.x * k[2 * c];
                acc += v.y * k[2 * c + 1];
            }
        }
        for (int J = I; J * TILE < D; ++J) {
            const double *q = rowp[J] + ROW_HDR + (size_t)I * TILE_ELEMS + r;
            const double *k = kf + J * TILE;
            const int nj = min(TILE, D - J * TILE);
#pragma unroll 8
            for (int jr = 0; jr < nj; ++jr) acc += q[jr * TILE] * k[jr];
        }
        ds[i] = acc;
    }
}

// K^-1 <- [[K^-1, 0], [0, 0]] + [d*; -1][d*; -1]^T / delta over the stored tiles (projectron.py:54-58); ds[D] == -1.
// Row D and column D are new: their old value is 0 (never read).  Diagonal tiles update both triangles with the same
// commutative products, so the matrix stays exactly symmetric.
template <int GROUP>
__device__ __forceinline__ void kinv_extend(double *const *rowp, int D, const double *ds, double delta, int gt) {
    const int m = D + 1, Tn = (m + TILE - 1) >> 5;
    for (int I = 0; I < Tn; ++I) {
        const int ni = min(TILE, m - I * TILE);
        for (int J = 0; J <= I; ++J) {
            double *t = tile_ptr(rowp, I, J);
            const int nj = min(TILE, m - J * TILE);
            for (int e = gt; e < ni * TILE; e += GROUP) {
                const int r = e >> 5, c = e & 31;
                if (c >= nj) continue;
                const int i = I * TILE + r, j = J * TILE + c;
                const double add = ds[i] * ds[j] / delta;
                const double old = (i == D || j == D) ? 0.0 : t[e];
                t[e] = old + add;
            }
        }
    }
}

// hands learner l to the large-dictionary launch.  That launch is persistent and its entries are very uneven (a 1000-landmark
// dictionary that errs on most of its candidates streams gigabytes), so the largest dictionaries are listed from the front,
// which is taken first, and the rest from the back.
constexpr int BIG_FIRST = 640;
__device__ __forceinline__ void list_big(const State &kb, int l, int D) {
    if (D >= BIG_FIRST) kb.big_list[atomicAdd(&kb.big_ctl[0], 1)] = l;
    else kb.big_list[kb.L - 1 - atomicAdd(&kb.big_ctl[2], 1)] = l;
}

// bump allocation of tile row I of learner l (thread 0 of the group); returns nullptr when the pool is exhausted
__device__ __forceinline__ double *alloc_row(const State &kb, int l, int I) {
    const unsigned long long need = row_doubles(I);
    const unsigned long long off = atomicAdd(kb.cursor, need);
    if (off + need > kb.pool_doubles) {
        atomicAdd(kb.cursor, 0ull - need);
        return nullptr;
    }
    kb.rows[(size_t)l * kb.tmax + I] = off;
    return kb.pool + off;
}

// ---------------------------------------------------------------------------------------------
// Projectron part of update_control (+ the E-learner when ctl.acc != nullptr) for one learner.  Two launches per step:
// the first (hand_over = true) stages at most capS = SMALL_CAP landmarks (14 KB of shared memory per learner, 256 threads)
// and lists a learner whose dictionary has reached that size for the second, which starts it (pend == PEND_FRESH) or
// resumes its augmentation loop at allocation `pend` with capS = dict_cap and 1024 threads.
template <int GROUP>
__device__ void update_learner(const State &kb, const int capS, const bool hand_over, const int l, const int pend,
                               const float *__restrict__ state, const int32_t *__restrict__ action,
                               const int32_t *__restrict__ labels, int32_t *y_pred, const Control &ctl, int32_t *hits,
                               unsigned char *raw, GroupVars &g, const int group, const int gt) {
    int D = kb.D[l];
    if (hand_over && D >= capS) {                             // large dictionary: the second launch takes it from the start
        if (gt == 0) { kb.pend[l] = PEND_FRESH; list_big(kb, l, D); }
        return;
    }
    double **rowp, *base, *cf, *ll, *kf, *ds;
    float4 *fast;
    carve<true>(raw, capS, kb.tmax, group, rowp, base, cf, ll, kf, ds, fast);
    double *xs = g.xs;
    const int env = l / kb.S, s = l - env * kb.S, d = kb.dims[s], n = kb.n_prbs;
    if (gt < d - 1) xs[gt] = (double)state[(size_t)env * kb.V + kb.offs[s] + gt];
    const int y = labels[l];
    const int a0 = min(max(action[l], 0), n);
    const int lo = y == 1 ? a0 : 0, hi = y == 1 ? n : a0;     // kbrl_control.py:103-112
    int cur = pend >= 0 ? pend : lo;
    bool first_round = pend < 0;
    unsigned n_updates = 0;
    uint32_t tie_n = kb.tie_ctr[l];
    load_rows<GROUP>(kb, l, D, rowp, gt);
    gsync<GROUP>(group);
    // The dictionary is staged from HBM once; a round's update is mirrored into the staged copy (the state part of
    // a new landmark equals the current state exactly, so its base distance is 0 and becomes the new minimum).
    bool fast_ok = stage_dictionary<GROUP>(kb, rowp, D, d, g, group, gt, base, cf, ll, fast);
    const double g2 = kb.gamma * 1.4426950408889634;
    // E-learner inputs, loaded early so that their latency hides behind the first evaluation round
    const int ctl_margin = ctl.acc ? max(0, ctl.margins[l]) : 0;
    const int ctl_adjusted = ctl.acc ? ctl.adjusted[env] : 0;
    double acc_v = 0.0;
    if (ctl.acc && gt < n) acc_v = ctl.acc[(size_t)l * n + gt];
    int handed_over = -1;
    while (cur <= hi) {
        if (gt == 0) g.first = 1 << 30;
        gsync<GROUP>(group);
        int first = 1 << 30;                                   // key: a << 1 | (f != 0)
        if (GROUP >= 1024) {                                   // at most 201 candidates: four threads each (whole warps stay in the loop)
            constexpr int P = 4;
            for (int a0w = cur; a0w <= hi; a0w += GROUP / P) {
                const int a = a0w + gt / P;
                const double f = eval_f_guarded_split<P>(a <= hi ? D : 0, kb.gamma, base, cf, ll, fast, fast_ok, (double)a / (double)n, gt % P);
                if (a <= hi && gt % P == 0) {
                    if (first_round && a == a0) g.fa0 = f;
                    const bool act = kb.plus ? f * (double)y < 1.0 : f * (double)y <= 0.0;
                    if (act) first = min(first, (a << 1) | (f != 0.0));
                }
            }
        } else
        for (int a = cur + gt; a <= hi; a += GROUP) {
            const double f = eval_f_guarded(D, kb.gamma, base, cf, ll, fast, fast_ok, (double)a / (double)n, nullptr);
            if (first_round && a == a0) g.fa0 = f;
            // Projectron.update acts on mistakes (f y <= 0, projectron.py:40); ProjectronPlus also inside the margin (:70-72)
            const bool act = kb.plus ? f * (double)y < 1.0 : f * (double)y <= 0.0;
            if (act) first = min(first, (a << 1) | (f != 0.0));
        }
        if (first != (1 << 30)) atomicMin(&g.first, first);
        gsync<GROUP>(group);
        if (first_round) {                                     // the predict of kbrl_control.py:89 (before any update)
            int yp = 0;
            if (D > 0) {
                yp = g.fa0 > 0.0 ? 1 : -1;
                if (g.fa0 == 0.0) yp = tie_draw(kb, env, s, tie_n++);      // kernel.py:26-27 (every thread: same draw)
            }
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
        const int key = g.first;
        if (key == (1 << 30)) break;
        const int astar = key >> 1;
        if (hand_over && D >= capS) { handed_over = astar; break; }   // staging is full: the second launch resumes at astar
        // ---- Projectron.update at x = [s, astar / n] (projectron.py:41-60); its predict drew a tie-break if f == 0
        const double xa = (double)astar / (double)n;
        ++n_updates;
        if (D > 0 && !(key & 1)) ++tie_n;
        if (D <= 1) {                                          // float32 stage: Kinv, K_f, coeff are float32 arrays of length 1
            if (gt == 0) {
                float kfv = 0.f, ki = 0.f, f32 = 0.f;
                if (D == 1) {
                    const double t = ll[0] - xa;
                    kfv = (float)exp(-kb.gamma * (base[0] + t * t)); ki = (float)tile_ptr(rowp, 0, 0)[0];
                    f32 = kfv * (float)cf[0];
                }
                const float dsv = ki * kfv;
                const double dot = (double)(float)(dsv * kfv);
                double delta = 1.0 - dot;
                if (delta < 0.0) delta = 0.0;
                const double margin = (double)y * (double)f32;                 // y * self.f (np.int64 * np.float32 -> float64)
                int newD = D;
                if (kb.plus && margin > 0.0) {                                 // ProjectronPlus inside the margin (:72-84); never reached with D == 0 (f = 0)
                    const double loss = 1.0 - margin;
                    double norm_xt = 1.0 - delta;
                    if (norm_xt < 0.0) norm_xt = 0.0;
                    if (loss - delta / kb.eta > 0.0) {
                        const double alpha = fmin(fmin(loss / norm_xt, 1.0), 2.0 * (loss - delta / kb.eta) / norm_xt);
                        // coeff (float32[1]) += alpha * y * d_star: the product is float64 (numpy scalars are not weak), the in-place add rounds to float32
                        *cf_ptr(rowp, 0) = (double)(float)(cf[0] + (alpha * (double)y) * (double)dsv);
                    }
                } else if (delta <= kb.eta) *cf_ptr(rowp, 0) = (double)(float)((float)cf[0] + (float)y * dsv);
                else {
                    double *row0 = D == 0 ? alloc_row(kb, l, 0) : rowp[0];
                    if (!row0) atomicOr(&kb.flags[l], (uint32_t)KB_FLAG_POOL);
                    else {
                        rowp[0] = row0;
                        double *lmn = lm_ptr(rowp, D);
                        for (int i = 0; i < d - 1; ++i) lmn[i] = xs[i];
                        lmn[d - 1] = xa;
                        *cf_ptr(rowp, D) = (double)y;
                        double *t00 = tile_ptr(rowp, 0, 0);
                        if (D == 0) t00[0] = (double)(float)(1.0 / 1.0);
                        else {
                            const double dse[2] = {(double)dsv, -1.0};
                            t00[1] = 0.0; t00[TILE] = 0.0; t00[TILE + 1] = 0.0;
                            for (int i = 0; i < 2; ++i)
                                for (int j = 0; j < 2; ++j) t00[i * TILE + j] += dse[i] * dse[j] / delta;
                        }
                        newD = D + 1;
                    }
                }
                g.D = newD;
            }
            gsync<GROUP>(group);
            D = g.D;
            fast_ok = stage_dictionary<GROUP>(kb, rowp, D, d, g, group, gt, base, cf, ll, fast);   // float32 stage: tiny, re-staged
        } else {
            for (int j = gt; j < D; j += GROUP) { const double t = ll[j] - xa; kf[j] = exp(-kb.gamma * (base[j] + t * t)); }
            gsync<GROUP>(group);
            kinv_matvec<GROUP>(rowp, D, kf, ds, gt);
            gsync<GROUP>(group);
            // delta = max(Kii - d* . k, 0): 32 strided partial sums + a shuffle tree in the group's first warp (a serial sum by one
            // thread was 40 % of an update at D = 500; numpy's own dot is a BLAS ddot with several SIMD accumulators)
            double dot = 0.0;
            if (gt < 32) {
                for (int i = gt; i < D; i += 32) dot += ds[i] * kf[i];
#pragma unroll
                for (int d = 16; d; d >>= 1) {
                    int lo = __double2loint(dot), hi = __double2hiint(dot);
                    lo = __shfl_xor_sync(0xffffffffu, lo, d); hi = __shfl_xor_sync(0xffffffffu, hi, d);
                    dot += __hiloint2double(hi, lo);
                }
            }
            if (gt == 0) {                                   // (f for ProjectronPlus)
                double delta = 1.0 - dot;
                g.delta = delta < 0.0 ? 0.0 : delta;
                g.ok = 1;
                if (kb.plus) {
                    double f = 0.0;
                    for (int j = 0; j < D; ++j) f += kf[j] * cf[j];
                    g.fa0 = f;                                // (fa0 is free after the first round)
                }
            }
            gsync<GROUP>(group);
            const double delta = g.delta;
            const double margin = kb.plus ? (double)y * g.fa0 : 0.0;
            if (kb.plus && margin > 0.0) {                            // ProjectronPlus inside the margin (projectron.py:72-84)
                const double loss = 1.0 - margin;
                double norm_xt = 1.0 - delta;
                if (norm_xt < 0.0) norm_xt = 0.0;
                if (loss - delta / kb.eta > 0.0) {
                    const double alpha = fmin(fmin(loss / norm_xt, 1.0), 2.0 * (loss - delta / kb.eta) / norm_xt);
                    const double ay = alpha * (double)y;
                    for (int i = gt; i < D; i += GROUP) {
                        const double c = cf[i] + ay * ds[i];
                        *cf_ptr(rowp, i) = c; cf[i] = c; fast[i].y = (float)c;
                    }
                }
            } else if (delta <= kb.eta) {                             // sv.update(y * d_star)
                for (int i = gt; i < D; i += GROUP) {
                    const double c = cf[i] + (double)y * ds[i];
                    *cf_ptr(rowp, i) = c; cf[i] = c; fast[i].y = (float)c;
                }
            } else if (D < kb.cap) {                                  // sv.extend / insert + rank-1 extension of K^-1
                if ((D & (TILE - 1)) == 0) {                          // the new landmark opens a tile row
                    if (gt == 0) {
                        double *row = alloc_row(kb, l, D >> 5);
                        if (row) rowp[D >> 5] = row; else { g.ok = 0; atomicOr(&kb.flags[l], (uint32_t)KB_FLAG_POOL); }
                    }
                    gsync<GROUP>(group);
                }
                if (g.ok) {
                    if (gt == 0) {
                        double *lmn = lm_ptr(rowp, D);
                        for (int i = 0; i < d - 1; ++i) lmn[i] = xs[i];
                        lmn[d - 1] = xa;
                        *cf_ptr(rowp, D) = (double)y;
                        ds[D] = -1.0;
                        base[D] = 0.0; cf[D] = (double)y; ll[D] = xa;          // staged copy of the new landmark
                    }
                    gsync<GROUP>(group);
                    kinv_extend<GROUP>(rowp, D, ds, delta, gt);
                    D += 1;
                    for (int j = gt; j < D; j += GROUP)                      // fp32 copy relative to the new minimum (0)
                        fast[j] = make_float4((float)(g2 * base[j]), (float)cf[j], (float)ll[j], 0.f);
                    fast_ok = !kb.exact_only && !kb.plus;
                }
            } else if (gt == 0) atomicOr(&kb.flags[l], (uint32_t)KB_FLAG_DICT_CAP);
            gsync<GROUP>(group);
        }
        cur = astar + 1;
    }
    if (gt == 0) {
        kb.D[l] = D;
        kb.tie_ctr[l] = tie_n;
        if (handed_over >= 0) { kb.pend[l] = handed_over; list_big(kb, l, D); }
        if (n_updates) atomicAdd(kb.updates, (unsigned long long)n_updates);
        if (D > *kb.max_d) atomicMax(kb.max_d, D);
    }
}

template <int GROUP>
__global__ void __launch_bounds__(Cfg<GROUP>::THREADS, GROUP == 256 ? 4 : 1) update_kernel(const State kb, const int capS, const int hand_over,
                                                         const float *__restrict__ state,
                                                         const int32_t *__restrict__ action,
                                                         const int32_t *__restrict__ labels, int32_t *y_pred,
                                                         const Control ctl, int32_t *hits) {
    extern __shared__ __align__(16) unsigned char raw[];
    constexpr int GROUPS = Cfg<GROUP>::GROUPS;
    __shared__ GroupVars gv[GROUPS];
    const int group = threadIdx.x / GROUP, gt = threadIdx.x % GROUP;
    const int l = blockIdx.x * GROUPS + group;
    if (l >= kb.L) return;                                    // whole group
    update_learner<GROUP>(kb, capS, hand_over != 0, l, PEND_FRESH, state, action, labels, y_pred, ctl, hits, raw, gv[group], group, gt);
}
// the learners listed by update_kernel, one CTA at a time (persistent: the list is short and its entries are uneven)
template <int GROUP>
__global__ void __launch_bounds__(GROUP, 1024 / GROUP) update_list_kernel(const State kb, const int capS, const float *__restrict__ state,
                                                               const int32_t *__restrict__ action, const int32_t *__restrict__ labels,
                                                               int32_t *y_pred, const Control ctl, int32_t *hits) {
    extern __shared__ __align__(16) unsigned char raw[];
    __shared__ GroupVars g;
    __shared__ int s_idx;
    const int front = kb.big_ctl[0], count = front + kb.big_ctl[2];
    for (;;) {
        if (threadIdx.x == 0) s_idx = atomicAdd(&kb.big_ctl[1], 1);
        __syncthreads();
        const int idx = s_idx;
        __syncthreads();
        if (idx >= count) return;
        const int l = kb.big_list[idx < front ? idx : kb.L - 1 - (idx - front)];
        update_learner<GROUP>(kb, capS, false, l, kb.pend[l], state, action, labels, y_pred, ctl, hits, raw, g, 0, threadIdx.x);
        __syncthreads();
    }
}

// dense copy of one learner's dictionary (kb_get_learner): landmarks [D][MAX_DIM], coeff [D], kinv [D][D]
__global__ void gather_kernel(const State kb, int l, double *lm, double *coeff, double *kinv) {
    const int D = kb.D[l], tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const unsigned long long *rows = kb.rows + (size_t)l * kb.tmax;
    for (int idx = tid; idx < D * D; idx += nt) {
        const int a = idx / D, b = idx - a * D, i = max(a, b), j = min(a, b);
        kinv[idx] = kb.pool[rows[i >> 5] + ROW_HDR + (size_t)(j >> 5) * TILE_ELEMS + (i & 31) * TILE + (j & 31)];
    }
    for (int idx = tid; idx < D * MAX_DIM; idx += nt) {
        const int q = idx / MAX_DIM;
        lm[idx] = kb.pool[rows[q >> 5] + (q & 31) * MAX_DIM + (idx - q * MAX_DIM)];
    }
    for (int idx = tid; idx < D; idx += nt) coeff[idx] = kb.pool[rows[idx >> 5] + TILE * MAX_DIM + (idx & 31)];
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
    int cap_small;                           // staging capacity of the small-dictionary update launch
    int sm_count;
    size_t smem_update_small, smem_update_big, smem_predict_small, smem_predict_mid, smem_predict_big;
    cudaStream_t stream;
    float *d_state;
    int32_t *d_action, *d_labels, *d_out;
    double *d_gather;                        // kb_get_learner scratch (grown on demand)
    size_t gather_doubles;
    uint64_t launches;
};

// error text is shared with ranslice_cabi.cu through rs_set_error
extern "C" void rs_set_error(const char *msg);
namespace {
constexpr int UG = kb::Cfg<KB_GROUP_UPDATE>::GROUPS, PG = kb::Cfg<KB_GROUP_PREDICT>::GROUPS;
constexpr int PREDICT_MID_GROUP = KB_GROUP_PREDICT_BIG, PREDICT_BIG_GROUP = 256;
int kfail(int code, const std::string &m) { rs_set_error(m.c_str()); return code; }
}
#define KCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return kfail(RS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

static int kb_create_impl(kb_handle *h, const kb_config *cfg, const int32_t *dims, const int32_t *offsets) {
    const int cap = h->cfg.dict_cap;
    kb::State &st = h->st;
    st.L = cfg->n_envs * cfg->n_slices; st.S = cfg->n_slices; st.V = cfg->n_variables; st.n_prbs = cfg->n_prbs;
    st.cap = cap; st.tmax = cap / kb::TILE;
    st.gamma = cfg->gamma; st.eta = cfg->eta; st.exact_only = 0; st.plus = cfg->algorithm == 1;
    st.k0 = (uint32_t)cfg->tie_seed; st.k1 = (uint32_t)(cfg->tie_seed >> 32); st.env0 = (uint32_t)cfg->first_env_id;
    for (int s = 0; s < 8; ++s) { st.dims[s] = s < cfg->n_slices ? dims[s] : 0; st.offs[s] = s < cfg->n_slices ? offsets[s] : 0; }
    const size_t L = (size_t)st.L;
    KCU(cudaMalloc(&st.D, L * sizeof(int)));
    KCU(cudaMalloc(&st.rows, L * st.tmax * sizeof(unsigned long long)));
    KCU(cudaMalloc(&st.flags, L * sizeof(uint32_t)));
    KCU(cudaMalloc(&st.pend, L * sizeof(int)));
    KCU(cudaMalloc(&st.tie_ctr, L * sizeof(uint32_t)));
    KCU(cudaMalloc(&st.updates, sizeof(unsigned long long)));
    KCU(cudaMalloc(&st.cursor, sizeof(unsigned long long)));
    KCU(cudaMalloc(&st.max_d, sizeof(int)));
    KCU(cudaMalloc(&st.big_list, L * sizeof(int)));
    KCU(cudaMalloc(&st.big_ctl, 3 * sizeof(int)));
    KCU(cudaMalloc(&st.pred_list, 2 * L * sizeof(int)));
    KCU(cudaMalloc(&st.pred_ctl, 4 * sizeof(int)));
    KCU(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    {   // dictionary pool: what every learner would need at full size, capped by pool_mb or (by default) 70 % of the free memory
        unsigned long long full = 0;
        for (int I = 0; I < st.tmax; ++I) full += kb::row_doubles(I);
        full *= L;
        size_t free_b = 0, total_b = 0;
        KCU(cudaMemGetInfo(&free_b, &total_b));
        unsigned long long want = cfg->pool_mb > 0 ? (unsigned long long)cfg->pool_mb * (1ull << 20) / sizeof(double)
                                                   : std::min<unsigned long long>(full, (unsigned long long)(0.7 * (double)free_b) / sizeof(double));
        want = std::max<unsigned long long>(std::min(want, full), kb::row_doubles(0));
        cudaError_t e = cudaMalloc(&st.pool, want * sizeof(double));
        if (e != cudaSuccess) return kfail(RS_E_NOMEM, "dictionary pool of " + std::to_string(want * sizeof(double) >> 20) + " MB: " + cudaGetErrorString(e));
        st.pool_doubles = want;
    }
    KCU(cudaMalloc(&h->d_state, (size_t)cfg->n_envs * cfg->n_variables * sizeof(float)));
    KCU(cudaMalloc(&h->d_action, L * sizeof(int32_t)));
    KCU(cudaMalloc(&h->d_labels, L * sizeof(int32_t)));
    KCU(cudaMalloc(&h->d_out, L * sizeof(int32_t)));
    KCU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->cap_small = std::min(cap, kb::SMALL_CAP);
    const size_t rowp = ((size_t)st.tmax * sizeof(double *) + 15) & ~size_t(15);
    h->smem_update_small = (size_t)UG * (rowp + (size_t)h->cap_small * kb::GROUP_SMEM_PER_LANDMARK);
    h->smem_update_big = rowp + (size_t)cap * kb::GROUP_SMEM_PER_LANDMARK;
    h->smem_predict_small = (size_t)PG * (rowp + (size_t)std::min(cap, kb::SMALL_CAP_PREDICT) * kb::PREDICT_SMEM_PER_LANDMARK);
    h->smem_predict_mid = rowp + (size_t)std::min(cap, kb::MID_CAP_PREDICT) * kb::PREDICT_SMEM_PER_LANDMARK;
    h->smem_predict_big = rowp + (size_t)cap * kb::PREDICT_SMEM_PER_LANDMARK;
    KCU(cudaFuncSetAttribute(kb::update_kernel<KB_GROUP_UPDATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_update_small));
    KCU(cudaFuncSetAttribute(kb::update_list_kernel<KB_GROUP_UPDATE_BIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_update_big));
    KCU(cudaFuncSetAttribute(kb::predict_kernel<KB_GROUP_PREDICT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_predict_small));
    KCU(cudaFuncSetAttribute(kb::predict_list_kernel<PREDICT_MID_GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_predict_mid));
    KCU(cudaFuncSetAttribute(kb::predict_list_kernel<PREDICT_BIG_GROUP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_predict_big));
    return kb_reset(h);
}

extern "C" {

int kb_create(const kb_config *cfg, const int32_t *dims, const int32_t *offsets, kb_handle **out) {
    if (!cfg || !dims || !offsets || !out) return kfail(RS_E_ARG, "null argument");
    if (cfg->abi_version != RS_ABI_VERSION) return kfail(RS_E_ARG, "abi_version mismatch");
    if (cfg->n_envs <= 0 || cfg->n_slices <= 0 || cfg->n_slices > 8 || cfg->n_prbs <= 0 || cfg->n_prbs + 1 > kb::MAX_CAND)
        return kfail(RS_E_ARG, "bad n_envs / n_slices / n_prbs");
    if (cfg->algorithm != 0 && cfg->algorithm != 1) return kfail(RS_E_ARG, "algorithm must be 0 (Projectron) or 1 (ProjectronPlus)");
    if (cfg->pool_mb < 0) return kfail(RS_E_ARG, "pool_mb must be >= 0");
    int cap = cfg->dict_cap ? cfg->dict_cap : 1024;
    if (cap < 2 || cap > kb::MAX_CAP) return kfail(RS_E_ARG, "dict_cap must be in [2, 2048]");
    cap = (cap + kb::TILE - 1) / kb::TILE * kb::TILE;                      // whole tile rows
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
    h->st = kb::State{};
    const int rc = kb_create_impl(h, cfg, dims, offsets);
    if (rc != RS_OK) { kb_destroy(h); return rc; }                         // frees whatever had been allocated
    *out = h;
    return RS_OK;
}

int kb_reset(kb_handle *h) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    KCU(cudaMemset(h->st.D, 0, (size_t)h->st.L * sizeof(int)));
    KCU(cudaMemset(h->st.flags, 0, (size_t)h->st.L * sizeof(uint32_t)));
    KCU(cudaMemset(h->st.tie_ctr, 0, (size_t)h->st.L * sizeof(uint32_t)));
    KCU(cudaMemset(h->st.updates, 0, sizeof(unsigned long long)));
    KCU(cudaMemset(h->st.cursor, 0, sizeof(unsigned long long)));          // the whole pool is free again
    KCU(cudaMemset(h->st.max_d, 0, sizeof(int)));
    return RS_OK;
}

int kb_destroy(kb_handle *h) {
    if (!h) return RS_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(h->st.D); cudaFree(h->st.rows); cudaFree(h->st.pool); cudaFree(h->st.flags); cudaFree(h->st.pend); cudaFree(h->st.tie_ctr);
    cudaFree(h->st.cursor); cudaFree(h->st.max_d); cudaFree(h->st.big_list); cudaFree(h->st.big_ctl); cudaFree(h->st.pred_list);
    cudaFree(h->st.pred_ctl);
    cudaFree(h->ctl.acc); cudaFree(h->ctl.sec); cudaFree(h->ctl.margins); cudaFree(h->ctl.adjusted); cudaFree(h->ctl.action);
    cudaFree(h->ctl.first);
    cudaFree(h->st.updates); cudaFree(h->d_state); cudaFree(h->d_action); cudaFree(h->d_labels); cudaFree(h->d_out); cudaFree(h->d_gather);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return RS_OK;
}

// the two launches of an update step: dictionaries below cap_small with a small staging area (it lists the others and
// those that outgrow it), then a persistent launch over that list with dict_cap staging and 1024 threads per learner
static int launch_update(kb_handle *h, const float *d_state, const int32_t *d_action, const int32_t *d_labels, int32_t *d_y_pred,
                         const kb::Control &ctl, int32_t *d_hits, cudaStream_t st) {
    KCU(cudaMemsetAsync(h->st.updates, 0, sizeof(unsigned long long), st));
    KCU(cudaMemsetAsync(h->st.big_ctl, 0, 3 * sizeof(int), st));
    const int blocks = (h->st.L + UG - 1) / UG, threads = kb::Cfg<KB_GROUP_UPDATE>::THREADS;
    const int two = h->st.cap > h->cap_small;                  // (with dict_cap <= SMALL_CAP the first launch does everything)
    kb::update_kernel<KB_GROUP_UPDATE><<<blocks, threads, h->smem_update_small, st>>>(h->st, h->cap_small, two, d_state, d_action, d_labels, d_y_pred, ctl, d_hits);
    h->launches += 1;
    if (two) {
        kb::update_list_kernel<KB_GROUP_UPDATE_BIG><<<h->sm_count * (1024 / KB_GROUP_UPDATE_BIG), KB_GROUP_UPDATE_BIG, h->smem_update_big, st>>>(h->st, h->st.cap, d_state, d_action, d_labels, d_y_pred, ctl, d_hits);
        h->launches += 1;
    }
    KCU(cudaGetLastError());
    return RS_OK;
}

static int launch_predict(kb_handle *h, const float *d_state, int32_t *d_first, cudaStream_t st) {
    const int cs = std::min(h->st.cap, kb::SMALL_CAP_PREDICT), cm = std::min(h->st.cap, kb::MID_CAP_PREDICT);
    KCU(cudaMemsetAsync(h->st.pred_ctl, 0, 4 * sizeof(int), st));
    kb::predict_kernel<KB_GROUP_PREDICT><<<(h->st.L + PG - 1) / PG, kb::Cfg<KB_GROUP_PREDICT>::THREADS, h->smem_predict_small, st>>>(h->st, cs, d_state, d_first);
    h->launches += 1;
    if (h->st.cap > cs) {
        kb::predict_list_kernel<PREDICT_MID_GROUP><<<8 * h->sm_count, PREDICT_MID_GROUP, h->smem_predict_mid, st>>>(h->st, cm, 0, d_state, d_first);
        h->launches += 1;
    }
    if (h->st.cap > cm) {
        kb::predict_list_kernel<PREDICT_BIG_GROUP><<<2 * h->sm_count, PREDICT_BIG_GROUP, h->smem_predict_big, st>>>(h->st, h->st.cap, 1, d_state, d_first);
        h->launches += 1;
    }
    KCU(cudaGetLastError());
    return RS_OK;
}

int kb_update_device(kb_handle *h, const float *d_state, const int32_t *d_action, const int32_t *d_labels,
                     int32_t *d_y_pred, void *stream) {
    if (!h || !d_state || !d_action || !d_labels || !d_y_pred) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    return launch_update(h, d_state, d_action, d_labels, d_y_pred, kb::Control{}, nullptr, (cudaStream_t)stream);
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
    return launch_update(h, d_state, d_action, d_labels, nullptr, h->ctl, d_hits, (cudaStream_t)stream);
}

int kb_control_select_device(kb_handle *h, const float *d_state, int32_t *d_action, int32_t *d_adjusted, void *stream) {
    if (!h || !d_state) return kfail(RS_E_ARG, "null argument");
    if (!h->ctl.acc) return kfail(RS_E_ARG, "kb_control_init has not been called");
    KCU(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = launch_predict(h, d_state, h->ctl.first, st);
    if (rc) return rc;
    kb::select_kernel<<<(h->cfg.n_envs + 127) / 128, 128, 0, st>>>(h->st, h->ctl, d_action, d_adjusted);
    h->launches += 1;
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
    return launch_predict(h, d_state, d_first_pos, (cudaStream_t)stream);
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

int kb_get_pool(kb_handle *h, uint64_t *used_bytes, uint64_t *total_bytes, int32_t *max_dictionary, uint64_t *tie_breaks) {
    if (!h) return kfail(RS_E_ARG, "null handle");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    unsigned long long cur = 0;
    int md = 0;
    KCU(cudaMemcpy(&cur, h->st.cursor, sizeof cur, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(&md, h->st.max_d, sizeof md, cudaMemcpyDeviceToHost));
    if (used_bytes) *used_bytes = cur * sizeof(double);
    if (total_bytes) *total_bytes = h->st.pool_doubles * sizeof(double);
    if (max_dictionary) *max_dictionary = md;
    if (tie_breaks) {
        std::vector<uint32_t> t((size_t)h->st.L);
        KCU(cudaMemcpy(t.data(), h->st.tie_ctr, t.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        uint64_t sum = 0;
        for (uint32_t v : t) sum += v;
        *tie_breaks = sum;
    }
    return RS_OK;
}

int kb_get_learner(kb_handle *h, int32_t l, double *landmarks, double *coeff, double *kinv, int32_t *D_out) {
    if (!h || l < 0 || l >= h->st.L) return kfail(RS_E_ARG, "bad handle / learner");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    int D = 0;
    KCU(cudaMemcpy(&D, h->st.D + l, sizeof(int), cudaMemcpyDeviceToHost));
    const int d = h->st.dims[l % h->st.S];
    if (D && (landmarks || coeff || kinv)) {
        const size_t need = (size_t)D * D + (size_t)D * kb::MAX_DIM + (size_t)D;
        if (need > h->gather_doubles) {
            cudaFree(h->d_gather); h->d_gather = nullptr; h->gather_doubles = 0;
            KCU(cudaMalloc(&h->d_gather, need * sizeof(double)));
            h->gather_doubles = need;
        }
        double *g_kinv = h->d_gather, *g_lm = g_kinv + (size_t)D * D, *g_cf = g_lm + (size_t)D * kb::MAX_DIM;
        kb::gather_kernel<<<64, 256>>>(h->st, l, g_lm, g_cf, g_kinv);
        KCU(cudaGetLastError());
        KCU(cudaDeviceSynchronize());
        if (landmarks) {
            std::vector<double> tmp((size_t)D * kb::MAX_DIM);
            KCU(cudaMemcpy(tmp.data(), g_lm, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (int i = 0; i < D; ++i) for (int j = 0; j < d; ++j) landmarks[(size_t)i * d + j] = tmp[(size_t)i * kb::MAX_DIM + j];
        }
        if (coeff) KCU(cudaMemcpy(coeff, g_cf, (size_t)D * sizeof(double), cudaMemcpyDeviceToHost));
        if (kinv) KCU(cudaMemcpy(kinv, g_kinv, (size_t)D * D * sizeof(double), cudaMemcpyDeviceToHost));
    }
    if (D_out) *D_out = D;
    return RS_OK;
}

// ---- checkpoint / restore (state_dict-style dump): dictionaries (only the part of the pool in use), per-learner
// bookkeeping and, when kb_control_init has been called, the controller state.  Blob layout: KbBlobHdr, D[L], rows[L][tmax],
// flags[L], tie_ctr[L], pool[cursor], then acc[L][n_prbs], sec[L], margins[L], action[L], adjusted[N].
struct KbBlobHdr { uint64_t magic, L, tmax, n_prbs, n_envs, cursor, has_ctl, max_d; };
static const uint64_t KB_BLOB_MAGIC = 0x4B42524C32303236ull;   // "KBRL2026"
static size_t kb_blob_bytes(const kb_handle *h, unsigned long long cursor) {
    const size_t L = (size_t)h->st.L;
    size_t b = sizeof(KbBlobHdr) + L * sizeof(int) + L * h->st.tmax * sizeof(unsigned long long) + 2 * L * sizeof(uint32_t) + (size_t)cursor * sizeof(double);
    if (h->ctl.acc) b += L * h->st.n_prbs * sizeof(double) + 3 * L * sizeof(int32_t) + (size_t)h->cfg.n_envs * sizeof(int32_t);
    return b;
}
int kb_state_size(kb_handle *h, size_t *bytes) {
    if (!h || !bytes) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    unsigned long long cur = 0;
    KCU(cudaMemcpy(&cur, h->st.cursor, sizeof cur, cudaMemcpyDeviceToHost));
    *bytes = kb_blob_bytes(h, cur);
    return RS_OK;
}
int kb_get_state(kb_handle *h, void *blob, size_t bytes) {
    if (!h || !blob) return kfail(RS_E_ARG, "null argument");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    unsigned long long cur = 0;
    int md = 0;
    KCU(cudaMemcpy(&cur, h->st.cursor, sizeof cur, cudaMemcpyDeviceToHost));
    KCU(cudaMemcpy(&md, h->st.max_d, sizeof md, cudaMemcpyDeviceToHost));
    if (bytes != kb_blob_bytes(h, cur)) return kfail(RS_E_ARG, "bad blob size (call kb_state_size first)");
    const size_t L = (size_t)h->st.L;
    char *p = static_cast<char *>(blob);
    const KbBlobHdr hdr{KB_BLOB_MAGIC, (uint64_t)L, (uint64_t)h->st.tmax, (uint64_t)h->st.n_prbs, (uint64_t)h->cfg.n_envs, cur, h->ctl.acc ? 1ull : 0ull, (uint64_t)md};
    std::memcpy(p, &hdr, sizeof hdr); p += sizeof hdr;
    auto out = [&](const void *d, size_t n) -> cudaError_t { cudaError_t e = cudaMemcpy(p, d, n, cudaMemcpyDeviceToHost); p += n; return e; };
    KCU(out(h->st.D, L * sizeof(int)));
    KCU(out(h->st.rows, L * h->st.tmax * sizeof(unsigned long long)));
    KCU(out(h->st.flags, L * sizeof(uint32_t)));
    KCU(out(h->st.tie_ctr, L * sizeof(uint32_t)));
    KCU(out(h->st.pool, (size_t)cur * sizeof(double)));
    if (h->ctl.acc) {
        KCU(out(h->ctl.acc, L * h->st.n_prbs * sizeof(double)));
        KCU(out(h->ctl.sec, L * sizeof(int32_t)));
        KCU(out(h->ctl.margins, L * sizeof(int32_t)));
        KCU(out(h->ctl.action, L * sizeof(int32_t)));
        KCU(out(h->ctl.adjusted, (size_t)h->cfg.n_envs * sizeof(int32_t)));
    }
    return RS_OK;
}
int kb_set_state(kb_handle *h, const void *blob, size_t bytes) {
    if (!h || !blob || bytes < sizeof(KbBlobHdr)) return kfail(RS_E_ARG, "null argument / short blob");
    KCU(cudaSetDevice(h->cfg.device));
    KCU(cudaDeviceSynchronize());
    KbBlobHdr hdr;
    std::memcpy(&hdr, blob, sizeof hdr);
    const size_t L = (size_t)h->st.L;
    if (hdr.magic != KB_BLOB_MAGIC || hdr.L != L || hdr.tmax != (uint64_t)h->st.tmax || hdr.n_prbs != (uint64_t)h->st.n_prbs ||
        hdr.n_envs != (uint64_t)h->cfg.n_envs)
        return kfail(RS_E_ARG, "blob does not belong to a handle of this shape (learners, dict_cap, n_prbs)");
    if (hdr.cursor > h->st.pool_doubles) return kfail(RS_E_NOMEM, "the dictionary pool of this handle is smaller than the checkpoint");
    if ((hdr.has_ctl != 0) != (h->ctl.acc != nullptr)) return kfail(RS_E_STATE, "controller state: call kb_control_init on both handles or on neither");
    if (bytes != kb_blob_bytes(h, hdr.cursor)) return kfail(RS_E_ARG, "bad blob size");
    const char *p = static_cast<const char *>(blob) + sizeof hdr;
    auto in = [&](void *d, size_t n) -> cudaError_t { cudaError_t e = cudaMemcpy(d, p, n, cudaMemcpyHostToDevice); p += n; return e; };
    KCU(in(h->st.D, L * sizeof(int)));
    KCU(in(h->st.rows, L * h->st.tmax * sizeof(unsigned long long)));
    KCU(in(h->st.flags, L * sizeof(uint32_t)));
    KCU(in(h->st.tie_ctr, L * sizeof(uint32_t)));
    KCU(in(h->st.pool, (size_t)hdr.cursor * sizeof(double)));
    if (h->ctl.acc) {
        KCU(in(h->ctl.acc, L * h->st.n_prbs * sizeof(double)));
        KCU(in(h->ctl.sec, L * sizeof(int32_t)));
        KCU(in(h->ctl.margins, L * sizeof(int32_t)));
        KCU(in(h->ctl.action, L * sizeof(int32_t)));
        KCU(in(h->ctl.adjusted, (size_t)h->cfg.n_envs * sizeof(int32_t)));
    }
    const unsigned long long cur = hdr.cursor;
    const int md = (int)hdr.max_d;
    KCU(cudaMemcpy(h->st.cursor, &cur, sizeof cur, cudaMemcpyHostToDevice));
    KCU(cudaMemcpy(h->st.max_d, &md, sizeof md, cudaMemcpyHostToDevice));
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
