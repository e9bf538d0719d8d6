// K2: mMTC slice step + end-of-period reduction (reward).
//
//   SliceL1mMTC.slot        slice_l1.py:87-125       FIFO of queued devices with repetition counters
//   SliceRANmMTC.slot/reset slice_ran.py:91-121      1000 periodic devices, deterministic inter-arrival
//   RanSlice.step reward    ran_slice.py:45-54
//
// The reference decrements 1000 countdowns per slot (slice_ran.py:107); here every device stores the
// ABSOLUTE slot of its next arrival, so one step scans the 1000 words once and only arriving devices
// are rewritten.  The scan is the HBM-bound part (4 KB per unit-step) and runs as its own kernel,
// mmtc_scan_kernel, with (unit, device strip) threads -- coalesced across units, 8 strips per unit so that
// 65 536 units put half a million threads in flight -- appending the arrivals of the period to a small
// per-unit list; mmtc_step_kernel (thread per unit) orders the list by (slot, device) -- the order in which
// the reference's np.where finds them -- and runs the 50 slots of FIFO processing.
// All arithmetic is integer except the per-slot means.
#include <cuda_runtime.h>

#include "philox.cuh"
#include "ranslice_state.cuh"

namespace rs {

__device__ __constant__ int c_REP_SET[7] = {2, 4, 8, 16, 32, 64, 128};                     // scenario_creator.py:88
__device__ __constant__ int c_PERIOD_SET[8] = {1000, 50000, 10000, 15000, 20000, 25000, 50000, 100000};  // :89 (50000 twice)

constexpr int MTC_STRIPS = 8;      // device strips per unit in the scan kernel (1000 / 8 = 125 devices each)

// SliceRANmMTC.reset (slice_ran.py:91-101) + SliceL1mMTC.reset (slice_l1.py:29-39)
__global__ void __launch_bounds__(128) mmtc_reset_kernel(const __grid_constant__ StepParams p,
                                                         const __grid_constant__ MmtcState st) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= st.U) return;
    const int env = u / p.n_mmtc, m = u - env * p.n_mmtc;
    const uint64_t seed = p.seed0;
    const uint32_t genv = p.env0 + (uint32_t)env;   // global env id: Philox counter word 3
    PhiloxStream r{(uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(p.n_l1e + m), STREAM_MTC, st.ctr[u], genv};   // stream of L1 slice n_l1e + m
    for (int i = 0; i < N_MTC_DEV; ++i) {
        const uint32_t rep_ix = r.integers(7);
        const uint32_t per_ix = r.integers(8);
        const uint32_t t = 1u + r.integers((uint32_t)c_PERIOD_SET[per_ix]);
        st.rep_ix[(size_t)i * st.U + u] = (uint8_t)rep_ix;
        st.period_ix[(size_t)i * st.U + u] = (uint8_t)per_ix;
        st.next_abs[(size_t)i * st.U + u] = t;               // time == 0 after reset
    }
    st.ctr[u] = r.n;
    st.q_n[u] = 0;
    st.time[u] = 0;
    st.acc[(size_t)u * 3 + 0] = st.acc[(size_t)u * 3 + 1] = st.acc[(size_t)u * 3 + 2] = 0.0;
}

// Phase 1: which devices fire in (t0, t0 + slots]?  Thread = (unit, strip of 125 devices).
__global__ void __launch_bounds__(128) mmtc_scan_kernel(const __grid_constant__ StepParams p,
                                                        const __grid_constant__ MmtcState st) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int U = st.U;
    if (u >= U) return;
    constexpr int PER = (N_MTC_DEV + MTC_STRIPS - 1) / MTC_STRIPS;
    const int i0 = blockIdx.y * PER, i1 = min(i0 + PER, N_MTC_DEV);
    const uint32_t t0 = st.time[u];
    bool overflow = false;
#pragma unroll 5
    for (int i = i0; i < i1; ++i) {
        const uint32_t d = st.next_abs[(size_t)i * U + u] - t0;     // wrap-safe: periods << 2^31
        if (d >= 1u && d <= (uint32_t)p.slots) {
            const uint32_t k = atomicAdd(&st.arr_n[u], 1u);
            if (k < (uint32_t)MTC_MAX_ARR) st.arr[(size_t)k * U + u] = (d << 16) | (uint32_t)i;
            else overflow = true;
            st.next_abs[(size_t)i * U + u] += (uint32_t)c_PERIOD_SET[st.period_ix[(size_t)i * U + u]];  // period >= 1000 > slots
        }
    }
    if (overflow) atomicOr(p.flags_acc + u / p.n_mmtc, 16u);
}

// The same scan for U % 4 == 0 (every BASELINE size): a thread owns FOUR consecutive units and reads their next-arrival
// words with one 128-bit load per device, ten devices (160 bytes per thread) in flight.  The scalar version above kept
// five 4-byte loads in flight and reached 1.7 TB/s, 26 % of the HBM peak (profiles/r01l_mmtc_scan_65536_full.txt); a first
// 128-bit version that handled an arrival where it found it was no faster (147 us): about ten lanes per warp-iteration find
// one, and each handler is a chain of L2 round trips (atomicAdd -> store, period load -> store) executed under divergence.
// So the scan only records WHERE arrivals are (one 50-bit mask per unit), and the arrivals are handled after the
// loop: one atomicAdd per unit reserves the list entries, then the independent loads / stores of all arrivals overlap.
constexpr int MTC_STRIPS_V4 = 10;  // blockIdx.y: two half-strips of 50 devices each (65 536 units: 1280 blocks = one wave of 10 blocks per SM)
__global__ void __launch_bounds__(128) mmtc_scan_kernel_v4(const __grid_constant__ StepParams p,
                                                           const __grid_constant__ MmtcState st) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int U = st.U;
    if (4 * q >= U) return;
    constexpr int PER = N_MTC_DEV / MTC_STRIPS_V4 / 2;
    static_assert(PER * MTC_STRIPS_V4 * 2 == N_MTC_DEV && PER % 10 == 0 && PER <= 64, "strip geometry");
    const uint4 t0 = *reinterpret_cast<const uint4 *>(st.time + 4 * q);
    const uint32_t slots = (uint32_t)p.slots;
    const size_t stride = (size_t)U / 4;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int i0 = (blockIdx.y * 2 + half) * PER;
    const uint4 *row = reinterpret_cast<const uint4 *>(st.next_abs + (size_t)i0 * U) + q;
    unsigned long long m0 = 0, m1 = 0, m2 = 0, m3 = 0;           // arrival masks of the four units over the half-strip
#pragma unroll 1
    for (int i = 0; i < PER; i += 10) {
        uint4 v[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) v[j] = __ldg(row + (size_t)(i + j) * stride);
#pragma unroll
        for (int j = 0; j < 10; ++j) {                           // d in [1, slots]  <=>  d - 1 < slots (wrap-safe: periods << 2^31)
            const unsigned long long bit = 1ull << (i + j);
            m0 |= (v[j].x - t0.x - 1u < slots) ? bit : 0ull;
            m1 |= (v[j].y - t0.y - 1u < slots) ? bit : 0ull;
            m2 |= (v[j].z - t0.z - 1u < slots) ? bit : 0ull;
            m3 |= (v[j].w - t0.w - 1u < slots) ? bit : 0ull;
        }
    }
    if ((m0 | m1 | m2 | m3) == 0ull) continue;
    const unsigned long long mk[4] = {m0, m1, m2, m3};
    const uint32_t tt[4] = {t0.x, t0.y, t0.z, t0.w};
    uint32_t base[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) base[k] = mk[k] ? atomicAdd(&st.arr_n[4 * q + k], (uint32_t)__popcll(mk[k])) : 0u;   // independent: in flight together
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned long long m = mk[k];
        const int u = 4 * q + k;
        uint32_t pos = base[k];
        bool overflow = false;
        while (m) {
            const int i = i0 + __ffsll((long long)m) - 1;
            m &= m - 1;
            const size_t ix = (size_t)i * U + u;
            const uint32_t nx = st.next_abs[ix];                 // (L1 / L2 hit: read a moment ago)
            if (pos < (uint32_t)MTC_MAX_ARR) st.arr[(size_t)pos * U + u] = ((nx - tt[k]) << 16) | (uint32_t)i;
            else overflow = true;
            ++pos;
            st.next_abs[ix] = nx + (uint32_t)c_PERIOD_SET[st.period_ix[ix]];      // period >= 1000 > slots
        }
        if (overflow) atomicOr(p.flags_acc + u / p.n_mmtc, 16u);
    }
  }
}

// Phase 2: the 50 slots of one mMTC slice.  Thread per unit.
__global__ void __launch_bounds__(128) mmtc_step_kernel(const __grid_constant__ StepParams p,
                                                        const __grid_constant__ MmtcState st) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int U = st.U;
    if (u >= U) return;
    const int env = u / p.n_mmtc, m = u - env * p.n_mmtc, s = p.n_l1e + m;   // s: L1 / action index of this slice
    uint32_t flags = 0;
    // PRBs of this slice after clamping (mMTC ignores i_prb, slice_l1.py:50-51)
    int n_prbs;
    {
        const int32_t *a = p.action + (size_t)env * p.S;
        int off = 0, mine = 0;
        for (int j = 0; j <= s; ++j) {
            int v = a[j];
            if (v < 0) { v = 0; flags |= 4u; }
            if (off + v > p.n_prbs) { v = p.n_prbs - off; flags |= 4u; }
            if (j == s) mine = v; else off += v;
        }
        n_prbs = mine;
    }
    st.cur_prbs[u] = n_prbs;

    const uint32_t t0 = st.time[u];
    // ---- arrivals of this period in (slot, device) order: insertion sort of the scan kernel's list (in place, coalesced)
    const int n_arr = (int)min(st.arr_n[u], (uint32_t)MTC_MAX_ARR);
    st.arr_n[u] = 0u;                                            // ready for the next step's scan
    for (int a = 1; a < n_arr; ++a) {
        const uint32_t key = st.arr[(size_t)a * U + u];
        int b = a - 1;
        while (b >= 0) {
            const uint32_t kb = st.arr[(size_t)b * U + u];
            if (kb <= key) break;
            st.arr[(size_t)(b + 1) * U + u] = kb;
            --b;
        }
        st.arr[(size_t)(b + 1) * U + u] = key;
    }
    int next_arr = 0;                                            // cursor into the ordered list
    uint32_t next_key = n_arr > 0 ? st.arr[u] : 0xFFFFFFFFu;
    int q_n = st.q_n[u];
    double a_delay = 0.0, a_rep = 0.0;
    long long a_dev = 0;
    for (int t = 1; t <= p.slots; ++t) {
        const uint32_t now = t0 + (uint32_t)t;                     // self.time += 1
        while ((next_key >> 16) == (uint32_t)t) {                  // add_users, ascending device index (np.where)
            if (q_n < st.Q) {
                st.q_rep[(size_t)q_n * U + u] = c_REP_SET[st.rep_ix[(size_t)(next_key & 0xFFFFu) * U + u]];
                st.q_t0[(size_t)q_n * U + u] = now;
                ++q_n;
            } else flags |= 16u;
            ++next_arr;
            next_key = next_arr < n_arr ? st.arr[(size_t)next_arr * U + u] : 0xFFFFFFFFu;
        }
        const int n_tx = min(n_prbs, q_n);                         // one carrier (PRB) per device
        int w = 0;
        long long sd = 0, sr = 0;
        for (int k = 0; k < q_n; ++k) {
            int rep = st.q_rep[(size_t)k * U + u];
            if (k < n_tx) rep -= 1;
            if (rep > 0) {
                const uint32_t t_start = st.q_t0[(size_t)k * U + u];
                if (w != k) st.q_t0[(size_t)w * U + u] = t_start;
                st.q_rep[(size_t)w * U + u] = rep;
                sd += (long long)(now - t_start);                  // np.maximum(0, time - t_start): never negative
                sr += rep;
                ++w;
            }
        }
        q_n = w;
        if (w > 0) {
            a_delay += (double)sd / (double)w;                     // delays.mean()
            a_rep += rint((double)sr / (double)w);                 // np.rint(repetitions.mean())
            a_dev += w;
        }
    }
    st.q_n[u] = q_n;
    st.time[u] = t0 + (uint32_t)p.slots;

    // ---- state (slice_ran.py:133-137), SLA (:145-148)
    const double acc[3] = {(double)a_dev, a_rep, a_delay};
    float *obs = p.obs + (size_t)env * p.V + p.n_embb * 10 + m * 3;
    for (int j = 0; j < 3; ++j) { obs[j] = (float)(acc[j] / p.norm_mmtc[j]); st.acc[(size_t)u * 3 + j] = acc[j]; }
    const int viol = !(a_delay / (double)p.slots < 300.0);
    p.violations[(size_t)env * p.S + s] = viol;
    p.labels[(size_t)env * p.S + s] = viol ? -1 : 1;
    if (flags) atomicOr(p.flags_acc + env, flags);
}

// Phase 2 for a MULTIPLEXED mMTC L1 (create_env(L1_level=False) with several mMTC slices, scenario_creator.py:173-176;
// slice_l1.py:87-125): the M RAN slices of an env share ONE queue (kept in the arrays of unit env * M) and one action
// entry.  Per slot the arrivals are appended RAN slice by RAN slice (each in device order), the first n_prbs queued
// devices transmit whatever their slice, and every RAN slice accumulates the statistics of its own devices.  The RAN
// slice of a queued device rides in bits 16.. of its repetition word.  Thread per env.
__global__ void __launch_bounds__(128) mmtc_step_mux_kernel(const __grid_constant__ StepParams p,
                                                            const __grid_constant__ MmtcState st) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    const int U = st.U, M = p.n_mmtc;
    if (env >= p.N) return;
    const int u0 = env * M, s = p.n_l1e;                         // queue unit; L1 / action index of the multiplexed mMTC slice
    uint32_t flags = 0;
    int n_prbs;
    {
        const int32_t *a = p.action + (size_t)env * p.S;
        int off = 0, mine = 0;
        for (int j = 0; j <= s; ++j) {
            int v = a[j];
            if (v < 0) { v = 0; flags |= 4u; }
            if (off + v > p.n_prbs) { v = p.n_prbs - off; flags |= 4u; }
            if (j == s) mine = v; else off += v;
        }
        n_prbs = mine;
    }
    const uint32_t t0 = st.time[u0];
    int n_arr[MAX_SLICES], next_arr[MAX_SLICES];
    uint32_t next_key[MAX_SLICES];
    double a_delay[MAX_SLICES], a_rep[MAX_SLICES];
    long long a_dev[MAX_SLICES];
    for (int m = 0; m < M; ++m) {                                // order every RAN slice's arrival list by (slot, device)
        const int u = u0 + m;
        st.cur_prbs[u] = n_prbs;
        const int n = (int)min(st.arr_n[u], (uint32_t)MTC_MAX_ARR);
        st.arr_n[u] = 0u;
        for (int a = 1; a < n; ++a) {
            const uint32_t key = st.arr[(size_t)a * U + u];
            int b = a - 1;
            while (b >= 0) {
                const uint32_t kb = st.arr[(size_t)b * U + u];
                if (kb <= key) break;
                st.arr[(size_t)(b + 1) * U + u] = kb;
                --b;
            }
            st.arr[(size_t)(b + 1) * U + u] = key;
        }
        n_arr[m] = n; next_arr[m] = 0; next_key[m] = n > 0 ? st.arr[u] : 0xFFFFFFFFu;
        a_delay[m] = 0.0; a_rep[m] = 0.0; a_dev[m] = 0;
    }
    int q_n = st.q_n[u0];
    for (int t = 1; t <= p.slots; ++t) {
        const uint32_t now = t0 + (uint32_t)t;
        for (int m = 0; m < M; ++m) {                            // slice_l1.py:91-95: arrivals of RAN slice m, ascending device index
            const int u = u0 + m;
            while ((next_key[m] >> 16) == (uint32_t)t) {
                if (q_n < st.Q) {
                    st.q_rep[(size_t)q_n * U + u0] = c_REP_SET[st.rep_ix[(size_t)(next_key[m] & 0xFFFFu) * U + u]] | (m << 16);
                    st.q_t0[(size_t)q_n * U + u0] = now;
                    ++q_n;
                } else flags |= 16u;
                ++next_arr[m];
                next_key[m] = next_arr[m] < n_arr[m] ? st.arr[(size_t)next_arr[m] * U + u] : 0xFFFFFFFFu;
            }
        }
        const int n_tx = min(n_prbs, q_n);
        int w = 0;
        long long sd[MAX_SLICES], sr[MAX_SLICES];
        int cnt[MAX_SLICES];
        for (int m = 0; m < M; ++m) { sd[m] = 0; sr[m] = 0; cnt[m] = 0; }
        for (int k = 0; k < q_n; ++k) {
            const int word = st.q_rep[(size_t)k * U + u0];
            int rep = word & 0xFFFF;
            const int m = word >> 16;
            if (k < n_tx) rep -= 1;
            if (rep > 0) {
                const uint32_t t_start = st.q_t0[(size_t)k * U + u0];
                if (w != k) st.q_t0[(size_t)w * U + u0] = t_start;
                st.q_rep[(size_t)w * U + u0] = rep | (m << 16);
                sd[m] += (long long)(now - t_start);
                sr[m] += rep;
                cnt[m] += 1;
                ++w;
            }
        }
        q_n = w;
        for (int m = 0; m < M; ++m)
            if (cnt[m] > 0) {
                a_delay[m] += (double)sd[m] / (double)cnt[m];
                a_rep[m] += rint((double)sr[m] / (double)cnt[m]);
                a_dev[m] += cnt[m];
            }
    }
    st.q_n[u0] = q_n;
    int l1_viol = 0;
    for (int m = 0; m < M; ++m) {
        const int u = u0 + m;
        st.time[u] = t0 + (uint32_t)p.slots;                     // (the scan kernel reads every unit's own clock)
        const double acc[3] = {(double)a_dev[m], a_rep[m], a_delay[m]};
        float *obs = p.obs + (size_t)env * p.V + p.n_embb * 10 + m * 3;
        for (int j = 0; j < 3; ++j) { obs[j] = (float)(acc[j] / p.norm_mmtc[j]); st.acc[(size_t)u * 3 + j] = acc[j]; }
        l1_viol += !(a_delay[m] / (double)p.slots < 300.0);
    }
    p.violations[(size_t)env * p.S + s] = l1_viol;              // slice_l1.py:65-75: the L1 adds its RAN slices' violations up
    p.labels[(size_t)env * p.S + s] = l1_viol ? -1 : 1;
    if (flags) atomicOr(p.flags_acc + env, flags);
}

// RanSlice.step epilogue (ran_slice.py:45-54): one thread per env
__global__ void __launch_bounds__(256) reward_kernel(const __grid_constant__ StepParams p) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.N) return;
    int tv = 0, asum = 0;
    for (int s = 0; s < p.S; ++s) {
        tv += p.violations[(size_t)env * p.S + s];
        asum += p.action[(size_t)env * p.S + s];
    }
    const double r = tv > 0 ? -1.0 * p.penalty * (double)tv : (double)max(0, p.n_prbs - asum);
    p.reward[env] = (float)r;
    const uint32_t f = p.flags_acc[env];
    p.flags[env] = f;
    if (f) p.flags_acc[env] = 0u;
}

void launch_mmtc_reset(const StepParams &p, const MmtcState &st, cudaStream_t stream) {
    mmtc_reset_kernel<<<(st.U + 127) / 128, 128, 0, stream>>>(p, st);
}
int launch_mmtc_step(const StepParams &p, const MmtcState &st, cudaStream_t stream, cudaEvent_t *prof) {
    if (st.U % 4 == 0) mmtc_scan_kernel_v4<<<dim3((st.U / 4 + 127) / 128, MTC_STRIPS_V4), 128, 0, stream>>>(p, st);
    else mmtc_scan_kernel<<<dim3((st.U + 127) / 128, MTC_STRIPS), 128, 0, stream>>>(p, st);
    if (prof) cudaEventRecord(prof[0], stream);               // profiling: end of the scan kernel
    if (p.n_l1m < p.n_mmtc) mmtc_step_mux_kernel<<<(p.N + 127) / 128, 128, 0, stream>>>(p, st);   // several RAN slices in one L1
    else mmtc_step_kernel<<<(st.U + 127) / 128, 128, 0, stream>>>(p, st);
    return 2;   // kernels launched
}
void launch_reward(const StepParams &p, cudaStream_t stream) {
    reward_kernel<<<(p.N + 255) / 256, 256, 0, stream>>>(p);
}

}  // namespace rs
