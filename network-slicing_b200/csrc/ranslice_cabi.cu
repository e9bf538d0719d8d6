// C ABI of libranslice_b200 (include/ranslice_b200.h): handle lifecycle, HBM state arena, table
// upload, step orchestration.  No torch types; the Python layer binds it with ctypes.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ranslice_b200.h"
#include "ranslice_state.cuh"

namespace rs {
void launch_embb_unit_thread(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream);
int launch_embb_fast(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream);
int launch_embb_smem(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream, cudaEvent_t *prof, const HeavyFork *fork);
int launch_embb_warp(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream, cudaEvent_t *prof);
void launch_embb_reset(const EmbbState &st, cudaStream_t stream);
void launch_mmtc_reset(const StepParams &p, const MmtcState &st, cudaStream_t stream);
int launch_mmtc_step(const StepParams &p, const MmtcState &st, cudaStream_t stream, cudaEvent_t *prof);
void launch_embb_mux(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream);
void launch_embb_mux_warp(const StepParams &p, const EmbbState &st, const Tables &tb, cudaStream_t stream);
void launch_embb_mux_reset(const EmbbState &st, cudaStream_t stream);
void launch_reward(const StepParams &p, cudaStream_t stream);
}  // namespace rs

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(RS_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

constexpr int PROF_EVENTS = 7;
#ifndef RS_HEAVY_PF_DEFAULT
#define RS_HEAVY_PF_DEFAULT 1000  // predicted contended PF chunks per step (embb_fast.cu is_heavy) above which a unit of a SMALL batch goes to the
#define RS_HEAVY_PF_DIL1 1500     // warp-per-unit kernel.  Measured (profiles/r02f_heavy_prediction.txt): 4096 envs (lane dilution 2) 2.13 ms/step
                                  // without, 1.70 / 1.60 / 1.61 at 800 / 1000 / 1200; 16 384 envs (dilution 1) 2.51 without, 2.33 / 2.45 / 2.50 at
                                  // 1500 / 2500 / 3500; 65 536 envs 4.02 without, 4.28 / 4.11 at 3000 / 4000 (a warp executes ~3x the instructions of
                                  // a lane), so it is only enabled for diluted batches
#endif
#ifndef RS_WARP_AUTO_UNITS_DEFAULT
#define RS_WARP_AUTO_UNITS_DEFAULT 16384   // batches of up to this many units run the warp-per-unit kernel: 2048 envs x 5 slices 1.18 vs 1.89 ms/step,
                                           // 4096 envs 2.05 vs 2.13 (1.94 with the heavy list), 8192 envs 3.9 vs 2.4 (profiles/r02d_*)
#endif
struct rs_handle {
    rs_config cfg;
    rs::StepParams p;
    rs::EmbbState embb;
    rs::MmtcState mmtc;
    rs::Tables tb;
    int sm_count;
    // persistent state arena (checkpointable)
    char *arena;
    size_t arena_bytes;
    // tables + I/O staging
    double *d_trace;
    int32_t *d_trace_fix, *d_trace_pre;
    rs::LutBlock *d_lut;
    char *scratch;
    unsigned long long *d_slow_paths;
    int32_t *d_action;
    float *d_obs, *d_reward;
    int32_t *d_labels, *d_violations;
    uint32_t *d_flags, *d_flags_acc;
    unsigned long long *d_trace_elems;
    cudaStream_t stream;
    cudaStream_t side_stream;                // mMTC kernels run here, concurrently with the eMBB kernels of the same step
    // rs_step_async: copy stream + double-buffered device I/O (set 0 aliases d_action / d_obs / ...)
    cudaStream_t copy_stream;
    cudaEvent_t ev_kernels[2], ev_copied[2];
    int32_t *a_action[2];
    float *a_obs[2], *a_reward[2];
    int32_t *a_labels[2], *a_violations[2];
    uint32_t *a_flags[2];
    int async_slot;
    cudaEvent_t ev_fork, ev_join;
    cudaStream_t warp_stream;                // heavy list of the default route (warp-per-unit kernel) next to the shared-memory kernel
    cudaEvent_t ev_wfork, ev_wjoin;
    // ordering between entry points that launch on different streams (rs_step_device on a caller stream, then rs_reset /
    // rs_step / rs_step_async on the handle's own stream, or the other way round): the last step's stream and an event at its end
    cudaEvent_t ev_last;
    cudaStream_t last_stream;
    bool has_last;
    uint64_t launches;
    bool was_reset;
    bool use_warp;                           // eMBB slices stepped by the warp-per-unit kernel (variant 3, or a batch too small to fill the GPU)
    bool profiling;
    std::vector<cudaEvent_t> *prof_events;   // PROF_EVENTS per profiled step
};

namespace {

struct Carver {
    size_t off = 0;
    char *base = nullptr;
    template <typename T>
    T *take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T *ptr = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return ptr;
    }
};

void carve(rs_handle *h, Carver &c) {
    rs::EmbbState &e = h->embb;
    const size_t U = (size_t)e.U, K = (size_t)e.K;
    e.hdr = c.take<rs::UnitHdr>(U);
    e.ue = c.take<rs::UeRec>(U * K);
    e.acc = c.take<double>(U * (size_t)e.R * 10);
    e.mux = e.R > 1 || h->cfg.l1_mux ? c.take<rs::MuxRan>(U * (size_t)e.R) : nullptr;
    e.cur_prbs = c.take<int32_t>(U);
    rs::MmtcState &m = h->mmtc;
    const size_t UM = (size_t)m.U, Q = (size_t)m.Q, D = (size_t)rs::N_MTC_DEV;
    m.next_abs = c.take<uint32_t>(D * UM);
    m.period_ix = c.take<uint8_t>(D * UM);
    m.rep_ix = c.take<uint8_t>(D * UM);
    m.q_rep = c.take<int32_t>(Q * UM);
    m.q_t0 = c.take<uint32_t>(Q * UM);
    m.q_n = c.take<int32_t>(UM);
    m.time = c.take<uint32_t>(UM);
    m.ctr = c.take<uint32_t>(UM);
    m.acc = c.take<double>(UM * 3);
    m.cur_prbs = c.take<int32_t>(UM);
}

double sigmoid(double x) { return 1.0 / (1.0 + std::exp(-x)); }

// MCSCodeset.compute_factors / mcs_rate_vs_error (channel_models.py:272-295) folded into an integer LUT
void build_tables(const rs_tables *t, rs::Tables &tb) {
    double A = 1.0 / 0.1;
    A = A * (std::log(1.0 / sigmoid(0.1) - 1.0) - std::log(1.0 / sigmoid(0.9) - 1.0));
    const double B = -std::log(1.0 / sigmoid(0.9) - 1.0);
    tb.A = A; tb.B = B;
    for (int m = 0; m < 26; ++m) { tb.snr_ref[m] = t->mcs_snr[m]; tb.mod[m] = (int8_t)t->mcs_mod[m]; }
    const double target = 1.0 - 0.1;
    for (int i = 0; i < 256; ++i) {
        const double snr = (double)(i - 128);
        int mcs = 0, sel = 25, rate_ix = 25;
        for (mcs = 0; mcs < 26; ++mcs)
            if (sigmoid(A * (snr - t->mcs_snr[mcs]) - B) < target) { sel = mcs - 1 > 0 ? mcs - 1 : 0; rate_ix = mcs; break; }
        const double bps = t->mcs_rate[rate_ix] * t->mcs_order[rate_ix];   // rate of the FAILING mcs
        tb.lut_mcs[i] = (int8_t)sel;
        tb.lut_rate[i] = (int16_t)(int)(158 * bps);                       // schedulers.py:45 truncation
    }
}

}  // namespace

extern "C" {

const char *rs_last_error(void) { return g_err.c_str(); }
void rs_set_error(const char *msg) { g_err = msg ? msg : ""; }   // shared with kbrl.cu (internal)

int rs_n_variables(const rs_handle *h) { return h ? h->p.V : 0; }
int rs_active_variant(const rs_handle *h) {                  // the kernel that steps the eMBB slices of this handle (1..4; 5 = multiplexed L1)
    if (!h) return -1;
    if (h->cfg.l1_mux) return 5;
    if (h->use_warp) return 3;
    if (h->cfg.kernel_variant == 1) return 1;
    if (h->cfg.kernel_variant == 2 || h->cfg.kernel_variant == 3 || h->embb.K > 16) return 2;
    return 4;
}

static int create_impl(rs_handle *h, const rs_config *cfg, const rs_tables *tables);

int rs_create(const rs_config *cfg, const rs_tables *tables, rs_handle **out) {
    if (!cfg || !tables || !out) return fail(RS_E_ARG, "null argument");
    if (cfg->abi_version != RS_ABI_VERSION) return fail(RS_E_ARG, "abi_version mismatch");
    if (cfg->n_envs <= 0 || cfg->n_embb < 0 || cfg->n_mmtc < 0 || cfg->n_embb + cfg->n_mmtc <= 0 ||
        cfg->n_embb + cfg->n_mmtc > rs::MAX_SLICES)
        return fail(RS_E_ARG, "bad n_envs / slice counts");
    if (cfg->n_prbs <= 0 || cfg->n_prbs > 2 * rs::TRACE_ROWS)
        return fail(RS_E_ARG, "n_prbs must be in [1, 200] (the trace rows wrap once, channel_models.py:144-148)");
    if (cfg->slots_per_step <= 0 || cfg->slots_per_step > 255) return fail(RS_E_ARG, "slots_per_step must be in [1,255]");
    if (cfg->kernel_variant < 0 || cfg->kernel_variant > 4) return fail(RS_E_ARG, "kernel_variant must be in [0, 4]");
    const bool mux = cfg->l1_mux != 0;
    const int K = mux ? 32 : (cfg->max_ues ? cfg->max_ues : 16), MB = cfg->max_bursts ? cfg->max_bursts : 8;   // a multiplexed L1 holds the UEs of all its RAN slices
    const int Q = cfg->mtc_queue_cap ? cfg->mtc_queue_cap : 128;
    if (K < 2 || K > 32 || MB < 1 || MB > 16 || Q < 1) return fail(RS_E_ARG, "caps out of range");
    if (!tables->trace || !tables->mcs_rate || !tables->mcs_snr || !tables->mcs_order || !tables->mcs_mod)
        return fail(RS_E_ARG, "null table pointer");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(RS_E_ARG, "bad device ordinal");
    CU(cudaSetDevice(cfg->device));

    rs_handle *h = new rs_handle();
    std::memset(static_cast<void *>(h), 0, sizeof(*h));
    h->cfg = *cfg;
    h->cfg.max_ues = K; h->cfg.max_bursts = MB; h->cfg.mtc_queue_cap = Q;
    const int rc = create_impl(h, cfg, tables);
    if (rc != RS_OK) { rs_destroy(h); return rc; }             // frees whatever had been allocated (the error text is kept)
    *out = h;
    return RS_OK;
}

static int create_impl(rs_handle *h, const rs_config *cfg, const rs_tables *tables) {
    const bool mux = cfg->l1_mux != 0;
    const int K = h->cfg.max_ues, MB = h->cfg.max_bursts, Q = h->cfg.mtc_queue_cap;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;

    rs::StepParams &p = h->p;
    p.N = cfg->n_envs; p.n_embb = cfg->n_embb; p.n_mmtc = cfg->n_mmtc;
    p.n_l1e = mux ? (cfg->n_embb > 0 ? 1 : 0) : cfg->n_embb;
    p.n_l1m = mux ? (cfg->n_mmtc > 0 ? 1 : 0) : cfg->n_mmtc;   // L1_level=False also multiplexes the mMTC RAN slices (scenario_creator.py:173-176)
    p.S = p.n_l1e + p.n_l1m;
    p.n_prbs = cfg->n_prbs; p.slots = cfg->slots_per_step; p.V = 10 * cfg->n_embb + 3 * cfg->n_mmtc;
    p.penalty = cfg->penalty; p.prop_A = cfg->prop_A; p.prop_B = cfg->prop_B;
    p.seed0 = cfg->base_seed; p.env0 = (uint32_t)cfg->first_env_id;
    {   // scenario_creator.py:106,115-134 (same fp64 expressions)
        const double tps = cfg->slots_per_step * 1e-3;
        const int sps = cfg->slots_per_step;
        const double ne[10] = {5e6 * tps, 10e6 * tps, 25.0 * sps, 10e4 * sps, 35.0 * sps,
                               5e6 * tps, 10e6 * tps, 35.0 * sps, 10e4 * sps, 35.0 * sps};
        std::memcpy(p.norm_embb, ne, sizeof ne);
        for (int i = 0; i < 3; ++i) p.norm_mmtc[i] = 100.0 * sps;
        p.obs_time = cfg->slots_per_step * 1e-3;
    }
    h->embb.U = cfg->n_envs * p.n_l1e; h->embb.K = K; h->embb.MB = MB;
    h->embb.route[0] = 6; h->embb.route[1] = 8; h->embb.route[2] = 14; h->embb.route[3] = 16;   // embb_smem.cu: SM_KS = 8 slots per lane
    h->embb.R = mux ? cfg->n_embb : 1;
    h->mmtc.U = cfg->n_envs * cfg->n_mmtc; h->mmtc.Q = Q;

    Carver sizing;
    carve(h, sizing);
    h->arena_bytes = (sizing.off + 255) & ~size_t(255);
    CU(cudaMalloc(&h->arena, h->arena_bytes));
    CU(cudaMemset(h->arena, 0, h->arena_bytes));
    Carver real; real.base = h->arena;
    carve(h, real);

    const size_t n_trace = (size_t)3 * rs::N_SAMPLES * rs::TRACE_ROWS;
    CU(cudaMalloc(&h->d_trace, n_trace * sizeof(double)));
    CU(cudaMemcpy(h->d_trace, tables->trace, n_trace * sizeof(double), cudaMemcpyHostToDevice));
    {   // 2^-22 fixed-point copy (|v| < 127 dB, embb_fastmath.cuh FIX_BITS): exact integer window sums on the fast path
        std::vector<int32_t> fix(n_trace);
        for (size_t i = 0; i < n_trace; ++i) {
            const double v = tables->trace[i];
            if (std::isnan(v)) { fix[i] = 0; continue; }
            if (std::fabs(v) >= 127.0) return fail(RS_E_ARG, "fading trace value out of the +-127 dB fixed-point range");
            fix[i] = (int32_t)std::llrint(v * 4194304.0);
        }
        CU(cudaMalloc(&h->d_trace_fix, n_trace * sizeof(int32_t)));
        CU(cudaMemcpy(h->d_trace_fix, fix.data(), n_trace * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    build_tables(tables, h->tb);
    h->tb.trace = h->d_trace; h->tb.trace_fix = h->d_trace_fix;
    {   // per-column prefix sums for the PRB-window mean (embb_fastmath.cuh window_sum_prefix): the largest fixed-point scale
        // that keeps every window sum of up to 300 wrapped rows inside int32, at most 2^22
        const size_t n_cols = (size_t)3 * rs::N_SAMPLES;
        double span = 0.0;                                    // max over columns of (max - min) of the prefix over 3 wraps
        for (size_t c = 0; c < n_cols; ++c) {
            const double *col = tables->trace + c * rs::TRACE_ROWS;
            if (std::isnan(col[0])) continue;
            double acc = 0.0, lo = 0.0, hi = 0.0;
            for (int r = 0; r < 3 * rs::TRACE_ROWS; ++r) { acc += col[r % rs::TRACE_ROWS]; lo = std::min(lo, acc); hi = std::max(hi, acc); }
            span = std::max(span, hi - lo);
        }
        int bits = 22;
        while (bits > 8 && (span + 300.0) * std::ldexp(1.0, bits) >= 2147483647.0) --bits;   // + 300: half a unit of rounding per row
        h->tb.pre_bits = bits; h->tb.pre_inv = std::ldexp(1.0, -bits); h->tb.pre_guard = 2.5 * std::ldexp(1.0, -(bits + 1));
        std::vector<int32_t> pre(n_cols * rs::PRE_STRIDE, 0);
        const double scale = std::ldexp(1.0, bits);
        for (size_t c = 0; c < n_cols; ++c) {
            const double *col = tables->trace + c * rs::TRACE_ROWS;
            if (std::isnan(col[0])) continue;
            uint32_t acc = 0;                                  // modular: only differences of prefixes are used
            for (int r = 0; r < rs::TRACE_ROWS; ++r) {
                pre[c * rs::PRE_STRIDE + r] = (int32_t)acc;
                acc += (uint32_t)(int32_t)std::llrint(col[r] * scale);
            }
            pre[c * rs::PRE_STRIDE + rs::TRACE_ROWS] = (int32_t)acc;
        }
        CU(cudaMalloc(&h->d_trace_pre, pre.size() * sizeof(int32_t)));
        CU(cudaMemcpy(h->d_trace_pre, pre.data(), pre.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        h->tb.trace_pre = h->d_trace_pre;
    }
    {   // packed lookup tables (one bulk copy per block in embb_warp.cu); same fp32 expressions as the kernels' own staging loops
        static const double MI_X0[3] = {-0.25040431, 5.12440916, 9.16962738}, MI_K[3] = {0.31591749, 0.25423209, 0.22298101};   // channel_models.py:268-270
        rs::LutBlock lb;
        std::memset(&lb, 0, sizeof lb);
        for (int i = 0; i < 256; ++i) { lb.rate[i] = h->tb.lut_rate[i]; lb.mcs[i] = h->tb.lut_mcs[i]; }
        for (int m = 0; m < 26; ++m) { lb.ref[m] = (float)h->tb.snr_ref[m]; lb.mod[m] = h->tb.mod[m]; }
        for (int m = 0; m < 3; ++m) {
            const float kf = (float)MI_K[m], x0f = (float)MI_X0[m], l2e = 1.4426950408889634f;
            lb.mi[m][0] = kf; lb.mi[m][1] = x0f; lb.mi[m][2] = -kf * l2e; lb.mi[m][3] = kf * x0f * l2e;
        }
        for (int i = 1; i <= 2 * rs::TRACE_ROWS; ++i) lb.inv[i] = 1.0f / (float)i;      // == __frcp_rn((float)i)
        CU(cudaMalloc(&h->d_lut, sizeof lb));
        CU(cudaMemcpy(h->d_lut, &lb, sizeof lb, cudaMemcpyHostToDevice));
        h->tb.lut = h->d_lut;
    }
    {   // tensor map of the fixed-point trace table for TMA column loads (embb_warp.cu); the driver API is reached through the
        // runtime's entry-point query, so the library does not link libcuda
        std::memset(h->tb.tmap_fix, 0, sizeof h->tb.tmap_fix);
        h->tb.tmap_ok = 0;
        typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                      const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && fn && q == cudaDriverEntryPointSuccess) {
            static_assert(sizeof(CUtensorMap) == sizeof(h->tb.tmap_fix), "CUtensorMap is 128 bytes");
            CUtensorMap tm;
            const cuuint64_t dims[2] = {(cuuint64_t)rs::TRACE_ROWS, (cuuint64_t)3 * rs::N_SAMPLES};
            const cuuint64_t strides[1] = {(cuuint64_t)rs::TRACE_ROWS * sizeof(int32_t)};
            const cuuint32_t box[2] = {(cuuint32_t)rs::TRACE_ROWS, 1u}, estr[2] = {1u, 1u};
            if (reinterpret_cast<encode_fn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, h->d_trace_fix, dims, strides, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                std::memcpy(h->tb.tmap_fix, &tm, sizeof tm);
                h->tb.tmap_ok = 1;
            }
        } else cudaGetLastError();
    }
    {   // per-step scheduling scratch (outside the checkpoint arena)
        Carver sc;
        const size_t U = (size_t)h->embb.U;
        // lane dilution of the shared-memory kernel (ranslice_state.cuh): every 2nd lane while the diluted list (with ~10 % of pair
        // entries) stays within 2.5x the lanes the GPU keeps resident (4 blocks of 128 threads per SM), every 4th within 1.75x.
        // warp-per-unit kernel (embb_warp.cu): kernel_variant 3, or automatically while the batch leaves the GPU underfilled
        // (measured crossover; RS_WARP_AUTO_UNITS overrides)
        size_t warp_auto = RS_WARP_AUTO_UNITS_DEFAULT;
        if (const char *e = std::getenv("RS_WARP_AUTO_UNITS")) warp_auto = (size_t)std::atoll(e);
        h->use_warp = h->embb.K <= 16 && U > 0 && !h->cfg.l1_mux && (h->cfg.kernel_variant == 3 || (h->cfg.kernel_variant == 0 && U <= warp_auto));
        int dil = 0;
        const bool smem_variant = h->cfg.kernel_variant == 0 || h->cfg.kernel_variant == 4;
        if (smem_variant && h->embb.K <= 16 && U > 0 && !h->cfg.l1_mux && !h->use_warp) {
            // (profiles/r02f_dilution_sweep.jsonl, ms/step at dilution 0 / 1 / 2: 4096 envs 2.38 / 1.88 / 1.57, 6144 envs 2.49 / 1.96 / 1.96,
            //  8192 envs 2.64 / 2.00 / 2.47, 16 384 envs 2.60 / 2.29 / 4.21, 24 576 envs 2.80 / 3.09 / 5.94)
            const double resident = 4.0 * 128.0 * (double)h->sm_count, list = 1.1 * (double)U;
            if (2.0 * list <= 2.5 * resident) dil = 1;
            if (4.0 * list <= 1.75 * resident) dil = 2;
            if (const char *e = std::getenv("RS_DILUTION")) dil = std::max(0, std::min(2, std::atoi(e)));
        }
        h->embb.dil = dil;
        h->embb.wide = dil == 2;      // smallest batches: latency variant (measured: 4096 envs 2.94 -> 2.74 ms/step; no gain at 16384)
        if (const char *e = std::getenv("RS_WIDE")) h->embb.wide = std::atoi(e) != 0;
        // heavy list of the default route (ranslice_state.cuh): threshold on last step's contended PF chunks; RS_HEAVY_PF overrides (0 = off)
        h->embb.heavy_thr = 0;
        if (h->cfg.kernel_variant == 0 && !h->use_warp && h->embb.K <= 16 && U > 0 && !h->cfg.l1_mux) {   // (variant 4 is the pure shared-memory route)
            h->embb.heavy_thr = dil == 2 ? RS_HEAVY_PF_DEFAULT : dil == 1 ? RS_HEAVY_PF_DIL1 : 0;   // only where the step is bound by its slowest lanes (small batches)
            if (const char *e = std::getenv("RS_HEAVY_PF")) h->embb.heavy_thr = std::max(0, std::atoi(e));
        }
        h->embb.heavy_cap = (int)std::min<size_t>(U, (size_t)8 * 4 * (size_t)h->sm_count);   // at most four 8-warp blocks per SM
        if (const char *e = std::getenv("RS_HEAVY_CAP")) h->embb.heavy_cap = (int)std::min<size_t>(U, (size_t)std::max(8, std::atoi(e)));
        h->embb.perm_len = (int)((2 * U) << dil);
        const size_t perm_len = (size_t)h->embb.perm_len;
        sc.take<uint32_t>(U); sc.take<int32_t>(perm_len); sc.take<uint32_t>(2 * rs::SORT_BINS + 4 + rs::SCAN_BLOCKS); sc.take<uint32_t>(U); sc.take<float>(8); sc.take<rs::ColdRec>(U * (size_t)h->embb.K); sc.take<int32_t>(U + 1);
        CU(cudaMalloc(&h->scratch, sc.off + 256));
        CU(cudaMemset(h->scratch, 0, sc.off + 256));
        Carver rc; rc.base = h->scratch;
        h->embb.win = rc.take<uint32_t>(U); h->embb.perm = rc.take<int32_t>(perm_len);
        h->embb.hist = rc.take<uint32_t>(2 * rs::SORT_BINS + 4 + rs::SCAN_BLOCKS); h->embb.hint = rc.take<uint32_t>(U); h->embb.dbg = rc.take<float>(8);
        h->embb.cold = rc.take<rs::ColdRec>(U * (size_t)h->embb.K);
        h->embb.wlist = rc.take<int32_t>(U + 1);
    }
    if (h->mmtc.U) {   // arrival scratch of the mMTC scan kernel
        const size_t UM = (size_t)h->mmtc.U;
        CU(cudaMalloc(&h->mmtc.arr_n, UM * sizeof(uint32_t)));
        CU(cudaMemset(h->mmtc.arr_n, 0, UM * sizeof(uint32_t)));
        CU(cudaMalloc(&h->mmtc.arr, UM * rs::MTC_MAX_ARR * sizeof(uint32_t)));
    }

    const size_t N = (size_t)p.N, S = (size_t)p.S, V = (size_t)p.V;
    CU(cudaMalloc(&h->d_action, N * S * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_obs, N * V * sizeof(float)));
    CU(cudaMalloc(&h->d_reward, N * sizeof(float)));
    CU(cudaMalloc(&h->d_labels, N * S * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_violations, N * S * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_flags, N * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_flags_acc, N * sizeof(uint32_t)));
    CU(cudaMemset(h->d_flags_acc, 0, N * sizeof(uint32_t)));
    CU(cudaMalloc(&h->d_trace_elems, 4 * sizeof(unsigned long long)));
    CU(cudaMemset(h->d_trace_elems, 0, 4 * sizeof(unsigned long long)));
    h->d_slow_paths = h->d_trace_elems + 1;
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&h->ev_kernels[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
    }
    h->a_action[0] = h->d_action; h->a_obs[0] = h->d_obs; h->a_reward[0] = h->d_reward; h->a_labels[0] = h->d_labels;
    h->a_violations[0] = h->d_violations; h->a_flags[0] = h->d_flags;
    CU(cudaMalloc(&h->a_action[1], N * S * sizeof(int32_t)));
    CU(cudaMalloc(&h->a_obs[1], N * V * sizeof(float)));
    CU(cudaMalloc(&h->a_reward[1], N * sizeof(float)));
    CU(cudaMalloc(&h->a_labels[1], N * S * sizeof(int32_t)));
    CU(cudaMalloc(&h->a_violations[1], N * S * sizeof(int32_t)));
    CU(cudaMalloc(&h->a_flags[1], N * sizeof(uint32_t)));
    h->async_slot = 1;
    CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&h->warp_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_wfork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_wjoin, cudaEventDisableTiming));
    p.flags_acc = h->d_flags_acc;
    p.trace_elems = h->d_trace_elems;
    p.slow_paths = h->d_slow_paths;
    p.debug_check = 0;
    return RS_OK;
}

int rs_destroy(rs_handle *h) {
    if (!h) return RS_OK;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(h->arena); cudaFree(h->d_trace); cudaFree(h->d_trace_fix); cudaFree(h->d_trace_pre); cudaFree(h->d_lut); cudaFree(h->scratch); cudaFree(h->d_action); cudaFree(h->d_obs);
    cudaFree(h->d_reward); cudaFree(h->d_labels); cudaFree(h->d_violations); cudaFree(h->d_flags);
    cudaFree(h->d_flags_acc); cudaFree(h->d_trace_elems);
    if (h->mmtc.U) { cudaFree(h->mmtc.arr_n); cudaFree(h->mmtc.arr); }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (int i = 0; i < 2; ++i) { if (h->ev_kernels[i]) cudaEventDestroy(h->ev_kernels[i]); if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); }
    cudaFree(h->a_action[1]); cudaFree(h->a_obs[1]); cudaFree(h->a_reward[1]); cudaFree(h->a_labels[1]); cudaFree(h->a_violations[1]);
    cudaFree(h->a_flags[1]);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_last) cudaEventDestroy(h->ev_last);
    if (h->warp_stream) cudaStreamDestroy(h->warp_stream);
    if (h->ev_wfork) cudaEventDestroy(h->ev_wfork);
    if (h->ev_wjoin) cudaEventDestroy(h->ev_wjoin);
    if (h->prof_events) { for (auto e : *h->prof_events) cudaEventDestroy(e); delete h->prof_events; }
    delete h;
    return RS_OK;
}

int rs_reset(rs_handle *h, float *obs) {
    if (!h) return fail(RS_E_ARG, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    // NodeB.reset (node_b.py:17-22): UEs, timers and accumulators cleared; RNG counters keep running
    // (the reference never reseeds on reset).  eMBB: zero everything but the counters.
    // Steps still queued on another stream (rs_step_device on the caller's stream, rs_step_async) finish first.
    if (h->has_last) CU(cudaStreamWaitEvent(h->stream, h->ev_last, 0));
    CU(cudaStreamWaitEvent(h->stream, h->ev_copied[0], 0));
    CU(cudaStreamWaitEvent(h->stream, h->ev_copied[1], 0));
    if (h->embb.U) { rs::launch_embb_reset(h->embb, h->stream); h->launches += 1; }
    if (h->embb.U && h->cfg.l1_mux) { rs::launch_embb_mux_reset(h->embb, h->stream); h->launches += 1; }
    if (h->mmtc.U) {
        rs::launch_mmtc_reset(h->p, h->mmtc, h->stream);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(h->stream));
    if (obs) std::memset(obs, 0, sizeof(float) * (size_t)h->p.N * h->p.V);
    h->was_reset = true;
    return RS_OK;
}

int rs_step_device(rs_handle *h, const int32_t *d_action, float *d_obs, float *d_reward, int32_t *d_labels,
                   int32_t *d_violations, uint32_t *d_flags, void *stream) {
    if (!h || !d_action) return fail(RS_E_ARG, "null handle/action");
    if (!h->was_reset) return fail(RS_E_STATE, "rs_step before rs_reset");
    CU(cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    // a step on another stream than the previous one (caller's stream vs the handle's own) is ordered after it
    if (h->has_last && h->last_stream != st) CU(cudaStreamWaitEvent(st, h->ev_last, 0));
    rs::StepParams p = h->p;
    p.action = d_action;
    p.obs = d_obs ? d_obs : h->d_obs;
    p.reward = d_reward ? d_reward : h->d_reward;
    p.labels = d_labels ? d_labels : h->d_labels;
    p.violations = d_violations ? d_violations : h->d_violations;
    p.flags = d_flags ? d_flags : h->d_flags;
    CU(cudaMemsetAsync(h->d_trace_elems, 0, 4 * sizeof(unsigned long long), st));
    // profiling events of one step: [0] start, [1] / [2] around the dominant eMBB kernel, [3] end of the eMBB kernels,
    // [4] end of the mMTC scan, [5] end of the mMTC kernels, [6] end of the step
    cudaEvent_t ev[PROF_EVENTS] = {};
    if (h->profiling) {
        for (auto &e : ev) CU(cudaEventCreate(&e));
        CU(cudaEventRecord(ev[0], st));
    }
    // eMBB and mMTC slices of a step are independent (disjoint state, disjoint output columns): when both exist the
    // mMTC kernels run on a side stream next to the eMBB kernels and join before the reward reduction.  With
    // per-kernel profiling on, everything stays on one stream so that the events bracket single kernels.
    const bool fork = h->embb.U && h->mmtc.U && !h->profiling;
    bool dominant = false;
    if (fork) {
        CU(cudaEventRecord(h->ev_fork, st));
        CU(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        h->launches += rs::launch_mmtc_step(p, h->mmtc, h->side_stream, nullptr);
        CU(cudaEventRecord(h->ev_join, h->side_stream));
    }
    if (h->embb.U) {
        if (h->cfg.l1_mux) {                                        // multiplexed L1: warp per env; kernel_variant 1 = the all-fp64 thread-per-env anchor
            if (h->cfg.kernel_variant == 1) rs::launch_embb_mux(p, h->embb, h->tb, st); else rs::launch_embb_mux_warp(p, h->embb, h->tb, st);
            h->launches += 1;
        }
        else if (h->cfg.kernel_variant == 1) { rs::launch_embb_unit_thread(p, h->embb, h->tb, st); h->launches += 1; }
        else if (h->use_warp) { h->launches += rs::launch_embb_warp(p, h->embb, h->tb, st, h->profiling ? ev + 1 : nullptr); dominant = true; }
        else if (h->cfg.kernel_variant == 2 || h->cfg.kernel_variant == 3 || h->embb.K > 16) h->launches += rs::launch_embb_fast(p, h->embb, h->tb, st);
        else {
            const rs::HeavyFork hf{h->warp_stream, h->ev_wfork, h->ev_wjoin};
            h->launches += rs::launch_embb_smem(p, h->embb, h->tb, st, h->profiling ? ev + 1 : nullptr, h->profiling ? nullptr : &hf);
            dominant = true;
        }
    }
    if (h->profiling) {
        if (!dominant) { CU(cudaEventRecord(ev[1], st)); CU(cudaEventRecord(ev[2], st)); }   // other variants: no single dominant kernel
        CU(cudaEventRecord(ev[3], st));
    }
    if (fork) CU(cudaStreamWaitEvent(st, h->ev_join, 0));
    else if (h->mmtc.U) h->launches += rs::launch_mmtc_step(p, h->mmtc, st, h->profiling ? ev + 4 : nullptr);
    if (h->profiling) {
        if (!h->mmtc.U) CU(cudaEventRecord(ev[4], st));
        CU(cudaEventRecord(ev[5], st));
    }
    rs::launch_reward(p, st);
    h->launches += 1;
    if (h->profiling) {
        CU(cudaEventRecord(ev[6], st));
        for (auto &e : ev) h->prof_events->push_back(e);
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_last, st));
    h->last_stream = st; h->has_last = true;
    return RS_OK;
}

int rs_step(rs_handle *h, const int32_t *action, float *obs, float *reward, int32_t *labels, int32_t *violations,
            uint32_t *flags) {
    if (!h || !action) return fail(RS_E_ARG, "null handle/action");
    CU(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)h->p.N, S = (size_t)h->p.S, V = (size_t)h->p.V;
    // device I/O set 0 is shared with rs_step_async: an outstanding ticket's copies finish before it is overwritten; a
    // step still queued on a caller's stream (rs_step_device) reads its own action buffer but shares the state
    CU(cudaStreamWaitEvent(h->stream, h->ev_copied[0], 0));
    if (h->has_last && h->last_stream != h->stream) CU(cudaStreamWaitEvent(h->stream, h->ev_last, 0));
    CU(cudaMemcpyAsync(h->d_action, action, N * S * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int rc = rs_step_device(h, h->d_action, nullptr, nullptr, nullptr, nullptr, nullptr, h->stream);
    if (rc) return rc;
    if (obs) CU(cudaMemcpyAsync(obs, h->d_obs, N * V * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (reward) CU(cudaMemcpyAsync(reward, h->d_reward, N * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (labels) CU(cudaMemcpyAsync(labels, h->d_labels, N * S * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (violations) CU(cudaMemcpyAsync(violations, h->d_violations, N * S * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    if (flags) CU(cudaMemcpyAsync(flags, h->d_flags, N * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return RS_OK;
}

int rs_step_async(rs_handle *h, const int32_t *action, float *obs, float *reward, int32_t *labels, int32_t *violations,
                  uint32_t *flags, int32_t *ticket) {
    if (!h || !action || !ticket) return fail(RS_E_ARG, "null handle/action/ticket");
    CU(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)h->p.N, S = (size_t)h->p.S, V = (size_t)h->p.V;
    const int k = h->async_slot ^= 1;
    // the kernels of this step overwrite device set k: wait until the copies of the step that used it are done
    CU(cudaStreamWaitEvent(h->stream, h->ev_copied[k], 0));
    CU(cudaMemcpyAsync(h->a_action[k], action, N * S * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
    int rc = rs_step_device(h, h->a_action[k], h->a_obs[k], h->a_reward[k], h->a_labels[k], h->a_violations[k], h->a_flags[k], h->stream);
    if (rc) return rc;
    CU(cudaEventRecord(h->ev_kernels[k], h->stream));
    CU(cudaStreamWaitEvent(h->copy_stream, h->ev_kernels[k], 0));
    cudaStream_t cs = h->copy_stream;
    if (obs) CU(cudaMemcpyAsync(obs, h->a_obs[k], N * V * sizeof(float), cudaMemcpyDeviceToHost, cs));
    if (reward) CU(cudaMemcpyAsync(reward, h->a_reward[k], N * sizeof(float), cudaMemcpyDeviceToHost, cs));
    if (labels) CU(cudaMemcpyAsync(labels, h->a_labels[k], N * S * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    if (violations) CU(cudaMemcpyAsync(violations, h->a_violations[k], N * S * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    if (flags) CU(cudaMemcpyAsync(flags, h->a_flags[k], N * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
    CU(cudaEventRecord(h->ev_copied[k], cs));
    *ticket = k;
    return RS_OK;
}

int rs_wait(rs_handle *h, int32_t ticket) {
    if (!h || ticket < 0 || ticket > 1) return fail(RS_E_ARG, "bad handle/ticket");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaEventSynchronize(h->ev_copied[ticket]));
    return RS_OK;
}

int rs_get_info(rs_handle *h, int32_t env, double *acc, int32_t *n_prbs) {
    if (!h || env < 0 || env >= h->p.N) return fail(RS_E_ARG, "bad handle/env");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    const int ne = h->p.n_embb, nm = h->p.n_mmtc, nl = h->p.n_l1e;   // accumulator rows = RAN slices (ne + nm), PRB entries = L1 slices (nl + nm)
    if (acc) {
        std::memset(acc, 0, sizeof(double) * 10 * (size_t)(ne + nm));
        if (ne) CU(cudaMemcpy(acc, h->embb.acc + (size_t)env * ne * 10, sizeof(double) * 10 * ne, cudaMemcpyDeviceToHost));
        for (int m = 0; m < nm; ++m)
            CU(cudaMemcpy(acc + (size_t)(ne + m) * 10, h->mmtc.acc + ((size_t)env * nm + m) * 3, sizeof(double) * 3,
                          cudaMemcpyDeviceToHost));
    }
    if (n_prbs) {
        if (nl) CU(cudaMemcpy(n_prbs, h->embb.cur_prbs + (size_t)env * nl, sizeof(int32_t) * nl, cudaMemcpyDeviceToHost));
        if (nm) CU(cudaMemcpy(n_prbs + nl, h->mmtc.cur_prbs + (size_t)env * nm, sizeof(int32_t) * h->p.n_l1m, cudaMemcpyDeviceToHost));
    }
    return RS_OK;
}

int rs_get_n_ues(rs_handle *h, int32_t *n_ues) {
    if (!h || !n_ues) return fail(RS_E_ARG, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    if (h->embb.U) {
        std::vector<rs::UnitHdr> hd((size_t)h->embb.U);
        CU(cudaMemcpy(hd.data(), h->embb.hdr, sizeof(rs::UnitHdr) * hd.size(), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < hd.size(); ++i) n_ues[i] = hd[i].n_ues;
    }
    return RS_OK;
}

int rs_state_size(rs_handle *h, size_t *bytes) {
    if (!h || !bytes) return fail(RS_E_ARG, "null argument");
    *bytes = h->arena_bytes;
    return RS_OK;
}
int rs_get_state(rs_handle *h, void *blob, size_t bytes) {
    if (!h || !blob || bytes != h->arena_bytes) return fail(RS_E_ARG, "bad blob size");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(blob, h->arena, bytes, cudaMemcpyDeviceToHost));
    return RS_OK;
}
int rs_set_state(rs_handle *h, const void *blob, size_t bytes) {
    if (!h || !blob || bytes != h->arena_bytes) return fail(RS_E_ARG, "bad blob size");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(h->arena, blob, bytes, cudaMemcpyHostToDevice));
    h->was_reset = true;
    return RS_OK;
}

// Host-side exhaustive check of the identities the fast kernel relies on (no GPU needed):
//  (1) the two-FMA division by slot_length equals the IEEE quotient for every bits value a TTI can carry
//      ((b * bits) / slot_length, schedulers.py:63 / slice_ran.py:55; bits <= 200 PRBs * 853 bits);
//  (2) PF_A / PF_B are the reference's 1 - 1/50 and 1/50.
int rs_selftest(void) {
    const double b = 1.0 / 50, a = 1 - b, slot = 1e-3;
    if (a != 0.98 || b != 0.02) return fail(RS_E_STATE, "EWMA constants differ from 1-1/50, 1/50");
    for (int bits = 0; bits <= 200 * 853 + 1000; ++bits) {
        const double y = b * (double)bits;
        const double q = y * 1000.0;
        const double r = std::fma(-q, slot, y);
        const double q2 = std::fma(r, 1000.0, q);
        if (q2 != y / slot) return fail(RS_E_STATE, "two-FMA division by slot_length is not exact for bits=" + std::to_string(bits));
    }
    // (3) x / n by two FMAs (update_info means, slice_ran.py:290-291,304-305) for counts 2..32
    unsigned long long lcg = 88172645463325252ull;
    for (int n = 2; n <= 32; ++n) {
        const double dn = (double)n, rcp = 1.0 / dn;
        for (int it = 0; it < 200000; ++it) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const double x = (it & 1) ? (double)(long long)(lcg >> (16 + (it % 40))) : -(double)(long long)(lcg >> 40);
            const double q = x * rcp;
            const double q2 = std::fma(std::fma(-q, dn, x), rcp, q);
            if (q2 != x / dn) return fail(RS_E_STATE, "two-FMA division by a count is not exact");
        }
    }
    return RS_OK;
}

int rs_set_route_limits(rs_handle *h, int32_t single_start_max, int32_t single_slots, int32_t pair_start_max, int32_t pair_slots) {
    if (!h) return fail(RS_E_ARG, "null handle");
    if (single_slots < 1 || single_slots > 8 || pair_slots < single_slots || pair_slots > 16 || single_start_max < 0 ||
        single_start_max > single_slots || pair_start_max < single_start_max || pair_start_max > pair_slots)
        return fail(RS_E_ARG, "route limits: 0 <= single_start_max <= single_slots <= 8, single_start_max <= pair_start_max <= pair_slots <= 16");
    h->embb.route[0] = single_start_max; h->embb.route[1] = single_slots; h->embb.route[2] = pair_start_max; h->embb.route[3] = pair_slots;
    return RS_OK;
}

int rs_set_heavy_threshold(rs_handle *h, int32_t contended_chunks_per_step, int32_t max_units) {
    if (!h || contended_chunks_per_step < 0 || max_units < 0) return fail(RS_E_ARG, "bad handle / negative argument");
    // only the lane-per-unit route has a heavy list; the other routes ignore the call
    if (!h->embb.U || h->cfg.l1_mux || h->cfg.kernel_variant != 0 || h->embb.K > 16 || h->use_warp) return RS_OK;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    h->embb.heavy_thr = contended_chunks_per_step;
    h->embb.heavy_cap = (int)std::min<size_t>((size_t)h->embb.U, max_units ? (size_t)max_units : (size_t)32 * (size_t)h->sm_count);
    return RS_OK;
}

int rs_get_routes(rs_handle *h, uint64_t *out5) {
    uint64_t *out4 = out5;
    if (!h || !out4) return fail(RS_E_ARG, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    out4[0] = out4[1] = out4[2] = out4[3] = out5[4] = 0;
    if (h->use_warp) { out5[4] = (uint64_t)h->embb.U; return RS_OK; }
    if (!h->embb.U || h->cfg.l1_mux || (h->cfg.kernel_variant != 0 && h->cfg.kernel_variant != 4) || h->embb.K > 16 || h->use_warp) return RS_OK;   // other variants do not route
    uint32_t t[4];
    CU(cudaMemcpy(t, h->embb.hist + 2 * rs::SORT_BINS, sizeof t, cudaMemcpyDeviceToHost));
    out4[0] = (uint64_t)t[0] - 2ull * t[3]; out4[1] = t[3]; out4[2] = t[2]; out4[3] = (uint64_t)t[1] - t[2];
    if (h->embb.heavy_thr > 0) {
        int32_t n = 0;
        CU(cudaMemcpy(&n, h->embb.wlist + h->embb.U, sizeof n, cudaMemcpyDeviceToHost));
        out5[4] = (uint64_t)n;
    }
    return RS_OK;
}

int rs_set_debug_check(rs_handle *h, int32_t enable) {
    if (!h) return fail(RS_E_ARG, "null handle");
    h->p.debug_check = enable != 0;
    return RS_OK;
}

int rs_get_diag(rs_handle *h, double *out, int32_t n) {
    if (!h || !out || n < 5) return fail(RS_E_ARG, "need room for 5 doubles");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    float dbg[8];
    unsigned long long ctr[4];
    CU(cudaMemcpy(dbg, h->embb.dbg, sizeof dbg, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ctr, h->d_trace_elems, sizeof ctr, cudaMemcpyDeviceToHost));
    unsigned mism;
    std::memcpy(&mism, &dbg[2], sizeof mism);
    out[0] = dbg[0]; out[1] = dbg[1]; out[2] = (double)mism; out[3] = (double)ctr[1]; out[4] = (double)ctr[2];
    if (n >= 6) out[5] = (double)ctr[3];
    return RS_OK;
}

int rs_set_profiling(rs_handle *h, int32_t enable) {
    if (!h) return fail(RS_E_ARG, "null handle");
    if (!h->prof_events) h->prof_events = new std::vector<cudaEvent_t>();
    h->profiling = enable != 0;
    return RS_OK;
}

int rs_get_profile(rs_handle *h, double *ms6, uint64_t *steps) {
    if (!h || !ms6) return fail(RS_E_ARG, "null argument");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaDeviceSynchronize());
    double s[PROF_EVENTS - 1] = {};
    uint64_t n = 0;
    if (h->prof_events) {
        std::vector<cudaEvent_t> &v = *h->prof_events;
        for (size_t i = 0; i + PROF_EVENTS <= v.size(); i += PROF_EVENTS) {
            for (int j = 0; j + 1 < PROF_EVENTS; ++j) {
                float ms = 0.f;
                CU(cudaEventElapsedTime(&ms, v[i + j], v[i + j + 1]));
                s[j] += ms;
            }
            ++n;
        }
        for (auto e : v) cudaEventDestroy(e);
        v.clear();
    }
    for (int j = 0; j + 1 < PROF_EVENTS; ++j) ms6[j] = s[j];
    if (steps) *steps = n;
    return RS_OK;
}

int rs_get_counters(rs_handle *h, uint64_t *kernel_launches, uint64_t *trace_elems_last_step) {
    if (!h) return fail(RS_E_ARG, "null handle");
    CU(cudaSetDevice(h->cfg.device));
    if (kernel_launches) *kernel_launches = h->launches;
    if (trace_elems_last_step) {
        CU(cudaDeviceSynchronize());
        unsigned long long v = 0;
        CU(cudaMemcpy(&v, h->d_trace_elems, sizeof v, cudaMemcpyDeviceToHost));
        *trace_elems_last_step = v;
    }
    return RS_OK;
}

}  // extern "C"
