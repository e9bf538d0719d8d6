/*
 * ranslice_oracle.c -- CPU ORACLE (test infrastructure, see ranslice_oracle.h).
 *
 * Scalar fp64 restatement of the reference step path; every function cites the reference
 * lines it follows (paths relative to the reference root).  Written for fidelity, not speed:
 * operation order of every fp64 expression that feeds a decision follows the Python source,
 * np.mean uses numpy's pairwise summation, round() is half-even.  Build with
 * -ffp-contract=off (no FMA contraction).
 */
#define _GNU_SOURCE
#include "ranslice_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define N_SAMPLES 10001         /* channel_models.py:165 samples[f].shape[1] (incl. NaN col) */
#define TRACE_ROWS 100
#define MAX_UE 64
#define MAX_BURST 64
#define MAX_PRB 256
#define N_MTC_DEV 1000          /* scenario_creator.py:87 */

static const int MTC_REP_SET[7] = {2, 4, 8, 16, 32, 64, 128};                 /* scenario_creator.py:88 */
static const int MTC_PERIOD_SET[8] = {1000, 50000, 10000, 15000, 20000, 25000, 50000, 100000}; /* :89 */

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct { uint32_t key[2]; uint32_t env; /* global env id: counter word 3 */ uint32_t *ctr; /* [S][ORC_N_STREAMS] */ } philox_ctx;

static void px_raw(philox_ctx *p, int slice, int stream, uint32_t out[4]) {
    uint32_t c[4] = {p->ctr[slice * ORC_N_STREAMS + stream]++, (uint32_t)stream, (uint32_t)slice, p->env};
    orc_philox(c, p->key, out);
}
static double px_random(void *ctx, int slice, int stream) {
    uint32_t x[4]; px_raw((philox_ctx *)ctx, slice, stream, x);
    return ((double)(x[0] >> 5) * 67108864.0 + (double)(x[1] >> 6)) / 9007199254740992.0;
}
static double px_exponential(void *ctx, int slice, int stream, double scale) {
    return -log(1.0 - px_random(ctx, slice, stream)) * scale;
}
static int64_t px_integers(void *ctx, int slice, int stream, int64_t n) {
    uint32_t x[4]; px_raw((philox_ctx *)ctx, slice, stream, x);
    return (int64_t)(((uint64_t)x[0] * (uint64_t)n) >> 32);
}
static double px_normal(void *ctx, int slice, int stream, double mu, double sigma) {
    double u1 = px_random(ctx, slice, stream), u2 = px_random(ctx, slice, stream);
    double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
    return mu + sigma * z;
}
static void px_random2(void *ctx, int slice, int stream, double *xy) {
    xy[0] = px_random(ctx, slice, stream); xy[1] = px_random(ctx, slice, stream);
}

/* ------------------------------------------------------------------ state */
typedef struct {
    int type;                   /* 0 CBR, 1 VBR (slice_ran.py:14-15) */
    int fading, step, index;    /* channel_models.py:168 */
    double nominal;
    int64_t remaining;          /* slice_ran.py:222 remaining_time[id] */
    int64_t queue, new_bits, bits;
    double th, p;
    int prbs, e_snr;
    double snr[MAX_PRB];        /* ue.snr = windowed vector, slice_ran.py:44 */
    int ran;                    /* RAN slice of the UE inside its L1 (0 unless the L1 multiplexes several, L1_level=False) */
    int64_t next_arrival;       /* VbrSource.steps_to_next_arrival */
    int nb;
    int64_t togo[MAX_BURST];    /* VbrSource.steps_to_go */
} ue_t;

#define ORC_MAX_RAN 8
typedef struct {                /* SliceRANeMBB (slice_ran.py:150-325) */
    int64_t cbr_next, vbr_next; /* slice_ran.py:185-186 */
    int64_t slot_counter;
    int64_t a_traffic[2], a_th[2], a_prb[2];
    double a_queue[2], a_snr[2];
} ran_t;
typedef struct {                /* SliceL1eMBB (slice_l1.py:128-228): one RAN slice, or n_embb of them when L1_level=False */
    ue_t *ues; int n_ues;
    int n_ran; ran_t ran[ORC_MAX_RAN];
    int i_prb, n_prbs;
} embb_t;

typedef struct {
    int32_t period[N_MTC_DEV], t_arr[N_MTC_DEV], reps[N_MTC_DEV];
    int64_t *q_rep, *q_t0; int q_n, q_cap;
    int32_t *q_id;               /* multiplexed mMTC L1 (L1_level=False): RAN slice of every queued device (slice_l1.py:33,80) */
    int64_t time;
    double a_delay, a_rep; int64_t a_dev;
    int n_prbs;
} mmtc_t;

struct orc_env {
    orc_config cfg;
    orc_tables tbl;
    int S;                       /* L1 slices = action entries */
    int n_l1_embb;               /* eMBB L1 slices: n_embb, or 1 when they are multiplexed (scenario_creator.py:156-177) */
    int n_l1_mmtc;               /* mMTC L1 slices: n_mmtc, or 1 when they are multiplexed (scenario_creator.py:173-176; queue in mmtc[0]) */
    double acc_ran[ORC_MAX_RAN * 2][10];   /* raw accumulators of every RAN slice after the last step, L1-major */
    embb_t *embb; mmtc_t *mmtc;
    orc_rng rng; philox_ctx px; uint32_t *ctr;
    double A, B;                 /* MCSCodeset.compute_factors(0.1) */
    double norm_embb[10], norm_mmtc[3];
    uint32_t flags;
    int lut_mcs[256], lut_rate[256];   /* memoised orc_mcs_lut for e_snr in [-128,127] (pure function of the integer) */
};

/* ------------------------------------------------------------------ leaf math */
static double sigmoid3(double x, double x0, double k) { return 1.0 / (1.0 + exp(-k * (x - x0))); } /* channel_models.py:35-37 */

static void compute_factors(double Delta, double *A, double *B) { /* channel_models.py:272-279 */
    double a = 1.0 / Delta;
    a = a * (log(1.0 / sigmoid3(0.1, 0, 1) - 1.0) - log(1.0 / sigmoid3(0.9, 0, 1) - 1.0));
    *A = a; *B = -log(1.0 / sigmoid3(0.9, 0, 1) - 1.0);
}

static double rx_prob(const orc_tables *t, double A, double B, int mcs, double snr) { /* :281-286 */
    double x = A * (snr - t->mcs_snr[mcs]) - B;
    return sigmoid3(x, 0, 1);
}

void orc_mcs_lut(const orc_tables *t, int e_snr, int *mcs_out, double *bps, int *rate) { /* :288-295 */
    double A, B; compute_factors(0.1, &A, &B);
    double target = 1.0 - 0.1;
    int mcs;
    for (mcs = 0; mcs < 26; ++mcs)
        if (rx_prob(t, A, B, mcs, (double)e_snr) < target) {
            *mcs_out = mcs - 1 > 0 ? mcs - 1 : 0;
            *bps = t->mcs_rate[mcs] * t->mcs_order[mcs];
            *rate = (int)(158 * *bps);                       /* schedulers.py:45 int truncation */
            return;
        }
    mcs = 25;
    *mcs_out = mcs; *bps = t->mcs_rate[mcs] * t->mcs_order[mcs]; *rate = (int)(158 * *bps);
}

static double np_pairwise_sum(const double *a, int n) { /* numpy pairwise_sum (contiguous add.reduce) */
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    } else if (n <= 128) {
        double r[8]; int i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int n2 = n / 2; n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

static const double MI_X0[3] = {-0.25040431, 5.12440916, 9.16962738};   /* channel_models.py:268-270 */
static const double MI_K[3] = {0.31591749, 0.25423209, 0.22298101};

static double response_AB(const orc_tables *t, double A, double B, int mcs, const double *snr, int n) { /* :297-313 */
    double s = snr[0];
    if (n > 1) {
        int m = t->mcs_mod[mcs];
        double mi[MAX_PRB];
        for (int i = 0; i < n; ++i) mi[i] = sigmoid3(snr[i], MI_X0[m], MI_K[m]);
        double avg = np_pairwise_sum(mi, n) / n;
        s = -(1.0 / MI_K[m]) * log(1.0 / avg - 1.0) + MI_X0[m];         /* inv_sigmoid :39-41 */
    }
    return rx_prob(t, A, B, mcs, s);
}
double orc_response(const orc_tables *t, int mcs, const double *snr, int n) {
    double A, B; compute_factors(0.1, &A, &B);
    return response_AB(t, A, B, mcs, snr, n);
}

static double find_y(double x1, double y1, double x2, double y2, double x) { /* :44-48 */
    double m = (y2 - y1) / (x2 - x1);
    double b = -m * x1 + y1;
    return m * x + b;
}
static int in_cell(double x, double y) {                                   /* :50-60, :74 */
    return (y > find_y(0, 0.5, 0.25, 0, x)) && (y > find_y(0.75, 0, 1, 0.5, x)) &&
           (y < find_y(0, 0.5, 0.25, 1, x)) && (y < find_y(0.75, 1, 1, .5, x));
}
double orc_macro_cell(double x, double y, double logf, double A, double B) { /* :62-68, 80-97 */
    double x_t = x - 0.5 / 2;
    double d = sqrt(x_t * x_t + y * y);
    double theta = acos(x_t / d) * (180.0 / M_PI) - 60;
    double R = d * 2 > 0.1 ? d * 2 : 0.1;
    double ap = 12 * ((theta / 65) * (theta / 65));
    double G = 15 + -1 * (ap < 20 ? ap : 20);
    double L = A + B * log10(R);
    double fspl = 20 * log10(2.0) + 92.45 + 2.6 * 10 * log10(R);
    L = L > fspl ? L : fspl;
    double att = L + logf - G;
    double rx = 30 - (att > 70 ? att : 70);
    return rx - (-110) - 9;
}

/* ------------------------------------------------------------------ lifecycle */
int orc_n_variables(const orc_env *e) { return 10 * e->cfg.n_embb + 3 * e->cfg.n_mmtc; }
int orc_n_ues(const orc_env *e, int s) { return e->embb[s].n_ues; }
void orc_get_acc_ran(const orc_env *e, double *acc) { memcpy(acc, e->acc_ran, sizeof(double) * 10 * (size_t)(e->cfg.n_embb + e->cfg.n_mmtc)); }

void orc_set_rng(orc_env *e, const orc_rng *rng) {
    if (rng) { e->rng = *rng; return; }
    e->rng.ctx = &e->px;
    e->rng.random = px_random; e->rng.exponential = px_exponential; e->rng.integers = px_integers;
    e->rng.normal = px_normal; e->rng.random2 = px_random2; e->rng.choice = px_integers;
}

orc_env *orc_create(const orc_config *cfg, const orc_tables *tbl, uint64_t seed, uint32_t env_id) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    e->cfg = *cfg; e->tbl = *tbl;
    e->n_l1_embb = cfg->l1_mux ? (cfg->n_embb > 0) : cfg->n_embb;
    e->n_l1_mmtc = cfg->l1_mux ? (cfg->n_mmtc > 0) : cfg->n_mmtc;
    e->S = e->n_l1_embb + e->n_l1_mmtc;
    e->embb = (embb_t *)calloc(e->n_l1_embb > 0 ? e->n_l1_embb : 1, sizeof(embb_t));
    for (int s = 0; s < e->n_l1_embb; ++s) {
        e->embb[s].ues = (ue_t *)calloc(MAX_UE, sizeof(ue_t));
        e->embb[s].n_ran = cfg->l1_mux ? cfg->n_embb : 1;
    }
    e->mmtc = (mmtc_t *)calloc(cfg->n_mmtc > 0 ? cfg->n_mmtc : 1, sizeof(mmtc_t));
    e->ctr = (uint32_t *)calloc((size_t)(cfg->n_embb + cfg->n_mmtc + 1) * ORC_N_STREAMS, sizeof(uint32_t));
    e->px.key[0] = (uint32_t)seed; e->px.key[1] = (uint32_t)(seed >> 32); e->px.env = env_id; e->px.ctr = e->ctr;
    orc_set_rng(e, NULL);
    compute_factors(0.1, &e->A, &e->B);
    for (int i = 0; i < 256; ++i) { double bps; orc_mcs_lut(tbl, i - 128, &e->lut_mcs[i], &bps, &e->lut_rate[i]); }
    double tps = cfg->slots_per_step * 1e-3; int sps = cfg->slots_per_step;   /* scenario_creator.py:106,115-134 */
    double ne[10] = {5e6 * tps, 10e6 * tps, 25.0 * sps, 10e4 * sps, 35.0 * sps,
                     5e6 * tps, 10e6 * tps, 35.0 * sps, 10e4 * sps, 35.0 * sps};
    memcpy(e->norm_embb, ne, sizeof ne);
    for (int i = 0; i < 3; ++i) e->norm_mmtc[i] = 100.0 * sps;
    for (int s = 0; s < e->n_l1_embb; ++s) e->embb[s].n_prbs = 20;         /* scenario_creator.py:160 */
    for (int s = 0; s < cfg->n_mmtc; ++s) e->mmtc[s].n_prbs = 5;           /* :165 */
    return e;
}

void orc_destroy(orc_env *e) {
    if (!e) return;
    for (int s = 0; s < e->n_l1_embb; ++s) free(e->embb[s].ues);
    for (int s = 0; s < e->cfg.n_mmtc; ++s) { free(e->mmtc[s].q_rep); free(e->mmtc[s].q_t0); free(e->mmtc[s].q_id); }
    free(e->embb); free(e->mmtc); free(e->ctr); free(e);
}

static void embb_reset_info(embb_t *l1) {                                   /* slice_ran.py:270-273, every RAN slice of the L1 */
    for (int r = 0; r < l1->n_ran; ++r) {
        ran_t *s = &l1->ran[r];
        for (int k = 0; k < 2; ++k) { s->a_traffic[k] = s->a_th[k] = s->a_prb[k] = 0; s->a_queue[k] = s->a_snr[k] = 0.0; }
        s->slot_counter = 0;
    }
}
static void mmtc_reset_info(mmtc_t *m) { m->a_delay = 0; m->a_rep = 0; m->a_dev = 0; }   /* :123-125 */

void orc_reset(orc_env *e, float *obs) {                                    /* node_b.py:17-22 */
    for (int s = 0; s < e->n_l1_embb; ++s) {                                /* slice_l1.py:145-148, slice_ran.py:182-190 */
        embb_t *sl = &e->embb[s];
        sl->n_ues = 0;
        for (int r = 0; r < sl->n_ran; ++r) { sl->ran[r].cbr_next = 0; sl->ran[r].vbr_next = 0; }
        embb_reset_info(sl);
    }
    for (int s = 0; s < e->cfg.n_mmtc; ++s) {                               /* slice_l1.py:29-39, slice_ran.py:91-101 */
        mmtc_t *m = &e->mmtc[s]; int gs = e->n_l1_embb + s;
        m->q_n = 0; m->time = 0; mmtc_reset_info(m);
        for (int i = 0; i < N_MTC_DEV; ++i) {
            m->reps[i] = MTC_REP_SET[e->rng.choice(e->rng.ctx, gs, ORC_STREAM_MTC, 7)];
            m->period[i] = MTC_PERIOD_SET[e->rng.choice(e->rng.ctx, gs, ORC_STREAM_MTC, 8)];
            m->t_arr[i] = 1 + (int32_t)e->rng.choice(e->rng.ctx, gs, ORC_STREAM_MTC, m->period[i]);
        }
    }
    if (obs) memset(obs, 0, sizeof(float) * orc_n_variables(e));
}

/* ------------------------------------------------------------------ eMBB slot */
static int64_t exp_slots(orc_env *e, int s, int stream, double scale, int div_slot) {
    double x = e->rng.exponential(e->rng.ctx, s, stream, scale);
    return (int64_t)rint(div_slot ? x / 1e-3 : x);     /* np.rint(x / slot_length) */
}

static int cbr_cac(const ran_t *sl) {                                       /* slice_ran.py:195-203 */
    int64_t slots = sl->slot_counter > 1 ? sl->slot_counter : 1;
    double time = slots * 1e-3;
    double cbr_prb = (double)sl->a_prb[0] / (double)slots;
    double cbr_th = (double)sl->a_th[0] / time;
    return !(cbr_prb >= 20 || cbr_th >= 10e6);
}

static void insert_user(orc_env *e, int s, ue_t *u) {                       /* channel_models.py:163-169 */
    u->fading = (int)e->rng.integers(e->rng.ctx, s, ORC_STREAM_CHAN, 3);
    u->index = (int)e->rng.integers(e->rng.ctx, s, ORC_STREAM_CHAN, N_SAMPLES);
    u->step = e->rng.choice(e->rng.ctx, s, ORC_STREAM_CHAN, 2) ? 1 : -1;   /* rng.choice([-1,1]) */
    double xy[2];
    do { e->rng.random2(e->rng.ctx, s, ORC_STREAM_CHAN, xy); } while (!in_cell(xy[0], xy[1]));   /* :70-76 */
    double logf = e->rng.normal(e->rng.ctx, s, ORC_STREAM_CHAN, 0.0, 10.0);
    u->nominal = orc_macro_cell(xy[0], xy[1], logf, e->cfg.prop_A, e->cfg.prop_B);
}

static void new_ue(ue_t *u, int type) {                                     /* slice_ran.py:23-39 */
    memset(u, 0, sizeof(*u)); u->type = type;
}

#ifdef ORC_PF_STATS
unsigned long long orc_pf_stat[12];
#endif
static void pf_allocate(orc_env *e, embb_t *sl) {                           /* schedulers.py:21-76 */
    int n = sl->n_ues, n_prb = sl->n_prbs;
    int rbs[MAX_UE], mcs[MAX_UE]; int64_t queue[MAX_UE], rate[MAX_UE], bits[MAX_UE]; double th[MAX_UE];
    const double b = 1.0 / 50, a = 1 - b, slot = 1e-3;
    for (int i = 0; i < n; ++i) {
        ue_t *u = &sl->ues[i];
        rbs[i] = 0; bits[i] = 0;
        th[i] = u->th > 1 ? u->th : 1;
        queue[i] = u->queue;
        int es = u->e_snr < -128 ? -128 : (u->e_snr > 127 ? 127 : u->e_snr);   /* LUT saturates far inside this range */
        mcs[i] = e->lut_mcs[es + 128]; rate[i] = e->lut_rate[es + 128];
    }
#ifdef ORC_PF_STATS
    int prev = -1, prev2 = -1, prev3 = -1;                                  /* winner / runner-up / third at the start of the current run */
    {   /* What-if: chunks handed out in BATCHES.  Every backlogged UE speculates its next R chunks on its own; event (k, j) = "UE k gets
         * its j-th chunk" has key min(head metrics before it); the greedy order of the RB loop is the order of decreasing keys, so every
         * event whose key is above the largest key any UE could still produce beyond its horizon can be applied at once (counts only:
         * a UE's state depends on how many chunks it got, not on the interleaving).  No such event: one run of the current loop.
         * stat[6] = warp-wide steps of that scheme, stat[7] = steps that needed a selection of the top-N events (budget cut). */
        extern unsigned long long orc_pf_stat[12];
        enum { R = ORC_PF_SPEC };
        double th2[MAX_UE]; int64_t q2[MAX_UE], b2[MAX_UE];
        for (int i = 0; i < n; ++i) { th2[i] = th[i]; q2[i] = queue[i]; b2[i] = 0; }
        int nrem = n_prb / 2;
        for (;;) {
            int nb = 0;
            for (int i = 0; i < n; ++i) nb += q2[i] > 0;
            if (nrem <= 0 || nb < 2) break;
            double key[MAX_UE][R], T = 0.0; int nev[MAX_UE];
#ifndef ORC_PF_NMIN
#define ORC_PF_NMIN 0
#define ORC_PF_BMIN 2
#endif
            const int use_batch = nrem >= ORC_PF_NMIN && nb >= ORC_PF_BMIN;
            for (int i = 0; i < n; ++i) {
                nev[i] = 0;
                if (q2[i] <= 0) continue;
                double t = th2[i], head = (double)rate[i] / t, kmin = head; int64_t q = q2[i], bb = b2[i];
                int j = 0;
#ifdef ORC_PF_NOCUT
                const int depth = nrem / nb < R ? nrem / nb : R;            /* nb * depth <= nrem: a batch never exceeds the budget */
#else
                const int depth = R;
#endif
                for (; j < depth; ++j) {
                    key[i][j] = kmin;
                    int64_t tx = 2 * rate[i] < q ? 2 * rate[i] : q;
                    q -= tx; bb += tx; t = a * t + b * (double)bb / slot;
                    head = q > 0 ? (double)rate[i] / t : 0.0;
                    if (head < kmin) kmin = head;
                    if (q <= 0) { ++j; break; }
                }
                nev[i] = j;
                if (q > 0 && kmin > T) T = kmin;                            /* more events beyond the horizon: all with key <= kmin */
            }
            int count = 0;
            for (int i = 0; i < n; ++i) for (int j = 0; j < nev[i]; ++j) count += key[i][j] > T;
            if (!use_batch) count = 0;
            orc_pf_stat[6] += 1;
            if (use_batch) orc_pf_stat[8] += 1;                             /* batch attempts */
            if (use_batch && count == 0) orc_pf_stat[9] += 1;               /* ... that found nothing to apply */
            if (count > nrem) {                                             /* budget cut: the nrem largest keys */
                orc_pf_stat[7] += 1;
                double all[MAX_UE * R]; int m = 0;
                for (int i = 0; i < n; ++i) for (int j = 0; j < nev[i]; ++j) if (key[i][j] > T) all[m++] = key[i][j];
                for (int x = 0; x < m; ++x) for (int y = x + 1; y < m; ++y) if (all[y] > all[x]) { double z = all[x]; all[x] = all[y]; all[y] = z; }
                T = all[nrem];                                              /* events with key > the (nrem+1)-th largest (ties: fewer) */
                count = 0;
                for (int i = 0; i < n; ++i) for (int j = 0; j < nev[i]; ++j) count += key[i][j] > T;
            }
            if (count == 0) {                                               /* one run of the current loop */
                int idx = 0; double best = -1.0, second = -1.0;
                for (int i = 0; i < n; ++i) { double m = q2[i] > 0 ? (double)rate[i] / th2[i] : 0.0; if (m > best) { second = best; best = m; idx = i; } else if (m > second) second = m; }
                do {
                    int64_t tx = 2 * rate[idx] < q2[idx] ? 2 * rate[idx] : q2[idx];
                    q2[idx] -= tx; b2[idx] += tx; th2[idx] = a * th2[idx] + b * (double)b2[idx] / slot; --nrem;
                } while (nrem > 0 && q2[idx] > 0 && (double)rate[idx] / th2[idx] > second);
                continue;
            }
            for (int i = 0; i < n; ++i) {
                int c = 0;
                for (int j = 0; j < nev[i]; ++j) c += key[i][j] > T;
                for (int j = 0; j < c; ++j) {
                    int64_t tx = 2 * rate[i] < q2[i] ? 2 * rate[i] : q2[i];
                    q2[i] -= tx; b2[i] += tx; th2[i] = a * th2[i] + b * (double)b2[i] / slot;
                }
                nrem -= c;
            }
        }
    }
#endif
    for (int r = 0; r < n_prb; r += 2) {
        int prbs = n_prb - r < 2 ? n_prb - r : 2;
        int idx = 0; double best = -INFINITY;
        for (int i = 0; i < n; ++i) {                                       /* np.argmax: first max */
            double m = (double)(rate[i] * (queue[i] > 0)) / th[i];
            if (m > best) { best = m; idx = i; }
        }
#ifdef ORC_PF_STATS
        {   /* statistics for the design of the warp kernel's PF loop: how often could the next winner
             * and runner-up be told from the top three of the previous warp-wide argmax alone? */
            extern unsigned long long orc_pf_stat[12];
            int nb = 0;
            for (int i = 0; i < n; ++i) nb += queue[i] > 0;
            orc_pf_stat[0] += 1;                                            /* chunks */
            if (nb >= 2) orc_pf_stat[1] += 1;                               /* contended chunks */
            if (nb >= 2 && idx != prev) {
                orc_pf_stat[2] += 1;                                        /* runs (warp-wide iterations of the current kernel) */
                if (prev >= 0) {
                    orc_pf_stat[3] += 1;                                    /* winner changes */
                    double mp = (double)(rate[prev] * (queue[prev] > 0)) / th[prev], v4 = 0.0;
                    for (int i = 0; i < n; ++i)
                        if (i != prev && i != prev2 && i != prev3) { double m = (double)(rate[i] * (queue[i] > 0)) / th[i]; if (m > v4) v4 = m; }
                    if (idx == prev2) orc_pf_stat[4] += 1;                  /* ... to the old runner-up */
                    if (idx == prev2 && mp >= v4) orc_pf_stat[5] += 1;      /* ... and the top-3 set is unchanged */
                }
                /* top three of this argmax */
                int i2 = -1, i3 = -1; double b2 = -1.0, b3 = -1.0;
                for (int i = 0; i < n; ++i) if (i != idx) { double m = (double)(rate[i] * (queue[i] > 0)) / th[i]; if (m > b2) { b2 = m; i2 = i; } }
                for (int i = 0; i < n; ++i) if (i != idx && i != i2) { double m = (double)(rate[i] * (queue[i] > 0)) / th[i]; if (m > b3) { b3 = m; i3 = i; } }
                prev = idx; prev2 = i2; prev3 = i3;
            }
        }
#endif
        rbs[idx] += prbs;
        int64_t tx = prbs * rate[idx] < queue[idx] ? prbs * rate[idx] : queue[idx];
        queue[idx] -= tx; bits[idx] += tx;
        th[idx] = a * th[idx] + b * (double)bits[idx] / slot;
    }
    int o = 0;
    for (int i = 0; i < n; ++i) {
        ue_t *u = &sl->ues[i];
        u->prbs = rbs[i]; u->bits = bits[i];
        u->p = rbs[i] ? response_AB(&e->tbl, e->A, e->B, mcs[i], u->snr + o, rbs[i]) : 0.0;
        o += rbs[i];
    }
}

static void embb_slot(orc_env *e, int s) {                                  /* slice_l1.py:193-228 */
    embb_t *sl = &e->embb[s];
    /* --- for slice_ran in self.slices_ran: slot(), extract_users(departures), add_users(arrivals)  (slice_l1.py:195-198).
     * RAN draws come from the stream of the RAN slice (slice index r when the L1 multiplexes several, else s); the
     * channel, reception and VbrSource draws from the streams of the L1 (slice index s). */
    for (int r = 0; r < sl->n_ran; ++r) {
        ran_t *rn = &sl->ran[r];
        const int rs = e->cfg.l1_mux ? r : s;
        ue_t arrivals[2]; int n_arr = 0;
        /* --- slice_ran.slot(), slice_ran.py:263-268 */
        rn->slot_counter += 1;
        if (rn->cbr_next == 0) {                                            /* :205-227 */
            rn->cbr_next = exp_slots(e, rs, ORC_STREAM_RAN, 1.0 / (2.0 / 60.0), 1);
            if (cbr_cac(rn)) {
                ue_t *u = &arrivals[n_arr++]; new_ue(u, 0);
                u->remaining = exp_slots(e, rs, ORC_STREAM_RAN, 30.0, 1);
            }
        } else rn->cbr_next -= 1;
        if (rn->vbr_next == 0) {                                            /* :229-249 */
            ue_t *u = &arrivals[n_arr++]; new_ue(u, 1);
            u->next_arrival = exp_slots(e, s, ORC_STREAM_VBR, (1.0 / 1) / 1e-3, 0);   /* traffic_generators.py:65-66 */
            u->remaining = exp_slots(e, rs, ORC_STREAM_RAN, 30.0, 1);
            rn->vbr_next = exp_slots(e, rs, ORC_STREAM_RAN, 1.0 / (5.0 / 60.0), 1);
        } else rn->vbr_next -= 1;
        /* departures(), :251-261 -- live timers of THIS RAN slice in arrival order, then this slot's arrivals */
        int w = 0;
        for (int i = 0; i < sl->n_ues; ++i) {
            if (sl->ues[i].ran == r) {
                sl->ues[i].remaining -= 1;
                if (sl->ues[i].remaining == 0) continue;                    /* extract_users keeps the order of the rest */
            }
            if (w != i) sl->ues[w] = sl->ues[i];
            ++w;
        }
        sl->n_ues = w;
        for (int k = 0; k < n_arr; ++k) {                                   /* slice_l1.py:183-186 */
            arrivals[k].remaining -= 1;
            if (arrivals[k].remaining == 0) { e->flags |= 8u; continue; }   /* reference crashes here */
            if (sl->n_ues >= MAX_UE) { e->flags |= 1u; continue; }
            insert_user(e, s, &arrivals[k]);
            arrivals[k].ran = r;
            sl->ues[sl->n_ues++] = arrivals[k];
        }
    }
    /* --- per-UE traffic + SNR estimate, slice_l1.py:200-213 */
    int64_t queued = 0;
    const double *trace = e->tbl.trace;
    for (int i = 0; i < sl->n_ues; ++i) {
        ue_t *u = &sl->ues[i];
        if (u->type == 0) u->new_bits = 500;                                /* CbrSource: 500000*1e-3 every slot */
        else {                                                              /* VbrSource.step, traffic_generators.py:70-99 */
            int64_t bits = 0; int k = 0;
            for (int j = 0; j < u->nb; ++j) {
                u->togo[j] -= 1;
                if (u->togo[j] == 0) continue;                              /* ending: contributes 0 */
                bits += 1000;
                u->togo[k++] = u->togo[j];
            }
            u->nb = k;
            u->next_arrival -= 1;
            if (u->next_arrival == 0) {
                int64_t len = exp_slots(e, s, ORC_STREAM_VBR, 500.0, 0);
                if (u->nb < MAX_BURST) u->togo[u->nb++] = len; else e->flags |= 2u;
                u->next_arrival = exp_slots(e, s, ORC_STREAM_VBR, 1000.0, 0);
            }
            u->new_bits = bits;
        }
        u->queue += u->new_bits;
        queued += u->queue;
        if (sl->n_prbs > 0) {
            for (;;) {                                                      /* get_snr, channel_models.py:171-191 */
                u->index += u->step;
                if (u->index >= N_SAMPLES || u->index < 0) {
                    u->index = (int)e->rng.integers(e->rng.ctx, s, ORC_STREAM_CHAN, N_SAMPLES);
                    u->step = e->rng.choice(e->rng.ctx, s, ORC_STREAM_CHAN, 2) ? 1 : -1;   /* rng.choice([-1,1]) */
                }
                if (u->index != N_SAMPLES - 1) break;                       /* column 10000 is NaN */
            }
            const double *col = trace + ((size_t)u->fading * N_SAMPLES + u->index) * TRACE_ROWS;
            for (int j = 0; j < sl->n_prbs; ++j) u->snr[j] = col[(sl->i_prb + j) % TRACE_ROWS] + u->nominal;
            u->e_snr = (int)rint(np_pairwise_sum(u->snr, sl->n_prbs) / sl->n_prbs);   /* slice_ran.py:43-45 */
        }
    }
    /* --- scheduling + reception, slice_l1.py:215-224 */
    if (queued > 0 && sl->n_prbs > 0) {
        pf_allocate(e, sl);
        const double b = 1.0 / 50, a = 1 - b;
        for (int i = 0; i < sl->n_ues; ++i) {
            ue_t *u = &sl->ues[i];
            int received = 0;
            if (u->prbs) received = e->rng.random(e->rng.ctx, s, ORC_STREAM_L1RX) < u->p;
            if (!received) u->bits = 0;                                     /* slice_ran.py:51-55 */
            u->queue = u->queue - u->bits > 0 ? u->queue - u->bits : 0;
            u->th = a * u->th + b * (double)u->bits / 1e-3;
        }
    }
    /* --- update_info of every RAN slice over its own UEs, slice_ran.py:278-305 */
    for (int r = 0; r < sl->n_ran; ++r) {
        ran_t *rn = &sl->ran[r];
        for (int t = 0; t < 2; ++t) {
            int64_t q = 0, snr = 0, n = 0;
            for (int i = 0; i < sl->n_ues; ++i) {
                ue_t *u = &sl->ues[i];
                if (u->type != t || u->ran != r) continue;
                rn->a_traffic[t] += u->new_bits; rn->a_th[t] += u->bits; rn->a_prb[t] += u->prbs;
                q += u->queue; snr += u->e_snr; n += 1;
            }
            if (n < 1) n = 1;
            rn->a_queue[t] += (double)q / (double)n;
            rn->a_snr[t] += (double)snr / (double)n;
        }
    }
}

/* ------------------------------------------------------------------ mMTC slot */
static void mmtc_slot(mmtc_t *m) {                                          /* slice_l1.py:87-125, slice_ran.py:103-121 */
    m->time += 1;
    for (int i = 0; i < N_MTC_DEV; ++i) m->t_arr[i] -= 1;
    for (int i = 0; i < N_MTC_DEV; ++i)
        if (m->t_arr[i] == 0) {
            if (m->q_n == m->q_cap) {
                m->q_cap = m->q_cap ? 2 * m->q_cap : 256;
                m->q_rep = (int64_t *)realloc(m->q_rep, sizeof(int64_t) * m->q_cap);
                m->q_t0 = (int64_t *)realloc(m->q_t0, sizeof(int64_t) * m->q_cap);
            }
            m->q_rep[m->q_n] = m->reps[i]; m->q_t0[m->q_n] = m->time; m->q_n++;
            m->t_arr[i] = m->period[i];
        }
    int n_tx = m->n_prbs < m->q_n ? m->n_prbs : m->q_n;
    for (int k = 0; k < n_tx; ++k) m->q_rep[k] -= 1;
    int w = 0;
    for (int k = 0; k < m->q_n; ++k)
        if (m->q_rep[k] > 0) { m->q_rep[w] = m->q_rep[k]; m->q_t0[w] = m->q_t0[k]; ++w; }
    m->q_n = w;
    double delay = 0, avg_rep = 0;
    if (w > 0) {
        int64_t sd = 0, sr = 0;
        for (int k = 0; k < w; ++k) { int64_t d = m->time - m->q_t0[k]; sd += d > 0 ? d : 0; sr += m->q_rep[k]; }
        delay = (double)sd / (double)w;
        avg_rep = rint((double)sr / (double)w);
    }
    m->a_delay += delay; m->a_rep += avg_rep; m->a_dev += w;               /* slice_ran.py:139-142 */
}

/* SliceL1mMTC.slot with SEVERAL RAN slices (create_env(L1_level=False), scenario_creator.py:173-176; slice_l1.py:87-125): one
 * queue for all of them (kept in mmtc[0]), arrivals appended RAN slice by RAN slice, the first n_prbs queued devices
 * transmit whatever their slice, statistics per RAN slice over its own devices. */
static void mmtc_slot_mux(orc_env *e) {
    const int M = e->cfg.n_mmtc;
    mmtc_t *q = &e->mmtc[0];
    q->time += 1;
    for (int m = 0; m < M; ++m) {
        mmtc_t *sl = &e->mmtc[m];
        for (int i = 0; i < N_MTC_DEV; ++i) sl->t_arr[i] -= 1;
        for (int i = 0; i < N_MTC_DEV; ++i)
            if (sl->t_arr[i] == 0) {
                if (q->q_n == q->q_cap) {
                    q->q_cap = q->q_cap ? 2 * q->q_cap : 256;
                    q->q_rep = (int64_t *)realloc(q->q_rep, sizeof(int64_t) * q->q_cap);
                    q->q_t0 = (int64_t *)realloc(q->q_t0, sizeof(int64_t) * q->q_cap);
                    q->q_id = (int32_t *)realloc(q->q_id, sizeof(int32_t) * q->q_cap);
                }
                q->q_rep[q->q_n] = sl->reps[i]; q->q_t0[q->q_n] = q->time; q->q_id[q->q_n] = m; q->q_n++;
                sl->t_arr[i] = sl->period[i];
            }
    }
    int n_tx = q->n_prbs < q->q_n ? q->n_prbs : q->q_n;
    for (int k = 0; k < n_tx; ++k) q->q_rep[k] -= 1;
    int w = 0;
    for (int k = 0; k < q->q_n; ++k)
        if (q->q_rep[k] > 0) { q->q_rep[w] = q->q_rep[k]; q->q_t0[w] = q->q_t0[k]; q->q_id[w] = q->q_id[k]; ++w; }
    q->q_n = w;
    for (int m = 0; m < M; ++m) {
        int64_t sd = 0, sr = 0; int n = 0;
        for (int k = 0; k < w; ++k)
            if (q->q_id[k] == m) { int64_t d = q->time - q->q_t0[k]; sd += d > 0 ? d : 0; sr += q->q_rep[k]; ++n; }
        double delay = 0, avg_rep = 0;
        if (n > 0) { delay = (double)sd / (double)n; avg_rep = rint((double)sr / (double)n); }
        e->mmtc[m].a_delay += delay; e->mmtc[m].a_rep += avg_rep; e->mmtc[m].a_dev += n;
    }
}

/* ------------------------------------------------------------------ step */
uint32_t orc_step(orc_env *e, const int64_t *action, float *obs, double *reward, int32_t *labels,
                  int32_t *violations, double *acc) {
    const orc_config *c = &e->cfg;
    e->flags = 0;
    for (int s = 0; s < e->n_l1_embb; ++s) embb_reset_info(&e->embb[s]);   /* node_b.py:64 */
    for (int s = 0; s < c->n_mmtc; ++s) mmtc_reset_info(&e->mmtc[s]);
    int64_t i_prb = 0, asum = 0;
    for (int s = 0; s < e->S; ++s) {                                        /* node_b.py:71-74 */
        int64_t a = action[s]; asum += a;
        if (a < 0) { a = 0; e->flags |= 4u; }
        if (i_prb + a > c->n_prbs) { a = c->n_prbs - i_prb; e->flags |= 4u; }   /* reference: undefined */
        if (s < e->n_l1_embb) { e->embb[s].i_prb = (int)i_prb; e->embb[s].n_prbs = (int)a; }
        else e->mmtc[s - e->n_l1_embb].n_prbs = (int)a;                   /* (a multiplexed mMTC L1 keeps its PRBs in mmtc[0]) */
        i_prb += a;
    }
    for (int t = 0; t < c->slots_per_step; ++t) {                           /* node_b.py:77-78, 35-38 */
        for (int s = 0; s < e->n_l1_embb; ++s) embb_slot(e, s);
        if (c->l1_mux && c->n_mmtc > 1) mmtc_slot_mux(e);
        else for (int s = 0; s < c->n_mmtc; ++s) mmtc_slot(&e->mmtc[s]);
    }
    int64_t tv = 0; int v = 0;
    const double obs_time = c->slots_per_step * 1e-3;                       /* slice_ran.py:165 */
    int row = 0;
    for (int s = 0; s < e->n_l1_embb; ++s) {                                /* slice_ran.py:307-325 per RAN slice, slice_l1.py:160-181 per L1 */
        embb_t *l1 = &e->embb[s];
        int l1_viol = 0;
        for (int r = 0; r < l1->n_ran; ++r, ++row) {
            ran_t *sl = &l1->ran[r];
            double a[10] = {(double)sl->a_traffic[0], (double)sl->a_th[0], (double)sl->a_prb[0], sl->a_queue[0], sl->a_snr[0],
                            (double)sl->a_traffic[1], (double)sl->a_th[1], (double)sl->a_prb[1], sl->a_queue[1], sl->a_snr[1]};
            for (int j = 0; j < 10; ++j) { obs[v++] = (float)(a[j] / e->norm_embb[j]); e->acc_ran[row][j] = a[j]; if (acc && r == 0) acc[s * 10 + j] = a[j]; }
            int cbr_ok = a[1] / obs_time > 10e6 || a[2] / c->slots_per_step > 20 || a[3] / c->slots_per_step < 10e4;
            int vbr_ok = a[6] / obs_time > 15e6 || a[7] / c->slots_per_step > 30 || a[8] / c->slots_per_step < 15e4;
            l1_viol += !(cbr_ok && vbr_ok);
        }
        violations[s] = l1_viol; labels[s] = l1_viol ? -1 : 1; tv += l1_viol;   /* slice_l1.py:160-171: sum of the RAN slices' violations */
    }
    const int mtc_mux = c->l1_mux && c->n_mmtc > 1;
    int mux_viol = 0;
    for (int s = 0; s < c->n_mmtc; ++s) {                                   /* slice_ran.py:133-148 */
        mmtc_t *m = &e->mmtc[s]; int gs = e->n_l1_embb + (mtc_mux ? 0 : s);     /* L1 row of this RAN slice */
        double a[3] = {(double)m->a_dev, m->a_rep, m->a_delay};
        for (int j = 0; j < 10; ++j) e->acc_ran[row][j] = j < 3 ? a[j] : 0.0;
        ++row;
        for (int j = 0; j < 3; ++j) { obs[v++] = (float)(a[j] / e->norm_mmtc[j]); if (acc && (!mtc_mux || s == 0)) acc[gs * 10 + j] = a[j]; }
        if (acc && (!mtc_mux || s == 0)) for (int j = 3; j < 10; ++j) acc[gs * 10 + j] = 0;
        int viol = !(m->a_delay / c->slots_per_step < 300);
        if (mtc_mux) mux_viol += viol;                                      /* slice_l1.py:65-75: the L1 adds its RAN slices' violations up */
        else { violations[gs] = viol; labels[gs] = viol ? -1 : 1; tv += viol; }
    }
    if (mtc_mux) { int gs = e->n_l1_embb; violations[gs] = mux_viol; labels[gs] = mux_viol ? -1 : 1; tv += mux_viol; }
    if (tv > 0) *reward = -1 * c->penalty * (double)tv;                     /* ran_slice.py:45-52 */
    else *reward = (double)(c->n_prbs - asum > 0 ? c->n_prbs - asum : 0);
    return e->flags;
}

/* batched convenience: envs are independent; pthreads pull env indices from a shared counter */
typedef struct { orc_env **envs; int n; const int64_t *actions; float *obs; double *reward;
                 int32_t *labels, *violations; uint32_t *flags; int next; pthread_mutex_t mu; } batch_job;

static void *batch_worker(void *arg) {
    batch_job *j = (batch_job *)arg;
    int S = j->envs[0]->S, V = orc_n_variables(j->envs[0]);
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int lo = j->next; j->next += 4;
        pthread_mutex_unlock(&j->mu);
        if (lo >= j->n) break;
        int hi = lo + 4 < j->n ? lo + 4 : j->n;
        for (int i = lo; i < hi; ++i) {
            uint32_t f = orc_step(j->envs[i], j->actions + (size_t)i * S, j->obs + (size_t)i * V, j->reward + i,
                                  j->labels + (size_t)i * S, j->violations + (size_t)i * S, NULL);
            if (j->flags) j->flags[i] = f;
        }
    }
    return NULL;
}

void orc_step_batch(orc_env **envs, int n, int n_threads, const int64_t *actions, float *obs, double *reward,
                    int32_t *labels, int32_t *violations, uint32_t *flags) {
    if (n <= 0) return;
    batch_job j = {envs, n, actions, obs, reward, labels, violations, flags, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads <= 1) { batch_worker(&j); return; }
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, batch_worker, &j);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
}
