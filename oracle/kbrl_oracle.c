/*
 * kbrl_oracle.c -- CPU ORACLE for kernel #2 (KBRL inner loop).  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's
 *   GaussianKernel.k_eval / k / predict      algorithms/kernel.py:8-28
 *   SVvariable.extend / update / insert      algorithms/projectron.py:3-21
 *   Projectron.predict / update              algorithms/projectron.py:32-60
 *   KBRL_Control.select_action / adjust_action / update_control   kbrl_control.py:41-114
 * including its dtype quirks: with a single landmark the kernel value, the coefficient and K^-1 are
 * float32 (kernel.py:15-16, projectron.py:10,59); from the second landmark on everything is float64.
 * np.sum over the 11 (or 4) squared differences follows numpy's pairwise order, in which the action
 * coordinate is added last.  Dot products are accumulated in index order (numpy defers to BLAS, whose
 * order is unspecified; decisions only depend on signs / thresholds, see tests/test_kbrl_oracle.py).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int d, D, cap;
    double *lm;      /* [cap][d] landmarks */
    double *coeff;   /* [cap]; float32-valued while D == 1 */
    double *kinv;    /* [cap][cap] */
    double *kf;      /* [cap] K_f of the last predict */
    double f;        /* f of the last predict */
    int off;         /* first state variable of this learner (Learner.indexes, kbrl_control.py:18) */
} learner_t;

typedef struct orc_kb {
    int S, n_prbs;
    double alfa, acc_lo, acc_hi, gamma, eta;
    learner_t *L;
    int64_t *action, *sec, *margins;
    double *acc;     /* [S][n_prbs] accuracies (kbrl_control.py:38-39) */
    int adjusted;
    long tie_breaks; /* kernel.py:26-27 random tie breaks taken (f == 0 with a non-empty dictionary) */
    /* tie-break stream: np.random.choice([-1, 1]) -> Philox (key = seed, counter = (n, STREAM_KBRL = 5, learner, env id)),
     * integers(2) -> index (RNG contract: network-slicing_b200/philox.py); off until orc_kb_set_tie_stream */
    int tie_on; uint32_t tie_key[2], tie_env; uint32_t *tie_ctr; /* [S] */
    int cur;         /* learner whose predict() is running */
    int plus;        /* 1: ProjectronPlus.update (algorithms/projectron.py:66-107) instead of Projectron.update */
} orc_kb;

void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);   /* ranslice_oracle.c */

static double sq_dist(const double *l, const double *x, int n) {   /* ((l - x)**2).sum(): numpy pairwise order */
    double a[16];
    for (int i = 0; i < n; ++i) a[i] = (l[i] - x[i]) * (l[i] - x[i]);
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    double res = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    for (int i = 8; i < n; ++i) res += a[i];
    return res;
}

static void grow(learner_t *h) {
    int nc = h->cap ? 2 * h->cap : 16;
    double *lm = (double *)calloc((size_t)nc * h->d, sizeof(double));
    double *cf = (double *)calloc(nc, sizeof(double));
    double *ki = (double *)calloc((size_t)nc * nc, sizeof(double));
    double *kf = (double *)calloc(nc, sizeof(double));
    for (int i = 0; i < h->D; ++i) {
        memcpy(lm + (size_t)i * h->d, h->lm + (size_t)i * h->d, sizeof(double) * h->d);
        cf[i] = h->coeff[i];
        kf[i] = h->kf[i];
        for (int j = 0; j < h->D; ++j) ki[(size_t)i * nc + j] = h->kinv[(size_t)i * h->cap + j];
    }
    free(h->lm); free(h->coeff); free(h->kinv); free(h->kf);
    h->lm = lm; h->coeff = cf; h->kinv = ki; h->kf = kf; h->cap = nc;
}

/* Projectron.predict (projectron.py:32-37) -> GaussianKernel.predict (kernel.py:22-28) */
static int predict(orc_kb *kb, learner_t *h, const double *x) {
    if (h->D == 0) { h->f = 0.0; h->kf[0] = 0.0; return 0; }
    if (h->D == 1) {                                         /* kernel.py:15-16: float32 kernel value */
        float k = (float)exp(-kb->gamma * sq_dist(h->lm, x, h->d));
        float f = k * (float)h->coeff[0];
        h->kf[0] = (double)k; h->f = (double)f;
    } else {
        double f = 0.0;
        for (int j = 0; j < h->D; ++j) {
            h->kf[j] = exp(-kb->gamma * sq_dist(h->lm + (size_t)j * h->d, x, h->d));
            f += h->kf[j] * h->coeff[j];
        }
        h->f = f;
    }
    if (h->f > 0) return 1;
    if (h->f < 0) return -1;
    kb->tie_breaks++;                                        /* np.random.choice([-1,1]) in the reference */
    if (kb->tie_on) {
        uint32_t c[4] = {kb->tie_ctr[kb->cur]++, 5u, (uint32_t)kb->cur, kb->tie_env}, out[4];
        orc_philox(c, kb->tie_key, out);
        return (((uint64_t)out[0] * 2u) >> 32) ? 1 : -1;     /* [-1, 1][integers(2)] */
    }
    return -1;
}

/* Projectron.update (projectron.py:39-60) */
static void update(orc_kb *kb, learner_t *h, const double *x, int y) {
    const double margin = (double)y * h->f;                  /* np.int64 * float32 / float64 -> float64 */
    const int inside = kb->plus && margin < 1 && margin > 0; /* ProjectronPlus: right side of the margin (projectron.py:72) */
    if (!inside && !(margin <= 0)) return;
    const double Kii = 1.0;                                  /* k_eval(x, x) = exp(-gamma * 0) */
    int D = h->D;
    if (D + 1 >= h->cap) grow(h);
    double dstar[4096];
    double dot = 0.0;
    if (D <= 1) {                                            /* float32 stage (D == 0: Kinv = [0.0], K_f = [0.0]) */
        float ki = D ? (float)h->kinv[0] : 0.0f, kf = (float)h->kf[0];
        float ds = ki * kf;
        dstar[0] = (double)ds;
        dot = (double)(float)(ds * kf);
    } else {
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            for (int j = 0; j < D; ++j) s += h->kinv[(size_t)i * h->cap + j] * h->kf[j];
            dstar[i] = s;
        }
        for (int i = 0; i < D; ++i) dot += dstar[i] * h->kf[i];
    }
    double delta = Kii - dot;
    if (delta < 0) delta = 0;
    if (inside) {                                            /* projectron.py:73-84 (never extends the dictionary) */
        const double loss = 1 - margin;
        double norm_xt = Kii - delta;
        if (norm_xt < 0) norm_xt = 0;
        if (loss - delta / kb->eta > 0) {
            double alpha = loss / norm_xt;
            if (alpha > 1) alpha = 1;
            const double a2 = 2 * (loss - delta / kb->eta) / norm_xt;
            if (a2 < alpha) alpha = a2;
            const double ay = alpha * y;                     /* alpha * y * d_star: float64 products; float32 coeff += rounds once */
            if (D <= 1) h->coeff[0] = (double)(float)(h->coeff[0] + ay * dstar[0]);
            else for (int i = 0; i < D; ++i) h->coeff[i] += ay * dstar[i];
        }
        return;
    }
    if (delta <= kb->eta) {                                  /* sv.update(y * d_star) */
        if (D <= 1) h->coeff[0] = (double)(float)((float)h->coeff[0] + (float)y * (float)dstar[0]);
        else for (int i = 0; i < D; ++i) h->coeff[i] += y * dstar[i];
        return;
    }
    h->coeff[D] = (double)y;                                 /* sv.extend(y); sv.insert(x) */
    memcpy(h->lm + (size_t)D * h->d, x, sizeof(double) * h->d);
    h->D = D + 1;
    if (D == 0) { h->kinv[0] = (double)(float)(1.0 / Kii); return; }
    int cap = h->cap;
    for (int i = 0; i <= D; ++i) { h->kinv[(size_t)i * cap + D] = 0.0; h->kinv[(size_t)D * cap + i] = 0.0; }
    dstar[D] = -1.0;                                         /* d_star_extend */
    for (int i = 0; i <= D; ++i)
        for (int j = 0; j <= D; ++j) h->kinv[(size_t)i * cap + j] += dstar[i] * dstar[j] / delta;
}

orc_kb *orc_kb_create(int S, const int32_t *dims, const int32_t *offsets, int n_prbs, double alfa, double acc_lo,
                      double acc_hi, const int64_t *init_action, const int64_t *init_sec, double gamma, double eta) {
    orc_kb *kb = (orc_kb *)calloc(1, sizeof(orc_kb));
    kb->S = S; kb->n_prbs = n_prbs; kb->alfa = alfa; kb->acc_lo = acc_lo; kb->acc_hi = acc_hi;
    kb->gamma = gamma; kb->eta = eta;
    kb->L = (learner_t *)calloc(S, sizeof(learner_t));
    kb->action = (int64_t *)calloc(S, sizeof(int64_t));
    kb->sec = (int64_t *)calloc(S, sizeof(int64_t));
    kb->margins = (int64_t *)calloc(S, sizeof(int64_t));
    kb->acc = (double *)calloc((size_t)S * n_prbs, sizeof(double));
    kb->tie_ctr = (uint32_t *)calloc(S, sizeof(uint32_t));
    for (int s = 0; s < S; ++s) {
        kb->L[s].d = dims[s]; kb->L[s].off = offsets[s];
        grow(&kb->L[s]);
        kb->action[s] = init_action[s]; kb->sec[s] = init_sec[s];
        for (int a = 0; a < n_prbs; ++a) kb->acc[(size_t)s * n_prbs + a] = (acc_lo + acc_hi) / 2;
    }
    return kb;
}

void orc_kb_destroy(orc_kb *kb) {
    if (!kb) return;
    for (int s = 0; s < kb->S; ++s) { free(kb->L[s].lm); free(kb->L[s].coeff); free(kb->L[s].kinv); free(kb->L[s].kf); }
    free(kb->tie_ctr);
    free(kb->L); free(kb->action); free(kb->sec); free(kb->margins); free(kb->acc); free(kb);
}

void orc_kb_set_algorithm(orc_kb *kb, int plus) { kb->plus = plus; }

void orc_kb_set_tie_stream(orc_kb *kb, uint64_t seed, uint32_t env_id) {
    kb->tie_on = 1; kb->tie_key[0] = (uint32_t)seed; kb->tie_key[1] = (uint32_t)(seed >> 32); kb->tie_env = env_id;
}

static void make_x(const learner_t *h, const float *state, int64_t a, int n_prbs, double *x) {
    for (int i = 0; i < h->d - 1; ++i) x[i] = (double)state[h->off + i];   /* np.append(f32 slice, float) -> f64 */
    x[h->d - 1] = (double)a / (double)n_prbs;
}

/* KBRL_Control.update_control (kbrl_control.py:80-114) */
void orc_kb_update_control(orc_kb *kb, const float *state, const int64_t *action, const int64_t *labels, int64_t *hits) {
    double x[16];
    int n = kb->n_prbs;
    for (int i = 0; i < kb->S; ++i) {
        learner_t *h = &kb->L[i];
        kb->cur = i;
        int64_t a0 = action[i];
        make_x(h, state, a0, n, x);
        int y_pred = predict(kb, h, x);
        int y = (int)labels[i];
        int hit = y == y_pred;
        int64_t margin = kb->margins[i] > 0 ? kb->margins[i] : 0;
        double *acc = kb->acc + (size_t)i * n;
        if (y_pred == 1) {
            if (!hit) for (int64_t m = 0; m < margin + 1 && m < n; ++m) acc[m] = (1 - kb->alfa) * acc[m];
            else for (int64_t m = margin; m < n; ++m) acc[m] = (1 - kb->alfa) * acc[m] + kb->alfa;
        }
        if (!kb->adjusted) {                                  /* np.argmax(accuracies[i,:] > lo): first True, else 0 */
            int64_t sf = 0;
            for (int m = 0; m < n; ++m) if (acc[m] > kb->acc_lo) { sf = m; break; }
            kb->sec[i] = sf;
        }
        hits[i] = hit;
        int64_t lo = y == 1 ? a0 : 0, hi = y == 1 ? n : a0;   /* sample augmentation, :103-112 */
        for (int64_t a = lo; a <= hi; ++a) {
            make_x(h, state, a, n, x);
            predict(kb, h, x);
            update(kb, h, x, y);
        }
    }
}

/* KBRL_Control.select_action + adjust_action (kbrl_control.py:41-78) */
void orc_kb_select_action(orc_kb *kb, const float *state, int64_t *action_out, int32_t *adjusted_out) {
    double x[16];
    int n = kb->n_prbs;
    int64_t assigned = 0;
    for (int i = 0; i < kb->S; ++i) {
        learner_t *h = &kb->L[i];
        kb->cur = i;
        int64_t offset = kb->sec[i], margin = 0, l1 = n;
        for (int64_t c = 0; c <= n; ++c) {
            make_x(h, state, c, n, x);
            if (predict(kb, h, x) == 1) {
                int64_t a = c + offset < n ? c + offset : n;
                margin = a - c; l1 = a;
                break;
            }
        }
        kb->action[i] = l1; kb->margins[i] = margin;
        assigned += l1;
    }
    int adjusted = 0;
    if (assigned > n) {
        adjusted = 1;
        for (int i = 0; i < kb->S; ++i) {
            double p = (double)kb->action[i] / (double)assigned;
            int64_t na = (int64_t)floor(n * p);
            kb->margins[i] -= kb->action[i] - na;
            kb->action[i] = na;
        }
    }
    kb->adjusted = adjusted;                                  /* run(): action, self.adjusted = select_action(...) */
    for (int i = 0; i < kb->S; ++i) action_out[i] = kb->action[i];
    *adjusted_out = adjusted;
}

void orc_kb_get_control(const orc_kb *kb, int64_t *sec, int64_t *margins, double *acc, int64_t *sizes, long *tie_breaks) {
    for (int i = 0; i < kb->S; ++i) {
        if (sec) sec[i] = kb->sec[i];
        if (margins) margins[i] = kb->margins[i];
        if (sizes) sizes[i] = kb->L[i].D;
    }
    if (acc) memcpy(acc, kb->acc, sizeof(double) * (size_t)kb->S * kb->n_prbs);
    if (tie_breaks) *tie_breaks = kb->tie_breaks;
}

/* dictionary of learner s: landmarks [D][d], coeff [D], kinv [D][D] (row-major, packed) */
int orc_kb_get_learner(const orc_kb *kb, int s, double *lm, double *coeff, double *kinv) {
    const learner_t *h = &kb->L[s];
    for (int i = 0; i < h->D; ++i) {
        if (lm) memcpy(lm + (size_t)i * h->d, h->lm + (size_t)i * h->d, sizeof(double) * h->d);
        if (coeff) coeff[i] = h->coeff[i];
        if (kinv) for (int j = 0; j < h->D; ++j) kinv[(size_t)i * h->D + j] = h->kinv[(size_t)i * h->cap + j];
    }
    return h->D;
}

/* single-learner entry points for unit tests of predict / update */
double orc_kb_predict(orc_kb *kb, int s, const double *x, int32_t *y) { kb->cur = s; *y = predict(kb, &kb->L[s], x); return kb->L[s].f; }
void orc_kb_update(orc_kb *kb, int s, const double *x, int32_t y) { update(kb, &kb->L[s], x, y); }
