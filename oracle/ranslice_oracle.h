/*
 * ranslice_oracle.h -- CPU ORACLE for the env.step() path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithm (node_b.py, slice_l1.py, slice_ran.py,
 * schedulers.py, channel_models.py, traffic_generators.py, ran_slice.py).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product (network-slicing_b200/csrc) never includes, links or calls anything from here.
 *
 * Parity pin: tests/test_oracle_golden.py drives this oracle (a) through numpy-Generator
 * callbacks against golden trace A (unmodified reference, native seeding) and (b) with its
 * built-in Philox streams against golden trace B (unmodified reference with injected streams).
 */
#ifndef RANSLICE_ORACLE_H
#define RANSLICE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_STREAM_RAN = 0, ORC_STREAM_CHAN = 1, ORC_STREAM_L1RX = 2, ORC_STREAM_VBR = 3,
       ORC_STREAM_MTC = 4, ORC_N_STREAMS = 5 };

/* RNG indirection: Philox streams (default) or host callbacks (numpy Generator in tests). */
typedef struct orc_rng {
    void *ctx;
    double (*random)(void *ctx, int slice, int stream);
    double (*exponential)(void *ctx, int slice, int stream, double scale);
    int64_t (*integers)(void *ctx, int slice, int stream, int64_t n);
    int64_t (*choice)(void *ctx, int slice, int stream, int64_t n);   /* rng.choice(seq of length n) -> index */
    double (*normal)(void *ctx, int slice, int stream, double mu, double sigma);
    void (*random2)(void *ctx, int slice, int stream, double *xy);
} orc_rng;

typedef struct orc_config {
    int32_t n_prbs, n_embb, n_mmtc, slots_per_step;
    double penalty;
    double prop_A, prop_B;          /* channel_models.py:117-124 */
    int32_t l1_mux;                 /* 1: create_env(L1_level=False), the eMBB RAN slices share ONE L1 scheduler (scenario_creator.py:168-177) */
    int32_t reserved;
} orc_config;

typedef struct orc_tables {
    const double *trace;            /* [3][10001][100] time-major, column 10000 = NaN */
    const double *mcs_rate, *mcs_snr; /* [26] */
    const int32_t *mcs_order, *mcs_mod; /* [26]; mod 0 qpsk, 1 16qam, 2 64qam */
} orc_tables;

typedef struct orc_env orc_env;

/* built-in Philox streams: key = seed (the batch's base seed), counter word 3 = env_id (global env id) */
orc_env *orc_create(const orc_config *cfg, const orc_tables *tbl, uint64_t seed, uint32_t env_id);
void orc_set_rng(orc_env *e, const orc_rng *rng);     /* NULL -> built-in Philox(seed) */
void orc_destroy(orc_env *e);
int orc_n_variables(const orc_env *e);
void orc_reset(orc_env *e, float *obs);
/* returns flags (bit0 UE cap, bit1 burst cap, bit2 sum(action)>n_prbs clamp, bit3 same-slot departure) */
uint32_t orc_step(orc_env *e, const int64_t *action, float *obs, double *reward, int32_t *labels,
                  int32_t *violations, double *acc /* [S][10] or NULL */);
/* batched convenience: envs[i] are independent, stepped by n_threads pthreads */
void orc_step_batch(orc_env **envs, int n, int n_threads, const int64_t *actions, float *obs, double *reward,
                    int32_t *labels, int32_t *violations, uint32_t *flags);

/* leaf functions exposed for known-answer tests */
void orc_mcs_lut(const orc_tables *tbl, int e_snr, int *mcs, double *bps, int *rate);
double orc_response(const orc_tables *tbl, int mcs, const double *snr, int n);
double orc_macro_cell(double x, double y, double logf, double A, double B);
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* raw accumulators of every RAN slice after the last step, L1-major, [n_embb + n_mmtc][10] (info['l1_info'][l1][ran]) */
void orc_get_acc_ran(const orc_env *e, double *acc);
/* debugging: number of live UEs of eMBB slice s */
int orc_n_ues(const orc_env *e, int s);

#ifdef __cplusplus
}
#endif
#endif
