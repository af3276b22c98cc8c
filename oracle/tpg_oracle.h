/*
 * tpg_oracle.h — C API of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load
 * this library.  It is a CPU restatement of GEGELATI's policy-evaluation path; the product
 * (libtpgx.so) never links, includes or calls anything in oracle/.
 *
 * PARITY STATUS: the engine part is "parity unpinned" in the strict sense — the GEGELATI library
 * is not vendored in /root/reference and no golden logs exist there.  What IS
 * pinned: (1) instruction semantics and environment dynamics against the reference's own sources,
 * compiled unmodified into oracle/_ref/libtpgref.so (tests/test_oracle_vs_ref.py and the golden
 * fixtures generated from it); (2) the three golden graphs of the reference (Appendix C KATs).
 */
#ifndef TPG_ORACLE_H
#define TPG_ORACLE_H

#include "../include/tpgx.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MATH_STD = 0 /* glibc std::exp ... what the real apps compute */,
       ORC_MATH_SHARED = 1 /* tpgx_math.h: the sequence the GPU also runs */ };

typedef struct orc_engine orc_engine;

/* Build an engine from the same flat inputs libtpgx takes. Returns NULL and fills err on failure. */
orc_engine* orc_create(const tpgx_config* cfg, int32_t n_ops, const int32_t* opcode,
                       int32_t n_src, const tpgx_source_desc* src, const tpgx_graph* g,
                       int32_t math_flavour, char* err, size_t err_len);
void orc_destroy(orc_engine* e);

/* [UP] Program::identifyIntrons restated: flags[line] = 1 for introns; returns the intron count. */
int64_t orc_identify_introns(const orc_engine* e, uint8_t* flags);
/* Effective (non-intron) line count per program. */
void orc_effective_lines(const orc_engine* e, int32_t* n_eff);

/* [UP] ProgramExecutionEngine::executeProgram on explicit observations, NaN -> -inf applied
 * ([UP] TPGExecutionEngine::evaluateEdge).  obs_f64[s]/obs_i32[s] : per source, [count][space]
 * (NULL when the source has the other element type).  bid : [P][count].  raw != 0 skips the
 * NaN filter.                                                                                    */
int32_t orc_execute_programs(const orc_engine* e, int64_t count, const double* const* obs_f64,
                             const int32_t* const* obs_i32, double* bid, int32_t raw);

/* Classification evaluateAllRoots on uint8 images ([UP] ClassificationLearningAgent::evaluateJob).
 * images: [n_images][h*w] u8, labels [n_images].  nb_threads <= 1 runs the sequential
 * LearningAgent path, otherwise the ParallelLearningAgent job queue.                             */
int32_t orc_evaluate_classification(const orc_engine* e, const uint8_t* images,
                                    const uint8_t* labels, int64_t n_images,
                                    const tpgx_classif_request* req, tpgx_classif_result* out,
                                    int32_t nb_threads);

/* Archive side effect of a TRAINING evaluateAllRoots ([UP] Archive::addRecording via evaluateEdge, per-job archives merged
 * in job order): the final <= archive_size recordings as (program id, observation index, bid). */
int32_t orc_archive_classification(const orc_engine* e, const uint8_t* images, int64_t n_images, const tpgx_classif_request* req,
                                   const uint64_t* archive_seed, double probability, int32_t archive_size,
                                   int32_t* out_program, int64_t* out_sample, double* out_bid, int32_t* nb_records);

/* Episodic evaluateAllRoots ([UP] LearningAgent::evaluateJob / AdversarialLearningAgent). */
int32_t orc_evaluate_episodes(const orc_engine* e, const tpgx_episode_request* req,
                              tpgx_episode_result* out, int32_t nb_threads);

/* Archive side effect of the same call in TRAINING mode: final <= archive_size recordings as (program, job, iteration,
 * step, bid) plus the observation (data sources concatenated, ints as doubles, obs_len <= 16 values). */
int32_t orc_archive_episodes(const orc_engine* e, const tpgx_episode_request* req, const uint64_t* archive_seed, double probability,
                             int32_t archive_size, int32_t* out_program, int32_t* out_job, int32_t* out_iteration, int32_t* out_step,
                             double* out_bid, double* out_obs, int32_t obs_len, int32_t* nb_records);

/* ---- single pieces exposed for pinning tests ---- */
/* Apply device-opcode semantics on scalar operands: a,b doubles / ia,ib ints / c constant /
 * win 3x3 row-major window. */
double orc_apply_op(int32_t opcode, int32_t math_flavour, double a, double b, int32_t ia,
                    int32_t ib, double c, const double* win);
void orc_math_array(int32_t fn /* tpgx_opcode EXP..TAN */, int32_t math_flavour, int64_t n,
                    const double* in, double* out);

void orc_math_probe(int32_t fn /* TPGX_PROBE_* */, int64_t n, const double* a, const double* b, double* out);

/* Environment restatements, stepped one action at a time. */
typedef struct orc_env orc_env;
orc_env* orc_env_create(int32_t kind, int32_t second_player_random, int32_t nb_pendulum_actions,
                        const double* pendulum_actions, int32_t math_flavour);
void orc_env_destroy(orc_env* v);
void orc_env_reset(orc_env* v, uint64_t seed, int32_t mode);
void orc_env_set_state(orc_env* v, const double* state, int32_t n); /* test hook: overwrite observable state */
int32_t orc_env_do_action(orc_env* v, double action); /* returns -1 if the reference would throw */
int32_t orc_env_is_terminal(const orc_env* v);
double orc_env_score(const orc_env* v, int32_t player);
/* Fills up to n doubles with the env's observable state + bookkeeping (layout per env, see .cpp). */
int32_t orc_env_state(const orc_env* v, double* state, int32_t n);

/* RNG restatement ([UP] Mutator::RNG on libstdc++): fills out with getUnsignedInt64(lo,hi) draws. */
void orc_rng_u64(uint64_t seed, int64_t n, uint64_t lo, uint64_t hi, uint64_t* out);
void orc_rng_f64(uint64_t seed, int64_t n, double lo, double hi, double* out);
void orc_rng_raw(uint64_t seed, int64_t n, uint64_t* out);
/* same draws through the product's restated distributions (tpgx_rng.h); n <= 256 raw draws */
void orc_rng_restated_u64(uint64_t seed, int64_t n, uint64_t lo, uint64_t hi, uint64_t* out);
void orc_rng_restated_f64(uint64_t seed, int64_t n, double lo, double hi, double* out);

#ifdef __cplusplus
}
#endif
#endif
