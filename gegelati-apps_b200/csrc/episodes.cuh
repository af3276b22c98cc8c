/*
 * episodes.cuh — parameter block of the episodic-environment kernel (kernels for SURVEY.md §8 a8).
 */
#ifndef TPGX_EPISODES_CUH
#define TPGX_EPISODES_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "tpgx_internal.h"
#include "tpgx_rng.h"

struct tpgx_episode_params {
    int env, mode, nb_iterations, max_actions, nb_jobs, roots_per_job, second_player_random, trace_steps;
    int tie_break, nb_constants, nb_pendulum_actions;
    int group_width;              /* lanes per episode: power of two >= the largest team's edge count, <= 32 */
    int groups_per_warp;          /* episodes per warp: 1..4, groups_per_warp * group_width <= 32          */
    int cta_warps;                /* > 0: the plain pass runs tpgx_episode_cta_kernel with that many edge warps */
    double pendulum_actions[16];
    const uint64_t* rng_raw;      /* [nb_iterations][TPGX_RNG_TABLE] raw mt19937_64 outputs          */
    const int32_t* job_teams;     /* [nb_jobs][roots_per_job] root team ids                          */
    const uint32_t* code;
    const int32_t* prog_offset;
    const double* constants;
    const int32_t* team_edge_offset;
    const int32_t* edge_program;
    const int32_t* edge_destination;
    double* episode_score;        /* [nb_jobs][nb_iterations][roots_per_job]                         */
    int32_t* episode_actions;     /* [nb_jobs][nb_iterations]                                        */
    double* final_state;          /* [nb_jobs][nb_iterations][4]                                     */
    int8_t* trace;                /* [nb_jobs][nb_iterations][trace_steps]                           */
    double* job_score;            /* [nb_jobs][roots_per_job]                                        */
    unsigned long long* counters; /* [4]                                                             */
    int* status;
    /* archive reproduction ([UP] Archive::addRecording): pass 1 counts the program executions of
       every episode; pass 2 (hit_offset != nullptr) re-runs the episodes that hold recorded executions and resolves
       them — the h-th recorded execution of an episode is execution number hit_exec[h] (0-based, in the reference's
       order: steps, then teams along the path, then admissible edges in list order) and goes to record hit_slot[h] */
    int32_t* episode_progs;       /* [nb_jobs][nb_iterations] out, optional                          */
    const int32_t* hit_offset;    /* [nb_jobs * nb_iterations + 1]                                   */
    const int32_t* hit_exec;
    const int32_t* hit_slot;
    int32_t* rec_program;         /* [records]                                                       */
    int32_t* rec_step;            /* [records] number of doAction calls before the observation       */
    double* rec_bid;              /* [records]                                                       */
    double* rec_obs;              /* [records][obs_len] data sources at that step, concatenated      */
    int obs_len;
};

cudaError_t tpgx_launch_episodes(const tpgx_episode_params& p, cudaStream_t st);
/* host: raw outputs of std::mt19937_64 seeded with `seed` */
void tpgx_rng_fill_raw(uint64_t seed, int n, uint64_t* out);

#endif
