/*
 * tpgx.h — C ABI of libtpgx.so, the B200-native evaluator for Tangled Program Graph policies.
 *
 * This is the drop-in boundary for GEGELATI's policy-evaluation hot path
 * (Learn::LearningAgent::evaluateAllRoots / ParallelLearningAgent).  Every entry
 * point names the reference interface it replaces.  "[UP]" = upstream GEGELATI library (not vendored
 * in /root/reference; behaviour restated from its published v1.x/v2.0.0 sources), "[REF]" = a file in
 * gegelati-apps.
 *
 * Conventions
 *   - plain C types only; every function returns a tpgx_status (0 = OK, negative = error);
 *     no C++ exception ever crosses this boundary (the reference throws std::runtime_error,
 *     e.g. [REF] pendulum/src/Learn/pendulum.cpp:66-68; the C++ adapter re-throws).
 *   - the caller owns every host buffer it passes in; inputs are copied before the call returns,
 *     outputs are written into caller-allocated buffers whose sizes follow from the request.
 *   - the library owns all device memory, streams and events; freed by tpgx_destroy.
 *   - a tpgx_ctx is used by one caller thread at a time (like one Learn::LearningAgent).
 *   - there is NO CPU fallback: without a CUDA device tpgx_create fails with TPGX_ERR_CUDA.
 */
#ifndef TPGX_H
#define TPGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPGX_ABI_VERSION 2

typedef int32_t tpgx_status;
enum {
    TPGX_OK = 0,
    TPGX_ERR_BAD_ARG = -1,
    TPGX_ERR_UNSUPPORTED = -2,   /* instruction / layout the device path does not implement (no fallback) */
    TPGX_ERR_GRAPH_INVALID = -3, /* e.g. a team with no admissible out-edge ([UP] evaluateTeam throws)   */
    TPGX_ERR_CUDA = -4,
    TPGX_ERR_NCCL = -5,
    TPGX_ERR_OOM = -6,
    TPGX_ERR_STATE = -7          /* call order (graph / observations not set yet)                        */
};

/* ---- device opcode table ------------------------------------------------------------------------
 * One entry per distinct behaviour of the 5 apps' Instructions::Set lambdas:
 * [REF] mnist/src/main.cpp:47-88, pendulum/src/Learn/instructions.cpp:7-32,
 * gridworld/src/instructions.cpp:7-27, stickgame/src/Learn/instructions.cpp:9-23,
 * tic-tac-toe/src/Learn/main.cpp:15-34.  Operand types are implied by the opcode:
 * D = double, I = int, C = Data::Constant, W = const double[3][3] window.                          */
typedef enum {
    TPGX_OP_SUB = 0,        /* (D,D)  a - b                                   */
    TPGX_OP_ADD = 1,        /* (D,D)  a + b                                   */
    TPGX_OP_MUL = 2,        /* (D,D)  a * b                                   */
    TPGX_OP_DIV = 3,        /* (D,D)  a / b   (no zero guard)                 */
    TPGX_OP_MAX = 4,        /* (D,D)  std::max(a,b) == (a < b) ? b : a        */
    TPGX_OP_EXP = 5,        /* (D)    exp(a)                                  */
    TPGX_OP_LOG = 6,        /* (D)    log(a)                                  */
    TPGX_OP_COS = 7,        /* (D)    cos(a)                                  */
    TPGX_OP_SIN = 8,        /* (D)    sin(a)                                  */
    TPGX_OP_TAN = 9,        /* (D)    tan(a)                                  */
    TPGX_OP_PI = 10,        /* (D)    M_PI  (operand fetched, ignored)        */
    TPGX_OP_MULC = 11,      /* (D,C)  a * (double)c / 10.0                    */
    TPGX_OP_FMODG = 12,     /* (D,D)  b != 0.0 ? fmod(a,b) : DBL_MIN          */
    TPGX_OP_SUB_II = 13,    /* (I,I)  (double)a - (double)b                   */
    TPGX_OP_CAST_I = 14,    /* (I)    (double)a                               */
    TPGX_OP_EQ0 = 15,       /* (D)    a == 0.0  ? 10.0 : 0.0                  */
    TPGX_OP_EQM1 = 16,      /* (D)    a == -1.0 ? 10.0 : 0.0                  */
    TPGX_OP_EQ1 = 17,       /* (D)    a == 1.0  ? 10.0 : 0.0                  */
    TPGX_OP_GE15 = 18,      /* (D)    a >= 15.0 ? 10.0 : 0.0                  */
    TPGX_OP_COND = 19,      /* (D,D)  a < b ? -a : a                          */
    TPGX_OP_SOBEL_MAGN = 20,/* (W)    sqrt(gx*gx + gy*gy)                     */
    TPGX_OP_SOBEL_DIR = 21, /* (W)    atan(gy / gx)                           */
    TPGX_OP_COUNT = 22
} tpgx_opcode;

/* Tie-break of [UP] TPG::TPGExecutionEngine::evaluateTeam.           */
enum { TPGX_TIE_LAST_MAX = 0 /* `bid >= best` */, TPGX_TIE_FIRST_MAX = 1 /* `bid > best` */ };

/* Element type of a data source as the *programs* see it ([UP] Data::PrimitiveTypeArray<T>,
 * Data::Array2DWrapper<T>).                                                                        */
enum { TPGX_ELEM_F64 = 0, TPGX_ELEM_I32 = 1 };
/* Storage type of an uploaded observation batch. U8 is exact for MNIST (pixels are the integers
 * 0..255 stored as double: [REF] mnist/src/mnist.h:16).                                            */
enum { TPGX_STORE_U8 = 0, TPGX_STORE_F64 = 1, TPGX_STORE_I32 = 2 };

typedef struct {
    int32_t elem;   /* TPGX_ELEM_*                                                     */
    int32_t height; /* 1 for a 1-D PrimitiveTypeArray, h for Array2DWrapper(h, w)      */
    int32_t width;  /* number of elements (1-D) or w                                    */
    int32_t is_2d;  /* 0: PrimitiveTypeArray<T>(width); 1: Array2DWrapper<T>(h, w)      */
} tpgx_source_desc;

typedef struct {
    int32_t device;       /* CUDA device ordinal                                                    */
    int32_t nb_registers; /* params.json nbRegisters (<= 8 on the device path)                      */
    int32_t nb_constants; /* params.json nbProgramConstant (0 => no constant source in the index map) */
    int32_t tie_break;    /* TPGX_TIE_*                                                             */
    int32_t reserved[4];
} tpgx_config;

/* One program line in the reference's own encoding ([UP] Program::Line; dot format SURVEY.md C.1):
 * instruction index into the Instructions::Set, destination register, two (source index, location)
 * pairs.  Source 0 = registers, 1 = constants iff nb_constants > 0, then the environment's data
 * sources in getDataSources() order.  Locations are un-reduced; the library takes them modulo the
 * address space of (source, operand type), as [UP] ProgramExecutionEngine::fetchCurrentOperands. */
typedef struct {
    uint32_t instruction;
    uint32_t destination;
    uint32_t src[2];
    uint32_t loc[2];
} tpgx_line;

/* The TPG, flattened by the caller in the reference's iteration order (edge order within a team =
 * the library's std::list order — it decides ties).                                               */
typedef struct {
    int32_t nb_programs;
    const int64_t* program_line_offset; /* [nb_programs + 1] into lines                              */
    const tpgx_line* lines;
    const uint8_t* intron;              /* optional, per line: 1 = intron ([UP] Program::identifyIntrons);
                                           NULL => the library recomputes them                       */
    const double* constants;            /* [nb_programs][nb_constants] (may be NULL if nb_constants==0) */
    int32_t nb_teams;
    const int32_t* team_edge_offset;    /* [nb_teams + 1] CSR                                        */
    const int32_t* edge_program;        /* [nb_edges] program id                                     */
    const int32_t* edge_destination;    /* [nb_edges] >= 0: team id;  < 0: action id = -1 - value    */
    int32_t nb_roots;
    const int32_t* root_team;           /* [nb_roots] team ids, in job-index order                   */
} tpgx_graph;

/* Work counters, in the reference's own units (what [UP] evaluateAllRoots would have executed).    */
typedef struct {
    uint64_t evaluations;        /* executeFromRoot + doAction pairs                                */
    uint64_t team_visits;        /* evaluateTeam calls                                              */
    uint64_t program_executions; /* evaluateEdge calls (admissible edges only)                      */
    uint64_t lines_executed;     /* non-intron lines run by those executions                        */
    uint64_t device_lines;       /* lines the GPU actually interpreted (dense program x sample)     */
    uint64_t device_program_executions;
} tpgx_counters;

typedef struct tpgx_ctx tpgx_ctx;

/* Replaces: construction of Learn::LearningAgent / Environment
 * ([REF] mnist/src/main.cpp:106-107, pendulum/src/Learn/main.cpp:38-39).                           */
tpgx_status tpgx_create(const tpgx_config* cfg, tpgx_ctx** out);
tpgx_status tpgx_destroy(tpgx_ctx* ctx);
/* Valid until the next call on ctx. ctx == NULL returns the last tpgx_create failure text.         */
const char* tpgx_last_error(const tpgx_ctx* ctx);
int32_t tpgx_abi_version(void);

/* Replaces: Instructions::Set + LambdaInstruction<...> registration
 * ([REF] mnist/src/main.cpp:79-88).  opcode[i] = device behaviour of instruction index i.           */
tpgx_status tpgx_set_instruction_table(tpgx_ctx* ctx, int32_t n, const int32_t* opcode);

/* Replaces: LearningEnvironment::getDataSources() ([REF] mnist/src/mnist.cpp:71-76,
 * stickgame/src/Learn/stickGameAdversarial.cpp:118-125).                                           */
tpgx_status tpgx_set_data_sources(tpgx_ctx* ctx, int32_t n, const tpgx_source_desc* src);

/* Replaces: the TPG::TPGGraph held by the agent (la.getTPGGraph(), [REF] mnist/src/main.cpp:110)
 * as seen by [UP] TPGExecutionEngine / ProgramExecutionEngine.  Flattens programs (introns
 * stripped, locations reduced) to packed device bytecode and uploads it with the CSR graph.        */
tpgx_status tpgx_set_graph(tpgx_ctx* ctx, const tpgx_graph* g);
/* Same for a context that is one rank of a communicator (tpgx_comm_init*): every rank is handed the WHOLE graph, as every
 * rank of a replicated host loop has it, but flattens, encodes and uploads only the programs its own root shard
 * (tpgx_comm_shard of g->nb_roots) can reach — 1/N of the host work per rank.  Evaluating a root outside the shard is then
 * TPGX_ERR_STATE.  Without a communicator it is tpgx_set_graph.                                                   */
tpgx_status tpgx_set_graph_sharded(tpgx_ctx* ctx, const tpgx_graph* g);

/* Replaces: the static dataset of [REF] mnist/src/mnist.cpp:6 / the Array2DWrapper pointer swap
 * of MNIST::changeCurrentImage (:8-34).  `data` is sample-major [count][height*width] of `store`
 * type, `labels` [count] class ids (may be NULL for bid-only use).  Copies host -> device and
 * re-tiles on the device.                                                                          */
tpgx_status tpgx_set_observations(tpgx_ctx* ctx, int32_t source, int32_t store, const void* data,
                                  const uint8_t* labels, int64_t count);
/* Same for PAGE-LOCKED host buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory): the copies are only
 * enqueued and the call returns at once, so the host can flatten the next graph while the observations cross
 * PCIe.  `data` / `labels` must stay valid and unmodified until the next tpgx_evaluate_* on ctx returns.       */
tpgx_status tpgx_set_observations_pinned(tpgx_ctx* ctx, int32_t source, int32_t store, const void* data,
                                         const uint8_t* labels, int64_t count);
/* Same, but `data` / `labels` are DEVICE pointers already resident in HBM (no host copy).          */
tpgx_status tpgx_set_observations_device(tpgx_ctx* ctx, int32_t source, int32_t store,
                                         const void* d_data, const uint8_t* d_labels, int64_t count);

/* ---- multi-GPU: roots shard across the GPUs of one box, per-root score records are combined with ONE
 * NCCL all-gather that the library enqueues on the evaluation stream (K3 writes the records straight into the collective's
 * buffer; no host synchronisation between the kernels and the collective).  The library owns the communicator.
 * Two ways to set it up, as NCCL itself offers:
 *   - one process per GPU: rank 0 calls tpgx_comm_unique_id, the caller distributes the TPGX_COMM_ID_BYTES by any means
 *     (MPI, a file, torch.distributed's store), every rank calls tpgx_comm_init on its context;
 *   - one process, N contexts on N devices (the reference's single-process model): tpgx_comm_init_all.
 * Sharding convention: with slot = ceil(total_roots / world), rank r evaluates roots [r * slot, min((r + 1) * slot, total)).
 * NCCL (libnccl.so.2) is loaded at the first of these calls; TPGX_ERR_NCCL when it is missing or reports an error.   */
#define TPGX_COMM_ID_BYTES 128
tpgx_status tpgx_comm_unique_id(uint8_t* id);
tpgx_status tpgx_comm_init(tpgx_ctx* ctx, const uint8_t* id, int32_t rank, int32_t world);
tpgx_status tpgx_comm_init_all(tpgx_ctx* const* ctxs, int32_t n);
/* this context's rank / world size (0 / 1 without a communicator) and its root shard of `total_roots`             */
tpgx_status tpgx_comm_shard(const tpgx_ctx* ctx, int32_t total_roots, int32_t* rank, int32_t* world, int32_t* root_begin, int32_t* root_end);

/* Classification evaluation request: one generation of evaluateAllRoots for a
 * ClassificationLearningEnvironment ([UP] ClassificationLearningAgent::evaluateJob; [REF]
 * mnist/src/mnist.cpp:50-69).  Every root in [root_begin, root_end) classifies the same sample
 * list (the reference seeds every root's episode identically: SURVEY.md §3.3).                     */
typedef struct {
    int64_t nb_samples;          /* number of doAction calls per root (MNIST: maxNbActionsPerEval)   */
    const int32_t* sample_index; /* [nb_samples] indices into the uploaded observations; NULL = 0..n-1 */
    int32_t nb_classes;          /* = nb actions for a classification LE (10)                        */
    int32_t root_begin;          /* shard of the root list evaluated by this context                 */
    int32_t root_end;
    int32_t reserved;
} tpgx_classif_request;

typedef struct {                 /* every pointer optional (NULL = not wanted); sizes for nR = root_end-root_begin */
    double* score;               /* [nR]            mean of per-class F1 (ClassificationEvaluationResult) */
    double* class_f1;            /* [nR][C]                                                          */
    uint64_t* confusion;         /* [nR][C][C]      table[true class][action]                        */
    uint8_t* action;             /* [nR][nb_samples]                                                 */
    double* bid;                 /* [nb_programs][nb_samples]  NaN already mapped to -inf            */
    tpgx_counters* counters;
    float* device_ms;            /* [3] K1 bid interpreter, K2 traversal, K3 score reduce (CUDA events) */
} tpgx_classif_result;

/* Replaces: Learn::LearningAgent::evaluateAllRoots(gen, mode) for classification environments
 * ([REF] direct call tic-tac-toe/src/Learn/main.cpp:82-83; via trainOneGeneration
 * mnist/src/main.cpp:145).                                                                         */
tpgx_status tpgx_evaluate_classification(tpgx_ctx* ctx, const tpgx_classif_request* req,
                                         tpgx_classif_result* out);
/* Device-resident variant for callers that keep results on the GPU (multi-GPU allgather):
 * d_records receives nR records of (1 + nb_classes) doubles {score, f1[0..C)} on `stream`
 * (a cudaStream_t passed as void*); no host synchronisation is performed.                          */
tpgx_status tpgx_evaluate_classification_device(tpgx_ctx* ctx, const tpgx_classif_request* req,
                                                double* d_records, void* stream);

/* Archive side effect of a TRAINING evaluateAllRoots ([UP] Archive::addRecording, called by evaluateEdge for every executed
 * program; archiveSize / archivingProbability: [REF] mnist/params.json:4,7).  Restated semantics:
 * every job j evaluates its root with a private Archive seeded with archive_seed[j]; an execution is recorded iff
 * `probability == 1.0 || probability >= rng.getDouble(0, 1)` (one draw per executed program, in execution order); a job keeps
 * its last archive_size recordings; the per-job archives are concatenated in job order and the LAST archive_size kept.
 * Output: records in final archive order as (program id, observation index — standing for the data-source hash —, bid).
 * The device supplies the per-(root, sample) execution counts and resolves the recorded executions; the RNG stream is
 * generated on the host (mt19937_64, as [UP] Mutator::RNG).  Only the jobs whose recordings survive the merge are evaluated. */
typedef struct {
    const uint64_t* archive_seed;  /* [root_end - root_begin] Job::archiveSeed of each root's job                     */
    double probability;            /* archivingProbability                                                             */
    int32_t archive_size;          /* archiveSize                                                                      */
    int32_t reserved;
} tpgx_archive_request;
typedef struct {
    int32_t* program;              /* [archive_size] */
    int64_t* sample;               /* [archive_size] */
    double* bid;                   /* [archive_size] */
    int32_t nb_records;            /* out */
    int32_t reserved;
} tpgx_archive_result;
tpgx_status tpgx_archive_classification(tpgx_ctx* ctx, const tpgx_classif_request* req, const tpgx_archive_request* arch,
                                        tpgx_archive_result* out);

/* ---- episodic environments ----------------------------------------------------------------------
 * Batched lock-step restatements of the apps' LearningEnvironment::{reset,doAction,isTerminal,
 * getScore(s)}.                                                             */
typedef enum {
    TPGX_ENV_GRIDWORLD = 0, /* [REF] gridworld/src/gridworld.cpp:3-83                               */
    TPGX_ENV_STICKGAME = 1, /* [REF] stickgame/src/Learn/stickGameAdversarial.cpp:38-171            */
    TPGX_ENV_TICTACTOE = 2, /* [REF] tic-tac-toe/src/Learn/TicTacToe.cpp:33-204                     */
    TPGX_ENV_PENDULUM = 3   /* [REF] pendulum/src/Learn/pendulum.cpp:42-141                         */
} tpgx_env_kind;

typedef struct {
    int32_t env;                 /* tpgx_env_kind                                                    */
    int32_t mode;                /* Learn::LearningMode: 0 TRAINING, 1 VALIDATION, 2 TESTING         */
    int32_t nb_iterations;       /* nbIterationsPerPolicyEvaluation (or nbIterationsPerJob)          */
    int32_t max_actions;         /* maxNbActionsPerEval                                              */
    const uint64_t* seeds;       /* [nb_iterations] value the agent passes to le.reset(seed, mode):
                                    Hash(gen) ^ Hash(it) ([UP] LearningAgent::evaluateJob)           */
    int32_t nb_jobs;
    int32_t roots_per_job;       /* 1; 2 for AdversarialLearningAgent jobs (tic-tac-toe)             */
    const int32_t* job_roots;    /* [nb_jobs][roots_per_job] indices into graph root list            */
    int32_t second_player_random;/* stickgame / tic-tac-toe constructor flag                         */
    int32_t nb_pendulum_actions; /* pendulum: size of availableActions (7)                           */
    const double* pendulum_actions; /* [nb_pendulum_actions] ([REF] pendulum/src/Learn/main.cpp:33)  */
    int32_t trace_steps;         /* >0: record the first trace_steps actions of every episode        */
    int32_t reserved;
    const uint64_t* rng_seeds;   /* optional [nb_iterations]: the value the environment's reset() hands to rng.setSeed(),
                                    i.e. Data::Hash<size_t>()(seed) ^ Data::Hash<LearningMode>()(mode) computed by the CALLER
                                    with the GEGELATI it links ([REF] pendulum/src/Learn/pendulum.cpp:45-48) — keeps the library
                                    independent of Data::Hash.  NULL: seeds[it] ^ mode (std::hash on libstdc++ is the identity) */
} tpgx_episode_request;

typedef struct {                 /* every pointer optional                                            */
    double* score;               /* [nb_jobs][roots_per_job] mean over iterations (EvaluationResult)  */
    double* episode_score;       /* [nb_jobs][nb_iterations][roots_per_job]                           */
    int32_t* episode_actions;    /* [nb_jobs][nb_iterations] number of doAction calls                 */
    int8_t* trace;               /* [nb_jobs][nb_iterations][trace_steps] action ids, -1 = not reached */
    double* final_state;         /* [nb_jobs][nb_iterations][4] env-specific observable state         */
    tpgx_counters* counters;
    float* device_ms;            /* [1] episode kernel                                                */
} tpgx_episode_result;

/* Replaces: Learn::LearningAgent::evaluateAllRoots(gen, mode) for sequential environments
 * ([UP] LearningAgent::evaluateJob / AdversarialLearningAgent::evaluateJob episode loop).          */
tpgx_status tpgx_evaluate_episodes(tpgx_ctx* ctx, const tpgx_episode_request* req,
                                   tpgx_episode_result* out);

/* Archive side effect of a TRAINING evaluateAllRoots of a sequential environment ([UP] Archive::addRecording called by
 * evaluateEdge; archiveSize / archivingProbability: [REF] pendulum/params.json:4,7, gridworld/params.json:2-3,
 * stickgame/params.json:4,7, tic-tac-toe/params.json:4,7).  Same restated semantics as tpgx_archive_classification
 * (one Archive per job seeded with archive_seed[job], one draw per executed program in execution order — iterations, steps,
 * teams along the path, admissible edges in list order; both roots of an adversarial job share the job's archive), with
 * the recorded observation identified by (job, iteration, step) and optionally copied out.  eval_out (optional) receives
 * the evaluation results of the same call: the archive is a side effect of evaluateAllRoots, not a second evaluation.   */
typedef struct {
    int32_t* program;              /* [archive_size] */
    int32_t* job;                  /* [archive_size] */
    int32_t* iteration;            /* [archive_size] */
    int32_t* step;                 /* [archive_size] doAction calls made before the recorded observation               */
    double* bid;                   /* [archive_size] */
    double* observation;           /* optional [archive_size][obs_len]: the data sources at that step, concatenated in
                                      getDataSources() order (int sources converted to double), zero padded           */
    int32_t obs_len;               /* in */
    int32_t nb_records;            /* out */
} tpgx_episode_archive_result;
tpgx_status tpgx_archive_episodes(tpgx_ctx* ctx, const tpgx_episode_request* req, const tpgx_archive_request* arch,
                                  tpgx_episode_result* eval_out, tpgx_episode_archive_result* out);

/* ---- host generation loop ------------------------------------------------------
 * Learn::LearningAgent::{init, trainOneGeneration} as the apps drive it ([REF] gridworld/src/main.cpp:33-34,57-65;
 * pendulum/src/Learn/main.cpp:38-39,83): initRandomTPG, then per generation populateTPG (clone + mutate roots up to
 * nbRoots, mutate the new programs until their behaviour changes against the archive) -> evaluateAllRoots in TRAINING
 * mode (scores + archive side effect) -> decimateWorstRoots.  The mutation operators are [UP] Mutator::TPGMutator /
 * ProgramMutator restated from the published algorithm with the params.json keys below; the library's exact draw order
 * is not in /root/reference, so graphs evolve like the reference's but not draw for draw ("parity unpinned",
 * DESIGN.md §5).  What IS exact: with the same trainer seed the loop is deterministic, and it depends on the evaluator
 * only through scores and archive records, which the GPU backend reproduces bit for bit — so the elite / mutation
 * sequence with the GPU evaluator equals the one with the CPU oracle (tests/test_gpu_trainer.py).
 * Evaluation and program execution go through a backend of two callbacks; tpgx_train_backend_episodes fills it with the
 * GPU implementation for the sequential environments.  There is no CPU backend in the library.                       */
typedef struct {
    int32_t nb_actions;                /* LearningEnvironment::getNbActions()                                       */
    int32_t nb_roots;                  /* mutation.tpg.nbRoots                                                      */
    int32_t max_init_outgoing_edges;   /* mutation.tpg.maxInitOutgoingEdges                                         */
    int32_t max_outgoing_edges;        /* mutation.tpg.maxOutgoingEdges                                             */
    double p_edge_deletion, p_edge_addition, p_program_mutation, p_edge_destination_change, p_edge_destination_is_action;
    int32_t force_program_behavior_change; /* mutation.tpg.forceProgramBehaviorChangeOnMutation                     */
    int32_t max_program_size;          /* mutation.prog.maxProgramSize                                              */
    double p_delete, p_add, p_mutate, p_swap, p_constant_mutation;
    int32_t min_const_value, max_const_value;
    double ratio_deleted_roots;        /* ratioDeletedRoots                                                         */
    int32_t archive_size;              /* archiveSize                                                               */
    int32_t reserved;
    double archiving_probability;      /* archivingProbability                                                      */
    int32_t max_nb_evaluation_per_policy;      /* maxNbEvaluationPerPolicy ([REF] gridworld/params.json:20): a root whose stored
                                                  result already counts that many evaluations is not evaluated again in
                                                  TRAINING mode ([UP] LearningAgent::isRootEvalSkipped); <= 0: never skipped  */
    int32_t do_validation;                     /* doValidation ([REF] mnist/params.json:11): evaluateAllRoots(VALIDATION) on the
                                                  surviving roots after the decimation                                       */
    int32_t nb_iterations_per_policy_evaluation; /* nbIterationsPerPolicyEvaluation: evaluations one plain job adds to a root     */
    int32_t nb_classes;                        /* > 0: Learn::ClassificationLearningAgent (results combine per class)        */
} tpgx_train_params;

typedef struct {
    void* user;
    /* evaluateAllRoots(generation, mode) of graph g, whose root list holds exactly the roots to evaluate (the trainer leaves
       out the roots [UP] evaluateJob skips), job j = root j: scores[nb_roots]; for a classification agent (nb_classes > 0)
       also class_scores[nb_roots][nb_classes] (per-class F1) and class_counts[nb_roots][nb_classes] (samples of each class
       seen).  Archive side effect when archive_size > 0 and probability > 0 (TRAINING only): one Archive per job seeded
       archive_seed[j], <= archive_size records (program id in g, bid, observation row of obs_len doubles) in archive order.
       Returns a tpgx_status.                                                                                        */
    tpgx_status (*evaluate)(void* user, const tpgx_graph* g, uint64_t generation, int32_t mode, const uint64_t* archive_seed,
                            double probability, int32_t archive_size, double* scores, int32_t nb_classes, double* class_scores,
                            uint64_t* class_counts, int32_t* rec_program, double* rec_bid, double* rec_obs, int32_t obs_len,
                            int32_t* nb_records);
    /* bid[nb_programs][count] of every program of g on `count` observation rows (obs_len doubles each)             */
    tpgx_status (*execute)(void* user, const tpgx_graph* g, int64_t count, const double* obs, int32_t obs_len, double* bid);
} tpgx_train_backend;

typedef struct {
    int32_t nb_teams, nb_programs, nb_roots, nb_new_programs; /* after populate                                      */
    int32_t nb_removed_roots, archive_records, behaviour_retries, reserved;
    double best_score, mean_score, worst_score;              /* over all roots: stored results of the skipped ones included */
    int32_t nb_evaluated_roots, nb_skipped_roots;            /* TRAINING pass: jobs run / roots at maxNbEvaluationPerPolicy  */
    double best_validation_score, mean_validation_score;     /* doValidation: over the surviving roots, else NaN            */
} tpgx_train_stats;

typedef struct tpgx_trainer tpgx_trainer;
/* [UP] LearningAgent::init(seed): RNG seeded, initRandomTPG.  obs_len = values per archived observation row (sum of
 * the data sources' address spaces).  Instruction table and data sources as for a context.                         */
tpgx_status tpgx_trainer_create(const tpgx_config* cfg, int32_t nb_instructions, const int32_t* opcode, int32_t nb_sources,
                                const tpgx_source_desc* sources, const tpgx_train_params* params, uint64_t seed,
                                tpgx_trainer** out);
void tpgx_trainer_destroy(tpgx_trainer* t);
const char* tpgx_trainer_last_error(const tpgx_trainer* t);
/* The current graph in the C ABI's layout (roots in creation order); pointers stay valid until the next call that
 * changes the trainer.                                                                                              */
tpgx_status tpgx_trainer_graph(tpgx_trainer* t, tpgx_graph* view);
/* [UP] LearningAgent::trainOneGeneration(generation) through the backend.                                          */
tpgx_status tpgx_trainer_train_generation(tpgx_trainer* t, const tpgx_train_backend* backend, uint64_t generation,
                                          tpgx_train_stats* stats);
/* The three steps on their own (tests, custom loops): populate needs the backend only for the behaviour-change test
 * (may be NULL when force_program_behavior_change is 0 or the archive is empty).                                    */
tpgx_status tpgx_trainer_populate(tpgx_trainer* t, const tpgx_train_backend* backend, tpgx_train_stats* stats);
tpgx_status tpgx_trainer_decimate(tpgx_trainer* t, const double* root_scores, tpgx_train_stats* stats);
/* 64-bit fingerprint of the graph (teams, edges, programs, constants): equal fingerprints <=> same evolution so far */
uint64_t tpgx_trainer_fingerprint(const tpgx_trainer* t);
/* [UP] resultsPerRoot: the stored result of every current root, in root order (tpgx_trainer_graph's root list):
 * score[r] (NaN when the root has not been evaluated yet) and nb_evaluation[r].  Returns the number of roots.          */
int32_t tpgx_trainer_root_results(tpgx_trainer* t, double* score, uint64_t* nb_evaluation, int32_t capacity);
/* [UP] LearningAgent::getBestRoot() / Selector::getBestRoot(): index in the current root list of the best root seen by
 * the last updateEvaluationRecords and its stored score; TPGX_ERR_STATE before the first generation or when that root
 * has left the graph.  [UP] keepBestPolicy(): every other root team is removed until the best root is the only root.  */
tpgx_status tpgx_trainer_best_root(tpgx_trainer* t, int32_t* root_index, double* score);
tpgx_status tpgx_trainer_keep_best_policy(tpgx_trainer* t);
/* GPU backend for the sequential environments: evaluate = tpgx_set_graph + tpgx_archive_episodes with the request
 * template `req` (nb_jobs / job_roots / seeds / mode are filled per call: job j = root j, iteration seeds
 * 1000 * generation + it — SURVEY.md §8d, the real Data::Hash is upstream), execute = tpgx_set_graph +
 * tpgx_execute_programs.  The backend borrows ctx and req (and req->pendulum_actions); free with
 * tpgx_train_backend_release.                                                                                       */
tpgx_status tpgx_train_backend_episodes(tpgx_ctx* ctx, const tpgx_episode_request* req, tpgx_train_backend* out);
/* GPU backend for the classification environment ([REF] mnist/src/main.cpp:106-107,145: ClassificationLearningAgent +
 * trainOneGeneration): every generation classifies the actions_per_evaluation (maxNbActionsPerEval) images that
 * tpgx_mnist_episode_indices draws from seed 1000 * generation, evaluate = tpgx_set_graph + tpgx_evaluate_classification +
 * tpgx_archive_classification (the archived observation is the image, 784 doubles), execute = tpgx_execute_programs.
 * `images` is the caller's host copy of the observations already uploaded to ctx (borrowed until release).             */
tpgx_status tpgx_train_backend_classification(tpgx_ctx* ctx, const uint8_t* images, int64_t nb_images, int32_t nb_classes,
                                              int64_t actions_per_evaluation, tpgx_train_backend* out);
void tpgx_train_backend_release(tpgx_train_backend* backend);

/* evaluateAllRoots over all GPUs: this context evaluates its shard of the `total_roots` roots of the graph it was given
 * (req->root_begin / root_end are ignored: the shard is tpgx_comm_shard's) and the records of ALL roots come back, in root
 * order, [total_roots][1 + nb_classes] = {mean F1, F1 per class}: records_all (host, may be NULL) after a synchronisation,
 * or — _device — d_records_all left on the device, ordered on `stream` like tpgx_evaluate_classification_device.
 * Every rank must make the call (it is a collective).  _all: the single-process form, one request per context.       */
tpgx_status tpgx_evaluate_classification_allgather(tpgx_ctx* ctx, const tpgx_classif_request* req, int32_t total_roots, double* records_all);
tpgx_status tpgx_evaluate_classification_allgather_device(tpgx_ctx* ctx, const tpgx_classif_request* req, int32_t total_roots,
                                                          double* d_records_all, void* stream);
tpgx_status tpgx_evaluate_classification_allgather_all(tpgx_ctx* const* ctxs, int32_t n, const tpgx_classif_request* req, int32_t total_roots,
                                                       double* records_all);
/* One generation in one call = tpgx_set_graph_sharded(ctx, g) + tpgx_evaluate_classification_allgather(ctx, req, ...),
 * i.e. [UP] LearningAgent::evaluateAllRoots on the graph the host has just mutated ([REF] mnist/src/main.cpp:145), with the host
 * side of the graph (introns, flattening, stream planning, uploads) pipelined behind the device: the rank's root shard is
 * cut into up to 3 slices of 512 roots or more on average (TPGX_PIPELINE in the environment overrides the count), the first one
 * half as long as the others, and slice k is prepared while the kernels of slice k - 1 run.  Results are identical to the two-call form.  `g` is only read during the call; ctx holds no graph
 * afterwards when more than one slice ran (call tpgx_set_graph before using the other entry points).  Collective.        */
tpgx_status tpgx_evaluate_graph_classification_allgather(tpgx_ctx* ctx, const tpgx_graph* g, const tpgx_classif_request* req,
                                                         int32_t total_roots, double* records_all);
/* number of kernels this library has launched in the calling process so far (bench.py's gpu_launches)               */
uint64_t tpgx_kernel_launches(void);

/* Direct access to single engine steps (used by parity tests; replaces
 * [UP] ProgramExecutionEngine::executeProgram on an explicit observation batch):
 * obs_f64 / obs_i32 are [count][address space] per source (NULL if absent); bid is [P][count].     */
tpgx_status tpgx_execute_programs(tpgx_ctx* ctx, int64_t count, const double* const* obs_f64,
                                  const int32_t* const* obs_i32, double* bid);

/* FP64 issue-rate microbenchmark used as the roofline denominator (non-FMA DADD/DMUL and DFMA
 * warp-instructions per second), measured on ctx's device.                                         */
tpgx_status tpgx_measure_fp64_peak(tpgx_ctx* ctx, double* dadd_per_s, double* dfma_per_s);

/* Host flattener on its own (no GPU needed): the work [UP] Program::identifyIntrons and
 * [UP] ProgramExecutionEngine::fetchCurrentOperands do — validates the graph against the instruction table and
 * data sources, marks introns by backward liveness from register 0 and packs the effective lines.
 * intron : optional [total lines], 1 = intron.   effective_lines : optional [nb_programs].
 * code / code_capacity : optional buffer for the packed 32-bit device words of all effective lines in program
 * order (csrc/tpgx_internal.h layout); *nb_code_words receives the count (also when code is NULL).
 * err / err_len : optional message buffer.  Returns the status tpgx_set_graph would return.                  */
tpgx_status tpgx_flatten_programs(const tpgx_config* cfg, int32_t nb_instructions, const int32_t* opcode, int32_t nb_sources,
                                  const tpgx_source_desc* sources, const tpgx_graph* g, uint8_t* intron, int32_t* effective_lines,
                                  uint32_t* code, int64_t code_capacity, int64_t* nb_code_words, char* err, size_t err_len);

/* Host only, no GPU needed: the sub-graph that roots [root_begin, root_end) of g reach, renumbered the way
 * tpgx_evaluate_graph_classification_allgather sets up each pipeline slice — teams breadth-first from the roots in root order,
 * programs in order of first appearance, every team's edges in the caller's order (the order decides ties in
 * [UP] TPGExecutionEngine::evaluateTeam), roots 0 .. root_end - root_begin - 1.  Evaluating root i of the sub-graph is
 * evaluating root root_begin + i of g.
 * Call once with the arrays NULL: *nb_teams / *nb_edges / *nb_programs receive the sizes.  Call again with the counts set to
 * the capacities: team_ids[nb_teams] and program_ids[nb_programs] map new ids to g's, team_edge_offset[nb_teams + 1],
 * edge_program[nb_edges] (new program ids), edge_destination[nb_edges] (new team ids, or the action code unchanged),
 * root_team[root_end - root_begin] (new team ids).                                                                     */
tpgx_status tpgx_slice_graph(const tpgx_graph* g, int32_t root_begin, int32_t root_end, int32_t* nb_teams, int32_t* nb_edges, int32_t* nb_programs,
                             int32_t* team_ids, int32_t* team_edge_offset, int32_t* edge_program, int32_t* edge_destination, int32_t* program_ids,
                             int32_t* root_team);

/* K1 v2 stream encoder on its own (host only; csrc/k1v2_encode.h — the device runs the same function when a graph is
 * planned): the entry stream of ONE program given its packed generic words (tpgx_flatten_programs' `code`), for a
 * 2-D uint8 source of `hw` elements in rows of `width` with `nwin` 3x3 window positions, 128 samples per tile.
 * quads: optional [quad_capacity][4] uint32 {handler, operand A, operand B, destination slot}; *nb_entries and
 * *nb_slots receive the entry count (incl. ST / NEXT / END entries) and the shared-memory slots the program needs.
 * TPGX_ERR_UNSUPPORTED: instruction outside the MNIST set or program too long for K1 v2 (K1 v1 runs it).
 * tpgx_k1v2_handler_name: name of handler id h ("ADD_AR", ...), NULL past the table.                         */
tpgx_status tpgx_k1v2_encode_program(const uint32_t* words, int32_t nb_words, int32_t width, int32_t hw, int32_t nwin,
                                     uint32_t* quads, int32_t quad_capacity, int32_t* nb_entries, int32_t* nb_slots);
const char* tpgx_k1v2_handler_name(int32_t h);

/* The image indices one MNIST evaluation classifies, restating [REF] mnist/src/mnist.cpp:58-69 (reset:
 * rng.setSeed(seed) with the RAW seed, currentIndex = -1, changeCurrentImage) and :8-34 / :50-56 (every
 * doAction classifies the current image, then draws the next): in TRAINING (0) and VALIDATION (1) mode
 * index i is the i-th draw of rng.getUnsignedInt64(0, dataset_size - 1) ([UP] Mutator::RNG = mt19937_64 +
 * uniform_int_distribution); in TESTING (2) mode the images are taken in order, (i) % dataset_size.
 * dataset_size is the size of the set the mode reads (training set in TRAINING, test set otherwise).
 * The result is the sample_index list of tpgx_classif_request.  Pure host function, no GPU needed.          */
tpgx_status tpgx_mnist_episode_indices(uint64_t seed, int32_t mode, int64_t dataset_size, int64_t nb_actions, int32_t* out);

/* Diagnostic entry point for parity tests: evaluates ONE primitive of the device arithmetic
 * (the division / square-root sequences and the shared libm of csrc/tpgx_math.h that stand in for
 * the reference's `/`, sqrt, std::exp, std::log, std::atan, ... — [REF] mnist/src/main.cpp:50-77)
 * elementwise on the GPU, so tests can compare it bit for bit with the host build of the same header
 * on millions of operands.  The *_K codes run the 4-wide variants the K1 interpreter uses.          */
enum {
    TPGX_PROBE_DIV = 0, TPGX_PROBE_DIV_K = 1, TPGX_PROBE_EXP = 2, TPGX_PROBE_EXP_K = 3, TPGX_PROBE_LOG = 4,
    TPGX_PROBE_LOG_K = 5, TPGX_PROBE_ATAN = 6, TPGX_PROBE_SQRT_NONNEG = 7 /* integer-valued a in [0, 2^31) */,
    TPGX_PROBE_DIV_SMALL_INT = 8 /* (int)a / (int)b, Sobel gradient range */, TPGX_PROBE_SIN = 9, TPGX_PROBE_COS = 10,
    TPGX_PROBE_TAN = 11, TPGX_PROBE_DIV10 = 12, TPGX_PROBE_DIV10_K = 13 /* a / 10.0 (multByConst) */
};
tpgx_status tpgx_math_probe(tpgx_ctx* ctx, int32_t fn, int64_t n, const double* a, const double* b, double* out);

#ifdef __cplusplus
}
#endif
#endif /* TPGX_H */
