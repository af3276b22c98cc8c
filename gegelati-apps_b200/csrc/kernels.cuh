/*
 * kernels.cuh — launch parameter blocks shared between kernels.cu and tpgx_api.cu.
 */
#ifndef TPGX_KERNELS_CUH
#define TPGX_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "tpgx_internal.h"

/* kernels launched by this library in this process (tpgx_kernel_launches) */
std::atomic<unsigned long long>& tpgx_launch_counter();
#define TPGX_COUNT_LAUNCH() ((void)tpgx_launch_counter().fetch_add(1ull, std::memory_order_relaxed))

/* K1: bid interpreter over a 2-D uint8 observation source (MNIST shape). */
struct tpgx_bid_params {
    const uint8_t* tiles;        /* [tiles][hw][TS] pixel-major sample tiles (first tile of the pass) */
    int hw, width;               /* elements per observation, row length                              */
    const uint4* stream;         /* per bid-matrix row: TPGX_K1_CQ constant quads, then its wide lines */
    const int2* desc;            /* [rows] {first quad of the row in stream, number of lines}           */
    const double* pixel_lut;     /* [2][256] exp / ln of every uint8 pixel value, then [2][1021][1021] sobelMagn / sobelDir
                                    of every |gy|, |gx| (filled on the device by the tpgx_math routines)          */
    const double* sobel;         /* [tiles][2][nwin][TS] sobelMagn / sobelDir of every window position (first tile of the pass) */
    int nwin;                    /* window positions per observation, (h-2)(w-2)                        */
    int nb_units;                /* work units: programs (bid mode) or teams (team mode)                */
    int units_per_chunk;
    double* bid;                 /* bid mode: [rows][row_stride]                                        */
    long long row_stride;        /* samples per row in this pass (multiple of TS)                      */
    /* team mode ([UP] evaluateTeam fused into the interpreter) */
    const int2* team;            /* [nb_units] {first row, number of edges}; rows index desc / row_dst  */
    const int32_t* row_dst;      /* [rows] destination of the edge: >= 0 team unit index, < 0 action = -1 - value */
    int32_t* decision;           /* [nb_units][row_stride] destination of the winning edge              */
    int tie_break;
};

/* K1 v2 (k1_bid2.cu): threaded interpreter, persistent grid; stream / desc in the v2 encoding (k1v2_encode.h). */
struct tpgx_bid2_params {
    const uint8_t* tiles;        /* [tiles][hw + 1][128] pixel-major sample tiles (first tile of the pass) */
    int hw;
    const uint4* stream;         /* per row: K1V2_CQ head quads (constants, row header), then its entries   */
    const int2* unit_first;      /* [units] {first quad, number of entries} of the unit's first row         */
    const double* pixel_lut;
    const double* sobel;
    int nwin;
    int slots;                   /* shared-memory register slots per warp: the largest need of the shard's programs (<= K1V2_SLOTS) */
    int nb_units, units_per_item, items_per_tile, nb_items;
    unsigned int* work_counter;  /* next unclaimed item (zeroed by the launcher)                           */
    double* bid;
    long long row_stride;
    int32_t* decision;
    int tie_break;
};
size_t tpgx_bid2_smem_bytes(int slots);
cudaError_t tpgx_launch_bid2(const tpgx_bid2_params& p, bool team_mode, int nb_ctas, cudaStream_t st);
cudaError_t tpgx_launch_k1v2_encode(const uint32_t* code, const int32_t* prog_offset, const double* constants, int nb_constants,
                                    const int32_t* active_prog, const int2* desc, const int32_t* row_dst, const uint8_t* row_flags, int nb_rows,
                                    int width, int hw, int nwin, uint4* stream, int* flags, cudaStream_t st);

/* K2: graph traversal per (root, sample) on the bid matrix. */
struct tpgx_trav_params {
    const double* bid;
    long long row_stride;
    int nb_samples;              /* valid samples in this pass                                          */
    const int32_t* team_edge_offset;
    const int32_t* edge_row;     /* bid-matrix row of each edge's program                               */
    const int32_t* edge_destination;
    const int32_t* edge_lines;   /* effective line count of each edge's program                         */
    const int32_t* roots;        /* [nb_roots] root team ids of this shard                              */
    int nb_roots;
    const uint8_t* labels;       /* [nb_samples] class of each sample of the pass (may be NULL)         */
    int nb_classes;
    uint8_t* action;             /* optional [nb_roots][action_stride], pass offset already applied     */
    long long action_stride;
    unsigned long long* confusion; /* [nb_roots][C*C]                                                   */
    unsigned long long* counters;  /* [4] evaluations, team visits, program executions, lines           */
    int tie_break;
    int* status;                 /* OR of TPGX_DEV_* error bits                                         */
};

/* K2 in team mode: walk of the per-team decisions. */
struct tpgx_walk_params {
    const int32_t* decision;     /* [units][row_stride]                                                 */
    long long row_stride;
    int nb_samples;
    const int2* team_stat;       /* [units] {number of edges, sum of their effective lines} (reference-equivalent counters) */
    const int32_t* root_unit;    /* [nb_roots] unit index of each root team                             */
    int nb_roots;
    const uint8_t* labels;
    int nb_classes;
    uint8_t* action;
    long long action_stride;
    unsigned long long* confusion;
    unsigned long long* counters;
    int* status;
};
cudaError_t tpgx_launch_walk(const tpgx_walk_params& p, cudaStream_t st);

/* Archive reproduction on the bid matrix (bid mode): the reference-order stream of program executions of each job. */
struct tpgx_arch_params {
    const double* bid;
    long long row_stride;
    int nb_samples;
    const int32_t* team_edge_offset;
    const int32_t* edge_row;
    const int32_t* edge_destination;
    const int32_t* edge_program;      /* global program id of each edge                                   */
    const int32_t* roots;             /* [nb_roots] root team ids of the batch                             */
    int nb_roots;
    int tie_break;
    unsigned long long* counts;       /* [nb_roots][nb_samples] executions of (root, sample); scanned in place to an inclusive prefix */
    const unsigned long long* hit_k;  /* [nb_hits] execution index within its job                          */
    const int32_t* hit_root;          /* [nb_hits] root index within the batch                             */
    int nb_hits;
    const int32_t* sample_obs;        /* [nb_samples] observation index of each sample, or NULL = identity */
    int32_t* out_program;
    long long* out_sample;
    double* out_bid;
    int* status;
};
cudaError_t tpgx_launch_arch_count(const tpgx_arch_params& p, cudaStream_t st);
cudaError_t tpgx_arch_scan(const tpgx_arch_params& p, void* temp, size_t* temp_bytes, cudaStream_t st);
cudaError_t tpgx_launch_arch_resolve(const tpgx_arch_params& p, cudaStream_t st);

#define TPGX_SOBEL_LUT_N 1021
#define TPGX_LUT_DOUBLES (512 + 2 * TPGX_SOBEL_LUT_N * TPGX_SOBEL_LUT_N)

#define TPGX_DEV_NO_EDGE 1
#define TPGX_DEV_PATH_TOO_LONG 2
#define TPGX_DEV_BAD_ACTION 4
#define TPGX_DEV_RNG_EXHAUSTED 8
#define TPGX_DEV_BAD_LABEL 16

/* Launchers (defined in kernels.cu). variant: 0 = default tile shape. Returns cudaError_t. */
int tpgx_bid_tile_samples(int variant);
int tpgx_bid_warps(int variant);
size_t tpgx_bid_smem_bytes(int variant, int hw);
cudaError_t tpgx_launch_bid(const tpgx_bid_params& p, int nb_tiles, int variant, bool extended_ops, bool team_mode, cudaStream_t st);
cudaError_t tpgx_launch_traverse(const tpgx_trav_params& p, cudaStream_t st);
cudaError_t tpgx_launch_score(const unsigned long long* confusion, int nb_roots, int nb_classes,
                              double* records, cudaStream_t st);
cudaError_t tpgx_launch_pack(const uint8_t* obs, const int32_t* sample_index, long long nb_samples,
                             long long nb_obs, int hw, int ts, uint8_t* tiles, const uint8_t* labels_in,
                             uint8_t* labels_out, cudaStream_t st);
cudaError_t tpgx_launch_exec_generic(const uint32_t* code, const int32_t* prog_offset, const double* constants,
                                     int nb_constants, int nb_programs, long long count,
                                     const double* f0, const double* f1, const int32_t* i0, const int32_t* i1,
                                     int space0, int space1, int width0, int width1, double* bid, cudaStream_t st);
cudaError_t tpgx_launch_k1_encode(const uint32_t* code, const int32_t* prog_offset, const double* constants, int nb_constants,
                                  const int32_t* active_prog, const int2* desc, int nb_rows, int ts, int width, int hw, uint4* stream,
                                  int* flags, cudaStream_t st);
cudaError_t tpgx_launch_pixel_lut(double* lut, cudaStream_t st);
cudaError_t tpgx_launch_sobel_planes(const uint8_t* tiles, long long nb_tiles, int hw, int height, int width, int ts, const double* lut,
                                     double* planes, cudaStream_t st);
cudaError_t tpgx_launch_math_probe(int fn, long long n, const double* a, const double* b, double* out, cudaStream_t st);
cudaError_t tpgx_launch_fp64_peak(int mode /*0 dadd+dmul, 1 dfma*/, int blocks, int iters, double* sink, cudaStream_t st);

#endif
