/*
 * tpgx_api.cu — the C ABI of libtpgx.so (include/tpgx.h): context, uploads, evaluation drivers.
 * Host-side orchestration only; all arithmetic on observations happens in the kernels.
 * There is no CPU fallback: every evaluation entry point launches CUDA kernels or fails.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "episodes.cuh"
#include "k1v2_encode.h"
#include "kernels.cuh"
#include "tpgx_internal.h"

#include <chrono>
namespace {

std::string g_create_error;
struct Trace {
    const char* what; std::chrono::steady_clock::time_point t0; bool on;
    explicit Trace(const char* w) : what(w), t0(std::chrono::steady_clock::now()), on(getenv("TPGX_TRACE") != nullptr) {}
    void lap(const char* step) { if (!on) return; auto t = std::chrono::steady_clock::now(); fprintf(stderr, "[tpgx] %s: %s %.3f ms\n", what, step, std::chrono::duration<double, std::milli>(t - t0).count()); t0 = t; }
};

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool borrowed = false; /* p belongs to another context's buffer (pipeline lanes share the observation tiles) */
    void alias(const DevBuf& o) { release(); p = o.p; cap = o.cap; borrowed = o.p != nullptr; }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p && !borrowed) cudaFree(p);
        borrowed = false;
        p = nullptr;
        cap = 0;
        /* head room: small buffers (graph arrays, scratch of a generation loop: they grow a little every generation while the graph
           does) double, so that a loop stops allocating after a few generations; large ones (matrices, planes) take an eighth */
        size_t want = bytes < ((size_t)64 << 20) ? 2 * bytes + 256 : bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p && !borrowed) cudaFree(p); p = nullptr; cap = 0; borrowed = false; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct ShardPlan { /* programs reachable from a root range, cached between evaluations */
    int root_begin = -1, root_end = -1;
    uint64_t graph_version = 0;
    int nb_active = 0;
    int k1_ts = 0;        /* tile shape the K1 stream was encoded for */
    bool k1_ext = false;  /* the shard uses instructions outside the MNIST set */
    bool team_mode = false; /* [UP] evaluateTeam fused into K1 (acyclic shard, no bid output, little program sharing) */
    bool want_bid = false;
    bool k1_v2 = false;   /* the threaded interpreter (k1_bid2.cu) runs this shard */
    int k1_slots = 1;     /* ... with this many shared-memory register slots per warp (the largest need of its programs) */
    int nb_units = 0;     /* teams in team mode, programs otherwise */
    int64_t active_lines = 0, active_flops = 0;
};

} // namespace

/* set_graph_slice: stamped renumbering tables (a table entry is valid when its stamp is the current one: no clearing
   between slices) and the slice's sub-graph arrays, reused from call to call */
struct SliceMaps {
    uint32_t stamp = 0;
    std::vector<uint32_t> team_stamp, prog_stamp;
    std::vector<int32_t> team_id, prog_id, teams, progs, team_off, edge_prog, edge_dst, roots;
    std::vector<double> consts;
};

struct tpgx_ctx {
    tpgx_config cfg{};
    std::string err;
    std::vector<int32_t> opcode;
    std::vector<tpgx_source_desc> src;
    tpgx_flat_graph flat;
    bool has_graph = false, uses_trig = false, uses_ext = false, uses_int = false;
    uint64_t graph_version = 0;
    cudaStream_t stream = nullptr;
    std::vector<cudaEvent_t> events;
    int variant = 0;
    int sm_count = 148;
    size_t bid_budget = (size_t)16 << 30;

    DevBuf d_pixel_lut, d_sobel, d_tm_team, d_tm_dst, d_tm_stat, d_tm_roots, d_decision;
    DevBuf d_code, d_k1_stream, d_k1_desc, d_prog_off, d_consts, d_team_off, d_edge_dst, d_edge_row, d_edge_lines, d_roots, d_active, d_edge_prog;
    DevBuf d_obs_raw, d_labels_raw, d_tiles, d_labels, d_index, d_bid, d_conf, d_records, d_action, d_counters, d_status;
    DevBuf d_scratch[16], d_work, d_k2_rflags, d_k2_ufirst, d_gather;
    ncclComm_t comm = nullptr;           /* owned; NCCL communicator of this context's rank */
    int comm_rank = 0, comm_world = 1;
    int flat_root_begin = 0, flat_root_end = -1; /* tpgx_set_graph_sharded: the only roots this context may evaluate (end < 0: all) */
    /* tpgx_set_observations_pinned enqueues its copy / re-tiling / Sobel-plane work on a stream of its own, so that
       the graph uploads that follow on the main stream are not queued behind 47 MB of observations; consumers of the
       tiles wait for obs_ready */
    cudaStream_t obs_stream = nullptr;
    cudaEvent_t obs_ready = nullptr;
    bool obs_pending = false;
    const uint8_t* obs_dev = nullptr;    /* raw observations [nb_obs][hw] (ours or the caller's) */
    const uint8_t* labels_dev = nullptr;
    int64_t nb_obs = 0;
    int obs_hw = 0, obs_width = 0, nwin = 0;
    bool tiles_identity_valid = false;   /* d_tiles holds the identity packing of all observations */
    int tiles_ts = 0;
    ShardPlan plan;
    std::vector<int32_t> h_edge_row, h_active;
    std::vector<uint8_t> h_stage_in, h_stage_out; /* episode calls: one transfer each way */
    std::vector<tpgx_ctx*> lanes;        /* tpgx_evaluate_graph_classification_allgather: contexts of the pipeline's slices (owned) */
    SliceMaps slice_maps;                /* ... and the renumbering tables of set_graph_slice */
    bool tiles_shared = false;           /* a lane: d_tiles / d_labels / d_sobel are the parent's, prepared for the request */
    void* pin_ring = nullptr;            /* page-locked staging ring of upload() */
    int* h_status = nullptr;             /* page-locked copy of d_status ({device status, -, K1 encoder flags, -}) */
    size_t pin_used = 0;
    bool pin_failed = false;

    cudaEvent_t event(size_t i) {
        while (events.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); events.push_back(e); }
        return events[i];
    }
};

namespace {

tpgx_status fail(tpgx_ctx* c, tpgx_status st, const std::string& m) {
    if (c) c->err = m; else g_create_error = m;
    return st;
}
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? TPGX_ERR_OOM : TPGX_ERR_CUDA,            \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                            \
    } while (0)

/* work on the main stream that reads the observation buffers is ordered after a pending pinned upload */
static cudaError_t wait_for_observations(tpgx_ctx* ctx) {
    if (!ctx->obs_pending) return cudaSuccess;
    ctx->obs_pending = false;
    return cudaStreamWaitEvent(ctx->stream, ctx->obs_ready, 0);
}

/* Small host -> device copies (graph arrays, plans, per-call inputs: a dozen per generation) go through a page-locked ring:
   a copy from pageable memory is staged by the driver and blocks the host, one from page-locked memory is only enqueued.
   The ring wraps after a stream synchronisation, so a slot is never reused while a copy may still read it.
   Contract: the source is consumed when upload() returns (copied into the ring, or staged by the driver when it is
   pageable memory going the direct way), so callers may pass locals and need no synchronisation of their own. */
tpgx_status upload_at(tpgx_ctx* ctx, DevBuf& b, size_t offset, const void* src, size_t bytes, bool through_ring = true);
tpgx_status upload(tpgx_ctx* ctx, DevBuf& b, const void* src, size_t bytes, bool through_ring = true) {
    CU(b.ensure(bytes ? bytes : 16));
    return upload_at(ctx, b, 0, src, bytes, through_ring);
}
/* ... into a buffer that is already large enough, at a byte offset */
tpgx_status upload_at(tpgx_ctx* ctx, DevBuf& b, size_t offset, const void* src, size_t bytes, bool through_ring) {
    if (!bytes) return TPGX_OK;
    if (!b.p || offset + bytes > b.cap) return fail(ctx, TPGX_ERR_STATE, "upload_at: destination buffer too small");
    uint8_t* const dst = static_cast<uint8_t*>(b.p) + offset;
    const size_t ring = (size_t)48 << 20, limit = (size_t)8 << 20;
    if (through_ring && bytes <= limit) {
        if (!ctx->pin_ring && !ctx->pin_failed) {
            if (cudaHostAlloc(&ctx->pin_ring, ring, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->pin_ring = nullptr; ctx->pin_failed = true; }
        }
        if (ctx->pin_ring) {
            const size_t need = (bytes + 255) & ~(size_t)255;
            if (ctx->pin_used + need > ring) { CU(cudaStreamSynchronize(ctx->stream)); ctx->pin_used = 0; }
            void* slot = static_cast<uint8_t*>(ctx->pin_ring) + ctx->pin_used;
            ctx->pin_used += need;
            memcpy(slot, src, bytes);
            CU(cudaMemcpyAsync(dst, slot, bytes, cudaMemcpyHostToDevice, ctx->stream));
            return TPGX_OK;
        }
    }
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return TPGX_OK;
}

/* Plan of one evaluation shard.  Teams reachable from roots [rb, re) by BFS.
 * Bid mode: one bid-matrix row per distinct program, in order of first appearance; K2 walks the graph on bids.
 * Team mode (the default when it applies): one row per EDGE, grouped by team, and K1 itself keeps the winning
 * edge per (team, sample).  It applies when no bid output is wanted, the reachable graph is acyclic (then the
 * visited-team exclusion of [UP] evaluateTeam never removes an edge) and re-running shared programs once per
 * referencing team costs at most 25 % more lines than running each distinct program once.              */
tpgx_status make_plan(tpgx_ctx* ctx, int rb, int re, int ts, bool want_bid) {
    ShardPlan& pl = ctx->plan;
    if (pl.root_begin == rb && pl.root_end == re && pl.graph_version == ctx->graph_version && pl.k1_ts == ts && pl.want_bid == want_bid) return TPGX_OK;
    const tpgx_flat_graph& f = ctx->flat;
    Trace tr("make_plan");
    std::vector<int32_t> row_of(f.nb_programs, -1);
    std::vector<int32_t> unit_of(f.nb_teams, -1);
    std::vector<int32_t> queue;
    std::vector<int32_t> uniq;
    ctx->h_edge_row.assign(f.nb_edges, -1);
    for (int r = rb; r < re; r++) {
        int t = f.root_team[r];
        if (unit_of[t] < 0) { unit_of[t] = (int32_t)queue.size(); queue.push_back(t); }
    }
    int64_t uniq_lines = 0, edge_lines_total = 0;
    for (size_t q = 0; q < queue.size(); q++) {
        const int t = queue[q];
        for (int e = f.team_edge_offset[t]; e < f.team_edge_offset[t + 1]; e++) {
            const int p = f.edge_program[e];
            const int n = f.prog_offset[p + 1] - f.prog_offset[p];
            if (row_of[p] < 0) { row_of[p] = (int32_t)uniq.size(); uniq.push_back(p); uniq_lines += n; }
            edge_lines_total += n;
            ctx->h_edge_row[e] = row_of[p];
            const int d = f.edge_destination[e];
            if (d >= 0 && unit_of[d] < 0) { unit_of[d] = (int32_t)queue.size(); queue.push_back(d); }
        }
    }
    /* acyclic?  Kahn's algorithm on the reachable team graph */
    bool acyclic = true;
    {
        std::vector<int32_t> indeg(queue.size(), 0), stack;
        for (int t : queue)
            for (int e = f.team_edge_offset[t]; e < f.team_edge_offset[t + 1]; e++)
                if (f.edge_destination[e] >= 0) indeg[unit_of[f.edge_destination[e]]]++;
        for (size_t u = 0; u < queue.size(); u++) if (indeg[u] == 0) stack.push_back((int32_t)u);
        size_t done = 0;
        while (!stack.empty()) {
            const int u = stack.back(); stack.pop_back(); done++;
            const int t = queue[u];
            for (int e = f.team_edge_offset[t]; e < f.team_edge_offset[t + 1]; e++)
                if (f.edge_destination[e] >= 0 && --indeg[unit_of[f.edge_destination[e]]] == 0) stack.push_back(unit_of[f.edge_destination[e]]);
        }
        acyclic = done == queue.size();
    }
    const char* force = getenv("TPGX_TEAM_MODE");
    pl.team_mode = !want_bid && acyclic && edge_lines_total * 4 <= uniq_lines * 5 && !(force && force[0] == '0');
    pl.want_bid = want_bid;
    tpgx_status st;
    std::vector<int32_t> team_rows; /* team mode: {first row, number of rows} per unit */
    if (pl.team_mode) {
        std::vector<int32_t> team(2 * queue.size()), stat(2 * queue.size()), rows, dsts, roots(re - rb);
        for (size_t u = 0; u < queue.size(); u++) {
            const int t = queue[u];
            const int eb = f.team_edge_offset[t], ee = f.team_edge_offset[t + 1];
            team[2 * u] = (int32_t)rows.size(); team[2 * u + 1] = ee - eb;
            int lines = 0;
            for (int e = eb; e < ee; e++) {
                const int p = f.edge_program[e];
                rows.push_back(p);
                dsts.push_back(f.edge_destination[e] >= 0 ? unit_of[f.edge_destination[e]] : f.edge_destination[e]);
                lines += f.prog_offset[p + 1] - f.prog_offset[p];
            }
            stat[2 * u] = ee - eb; stat[2 * u + 1] = lines;
        }
        for (int r = rb; r < re; r++) roots[r - rb] = unit_of[f.root_team[r]];
        ctx->h_active = rows;
        team_rows = team;
        pl.nb_units = (int)queue.size();
        if ((st = upload(ctx, ctx->d_tm_team, team.data(), team.size() * 4)) != TPGX_OK) return st;
        if ((st = upload(ctx, ctx->d_tm_stat, stat.data(), stat.size() * 4)) != TPGX_OK) return st;
        if ((st = upload(ctx, ctx->d_tm_dst, dsts.data(), dsts.size() * 4)) != TPGX_OK) return st;
        if ((st = upload(ctx, ctx->d_tm_roots, roots.data(), roots.size() * 4)) != TPGX_OK) return st;
    } else {
        ctx->h_active = uniq;
        pl.nb_units = (int)uniq.size();
    }
    pl.nb_active = (int)ctx->h_active.size();
    pl.active_lines = 0;
    pl.active_flops = 0;
    for (int p : ctx->h_active) { pl.active_lines += f.prog_offset[p + 1] - f.prog_offset[p]; pl.active_flops += f.prog_flops[p]; }
    if ((st = upload(ctx, ctx->d_active, ctx->h_active.data(), ctx->h_active.size() * 4)) != TPGX_OK) return st;
    if ((st = upload(ctx, ctx->d_edge_row, ctx->h_edge_row.data(), ctx->h_edge_row.size() * 4)) != TPGX_OK) return st;
    if ((st = upload(ctx, ctx->d_roots, f.root_team.data() + rb, (size_t)(re - rb) * 4)) != TPGX_OK) return st;
    tr.lap("reachability + uploads");
    /* K1 stream: per row its constants followed by its pre-decoded lines.  The host ships only the row
       descriptors; the quads are expanded on the device from the generic code.  K1 v2 (threaded interpreter,
       k1_bid2.cu) takes the shard when every program of it is in the MNIST instruction set and fits its slots;
       otherwise K1 v1 (k1_bid.cu). */
    {
        const int nrows = (int)ctx->h_active.size();
        const char* v2env = getenv("TPGX_K1_V2");
        bool v2 = !(v2env && v2env[0] == '0') && !ctx->uses_ext && !ctx->uses_int && ts == 128 && ctx->variant == 0 &&
                  f.v2_entries.size() == (size_t)f.nb_programs;
        int slots = 1;
        for (int row = 0; v2 && row < nrows; row++) {
            const int p = ctx->h_active[row];
            if (f.v2_entries[p] < 0 || f.v2_slots[p] > K1V2_SLOTS) v2 = false;
            else slots = std::max(slots, (int)f.v2_slots[p]);
        }
        pl.k1_v2 = v2;
        pl.k1_slots = slots;
        std::vector<int32_t> desc(2 * (size_t)nrows);
        int64_t nq = 0;
        for (int row = 0; row < nrows; row++) {
            const int p = ctx->h_active[row];
            desc[2 * row] = (int32_t)nq;
            desc[2 * row + 1] = v2 ? f.v2_entries[p] : f.prog_offset[p + 1] - f.prog_offset[p];
            nq += (v2 ? K1V2_CQ : TPGX_K1_CQ) + desc[2 * row + 1];
        }
        if ((st = upload(ctx, ctx->d_k1_desc, desc.data(), desc.size() * 4)) != TPGX_OK) return st;
        CU(ctx->d_k1_stream.ensure(((size_t)nq + 64) * sizeof(tpgx_k1_quad))); /* slack: a lane may address up to 32 quads past a row's start */
        CU(ctx->d_status.ensure(16));
        int* d_flags = ctx->d_status.as<int>() + 2;
        CU(cudaMemsetAsync(d_flags, 0, 4, ctx->stream));
        if (v2) {
            /* row headers: edge destination (team mode: already on the device), entries of the next row, first / last row of
               its unit; and per unit where its first row starts */
            const uint8_t* d_row_flags = nullptr;
            if (pl.team_mode) {
                std::vector<uint8_t> rflags((size_t)nrows, 0);
                std::vector<int32_t> ufirst(2 * (size_t)pl.nb_units);
                for (int u = 0; u < pl.nb_units; u++) {
                    const int r0 = team_rows[2 * u], nr = team_rows[2 * u + 1];
                    rflags[r0] |= K1V2_ROW_FIRST;
                    rflags[r0 + nr - 1] |= K1V2_ROW_LAST;
                    ufirst[2 * u] = desc[2 * r0]; ufirst[2 * u + 1] = desc[2 * r0 + 1];
                }
                if ((st = upload(ctx, ctx->d_k2_rflags, rflags.data(), rflags.size())) != TPGX_OK) return st;
                if ((st = upload(ctx, ctx->d_k2_ufirst, ufirst.data(), ufirst.size() * 4)) != TPGX_OK) return st;
                d_row_flags = ctx->d_k2_rflags.as<uint8_t>();
            }
            CU(tpgx_launch_k1v2_encode(ctx->d_code.as<uint32_t>(), ctx->d_prog_off.as<int32_t>(), ctx->d_consts.as<double>(), ctx->cfg.nb_constants,
                                       ctx->d_active.as<int32_t>(), ctx->d_k1_desc.as<int2>(), pl.team_mode ? ctx->d_tm_dst.as<int32_t>() : nullptr,
                                       d_row_flags, nrows, ctx->obs_width, ctx->obs_hw,
                                       std::max(0, (ctx->src[0].height - 2) * (ctx->src[0].width - 2)), ctx->d_k1_stream.as<uint4>(), d_flags, ctx->stream));
        }
        else
            CU(tpgx_launch_k1_encode(ctx->d_code.as<uint32_t>(), ctx->d_prog_off.as<int32_t>(), ctx->d_consts.as<double>(), ctx->cfg.nb_constants,
                                     ctx->d_active.as<int32_t>(), ctx->d_k1_desc.as<int2>(), nrows, ts, ctx->obs_width, ctx->obs_hw, ctx->d_k1_stream.as<uint4>(),
                                     d_flags, ctx->stream));
        /* the encoder's verdict (flag 4: its entry / slot counts disagree with the host's) is read back with the device status
           at the end of the evaluation: no synchronisation here */
        /* which build of K1 (MNIST set only / extended) is known on the host from the opcode table */
        if (ctx->uses_int) return fail(ctx, TPGX_ERR_UNSUPPORTED, "int-typed instruction in a program evaluated on a uint8 image source");
        pl.k1_ext = ctx->uses_ext;
    }
    tr.lap("K1 stream encode + upload");
    pl.root_begin = rb; pl.root_end = re; pl.graph_version = ctx->graph_version; pl.k1_ts = ts;
    return TPGX_OK;
}

/* K0b after every re-tiling: sobelMagn / sobelDir planes of the tiles (2-D sources of at least 3x3 only) */
tpgx_status sobel_planes(tpgx_ctx* ctx, int64_t nTiles, int ts) {
    const tpgx_source_desc& d = ctx->src[0];
    ctx->nwin = 0;
    if (!d.is_2d || d.height < 3 || d.width < 3) return TPGX_OK;
    ctx->nwin = (d.height - 2) * (d.width - 2);
    CU(ctx->d_sobel.ensure((size_t)nTiles * 2 * ctx->nwin * ts * 8));
    CU(tpgx_launch_sobel_planes(ctx->d_tiles.as<uint8_t>(), nTiles, ctx->obs_hw, d.height, d.width, ts, ctx->d_pixel_lut.as<double>(),
                                ctx->d_sobel.as<double>(), ctx->stream));
    return TPGX_OK;
}

/* Observation tiles of a request: the identity packing made at upload time, or a re-tiling of the indexed samples (and their
   Sobel planes).  A pipeline lane uses its parent's, prepared once for the request. */
tpgx_status sobel_planes(tpgx_ctx* ctx, int64_t nTiles, int ts);
tpgx_status prepare_tiles(tpgx_ctx* ctx, const tpgx_classif_request* req, int ts) {
    const int64_t nS = req->nb_samples;
    const int64_t nTiles = (nS + ts - 1) / ts;
    const int hw = ctx->obs_hw;
    tpgx_status st;
    const bool identity = (req->sample_index == nullptr) && nS == ctx->nb_obs;
    if (!ctx->tiles_shared && !(identity && ctx->tiles_identity_valid && ctx->tiles_ts == ts)) {
        const int32_t* d_index = nullptr;
        if (req->sample_index) {
            for (int64_t i = 0; i < nS; i++)
                if (req->sample_index[i] < 0 || req->sample_index[i] >= ctx->nb_obs) return fail(ctx, TPGX_ERR_BAD_ARG, "sample_index out of range");
            if ((st = upload(ctx, ctx->d_index, req->sample_index, (size_t)nS * 4)) != TPGX_OK) return st;
            d_index = ctx->d_index.as<int32_t>();
        }
        CU(ctx->d_tiles.ensure((size_t)nTiles * (hw + 1) * ts));
        CU(ctx->d_labels.ensure((size_t)nTiles * ts));
        CU(tpgx_launch_pack(ctx->obs_dev, d_index, nS, ctx->nb_obs, hw, ts, ctx->d_tiles.as<uint8_t>(), ctx->labels_dev,
                            ctx->d_labels.as<uint8_t>(), ctx->stream));
        if ((st = sobel_planes(ctx, nTiles, ts)) != TPGX_OK) return st;
        ctx->tiles_identity_valid = identity;
        ctx->tiles_ts = ts;
    }

    return TPGX_OK;
}

struct ClassifRun {
    int nR = 0, C = 0, ts = 0;
    int64_t nS = 0, nTiles = 0;
    int passes = 0;
};

/* Enqueues one classification evaluateAllRoots on ctx->stream; no host synchronisation.
 * records_out: device buffer of nR*(C+1) doubles. */
tpgx_status enqueue_classification(tpgx_ctx* ctx, const tpgx_classif_request* req, double* records_out,
                                   bool want_action, bool keep_bid_single_pass, ClassifRun* run, bool timing, bool k1_only = false) {
    if (!ctx->has_graph) return fail(ctx, TPGX_ERR_STATE, "tpgx_evaluate_classification: no graph set");
    if (!ctx->obs_dev) return fail(ctx, TPGX_ERR_STATE, "tpgx_evaluate_classification: no observations set");
    if (!req || req->nb_samples <= 0) return fail(ctx, TPGX_ERR_BAD_ARG, "nb_samples must be positive");
    if (req->nb_classes < 1 || req->nb_classes > 16) return fail(ctx, TPGX_ERR_UNSUPPORTED, "nb_classes must be in [1,16]");
    if (req->root_begin < 0 || req->root_end > ctx->flat.nb_roots || req->root_begin >= req->root_end)
        return fail(ctx, TPGX_ERR_BAD_ARG, "root range out of bounds");
    if (ctx->flat_root_end >= 0 && (req->root_begin < ctx->flat_root_begin || req->root_end > ctx->flat_root_end))
        return fail(ctx, TPGX_ERR_STATE, "the graph was set with tpgx_set_graph_sharded: only this context's root shard can be evaluated");
    if (!req->sample_index && req->nb_samples > ctx->nb_obs) return fail(ctx, TPGX_ERR_BAD_ARG, "nb_samples exceeds uploaded observations");
    if (!ctx->labels_dev) return fail(ctx, TPGX_ERR_STATE, "classification needs labels");
    Trace etr("enqueue_classification");
    CU(wait_for_observations(ctx));
    etr.lap("wait for observations");
    const int nR = req->root_end - req->root_begin, C = req->nb_classes;
    const int64_t nS = req->nb_samples;
    const int ts = tpgx_bid_tile_samples(ctx->variant);
    const int64_t nTiles = (nS + ts - 1) / ts;
    const int hw = ctx->obs_hw;
    tpgx_status st = make_plan(ctx, req->root_begin, req->root_end, ts, keep_bid_single_pass);
    if (st != TPGX_OK) return st;
    etr.lap("plan");
    const int nA = ctx->plan.nb_active;

    if ((st = prepare_tiles(ctx, req, ts)) != TPGX_OK) return st;

    etr.lap("tiles + Sobel planes");
    /* pass planning: keep the bid matrix (or, in team mode, the decision matrix) of one pass within the budget */
    const bool team_mode = ctx->plan.team_mode;
    const int nU = ctx->plan.nb_units;
    int64_t tiles_per_pass = nTiles;
    const size_t per_tile = team_mode ? (size_t)std::max(nU, 1) * ts * 4 : (size_t)std::max(nA, 1) * ts * 8;
    if (per_tile * (size_t)nTiles > ctx->bid_budget) tiles_per_pass = std::max<int64_t>(1, (int64_t)(ctx->bid_budget / per_tile));
    if (keep_bid_single_pass && tiles_per_pass < nTiles) return fail(ctx, TPGX_ERR_OOM, "bid output requested but the bid matrix exceeds the budget");
    const int passes = (int)((nTiles + tiles_per_pass - 1) / tiles_per_pass);
    if (team_mode) CU(ctx->d_decision.ensure(per_tile * (size_t)tiles_per_pass));
    else CU(ctx->d_bid.ensure(per_tile * (size_t)tiles_per_pass));
    CU(ctx->d_conf.ensure((size_t)nR * C * C * 8));
    CU(ctx->d_counters.ensure(8 * 8));
    CU(ctx->d_status.ensure(16));
    CU(cudaMemsetAsync(ctx->d_conf.p, 0, (size_t)nR * C * C * 8, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_counters.p, 0, 64, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_status.p, 0, 8, ctx->stream)); /* words 2.. are the K1 encoder's flags, set when the plan was made */
    if (want_action) CU(ctx->d_action.ensure((size_t)nR * nS));
    etr.lap("buffers");

    size_t ev = 0;
    for (int ps = 0; ps < passes; ps++) {
        const int64_t t0 = (int64_t)ps * tiles_per_pass;
        const int nt = (int)std::min<int64_t>(tiles_per_pass, nTiles - t0);
        const int64_t s0 = t0 * ts;
        const int ns = (int)std::min<int64_t>((int64_t)nt * ts, nS - s0);
        if (timing) CU(cudaEventRecord(ctx->event(ev++), ctx->stream));
        if (nA > 0 && ctx->plan.k1_v2) {
            tpgx_bid2_params bp{};
            bp.tiles = ctx->d_tiles.as<uint8_t>() + (size_t)t0 * (hw + 1) * ts;
            bp.hw = hw;
            bp.stream = ctx->d_k1_stream.as<uint4>();
            bp.unit_first = team_mode ? ctx->d_k2_ufirst.as<int2>() : ctx->d_k1_desc.as<int2>();
            bp.pixel_lut = ctx->d_pixel_lut.as<double>();
            bp.sobel = ctx->nwin ? ctx->d_sobel.as<double>() + (size_t)t0 * 2 * ctx->nwin * ts : nullptr;
            bp.nwin = ctx->nwin;
            bp.nb_units = nU;
            bp.slots = ctx->plan.k1_slots;
            if (const char* v = getenv("TPGX_K1_SLOTS")) bp.slots = std::min(K1V2_SLOTS, std::max(bp.slots, atoi(v))); /* experiments: more than needed */
            /* work items = (tile, group of units), claimed warp by warp: ~20 items per resident warp when there is that
               much work (dynamic balance to the last item), at most 2 teams / 8 programs per item — an item of 8 teams is 40 rows,
               0.33 ms of a warp's time, and the last ones leave warps idle (C2 K1 7.97 -> 7.90 ms, a 1024-root shard 4.13 -> 4.08 with
               2; the claim itself — one atomic and one descriptor load per item — does not show: 1 team per item measures the same) */
            const int64_t warps = (int64_t)ctx->sm_count * 32;
            int g = (int)std::min<int64_t>(team_mode ? 2 : 8, std::max<int64_t>(1, (int64_t)nt * nU / (20 * warps)));
            if (const char* v = getenv("TPGX_K1_GROUP")) g = std::max(1, atoi(v)); /* experiments */
            bp.units_per_item = g;
            bp.items_per_tile = (nU + g - 1) / g;
            bp.nb_items = bp.items_per_tile * nt;
            CU(ctx->d_work.ensure(16));
            bp.work_counter = ctx->d_work.as<unsigned int>();
            bp.bid = ctx->d_bid.as<double>();
            bp.row_stride = (long long)nt * ts;
            bp.decision = ctx->d_decision.as<int32_t>();
            bp.tie_break = ctx->cfg.tie_break;
            int ctas = (int)std::min<int64_t>(ctx->sm_count, ((int64_t)bp.nb_items + 31) / 32);
            if (const char* v = getenv("TPGX_K1_CTAS")) ctas = std::max(1, atoi(v));
            CU(tpgx_launch_bid2(bp, team_mode, ctas, ctx->stream));
            etr.lap(team_mode ? "K1 v2 launch (team mode)" : "K1 v2 launch (bid mode)");
        } else if (nA > 0) {
            tpgx_bid_params bp{};
            bp.tiles = ctx->d_tiles.as<uint8_t>() + (size_t)t0 * (hw + 1) * ts;
            bp.hw = hw; bp.width = ctx->obs_width;
            bp.stream = ctx->d_k1_stream.as<uint4>();
            bp.desc = ctx->d_k1_desc.as<int2>();
            bp.pixel_lut = ctx->d_pixel_lut.as<double>();
            bp.sobel = ctx->nwin ? ctx->d_sobel.as<double>() + (size_t)t0 * 2 * ctx->nwin * ts : nullptr;
            bp.nwin = ctx->nwin;
            bp.nb_units = nU;
            /* chunks of work units per tile: ~12 waves of 148 CTAs when there is enough work, with >= ~14 teams (or 4
               programs) per warp for load balance; small problems trade balance for occupancy (>= 2 CTAs per SM in
               flight if every warp can still get a unit) */
            const int nw = tpgx_bid_warps(ctx->variant);
            int chunks = (int)std::max<int64_t>(1, (148 * 12 + nt - 1) / nt);
            chunks = std::min(chunks, std::max(1, nU / (team_mode ? 14 * nw : 4 * nw)));
            if ((int64_t)nt * chunks < 2 * 148) chunks = std::max(chunks, std::min((int)((2 * 148 + nt - 1) / nt), std::max(1, nU / nw)));
            else { /* one CTA per SM: prefer a nearby chunk count whose last wave of 148 CTAs is nearly full */
                const int cmax = std::min(chunks + 2, std::max(1, nU / (team_mode ? 10 * nw : 3 * nw)));
                double best_fill = 0.0;
                int best_c = chunks;
                for (int c = std::max(1, chunks - 1); c <= std::max(chunks, cmax); c++) {
                    const double waves = (double)nt * c / 148.0, fill = waves / std::ceil(waves);
                    if (fill > best_fill + 0.01) { best_fill = fill; best_c = c; }
                }
                chunks = best_c;
            }
            if (const char* v = getenv("TPGX_K1_CHUNKS")) chunks = std::max(1, std::min(atoi(v), nU)); /* experiments */
            bp.units_per_chunk = (nU + chunks - 1) / chunks;
            bp.bid = ctx->d_bid.as<double>();
            bp.row_stride = (long long)nt * ts;
            bp.team = ctx->d_tm_team.as<int2>();
            bp.row_dst = ctx->d_tm_dst.as<int32_t>();
            bp.decision = ctx->d_decision.as<int32_t>();
            bp.tie_break = ctx->cfg.tie_break;
            CU(tpgx_launch_bid(bp, nt, ctx->variant, ctx->plan.k1_ext, team_mode, ctx->stream));
            etr.lap(team_mode ? "K1 launch (team mode)" : "K1 launch (bid mode)");
        }
        if (timing) CU(cudaEventRecord(ctx->event(ev++), ctx->stream));
        if (k1_only) break; /* bid matrix of the (single) pass is all the caller wants */
        if (team_mode) {
            tpgx_walk_params wp{};
            wp.decision = ctx->d_decision.as<int32_t>();
            wp.row_stride = (long long)nt * ts;
            wp.nb_samples = ns;
            wp.team_stat = ctx->d_tm_stat.as<int2>();
            wp.root_unit = ctx->d_tm_roots.as<int32_t>();
            wp.nb_roots = nR;
            wp.labels = ctx->d_labels.as<uint8_t>() + s0;
            wp.nb_classes = C;
            wp.action = want_action ? ctx->d_action.as<uint8_t>() + s0 : nullptr;
            wp.action_stride = nS;
            wp.confusion = ctx->d_conf.as<unsigned long long>();
            wp.counters = ctx->d_counters.as<unsigned long long>();
            wp.status = ctx->d_status.as<int>();
            CU(tpgx_launch_walk(wp, ctx->stream));
        } else {
            tpgx_trav_params tp{};
            tp.bid = ctx->d_bid.as<double>();
            tp.row_stride = (long long)nt * ts;
            tp.nb_samples = ns;
            tp.team_edge_offset = ctx->d_team_off.as<int32_t>();
            tp.edge_row = ctx->d_edge_row.as<int32_t>();
            tp.edge_destination = ctx->d_edge_dst.as<int32_t>();
            tp.edge_lines = ctx->d_edge_lines.as<int32_t>();
            tp.roots = ctx->d_roots.as<int32_t>();
            tp.nb_roots = nR;
            tp.labels = ctx->d_labels.as<uint8_t>() + s0;
            tp.nb_classes = C;
            tp.action = want_action ? ctx->d_action.as<uint8_t>() + s0 : nullptr;
            tp.action_stride = nS;
            tp.confusion = ctx->d_conf.as<unsigned long long>();
            tp.counters = ctx->d_counters.as<unsigned long long>();
            tp.tie_break = ctx->cfg.tie_break;
            tp.status = ctx->d_status.as<int>();
            CU(tpgx_launch_traverse(tp, ctx->stream));
        }
        if (timing) CU(cudaEventRecord(ctx->event(ev++), ctx->stream));
    }
    if (!k1_only) CU(tpgx_launch_score(ctx->d_conf.as<unsigned long long>(), nR, C, records_out, ctx->stream));
    if (timing) CU(cudaEventRecord(ctx->event(ev++), ctx->stream));
    if (run) { run->nR = nR; run->C = C; run->ts = ts; run->nS = nS; run->nTiles = nTiles; run->passes = passes; }
    return TPGX_OK;
}

tpgx_status check_device_status(tpgx_ctx* ctx, int status, int encoder_flags = 0) {
    if (encoder_flags & 4) return fail(ctx, TPGX_ERR_CUDA, "K1 v2 encoder: device stream disagrees with the host's entry / slot counts");
    if (status & TPGX_DEV_NO_EDGE) return fail(ctx, TPGX_ERR_GRAPH_INVALID, "a team had no admissible outgoing edge (the reference throws \"No outgoing edge to evaluate\")");
    if (status & TPGX_DEV_PATH_TOO_LONG) return fail(ctx, TPGX_ERR_GRAPH_INVALID, "root->action path longer than TPGX_MAX_PATH teams");
    if (status & TPGX_DEV_BAD_ACTION) return fail(ctx, TPGX_ERR_GRAPH_INVALID, "action id outside the environment's action range (the reference throws in doAction)");
    if (status & TPGX_DEV_BAD_LABEL) return fail(ctx, TPGX_ERR_BAD_ARG, "observation label outside [0, nb_classes) (the reference's classificationTable.at() throws)");
    if (status & TPGX_DEV_RNG_EXHAUSTED) return fail(ctx, TPGX_ERR_STATE, "episode consumed more RNG draws than the table holds");
    return TPGX_OK;
}

/* End of an evaluation: device status and encoder flags come back with the stream's last copy (one synchronisation) */
static tpgx_status finish_stream(tpgx_ctx* ctx) {
    if (!ctx->h_status) CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->h_status), 64, cudaHostAllocDefault));
    CU(cudaMemcpyAsync(ctx->h_status, ctx->d_status.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return check_device_status(ctx, ctx->h_status[0], ctx->h_status[2]);
}

/* ---- NCCL, bound at run time: a Python process that imported torch already has libnccl.so.2 mapped (the loader hands
   that one back), a C++ host gets the system's; libtpgx.so itself carries no NCCL dependency */
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (h) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) { h = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
        if (!h) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
#define TPGX_NCCL_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(h, sym)); if (!field) { err = "libnccl lacks " sym; h = nullptr; return false; }
        TPGX_NCCL_SYM(GetUniqueId, "ncclGetUniqueId") TPGX_NCCL_SYM(CommInitRank, "ncclCommInitRank") TPGX_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        TPGX_NCCL_SYM(CommDestroy, "ncclCommDestroy") TPGX_NCCL_SYM(AllGather, "ncclAllGather") TPGX_NCCL_SYM(GroupStart, "ncclGroupStart")
        TPGX_NCCL_SYM(GroupEnd, "ncclGroupEnd") TPGX_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef TPGX_NCCL_SYM
        return true;
    }
};
NcclApi g_nccl;
#define NC(call)                                                                                                      \
    do {                                                                                                              \
        ncclResult_t r__ = (call);                                                                                    \
        if (r__ != ncclSuccess) return fail(ctx, TPGX_ERR_NCCL, std::string(#call) + ": " + g_nccl.GetErrorString(r__)); \
    } while (0)

void tpgx_comm_release(tpgx_ctx* ctx) {
    if (ctx->comm && g_nccl.h) g_nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr; ctx->comm_rank = 0; ctx->comm_world = 1;
}

void shard_of(const tpgx_ctx* ctx, int total, int* slot, int* rb, int* re) {
    const int s = (total + ctx->comm_world - 1) / ctx->comm_world;
    *slot = s;
    *rb = std::min(total, ctx->comm_rank * s);
    *re = std::min(total, (ctx->comm_rank + 1) * s);
}

/* K1 -> K2 -> K3 of this rank's shard with K3 writing into the rank's slot of the gather buffer; returns the buffer */
tpgx_status enqueue_shard_for_gather(tpgx_ctx* ctx, const tpgx_classif_request* req, int total, double** gather, size_t* count) {
    if (!req) return fail(ctx, TPGX_ERR_BAD_ARG, "null request");
    if (total < 1 || total > ctx->flat.nb_roots) return fail(ctx, TPGX_ERR_BAD_ARG, "total_roots exceeds the roots of the graph");
    int slot, rb, re;
    shard_of(ctx, total, &slot, &rb, &re);
    const int C = req->nb_classes;
    if (C < 1 || C > 16) return fail(ctx, TPGX_ERR_UNSUPPORTED, "nb_classes must be in [1,16]");
    *count = (size_t)slot * (C + 1);
    CU(ctx->d_gather.ensure(*count * ctx->comm_world * 8));
    double* mine = ctx->d_gather.as<double>() + (size_t)ctx->comm_rank * *count;
    CU(cudaMemsetAsync(mine, 0, *count * 8, ctx->stream)); /* the padding of an uneven last shard */
    if (re > rb) {
        tpgx_classif_request r = *req;
        r.root_begin = rb; r.root_end = re;
        const tpgx_status st = enqueue_classification(ctx, &r, mine, false, false, nullptr, false);
        if (st != TPGX_OK) return st;
    }
    *gather = ctx->d_gather.as<double>();
    return TPGX_OK;
}

} // namespace

extern "C" {

int32_t tpgx_abi_version(void) { return TPGX_ABI_VERSION; }
uint64_t tpgx_kernel_launches(void) { return tpgx_launch_counter().load(); }

tpgx_status tpgx_comm_unique_id(uint8_t* id) {
    tpgx_ctx* ctx = nullptr;
    if (!id) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_comm_unique_id: null id");
    static_assert(sizeof(ncclUniqueId) <= TPGX_COMM_ID_BYTES, "NCCL unique id does not fit");
    if (!g_nccl.load()) return fail(ctx, TPGX_ERR_NCCL, g_nccl.err);
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memset(id, 0, TPGX_COMM_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return TPGX_OK;
}

tpgx_status tpgx_comm_init(tpgx_ctx* ctx, const uint8_t* id, int32_t rank, int32_t world) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_comm_init: bad argument");
    if (!g_nccl.load()) return fail(ctx, TPGX_ERR_NCCL, g_nccl.err);
    cudaSetDevice(ctx->cfg.device);
    tpgx_comm_release(ctx);
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NC(g_nccl.CommInitRank(&ctx->comm, world, u, rank));
    ctx->comm_rank = rank; ctx->comm_world = world;
    return TPGX_OK;
}

tpgx_status tpgx_comm_init_all(tpgx_ctx* const* ctxs, int32_t n) {
    tpgx_ctx* ctx = (ctxs && n > 0) ? ctxs[0] : nullptr;
    if (!ctx) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_comm_init_all: bad argument");
    if (!g_nccl.load()) return fail(ctx, TPGX_ERR_NCCL, g_nccl.err);
    std::vector<int> devs((size_t)n);
    std::vector<ncclComm_t> comms((size_t)n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_comm_init_all: null context");
        for (int j = 0; j < i; j++) if (ctxs[j]->cfg.device == ctxs[i]->cfg.device) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_comm_init_all: two contexts on one device");
        devs[i] = ctxs[i]->cfg.device;
        tpgx_comm_release(ctxs[i]);
    }
    NC(g_nccl.CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->comm_rank = i; ctxs[i]->comm_world = n; }
    return TPGX_OK;
}

tpgx_status tpgx_comm_shard(const tpgx_ctx* ctx, int32_t total_roots, int32_t* rank, int32_t* world, int32_t* root_begin, int32_t* root_end) {
    if (!ctx || total_roots < 0) return TPGX_ERR_BAD_ARG;
    int slot, rb, re;
    shard_of(ctx, total_roots, &slot, &rb, &re);
    if (rank) *rank = ctx->comm_rank;
    if (world) *world = ctx->comm_world;
    if (root_begin) *root_begin = rb;
    if (root_end) *root_end = re;
    return TPGX_OK;
}

tpgx_status tpgx_evaluate_classification_allgather_device(tpgx_ctx* ctx, const tpgx_classif_request* req, int32_t total_roots,
                                                          double* d_records_all, void* stream) {
    if (!ctx || !req || !d_records_all) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_classification_allgather_device: null argument");
    if (ctx->comm_world > 1 && !ctx->comm) return fail(ctx, TPGX_ERR_STATE, "no communicator: call tpgx_comm_init first");
    cudaSetDevice(ctx->cfg.device);
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    cudaEvent_t ev_in = ctx->event(62), ev_out = ctx->event(63);
    CU(cudaEventRecord(ev_in, user));
    CU(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
    double* gather = nullptr;
    size_t count = 0;
    tpgx_status st = enqueue_shard_for_gather(ctx, req, total_roots, &gather, &count);
    if (st != TPGX_OK) return st;
    if (ctx->comm_world > 1) NC(g_nccl.AllGather(gather + (size_t)ctx->comm_rank * count, gather, count, ncclDouble, ctx->comm, ctx->stream));
    /* slots are per rank; with slot * world >= total the first total_roots records are the roots in order */
    CU(cudaMemcpyAsync(d_records_all, gather, (size_t)total_roots * (req->nb_classes + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaEventRecord(ev_out, ctx->stream));
    CU(cudaStreamWaitEvent(user, ev_out, 0));
    return TPGX_OK;
}

tpgx_status tpgx_evaluate_classification_allgather(tpgx_ctx* ctx, const tpgx_classif_request* req, int32_t total_roots, double* records_all) {
    if (!ctx || !req) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_classification_allgather: null argument");
    if (ctx->comm_world > 1 && !ctx->comm) return fail(ctx, TPGX_ERR_STATE, "no communicator: call tpgx_comm_init first");
    cudaSetDevice(ctx->cfg.device);
    double* gather = nullptr;
    size_t count = 0;
    tpgx_status st = enqueue_shard_for_gather(ctx, req, total_roots, &gather, &count);
    if (st != TPGX_OK) return st;
    if (ctx->comm_world > 1) NC(g_nccl.AllGather(gather + (size_t)ctx->comm_rank * count, gather, count, ncclDouble, ctx->comm, ctx->stream));
    if (records_all) CU(cudaMemcpyAsync(records_all, gather, (size_t)total_roots * (req->nb_classes + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return finish_stream(ctx);
}

tpgx_status tpgx_evaluate_classification_allgather_all(tpgx_ctx* const* ctxs, int32_t n, const tpgx_classif_request* req, int32_t total_roots,
                                                       double* records_all) {
    tpgx_ctx* ctx = (ctxs && n > 0) ? ctxs[0] : nullptr;
    if (!ctx || !req) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_classification_allgather_all: bad argument");
    std::vector<double*> gather((size_t)n, nullptr);
    size_t count = 0;
    for (int i = 0; i < n; i++) { /* every device works on its shard at the same time */
        tpgx_ctx* c = ctxs[i];
        if (!c || c->comm_world != n || c->comm_rank != i || (n > 1 && !c->comm)) return fail(ctx, TPGX_ERR_STATE, "contexts are not the ranks of one tpgx_comm_init_all");
        cudaSetDevice(c->cfg.device);
        const tpgx_status st = enqueue_shard_for_gather(c, req, total_roots, &gather[i], &count);
        if (st != TPGX_OK) { if (c != ctx) ctx->err = c->err; return st; }
    }
    if (n > 1) {
        NC(g_nccl.GroupStart());
        for (int i = 0; i < n; i++)
            NC(g_nccl.AllGather(gather[i] + (size_t)i * count, gather[i], count, ncclDouble, ctxs[i]->comm, ctxs[i]->stream));
        NC(g_nccl.GroupEnd());
    }
    for (int i = 0; i < n; i++) {
        tpgx_ctx* c = ctxs[i];
        cudaSetDevice(c->cfg.device);
        if (i == 0 && records_all) CU(cudaMemcpyAsync(records_all, gather[0], (size_t)total_roots * (req->nb_classes + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        const tpgx_status st = finish_stream(c);
        if (st != TPGX_OK) { if (c != ctx) ctx->err = c->err; return st; }
    }
    return TPGX_OK;
}


const char* tpgx_last_error(const tpgx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

tpgx_status tpgx_create(const tpgx_config* cfg, tpgx_ctx** out) {
    tpgx_ctx* ctx = nullptr;
    if (!cfg || !out) return fail(nullptr, TPGX_ERR_BAD_ARG, "tpgx_create: null argument");
    *out = nullptr;
    if (cfg->nb_registers < 1 || cfg->nb_registers > TPGX_MAX_REGS) return fail(nullptr, TPGX_ERR_UNSUPPORTED, "nb_registers must be in [1,8]");
    if (cfg->nb_constants < 0 || cfg->nb_constants > TPGX_MAX_CONSTS) return fail(nullptr, TPGX_ERR_UNSUPPORTED, "nb_constants must be in [0,8]");
    if (cfg->tie_break != TPGX_TIE_LAST_MAX && cfg->tie_break != TPGX_TIE_FIRST_MAX) return fail(nullptr, TPGX_ERR_BAD_ARG, "unknown tie_break");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TPGX_ERR_CUDA, std::string("tpgx_create: no CUDA device (") + cudaGetErrorString(e) + "); libtpgx has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TPGX_ERR_BAD_ARG, "device ordinal out of range");
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) return fail(nullptr, TPGX_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major < 10) return fail(nullptr, TPGX_ERR_UNSUPPORTED, "libtpgx is built for sm_100a (Blackwell) only");
    ctx = new tpgx_ctx();
    ctx->cfg = *cfg;
    ctx->sm_count = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->obs_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->obs_ready, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete ctx; return fail(nullptr, TPGX_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
    if (ctx->d_pixel_lut.ensure((size_t)TPGX_LUT_DOUBLES * 8) != cudaSuccess || tpgx_launch_pixel_lut(ctx->d_pixel_lut.as<double>(), ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, TPGX_ERR_CUDA, "tpgx_create: pixel table initialisation failed");
    }
    if (const char* v = getenv("TPGX_K1_VARIANT")) ctx->variant = atoi(v);
    if (const char* v = getenv("TPGX_BID_BUDGET_MB")) ctx->bid_budget = (size_t)atoll(v) << 20;
    *out = ctx;
    return TPGX_OK;
}

tpgx_status tpgx_destroy(tpgx_ctx* ctx) {
    if (!ctx) return TPGX_OK;
    for (tpgx_ctx* lane : ctx->lanes) tpgx_destroy(lane);
    ctx->lanes.clear();
    cudaSetDevice(ctx->cfg.device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->obs_stream) cudaStreamSynchronize(ctx->obs_stream);
    DevBuf* all[] = {&ctx->d_pixel_lut, &ctx->d_sobel, &ctx->d_tm_team, &ctx->d_tm_dst, &ctx->d_tm_stat, &ctx->d_tm_roots, &ctx->d_decision, &ctx->d_code, &ctx->d_k1_stream, &ctx->d_k1_desc, &ctx->d_prog_off, &ctx->d_consts, &ctx->d_team_off, &ctx->d_edge_dst, &ctx->d_edge_row,
                     &ctx->d_edge_lines, &ctx->d_roots, &ctx->d_active, &ctx->d_obs_raw, &ctx->d_labels_raw, &ctx->d_tiles,
                     &ctx->d_labels, &ctx->d_index, &ctx->d_bid, &ctx->d_conf, &ctx->d_records, &ctx->d_action,
                     &ctx->d_counters, &ctx->d_status, &ctx->d_work, &ctx->d_k2_rflags, &ctx->d_k2_ufirst, &ctx->d_gather, &ctx->d_edge_prog};
    for (DevBuf* b : all) b->release();
    for (DevBuf& b : ctx->d_scratch) b.release();
    tpgx_comm_release(ctx);
    if (ctx->pin_ring) cudaFreeHost(ctx->pin_ring);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    if (ctx->obs_stream) { cudaStreamSynchronize(ctx->obs_stream); cudaStreamDestroy(ctx->obs_stream); }
    if (ctx->obs_ready) cudaEventDestroy(ctx->obs_ready);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TPGX_OK;
}

tpgx_status tpgx_set_instruction_table(tpgx_ctx* ctx, int32_t n, const int32_t* opcode) {
    if (!ctx || !opcode || n < 1) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_instruction_table: bad argument");
    for (int i = 0; i < n; i++)
        if (opcode[i] < 0 || opcode[i] >= TPGX_OP_COUNT)
            return fail(ctx, TPGX_ERR_UNSUPPORTED, "instruction " + std::to_string(i) + " has no device opcode (no CPU fallback by design)");
    ctx->opcode.assign(opcode, opcode + n);
    ctx->has_graph = false;
    return TPGX_OK;
}

tpgx_status tpgx_set_data_sources(tpgx_ctx* ctx, int32_t n, const tpgx_source_desc* src) {
    if (!ctx || n < 1 || !src) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_data_sources: bad argument");
    if (n > 2) return fail(ctx, TPGX_ERR_UNSUPPORTED, "at most 2 data sources");
    for (int i = 0; i < n; i++) {
        if (src[i].height < 1 || src[i].width < 1) return fail(ctx, TPGX_ERR_BAD_ARG, "source dimensions must be positive");
        if ((int64_t)src[i].height * src[i].width > TPGX_MAX_LOC) return fail(ctx, TPGX_ERR_UNSUPPORTED, "source address space above 1024 elements");
        if (src[i].elem != TPGX_ELEM_F64 && src[i].elem != TPGX_ELEM_I32) return fail(ctx, TPGX_ERR_BAD_ARG, "unknown element type");
    }
    ctx->src.assign(src, src + n);
    ctx->has_graph = false;
    /* everything derived from the previous source shape goes: uploaded observations, their tiles and Sobel planes, the plan */
    ctx->obs_dev = nullptr;
    ctx->labels_dev = nullptr;
    ctx->nb_obs = 0;
    ctx->obs_hw = ctx->obs_width = ctx->nwin = 0;
    ctx->tiles_identity_valid = false;
    ctx->plan = ShardPlan();
    return TPGX_OK;
}

static tpgx_status set_graph_impl(tpgx_ctx* ctx, const tpgx_graph* g, bool shard_only, int range_begin, int range_end);
static tpgx_status upload_flat_graph(tpgx_ctx* ctx, const double* constants, const std::vector<uint8_t>& mask);
tpgx_status tpgx_set_graph(tpgx_ctx* ctx, const tpgx_graph* g) { return set_graph_impl(ctx, g, false, -1, -1); }
tpgx_status tpgx_set_graph_sharded(tpgx_ctx* ctx, const tpgx_graph* g) { return set_graph_impl(ctx, g, true, -1, -1); }

/* range_begin >= 0: only the programs reachable from roots [range_begin, range_end) (a pipeline slice) */
static tpgx_status set_graph_impl(tpgx_ctx* ctx, const tpgx_graph* g, bool shard_only, int range_begin, int range_end) {
    if (!ctx) return TPGX_ERR_BAD_ARG;
    if (ctx->opcode.empty()) return fail(ctx, TPGX_ERR_STATE, "tpgx_set_graph: instruction table not set");
    if (ctx->src.empty()) return fail(ctx, TPGX_ERR_STATE, "tpgx_set_graph: data sources not set");
    cudaSetDevice(ctx->cfg.device);
    std::string msg;
    tpgx_flat_graph f;
    Trace tr("set_graph");
    /* sharded: only the programs this context's root shard can reach are flattened and encoded (a rank of N pays 1/N of the
       host work); the rest of the graph stays empty and evaluating a root outside the shard is refused */
    std::vector<uint8_t> mask;
    ctx->flat_root_begin = 0; ctx->flat_root_end = -1;
    if (range_begin >= 0) shard_only = true;
    if (shard_only && (ctx->comm_world > 1 || range_begin >= 0) && g && g->nb_roots > 0 && g->nb_teams > 0 && g->nb_programs > 0 && g->team_edge_offset && g->edge_program &&
        g->edge_destination && g->root_team) {
        int slot, rb, re;
        shard_of(ctx, g->nb_roots, &slot, &rb, &re);
        if (range_begin >= 0) { rb = std::min(range_begin, g->nb_roots); re = std::min(std::max(range_end, rb), g->nb_roots); }
        mask.assign((size_t)g->nb_programs, 0);
        std::vector<uint8_t> seen((size_t)g->nb_teams, 0);
        std::vector<int32_t> queue;
        bool ok = true;
        for (int r = rb; r < re && ok; r++) {
            const int t = g->root_team[r];
            if (t < 0 || t >= g->nb_teams) { ok = false; break; }
            if (!seen[(size_t)t]) { seen[(size_t)t] = 1; queue.push_back(t); }
        }
        for (size_t q = 0; q < queue.size() && ok; q++) {
            const int t = queue[q];
            const int eb = g->team_edge_offset[t], ee = g->team_edge_offset[t + 1];
            if (eb < 0 || ee < eb) { ok = false; break; }
            for (int e = eb; e < ee; e++) {
                const int p = g->edge_program[e], d = g->edge_destination[e];
                if (p < 0 || p >= g->nb_programs || d >= g->nb_teams) { ok = false; break; }
                mask[(size_t)p] = 1;
                if (d >= 0 && !seen[(size_t)d]) { seen[(size_t)d] = 1; queue.push_back(d); }
            }
        }
        if (ok) { ctx->flat_root_begin = rb; ctx->flat_root_end = re; }
        else mask.clear(); /* malformed graph: the full flattener reports it */
    }
    tpgx_status st = tpgx_flatten_graph(ctx->cfg, ctx->opcode, ctx->src, g, &f, &msg, nullptr, mask.empty() ? nullptr : &mask);
    if (st != TPGX_OK) return fail(ctx, st, msg);
    tr.lap("flatten");
    ctx->flat = std::move(f);
    st = upload_flat_graph(ctx, g->constants, mask);
    tr.lap("uploads enqueued"); /* no synchronisation: upload() has consumed its sources */
    return st;
}

/* The flattened graph (ctx->flat) to the device.  constants: [nb_programs][nb_constants] in the numbering of ctx->flat; mask (may be
   empty): the programs that were flattened — only their constants travel. */
static tpgx_status upload_flat_graph(tpgx_ctx* ctx, const double* constants, const std::vector<uint8_t>& mask) {
    tpgx_status st;
    const tpgx_flat_graph& F = ctx->flat;
    ctx->uses_trig = ctx->uses_ext = ctx->uses_int = false;
    { /* instruction classes present among the effective lines (selects the K1 build, rejects int ops on image sources); the
         flattener also ran the count pass of the K1 v2 encoder (entries and shared-memory slots per program; the stream
         itself is written on the device) */
        auto present = [&](int op) { return ((F.op_mask >> op) & 1u) != 0; };
        ctx->uses_trig = present(TPGX_OP_SIN) || present(TPGX_OP_COS) || present(TPGX_OP_TAN);
        ctx->uses_int = present(TPGX_OP_SUB_II) || present(TPGX_OP_CAST_I);
        ctx->uses_ext = ctx->uses_trig || present(TPGX_OP_FMODG) || present(TPGX_OP_COND) || present(TPGX_OP_EQ0) || present(TPGX_OP_EQM1) ||
                        present(TPGX_OP_EQ1) || present(TPGX_OP_GE15) || present(TPGX_OP_PI);
    }
    std::vector<int32_t> edge_lines(F.nb_edges);
    for (int e = 0; e < F.nb_edges; e++) { int p = F.edge_program[e]; edge_lines[e] = F.prog_offset[p + 1] - F.prog_offset[p]; }
    if ((st = upload(ctx, ctx->d_code, F.code.data(), F.code.size() * 4)) != TPGX_OK) return st;
    if ((st = upload(ctx, ctx->d_prog_off, F.prog_offset.data(), F.prog_offset.size() * 4)) != TPGX_OK) return st;
    /* program constants straight from the caller's array; a shard sends the runs of programs it reaches only */
    if (ctx->cfg.nb_constants > 0 && F.nb_programs > 0) {
        const size_t row = (size_t)ctx->cfg.nb_constants * 8;
        CU(ctx->d_consts.ensure((size_t)F.nb_programs * row));
        if (mask.empty()) {
            if ((st = upload_at(ctx, ctx->d_consts, 0, constants, (size_t)F.nb_programs * row)) != TPGX_OK) return st;
        } else {
            for (int p = 0; p < F.nb_programs;) {
                if (!mask[(size_t)p]) { p++; continue; }
                int q = p;
                while (q < F.nb_programs && mask[(size_t)q]) q++;
                if ((st = upload_at(ctx, ctx->d_consts, (size_t)p * row, constants + (size_t)p * ctx->cfg.nb_constants, (size_t)(q - p) * row)) != TPGX_OK) return st;
                p = q;
            }
        }
    } else CU(ctx->d_consts.ensure(16));
    if ((st = upload(ctx, ctx->d_team_off, F.team_edge_offset.data(), F.team_edge_offset.size() * 4)) != TPGX_OK) return st;
    if ((st = upload(ctx, ctx->d_edge_dst, F.edge_destination.data(), F.edge_destination.size() * 4)) != TPGX_OK) return st;
    if ((st = upload(ctx, ctx->d_edge_lines, edge_lines.data(), edge_lines.size() * 4)) != TPGX_OK) return st;
    if ((st = upload(ctx, ctx->d_edge_prog, F.edge_program.data(), F.edge_program.size() * 4)) != TPGX_OK) return st;
    ctx->graph_version++;
    ctx->has_graph = true;
    return TPGX_OK;
}

/* The sub-graph roots [rb, re) of g reach, renumbered: teams breadth-first from the roots in root order (M.teams: new id ->
   the caller's team), programs in order of first appearance (M.progs), every team's edges in the caller's order (it decides
   ties).  Stamped tables: nothing is cleared between calls. */
static tpgx_status build_slice(SliceMaps& M, const tpgx_graph* g, int rb, int re, std::string* why) {
    auto bad = [&](const std::string& m) { if (why) *why = m; return TPGX_ERR_BAD_ARG; };
    if ((int)M.team_stamp.size() < g->nb_teams) { M.team_stamp.assign((size_t)g->nb_teams, 0u); M.team_id.resize((size_t)g->nb_teams); }
    if ((int)M.prog_stamp.size() < g->nb_programs) { M.prog_stamp.assign((size_t)g->nb_programs, 0u); M.prog_id.resize((size_t)g->nb_programs); }
    if (++M.stamp == 0) { std::fill(M.team_stamp.begin(), M.team_stamp.end(), 0u); std::fill(M.prog_stamp.begin(), M.prog_stamp.end(), 0u); M.stamp = 1; }
    const uint32_t stamp = M.stamp;
    std::vector<int32_t>&teams = M.teams, &progs = M.progs, &team_off = M.team_off, &edge_prog = M.edge_prog, &edge_dst = M.edge_dst, &roots = M.roots;
    teams.clear(); progs.clear(); team_off.clear(); edge_prog.clear(); edge_dst.clear(); roots.clear();
    auto team_of = [&](int t) {
        if (M.team_stamp[(size_t)t] != stamp) { M.team_stamp[(size_t)t] = stamp; M.team_id[(size_t)t] = (int32_t)teams.size(); teams.push_back(t); }
        return M.team_id[(size_t)t];
    };
    for (int r = rb; r < re; r++) {
        const int t = g->root_team[r];
        if (t < 0 || t >= g->nb_teams) return bad("root team out of range");
        roots.push_back(team_of(t));
    }
    team_off.push_back(0);
    for (size_t q = 0; q < teams.size(); q++) {
        const int t = teams[q];
        const int eb = g->team_edge_offset[t], ee = g->team_edge_offset[t + 1];
        if (eb < 0 || ee < eb) return bad("team_edge_offset not monotone");
        for (int e = eb; e < ee; e++) {
            const int p = g->edge_program[e], d = g->edge_destination[e];
            if (p < 0 || p >= g->nb_programs) return bad("edge " + std::to_string(e) + ": program out of range");
            if (d >= g->nb_teams) return bad("edge " + std::to_string(e) + ": destination team out of range");
            if (M.prog_stamp[(size_t)p] != stamp) { M.prog_stamp[(size_t)p] = stamp; M.prog_id[(size_t)p] = (int32_t)progs.size(); progs.push_back(p); }
            edge_prog.push_back(M.prog_id[(size_t)p]);
            edge_dst.push_back(d >= 0 ? team_of(d) : d);
        }
        team_off.push_back((int32_t)edge_prog.size());
    }
    return TPGX_OK;
}

/* A slice of the caller's graph — roots [rb, re) and what they reach — as a graph of its own on `ctx`: teams, edges and
   programs renumbered in breadth-first order (edge order within a team kept: it decides ties), roots 0 .. re - rb - 1.  The
   lines stay where they are (the flattener follows prog_ids); the constants are gathered.  Every cost — flatten, plan,
   uploads — scales with the slice, none with the graph: a 1 / 16 slice of an 8192-root graph paid 0.5 ms of fixed
   full-graph work (array copies, zero fills, 0.7 MB of uploads) per call before.  `owner` keeps the renumbering tables
   (stamped, never cleared). */
static tpgx_status set_graph_slice(tpgx_ctx* ctx, tpgx_ctx* owner, const tpgx_graph* g, int rb, int re) {
    if (ctx->opcode.empty()) return fail(ctx, TPGX_ERR_STATE, "tpgx_set_graph: instruction table not set");
    if (ctx->src.empty()) return fail(ctx, TPGX_ERR_STATE, "tpgx_set_graph: data sources not set");
    if (!g || !g->program_line_offset || !g->lines || !g->team_edge_offset || !g->edge_program || !g->edge_destination || !g->root_team)
        return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_graph: null array in graph");
    if (g->nb_programs < 1 || g->nb_teams < 1 || rb < 0 || re > g->nb_roots || rb >= re) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_graph: empty graph slice");
    if (ctx->cfg.nb_constants > 0 && !g->constants) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_graph: constants missing");
    cudaSetDevice(ctx->cfg.device);
    Trace tr("set_graph (slice)");
    SliceMaps& M = owner->slice_maps;
    {
        std::string why;
        const tpgx_status bst = build_slice(M, g, rb, re, &why);
        if (bst != TPGX_OK) return fail(ctx, bst, why);
    }
    std::vector<int32_t>&progs = M.progs, &teams = M.teams, &team_off = M.team_off, &edge_prog = M.edge_prog, &edge_dst = M.edge_dst, &roots = M.roots;
    const int nc = ctx->cfg.nb_constants;
    std::vector<double>& consts = M.consts;
    consts.resize((size_t)progs.size() * (size_t)std::max(nc, 0));
    if (nc > 0)
        for (size_t p = 0; p < progs.size(); p++) memcpy(consts.data() + p * nc, g->constants + (size_t)progs[p] * nc, (size_t)nc * 8);
    tpgx_graph cg = *g;
    cg.nb_programs = (int32_t)progs.size();
    cg.intron = nullptr;
    cg.constants = nc > 0 ? consts.data() : nullptr;
    cg.nb_teams = (int32_t)teams.size();
    cg.team_edge_offset = team_off.data();
    cg.edge_program = edge_prog.data();
    cg.edge_destination = edge_dst.data();
    cg.nb_roots = re - rb;
    cg.root_team = roots.data();
    tr.lap("sub-graph");
    std::string msg;
    tpgx_flat_graph f;
    tpgx_status st = tpgx_flatten_graph(ctx->cfg, ctx->opcode, ctx->src, &cg, &f, &msg, nullptr, nullptr, progs.data());
    if (st != TPGX_OK) return fail(ctx, st, msg);
    tr.lap("flatten");
    ctx->flat = std::move(f);
    ctx->flat_root_begin = 0; ctx->flat_root_end = -1;
    st = upload_flat_graph(ctx, cg.constants, std::vector<uint8_t>());
    tr.lap("uploads enqueued");
    return st;
}

/* Host only (include/tpgx.h): the renumbered sub-graph of set_graph_slice on its own. */
tpgx_status tpgx_slice_graph(const tpgx_graph* g, int32_t root_begin, int32_t root_end, int32_t* nb_teams, int32_t* nb_edges, int32_t* nb_programs,
                             int32_t* team_ids, int32_t* team_edge_offset, int32_t* edge_program, int32_t* edge_destination, int32_t* program_ids,
                             int32_t* root_team) {
    if (!g || !g->team_edge_offset || !g->edge_program || !g->edge_destination || !g->root_team || !nb_teams || !nb_edges || !nb_programs) return TPGX_ERR_BAD_ARG;
    if (g->nb_programs < 1 || g->nb_teams < 1 || root_begin < 0 || root_end > g->nb_roots || root_begin >= root_end) return TPGX_ERR_BAD_ARG;
    SliceMaps M;
    const tpgx_status st = build_slice(M, g, root_begin, root_end, nullptr);
    if (st != TPGX_OK) return st;
    /* sizes first; the arrays are filled when the caller's counts (from a first call) say they are large enough */
    const bool fits = *nb_teams >= (int32_t)M.teams.size() && *nb_edges >= (int32_t)M.edge_prog.size() && *nb_programs >= (int32_t)M.progs.size();
    *nb_teams = (int32_t)M.teams.size(); *nb_edges = (int32_t)M.edge_prog.size(); *nb_programs = (int32_t)M.progs.size();
    if (!fits || !team_ids || !team_edge_offset || !edge_program || !edge_destination || !program_ids || !root_team) return TPGX_OK;
    std::copy(M.teams.begin(), M.teams.end(), team_ids);
    std::copy(M.team_off.begin(), M.team_off.end(), team_edge_offset);
    std::copy(M.edge_prog.begin(), M.edge_prog.end(), edge_program);
    std::copy(M.edge_dst.begin(), M.edge_dst.end(), edge_destination);
    std::copy(M.progs.begin(), M.progs.end(), program_ids);
    std::copy(M.roots.begin(), M.roots.end(), root_team);
    return TPGX_OK;
}

/* One generation in one call: tpgx_set_graph_sharded + tpgx_evaluate_classification_allgather, with the host side of the
   graph (flatten, stream encoding counts, plan, uploads: ~0.1 us per line) pipelined behind the device work.  The rank's
   root shard is cut into slices; slice k is flattened and planned on the host while the kernels of slice k - 1 run.  Every
   slice has a context of its own ("lane": stream, graph buffers, plan) that shares the parent's observation tiles and
   writes its records straight into the parent's gather buffer; the collective and the read-back stay on the parent. */
tpgx_status tpgx_evaluate_graph_classification_allgather(tpgx_ctx* ctx, const tpgx_graph* g, const tpgx_classif_request* req, int32_t total_roots,
                                                         double* records_all) {
    if (!ctx || !g || !req) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_graph_classification_allgather: null argument");
    if (ctx->comm_world > 1 && !ctx->comm) return fail(ctx, TPGX_ERR_STATE, "no communicator: call tpgx_comm_init first");
    if (total_roots < 1 || total_roots > g->nb_roots) return fail(ctx, TPGX_ERR_BAD_ARG, "total_roots exceeds the roots of the graph");
    if (!ctx->obs_dev) return fail(ctx, TPGX_ERR_STATE, "tpgx_evaluate_graph_classification_allgather: no observations set");
    const int C = req->nb_classes;
    if (C < 1 || C > 16) return fail(ctx, TPGX_ERR_UNSUPPORTED, "nb_classes must be in [1,16]");
    cudaSetDevice(ctx->cfg.device);
    int slot, rb, re;
    shard_of(ctx, total_roots, &slot, &rb, &re);
    /* slices of 256 roots or more, at most 4.  Measured with every slice flattened as a sub-graph of its own (no cost of a slice
       scales with the whole graph any more): 1024 roots per GPU on 8 GPUs 6.26 / 5.55 / 5.52 / 5.43 / 5.60 ms with 1 / 2 / 3 / 4 / 6
       slices against 6.93 ms in two calls; 2000 roots on one GPU 9.06 / 9.06 ms with 3 / 4 */
    int want = 4, min_roots = 256;
    if (const char* v = getenv("TPGX_PIPELINE")) { want = std::max(1, std::min(8, atoi(v))); min_roots = 128; }
    const int K = std::max(1, std::min(want, (re - rb) / min_roots)); /* one slice: no pipelining, but still the slice's own sub-graph */
    int first_pm = 0;
    if (const char* v = getenv("TPGX_PIPELINE_FIRST")) first_pm = std::max(0, std::min(900, atoi(v)));
    ctx->has_graph = false; /* the parent keeps no graph after this call */
    if (re <= rb) { /* an empty shard (more ranks than roots): nothing of the graph to set, the collective still runs */
        tpgx_status st = set_graph_impl(ctx, g, true, rb, std::max(re, rb));
        if (st != TPGX_OK) return st;
        return tpgx_evaluate_classification_allgather(ctx, req, total_roots, records_all);
    }
    Trace tr("evaluate_graph");
    /* lanes: created once, configured like the parent */
    while ((int)ctx->lanes.size() < K) {
        tpgx_ctx* lane = nullptr;
        const tpgx_status st = tpgx_create(&ctx->cfg, &lane);
        if (st != TPGX_OK) return fail(ctx, st, std::string("pipeline lane: ") + tpgx_last_error(nullptr));
        ctx->lanes.push_back(lane);
    }
    const size_t count = (size_t)slot * (C + 1);
    CU(ctx->d_gather.ensure(count * ctx->comm_world * 8));
    double* mine = ctx->d_gather.as<double>() + (size_t)ctx->comm_rank * count;
    CU(cudaMemsetAsync(mine, 0, count * 8, ctx->stream)); /* the padding of an uneven last shard */
    CU(wait_for_observations(ctx));
    /* the request's sample tiles (a sample_index re-tiles) are prepared once, on the parent */
    const int ts = tpgx_bid_tile_samples(ctx->variant);
    {
        const int64_t nS = req->nb_samples;
        if (nS <= 0) return fail(ctx, TPGX_ERR_BAD_ARG, "nb_samples must be positive");
        if (!req->sample_index && nS > ctx->nb_obs) return fail(ctx, TPGX_ERR_BAD_ARG, "nb_samples exceeds uploaded observations");
        if (!ctx->labels_dev) return fail(ctx, TPGX_ERR_STATE, "classification needs labels");
        tpgx_status st = prepare_tiles(ctx, req, ts);
        if (st != TPGX_OK) return st;
    }
    cudaEvent_t start = ctx->event(58);
    CU(cudaEventRecord(start, ctx->stream));
    tr.lap("lanes + tiles");
    for (int k = 0; k < K; k++) {
        tpgx_ctx* lane = ctx->lanes[k];
        /* the first slice is half as long as the others: the device starts after 1 / (2K - 1) of the host work */
        auto cut = [&](int i) {
            if (i <= 0) return rb;
            if (i >= K) return re;
            if (first_pm > 0) { /* experiments: the first slice takes first_pm / 1000 of the shard, the others share the rest */
                const int first = std::max(1, (int)((int64_t)(re - rb) * first_pm / 1000));
                return rb + first + (int)((int64_t)(re - rb - first) * (i - 1) / (K - 1));
            }
            return rb + (int)((int64_t)(re - rb) * (2 * i - 1) / (2 * K - 1));
        };
        const int b = cut(k), e = cut(k + 1);
        lane->opcode = ctx->opcode; lane->src = ctx->src; lane->variant = ctx->variant; lane->bid_budget = ctx->bid_budget;
        lane->d_tiles.alias(ctx->d_tiles); lane->d_labels.alias(ctx->d_labels); lane->d_sobel.alias(ctx->d_sobel);
        lane->obs_dev = ctx->obs_dev; lane->labels_dev = ctx->labels_dev; lane->nb_obs = ctx->nb_obs; lane->obs_hw = ctx->obs_hw;
        lane->obs_width = ctx->obs_width; lane->nwin = ctx->nwin; lane->tiles_ts = ts; lane->tiles_shared = true; lane->obs_pending = false;
        if (cudaStreamWaitEvent(lane->stream, start, 0) != cudaSuccess) return fail(ctx, TPGX_ERR_CUDA, "cudaStreamWaitEvent (lane)");
        tpgx_status st = set_graph_slice(lane, ctx, g, b, e); /* roots b .. e - 1 are roots 0 .. e - b - 1 of the lane's graph */
        if (st == TPGX_OK) {
            tpgx_classif_request r = *req;
            r.root_begin = 0; r.root_end = e - b;
            st = enqueue_classification(lane, &r, mine + (size_t)(b - rb) * (C + 1), false, false, nullptr, false);
        }
        if (st != TPGX_OK) { ctx->err = lane->err; return st; }
        cudaEvent_t done = lane->event(61);
        if (cudaEventRecord(done, lane->stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, done, 0) != cudaSuccess)
            return fail(ctx, TPGX_ERR_CUDA, "cudaEventRecord (lane)");
        tr.lap("slice enqueued");
    }
    if (ctx->comm_world > 1) NC(g_nccl.AllGather(mine, ctx->d_gather.as<double>(), count, ncclDouble, ctx->comm, ctx->stream));
    if (records_all) CU(cudaMemcpyAsync(records_all, ctx->d_gather.p, (size_t)total_roots * (C + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx->d_status.ensure(16));
    CU(cudaMemsetAsync(ctx->d_status.p, 0, 16, ctx->stream));
    tpgx_status st = finish_stream(ctx);
    for (int k = 0; k < K && st == TPGX_OK; k++) {
        st = finish_stream(ctx->lanes[k]);
        if (st != TPGX_OK) ctx->err = ctx->lanes[k]->err;
    }
    tr.lap("device + read-back");
    return st;
}

static tpgx_status set_obs_common(tpgx_ctx* ctx, int32_t source, int32_t store, int64_t count) {
    if (!ctx) return TPGX_ERR_BAD_ARG;
    if (ctx->src.empty()) return fail(ctx, TPGX_ERR_STATE, "tpgx_set_observations: data sources not set");
    if (source != 0 || ctx->src.size() != 1) return fail(ctx, TPGX_ERR_UNSUPPORTED, "batched observations are supported for a single data source (index 0)");
    if (store != TPGX_STORE_U8) return fail(ctx, TPGX_ERR_UNSUPPORTED, "batched observations must be stored as uint8 (exact for MNIST pixels)");
    if (ctx->src[0].elem != TPGX_ELEM_F64) return fail(ctx, TPGX_ERR_BAD_ARG, "uint8 observations stand in for a double source");
    if (count < 1) return fail(ctx, TPGX_ERR_BAD_ARG, "count must be positive");
    ctx->obs_hw = ctx->src[0].height * ctx->src[0].width;
    ctx->obs_width = ctx->src[0].width;
    if (tpgx_bid_smem_bytes(ctx->variant, ctx->obs_hw) > 227 * 1024)
        return fail(ctx, TPGX_ERR_UNSUPPORTED, "observation tile does not fit in shared memory");
    ctx->nb_obs = count;
    ctx->tiles_identity_valid = false;
    return TPGX_OK;
}

/* identity packing of all observations, done once at upload so evaluations over the full set start
 * from the tiled layout that lives in HBM */
static tpgx_status pack_identity(tpgx_ctx* ctx) {
    const int ts = tpgx_bid_tile_samples(ctx->variant);
    const int64_t nTiles = (ctx->nb_obs + ts - 1) / ts;
    CU(ctx->d_tiles.ensure((size_t)nTiles * (ctx->obs_hw + 1) * ts));
    CU(ctx->d_labels.ensure((size_t)nTiles * ts));
    CU(tpgx_launch_pack(ctx->obs_dev, nullptr, ctx->nb_obs, ctx->nb_obs, ctx->obs_hw, ts, ctx->d_tiles.as<uint8_t>(),
                        ctx->labels_dev, ctx->d_labels.as<uint8_t>(), ctx->stream));
    tpgx_status st = sobel_planes(ctx, nTiles, ts);
    if (st != TPGX_OK) return st;
    ctx->tiles_identity_valid = true;
    ctx->tiles_ts = ts;
    return TPGX_OK;
}

static tpgx_status set_observations_host(tpgx_ctx* ctx, int32_t source, int32_t store, const void* data, const uint8_t* labels, int64_t count, bool wait_for_copies) {
    tpgx_status st = set_obs_common(ctx, source, store, count);
    if (st != TPGX_OK) return st;
    if (!data) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_observations: null data");
    cudaSetDevice(ctx->cfg.device);
    CU(wait_for_observations(ctx));
    /* enqueue-only (pinned) uploads run on obs_stream, after whatever the main stream still does with the old tiles */
    struct StreamSwap {
        tpgx_ctx* c; cudaStream_t main; bool on;
        ~StreamSwap() { if (on) c->stream = main; }
    } swap{ctx, ctx->stream, !wait_for_copies};
    if (swap.on) {
        CU(cudaEventRecord(ctx->event(60), ctx->stream));
        CU(cudaStreamWaitEvent(ctx->obs_stream, ctx->event(60), 0));
        ctx->stream = ctx->obs_stream;
    }
    if ((st = upload(ctx, ctx->d_obs_raw, data, (size_t)count * ctx->obs_hw, false)) != TPGX_OK) return st;
    ctx->obs_dev = ctx->d_obs_raw.as<uint8_t>();
    ctx->labels_dev = nullptr;
    if (labels) {
        if ((st = upload(ctx, ctx->d_labels_raw, labels, (size_t)count, false)) != TPGX_OK) return st;
        ctx->labels_dev = ctx->d_labels_raw.as<uint8_t>();
    }
    /* the caller may free its buffers on return: wait for the host -> device copies only; the re-tiling and
       Sobel-plane kernels keep running behind the host (later work is ordered on the same stream) */
    cudaEvent_t copied = ctx->event(61);
    CU(cudaEventRecord(copied, ctx->stream));
    if ((st = pack_identity(ctx)) != TPGX_OK) return st;
    if (wait_for_copies) CU(cudaEventSynchronize(copied));
    if (swap.on) {
        CU(cudaEventRecord(ctx->obs_ready, ctx->obs_stream));
        ctx->obs_pending = true;
    }
    return TPGX_OK;
}

tpgx_status tpgx_set_observations(tpgx_ctx* ctx, int32_t source, int32_t store, const void* data, const uint8_t* labels, int64_t count) {
    return set_observations_host(ctx, source, store, data, labels, count, true);
}

tpgx_status tpgx_set_observations_pinned(tpgx_ctx* ctx, int32_t source, int32_t store, const void* data, const uint8_t* labels, int64_t count) {
    if (ctx && data) { /* must really be page-locked: an asynchronous copy from pageable memory is staged and may read it late */
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, data) != cudaSuccess || at.type != cudaMemoryTypeHost) {
            cudaGetLastError();
            return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_observations_pinned: data is not page-locked host memory");
        }
    }
    return set_observations_host(ctx, source, store, data, labels, count, false);
}

tpgx_status tpgx_set_observations_device(tpgx_ctx* ctx, int32_t source, int32_t store, const void* d_data, const uint8_t* d_labels, int64_t count) {
    tpgx_status st = set_obs_common(ctx, source, store, count);
    if (st != TPGX_OK) return st;
    if (!d_data) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_set_observations_device: null data");
    cudaSetDevice(ctx->cfg.device);
    CU(wait_for_observations(ctx));
    ctx->obs_dev = static_cast<const uint8_t*>(d_data);
    ctx->labels_dev = d_labels;
    if ((st = pack_identity(ctx)) != TPGX_OK) return st;
    CU(cudaStreamSynchronize(ctx->stream));
    return TPGX_OK;
}

tpgx_status tpgx_evaluate_classification(tpgx_ctx* ctx, const tpgx_classif_request* req, tpgx_classif_result* out) {
    if (!ctx || !req || !out) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_classification: null argument");
    cudaSetDevice(ctx->cfg.device);
    const int nRq = req->root_end - req->root_begin;
    if (nRq > 0 && req->nb_classes > 0) CU(ctx->d_records.ensure((size_t)nRq * (req->nb_classes + 1) * 8));
    ClassifRun run;
    Trace tr("evaluate_classification");
    tpgx_status st = enqueue_classification(ctx, req, ctx->d_records.as<double>(), out->action != nullptr, out->bid != nullptr, &run, true);
    if (st != TPGX_OK) return st;
    tr.lap("enqueue (plan + launches)");
    if ((st = finish_stream(ctx)) != TPGX_OK) return st;
    tr.lap("device");
    const int nR = run.nR, C = run.C;
    if (out->score || out->class_f1) {
        std::vector<double> rec((size_t)nR * (C + 1));
        CU(cudaMemcpy(rec.data(), ctx->d_records.p, rec.size() * 8, cudaMemcpyDeviceToHost));
        for (int r = 0; r < nR; r++) {
            if (out->score) out->score[r] = rec[(size_t)r * (C + 1)];
            if (out->class_f1) memcpy(out->class_f1 + (size_t)r * C, &rec[(size_t)r * (C + 1) + 1], (size_t)C * 8);
        }
    }
    if (out->confusion) CU(cudaMemcpy(out->confusion, ctx->d_conf.p, (size_t)nR * C * C * 8, cudaMemcpyDeviceToHost));
    if (out->action) CU(cudaMemcpy(out->action, ctx->d_action.p, (size_t)nR * run.nS, cudaMemcpyDeviceToHost));
    if (out->bid) {
        /* rows are in shard order of first appearance; scatter to program ids. Unreached programs get NaN. */
        const int64_t stride = run.nTiles * run.ts;
        std::vector<double> tmp((size_t)ctx->plan.nb_active * stride);
        CU(cudaMemcpy(tmp.data(), ctx->d_bid.p, tmp.size() * 8, cudaMemcpyDeviceToHost));
        const double qnan = __builtin_nan("");
        for (size_t i = 0; i < (size_t)ctx->flat.nb_programs * run.nS; i++) out->bid[i] = qnan;
        for (int row = 0; row < ctx->plan.nb_active; row++)
            memcpy(out->bid + (size_t)ctx->h_active[row] * run.nS, tmp.data() + (size_t)row * stride, (size_t)run.nS * 8);
    }
    if (out->counters) {
        unsigned long long c[4];
        CU(cudaMemcpy(c, ctx->d_counters.p, 32, cudaMemcpyDeviceToHost));
        memset(out->counters, 0, sizeof(*out->counters));
        out->counters->evaluations = c[0];
        out->counters->team_visits = c[1];
        out->counters->program_executions = c[2];
        out->counters->lines_executed = c[3];
        out->counters->device_lines = (uint64_t)ctx->plan.active_lines * (uint64_t)run.nS;
        out->counters->device_program_executions = (uint64_t)ctx->plan.nb_active * (uint64_t)run.nS;
    }
    tr.lap("read-back");
    if (out->device_ms) {
        float k1 = 0, k2 = 0, k3 = 0, t;
        for (int ps = 0; ps < run.passes; ps++) {
            cudaEventElapsedTime(&t, ctx->event(3 * ps), ctx->event(3 * ps + 1)); k1 += t;
            cudaEventElapsedTime(&t, ctx->event(3 * ps + 1), ctx->event(3 * ps + 2)); k2 += t;
        }
        cudaEventElapsedTime(&k3, ctx->event(3 * run.passes - 1), ctx->event(3 * run.passes));
        out->device_ms[0] = k1; out->device_ms[1] = k2; out->device_ms[2] = k3;
    }
    return TPGX_OK;
}

tpgx_status tpgx_evaluate_classification_device(tpgx_ctx* ctx, const tpgx_classif_request* req, double* d_records, void* stream) {
    if (!ctx || !req || !d_records) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_classification_device: null argument");
    cudaSetDevice(ctx->cfg.device);
    /* order our stream after the caller's, run, then make the caller's stream wait for us */
    cudaStream_t user = static_cast<cudaStream_t>(stream);
    cudaEvent_t ev_in = ctx->event(62), ev_out = ctx->event(63);
    CU(cudaEventRecord(ev_in, user));
    CU(cudaStreamWaitEvent(ctx->stream, ev_in, 0));
    tpgx_status st = enqueue_classification(ctx, req, d_records, false, false, nullptr, false);
    if (st != TPGX_OK) return st;
    CU(cudaEventRecord(ev_out, ctx->stream));
    CU(cudaStreamWaitEvent(user, ev_out, 0));
    return TPGX_OK;
}

tpgx_status tpgx_execute_programs(tpgx_ctx* ctx, int64_t count, const double* const* obs_f64, const int32_t* const* obs_i32, double* bid) {
    if (!ctx || count < 1 || !bid) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_execute_programs: bad argument");
    if (!ctx->has_graph) return fail(ctx, TPGX_ERR_STATE, "tpgx_execute_programs: no graph set");
    cudaSetDevice(ctx->cfg.device);
    const double* f[2] = {nullptr, nullptr};
    const int32_t* iv[2] = {nullptr, nullptr};
    int space[2] = {0, 0}, width[2] = {0, 0};
    tpgx_status st;
    for (size_t k = 0; k < ctx->src.size(); k++) {
        space[k] = ctx->src[k].height * ctx->src[k].width;
        width[k] = ctx->src[k].width;
        if (ctx->src[k].elem == TPGX_ELEM_F64) {
            if (!obs_f64 || !obs_f64[k]) return fail(ctx, TPGX_ERR_BAD_ARG, "missing f64 observations");
            if ((st = upload(ctx, ctx->d_scratch[k], obs_f64[k], (size_t)count * space[k] * 8)) != TPGX_OK) return st;
            f[k] = ctx->d_scratch[k].as<double>();
        } else {
            if (!obs_i32 || !obs_i32[k]) return fail(ctx, TPGX_ERR_BAD_ARG, "missing i32 observations");
            if ((st = upload(ctx, ctx->d_scratch[k], obs_i32[k], (size_t)count * space[k] * 4)) != TPGX_OK) return st;
            iv[k] = ctx->d_scratch[k].as<int32_t>();
        }
    }
    const int P = ctx->flat.nb_programs;
    CU(ctx->d_scratch[2].ensure((size_t)P * count * 8));
    CU(tpgx_launch_exec_generic(ctx->d_code.as<uint32_t>(), ctx->d_prog_off.as<int32_t>(), ctx->d_consts.as<double>(),
                                ctx->cfg.nb_constants, P, count, f[0], f[1], iv[0], iv[1], space[0], space[1], width[0], width[1],
                                ctx->d_scratch[2].as<double>(), ctx->stream));
    CU(cudaMemcpyAsync(bid, ctx->d_scratch[2].p, (size_t)P * count * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TPGX_OK;
}

/* archive pass of the episode kernel: the recorded executions to resolve, per episode, and where the records go */
struct EpHitPass {
    std::vector<int32_t> hit_offset, hit_exec, hit_slot;
    int nb_records = 0, obs_len = 0;
    std::vector<int32_t> rec_program, rec_step;
    std::vector<double> rec_bid, rec_obs;
};

static tpgx_status episodes_impl(tpgx_ctx* ctx, const tpgx_episode_request* req, tpgx_episode_result* out,
                                 std::vector<int32_t>* progs_out, EpHitPass* hp) {
    if (!ctx || !req || !out) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_episodes: null argument");
    if (!ctx->has_graph) return fail(ctx, TPGX_ERR_STATE, "tpgx_evaluate_episodes: no graph set");
    cudaSetDevice(ctx->cfg.device);
    const int K = req->roots_per_job, NI = req->nb_iterations, J = req->nb_jobs;
    if (K < 1 || K > 2 || NI < 1 || J < 1 || req->max_actions < 1 || !req->seeds || !req->job_roots)
        return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_evaluate_episodes: bad request");
    for (int i = 0; i < J * K; i++)
        if (req->job_roots[i] < 0 || req->job_roots[i] >= ctx->flat.nb_roots) return fail(ctx, TPGX_ERR_BAD_ARG, "job root out of range");
    /* expected data-source shapes per environment (getDataSources() of the app) */
    const std::vector<tpgx_source_desc>& s = ctx->src;
    bool ok = false;
    switch (req->env) {
    case TPGX_ENV_GRIDWORLD: case TPGX_ENV_PENDULUM:
        ok = s.size() == 1 && s[0].elem == TPGX_ELEM_F64 && !s[0].is_2d && s[0].width == 2; break;
    case TPGX_ENV_TICTACTOE:
        ok = s.size() == 1 && s[0].elem == TPGX_ELEM_F64 && !s[0].is_2d && s[0].width == 9; break;
    case TPGX_ENV_STICKGAME:
        ok = s.size() == 2 && s[0].elem == TPGX_ELEM_I32 && s[0].width == 4 && s[1].elem == TPGX_ELEM_I32 && s[1].width == 1; break;
    default: return fail(ctx, TPGX_ERR_BAD_ARG, "unknown environment kind");
    }
    if (!ok) return fail(ctx, TPGX_ERR_BAD_ARG, "data sources do not match the environment's getDataSources()");
    if (req->env == TPGX_ENV_PENDULUM && (req->nb_pendulum_actions < 1 || req->nb_pendulum_actions > 16 || !req->pendulum_actions))
        return fail(ctx, TPGX_ERR_BAD_ARG, "pendulum needs its availableActions list (<= 16 entries)");

    tpgx_episode_params ep{};
    ep.env = req->env; ep.mode = req->mode; ep.nb_iterations = NI; ep.max_actions = req->max_actions;
    ep.nb_jobs = J; ep.roots_per_job = K; ep.second_player_random = req->second_player_random;
    ep.trace_steps = out->trace ? req->trace_steps : 0;
    ep.tie_break = ctx->cfg.tie_break;
    ep.nb_constants = ctx->cfg.nb_constants;
    ep.nb_pendulum_actions = req->nb_pendulum_actions;
    for (int i = 0; i < req->nb_pendulum_actions && i < 16; i++) ep.pendulum_actions[i] = req->pendulum_actions[i];
    {   /* lanes per episode: a team's edges are spread over a power-of-two lane group.  Several episodes share a warp
           only when there are enough of them to keep ~4 warps on every scheduler anyway: the kernel is bound by the
           latency of each episode's dependent steps, and a packed warp runs its groups' divergent programs one after
           the other (measured: stick game 100 x 100 episodes 1.21 -> 0.80 ms packed by 4; pendulum 500 x 5 episodes
           49 -> 57 ms, so that one stays at one episode per warp) */
        int deg = 1;
        for (size_t t = 0; t + 1 < ctx->flat.team_edge_offset.size(); t++) deg = std::max(deg, ctx->flat.team_edge_offset[t + 1] - ctx->flat.team_edge_offset[t]);
        int w = 1;
        while (w < deg && w < 32) w <<= 1;
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->cfg.device);
        int g = std::min(4, 32 / w);
        while (g > 1 && (long long)J * NI / g < 16ll * sms) g >>= 1;
        if (const char* v = getenv("TPGX_EP_GROUPS")) { const int f = atoi(v); if (f >= 1 && f <= 4 && f * w <= 32) g = f; }
        ep.group_width = w;
        ep.groups_per_warp = g;
        ep.cta_warps = std::min(deg, 16);
        if (const char* v = getenv("TPGX_EP_CTA_WARPS")) { const int f = atoi(v); if (f >= 0 && f <= 16) ep.cta_warps = f; }
    }
    /* RNG streams: every root sees the same seed per iteration ([UP] evaluateJob), so the raw
       mt19937_64 outputs are generated once per iteration on the host and shared by all jobs */
    std::vector<uint64_t> raw((size_t)NI * TPGX_RNG_TABLE);
    for (int it = 0; it < NI; it++)
        tpgx_rng_fill_raw(req->rng_seeds ? req->rng_seeds[it] : tpgx_env_seed(req->seeds[it], req->mode), TPGX_RNG_TABLE, raw.data() + (size_t)it * TPGX_RNG_TABLE);
    tpgx_status st;
    DevBuf* sc = ctx->d_scratch;
    const size_t nE = (size_t)J * NI;
    /* ONE host -> device transfer for the per-call inputs and ONE device -> host transfer for everything the call returns
       (a generation of the training loop makes two of these calls; a dozen small synchronous copies each used to cost more
       than the kernels): both sides are laid out as blocks of 8-byte-aligned sections */
    struct Section { size_t off = 0, bytes = 0; };
    auto carve = [](size_t& total, size_t bytes) { Section s; s.off = total; s.bytes = bytes; total += (bytes + 15) & ~(size_t)15; return s; };
    size_t in_total = 0;
    const Section i_raw = carve(in_total, raw.size() * 8), i_teams = carve(in_total, (size_t)J * K * 4);
    Section i_hoff, i_hexec, i_hslot;
    if (hp) { i_hoff = carve(in_total, hp->hit_offset.size() * 4); i_hexec = carve(in_total, std::max<size_t>(1, hp->hit_exec.size()) * 4); i_hslot = carve(in_total, std::max<size_t>(1, hp->hit_slot.size()) * 4); }
    std::vector<uint8_t>& hin = ctx->h_stage_in;
    hin.resize(in_total);
    memcpy(hin.data() + i_raw.off, raw.data(), i_raw.bytes);
    int32_t* job_teams = reinterpret_cast<int32_t*>(hin.data() + i_teams.off);
    for (int i = 0; i < J * K; i++) job_teams[i] = ctx->flat.root_team[req->job_roots[i]];
    if (hp) {
        memcpy(hin.data() + i_hoff.off, hp->hit_offset.data(), hp->hit_offset.size() * 4);
        memcpy(hin.data() + i_hexec.off, hp->hit_exec.data(), hp->hit_exec.size() * 4);
        memcpy(hin.data() + i_hslot.off, hp->hit_slot.data(), hp->hit_slot.size() * 4);
    }
    if ((st = upload(ctx, sc[0], hin.data(), in_total)) != TPGX_OK) return st;
    const size_t nr = hp ? (size_t)std::max(1, hp->nb_records) : 0;
    size_t out_total = 0;
    const Section o_status = carve(out_total, 16), o_counters = carve(out_total, 64), o_job = carve(out_total, (size_t)J * K * 8),
                  o_escore = carve(out_total, nE * K * 8), o_final = carve(out_total, nE * 4 * 8), o_actions = carve(out_total, nE * 4),
                  o_progs = carve(out_total, progs_out ? nE * 4 : 0), o_trace = carve(out_total, nE * (size_t)ep.trace_steps),
                  o_rbid = carve(out_total, nr * 8), o_robs = carve(out_total, hp ? nr * (size_t)std::max(1, hp->obs_len) * 8 : 0),
                  o_rprog = carve(out_total, nr * 4), o_rstep = carve(out_total, nr * 4);
    CU(sc[2].ensure(out_total));
    uint8_t* dout = sc[2].as<uint8_t>();
    const uint8_t* din = sc[0].as<uint8_t>();
    CU(cudaMemsetAsync(dout, 0, o_job.off, ctx->stream)); /* status + counters */
    ep.rng_raw = reinterpret_cast<const uint64_t*>(din + i_raw.off);
    ep.job_teams = reinterpret_cast<const int32_t*>(din + i_teams.off);
    ep.code = ctx->d_code.as<uint32_t>();
    ep.prog_offset = ctx->d_prog_off.as<int32_t>();
    ep.constants = ctx->d_consts.as<double>();
    ep.team_edge_offset = ctx->d_team_off.as<int32_t>();
    ep.edge_program = ctx->d_edge_prog.as<int32_t>();
    ep.edge_destination = ctx->d_edge_dst.as<int32_t>();
    ep.episode_score = reinterpret_cast<double*>(dout + o_escore.off);
    ep.episode_actions = reinterpret_cast<int32_t*>(dout + o_actions.off);
    ep.final_state = reinterpret_cast<double*>(dout + o_final.off);
    ep.trace = reinterpret_cast<int8_t*>(dout + o_trace.off);
    ep.job_score = reinterpret_cast<double*>(dout + o_job.off);
    ep.counters = reinterpret_cast<unsigned long long*>(dout + o_counters.off);
    ep.status = reinterpret_cast<int*>(dout + o_status.off);
    if (progs_out) ep.episode_progs = reinterpret_cast<int32_t*>(dout + o_progs.off);
    if (hp) {
        ep.hit_offset = reinterpret_cast<const int32_t*>(din + i_hoff.off);
        ep.hit_exec = reinterpret_cast<const int32_t*>(din + i_hexec.off);
        ep.hit_slot = reinterpret_cast<const int32_t*>(din + i_hslot.off);
        ep.rec_program = reinterpret_cast<int32_t*>(dout + o_rprog.off);
        ep.rec_step = reinterpret_cast<int32_t*>(dout + o_rstep.off);
        ep.rec_bid = reinterpret_cast<double*>(dout + o_rbid.off);
        ep.rec_obs = hp->obs_len > 0 ? reinterpret_cast<double*>(dout + o_robs.off) : nullptr;
        ep.obs_len = hp->obs_len;
    }
    CU(cudaEventRecord(ctx->event(0), ctx->stream));
    CU(tpgx_launch_episodes(ep, ctx->stream));
    CU(cudaEventRecord(ctx->event(1), ctx->stream));
    std::vector<uint8_t>& hout = ctx->h_stage_out;
    hout.resize(out_total);
    CU(cudaMemcpyAsync(hout.data(), dout, out_total, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    int status = 0;
    memcpy(&status, hout.data() + o_status.off, 4);
    if ((st = check_device_status(ctx, status)) != TPGX_OK) return st;
    if (progs_out) {
        progs_out->resize(nE);
        memcpy(progs_out->data(), hout.data() + o_progs.off, nE * 4);
    }
    if (hp) {
        const size_t n = (size_t)hp->nb_records;
        hp->rec_program.resize(n); hp->rec_step.resize(n); hp->rec_bid.resize(n); hp->rec_obs.resize(n * (size_t)hp->obs_len);
        if (n) {
            memcpy(hp->rec_program.data(), hout.data() + o_rprog.off, n * 4);
            memcpy(hp->rec_step.data(), hout.data() + o_rstep.off, n * 4);
            memcpy(hp->rec_bid.data(), hout.data() + o_rbid.off, n * 8);
            if (hp->obs_len > 0) memcpy(hp->rec_obs.data(), hout.data() + o_robs.off, n * (size_t)hp->obs_len * 8);
        }
        return TPGX_OK; /* the archive pass writes records only */
    }
    if (out->score) memcpy(out->score, hout.data() + o_job.off, (size_t)J * K * 8);
    if (out->episode_score) memcpy(out->episode_score, hout.data() + o_escore.off, nE * K * 8);
    if (out->episode_actions) memcpy(out->episode_actions, hout.data() + o_actions.off, nE * 4);
    if (out->final_state) memcpy(out->final_state, hout.data() + o_final.off, nE * 4 * 8);
    if (out->trace && ep.trace_steps > 0) memcpy(out->trace, hout.data() + o_trace.off, nE * (size_t)ep.trace_steps);
    if (out->counters) {
        unsigned long long c[4];
        memcpy(c, hout.data() + o_counters.off, 32);
        memset(out->counters, 0, sizeof(*out->counters));
        out->counters->evaluations = c[0]; out->counters->team_visits = c[1];
        out->counters->program_executions = c[2]; out->counters->lines_executed = c[3];
        out->counters->device_lines = c[3]; out->counters->device_program_executions = c[2];
    }
    if (out->device_ms) cudaEventElapsedTime(out->device_ms, ctx->event(0), ctx->event(1));
    return TPGX_OK;
}

tpgx_status tpgx_evaluate_episodes(tpgx_ctx* ctx, const tpgx_episode_request* req, tpgx_episode_result* out) {
    return episodes_impl(ctx, req, out, nullptr, nullptr);
}

/* [UP] Archive::addRecording during a TRAINING evaluateAllRoots of a sequential environment.  Pass 1 is the evaluation
   itself and yields the number of program executions of every episode; the host replays every job's draw stream over
   the job's executions (iterations in order), keeps the recordings that survive the per-job and the merged ring, and
   pass 2 re-runs only the episodes that hold such an execution to resolve (program, step, bid, observation). */
tpgx_status tpgx_archive_episodes(tpgx_ctx* ctx, const tpgx_episode_request* req, const tpgx_archive_request* arch,
                                  tpgx_episode_result* eval_out, tpgx_episode_archive_result* out) {
    if (!ctx || !req || !arch || !out || !arch->archive_seed || arch->archive_size < 0 || !(arch->probability >= 0.0 && arch->probability <= 1.0))
        return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_archive_episodes: bad argument");
    if (arch->archive_size > 0 && (!out->program || !out->job || !out->iteration || !out->step || !out->bid))
        return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_archive_episodes: null output");
    if (out->observation && out->obs_len < 1) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_archive_episodes: obs_len missing");
    tpgx_episode_result scratch_out{};
    std::vector<int32_t> progs;
    Trace tr("archive_episodes");
    tpgx_status st = episodes_impl(ctx, req, eval_out ? eval_out : &scratch_out, &progs, nullptr);
    if (st != TPGX_OK) return st;
    tr.lap("evaluation pass");
    out->nb_records = 0;
    const int J = req->nb_jobs, NI = req->nb_iterations, cap = arch->archive_size;
    if (cap == 0 || arch->probability == 0.0) return TPGX_OK;
    /* per job: the job-local indices of its recorded executions (its own ring keeps the last `cap`) */
    std::vector<std::vector<int64_t>> job_hits((size_t)J);
    tpgx_parallel_blocks(J, 8, [&](int jb, int je, int) {
        for (int j = jb; j < je; j++) {
            int64_t N = 0;
            for (int it = 0; it < NI; it++) N += progs[(size_t)j * NI + it];
            std::vector<int64_t>& hits = job_hits[j];
            if (arch->probability == 1.0) {
                for (int64_t k = std::max<int64_t>(0, N - cap); k < N; k++) hits.push_back(k);
            } else {
                std::mt19937_64 eng(arch->archive_seed[j]);
                for (int64_t k = 0; k < N; k++) {
                    double canon = (double)eng() / 18446744073709551616.0;     /* libstdc++ uniform_real_distribution<double>(0, 1) */
                    if (canon >= 1.0) canon = 0.99999999999999988898;
                    if (arch->probability >= canon) hits.push_back(k);
                }
                if (hits.size() > (size_t)cap) hits.erase(hits.begin(), hits.end() - cap);
            }
        }
    });
    tr.lap("host replay of the draw streams");
    /* merged archive = the last `cap` recordings of the per-job archives concatenated in job order */
    int64_t total = 0;
    for (int j = 0; j < J; j++) total += (int64_t)job_hits[j].size();
    int64_t drop = std::max<int64_t>(0, total - cap);
    EpHitPass hp;
    hp.obs_len = out->observation ? out->obs_len : 0;
    hp.hit_offset.assign((size_t)J * NI + 1, 0);
    std::vector<int32_t> rec_job, rec_it;
    for (int j = 0; j < J; j++) {
        std::vector<int64_t>& hits = job_hits[j];
        size_t first = 0;
        if (drop > 0) { first = (size_t)std::min<int64_t>(drop, (int64_t)hits.size()); drop -= (int64_t)first; }
        int it = 0;
        int64_t it_begin = 0;
        for (size_t h = first; h < hits.size(); h++) {
            while (it < NI && hits[h] >= it_begin + progs[(size_t)j * NI + it]) { it_begin += progs[(size_t)j * NI + it]; it++; }
            if (it >= NI) return fail(ctx, TPGX_ERR_STATE, "tpgx_archive_episodes: internal execution-count mismatch");
            hp.hit_offset[(size_t)j * NI + it + 1]++;
            hp.hit_exec.push_back((int32_t)(hits[h] - it_begin));
            hp.hit_slot.push_back((int32_t)rec_job.size());
            rec_job.push_back(j);
            rec_it.push_back(it);
        }
    }
    for (size_t i = 1; i < hp.hit_offset.size(); i++) hp.hit_offset[i] += hp.hit_offset[i - 1];
    hp.nb_records = (int)rec_job.size();
    if (hp.nb_records == 0) return TPGX_OK;
    tr.lap("hit lists");
    if ((st = episodes_impl(ctx, req, &scratch_out, nullptr, &hp)) != TPGX_OK) return st;
    tr.lap("resolution pass");
    out->nb_records = hp.nb_records;
    for (int i = 0; i < hp.nb_records; i++) {
        out->program[i] = hp.rec_program[i]; out->job[i] = rec_job[i]; out->iteration[i] = rec_it[i];
        out->step[i] = hp.rec_step[i]; out->bid[i] = hp.rec_bid[i];
    }
    if (out->observation) memcpy(out->observation, hp.rec_obs.data(), hp.rec_obs.size() * 8);
    return TPGX_OK;
}

/* ---- GPU backend of the host generation loop (csrc/train.cpp) for the sequential environments ---- */
namespace {
struct EpBackend {
    tpgx_ctx* ctx;
    tpgx_episode_request req;
};
tpgx_status epbe_evaluate(void* user, const tpgx_graph* g, uint64_t generation, int32_t mode, const uint64_t* archive_seed, double probability,
                          int32_t archive_size, double* scores, int32_t /*nb_classes*/, double* /*class_scores*/, uint64_t* /*class_counts*/,
                          int32_t* rec_program, double* rec_bid, double* rec_obs, int32_t obs_len, int32_t* nb_records) {
    EpBackend* b = static_cast<EpBackend*>(user);
    tpgx_status st = tpgx_set_graph(b->ctx, g);
    if (st != TPGX_OK) return st;
    const int R = g->nb_roots, NI = b->req.nb_iterations;
    std::vector<int32_t> job_roots((size_t)R);
    for (int j = 0; j < R; j++) job_roots[j] = j;
    std::vector<uint64_t> seeds((size_t)NI);
    for (int it = 0; it < NI; it++) seeds[it] = 1000ull * generation + (uint64_t)it; /* stands for Hash(gen) ^ Hash(it) */
    tpgx_episode_request req = b->req;
    req.mode = mode;
    req.nb_jobs = R; req.roots_per_job = 1; req.job_roots = job_roots.data(); req.seeds = seeds.data(); req.trace_steps = 0;
    tpgx_episode_result ev{};
    ev.score = scores;
    *nb_records = 0;
    if (archive_size <= 0 || probability == 0.0 || mode != 0 /* [UP] only a TRAINING evaluation archives */)
        return tpgx_evaluate_episodes(b->ctx, &req, &ev);
    const size_t cap = (size_t)std::max(1, archive_size);
    std::vector<int32_t> job(cap), iteration(cap), step(cap);
    tpgx_archive_request ar{archive_seed, probability, archive_size, 0};
    tpgx_episode_archive_result out{rec_program, job.data(), iteration.data(), step.data(), rec_bid, rec_obs, obs_len, 0};
    st = tpgx_archive_episodes(b->ctx, &req, &ar, &ev, &out);
    *nb_records = out.nb_records;
    return st;
}
tpgx_status epbe_execute(void* user, const tpgx_graph* g, int64_t count, const double* obs, int32_t obs_len, double* bid) {
    EpBackend* b = static_cast<EpBackend*>(user);
    tpgx_status st = tpgx_set_graph(b->ctx, g);
    if (st != TPGX_OK) return st;
    const std::vector<tpgx_source_desc>& src = b->ctx->src;
    std::vector<std::vector<double>> f(src.size());
    std::vector<std::vector<int32_t>> iv(src.size());
    const double* pf[2] = {nullptr, nullptr};
    const int32_t* pi[2] = {nullptr, nullptr};
    int col = 0;
    for (size_t k = 0; k < src.size(); k++) {
        const int sp = src[k].height * src[k].width;
        if (col + sp > obs_len) return fail(b->ctx, TPGX_ERR_BAD_ARG, "observation rows shorter than the data sources");
        if (src[k].elem == TPGX_ELEM_F64) {
            f[k].resize((size_t)count * sp);
            for (int64_t i = 0; i < count; i++) for (int q = 0; q < sp; q++) f[k][(size_t)i * sp + q] = obs[(size_t)i * obs_len + col + q];
            pf[k] = f[k].data();
        } else {
            iv[k].resize((size_t)count * sp);
            for (int64_t i = 0; i < count; i++) for (int q = 0; q < sp; q++) iv[k][(size_t)i * sp + q] = (int32_t)obs[(size_t)i * obs_len + col + q];
            pi[k] = iv[k].data();
        }
        col += sp;
    }
    return tpgx_execute_programs(b->ctx, count, pf, pi, bid);
}
} // namespace

/* ---- GPU backend of the host generation loop for the classification environment (the mnist app) ---- */
namespace {
struct ClBackend {
    tpgx_ctx* ctx;
    const uint8_t* images; /* host copy of what was uploaded: [nb_images][height * width], borrowed */
    int64_t nb_images;
    int32_t nb_classes;
    int64_t actions;
};
tpgx_status clbe_evaluate(void* user, const tpgx_graph* g, uint64_t generation, int32_t mode, const uint64_t* archive_seed, double probability,
                          int32_t archive_size, double* scores, int32_t nb_classes, double* class_scores, uint64_t* class_counts,
                          int32_t* rec_program, double* rec_bid, double* rec_obs, int32_t obs_len, int32_t* nb_records) {
    ClBackend* b = static_cast<ClBackend*>(user);
    tpgx_status st = tpgx_set_graph(b->ctx, g);
    if (st != TPGX_OK) return st;
    if (nb_classes != 0 && nb_classes != b->nb_classes) return fail(b->ctx, TPGX_ERR_BAD_ARG, "trainer and backend disagree on the number of classes");
    /* [REF] mnist/src/mnist.cpp:58-69 + :8-34: the evaluation classifies the images drawn from the episode seed; TRAINING and
       VALIDATION both draw from the training set, the mode does not enter the draw (rng.setSeed(seed): mnist.cpp:64) */
    std::vector<int32_t> idx((size_t)b->actions);
    if ((st = tpgx_mnist_episode_indices(1000ull * generation, mode, b->nb_images, b->actions, idx.data())) != TPGX_OK) return st;
    const int R = g->nb_roots, C = b->nb_classes;
    tpgx_classif_request req{};
    req.nb_samples = b->actions; req.sample_index = idx.data(); req.nb_classes = C; req.root_begin = 0; req.root_end = R;
    tpgx_classif_result ev{};
    ev.score = scores;
    std::vector<uint64_t> conf;
    if (class_scores) ev.class_f1 = class_scores;
    if (class_counts) { conf.resize((size_t)R * C * C); ev.confusion = conf.data(); }
    if ((st = tpgx_evaluate_classification(b->ctx, &req, &ev)) != TPGX_OK) return st;
    if (class_counts) /* [UP] nbEvaluationPerClass: the samples of each class the root has classified */
        for (int r = 0; r < R; r++)
            for (int c = 0; c < C; c++) {
                uint64_t n = 0;
                for (int a = 0; a < C; a++) n += conf[((size_t)r * C + c) * C + a];
                class_counts[(size_t)r * C + c] = n;
            }
    *nb_records = 0;
    if (archive_size <= 0 || probability == 0.0 || mode != 0) return TPGX_OK;
    const int px = b->ctx->obs_hw;
    if (obs_len < px) return fail(b->ctx, TPGX_ERR_BAD_ARG, "observation rows shorter than an image");
    std::vector<int64_t> sample((size_t)archive_size);
    tpgx_archive_request ar{archive_seed, probability, archive_size, 0};
    tpgx_archive_result res{rec_program, sample.data(), rec_bid, 0, 0};
    if ((st = tpgx_archive_classification(b->ctx, &req, &ar, &res)) != TPGX_OK) return st;
    for (int i = 0; i < res.nb_records; i++) { /* the recorded observation: the image itself, as the doubles the programs see */
        const uint8_t* im = b->images + (size_t)sample[i] * px;
        double* row = rec_obs + (size_t)i * obs_len;
        for (int q = 0; q < px; q++) row[q] = (double)im[q];
        for (int q = px; q < obs_len; q++) row[q] = 0.0;
    }
    *nb_records = res.nb_records;
    return TPGX_OK;
}
tpgx_status clbe_execute(void* user, const tpgx_graph* g, int64_t count, const double* obs, int32_t obs_len, double* bid) {
    ClBackend* b = static_cast<ClBackend*>(user);
    tpgx_status st = tpgx_set_graph(b->ctx, g);
    if (st != TPGX_OK) return st;
    if (obs_len != b->ctx->obs_hw) return fail(b->ctx, TPGX_ERR_BAD_ARG, "observation rows must be one image long");
    const double* pf[2] = {obs, nullptr};
    const int32_t* pi[2] = {nullptr, nullptr};
    return tpgx_execute_programs(b->ctx, count, pf, pi, bid);
}
} // namespace

tpgx_status tpgx_train_backend_classification(tpgx_ctx* ctx, const uint8_t* images, int64_t nb_images, int32_t nb_classes,
                                              int64_t actions_per_evaluation, tpgx_train_backend* out) {
    if (!ctx || !images || !out || nb_images < 1 || nb_classes < 2 || actions_per_evaluation < 1)
        return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_train_backend_classification: bad argument");
    if (!ctx->obs_dev || ctx->nb_obs != nb_images) return fail(ctx, TPGX_ERR_STATE, "tpgx_train_backend_classification: upload the same images with tpgx_set_observations first");
    ClBackend* b = new ClBackend{ctx, images, nb_images, nb_classes, actions_per_evaluation};
    out->user = b;
    out->evaluate = clbe_evaluate;
    out->execute = clbe_execute;
    return TPGX_OK;
}

tpgx_status tpgx_train_backend_episodes(tpgx_ctx* ctx, const tpgx_episode_request* req, tpgx_train_backend* out) {
    if (!ctx || !req || !out) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_train_backend_episodes: null argument");
    if (req->nb_iterations < 1 || req->max_actions < 1) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_train_backend_episodes: bad request template");
    EpBackend* b = new EpBackend{ctx, *req};
    out->user = b;
    out->evaluate = epbe_evaluate;
    out->execute = epbe_execute;
    return TPGX_OK;
}
void tpgx_train_backend_release(tpgx_train_backend* backend) {
    if (!backend || !backend->user) return;
    if (backend->evaluate == epbe_evaluate) delete static_cast<EpBackend*>(backend->user);
    else if (backend->evaluate == clbe_evaluate) delete static_cast<ClBackend*>(backend->user);
    else return;
    backend->user = nullptr;
}

tpgx_status tpgx_mnist_episode_indices(uint64_t seed, int32_t mode, int64_t dataset_size, int64_t nb_actions, int32_t* out) {
    if (!out || dataset_size < 1 || dataset_size > 0x7fffffff || nb_actions < 0 || mode < 0 || mode > 2) return TPGX_ERR_BAD_ARG;
    if (mode == 2) { /* TESTING: currentIndex = (currentIndex + 1) % size, starting from -1 */
        for (int64_t i = 0; i < nb_actions; i++) out[i] = (int32_t)(i % dataset_size);
        return TPGX_OK;
    }
    /* libstdc++ uniform_int_distribution<uint64_t>(0, size-1) on a 64-bit engine: Lemire's nearly
       divisionless method on the 64x64->128 product (the restatement of csrc/tpgx_rng.h, unbounded stream) */
    std::mt19937_64 eng(seed);
    const uint64_t range = (uint64_t)dataset_size;
    for (int64_t i = 0; i < nb_actions; i++) {
        unsigned __int128 prod = (unsigned __int128)eng() * range;
        uint64_t low = (uint64_t)prod;
        if (low < range) {
            const uint64_t threshold = (0 - range) % range;
            while (low < threshold) { prod = (unsigned __int128)eng() * range; low = (uint64_t)prod; }
        }
        out[i] = (int32_t)(uint64_t)(prod >> 64);
    }
    return TPGX_OK;
}

tpgx_status tpgx_archive_classification(tpgx_ctx* ctx, const tpgx_classif_request* req, const tpgx_archive_request* arch,
                                        tpgx_archive_result* out) {
    if (!ctx || !req || !arch || !out || !arch->archive_seed || arch->archive_size < 0 || !(arch->probability >= 0.0))
        return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_archive_classification: bad argument");
    if (arch->archive_size > 0 && (!out->program || !out->sample || !out->bid)) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_archive_classification: null output");
    if (!ctx->has_graph) return fail(ctx, TPGX_ERR_STATE, "tpgx_archive_classification: no graph set");
    cudaSetDevice(ctx->cfg.device);
    const int R0 = req->root_begin, R1 = req->root_end, cap = arch->archive_size;
    if (R0 < 0 || R1 > ctx->flat.nb_roots || R0 >= R1) return fail(ctx, TPGX_ERR_BAD_ARG, "root range out of bounds");
    out->nb_records = 0;
    if (cap == 0 || arch->probability == 0.0) {
        /* p == 0: `0 >= draw` only for a draw of exactly 0.0 (1 in 2^64) — still honoured below unless the archive is empty by size */
        if (cap == 0) return TPGX_OK;
    }
    struct Rec { int32_t program; int64_t sample; double bid; };
    std::vector<Rec> tail;                  /* collected from the LAST job backwards; reversed at the end */
    int remaining = cap;
    const int BATCH = 16;
    tpgx_status st;
    if ((st = upload(ctx, ctx->d_scratch[7], ctx->flat.edge_program.data(), ctx->flat.edge_program.size() * 4)) != TPGX_OK) return st;
    for (int be = R1; be > R0 && remaining > 0; be -= BATCH) {
        const int bb = std::max(R0, be - BATCH), nb = be - bb;
        /* bid matrix of the programs these roots reach, all samples (bid mode, K1 only) */
        tpgx_classif_request sub = *req;
        sub.root_begin = bb; sub.root_end = be;
        ClassifRun run;
        if ((st = enqueue_classification(ctx, &sub, nullptr, false, true, &run, false, true)) != TPGX_OK) return st;
        const int nS = (int)run.nS;
        tpgx_arch_params ap{};
        ap.bid = ctx->d_bid.as<double>();
        ap.row_stride = (long long)run.nTiles * run.ts;
        ap.nb_samples = nS;
        ap.team_edge_offset = ctx->d_team_off.as<int32_t>();
        ap.edge_row = ctx->d_edge_row.as<int32_t>();
        ap.edge_destination = ctx->d_edge_dst.as<int32_t>();
        ap.edge_program = ctx->d_scratch[7].as<int32_t>();
        ap.roots = ctx->d_roots.as<int32_t>();
        ap.nb_roots = nb;
        ap.tie_break = ctx->cfg.tie_break;
        ap.sample_obs = req->sample_index ? ctx->d_index.as<int32_t>() : nullptr;
        ap.status = ctx->d_status.as<int>();
        CU(ctx->d_scratch[3].ensure((size_t)nb * nS * 8));
        ap.counts = ctx->d_scratch[3].as<unsigned long long>();
        CU(tpgx_launch_arch_count(ap, ctx->stream));
        size_t temp_bytes = 0;
        CU(tpgx_arch_scan(ap, nullptr, &temp_bytes, ctx->stream));
        CU(ctx->d_scratch[4].ensure(temp_bytes + 16));
        CU(tpgx_arch_scan(ap, ctx->d_scratch[4].p, &temp_bytes, ctx->stream));
        std::vector<unsigned long long> totals(nb);
        for (int r = 0; r < nb; r++)
            CU(cudaMemcpyAsync(&totals[r], ap.counts + (size_t)r * nS + (nS - 1), 8, cudaMemcpyDeviceToHost, ctx->stream));
        int status = 0;
        CU(cudaMemcpyAsync(&status, ctx->d_status.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if ((st = check_device_status(ctx, status)) != TPGX_OK) return st;
        /* the jobs of the batch, last first: replay the job's draw stream, keep its last recordings that still fit */
        std::vector<unsigned long long> hit_k;
        std::vector<int32_t> hit_root;
        for (int r = nb - 1; r >= 0 && remaining > 0; r--) {
            const unsigned long long N = totals[r];
            std::vector<unsigned long long> hits;
            if (arch->probability == 1.0) {
                const unsigned long long take = std::min<unsigned long long>(N, (unsigned long long)remaining);
                for (unsigned long long k = N - take; k < N; k++) hits.push_back(k);
            } else {
                std::mt19937_64 eng(arch->archive_seed[bb + r - R0]);
                for (unsigned long long k = 0; k < N; k++) {
                    double canon = (double)eng() / 18446744073709551616.0;     /* libstdc++ uniform_real_distribution<double>(0, 1) */
                    if (canon >= 1.0) canon = 0.99999999999999988898;
                    if (arch->probability >= canon) hits.push_back(k);
                }
                const size_t keep = std::min<size_t>(hits.size(), (size_t)std::min(remaining, cap));
                hits.erase(hits.begin(), hits.end() - keep);
            }
            remaining -= (int)hits.size();
            /* batch hit list is assembled job-descending; inside a job ascending */
            hit_k.insert(hit_k.begin(), hits.begin(), hits.end());
            hit_root.insert(hit_root.begin(), hits.size(), r);
        }
        const int nh = (int)hit_k.size();
        if (nh > 0) {
            if ((st = upload(ctx, ctx->d_scratch[5], hit_k.data(), (size_t)nh * 8)) != TPGX_OK) return st;
            if ((st = upload(ctx, ctx->d_scratch[6], hit_root.data(), (size_t)nh * 4)) != TPGX_OK) return st;
            CU(ctx->d_scratch[0].ensure((size_t)nh * 4));
            CU(ctx->d_scratch[1].ensure((size_t)nh * 8));
            CU(ctx->d_scratch[2].ensure((size_t)nh * 8));
            ap.hit_k = ctx->d_scratch[5].as<unsigned long long>();
            ap.hit_root = ctx->d_scratch[6].as<int32_t>();
            ap.nb_hits = nh;
            ap.out_program = ctx->d_scratch[0].as<int32_t>();
            ap.out_sample = ctx->d_scratch[1].as<long long>();
            ap.out_bid = ctx->d_scratch[2].as<double>();
            CU(tpgx_launch_arch_resolve(ap, ctx->stream));
            std::vector<int32_t> hp(nh);
            std::vector<long long> hs(nh);
            std::vector<double> hb(nh);
            CU(cudaMemcpyAsync(hp.data(), ap.out_program, (size_t)nh * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(hs.data(), ap.out_sample, (size_t)nh * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaMemcpyAsync(hb.data(), ap.out_bid, (size_t)nh * 8, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            for (int i = nh - 1; i >= 0; i--) tail.push_back(Rec{hp[i], (int64_t)hs[i], hb[i]});   /* newest first */
        }
    }
    out->nb_records = (int32_t)tail.size();
    for (size_t i = 0; i < tail.size(); i++) {
        const Rec& rc = tail[tail.size() - 1 - i];
        out->program[i] = rc.program; out->sample[i] = rc.sample; out->bid[i] = rc.bid;
    }
    return TPGX_OK;
}

tpgx_status tpgx_math_probe(tpgx_ctx* ctx, int32_t fn, int64_t n, const double* a, const double* b, double* out) {
    if (!ctx || n < 1 || !a || !out || fn < 0 || fn > TPGX_PROBE_DIV10_K) return fail(ctx, TPGX_ERR_BAD_ARG, "tpgx_math_probe: bad argument");
    cudaSetDevice(ctx->cfg.device);
    tpgx_status st;
    if ((st = upload(ctx, ctx->d_scratch[0], a, (size_t)n * 8)) != TPGX_OK) return st;
    if (b && (st = upload(ctx, ctx->d_scratch[1], b, (size_t)n * 8)) != TPGX_OK) return st;
    CU(ctx->d_scratch[2].ensure((size_t)n * 8));
    CU(tpgx_launch_math_probe(fn, n, ctx->d_scratch[0].as<double>(), b ? ctx->d_scratch[1].as<double>() : nullptr,
                              ctx->d_scratch[2].as<double>(), ctx->stream));
    CU(cudaMemcpyAsync(out, ctx->d_scratch[2].p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return TPGX_OK;
}

tpgx_status tpgx_measure_fp64_peak(tpgx_ctx* ctx, double* dadd_per_s, double* dfma_per_s) {
    if (!ctx) return TPGX_ERR_BAD_ARG;
    cudaSetDevice(ctx->cfg.device);
    CU(ctx->d_scratch[0].ensure(64));
    const int blocks = 148 * 8, iters = 20000;
    double best[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 4; rep++) {
            CU(cudaEventRecord(ctx->event(0), ctx->stream));
            CU(tpgx_launch_fp64_peak(mode, blocks, iters, ctx->d_scratch[0].as<double>(), ctx->stream));
            CU(cudaEventRecord(ctx->event(1), ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->event(0), ctx->event(1));
            /* thread-level FP64 instructions per second */
            const double ops = (double)blocks * 256.0 * iters * 16.0;
            if (rep > 0) best[mode] = std::max(best[mode], ops / (ms * 1e-3));
        }
    }
    if (dadd_per_s) *dadd_per_s = best[0];
    if (dfma_per_s) *dfma_per_s = best[1];
    return TPGX_OK;
}

} /* extern "C" */
