/*
 * tpgx_internal.h — device bytecode format and host-side context of libtpgx.
 *
 * Device line word (32 bit), produced by the flattener (flatten.cpp) from the reference's
 * Program::Line encoding with introns removed and every location already reduced modulo the
 * address space of (source, operand type) — the modulo of [UP] fetchCurrentOperands is static per
 * line, so the device never divides:
 *
 *   [ 4: 0] device opcode (tpgx_opcode)
 *   [ 7: 5] destination register
 *   [19: 8] operand A : kind(2) << 10 | loc(10)
 *   [31:20] operand B : kind(2) << 10 | loc(10)
 *
 *   kind 0 REG   loc < 8 : register loc;  loc == TPGX_LOC_ZERO : a register no earlier effective
 *                line wrote, i.e. still 0.0 from [UP] executeProgram's register reset — this
 *                removes the per-program register clear on the device
 *   kind 1 CONST loc : program constant index
 *   kind 2 SRC0  loc : element index in data source 0 (for a 3x3 window: index of its top-left
 *                element, row * width + col)
 *   kind 3 SRC1  loc : element index in data source 1
 */
#ifndef TPGX_INTERNAL_H
#define TPGX_INTERNAL_H

#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "../../include/tpgx.h"

#define TPGX_KIND_REG 0u
#define TPGX_KIND_CONST 1u
#define TPGX_KIND_SRC0 2u
#define TPGX_KIND_SRC1 3u
#define TPGX_LOC_ZERO 8u
#define TPGX_MAX_REGS 8
#define TPGX_MAX_CONSTS 8
#define TPGX_MAX_LOC 1024
#define TPGX_MAX_PATH 64 /* teams on one root->action path; longer paths raise GRAPH_INVALID */

static inline uint32_t tpgx_pack_line(uint32_t op, uint32_t dst, uint32_t ka, uint32_t la, uint32_t kb, uint32_t lb) {
    return (op & 31u) | ((dst & 7u) << 5) | (((ka << 10) | (la & 1023u)) << 8) | (((kb << 10) | (lb & 1023u)) << 20);
}

/* ---- K1 wide line (16 bytes), derived from the generic word by tpgx_k1_line for one tile shape -----
 * Everything the interpreter would otherwise shift, mask and multiply per line is precomputed:
 *   x  handler id (below): for the cheap instructions the operand kinds (register / pixel) are part
 *      of the id, so a line costs one dispatch and no kind tests
 *   y  operand A: byte offset from the lane's register base (slot * TS * 8), or — pixel — index of the
 *      lane-0 packed word (K pixels) in the tile (loc * 32); for a 3x3 window: index of the lane-0 double2
 *      of its window in a Sobel plane (wp * TS / 2).  Indices rather than byte offsets, so that a 64-bit
 *      global address is one IMAD.WIDE (index * element size + base)
 *   z  operand B: same, or the byte offset of the program constant (index * 8) for MULC
 *   w  destination register: byte offset from the lane's register base
 * TS = samples per tile (32 * K).  A program's stream entry is TPGX_K1_CQ quads holding its constants
 * followed by its lines; TPGX_K1_CAPL lines are staged in shared memory at a time.                  */
#define TPGX_K1_CQ 4
#define TPGX_K1_CAPL 28
/* handler id bit layout: [0] operand A is a pixel, [1] operand B is a pixel, [5:2] operation, [6] heavy class
 * (own operand handling / libm), so the interpreter dispatches with a few single-bit tests; [7] set by the stream
 * encoder on the last line of a block */
#define TPGX_K1H_HEAVY 64u
#define TPGX_K1H_LAST 128u /* last line of a staged block (TPGX_K1_CAPL lines) or of the program: ends K1's line loop */
enum { /* cheap class: (op << 2) | kinds */
    TPGX_K1C_SUB = 0, TPGX_K1C_ADD, TPGX_K1C_MUL, TPGX_K1C_MAX, TPGX_K1C_COND, TPGX_K1C_EQ0, TPGX_K1C_EQM1, TPGX_K1C_EQ1,
    TPGX_K1C_GE15, TPGX_K1C_PI
};
enum { /* heavy class: 64 | (op << 2) | kinds */
    TPGX_K1X_DIV = 0, TPGX_K1X_MULC, TPGX_K1X_EXP, TPGX_K1X_LOG, TPGX_K1X_SOBEL_MAGN, TPGX_K1X_SOBEL_DIR,
    TPGX_K1X_FMODG, TPGX_K1X_COS, TPGX_K1X_SIN, TPGX_K1X_TAN
};
/* operations outside the MNIST set ([REF] mnist/src/main.cpp:79-88): compiled into the "extended" kernel only */
#if defined(__CUDACC__)
#define TPGX_HOSTDEV __host__ __device__
#else
#define TPGX_HOSTDEV
#endif
TPGX_HOSTDEV static inline bool tpgx_k1_is_extended(uint32_t h) {
    const uint32_t op = (h >> 2) & 15u;
    return (h & TPGX_K1H_HEAVY) ? op >= TPGX_K1X_FMODG : op >= TPGX_K1C_COND;
}

struct tpgx_k1_quad { uint32_t x, y, z, w; };

TPGX_HOSTDEV static inline tpgx_k1_quad tpgx_k1_line(uint32_t word, uint32_t ts, uint32_t width, uint32_t hw) {
    const uint32_t op = word & 31u, dst = (word >> 5) & 7u;
    const uint32_t fa = (word >> 8) & 0xfffu, fb = word >> 20;
    /* a never-written TPG register (TPGX_LOC_ZERO) reads the all-zero pixel row hw that K0 appends to every tile */
    const uint32_t za = (fa >> 10) == TPGX_KIND_REG && (fa & 1023u) >= TPGX_LOC_ZERO;
    const uint32_t zb = (fb >> 10) == TPGX_KIND_REG && (fb & 1023u) >= TPGX_LOC_ZERO;
    const uint32_t pa = (fa >> 10) == TPGX_KIND_SRC0 || za, pb = (fb >> 10) == TPGX_KIND_SRC0 || zb;
    const uint32_t la = za ? hw : (fa & 1023u), lb = zb ? hw : (fb & 1023u);
    tpgx_k1_quad q;
    q.y = pa ? la * 32u : la * ts * 8u;
    q.z = pb ? lb * 32u : lb * ts * 8u;
    q.w = dst * ts * 8u;
    const uint32_t kk = pa + 2 * pb;
    switch (op) {
    case TPGX_OP_SUB: q.x = (TPGX_K1C_SUB << 2) | kk; break;
    case TPGX_OP_ADD: q.x = (TPGX_K1C_ADD << 2) | kk; break;
    case TPGX_OP_MUL: q.x = (TPGX_K1C_MUL << 2) | kk; break;
    case TPGX_OP_MAX: q.x = (TPGX_K1C_MAX << 2) | kk; break;
    case TPGX_OP_COND: q.x = (TPGX_K1C_COND << 2) | kk; break;
    case TPGX_OP_EQ0: q.x = (TPGX_K1C_EQ0 << 2) | pa; break;
    case TPGX_OP_EQM1: q.x = (TPGX_K1C_EQM1 << 2) | pa; break;
    case TPGX_OP_EQ1: q.x = (TPGX_K1C_EQ1 << 2) | pa; break;
    case TPGX_OP_GE15: q.x = (TPGX_K1C_GE15 << 2) | pa; break;
    case TPGX_OP_PI: q.x = (TPGX_K1C_PI << 2) | pa; break;
    case TPGX_OP_DIV: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_DIV << 2) | kk; break;
    case TPGX_OP_FMODG: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_FMODG << 2) | kk; break;
    case TPGX_OP_MULC: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_MULC << 2) | pa; q.z = lb * 8u; break;
    case TPGX_OP_EXP: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_EXP << 2) | pa; break;
    case TPGX_OP_LOG: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_LOG << 2) | pa; break;
    case TPGX_OP_COS: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_COS << 2) | pa; break;
    case TPGX_OP_SIN: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_SIN << 2) | pa; break;
    case TPGX_OP_TAN: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_TAN << 2) | pa; break;
    /* window operand: la is the top-left pixel; the handlers read the precomputed plane at window index wp */
    case TPGX_OP_SOBEL_MAGN: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_SOBEL_MAGN << 2) | 1u; q.y = ((la / width) * (width - 2u) + la % width) * (ts / 2u); break;
    case TPGX_OP_SOBEL_DIR: q.x = TPGX_K1H_HEAVY | (TPGX_K1X_SOBEL_DIR << 2) | 1u; q.y = ((la / width) * (width - 2u) + la % width) * (ts / 2u); break;
    default: q.x = 0xffffffffu; break; /* int-typed instructions never reach K1 (no i32 source) */
    }
    return q;
}

enum tpgx_operand_type { TPGX_T_NONE = 0, TPGX_T_D = 1, TPGX_T_I = 2, TPGX_T_C = 3, TPGX_T_W = 4 };

struct tpgx_op_info {
    int nb_operands;
    int type[2];
    int flops; /* algorithmic FP64 operations per execution (SURVEY.md §8d convention) */
};

/* Operand signature per device opcode (= LambdaInstruction<...> template arguments of the apps). */
TPGX_HOSTDEV static inline tpgx_op_info tpgx_get_op_info(int op) {
    switch (op) {
    case TPGX_OP_SUB: case TPGX_OP_ADD: case TPGX_OP_MUL: case TPGX_OP_DIV: case TPGX_OP_MAX:
        return {2, {TPGX_T_D, TPGX_T_D}, 1};
    case TPGX_OP_FMODG: case TPGX_OP_COND:
        return {2, {TPGX_T_D, TPGX_T_D}, 2};
    case TPGX_OP_EXP: case TPGX_OP_LOG: case TPGX_OP_COS: case TPGX_OP_SIN: case TPGX_OP_TAN:
    case TPGX_OP_EQ0: case TPGX_OP_EQM1: case TPGX_OP_EQ1: case TPGX_OP_GE15:
        return {1, {TPGX_T_D, TPGX_T_NONE}, 1};
    case TPGX_OP_PI:
        return {1, {TPGX_T_D, TPGX_T_NONE}, 0};
    case TPGX_OP_MULC:
        return {2, {TPGX_T_D, TPGX_T_C}, 2};
    case TPGX_OP_SUB_II:
        return {2, {TPGX_T_I, TPGX_T_I}, 1};
    case TPGX_OP_CAST_I:
        return {1, {TPGX_T_I, TPGX_T_NONE}, 0};
    case TPGX_OP_SOBEL_MAGN:
        return {1, {TPGX_T_W, TPGX_T_NONE}, 20};
    case TPGX_OP_SOBEL_DIR:
        return {1, {TPGX_T_W, TPGX_T_NONE}, 18};
    default:
        return {0, {TPGX_T_NONE, TPGX_T_NONE}, 0};
    }
}

/* Address space of handler `s` of [registers, constants iff nb_constants > 0, data sources...] for operand type `t`;
 * 0 = the handler cannot provide it ([UP] DataHandler::getAddressSpace; PrimitiveTypeArray / Array2DWrapper rules,
 * SURVEY.md App. F U6/U7). */
static inline uint32_t tpgx_source_space(const tpgx_config& cfg, const std::vector<tpgx_source_desc>& src, int s, int t) {
    if (s == 0) return t == TPGX_T_D ? (uint32_t)cfg.nb_registers : 0u;
    const int first_env = cfg.nb_constants > 0 ? 2 : 1;
    if (cfg.nb_constants > 0 && s == 1) return t == TPGX_T_C ? (uint32_t)cfg.nb_constants : 0u;
    const int k = s - first_env;
    if (k < 0 || k >= (int)src.size()) return 0u;
    const tpgx_source_desc& d = src[k];
    if (d.elem == TPGX_ELEM_I32) return (t == TPGX_T_I && !d.is_2d) ? (uint32_t)d.width : 0u;
    if (t == TPGX_T_D) return (uint32_t)(d.height * d.width);
    if (t == TPGX_T_W && d.is_2d && d.height >= 3 && d.width >= 3) return (uint32_t)((d.height - 2) * (d.width - 2));
    return 0u;
}

/* Result of flattening a tpgx_graph for the device. */
struct tpgx_flat_graph {
    int nb_programs = 0, nb_teams = 0, nb_edges = 0, nb_roots = 0;
    std::vector<uint32_t> code;         /* packed effective lines of all programs              */
    std::vector<int32_t> prog_offset;   /* [P+1] into code                                      */
    std::vector<int32_t> prog_flops;    /* [P] algorithmic flops of one execution               */
    std::vector<double> constants;      /* [P][nb_constants]                                    */
    std::vector<int32_t> team_edge_offset, edge_program, edge_destination, root_team;
    int64_t total_lines = 0, total_introns = 0;
    uint32_t op_mask = 0;               /* bit o set: device opcode o occurs among the effective lines */
    std::vector<int32_t> v2_entries;    /* [P] entries of the program's K1 v2 stream (k1v2_encode.h), -1: not encodable for v2 */
    std::vector<uint8_t> v2_slots;      /* [P] shared-memory slots it needs there                */
};

/* Persistent host worker pool (flatten.cpp): run(nt, fn) calls fn(t) for t in [0, nt), t = 0 on the caller, the rest on
 * pool threads that sleep between calls — a generation's host work is a few hundred microseconds, less than spawning
 * threads for it costs. */
void tpgx_pool_run(int nt, void (*fn)(int, void*), void* arg);

/* Splits [0, n) into contiguous blocks over up to 16 host threads (the host side of a generation is
 * pure per-program work: intron marking, encoding); fn(begin, end, worker) runs once per block. */
static inline int tpgx_host_threads() {
    /* host threads of this process: TPGX_HOST_THREADS, else the cores divided by the ranks sharing the node
       (torchrun exports LOCAL_WORLD_SIZE), at most 16 */
    unsigned hw = std::thread::hardware_concurrency();
    if (const char* lw = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(lw); if (w > 1) hw = std::max(1u, hw / (unsigned)w); }
    int nt = (int)std::min<unsigned>(hw ? hw : 1u, 16u);
    if (const char* ht = getenv("TPGX_HOST_THREADS")) { const int v = atoi(ht); if (v >= 1) nt = std::min(v, 16); }
    return nt;
}
template <class F> static inline void tpgx_parallel_blocks(int n, int min_per_thread, F fn) {
    const int nt = std::max(1, std::min(tpgx_host_threads(), n / std::max(1, min_per_thread)));
    if (nt <= 1) { fn(0, n, 0); return; }
    struct Ctx { F* fn; int n, per; } c{&fn, n, (n + nt - 1) / nt};
    tpgx_pool_run(nt, [](int t, void* a) {
        Ctx& c = *static_cast<Ctx*>(a);
        (*c.fn)(std::min(c.n, t * c.per), std::min(c.n, (t + 1) * c.per), t);
    }, &c);
}
static inline int tpgx_parallel_workers() { return 16; }

/* Validates and flattens; returns TPGX_OK or an error code with a message in err. */
tpgx_status tpgx_flatten_graph(const tpgx_config& cfg, const std::vector<int32_t>& opcode,
                               const std::vector<tpgx_source_desc>& src, const tpgx_graph* g,
                               tpgx_flat_graph* out, std::string* err, std::vector<uint8_t>* keep_out = nullptr,
                               const std::vector<uint8_t>* prog_mask = nullptr, const int32_t* prog_ids = nullptr);

#endif
