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

#include <string>
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

/* ---- K1 "fused" line word (built from the generic word by tpgx_k1_word) ---------------------------
 *   [ 6: 0] K1 opcode   [ 9: 7] destination register   [19:10] operand A loc   [29:20] operand B loc
 *   [30] A is a pixel (heavy opcodes)   [31] B is a pixel (heavy opcodes)
 * Cheap instructions carry the operand kinds in the opcode: binary id*4 + (A pixel) + 2*(B pixel),
 * unary TPGX_K1_UNARY0 + id*2 + (A pixel).  A register operand's loc is its slot (8 = zero slot). */
#define TPGX_K1_UNARY0 20
enum {
    TPGX_K1_DIV = 32, TPGX_K1_MULC, TPGX_K1_FMODG, TPGX_K1_EXP, TPGX_K1_LOG, TPGX_K1_COS, TPGX_K1_SIN, TPGX_K1_TAN,
    TPGX_K1_SOBEL_MAGN = 48, TPGX_K1_SOBEL_DIR = 49, TPGX_K1_INVALID = 127
};

static inline uint32_t tpgx_k1_word(uint32_t w) {
    const uint32_t op = w & 31u, dst = (w >> 5) & 7u;
    const uint32_t fa = (w >> 8) & 0xfffu, fb = w >> 20;
    const uint32_t pa = (fa >> 10) == TPGX_KIND_SRC0, pb = (fb >> 10) == TPGX_KIND_SRC0;
    uint32_t k1;
    switch (op) {
    case TPGX_OP_SUB: k1 = 0 * 4 + pa + 2 * pb; break;
    case TPGX_OP_ADD: k1 = 1 * 4 + pa + 2 * pb; break;
    case TPGX_OP_MUL: k1 = 2 * 4 + pa + 2 * pb; break;
    case TPGX_OP_MAX: k1 = 3 * 4 + pa + 2 * pb; break;
    case TPGX_OP_COND: k1 = 4 * 4 + pa + 2 * pb; break;
    case TPGX_OP_EQ0: k1 = TPGX_K1_UNARY0 + 0 * 2 + pa; break;
    case TPGX_OP_EQM1: k1 = TPGX_K1_UNARY0 + 1 * 2 + pa; break;
    case TPGX_OP_EQ1: k1 = TPGX_K1_UNARY0 + 2 * 2 + pa; break;
    case TPGX_OP_GE15: k1 = TPGX_K1_UNARY0 + 3 * 2 + pa; break;
    case TPGX_OP_PI: k1 = TPGX_K1_UNARY0 + 4 * 2 + pa; break;
    case TPGX_OP_DIV: k1 = TPGX_K1_DIV; break;
    case TPGX_OP_MULC: k1 = TPGX_K1_MULC; break;
    case TPGX_OP_FMODG: k1 = TPGX_K1_FMODG; break;
    case TPGX_OP_EXP: k1 = TPGX_K1_EXP; break;
    case TPGX_OP_LOG: k1 = TPGX_K1_LOG; break;
    case TPGX_OP_COS: k1 = TPGX_K1_COS; break;
    case TPGX_OP_SIN: k1 = TPGX_K1_SIN; break;
    case TPGX_OP_TAN: k1 = TPGX_K1_TAN; break;
    case TPGX_OP_SOBEL_MAGN: k1 = TPGX_K1_SOBEL_MAGN; break;
    case TPGX_OP_SOBEL_DIR: k1 = TPGX_K1_SOBEL_DIR; break;
    default: k1 = TPGX_K1_INVALID; break; /* int-typed instructions never reach K1 (no i32 source) */
    }
    return k1 | (dst << 7) | ((fa & 1023u) << 10) | ((fb & 1023u) << 20) | (pa << 30) | (pb << 31);
}

enum tpgx_operand_type { TPGX_T_NONE = 0, TPGX_T_D = 1, TPGX_T_I = 2, TPGX_T_C = 3, TPGX_T_W = 4 };

struct tpgx_op_info {
    int nb_operands;
    int type[2];
    int flops; /* algorithmic FP64 operations per execution (SURVEY.md §8d convention) */
};

/* Operand signature per device opcode (= LambdaInstruction<...> template arguments of the apps). */
static inline tpgx_op_info tpgx_get_op_info(int op) {
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

/* Result of flattening a tpgx_graph for the device. */
struct tpgx_flat_graph {
    int nb_programs = 0, nb_teams = 0, nb_edges = 0, nb_roots = 0;
    std::vector<uint32_t> code;         /* packed effective lines of all programs              */
    std::vector<int32_t> prog_offset;   /* [P+1] into code                                      */
    std::vector<int32_t> prog_flops;    /* [P] algorithmic flops of one execution               */
    std::vector<double> constants;      /* [P][nb_constants]                                    */
    std::vector<int32_t> team_edge_offset, edge_program, edge_destination, root_team;
    int64_t total_lines = 0, total_introns = 0;
};

/* Validates and flattens; returns TPGX_OK or an error code with a message in err. */
tpgx_status tpgx_flatten_graph(const tpgx_config& cfg, const std::vector<int32_t>& opcode,
                               const std::vector<tpgx_source_desc>& src, const tpgx_graph* g,
                               tpgx_flat_graph* out, std::string* err);

#endif
