/*
 * k1v2_encode.h — program encoder of the threaded bid interpreter (csrc/k1_bid2.cu).
 *
 * Input: the generic 32-bit line words of ONE program (tpgx_internal.h: introns removed, locations reduced,
 * never-written registers marked).  Output: the program's entry stream for K1 v2, 16 bytes per entry
 * {x = handler id, y = operand A, z = operand B, w = destination slot}, ready to be dispatched by ONE
 * indirect branch per entry.  What the encoder decides, statically and per program:
 *
 *   accumulator forwarding  the result of a line stays in the interpreter's machine registers ("the
 *       accumulator", kind A).  A later operand that names the PREVIOUS line's destination reads the
 *       accumulator: no shared-memory load.  A result only the next line reads is never stored.
 *   slot allocation         a result that a line after the next one still reads gets a slot of the warp's
 *       shared-memory register file by live range (lowest free slot; freed at its last use), written by an
 *       explicit ST entry.  [UP] ProgramExecutionEngine's 8 TPG registers thus become at most
 *       `slots needed` <= 8 physical slots; programs that fit K1V2_SLOTS run on K1 v2, the others on K1 v1.
 *   handler specialisation  one handler per (operation, kind of A, kind of B), kinds being accumulator /
 *       slot / pixel; commutative operations are canonicalised (the accumulator or the slot first).
 *   blocks                  K1 v2 stages K1V2_CAP entries at a time in shared memory; longer programs are
 *       chained with NEXT entries (y = row-relative index of the next block's first entry), every program ends with END.
 *
 * The function compiles for the host (slot / entry counts during planning, unit tests) and for the device
 * (the stream itself is written by tpgx_k1v2_encode_kernel from the generic code already in HBM).
 */
#ifndef TPGX_K1V2_ENCODE_H
#define TPGX_K1V2_ENCODE_H

#include "tpgx_internal.h"

#define K1V2_SLOTS 5      /* shared-memory register slots per warp */
#define K1V2_CQ 5         /* quads at the head of a row's stream: 4 of program constants (8 doubles), then the row header */
#define K1V2_HDR 4        /* header quad {x = destination of the edge, y = first quad of the NEXT row of the stream, z = flags, w = the row's first entry in the stream} */
#define K1V2_ROW_LAST 1u  /* flags: last row (edge) of its unit (team) */
#define K1V2_ROW_FIRST 2u /* first row of its unit */
#define K1V2_CAP 36       /* entries staged at a time */
#define K1V2_MAX_LINES 256 /* longer programs run on K1 v1 */

/* handler ids, in the order of the interpreter's branch-target table.  Binary operations: <op>_<A><B> with
 * A = accumulator, R = slot, P = pixel; the operand that is not the accumulator travels in y (operand A) or z
 * (operand B) as the byte offset of its slot, or as the index of lane 0's packed pixel word in the tile. */
#define K1V2_HANDLERS(X)                                                                                   \
    X(END) X(NEXT) X(ST)                                                                                   \
    X(ADD_AA) X(ADD_AR) X(ADD_RR) X(ADD_AP) X(ADD_RP) X(ADD_PP)                                            \
    X(MUL_AA) X(MUL_AR) X(MUL_RR) X(MUL_AP) X(MUL_RP) X(MUL_PP)                                            \
    X(SUB_AA) X(SUB_AR) X(SUB_RR) X(SUB_PR) X(SUB_AP) X(SUB_RP) X(SUB_PP) X(SUB_RA) X(SUB_PA)              \
    X(MAX_AA) X(MAX_AR) X(MAX_RR) X(MAX_PR) X(MAX_AP) X(MAX_RP) X(MAX_PP) X(MAX_RA) X(MAX_PA)              \
    X(DIV_AA) X(DIV_AR) X(DIV_RR) X(DIV_PR) X(DIV_AP) X(DIV_RP) X(DIV_PP) X(DIV_RA) X(DIV_PA)              \
    X(EXP_A) X(EXP_R) X(LN_A) X(LN_R) X(LUTE) X(LUTL)                                                      \
    X(MULC_A) X(MULC_R) X(MULC_P)                                                                          \
    X(SOBM) X(SOBD)
enum {
#define K1V2_ENUM(n) K2H_##n,
    K1V2_HANDLERS(K1V2_ENUM)
#undef K1V2_ENUM
    K2H_COUNT
};

struct tpgx_k1v2_info {
    int entries;  /* stream entries incl. ST, NEXT and END (the stream is K1V2_CQ + entries quads) */
    int slots;    /* shared-memory slots the program needs */
    int ok;       /* 0: not encodable for K1 v2 (instruction outside the MNIST set, too long) */
};

/* kinds of an operand after forwarding */
#define K2K_A 0u
#define K2K_R 1u
#define K2K_P 2u

/* handler of a binary operation given the operand kinds; *swap = 1 when y / z must be exchanged (commutative
 * canonicalisation) */
TPGX_HOSTDEV static inline int tpgx_k1v2_binary(int op, uint32_t ka, uint32_t kb, int* swap) {
    *swap = 0;
    if (op == TPGX_OP_ADD || op == TPGX_OP_MUL) {
        const int base = op == TPGX_OP_ADD ? K2H_ADD_AA : K2H_MUL_AA;
        /* order: A before R before P */
        if (ka > kb) { const uint32_t t = ka; ka = kb; kb = t; *swap = 1; }
        if (ka == K2K_A) return base + (kb == K2K_A ? 0 : kb == K2K_R ? 1 : 3);             /* AA AR AP */
        if (ka == K2K_R) return base + (kb == K2K_R ? 2 : 4);                                  /* RR RP */
        return base + 5;                                                                       /* PP */
    }
    const int base = op == TPGX_OP_SUB ? K2H_SUB_AA : op == TPGX_OP_MAX ? K2H_MAX_AA : K2H_DIV_AA;
    /* AA AR RR PR AP RP PP RA PA */
    if (kb == K2K_A) return base + (ka == K2K_A ? 0 : ka == K2K_R ? 7 : 8);
    if (kb == K2K_R) return base + (ka == K2K_A ? 1 : ka == K2K_R ? 2 : 3);
    return base + (ka == K2K_A ? 4 : ka == K2K_R ? 5 : 6);
}

/* Encodes one program.  words[n]: generic line words; ts: samples per tile (128); width, hw: row length and
 * element count of the 2-D source; nwin: window positions per observation.  out == nullptr: count only. */
TPGX_HOSTDEV static inline tpgx_k1v2_info tpgx_k1v2_encode(const uint32_t* words, int n, uint32_t ts, uint32_t width, uint32_t hw,
                                                          uint32_t nwin, tpgx_k1_quad* out) {
    tpgx_k1v2_info info = {0, 0, 1};
    if (n > K1V2_MAX_LINES) { info.ok = 0; return info; }
    int pos = 0; /* entry position within the staged block */
    auto emit = [&](uint32_t h, uint32_t y, uint32_t z, uint32_t w) {
        if (pos == K1V2_CAP - 1 && h != K2H_END) { /* the block's last entry chains to the next block */
            if (out) { out[info.entries].x = K2H_NEXT; out[info.entries].y = (uint32_t)info.entries + 1u; out[info.entries].z = out[info.entries].w = 0; }
            info.entries++;
            pos = 0;
        }
        if (out) { out[info.entries].x = h; out[info.entries].y = y; out[info.entries].z = z; out[info.entries].w = w; }
        info.entries++;
        pos++;
    };
    /* last_read[i]: the last line reading the value line i defines (-1: none), in one forward pass: a line's register
       operands name the lines that defined them */
    int16_t last_read[K1V2_MAX_LINES];
    {
        int def_line[TPGX_MAX_REGS];
        for (int r = 0; r < TPGX_MAX_REGS; r++) def_line[r] = -1;
        for (int i = 0; i < n; i++) {
            const uint32_t w = words[i];
            const tpgx_op_info oi = tpgx_get_op_info((int)(w & 31u));
            const uint32_t f[2] = {(w >> 8) & 0xfffu, w >> 20};
            for (int k = 0; k < oi.nb_operands && k < 2; k++)
                if (oi.type[k] == TPGX_T_D && (f[k] >> 10) == TPGX_KIND_REG && (f[k] & 1023u) < TPGX_MAX_REGS && def_line[f[k] & 1023u] >= 0)
                    last_read[def_line[f[k] & 1023u]] = (int16_t)i;
            last_read[i] = -1;
            def_line[(w >> 5) & 7u] = i;
        }
    }
    int slot_of[TPGX_MAX_REGS];   /* slot holding the current value of each TPG register, -1: none */
    int def_of[TPGX_MAX_REGS];    /* line that wrote it */
    int last_use_of[TPGX_MAX_REGS]; /* last line reading the current value */
    for (int r = 0; r < TPGX_MAX_REGS; r++) { slot_of[r] = -1; def_of[r] = -2; last_use_of[r] = -1; }
    uint32_t free_mask = 0xffu;
    const uint32_t zero_word = hw * (ts / 4u); /* index of lane 0's packed word of the all-zero pixel row */
    if (n == 0) emit(K2H_ADD_PP, zero_word, zero_word, 0); /* [UP] executeProgram: register 0 is still 0.0 */
    for (int i = 0; i < n; i++) {
        const uint32_t word = words[i];
        const int op = (int)(word & 31u);
        const uint32_t dst = (word >> 5) & 7u;
        const uint32_t f[2] = {(word >> 8) & 0xfffu, word >> 20};
        const tpgx_op_info oi = tpgx_get_op_info(op);
        /* operands after forwarding */
        uint32_t kind[2] = {K2K_P, K2K_P}, val[2] = {zero_word, zero_word};
        int reg[2] = {-1, -1};
        for (int k = 0; k < 2; k++) {
            if (k >= oi.nb_operands) continue;
            const uint32_t kd = f[k] >> 10, loc = f[k] & 1023u;
            if (oi.type[k] == TPGX_T_C) { val[k] = loc * 8u; continue; }
            if (oi.type[k] == TPGX_T_W) { /* window: index of lane 0's 32-byte group in the tile's Sobel planes */
                val[k] = ((loc / width) * (width - 2u) + loc % width) * (ts / 4u);
                continue;
            }
            if (oi.type[k] != TPGX_T_D) { info.ok = 0; continue; }
            if (kd == TPGX_KIND_REG && loc < TPGX_LOC_ZERO) {
                reg[k] = (int)loc;
                if (def_of[loc] == i - 1) kind[k] = K2K_A;
                else { kind[k] = K2K_R; val[k] = (uint32_t)slot_of[loc] * ts * 8u; if (slot_of[loc] < 0) info.ok = 0; }
            } else if (kd == TPGX_KIND_REG) { kind[k] = K2K_P; val[k] = zero_word; }
            else if (kd == TPGX_KIND_SRC0) { kind[k] = K2K_P; val[k] = loc * (ts / 4u); }
            else info.ok = 0;
        }
        /* slots whose value dies here are free for this line's own result */
        for (int k = 0; k < 2; k++) {
            const int r = reg[k];
            if (r >= 0 && last_use_of[r] == i && slot_of[r] >= 0) { free_mask |= 1u << slot_of[r]; slot_of[r] = -1; }
        }
        switch (op) {
        case TPGX_OP_ADD: case TPGX_OP_MUL: case TPGX_OP_SUB: case TPGX_OP_MAX: case TPGX_OP_DIV: {
            int swap;
            const int h = tpgx_k1v2_binary(op, kind[0], kind[1], &swap);
            emit((uint32_t)h, swap ? val[1] : val[0], swap ? val[0] : val[1], 0);
            break;
        }
        case TPGX_OP_EXP: case TPGX_OP_LOG:
            if (kind[0] == K2K_P) emit(op == TPGX_OP_EXP ? K2H_LUTE : K2H_LUTL, val[0], 0, 0);
            else emit((op == TPGX_OP_EXP ? K2H_EXP_A : K2H_LN_A) + (kind[0] == K2K_R ? 1 : 0), val[0], 0, 0);
            break;
        case TPGX_OP_MULC:
            emit(K2H_MULC_A + kind[0], val[0], val[1], 0);
            break;
        case TPGX_OP_SOBEL_MAGN: emit(K2H_SOBM, val[0], 0, 0); break;
        case TPGX_OP_SOBEL_DIR: emit(K2H_SOBD, val[0] + nwin * (ts / 4u), 0, 0); break;
        default: info.ok = 0; break;
        }
        /* the new value of dst: the last line that reads it before it is overwritten */
        const int last = last_read[i];
        if (slot_of[dst] >= 0) { free_mask |= 1u << slot_of[dst]; slot_of[dst] = -1; } /* stale value overwritten (cannot be live) */
        def_of[dst] = i;
        last_use_of[dst] = last;
        if (last > i + 1) { /* read by a line after the next one: needs a slot */
            int s = 0;
            while (s < 8 && !((free_mask >> s) & 1u)) s++;
            if (s >= 8) { info.ok = 0; s = 0; }
            free_mask &= ~(1u << s);
            slot_of[dst] = s;
            if (s + 1 > info.slots) info.slots = s + 1;
            emit(K2H_ST, 0, 0, (uint32_t)s * ts * 8u);
        }
    }
    emit(K2H_END, 0, 0, 0);
    return info;
}

/* Count pass on its own, for the host planner: entries and slots of tpgx_k1v2_encode(words, n, ..., nullptr) without the
 * operand / handler work (the stream itself is written by the device, which reports any disagreement with these counts).
 * Same liveness and slot rules, statement for statement. */
static inline tpgx_k1v2_info tpgx_k1v2_count(const uint32_t* words, int n) {
    tpgx_k1v2_info info = {0, 0, 1};
    if (n > K1V2_MAX_LINES) { info.ok = 0; return info; }
    /* per opcode: bit 0-1 number of operands, bit 2 / 3 operand 0 / 1 is a register-file (double) operand, bit 7 in the MNIST set */
    static const uint8_t sig[32] = {
#define K1V2_SIG(op) ((op) == TPGX_OP_ADD || (op) == TPGX_OP_MUL || (op) == TPGX_OP_SUB || (op) == TPGX_OP_MAX || (op) == TPGX_OP_DIV ? (2 | 4 | 8 | 128) \
                      : (op) == TPGX_OP_EXP || (op) == TPGX_OP_LOG ? (1 | 4 | 128) : (op) == TPGX_OP_MULC ? (2 | 4 | 128)                                    \
                      : (op) == TPGX_OP_SOBEL_MAGN || (op) == TPGX_OP_SOBEL_DIR ? (1 | 128) : 0)
        K1V2_SIG(0), K1V2_SIG(1), K1V2_SIG(2), K1V2_SIG(3), K1V2_SIG(4), K1V2_SIG(5), K1V2_SIG(6), K1V2_SIG(7), K1V2_SIG(8), K1V2_SIG(9), K1V2_SIG(10),
        K1V2_SIG(11), K1V2_SIG(12), K1V2_SIG(13), K1V2_SIG(14), K1V2_SIG(15), K1V2_SIG(16), K1V2_SIG(17), K1V2_SIG(18), K1V2_SIG(19), K1V2_SIG(20),
        K1V2_SIG(21), K1V2_SIG(22), K1V2_SIG(23), K1V2_SIG(24), K1V2_SIG(25), K1V2_SIG(26), K1V2_SIG(27), K1V2_SIG(28), K1V2_SIG(29), K1V2_SIG(30), K1V2_SIG(31)
#undef K1V2_SIG
    };
    /* What the encoder's slot allocator and entry emitter amount to, without replaying them (the replay is a chain of dependent
       loads and stores through the per-register tables, 30 ns per line: it was 60 % of the host time of a generation's flatten):
         - the value line i defines needs a slot and an ST entry iff its last reader is a line after the next one;
         - such a value holds its slot from line i's allocation until line `last` frees it BEFORE allocating its own, i.e. over
           the allocation steps [i, last); lowest-free-slot allocation in order of definition is greedy colouring of an interval
           graph by left end point, which uses exactly as many slots as the largest number of values live at once;
         - a NEXT entry precedes every (K1V2_CAP - 1)-th entry that is not the first; END closes the row.
       The encoder's own consistency checks (a register operand is the accumulator or holds a slot) cannot fail on flattener
       output and stay with the encoder, which runs on the device and reports any disagreement with these counts (flag 4).
       Index TPGX_MAX_REGS of def_line and index -1 of last_read are dummies that absorb "not a register operand". */
    int16_t last_read_store[K1V2_MAX_LINES + 1];
    int16_t* const last_read = last_read_store + 1;
    uint32_t ok = 1u;
    {
        int def_line[TPGX_MAX_REGS + 1];
        for (int r = 0; r <= TPGX_MAX_REGS; r++) def_line[r] = -1;
        for (int i = 0; i < n; i++) {
            const uint32_t w = words[i];
            const uint32_t sg = sig[w & 31u];
            ok &= sg >> 7;
            const uint32_t f[2] = {(w >> 8) & 0xfffu, w >> 20};
            for (uint32_t k = 0; k < 2; k++) {
                const uint32_t kd = f[k] >> 10, loc = f[k] & 1023u;
                const uint32_t active = (uint32_t)(k < (sg & 3u)) & ((sg >> (2u + k)) & 1u);
                const uint32_t isreg = active & (uint32_t)(kd == TPGX_KIND_REG) & (uint32_t)(loc < TPGX_LOC_ZERO);
                ok &= 1u ^ (active & (uint32_t)(kd != TPGX_KIND_REG) & (uint32_t)(kd != TPGX_KIND_SRC0)); /* an operand kind K1 v2 has no handler for */
                last_read[def_line[TPGX_MAX_REGS ^ ((TPGX_MAX_REGS ^ loc) & (0u - isreg))]] = (int16_t)i; /* isreg ? loc : the dummy, without a branch */
            }
            last_read[i] = -1;
            def_line[(w >> 5) & 7u] = i;
        }
    }
    int8_t delta[K1V2_MAX_LINES + 1]; /* slots taken (+1 at the defining line) and given back (-1 at the last reader) */
    for (int i = 0; i <= n; i++) delta[i] = 0;
    int stores = 0;
    for (int i = 0; i < n; i++) {
        const int last = last_read[i];
        const int need = last > i + 1;
        stores += need;
        delta[i] += (int8_t)need;
        delta[n ^ ((n ^ last) & -need)] -= (int8_t)need; /* need ? last : n */
    }
    int live = 0, slots = 0;
    for (int i = 0; i < n; i++) {
        live += delta[i];
        slots = std::max(slots, live);
    }
    const int real = n == 0 ? 1 : n + stores; /* an empty program is one ADD of two zero rows */
    info.entries = real + (real - 1) / (K1V2_CAP - 1) + 1;
    info.slots = slots;
    info.ok = (int)(ok & (uint32_t)(slots <= 8));
    return info;
}

#endif
