/*
 * dev_ops.cuh — device-side instruction semantics and the per-thread program executor.
 *
 * apply_op() is the device opcode table: one case per distinct lambda of the five apps'
 * Instructions::Set ([REF] mnist/src/main.cpp:47-77,
 * pendulum/src/Learn/instructions.cpp:7-18, stickgame/src/Learn/instructions.cpp:9-16,
 * tic-tac-toe/src/Learn/main.cpp:15-23).  Compiled with -fmad=false: every + - * / below is one
 * IEEE-754 operation in the reference's source order; transcendentals come from tpgx_math.h.
 */
#ifndef TPGX_DEV_OPS_CUH
#define TPGX_DEV_OPS_CUH

#include <cfloat>

#include "tpgx_internal.h"
#include "tpgx_math.h"

#define TPGX_PI 3.14159265358979323846 /* M_PI */

/* Masks over opcodes: which ones read operand A / B as a double value. */
#define TPGX_MASK_A_DOUBLE                                                                                   \
    ((1u << TPGX_OP_SUB) | (1u << TPGX_OP_ADD) | (1u << TPGX_OP_MUL) | (1u << TPGX_OP_DIV) | (1u << TPGX_OP_MAX) |   \
     (1u << TPGX_OP_EXP) | (1u << TPGX_OP_LOG) | (1u << TPGX_OP_COS) | (1u << TPGX_OP_SIN) | (1u << TPGX_OP_TAN) |   \
     (1u << TPGX_OP_MULC) | (1u << TPGX_OP_FMODG) | (1u << TPGX_OP_EQ0) | (1u << TPGX_OP_EQM1) |                  \
     (1u << TPGX_OP_EQ1) | (1u << TPGX_OP_GE15) | (1u << TPGX_OP_COND))
#define TPGX_MASK_B_DOUBLE                                                                                   \
    ((1u << TPGX_OP_SUB) | (1u << TPGX_OP_ADD) | (1u << TPGX_OP_MUL) | (1u << TPGX_OP_DIV) | (1u << TPGX_OP_MAX) |   \
     (1u << TPGX_OP_FMODG) | (1u << TPGX_OP_COND))

/* Scalar ops on already-fetched double operands (everything except int and window ops). */
template <bool TRIG>
__device__ __forceinline__ double tpgx_apply_dd(uint32_t op, double a, double b) {
    switch (op) {
    case TPGX_OP_SUB: return a - b;
    case TPGX_OP_ADD: return a + b;
    case TPGX_OP_MUL: return a * b;
    case TPGX_OP_DIV: return tpgx_math::xdiv(a, b);
    case TPGX_OP_MAX: return (a < b) ? b : a;
    case TPGX_OP_EXP: return tpgx_math::exp(a);
    case TPGX_OP_LOG: return tpgx_math::log(a);
    case TPGX_OP_COS: if (TRIG) return tpgx_math::cos(a); else return 0.0;
    case TPGX_OP_SIN: if (TRIG) return tpgx_math::sin(a); else return 0.0;
    case TPGX_OP_TAN: if (TRIG) return tpgx_math::tan(a); else return 0.0;
    case TPGX_OP_PI: return TPGX_PI;
    case TPGX_OP_MULC: return tpgx_math::xdiv10(a * b);  /* b carries the constant */
    case TPGX_OP_FMODG: return b != 0.0 ? fmod(a, b) : DBL_MIN;
    case TPGX_OP_EQ0: return (a == 0.0) ? 10.0 : 0.0;
    case TPGX_OP_EQM1: return (a == -1.0) ? 10.0 : 0.0;
    case TPGX_OP_EQ1: return (a == 1.0) ? 10.0 : 0.0;
    case TPGX_OP_GE15: return (a >= 15.0) ? 10.0 : 0.0;
    case TPGX_OP_COND: return a < b ? -a : a;
    default: return 0.0;
    }
}

/* Sobel on a double window in the reference's literal operation order. */
__device__ __forceinline__ double tpgx_sobel_f64(uint32_t op, const double* w) {
    double gx = -w[0] + w[2] - 2.0 * w[3] + 2.0 * w[5] - w[6] + w[8];
    double gy = -w[0] - 2.0 * w[1] - w[2] + w[6] + 2.0 * w[7] + w[8];
    if (op == TPGX_OP_SOBEL_MAGN) return sqrt(gx * gx + gy * gy);
    return tpgx_math::atan(tpgx_math::xdiv(gy, gx));
}

/* Observation of one episode / sample for the per-thread executor. */
struct tpgx_obs_view {
    const double* f64[2];
    const int32_t* i32[2];
    int width[2]; /* row length of 2-D f64 sources (for 3x3 windows) */
};

/* Per-thread restatement of [UP] ProgramExecutionEngine::executeProgram on packed bytecode:
 * used where lanes may run different programs (episode kernels) and by tpgx_execute_programs.
 * Returns register 0 = result of the last effective line (0.0 for an empty program). */
__device__ __forceinline__ double tpgx_run_program(const uint32_t* __restrict__ code, int n,
                                                   const double* __restrict__ consts,
                                                   const tpgx_obs_view& o) {
    double reg[TPGX_MAX_REGS];
    double res = 0.0;
    uint32_t next = n > 0 ? __ldg(code) : 0u, prev = 0xffffffffu;
    for (int i = 0; i < n; i++) {
        const uint32_t w = next;
        if (i + 1 < n) next = __ldg(code + i + 1); /* the next line's word is in flight while this line runs */
        const uint32_t op = w & 31u, dst = (w >> 5) & 7u;
        const uint32_t fa = (w >> 8) & 0xfffu, fb = w >> 20;
        const uint32_t ka = fa >> 10, la = fa & 1023u, kb = fb >> 10, lb = fb & 1023u;
        if ((TPGX_MASK_A_DOUBLE >> op) & 1u) { /* the common case first: double operands */
            /* a register operand that names the previous line's destination is that line's result, still in `res`:
               most operands of an effective program are, and it keeps the local-memory round trip off the chain */
            double a, b = 0.0;
            if (fa == prev) a = res;
            else if (ka == TPGX_KIND_REG) a = (la < TPGX_LOC_ZERO) ? reg[la] : 0.0;
            else a = o.f64[ka - 2][la];
            if ((TPGX_MASK_B_DOUBLE >> op) & 1u) {
                if (fb == prev) b = res;
                else if (kb == TPGX_KIND_REG) b = (lb < TPGX_LOC_ZERO) ? reg[lb] : 0.0;
                else b = o.f64[kb - 2][lb];
            } else if (op == TPGX_OP_MULC) b = consts[lb];
            res = tpgx_apply_dd<true>(op, a, b);
        } else if (op == TPGX_OP_SOBEL_MAGN || op == TPGX_OP_SOBEL_DIR) {
            const int s = (int)ka - 2;
            const double* base = o.f64[s] + la;
            const int wd = o.width[s];
            double win[9];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) win[r * 3 + c] = base[r * wd + c];
            res = tpgx_sobel_f64(op, win);
        } else if (op == TPGX_OP_SUB_II) {
            const int ia = o.i32[ka - 2][la], ib = o.i32[kb - 2][lb];
            res = (double)ia - (double)ib;
        } else if (op == TPGX_OP_CAST_I) {
            res = (double)o.i32[ka - 2][la];
        } else {
            res = tpgx_apply_dd<true>(op, 0.0, 0.0); /* no operand (pi) */
        }
        reg[dst] = res;
        prev = dst; /* as an operand field: kind REG (0) << 10 | dst */
    }
    return res;
}

#endif
