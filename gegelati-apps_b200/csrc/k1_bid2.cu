/*
 * k1_bid2.cu — K1 v2, the threaded bid interpreter (sm_100a) for the MNIST instruction set.
 *
 * Replaces [UP] Program::ProgramExecutionEngine::executeProgram + Data::*::getDataAt + the mnist app's
 * instruction lambdas ([REF] mnist/src/main.cpp:47-88) and, in team mode, [UP] TPGExecutionEngine::evaluateTeam,
 * like K1 v1 (k1_bid.cu) — same results bit for bit, about half the instructions per line.  What changed:
 *
 *   dispatch     the line loop is ONE inline-PTX block: fetch the 16-byte entry (broadcast LDS.128), then a
 *                single warp-uniform indirect branch (brx.idx.uni -> LDC + BRX) to a handler specialised for
 *                (operation, kind of A, kind of B).  No compare chains, no BSSY / BSYNC, no predicated loads of
 *                the operand kind that is not there.
 *   accumulator  the result of a line stays in 8 machine registers (4 samples per lane).  The encoder
 *                (k1v2_encode.h) forwards it to the next line — 77 % of the register operands of effective
 *                programs — and emits a store only for results that are read later than that (26 %).
 *                Handler entry points fall through: <op>_RR loads A from its slot and continues into <op>_AR.
 *   slots        live-range allocation leaves at most K1V2_SLOTS (5) shared-memory slots per warp instead of 8
 *                fixed registers: 6.25 KB per warp incl. code buffers -> 32 warps per SM (v1: 24) and ~20 KB of L1 left.  Programs that need more run on v1.
 *   scheduling   a persistent grid, one CTA per SM; every WARP pulls (tile, group of units) items from a global
 *                counter, tile-major, so no CTA waits for its slowest warp and a shard of any size fills the
 *                machine to the last item (v1 launched (chunks x tiles) CTAs in waves of 148).
 *   arithmetic   identical sequences to tpgx_math.h (division, exp, ln, a*c/10), hand-scheduled in PTX so that
 *                operands arrive in, and results leave in, the accumulator's registers; the rare operands
 *                (subnormals, the slivers next to exp's overflow / underflow) leave the block and run the
 *                C++ routines.
 */
#include <cuda_runtime.h>

#include "dev_ops.cuh"
#include "k1v2_encode.h"
#include "kernels.cuh"

#ifndef K2_MINB
#define K2_MINB 1 /* CTAs per SM */
#endif
#ifndef K2_NW
#define K2_NW 32
#endif
#define K2_TS 128
#define K2_BUF (K1V2_CQ + K1V2_CAP) /* quads per code buffer */
#define K2_TAB_BYTES (256 * 8 + 512 * 8 + 512 * 8) /* exp table, ln table, exp / ln of the 256 pixel values */
#define K2_HEAD_BYTES (K2_TAB_BYTES + 16)        /* ... and the stream's address */
#define K2_WST_BYTES 64 /* per-warp row cursor record */
#define K2_WARP_FIXED (K2_WST_BYTES + 2 * K2_BUF * 16)
static_assert(K1V2_CQ == 5 && K1V2_HDR == 4 && K1V2_CAP == 36 && K2_TAB_BYTES == 10240, "the NEXT handler of the interpreter block spells these out");

__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k2_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void k2_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void k2_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

/* ---- the interpreter block --------------------------------------------------------------------------------
 * operands of the asm statement */
#define A0 "%0"
#define A1 "%1"
#define A2 "%2"
#define A3 "%3"
#define B0 "%4"
#define B1 "%5"
#define B2 "%6"
#define B3 "%7"
#define LP "%8"
#define CODE "%9"
#define REGS "%10"
#define GTILE "%11" /* the lane's 4 samples in pixel row 0 of the tile */
#define SOBEL "%12" /* the lane's 4 samples at window 0 of the tile's first plane */
#define CBASE "%13"
#define TAB "%14" /* exp table; +2048 ln table; +6144 exp / ln of the 256 pixel values */


/* IEEE-754 double constants of tpgx_math.h, as PTX literals */
#define KD_INV_LN2N "0d40671547652B82FE"    /* 128 / ln2 */
#define KD_MAGIC "0d4338000000000000"       /* 1.5 * 2^52 */
#define KD_NEG_LN2N_HI "0dBF762E42FF000000" /* -(ln2 / 128) high part */
#define KD_NEG_LN2N_LO "0d3D5718432A1B0E26" /* -(ln2 / 128) low part */
#define KD_C1_6 "0d3FC5555555555555"
#define KD_C1_24 "0d3FA5555555555555"
#define KD_C1_120 "0d3F81111111111111"
#define KD_HALF "0d3FE0000000000000"
#define KD_X710 "0d4086300000000000"
#define KD_XM746 "0dC087500000000000"
#define KD_LN2_HI "0d3FE62E42FEFA3800"
#define KD_LN2_LO "0d3D2EF35793C76730"
#define KD_C1_7 "0d3FC2492492492492"
#define KD_CM1_6 "0dBFC5555555555555"
#define KD_C0_2 "0d3FC999999999999A"
#define KD_CM0_25 "0dBFD0000000000000"
#define KD_C1_3 "0d3FD5555555555555"
#define KD_CM0_5 "0dBFE0000000000000"
#define KD_ONE "0d3FF0000000000000"
#define KD_MONE "0dBFF0000000000000"
#define KD_T01 "0d3FB999999999999A"         /* RN(0.1) */
#define KD_M10 "0dC024000000000000"
#define KD_ZERO "0d0000000000000000"
#define KD_INF "0d7FF0000000000000"
#define KD_NINF "0dFFF0000000000000"
#define KD_NAN "0d7FF8000000000000"

/* exit codes of the interpreter block beyond the handler ids: operands of a heavy operation that need the
   generic routine (tpgx_math.h).  Operands are left in the accumulator (and b) registers. */
#define K2X_SLOW_DIV 1000
#define K2X_SLOW_EXP 1001
#define K2X_SLOW_LN 1002
#define K2X_SLOW_MULC 1003

/* tpgx_math::xdiv_k, one sample k.  DIV_CLASS: reciprocal seed sd<k>, special-operand predicate psp<k>, pany |= !special.
   DIV_SEQ: quotient of a / b into qr<k> (a * seed when special); pall &= (special | fast path accepted). */
#define DIV_CLASS(a, b, k)                                                                       \
    "mov.b64 {alo, ahi}, " a ";\n mov.b64 {blo, bhi}, " b ";\n"                                  \
    "and.b32 ha, ahi, 0x7fffffff;\n and.b32 hb, bhi, 0x7fffffff;\n"                              \
    "or.b32 t0, ha, alo;\n setp.eq.u32 p1, t0, 0;\n setp.ge.u32 p2, ha, 0x7ff00000;\n or.pred pas, p1, p2;\n" \
    "or.b32 t0, hb, blo;\n setp.eq.u32 p1, t0, 0;\n setp.ge.u32 p2, hb, 0x7ff00000;\n or.pred pbs, p1, p2;\n" \
    "sub.u32 t0, hb, 0x00100000;\n setp.lt.u32 p1, t0, 0x7fb00000;\n and.pred pas, pas, p1;\n or.pred psp" k ", pbs, pas;\n" \
    "not.pred p1, psp" k ";\n or.pred pany, pany, p1;\n"
#define DIV_SEQ(a, b, k)                                                                         \
    "rcp.approx.ftz.f64 sd, " b ";\n"                                                            \
    "mov.b64 {slo, shi}, sd;\n mov.b32 slo, 1;\n mov.b64 rr, {slo, shi};\n"                      \
    "neg.f64 nb, " b ";\n"                                                                       \
    "fma.rn.f64 t1, nb, rr, " KD_ONE ";\n fma.rn.f64 t1, t1, t1, t1;\n fma.rn.f64 r1, rr, t1, rr;\n" \
    "fma.rn.f64 t2, nb, r1, " KD_ONE ";\n fma.rn.f64 r2, r1, t2, r1;\n"                           \
    "mul.rn.f64 q0, " a ", r2;\n fma.rn.f64 rem, nb, q0, " a ";\n fma.rn.f64 qr" k ", r2, rem, q0;\n" \
    "mov.b64 {qlo, qhi}, qr" k ";\n mov.b64 {alo, ahi}, " a ";\n mov.b64 {blo, bhi}, " b ";\n"   \
    "mov.b32 fq, qhi;\n mov.b32 fb, bhi;\n mov.b32 fa, ahi;\n"                                   \
    "fma.rn.f32 ft, 0f00000000, fb, fq;\n abs.f32 ft, ft;\n setp.gt.f32 p1, ft, 0f00100000;\n"   \
    "abs.f32 fa, fa;\n setp.ge.f32 p2, fa, 0f03600000;\n and.pred p1, p1, p2;\n"                 \
    "or.pred p1, p1, psp" k ";\n and.pred pall, pall, p1;\n"                                     \
    "mul.rn.f64 sp, " a ", sd;\n selp.f64 qr" k ", sp, qr" k ", psp" k ";\n"
/* every sample of the warp has a special operand: the quotient is a * seed */
#define DIV_SP(a, b) "rcp.approx.ftz.f64 sd, " b ";\n mul.rn.f64 " a ", " a ", sd;\n"

/* tpgx_math::exp_k, one sample (x in place).  EXP_RARE accumulates the sliver test (prare |= exp_is_rare(x), on the high word:
   |x| in [700, 710) for x > 0, [700, 746) for x < 0), EXP_MAIN computes.  The range predicates of a sample are made where they
   are used, after its polynomial: 8 predicates kept across the four samples' arithmetic do not fit the 7 predicate registers and
   ptxas spilled them into a GPR bit field (~20 extra instructions per line). */
#define EXP_RARE(x)                                                                               \
    "mov.b64 {xlo, xhi}, " x ";\n and.b32 ax, xhi, 0x7fffffff;\n sub.u32 t0, ax, 0x4085e000;\n"    \
    "shr.u32 ee, xhi, 31;\n mad.lo.u32 ee, ee, 0x12000, 0x5000;\n setp.lt.or.u32 prare, t0, ee, prare;\n"
#define EXP_MAIN(x)                                                                               \
    "mul.rn.f64 z, " x ", " KD_INV_LN2N ";\n add.rn.f64 tm, z, " KD_MAGIC ";\n sub.rn.f64 kd, tm, " KD_MAGIC ";\n" \
    "mov.b64 {ki, t0}, tm;\n"                                                                    \
    "fma.rn.f64 r, kd, " KD_NEG_LN2N_HI ", " x ";\n fma.rn.f64 r, kd, " KD_NEG_LN2N_LO ", r;\n"   \
    "and.b32 ii, ki, 127;\n shr.s32 ee, ki, 7;\n shl.b32 ii, ii, 4;\n add.u32 ii, ii, " TAB ";\n" \
    "ld.shared.v2.f64 {tail, TT}, [ii];\n"                                                       \
    "mul.rn.f64 r2, r, r;\n fma.rn.f64 pa, r, " KD_C1_6 ", " KD_HALF ";\n fma.rn.f64 pb, r, " KD_C1_120 ", " KD_C1_24 ";\n" \
    "add.rn.f64 tmp, tail, r;\n fma.rn.f64 tmp, r2, pa, tmp;\n mul.rn.f64 r4, r2, r2;\n fma.rn.f64 tmp, r4, pb, tmp;\n" \
    "fma.rn.f64 y, TT, tmp, TT;\n"                                                               \
    "mov.b64 {ylo, yhi}, y;\n shl.b32 ee, ee, 20;\n add.u32 yhi, yhi, ee;\n mov.b64 y, {ylo, yhi};\n" \
    "mov.b64 {xlo, xhi}, " x ";\n and.b32 ax, xhi, 0x7fffffff;\n setp.lt.u32 p1, ax, 0x4085e000;\n setp.le.f64 p2, " x ", " KD_XM746 ";\n" \
    "add.rn.f64 ev, " x ", " KD_INF ";\n selp.f64 ev, " KD_ZERO ", ev, p2;\n selp.f64 " x ", y, ev, p1;\n"

/* tpgx_math::log_k, one sample (x in place).  LN_RARE: prare |= positive subnormal, pany |= main range (positive, normal,
   finite).  LN_MAIN: the table-driven path + edge selects (zero, negative, NaN / inf), their predicates made at the end.
   LN_EDGE: the edge answers alone (no sample of the warp is in the main range). */
#define LN_RARE(x)                                                                                \
    "mov.b64 {xlo, xhi}, " x ";\n or.b32 t0, xhi, xlo;\n setp.ne.u32 p1, t0, 0;\n setp.lt.and.u32 p1, xhi, 0x00100000, p1;\n" \
    "or.pred prare, prare, p1;\n"                                                                \
    "sub.u32 t0, xhi, 0x00100000;\n setp.lt.or.u32 pany, t0, 0x7fe00000, pany;\n"
#define LN_EDGE(x)                                                                                \
    "add.rn.f64 ev, " x ", " x ";\n setp.lt.f64 p2, " x ", " KD_ZERO ";\n selp.f64 ev, " KD_NAN ", ev, p2;\n" \
    "setp.eq.f64 p1, " x ", " KD_ZERO ";\n selp.f64 " x ", " KD_NINF ", ev, p1;\n"
#define LN_MAIN(x)                                                                                \
    "mov.b64 {xlo, xhi}, " x ";\n"                                                               \
    "sub.u32 t0, xhi, 0x3FE6B000;\n shr.u32 ii, t0, 13;\n and.b32 ii, ii, 127;\n shr.s32 kk, t0, 20;\n" \
    "and.b32 t0, t0, 0xfff00000;\n sub.u32 t0, xhi, t0;\n mov.b64 z, {xlo, t0};\n"               \
    "mad.lo.u32 t0, ii, 8, " TAB ";\n mad.lo.u32 ii, ii, 16, " TAB ";\n ld.shared.v2.f64 {invc, lchi}, [ii+2048];\n ld.shared.f64 lclo, [t0+4096];\n" \
    "fma.rn.f64 r, z, invc, " KD_MONE ";\n cvt.rn.f64.s32 kd, kk;\n fma.rn.f64 w, kd, " KD_LN2_HI ", lchi;\n add.rn.f64 hi, w, r;\n" \
    "sub.rn.f64 lo, w, hi;\n add.rn.f64 lo, lo, r;\n fma.rn.f64 tmp, kd, " KD_LN2_LO ", lclo;\n add.rn.f64 lo, lo, tmp;\n" \
    "mul.rn.f64 r2, r, r;\n fma.rn.f64 pa, r, " KD_C1_7 ", " KD_CM1_6 ";\n fma.rn.f64 pa, r, pa, " KD_C0_2 ";\n" \
    "fma.rn.f64 pa, r, pa, " KD_CM0_25 ";\n fma.rn.f64 pa, r, pa, " KD_C1_3 ";\n fma.rn.f64 pa, r, pa, " KD_CM0_5 ";\n" \
    "fma.rn.f64 tmp, r2, pa, lo;\n add.rn.f64 y, hi, tmp;\n"                                     \
    "add.rn.f64 ev, " x ", " x ";\n setp.lt.f64 p2, " x ", " KD_ZERO ";\n selp.f64 ev, " KD_NAN ", ev, p2;\n" \
    "setp.eq.f64 p1, " x ", " KD_ZERO ";\n selp.f64 ev, " KD_NINF ", ev, p1;\n"                  \
    "sub.u32 t0, xhi, 0x00100000;\n setp.lt.u32 p1, t0, 0x7fe00000;\n selp.f64 " x ", y, ev, p1;\n"

/* tpgx_math::xdiv10_k, one sample: x / 10 into q; pall &= accepted */
#define DIV10_1(x, q)                                                                             \
    "mul.rn.f64 q0, " x ", " KD_T01 ";\n fma.rn.f64 rem, " KD_M10 ", q0, " x ";\n fma.rn.f64 " q ", rem, " KD_T01 ", q0;\n" \
    "mov.b64 {xlo, xhi}, " x ";\n and.b32 ha, xhi, 0x7fffffff;\n"                                \
    "or.b32 t0, ha, xlo;\n setp.eq.u32 p1, t0, 0;\n setp.ge.u32 p2, ha, 0x7ff00000;\n or.pred psp, p1, p2;\n" \
    "sub.u32 t0, ha, 0x01700000;\n setp.lt.u32 p1, t0, 0x7d000000;\n"                            \
    "mov.b64 {qlo, qhi}, " q ";\n mov.b64 {slo, shi}, q0;\n xor.b32 t0, qhi, shi;\n and.b32 t0, t0, 0x7ff00000;\n setp.eq.u32 p2, t0, 0;\n" \
    "and.pred p1, p1, p2;\n or.pred p1, p1, psp;\n and.pred pall, pall, p1;\n"                   \
    "selp.f64 " q ", q0, " q ", psp;\n"

/* Dispatch: one fetch + indirect branch, every handler ends with a branch back to it.  The measured alternative (every
 * handler fetches the NEXT entry and ends with its own indirect branch) executes 4 % fewer instructions and is 21 % SLOWER on
 * C2: 28 BRX sites plus 3.5 KB more code drop the instruction-cache hit rate from 99.3 % to 93 % (profiles/r2_k1v2_experiments.md). */
#define PF ""
/* Sobel planes: the lane's 4 samples are one 32-byte sector, read once per (program, tile) by one 256-bit load that does
   not displace anything in what is left of L1 (8.22 -> 8.15 ms on C2).  As two 16-byte loads, L1::no_allocate was 4.6 %
   SLOWER: the second load missed the sector the first one had brought in. */
#define LDSOB "ld.global.nc.L1::no_allocate.v4.f64"
#define GO "bra.uni FETCH;\n"
#define UNFETCH ""

/* address of a shared-memory slot operand (byte offset in field f) / of the lane's packed pixel word (word index in f) */
#define ADDR_R(f) "add.u32 ra, " f ", " REGS ";\n"
#define ADDR_P(f) "mad.wide.u32 ga, " f ", 4, " GTILE ";\n"
/* 4 doubles of the lane from the slot at ra */
#define LD_R(d0, d1, d2, d3)                               \
    "ld.shared.v2.f64 {" d0 ", " d1 "}, [ra];\n"           \
    "ld.shared.v2.f64 {" d2 ", " d3 "}, [ra+512];\n"
/* 4 pixels of the lane from ga, uint8 -> double (exact) */
#define LD_P(d0, d1, d2, d3)                               \
    "ld.global.nc.u32 pw, [ga];\n"                         \
    "cvt.rn.f64.u8 " d0 ", pw;\n"                          \
    "shr.u32 t0, pw, 8;\n"                                 \
    "cvt.rn.f64.u8 " d1 ", t0;\n"                          \
    "shr.u32 t0, pw, 16;\n"                                \
    "cvt.rn.f64.u8 " d2 ", t0;\n"                          \
    "shr.u32 t0, pw, 24;\n"                                \
    "cvt.rn.f64.u8 " d3 ", t0;\n"
/* operand A of the entry into the accumulator (not the last use of the entry's fields: no PF) */
#define LOAD_AR ADDR_R("ya") LD_R(A0, A1, A2, A3)
#define LOAD_AP ADDR_P("ya") LD_P(A0, A1, A2, A3)
/* last operand of the entry into b: address, then the next entry's fetch, then the load */
#define LAST_BR(f) ADDR_R(f) PF LD_R(B0, B1, B2, B3)
#define LAST_BP(f) ADDR_P(f) PF LD_P(B0, B1, B2, B3)
#define LAST_AR ADDR_R("ya") PF LD_R(A0, A1, A2, A3)
#define LAST_AP ADDR_P("ya") PF LD_P(A0, A1, A2, A3)
#define MOV_B_A "mov.f64 " B0 ", " A0 ";\n mov.f64 " B1 ", " A1 ";\n mov.f64 " B2 ", " A2 ";\n mov.f64 " B3 ", " A3 ";\n"
/* acc = acc op b / acc = b op acc / acc = acc op acc */
#define OP_AB(op) op " " A0 ", " A0 ", " B0 ";\n" op " " A1 ", " A1 ", " B1 ";\n" op " " A2 ", " A2 ", " B2 ";\n" op " " A3 ", " A3 ", " B3 ";\n"
#define OP_BA(op) op " " A0 ", " B0 ", " A0 ";\n" op " " A1 ", " B1 ", " A1 ";\n" op " " A2 ", " B2 ", " A2 ";\n" op " " A3 ", " B3 ", " A3 ";\n"
#define OP_AA(op) op " " A0 ", " A0 ", " A0 ";\n" op " " A1 ", " A1 ", " A1 ";\n" op " " A2 ", " A2 ", " A2 ";\n" op " " A3 ", " A3 ", " A3 ";\n"
/* [REF] mnist/src/main.cpp: max = (a < b) ? b : a  (NaN in a stays, NaN in b is dropped) */
#define MAX1(a, b) "setp.lt.f64 pq, " a ", " b ";\n selp.f64 " a ", " b ", " a ", pq;\n"
#define RMAX1(a, b) "setp.lt.f64 pq, " b ", " a ";\n selp.f64 " a ", " a ", " b ", pq;\n" /* a = (b < a) ? a : b, operand A in b */
#define MAX_AB MAX1(A0, B0) MAX1(A1, B1) MAX1(A2, B2) MAX1(A3, B3)
#define MAX_BA RMAX1(A0, B0) RMAX1(A1, B1) RMAX1(A2, B2) RMAX1(A3, B3)

/* commutative binary operation: AA AR RR AP RP PP */
#define BIN_COMM(N, op)                                                   \
    "H_" N "_PP:\n" LOAD_AP "bra.uni H_" N "_AP;\n"                       \
    "H_" N "_RP:\n" LOAD_AR                                               \
    "H_" N "_AP:\n" LAST_BP("zb") OP_AB(op) GO                            \
    "H_" N "_RR:\n" LOAD_AR                                               \
    "H_" N "_AR:\n" LAST_BR("zb") OP_AB(op) GO                            \
    "H_" N "_AA:\n" PF OP_AA(op) GO
/* ordered binary operation: AA AR RR PR AP RP PP RA PA, with the given bodies */
#define BIN_ORD(N, body_ab, body_ba, body_aa)                             \
    "H_" N "_PP:\n" LOAD_AP "bra.uni H_" N "_AP;\n"                       \
    "H_" N "_RP:\n" LOAD_AR                                               \
    "H_" N "_AP:\n" LAST_BP("zb") body_ab GO                              \
    "H_" N "_PR:\n" LOAD_AP "bra.uni H_" N "_AR;\n"                       \
    "H_" N "_RR:\n" LOAD_AR                                               \
    "H_" N "_AR:\n" LAST_BR("zb") body_ab GO                              \
    "H_" N "_RA:\n" LAST_BR("ya") body_ba GO                              \
    "H_" N "_PA:\n" LAST_BP("ya") body_ba GO                              \
    "H_" N "_AA:\n" PF body_aa GO

#define K2_TARGET(n) "H_" #n ", "

template <bool TEAM>
__global__ void __launch_bounds__(K2_NW * 32, K2_MINB) tpgx_bid2_kernel(const tpgx_bid2_params p) {
    /* shared memory: [exp, ln, pixel exp / ln tables] then per warp [row cursor record | 2 code buffers | p.slots register slots] */
    extern __shared__ __align__(128) uint8_t smem[];
    double* s_tab = reinterpret_cast<double*>(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    /* the ln table {1/c, ln c high, ln c low, -} is split into 16-byte pairs {1/c, ln c high} and a plane of ln c low: 32-byte entries
       put the 32 lanes' lookups on 4 groups of 8 banks (~11 wavefronts per load), 16-byte entries on 8 groups and the 8-byte plane on 16 */
    for (int i = threadIdx.x; i < 256 + 512 + 512; i += K2_NW * 32) {
        if (i < 256) s_tab[i] = tpgx_math::tpgx_exp_tab_dev[i];
        else if (i < 768) {
            const int e = (i - 256) >> 2, c = (i - 256) & 3;
            if (c < 2) s_tab[256 + 2 * e + c] = tpgx_math::tpgx_log_tab_dev[i - 256];
            else if (c == 2) s_tab[512 + e] = tpgx_math::tpgx_log_tab_dev[i - 256];
        } else s_tab[i] = __ldg(p.pixel_lut + (i - 768));
    }
    if (threadIdx.x == 0) *reinterpret_cast<const uint4**>(smem + K2_TAB_BYTES) = p.stream;
    const uint32_t tab = k2_smem_u32(s_tab);
    __syncthreads();
    /* from here on every warp is on its own */
    const uint32_t wst = tab + K2_HEAD_BYTES + warp * (K2_WARP_FIXED + p.slots * (K2_TS * 8));
    const uint32_t mycode = wst + K2_WST_BYTES;
    const uint32_t regs_lane = wst + K2_WARP_FIXED + lane * 16;
    const uint4* __restrict__ stream = p.stream;
    const unsigned int tile_bytes = (unsigned int)(p.hw + 1) * K2_TS;
    const unsigned int tile_doubles = 2u * (unsigned int)p.nwin * K2_TS; /* Sobel planes of a tile */

    /* cursor over the rows (programs) this warp runs.  item = (tile, group of consecutive units); the rows of consecutive
       units are consecutive in the stream, and every row's header quad names the edge's destination, where the row after
       it starts and whether it opens / closes its unit — so a warp reads global metadata once per ITEM (unit_first) and
       nothing per row.  The cursor is warp-uniform and needed only between two rows: it lives in a 64-byte per-warp
       record in shared memory, not in registers across the interpreter block (whose heavy handlers need them: the
       register file is 64 x 1024 threads). */
    struct Cur { int tile, unit, unit_end, q; }; /* q: first quad of the row in the stream */
    auto claim = [&]() {
        unsigned int i = 0;
        if (lane == 0) i = atomicAdd(p.work_counter, 1u);
        return (int)__shfl_sync(0xffffffffu, i, 0);
    };
    auto open_item = [&](Cur& c) {
        const int item = claim();
        if (item >= p.nb_items) { c.tile = -1; c.unit = 0; c.unit_end = 0; c.q = 0; return; }
        c.tile = item / p.items_per_tile;
        const int g = item - c.tile * p.items_per_tile;
        c.unit = g * p.units_per_item;
        c.unit_end = min(c.unit + p.units_per_item, p.nb_units);
        c.q = __ldg(&p.unit_first[c.unit].x);
    };
    /* asynchronous copy of a row's head (constants + header) and first block of entries into a code buffer: always the
       K2_BUF quads of a full buffer (past a short row that is the next row, past the last row the stream's slack) */
    auto stage = [&](uint32_t buf, const Cur& c) {
        if (c.tile >= 0) {
            const uint4* src = stream + (unsigned int)c.q + lane;
            k2_cp_async16(buf + lane * 16, src);
            if (lane < K2_BUF - 32) k2_cp_async16(buf + (32 + lane) * 16, src + 32);
        }
        k2_cp_async_commit();
    };
    /* the record: [0..3] the row to run next, [8..9] {tile, unit} of the row being run.  Every lane
       writes the same words and reads them back itself: no synchronisation needed */
    auto put_next = [&](const Cur& c) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(wst), "r"(c.tile), "r"(c.unit), "r"(c.unit_end), "r"(c.q) : "memory");
    };
    auto get_next = [&](Cur& c) {
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(c.tile), "=r"(c.unit), "=r"(c.unit_end), "=r"(c.q) : "r"(wst) : "memory");
    };

    {
        Cur c0;
        open_item(c0);
        stage(mycode, c0);
        put_next(c0);
    }
    uint32_t code = mycode;
    double best[4];
    int best_dst[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { best[k] = 0.0; best_dst[k] = 0; }

    for (;;) {
        const uint8_t* __restrict__ gtile;
        const double* __restrict__ sobel_tile;
        {
            Cur c;
            get_next(c);
            if (c.tile < 0) break;
            k2_cp_async_wait_all();
            __syncwarp();
            uint32_t h_dst, h_next, h_flags, h_first;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h_dst), "=r"(h_next), "=r"(h_flags), "=r"(h_first) : "r"(code + K1V2_HDR * 16));
            /* the row after this one */
            Cur nx = c;
            nx.q = (int)h_next;
            if (h_flags & K1V2_ROW_LAST) {
                if (++nx.unit >= c.unit_end) open_item(nx);
            }
            stage(2 * mycode + K2_BUF * 16 - code, nx); /* next row's code: in flight while this one runs */
            put_next(nx);
            asm volatile("st.shared.v2.u32 [%0+32], {%1, %2};" ::"r"(wst), "r"(c.tile), "r"(c.unit) : "memory");
            gtile = p.tiles + (size_t)((unsigned long long)(unsigned int)c.tile * tile_bytes) + lane * 4;
            sobel_tile = p.sobel + (size_t)((unsigned long long)(unsigned int)c.tile * tile_doubles) + lane * 4;
        }
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        uint32_t lp = code + K1V2_CQ * 16;
        for (;;) {
            uint32_t xc;
            double b0, b1, b2, b3; /* second operand of the heavy handlers; meaningful after a slow-path exit only */
            asm volatile(
                "{\n"
                ".reg .b32 hx, ya, zb, wd, ra, pw, t0;\n"
                ".reg .b64 ga;\n"
                ".reg .pred pq, p1, p2, pas, pbs, psp, psp0, psp1, psp2, psp3, pall, pany, prare, pslow;\n"
                ".reg .b32 alo, ahi, blo, bhi, ha, hb, slo, shi, qlo, qhi, xlo, xhi, ax, ki, ii, ee, kk, ylo, yhi;\n"
                ".reg .f32 fq, fb, fa, ft;\n"
                ".reg .f64 sd, rr, nb, t1, t2, r1, r2, q0, rem, sp, qr0, qr1, qr2, qr3, cst;\n"
                ".reg .f64 z, tm, kd, r, tail, TT, pa, pb, tmp, r4, y, ev, invc, lchi, lclo, w, hi, lo;\n"
                "tlist: .branchtargets " K1V2_HANDLERS(K2_TARGET) "H_END;\n"
                "FETCH:\n"
                "ld.shared.v4.u32 {hx, ya, zb, wd}, [" LP "];\n"
                "add.u32 " LP ", " LP ", 16;\n"
                "brx.idx.uni hx, tlist;\n"
                "H_ST:\n"
                "add.u32 ra, wd, " REGS ";\n"
                PF
                "st.shared.v2.f64 [ra], {" A0 ", " A1 "};\n"
                "st.shared.v2.f64 [ra+512], {" A2 ", " A3 "};\n"
                GO
                BIN_COMM("ADD", "add.rn.f64")
                BIN_COMM("MUL", "mul.rn.f64")
                BIN_ORD("SUB", OP_AB("sub.rn.f64"), OP_BA("sub.rn.f64"), OP_AA("sub.rn.f64"))
                BIN_ORD("MAX", MAX_AB, MAX_BA, "")
                /* Sobel planes: one 32-byte group per lane */
                "H_SOBM:\n"
                "H_SOBD:\n"
                "mad.wide.u32 ga, ya, 32, " SOBEL ";\n"
                PF
                LDSOB " {" A0 ", " A1 ", " A2 ", " A3 "}, [ga];\n"
                GO
                /* exp / ln of a pixel: 256-entry tables */
                "H_LUTL:\n"
                ADDR_P("ya") PF
                "ld.global.nc.u32 pw, [ga];\n"
                "and.b32 t0, pw, 255;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A0 ", [t0+8192];\n"
                "bfe.u32 t0, pw, 8, 8;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A1 ", [t0+8192];\n"
                "bfe.u32 t0, pw, 16, 8;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A2 ", [t0+8192];\n"
                "shr.u32 t0, pw, 24;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A3 ", [t0+8192];\n"
                GO
                "H_LUTE:\n"
                ADDR_P("ya") PF
                "ld.global.nc.u32 pw, [ga];\n"
                "and.b32 t0, pw, 255;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A0 ", [t0+6144];\n"
                "bfe.u32 t0, pw, 8, 8;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A1 ", [t0+6144];\n"
                "bfe.u32 t0, pw, 16, 8;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A2 ", [t0+6144];\n"
                "shr.u32 t0, pw, 24;\n mad.lo.u32 t0, t0, 8, " TAB ";\n ld.shared.f64 " A3 ", [t0+6144];\n"
                GO
                /* division: operands into (acc, b) in order, then the body */
                "H_DIV_PP:\n" LOAD_AP "bra.uni H_DIV_AP;\n"
                "H_DIV_RP:\n" LOAD_AR
                "H_DIV_AP:\n" LAST_BP("zb") "bra.uni DIVBODY;\n"
                "H_DIV_PR:\n" LOAD_AP "bra.uni H_DIV_AR;\n"
                "H_DIV_RR:\n" LOAD_AR
                "H_DIV_AR:\n" LAST_BR("zb") "bra.uni DIVBODY;\n"
                "H_DIV_RA:\n" MOV_B_A LAST_AR "bra.uni DIVBODY;\n"
                "H_DIV_PA:\n" MOV_B_A LAST_AP "bra.uni DIVBODY;\n"
                "H_DIV_AA:\n" MOV_B_A PF
                "DIVBODY:\n"
                /* classes first: when every sample of the warp has a zero / infinite / NaN operand (a fifth of the
                   divisions of evolved programs) the quotients are a * seed and the Newton sequence is skipped */
                "setp.ne.u32 pany, 0, 0;\n"
                DIV_CLASS(A0, B0, "0") DIV_CLASS(A1, B1, "1") DIV_CLASS(A2, B2, "2") DIV_CLASS(A3, B3, "3")
                "vote.sync.any.pred p1, pany, 0xffffffff;\n"
                "@!p1 bra.uni DIV_ALLSP;\n"
                "setp.eq.u32 pall, 0, 0;\n"
                DIV_SEQ(A0, B0, "0") DIV_SEQ(A1, B1, "1") DIV_SEQ(A2, B2, "2") DIV_SEQ(A3, B3, "3")
                "vote.sync.all.pred p1, pall, 0xffffffff;\n"
                "@!p1 bra.uni X_SLOW_DIV;\n"
                "mov.f64 " A0 ", qr0;\n mov.f64 " A1 ", qr1;\n mov.f64 " A2 ", qr2;\n mov.f64 " A3 ", qr3;\n"
                GO
                "DIV_ALLSP:\n"
                DIV_SP(A0, B0) DIV_SP(A1, B1) DIV_SP(A2, B2) DIV_SP(A3, B3)
                GO
                /* exp / ln: the rare operands (the slivers next to exp's thresholds, positive subnormals for ln) are known
                   from the inputs; when a lane has one, the inputs are kept in b, the fast path runs for everybody and
                   the block is left afterwards so that the C++ side patches just those samples */
                "H_EXP_R:\n" LOAD_AR
                "H_EXP_A:\n"
                PF
                "setp.ne.u32 prare, 0, 0;\n"
                EXP_RARE(A0) EXP_RARE(A1) EXP_RARE(A2) EXP_RARE(A3)
                "vote.sync.any.pred pslow, prare, 0xffffffff;\n"
                "@pslow bra.uni EXP_KEEP;\n"
                "EXP_GO:\n"
                EXP_MAIN(A0) EXP_MAIN(A1) EXP_MAIN(A2) EXP_MAIN(A3)
                "@pslow bra.uni X_SLOW_EXP;\n"
                GO
                "EXP_KEEP:\n" MOV_B_A "bra.uni EXP_GO;\n"
                "H_LN_R:\n" LOAD_AR
                "H_LN_A:\n"
                PF
                "setp.ne.u32 prare, 0, 0;\n setp.ne.u32 pany, 0, 0;\n"
                LN_RARE(A0) LN_RARE(A1) LN_RARE(A2) LN_RARE(A3)
                "vote.sync.any.pred pslow, prare, 0xffffffff;\n"
                "@pslow bra.uni LN_KEEP;\n"
                "LN_GO:\n"
                /* no sample of the warp is positive, normal and finite (a fifth of the ln lines): edge answers only */
                "vote.sync.any.pred p1, pany, 0xffffffff;\n"
                "@!p1 bra.uni LN_EDGES;\n"
                LN_MAIN(A0) LN_MAIN(A1) LN_MAIN(A2) LN_MAIN(A3)
                "@pslow bra.uni X_SLOW_LN;\n"
                GO
                "LN_EDGES:\n"
                LN_EDGE(A0) LN_EDGE(A1) LN_EDGE(A2) LN_EDGE(A3)
                "@pslow bra.uni X_SLOW_LN;\n"
                GO
                "LN_KEEP:\n" MOV_B_A "bra.uni LN_GO;\n"
                /* multByConst: a * c / 10 ([REF] mnist/src/main.cpp:54) */
                "H_MULC_P:\n" LOAD_AP "bra.uni H_MULC_A;\n"
                "H_MULC_R:\n" LOAD_AR
                "H_MULC_A:\n"
                "add.u32 ra, zb, " CBASE ";\n"
                PF
                "ld.shared.f64 cst, [ra];\n"
                "mul.rn.f64 " A0 ", " A0 ", cst;\n mul.rn.f64 " A1 ", " A1 ", cst;\n mul.rn.f64 " A2 ", " A2 ", cst;\n mul.rn.f64 " A3 ", " A3 ", cst;\n"
                "setp.eq.u32 pall, 0, 0;\n"
                DIV10_1(A0, "qr0") DIV10_1(A1, "qr1") DIV10_1(A2, "qr2") DIV10_1(A3, "qr3")
                "vote.sync.all.pred pslow, pall, 0xffffffff;\n"
                "@!pslow bra.uni MULC_KEEP;\n"
                "MULC_FIN:\n"
                "mov.f64 " A0 ", qr0;\n mov.f64 " A1 ", qr1;\n mov.f64 " A2 ", qr2;\n mov.f64 " A3 ", qr3;\n"
                "@!pslow bra.uni X_SLOW_MULC;\n"
                GO
                "MULC_KEEP:\n" MOV_B_A "bra.uni MULC_FIN;\n"
                "X_SLOW_DIV:\n" UNFETCH "mov.u32 " CODE ", 1000;\n bra.uni XEND;\n"
                "X_SLOW_EXP:\n" UNFETCH "mov.u32 " CODE ", 1001;\n bra.uni XEND;\n"
                "X_SLOW_LN:\n" UNFETCH "mov.u32 " CODE ", 1002;\n bra.uni XEND;\n"
                "X_SLOW_MULC:\n" UNFETCH "mov.u32 " CODE ", 1003;\n bra.uni XEND;\n"
                /* programs longer than one block: the NEXT entry names the row-relative index of the next block's first
                   entry (y); the row header's w is the row's first entry in the stream.  K1V2_CAP quads are copied (past the
                   row's end that is the next row / the stream's slack: never dispatched) */
                "H_NEXT:\n"
                "vote.sync.all.pred p1, p1, 0xffffffff;\n" /* every lane has fetched the NEXT entry before the buffer is overwritten.  (vote, not
                                                               bar.warp.sync: a warp barrier inside the loop makes ptxas wrap every
                                                               dispatch in a BSSY / BSYNC pair — +3 instructions per entry; the warp never
                                                               diverges in this block and its shared-memory accesses execute in order) */
                "ld.shared.u32 t0, [" CBASE "+76];\n"     /* header quad (K1V2_HDR * 16), word 3 */
                "add.u32 ya, ya, t0;\n"
                "mov.u32 t0, %%laneid;\n"
                "add.u32 ya, ya, t0;\n"
                "ld.shared.u64 ga, [" TAB "+10240];\n"    /* K2_TAB_BYTES: the stream's address */
                "mad.wide.u32 ga, ya, 16, ga;\n"
                "shl.b32 t0, t0, 4;\n add.u32 ra, t0, " CBASE ";\n"
                "ld.global.nc.v4.u32 {hx, ya, zb, wd}, [ga];\n"
                "st.shared.v4.u32 [ra+80], {hx, ya, zb, wd};\n"   /* K1V2_CQ * 16 */
                "setp.lt.u32 p1, t0, 64;\n"                       /* (K1V2_CAP - 32) lanes copy the tail */
                "@p1 ld.global.nc.v4.u32 {hx, ya, zb, wd}, [ga+512];\n"
                "@p1 st.shared.v4.u32 [ra+592], {hx, ya, zb, wd};\n"
                "vote.sync.all.pred p1, p1, 0xffffffff;\n"
                "add.u32 " LP ", " CBASE ", 80;\n"
                "bra.uni FETCH;\n"
                "H_END:\n"
                "mov.u32 " CODE ", hx;\n"
                "XEND:\n"
                "}\n"
                : "+d"(a0), "+d"(a1), "+d"(a2), "+d"(a3), "=&d"(b0), "=&d"(b1), "=&d"(b2), "=&d"(b3), "+r"(lp), "=&r"(xc)
                : "r"(regs_lane), "l"(gtile), "l"(sobel_tile), "r"(code), "r"(tab)
                : "memory");
            if (xc == K2H_END) break;
            /* rare operands of a heavy operation (a subnormal somewhere, the slivers next to exp's overflow / underflow
               thresholds, |a * c| outside 2^+-1000): the fast path has answered every other sample; patch these */
            if (xc == K2X_SLOW_DIV) { /* operands intact in (a, b) */
                a0 = tpgx_math::xdiv(a0, b0); a1 = tpgx_math::xdiv(a1, b1); a2 = tpgx_math::xdiv(a2, b2); a3 = tpgx_math::xdiv(a3, b3);
            } else if (xc == K2X_SLOW_EXP) { /* inputs in b, fast results in a */
                if (tpgx_math::exp_is_rare(b0)) a0 = tpgx_math::exp_slow(b0);
                if (tpgx_math::exp_is_rare(b1)) a1 = tpgx_math::exp_slow(b1);
                if (tpgx_math::exp_is_rare(b2)) a2 = tpgx_math::exp_slow(b2);
                if (tpgx_math::exp_is_rare(b3)) a3 = tpgx_math::exp_slow(b3);
            } else if (xc == K2X_SLOW_LN) {
                if (tpgx_math::log_is_rare(b0)) a0 = tpgx_math::log_slow(b0);
                if (tpgx_math::log_is_rare(b1)) a1 = tpgx_math::log_slow(b1);
                if (tpgx_math::log_is_rare(b2)) a2 = tpgx_math::log_slow(b2);
                if (tpgx_math::log_is_rare(b3)) a3 = tpgx_math::log_slow(b3);
            } else { /* K2X_SLOW_MULC: the products a * c in b, fast quotients in a */
                if (!tpgx_math::div10_fast_ok(b0)) a0 = tpgx_math::xdiv(b0, 10.0);
                if (!tpgx_math::div10_fast_ok(b1)) a1 = tpgx_math::xdiv(b1, 10.0);
                if (!tpgx_math::div10_fast_ok(b2)) a2 = tpgx_math::xdiv(b2, 10.0);
                if (!tpgx_math::div10_fast_ok(b3)) a3 = tpgx_math::xdiv(b3, 10.0);
            }
        }
        /* the last effective line of a program writes register 0: the bid is the accumulator.
           [UP] TPGExecutionEngine::evaluateEdge: NaN bid -> -inf */
        double res[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (res[k] != res[k]) res[k] = __longlong_as_double(0xfff0000000000000LL);
        uint32_t h_dst, h_next, h_flags, h_first;
        int c_tile, c_unit;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(h_dst), "=r"(h_next), "=r"(h_flags), "=r"(h_first) : "r"(code + K1V2_HDR * 16) : "memory");
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+32];" : "=r"(c_tile), "=r"(c_unit) : "r"(wst) : "memory");
        if (!TEAM) {
            double* out = p.bid + (size_t)c_unit * p.row_stride + (size_t)c_tile * K2_TS + lane * 4; /* bid mode: unit = row */
            *reinterpret_cast<double2*>(out) = make_double2(res[0], res[1]);
            *reinterpret_cast<double2*>(out + 2) = make_double2(res[2], res[3]);
        } else {
            /* [UP] evaluateTeam: the first edge is the initial best; a later edge replaces it if
               bid >= best (last-max) / bid > best (first-max) */
            const int dst = (int)h_dst;
            const unsigned int first = h_flags & K1V2_ROW_FIRST;
            if (p.tie_break == TPGX_TIE_LAST_MAX) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    asm("{\n .reg .pred pf, pt;\n setp.ne.u32 pf, %3, 0;\n setp.ge.or.f64 pt, %2, %0, pf;\n selp.f64 %0, %2, %0, pt;\n selp.b32 %1, %4, %1, pt;\n}"
                        : "+d"(best[k]), "+r"(best_dst[k]) : "d"(res[k]), "r"(first), "r"(dst));
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    asm("{\n .reg .pred pf, pt;\n setp.ne.u32 pf, %3, 0;\n setp.gt.or.f64 pt, %2, %0, pf;\n selp.f64 %0, %2, %0, pt;\n selp.b32 %1, %4, %1, pt;\n}"
                        : "+d"(best[k]), "+r"(best_dst[k]) : "d"(res[k]), "r"(first), "r"(dst));
            }
            if (h_flags & K1V2_ROW_LAST) {
                int* out = p.decision + (size_t)c_unit * p.row_stride + (size_t)c_tile * K2_TS + lane * 4;
                *reinterpret_cast<int4*>(out) = make_int4(best_dst[0], best_dst[1], best_dst[2], best_dst[3]);
            }
        }
        code = 2 * mycode + K2_BUF * 16 - code; /* the other buffer */
    }
    k2_cp_async_wait_all();
}

size_t tpgx_bid2_smem_bytes(int slots) { return (size_t)K2_HEAD_BYTES + (size_t)K2_NW * (K2_WARP_FIXED + (size_t)slots * K2_TS * 8); }

cudaError_t tpgx_launch_bid2(const tpgx_bid2_params& p, bool team, int nb_ctas, cudaStream_t st) {
    if (p.slots < 1 || p.slots > K1V2_SLOTS) return cudaErrorInvalidValue;
    const size_t smem = tpgx_bid2_smem_bytes(p.slots);
    cudaError_t e = cudaMemsetAsync(p.work_counter, 0, 4, st);
    if (e != cudaSuccess) return e;
    if (team) {
        e = cudaFuncSetAttribute(tpgx_bid2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tpgx_bid2_kernel<true><<<nb_ctas, K2_NW * 32, smem, st>>>(p); TPGX_COUNT_LAUNCH();
    } else {
        e = cudaFuncSetAttribute(tpgx_bid2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tpgx_bid2_kernel<false><<<nb_ctas, K2_NW * 32, smem, st>>>(p); TPGX_COUNT_LAUNCH();
    }
    return cudaGetLastError();
}

/* the stream of K1 v2, written on the device from the generic code: one thread per row */
__global__ void tpgx_k1v2_encode_kernel(const uint32_t* __restrict__ code, const int32_t* __restrict__ prog_offset,
                                        const double* __restrict__ constants, int nb_constants, const int32_t* __restrict__ active_prog,
                                        const int2* __restrict__ desc, const int32_t* __restrict__ row_dst, const uint8_t* __restrict__ row_flags,
                                        int nb_rows, int width, int hw, int nwin, uint4* __restrict__ stream, int* __restrict__ flags) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nb_rows) return;
    const int pidx = active_prog[row];
    const int2 d = desc[row];
    double* cq = reinterpret_cast<double*>(stream + d.x);
    for (int k = 0; k < 2 * K1V2_HDR; k++) cq[k] = k < nb_constants ? constants[(size_t)pidx * nb_constants + k] : 0.0;
    stream[d.x + K1V2_HDR] = make_uint4(row_dst ? (uint32_t)row_dst[row] : 0u, (uint32_t)(d.x + K1V2_CQ + d.y),
                                        row_flags ? (uint32_t)row_flags[row] : (K1V2_ROW_FIRST | K1V2_ROW_LAST), (uint32_t)(d.x + K1V2_CQ));
    const tpgx_k1v2_info info = tpgx_k1v2_encode(code + prog_offset[pidx], prog_offset[pidx + 1] - prog_offset[pidx], K2_TS, (uint32_t)width,
                                                 (uint32_t)hw, (uint32_t)nwin, reinterpret_cast<tpgx_k1_quad*>(stream + d.x + K1V2_CQ));
    if (!info.ok || info.entries != d.y || info.slots > K1V2_SLOTS) atomicOr(flags, 4);
}
cudaError_t tpgx_launch_k1v2_encode(const uint32_t* code, const int32_t* prog_offset, const double* constants, int nb_constants,
                                    const int32_t* active_prog, const int2* desc, const int32_t* row_dst, const uint8_t* row_flags, int nb_rows,
                                    int width, int hw, int nwin, uint4* stream, int* flags, cudaStream_t st) {
    if (nb_rows <= 0) return cudaSuccess;
    tpgx_k1v2_encode_kernel<<<(nb_rows + 127) / 128, 128, 0, st>>>(code, prog_offset, constants, nb_constants, active_prog, desc, row_dst, row_flags,
                                                                  nb_rows, width, hw, nwin, stream, flags); TPGX_COUNT_LAUNCH();
    return cudaGetLastError();
}
