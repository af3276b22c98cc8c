/*
 * k1_bid.cu — K1, the bid interpreter (sm_100a).
 *
 * Replaces [UP] Program::ProgramExecutionEngine::executeProgram + Data::*::getDataAt + the apps'
 * instruction lambdas ([REF] mnist/src/main.cpp:47-88) for a whole generation at once:
 * bid[row][s] = register 0 of program `row` on observation s, NaN mapped to -inf like
 * [UP] TPGExecutionEngine::evaluateEdge.
 *
 * Mapping.  CTA = one sample tile (TS = 32*K samples, pixel-major uint8; read in place through L1/L2
 * by the default variant, staged by one bulk TMA copy by the TILE_SMEM variants) x one chunk of work
 * units.  Each warp repeatedly claims a unit (program, or team in team mode) of the chunk and
 * interprets it for all TS samples of the tile: lane = K consecutive samples, so instruction dispatch
 * is warp-uniform and never diverges.  Two builds of the kernel: the MNIST instruction set only
 * (EXT = false: sub add mul div max exp ln multByConst sobelMagn sobelDir) and the extended one with
 * every double-typed instruction of the five apps.
 *
 * Where the instructions went (ncu, round 1: the previous version issued 137 warp instructions per
 * line of 64 samples, 26 of them FP64 arithmetic) and what this version does about it:
 *   - decode: lines arrive pre-decoded as 16-byte quads (tpgx_k1_line): handler id, operand byte
 *     offsets, destination byte offset.  No shifts, masks or multiplies per line.
 *   - fetch: the quads of a program sit in a per-warp shared-memory buffer and are read with one
 *     broadcast LDS.128 per line.  The NEXT program's quads are copied global -> shared with
 *     cp.async into the warp's second buffer while the current program runs, so switching programs
 *     costs a wait + __syncwarp and holds no registers.
 *   - every handler loads its operands, computes and stores its result by itself, so no values
 *     merge after the dispatch (the previous version spent ~8 register moves per line there); the
 *     bid is read back from register slot 0 at the end of the program.
 *   - TPG registers live in shared memory as [warp][8 slots][K/2][32 lanes][2] doubles: every access
 *     is a conflict-free 16-byte-per-lane vector.  There is no per-program clear: a read of a register
 *     no earlier line wrote is encoded as a read of the all-zero pixel row K0 appends to every tile.
 *   - pixel and Sobel-plane operands are addressed as block-uniform base + 32-bit element index, which
 *     is one IMAD.WIDE per global address.
 *   - libm and division are straight-line per sample with one shared rare branch per K samples
 *     (tpgx_math.h), so the K chains of a lane overlap; coefficients come from constant memory.
 */
#include <cuda_runtime.h>

#include <type_traits>

#include "dev_ops.cuh"
#include "kernels.cuh"

/* ------------------------------------------------------------------------------------------------
 * small PTX helpers: mbarrier + 1-D bulk TMA copy (cp.async.bulk -> SASS UBLKCP)                  */
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

#define K1_SLOTS 8
#define K1_CODE_QUADS 32 /* TPGX_K1_CQ constant quads + TPGX_K1_CAPL lines; two such buffers per warp */

/* Shared-memory accesses of the interpreter use explicit 32-bit shared addresses: with generic
 * pointers the compiler re-derived the shared window base (S2UR + ULEA + IADD3) on every line. */
__device__ __forceinline__ void lds_f64x2(uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <int K> __device__ __forceinline__ void ld_regs(uint32_t p, double (&v)[K]) {
    lds_f64x2(p, v[0], v[1]);
    if (K == 4) lds_f64x2(p + 512, v[2], v[3]);
}
template <int K> __device__ __forceinline__ void st_regs(uint32_t p, const double (&v)[K]) {
    sts_f64x2(p, v[0], v[1]);
    if (K == 4) sts_f64x2(p + 512, v[2], v[3]);
}
/* pixel operand address: a shared-memory address (tile staged in shared memory) or a block-uniform global base plus a
   32-bit per-thread offset (tile left in global memory) */
template <bool TILE_SMEM> struct px_addr;
template <> struct px_addr<true> { uint32_t a; };
template <> struct px_addr<false> { const uint8_t* base; uint32_t idx; };
/* K consecutive uint8 pixels of the lane -> packed word */
template <int K> __device__ __forceinline__ uint32_t ld_px_word(px_addr<true> p) { return K == 2 ? lds_u16(p.a) : lds_u32(p.a); }
/* same, tile left in global memory (read-only path, L1/L2) */
template <int K> __device__ __forceinline__ uint32_t ld_px_word(px_addr<false> p) {
    return K == 2 ? (uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p.base) + p.idx) : __ldg(reinterpret_cast<const unsigned int*>(p.base) + p.idx);
}
template <int K, class Addr> __device__ __forceinline__ void ld_px(Addr p, uint32_t (&v)[K]) {
    const uint32_t w = ld_px_word<K>(p);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = __byte_perm(w, 0, 0x4440 + k); /* byte k, zero-extended: one PRMT */
}
/* uint8 -> double is exact; cvt.rn.f64.u8 of (w >> 8k) compiles to ONE I2F.F64.U8 with a byte selector (R.Bk) per sample
   (the C++ cast goes through PRMT + I2F.F64.U32) */
template <int K, class Addr> __device__ __forceinline__ void ld_px_f64(Addr p, double (&v)[K]) {
    const uint32_t w = ld_px_word<K>(p);
#pragma unroll
    for (int k = 0; k < K; k++) asm("cvt.rn.f64.u8 %0, %1;" : "=d"(v[k]) : "r"(w >> (8 * k)));
}

/* exp / ln tables of tpgx_math.h read from the CTA's shared-memory copy (K1 leaves almost no L1, so the global
   tables would be fetched from L2 on most lookups) */
#define K1_TAB_BYTES (256 * 8 + 512 * 8)
struct tab_shared {
    uint32_t exp_addr, log_addr;
    __device__ __forceinline__ void exp_entry(int i, double& tail, double& T) const { lds_f64x2(exp_addr + (uint32_t)i * 16u, tail, T); }
    __device__ __forceinline__ void log_entry(int i, double& invc, double& logc_hi, double& logc_lo) const {
        lds_f64x2(log_addr + (uint32_t)i * 32u, invc, logc_hi);
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(logc_lo) : "r"(log_addr + (uint32_t)i * 32u + 16u));
    }
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int K, int NW, bool EXT, bool TEAM, bool TILE_SMEM>
__global__ void __launch_bounds__(NW * 32, 1) tpgx_bid_kernel(const tpgx_bid_params p) {
    constexpr int TS = 32 * K;
    static_assert(K == 2 || K == 4, "lane owns 2 or 4 samples");
    extern __shared__ __align__(128) uint8_t smem[];
    /* TILE_SMEM: the sample tile is staged in shared memory by one bulk TMA copy.  Otherwise pixel operands are read
       from the tile in global memory (L1/L2) and all of shared memory goes to TPG registers, i.e. to more warps. */
    const int tile_bytes = (p.hw + 1) * TS; /* + the all-zero row */
    uint8_t* tile = smem;
    uint8_t* regs = smem + (TILE_SMEM ? ((tile_bytes + 127) & ~127) : 0);
    uint4* codeb = reinterpret_cast<uint4*>(regs + NW * K1_SLOTS * TS * 8);
    uint64_t* bar = reinterpret_cast<uint64_t*>(codeb + NW * 2 * K1_CODE_QUADS);
    int* s_next = reinterpret_cast<int*>(bar + 1);
    double* s_tab = reinterpret_cast<double*>(bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        *s_next = 0;
    }
    for (int i = threadIdx.x; i < 256 + 512; i += NW * 32)
        s_tab[i] = i < 256 ? tpgx_math::tpgx_exp_tab_dev[i] : tpgx_math::tpgx_log_tab_dev[i - 256];
    const tab_shared tabs = {smem_u32(s_tab), smem_u32(s_tab + 256)};
    __syncthreads();
    if (TILE_SMEM && threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)tile_bytes);
        tma_load_1d(tile, p.tiles + (size_t)blockIdx.y * tile_bytes, (uint32_t)tile_bytes, bar);
    }
    const uint32_t regs_lane = smem_u32(regs) + warp * (K1_SLOTS * TS * 8) + lane * 16;
    const uint32_t tile_lane = smem_u32(tile) + lane * K;
    /* tile in global memory: block-uniform base + 32-bit index of the lane's packed word (K pixels) */
    const uint8_t* __restrict__ const gtile = p.tiles + (size_t)blockIdx.y * tile_bytes;
    uint4* const mycode = codeb + warp * (2 * K1_CODE_QUADS);

    const int chunk_begin = blockIdx.x * p.units_per_chunk; /* chunks vary fastest: CTAs resident together share tiles (L2) */
    const int chunk_n = min(p.units_per_chunk, p.nb_units - chunk_begin);
    /* this tile's Sobel planes: [which][window][TS samples] doubles; the lane's K samples are consecutive */
    const long long plane_stride = (long long)p.nwin * TS;
    const double* __restrict__ const sobel_tile = p.sobel + (size_t)blockIdx.y * 2 * plane_stride;
    const uint32_t sobel_lane = lane * (K / 2); /* double2 index */
    const uint4* __restrict__ stream = p.stream;

    /* next unclaimed program of the chunk: one shared-memory atomic by lane 0 (plain atom.shared — the
       compiler's warp-aggregation wrapper around atomicAdd costs ~35 instructions here), broadcast by shuffle */
    const uint32_t s_next_addr = smem_u32(s_next);
    auto claim = [&]() {
        int i = 0;
        if (lane == 0) asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(i) : "r"(s_next_addr) : "memory");
        return __shfl_sync(0xffffffffu, i, 0);
    };
    /* A work unit is a program (TEAM = false: its bid row is written out) or a whole team (TEAM = true: the
       programs of its outgoing edges run back to back, the winning edge per sample is kept in registers with
       the reference's order-dependent tie-break and only its destination is written: [UP] evaluateTeam fused
       into the interpreter, 4 bytes per (team, sample) instead of 8 bytes per (edge, sample)).
       unit -> {first row, number of rows}; rows index p.desc.                                              */
    auto load_unit = [&](int u) {
        if (u >= chunk_n) return make_int2(0, 0);
        return TEAM ? __ldg(p.team + chunk_begin + u) : make_int2(chunk_begin + u, 1);
    };
    auto load_desc = [&](int row, bool valid) { return valid ? __ldg(p.desc + row) : make_int2(0, 0); };
    /* asynchronous copy of a program's constants + first TPGX_K1_CAPL lines into a code buffer */
    auto stage = [&](uint4* buf, const int2 d) {
        if (lane < TPGX_K1_CQ + min(d.y, TPGX_K1_CAPL)) cp_async16(buf + lane, stream + d.x + lane);
        cp_async_commit();
    };
    int u0 = claim();
    int2 un0 = load_unit(u0);
    int row = un0.x, row_end = un0.x + un0.y;
    int2 d0 = load_desc(row, u0 < chunk_n);
    int cur = 0;
    stage(mycode, d0);
    if (TILE_SMEM) mbar_wait(bar, 0);
    double best[K];
    int best_dst[K];
#pragma unroll
    for (int k = 0; k < K; k++) { best[k] = 0.0; best_dst[k] = 0; }

    while (u0 < chunk_n) {
        /* the row after this one: next edge of the team, or the first row of the next claimed unit */
        int u1 = u0, nrow = row + 1, nrow_end = row_end;
        if (nrow >= row_end) {
            u1 = claim();
            const int2 un1 = load_unit(u1);
            nrow = un1.x; nrow_end = un1.x + un1.y;
        }
        const int2 d1 = load_desc(nrow, u1 < chunk_n);
        cp_async_wait_all();
        __syncwarp();
        uint4* const code = mycode + cur * K1_CODE_QUADS;
        stage(mycode + (cur ^ 1) * K1_CODE_QUADS, d1); /* next program's code: in flight while this one runs */

        const int n = d0.y;
        const uint8_t* const cbase = reinterpret_cast<const uint8_t*>(code); /* program constants at the head of the buffer */
        for (int done = 0; done < n;) {
            const int m = min(n - done, TPGX_K1_CAPL);
            if (done > 0) { /* programs longer than one buffer: stage the next TPGX_K1_CAPL lines synchronously */
                __syncwarp();
                if (lane < m) code[TPGX_K1_CQ + lane] = __ldg(stream + d0.x + TPGX_K1_CQ + done + lane);
                __syncwarp();
            }
            /* no trip counter: the encoder flags the last line of every staged block (TPGX_K1H_LAST), so the loop costs a
               pointer increment, one bit test and a branch (m >= 1 here) */
            uint32_t lp = smem_u32(code + TPGX_K1_CQ);
            uint32_t last;
#pragma unroll 1
            do {
                uint4 ln;                               /* broadcast LDS.128: handler, operand offsets, destination offset */
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ln.x), "=r"(ln.y), "=r"(ln.z), "=r"(ln.w) : "r"(lp));
                lp += 16u;
                last = ln.x & TPGX_K1H_LAST;
                const uint32_t ra = regs_lane + ln.y;  /* operand as a register slot */
                const uint32_t rb = regs_lane + ln.z;
                px_addr<TILE_SMEM> pa, pb;              /* operand as a pixel */
                if constexpr (TILE_SMEM) { pa.a = tile_lane + ln.y * K; pb.a = tile_lane + ln.z * K; }
                else { pa.base = pb.base = gtile; pa.idx = lane + ln.y; pb.idx = lane + ln.z; }
                const uint32_t rd = regs_lane + ln.w;  /* destination register slot */
                /* dispatch on single bits of the handler id; every handler loads, computes and stores by
                   itself, so no values merge afterwards */
                const uint32_t h = ln.x;
                const uint32_t op = (h >> 2) & 15u;
                double a[K], b[K], r[K];
                if (!(h & TPGX_K1H_HEAVY)) {
                    /* ---- cheap class: operands by kind, then one IEEE operation ---- */
                    if (h & 1u) ld_px_f64<K>(pa, a); else ld_regs<K>(ra, a);
                    if (EXT && op >= TPGX_K1C_EQ0) {
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const double x = a[k];
                            r[k] = op == TPGX_K1C_EQ0 ? ((x == 0.0) ? 10.0 : 0.0) : op == TPGX_K1C_EQM1 ? ((x == -1.0) ? 10.0 : 0.0)
                                 : op == TPGX_K1C_EQ1 ? ((x == 1.0) ? 10.0 : 0.0) : op == TPGX_K1C_GE15 ? ((x >= 15.0) ? 10.0 : 0.0) : TPGX_PI;
                        }
                    } else {
                        if (h & 2u) ld_px_f64<K>(pb, b); else ld_regs<K>(rb, b);
                        if (op == TPGX_K1C_SUB) {
#pragma unroll
                            for (int k = 0; k < K; k++) r[k] = a[k] - b[k];
                        } else if (op == TPGX_K1C_ADD) {
#pragma unroll
                            for (int k = 0; k < K; k++) r[k] = a[k] + b[k];
                        } else if (op == TPGX_K1C_MUL) {
#pragma unroll
                            for (int k = 0; k < K; k++) r[k] = a[k] * b[k];
                        } else if (!EXT || op == TPGX_K1C_MAX) {
#pragma unroll
                            for (int k = 0; k < K; k++) r[k] = (a[k] < b[k]) ? b[k] : a[k];
                        } else {
#pragma unroll
                            for (int k = 0; k < K; k++) r[k] = (a[k] < b[k]) ? -a[k] : a[k];
                        }
                    }
                    st_regs<K>(rd, r);
                } else if (op == TPGX_K1X_SOBEL_MAGN || op == TPGX_K1X_SOBEL_DIR) {
                    /* window instructions read only the image: their values were computed once per upload
                       (K0b, tpgx_sobel_planes_kernel) and are fetched as one coalesced 16-byte load per lane */
                    const double2* src = reinterpret_cast<const double2*>(sobel_tile)
                                       + (uint32_t)(sobel_lane + (op == TPGX_K1X_SOBEL_DIR ? (uint32_t)(plane_stride >> 1) : 0u) + ln.y);
#pragma unroll
                    for (int k = 0; k < K; k += 2) {
                        const double2 v = __ldg(src + (k >> 1));
                        r[k] = v.x; r[k + 1] = v.y;
                    }
                    st_regs<K>(rd, r);
                } else {
                    /* ---- heavy class: division and libm ---- */
                    if ((op == TPGX_K1X_EXP || op == TPGX_K1X_LOG) && (h & 1u)) {
                        /* exp / ln of a PIXEL: a uint8 has 256 possible values, so the result is memoised in a
                           table filled once per context by the same tpgx_math routines (bit-identical) */
                        const uint32_t w = ld_px_word<K>(pa);
                        const double* __restrict__ t = p.pixel_lut + (op == TPGX_K1X_LOG ? 256 : 0);
#pragma unroll
                        for (int k = 0; k < K; k++) r[k] = __ldg(t + __byte_perm(w, 0, 0x4440 + k));
                        st_regs<K>(rd, r);
                        continue;
                    }
                    if (h & 1u) ld_px_f64<K>(pa, a); else ld_regs<K>(ra, a);
                    if (op == TPGX_K1X_EXP) {
                        tpgx_math::exp_k<K>(a, r, tabs);
                    } else if (op == TPGX_K1X_LOG) {
                        tpgx_math::log_k<K>(a, r, tabs);
                    } else if (op == TPGX_K1X_MULC) {
                        const double c = *reinterpret_cast<const double*>(cbase + ln.z);
#pragma unroll
                        for (int k = 0; k < K; k++) a[k] = a[k] * c;
                        tpgx_math::xdiv10_k<K>(a, r);                               /* (a * c) / 10.0 */
                    } else if (op == TPGX_K1X_DIV) {
                        if (h & 2u) ld_px_f64<K>(pb, b); else ld_regs<K>(rb, b);
                        tpgx_math::xdiv_k<K>(a, b, r);
                    } else if (EXT) {
                        if (op == TPGX_K1X_FMODG) {
                            if (h & 2u) ld_px_f64<K>(pb, b); else ld_regs<K>(rb, b);
#pragma unroll
                            for (int k = 0; k < K; k++) r[k] = (b[k] != 0.0) ? fmod(a[k], b[k]) : DBL_MIN;
                        } else {
#pragma unroll
                            for (int k = 0; k < K; k++)
                                r[k] = op == TPGX_K1X_COS ? tpgx_math::cos(a[k]) : op == TPGX_K1X_SIN ? tpgx_math::sin(a[k]) : tpgx_math::tan(a[k]);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < K; k++) r[k] = 0.0;
                    }
                    st_regs<K>(rd, r);
                }
            } while (!last);
            done += m;
        }
        /* the last effective line of a program writes register 0 (that is what makes it effective);
           an empty program leaves the 0.0 of [UP] executeProgram's register reset */
        double res[K];
        if (n > 0) ld_regs<K>(regs_lane, res);
        else {
#pragma unroll
            for (int k = 0; k < K; k++) res[k] = 0.0;
        }
        /* [UP] TPGExecutionEngine::evaluateEdge: NaN bid -> -inf */
#pragma unroll
        for (int k = 0; k < K; k++)
            if (res[k] != res[k]) res[k] = __longlong_as_double(0xfff0000000000000LL);
        if (!TEAM) {
            double* out = p.bid + (size_t)row * p.row_stride + (size_t)blockIdx.y * TS + lane * K;
#pragma unroll
            for (int k = 0; k < K; k += 2) *reinterpret_cast<double2*>(out + k) = make_double2(res[k], res[k + 1]);
        } else {
            /* [UP] evaluateTeam: the first edge is the initial best; a later edge replaces it if
               bid >= best (last-max) / bid > best (first-max) */
            const int dst = __ldg(p.row_dst + row);
            const bool first = (row == row_end - (TEAM ? un0.y : 1));
#pragma unroll
            for (int k = 0; k < K; k++) {
                const bool take = first | (p.tie_break == TPGX_TIE_LAST_MAX ? (res[k] >= best[k]) : (res[k] > best[k]));
                best[k] = take ? res[k] : best[k];
                best_dst[k] = take ? dst : best_dst[k];
            }
            if (row + 1 == row_end) {
                int* out = p.decision + (size_t)(chunk_begin + u0) * p.row_stride + (size_t)blockIdx.y * TS + lane * K;
#pragma unroll
                for (int k = 0; k < K; k += 2) *reinterpret_cast<int2*>(out + k) = make_int2(best_dst[k], best_dst[k + 1]);
            }
        }

        if (u1 != u0) un0 = make_int2(nrow, nrow_end - nrow);
        u0 = u1; row = nrow; row_end = nrow_end; d0 = d1; cur ^= 1;
    }
    cp_async_wait_all();
}

struct bid_variant { int K, NW; bool tile_smem; };
/* variant 0 is the default: 4 samples per lane amortise the per-line work (fetch, dispatch, addressing) over 128 samples;
   the tile stays in global memory (pixel operands through L1/L2) so that shared memory holds the TPG registers of 24 warps
   (9 KB each incl. code buffers; 25 would fit but leave 72 registers per thread) at 80 registers per thread. */
static const bid_variant g_variants[] = {{4, 24, false}, {2, 32, true}, {2, 28, true}, {2, 24, true}, {4, 12, true},
                                         {4, 16, false}, {2, 32, false}, {4, 20, false}, {4, 22, false}};
static const int g_nb_variants = sizeof(g_variants) / sizeof(g_variants[0]);
int tpgx_bid_warps(int variant) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    return g_variants[variant].NW;
}

int tpgx_bid_tile_samples(int variant) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    return 32 * g_variants[variant].K;
}
size_t tpgx_bid_smem_bytes(int variant, int hw) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    const bid_variant& v = g_variants[variant];
    const int ts = 32 * v.K;
    const size_t tile = v.tile_smem ? (((size_t)(hw + 1) * ts + 127) & ~(size_t)127) : 0;
    return tile + (size_t)v.NW * K1_SLOTS * ts * 8 + (size_t)v.NW * 2 * K1_CODE_QUADS * 16 + 16 + K1_TAB_BYTES;
}

template <int K, int NW, bool EXT, bool TEAM, bool TS_>
static cudaError_t launch_bid_k(const tpgx_bid_params& p, dim3 grid, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(tpgx_bid_kernel<K, NW, EXT, TEAM, TS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    tpgx_bid_kernel<K, NW, EXT, TEAM, TS_><<<grid, NW * 32, smem, st>>>(p); TPGX_COUNT_LAUNCH();
    return cudaGetLastError();
}
template <int K, int NW, bool TS_>
static cudaError_t launch_bid_t(const tpgx_bid_params& p, int nb_tiles, size_t smem, bool ext, bool team, cudaStream_t st) {
    const int chunks = (p.nb_units + p.units_per_chunk - 1) / p.units_per_chunk;
    dim3 grid(chunks, nb_tiles);
    if (ext) return team ? launch_bid_k<K, NW, true, true, TS_>(p, grid, smem, st) : launch_bid_k<K, NW, true, false, TS_>(p, grid, smem, st);
    return team ? launch_bid_k<K, NW, false, true, TS_>(p, grid, smem, st) : launch_bid_k<K, NW, false, false, TS_>(p, grid, smem, st);
}

cudaError_t tpgx_launch_bid(const tpgx_bid_params& p, int nb_tiles, int variant, bool ext, bool team, cudaStream_t st) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    const size_t smem = tpgx_bid_smem_bytes(variant, p.hw);
    switch (variant) {
    case 1: return launch_bid_t<2, 32, true>(p, nb_tiles, smem, ext, team, st);
    case 2: return launch_bid_t<2, 28, true>(p, nb_tiles, smem, ext, team, st);
    case 3: return launch_bid_t<2, 24, true>(p, nb_tiles, smem, ext, team, st);
    case 4: return launch_bid_t<4, 12, true>(p, nb_tiles, smem, ext, team, st);
    case 5: return launch_bid_t<4, 16, false>(p, nb_tiles, smem, ext, team, st);
    case 6: return launch_bid_t<2, 32, false>(p, nb_tiles, smem, ext, team, st);
    case 7: return launch_bid_t<4, 20, false>(p, nb_tiles, smem, ext, team, st);
    case 8: return launch_bid_t<4, 22, false>(p, nb_tiles, smem, ext, team, st);
    default: return launch_bid_t<4, 24, false>(p, nb_tiles, smem, ext, team, st);
    }
}
