/*
 * kernels.cu — sm_100a kernels of the classification (MNIST-shape) path.
 *
 *   K0 pack      : gather + transpose sample-major uint8 observations into pixel-major tiles
 *   K1 bid       : warp = one program, lane = K samples; replaces
 *                  [UP] Program::ProgramExecutionEngine::executeProgram (+ Data::*::getDataAt)
 *   K2 traverse  : thread = (root, sample); replaces [UP] TPG::TPGExecutionEngine::executeFromRoot /
 *                  evaluateTeam and [UP] ClassificationLearningEnvironment::doAction (confusion table)
 *   K3 score     : thread = root; replaces the F1 reduction of
 *                  [UP] ClassificationLearningAgent::evaluateJob
 *
 * Compiled with -fmad=false (no contraction): arithmetic keeps the reference's operation order.
 */
#include <cuda_runtime.h>

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

/* exact uint8 -> double without the conversion pipe: 2^52 + v has v in its low mantissa bits */
__device__ __forceinline__ double u8_to_f64(uint32_t v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }

/* ================================================================================================
 * K1 — bid interpreter.
 * CTA = one sample tile (TS = 32*K samples, staged once by a bulk TMA copy) x one chunk of
 * programs; each warp repeatedly claims a program of the chunk and interprets it for the 32*K
 * samples of the tile (lane = K consecutive samples), so opcode dispatch is warp-uniform.
 * TPG registers live in shared memory as [warp][9 slots][K/2][32 lanes][2] doubles (slot 8 is a
 * constant 0.0 that stands for "never written" registers): every access is a conflict-free
 * 16-byte-per-lane vector.  The line words of a program are held one per lane and broadcast by
 * shuffle, one line ahead of execution.  Lines use the K1 "fused" opcodes (tpgx_internal.h): for
 * the cheap instructions the operand kinds (register / pixel) are part of the opcode, so a line
 * costs one indirect branch instead of a chain of kind tests.                                      */
#define K1_SLOTS 9

template <int K> __device__ __forceinline__ void ld_regs(const double* p, double (&v)[K]);
template <> __device__ __forceinline__ void ld_regs<1>(const double* p, double (&v)[1]) { v[0] = *p; }
template <> __device__ __forceinline__ void ld_regs<2>(const double* p, double (&v)[2]) {
    double2 t = *reinterpret_cast<const double2*>(p); v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_regs<4>(const double* p, double (&v)[4]) {
    double2 t = *reinterpret_cast<const double2*>(p); double2 u = *reinterpret_cast<const double2*>(p + 64);
    v[0] = t.x; v[1] = t.y; v[2] = u.x; v[3] = u.y;
}
template <int K> __device__ __forceinline__ void st_regs(double* p, const double (&v)[K]);
template <> __device__ __forceinline__ void st_regs<1>(double* p, const double (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void st_regs<2>(double* p, const double (&v)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}
template <> __device__ __forceinline__ void st_regs<4>(double* p, const double (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 64) = make_double2(v[2], v[3]);
}
/* global bid row: the lane's K samples are consecutive */
template <int K> __device__ __forceinline__ void st_bid(double* p, const double (&v)[K]) {
#pragma unroll
    for (int k = 0; k < K; k += 2) {
        if (K == 1) p[0] = v[0];
        else *reinterpret_cast<double2*>(p + k) = make_double2(v[k], v[k + 1 < K ? k + 1 : k]);
    }
}
template <int K> __device__ __forceinline__ void ld_px(const uint8_t* p, uint32_t (&v)[K]);
template <> __device__ __forceinline__ void ld_px<1>(const uint8_t* p, uint32_t (&v)[1]) { v[0] = *p; }
template <> __device__ __forceinline__ void ld_px<2>(const uint8_t* p, uint32_t (&v)[2]) {
    uchar2 t = *reinterpret_cast<const uchar2*>(p); v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_px<4>(const uint8_t* p, uint32_t (&v)[4]) {
    uchar4 t = *reinterpret_cast<const uchar4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <int K> __device__ __forceinline__ void ld_px_f64(const uint8_t* p, double (&v)[K]) {
    uint32_t px[K];
    ld_px<K>(p, px);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = u8_to_f64(px[k]);
}

template <int K, int NW, int MINB, bool TRIG>
__global__ void __launch_bounds__(NW * 32, MINB) tpgx_bid_kernel(const tpgx_bid_params p) {
    constexpr int TS = 32 * K;
    extern __shared__ __align__(128) uint8_t smem[];
    const int tile_bytes = p.hw * TS;
    uint8_t* tile = smem;
    double* regs = reinterpret_cast<double*>(smem + ((tile_bytes + 127) & ~127));
    double* sconst = regs + NW * K1_SLOTS * TS;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sconst + NW * TPGX_MAX_CONSTS);
    int* s_next = reinterpret_cast<int*>(bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        *s_next = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)tile_bytes);
        tma_load_1d(tile, p.tiles + (size_t)blockIdx.x * tile_bytes, (uint32_t)tile_bytes, bar);
    }
    double* regs_lane = regs + warp * (K1_SLOTS * TS) + lane * (K == 1 ? 1 : 2);
    {
        double z[K];
#pragma unroll
        for (int k = 0; k < K; k++) z[k] = 0.0;
        st_regs<K>(regs_lane + TPGX_LOC_ZERO * TS, z);
    }
    mbar_wait(bar, 0);

    const uint8_t* tile_lane = tile + lane * K;
    double* my_const = sconst + warp * TPGX_MAX_CONSTS;
    const int chunk_begin = blockIdx.y * p.progs_per_chunk;
    const int chunk_n = min(p.progs_per_chunk, p.nb_active - chunk_begin);
    const int wrow = p.width * TS;

    for (;;) {
        int i = 0;
        if (lane == 0) i = atomicAdd(s_next, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= chunk_n) break;
        const int row = chunk_begin + i;
        const int prog = __ldg(p.active_prog + row);
        const int cb = __ldg(p.prog_offset + prog);
        const int n = __ldg(p.prog_offset + prog + 1) - cb;
        __syncwarp();
        if (lane < p.nb_constants) my_const[lane] = __ldg(p.constants + (size_t)prog * p.nb_constants + lane);
        __syncwarp();

        double res[K];
#pragma unroll
        for (int k = 0; k < K; k++) res[k] = 0.0;

        for (int base = 0; base < n; base += 32) {
            const uint32_t mine = (base + lane < n) ? __ldg(p.code + cb + base + lane) : 0u;
            const int m = min(32, n - base);
            uint32_t w = __shfl_sync(0xffffffffu, mine, 0);
#pragma unroll 1
            for (int j = 0; j < m; j++) {
                const uint32_t w_next = __shfl_sync(0xffffffffu, mine, j + 1); /* next line, one iteration ahead */
                const uint32_t la = (w >> 10) & 1023u, lb = (w >> 20) & 1023u;
                const double* ra = regs_lane + la * TS;   /* operand as a register slot */
                const double* rb = regs_lane + lb * TS;
                const uint8_t* pa = tile_lane + la * TS;  /* operand as a pixel */
                const uint8_t* pb = tile_lane + lb * TS;
                double a[K], b[K];
#define K1_RR ld_regs<K>(ra, a); ld_regs<K>(rb, b);
#define K1_PR ld_px_f64<K>(pa, a); ld_regs<K>(rb, b);
#define K1_RP ld_regs<K>(ra, a); ld_px_f64<K>(pb, b);
#define K1_PP ld_px_f64<K>(pa, a); ld_px_f64<K>(pb, b);
#define K1_R ld_regs<K>(ra, a);
#define K1_P ld_px_f64<K>(pa, a);
#define K1_EACH(expr) { _Pragma("unroll") for (int k = 0; k < K; k++) { const double x = a[k], y = b[k]; (void)x; (void)y; res[k] = (expr); } } break;
#define K1_BIN(id, expr)                                     \
    case (id) * 4 + 0: K1_RR K1_EACH(expr)                    \
    case (id) * 4 + 1: K1_PR K1_EACH(expr)                    \
    case (id) * 4 + 2: K1_RP K1_EACH(expr)                    \
    case (id) * 4 + 3: K1_PP K1_EACH(expr)
#define K1_UN(id, expr)                                      \
    case TPGX_K1_UNARY0 + (id) * 2 + 0: K1_R K1_EACH(expr)    \
    case TPGX_K1_UNARY0 + (id) * 2 + 1: K1_P K1_EACH(expr)
                switch (w & 127u) {
                K1_BIN(0, x - y)
                K1_BIN(1, x + y)
                K1_BIN(2, x * y)
                K1_BIN(3, (x < y) ? y : x)
                K1_BIN(4, x < y ? -x : x)
                K1_UN(0, (x == 0.0) ? 10.0 : 0.0)
                K1_UN(1, (x == -1.0) ? 10.0 : 0.0)
                K1_UN(2, (x == 1.0) ? 10.0 : 0.0)
                K1_UN(3, (x >= 15.0) ? 10.0 : 0.0)
                K1_UN(4, TPGX_PI)
                default: {
                    /* heavy instructions: operand kinds in bits 30/31, then one shared body */
                    const uint32_t hop = w & 127u;
                    if (hop >= TPGX_K1_SOBEL_MAGN) {
                        /* window ops in exact integer arithmetic: pixels are integers, so every partial sum
                           of the reference's double expression is an exactly representable integer */
                        uint32_t a00[K], a01[K], a02[K], a10[K], a12[K], a20[K], a21[K], a22[K];
                        ld_px<K>(pa, a00); ld_px<K>(pa + TS, a01); ld_px<K>(pa + 2 * TS, a02);
                        ld_px<K>(pa + wrow, a10); ld_px<K>(pa + wrow + 2 * TS, a12);
                        ld_px<K>(pa + 2 * wrow, a20); ld_px<K>(pa + 2 * wrow + TS, a21); ld_px<K>(pa + 2 * wrow + 2 * TS, a22);
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int gx = -(int)a00[k] + (int)a02[k] - 2 * (int)a10[k] + 2 * (int)a12[k] - (int)a20[k] + (int)a22[k];
                            const int gy = -(int)a00[k] - 2 * (int)a01[k] - (int)a02[k] + (int)a20[k] + 2 * (int)a21[k] + (int)a22[k];
                            if (hop == TPGX_K1_SOBEL_MAGN) res[k] = tpgx_math::xsqrt_nonneg((double)(gx * gx + gy * gy));
                            else res[k] = tpgx_math::atan(tpgx_math::xdiv((double)gy, (double)gx));
                        }
                        break;
                    }
                    if (w & 0x40000000u) { K1_P } else { K1_R }
                    if (hop == TPGX_K1_DIV || hop == TPGX_K1_FMODG) {
                        if (w & 0x80000000u) ld_px_f64<K>(pb, b); else ld_regs<K>(rb, b);
                    } else if (hop == TPGX_K1_MULC) {
                        const double c = my_const[lb & 7u];
#pragma unroll
                        for (int k = 0; k < K; k++) b[k] = c;
                    }
                    switch (hop) {
                    case TPGX_K1_DIV: K1_EACH(tpgx_math::xdiv(x, y))
                    case TPGX_K1_MULC: K1_EACH(tpgx_math::xdiv(x * y, 10.0))
                    case TPGX_K1_FMODG: K1_EACH(y != 0.0 ? fmod(x, y) : DBL_MIN)
                    case TPGX_K1_EXP: K1_EACH(tpgx_math::exp(x))
                    case TPGX_K1_LOG: K1_EACH(tpgx_math::log(x))
                    case TPGX_K1_COS: K1_EACH(TRIG ? tpgx_math::cos(x) : 0.0)
                    case TPGX_K1_SIN: K1_EACH(TRIG ? tpgx_math::sin(x) : 0.0)
                    case TPGX_K1_TAN: K1_EACH(TRIG ? tpgx_math::tan(x) : 0.0)
                    default: break;
                    }
                    break;
                }
                }
#undef K1_RR
#undef K1_PR
#undef K1_RP
#undef K1_PP
#undef K1_R
#undef K1_P
#undef K1_EACH
#undef K1_BIN
#undef K1_UN
                st_regs<K>(regs_lane + ((w >> 7) & 7u) * TS, res);
                w = w_next;
            }
        }
        /* [UP] TPGExecutionEngine::evaluateEdge: NaN bid -> -inf */
#pragma unroll
        for (int k = 0; k < K; k++)
            if (res[k] != res[k]) res[k] = __longlong_as_double(0xfff0000000000000LL);
        st_bid<K>(p.bid + (size_t)row * p.row_stride + (size_t)blockIdx.x * TS + lane * K, res);
    }
}

struct bid_variant { int K, NW, MINB; };
static const bid_variant g_variants[] = {{2, 32, 1}, {4, 12, 1}, {2, 16, 1}, {1, 16, 2}, {2, 12, 2}, {4, 8, 1}, {2, 24, 1}, {4, 10, 1}};
static const int g_nb_variants = sizeof(g_variants) / sizeof(g_variants[0]);

int tpgx_bid_tile_samples(int variant) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    return 32 * g_variants[variant].K;
}
size_t tpgx_bid_smem_bytes(int variant, int hw) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    const bid_variant& v = g_variants[variant];
    const int ts = 32 * v.K;
    size_t tile = ((size_t)hw * ts + 127) & ~(size_t)127;
    return tile + (size_t)v.NW * K1_SLOTS * ts * 8 + (size_t)v.NW * TPGX_MAX_CONSTS * 8 + 16;
}

template <int K, int NW, int MINB>
static cudaError_t launch_bid_t(const tpgx_bid_params& p, int nb_tiles, size_t smem, bool trig, cudaStream_t st) {
    const int chunks = (p.nb_active + p.progs_per_chunk - 1) / p.progs_per_chunk;
    dim3 grid(nb_tiles, chunks);
    cudaError_t e;
    if (trig) {
        e = cudaFuncSetAttribute(tpgx_bid_kernel<K, NW, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tpgx_bid_kernel<K, NW, MINB, true><<<grid, NW * 32, smem, st>>>(p);
    } else {
        e = cudaFuncSetAttribute(tpgx_bid_kernel<K, NW, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tpgx_bid_kernel<K, NW, MINB, false><<<grid, NW * 32, smem, st>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t tpgx_launch_bid(const tpgx_bid_params& p, int nb_tiles, int variant, bool trig, cudaStream_t st) {
    if (variant < 0 || variant >= g_nb_variants) variant = 0;
    const size_t smem = tpgx_bid_smem_bytes(variant, p.hw);
    switch (variant) {
    case 1: return launch_bid_t<4, 12, 1>(p, nb_tiles, smem, trig, st);
    case 2: return launch_bid_t<2, 16, 1>(p, nb_tiles, smem, trig, st);
    case 3: return launch_bid_t<1, 16, 2>(p, nb_tiles, smem, trig, st);
    case 4: return launch_bid_t<2, 12, 2>(p, nb_tiles, smem, trig, st);
    case 5: return launch_bid_t<4, 8, 1>(p, nb_tiles, smem, trig, st);
    case 6: return launch_bid_t<2, 24, 1>(p, nb_tiles, smem, trig, st);
    case 7: return launch_bid_t<4, 10, 1>(p, nb_tiles, smem, trig, st);
    default: return launch_bid_t<2, 32, 1>(p, nb_tiles, smem, trig, st);
    }
}

/* ================================================================================================
 * K2 — traversal.  block = (span of samples, one root); thread = one (root, sample) walk.
 * [UP] executeFromRoot: from the root team follow the best admissible edge until an action;
 * [UP] evaluateTeam: skip edges to already visited teams, first admissible edge is the initial
 * best, later edges replace it if bid >= best (last-max) / bid > best (first-max).                 */
#define TRAV_THREADS 256
#define TRAV_SPT 8

__global__ void __launch_bounds__(TRAV_THREADS) tpgx_traverse_kernel(const tpgx_trav_params p) {
    __shared__ unsigned int s_table[256];
    __shared__ unsigned long long s_cnt[4];
    const int C = p.nb_classes;
    for (int i = threadIdx.x; i < C * C; i += TRAV_THREADS) s_table[i] = 0;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();

    const int r = blockIdx.y;
    const int root = __ldg(p.roots + r);
    unsigned int c_eval = 0, c_team = 0, c_prog = 0, c_line = 0;
    int err = 0;
    const int s_begin = blockIdx.x * (TRAV_THREADS * TRAV_SPT);
    for (int it = 0; it < TRAV_SPT; it++) {
        const int s = s_begin + it * TRAV_THREADS + threadIdx.x;
        if (s >= p.nb_samples) break;
        int visited[TPGX_MAX_PATH];
        int depth = 0;
        int cur = root;
        int action = -1;
        for (;;) {
            if (depth >= TPGX_MAX_PATH) { err |= TPGX_DEV_PATH_TOO_LONG; break; }
            visited[depth++] = cur;
            c_team++;
            const int eb = __ldg(p.team_edge_offset + cur), ee = __ldg(p.team_edge_offset + cur + 1);
            bool have = false;
            int best_dst = 0;
            double best_bid = 0.0;
            for (int e = eb; e < ee; e++) {
                const int dst = __ldg(p.edge_destination + e);
                if (dst >= 0) {
                    bool seen = false;
                    for (int v = 0; v < depth; v++) seen |= (visited[v] == dst);
                    if (seen) continue;
                }
                const double bid = __ldg(p.bid + (size_t)__ldg(p.edge_row + e) * p.row_stride + s);
                c_prog++;
                c_line += (unsigned int)__ldg(p.edge_lines + e);
                const bool take = !have || (p.tie_break == TPGX_TIE_LAST_MAX ? (bid >= best_bid) : (bid > best_bid));
                if (take) { have = true; best_dst = dst; best_bid = bid; }
            }
            if (!have) { err |= TPGX_DEV_NO_EDGE; break; }
            if (best_dst < 0) { action = -1 - best_dst; break; } /* destination is an action */
            cur = best_dst;
        }
        c_eval++;
        if (action >= 0) {
            if (p.action) p.action[(size_t)r * p.action_stride + s] = (uint8_t)action;
            if (p.labels) {
                if (action < C) atomicAdd(&s_table[(int)__ldg(p.labels + s) * C + action], 1u);
                else err |= TPGX_DEV_BAD_ACTION;
            }
        }
    }
    /* counters: warp reduce then one shared atomic per warp */
    for (int o = 16; o > 0; o >>= 1) {
        c_eval += __shfl_xor_sync(0xffffffffu, c_eval, o);
        c_team += __shfl_xor_sync(0xffffffffu, c_team, o);
        c_prog += __shfl_xor_sync(0xffffffffu, c_prog, o);
        c_line += __shfl_xor_sync(0xffffffffu, c_line, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_cnt[0], (unsigned long long)c_eval);
        atomicAdd(&s_cnt[1], (unsigned long long)c_team);
        atomicAdd(&s_cnt[2], (unsigned long long)c_prog);
        atomicAdd(&s_cnt[3], (unsigned long long)c_line);
    }
    if (err) atomicOr(p.status, err);
    __syncthreads();
    if (p.labels)
        for (int i = threadIdx.x; i < C * C; i += TRAV_THREADS)
            if (s_table[i]) atomicAdd(p.confusion + (size_t)r * C * C + i, (unsigned long long)s_table[i]);
    if (threadIdx.x < 4 && p.counters) atomicAdd(p.counters + threadIdx.x, s_cnt[threadIdx.x]);
}

cudaError_t tpgx_launch_traverse(const tpgx_trav_params& p, cudaStream_t st) {
    dim3 grid((p.nb_samples + TRAV_THREADS * TRAV_SPT - 1) / (TRAV_THREADS * TRAV_SPT), p.nb_roots);
    tpgx_traverse_kernel<<<grid, TRAV_THREADS, 0, st>>>(p);
    return cudaGetLastError();
}

/* ================================================================================================
 * K3 — per-root score record {mean F1, F1[0..C)} from the confusion table, in the operation order
 * of [UP] ClassificationLearningAgent::evaluateJob: recall = TP/(TP+FN), precision = TP/(TP+FP),
 * F = TP != 0 ? 2*(P*R)/(P+R) : 0; score = (sum of F in class order starting from 0.0) / C.       */
__global__ void tpgx_score_kernel(const unsigned long long* __restrict__ confusion, int nb_roots, int C,
                                  double* __restrict__ records) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nb_roots) return;
    const unsigned long long* t = confusion + (size_t)r * C * C;
    double sum = 0.0;
    for (int c = 0; c < C; c++) {
        const unsigned long long tp = t[c * C + c];
        unsigned long long row = 0, col = 0;
        for (int k = 0; k < C; k++) { row += t[c * C + k]; col += t[k * C + c]; }
        const unsigned long long fn = row - tp, fp = col - tp;
        const double recall = (double)tp / (double)(tp + fn);
        const double precision = (double)tp / (double)(tp + fp);
        const double f = (tp != 0) ? 2 * (precision * recall) / (precision + recall) : 0.0;
        records[(size_t)r * (C + 1) + 1 + c] = f;
        sum += f;
    }
    records[(size_t)r * (C + 1)] = sum / (double)C;
}

cudaError_t tpgx_launch_score(const unsigned long long* confusion, int nb_roots, int nb_classes, double* records, cudaStream_t st) {
    tpgx_score_kernel<<<(nb_roots + 127) / 128, 128, 0, st>>>(confusion, nb_roots, nb_classes, records);
    return cudaGetLastError();
}

/* ================================================================================================
 * K0 — pack: tiles[t][px][j] = obs[index[t*TS + j]][px]; samples past the end replicate zeros.
 * Block = (32 pixels x TS samples) sub-tile staged through shared memory so both the global read
 * (along pixels) and the global write (along samples) are coalesced.                               */
__global__ void __launch_bounds__(256) tpgx_pack_kernel(const uint8_t* __restrict__ obs, const int32_t* __restrict__ index,
                                                        long long nb_samples, long long nb_obs, int hw, int ts,
                                                        uint8_t* __restrict__ tiles, const uint8_t* __restrict__ labels_in,
                                                        uint8_t* __restrict__ labels_out) {
    __shared__ uint8_t sm[128][33];
    const long long tile = blockIdx.x;
    const int p0 = blockIdx.y * 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int j = wid; j < ts; j += 8) {
        const long long s = tile * ts + j;
        uint8_t v = 0;
        if (s < nb_samples && p0 + lane < hw) {
            const long long src = index ? (long long)index[s] : s;
            v = obs[src * hw + p0 + lane];
            if (blockIdx.y == 0 && lane == 0 && labels_in) labels_out[s] = labels_in[src];
        }
        sm[j][lane] = v;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < 32 * ts; q += 256) {
        const int px = q / ts, j = q % ts;
        if (p0 + px < hw) tiles[(tile * hw + p0 + px) * ts + j] = sm[j][px];
    }
}

cudaError_t tpgx_launch_pack(const uint8_t* obs, const int32_t* sample_index, long long nb_samples, long long nb_obs, int hw,
                             int ts, uint8_t* tiles, const uint8_t* labels_in, uint8_t* labels_out, cudaStream_t st) {
    const long long nb_tiles = (nb_samples + ts - 1) / ts;
    dim3 grid((unsigned)nb_tiles, (hw + 31) / 32);
    tpgx_pack_kernel<<<grid, 256, 0, st>>>(obs, sample_index, nb_samples, nb_obs, hw, ts, tiles, labels_in, labels_out);
    return cudaGetLastError();
}

/* ================================================================================================
 * Generic executor: thread = (program, observation) on explicit f64 / i32 observations.
 * Replaces [UP] ProgramExecutionEngine::executeProgram for tpgx_execute_programs (parity tests of
 * every opcode on every source shape); NaN -> -inf applied like evaluateEdge.                      */
__global__ void tpgx_exec_generic_kernel(const uint32_t* __restrict__ code, const int32_t* __restrict__ prog_offset,
                                         const double* __restrict__ constants, int nb_constants, int nb_programs,
                                         long long count, const double* f0, const double* f1, const int32_t* i0,
                                         const int32_t* i1, int space0, int space1, int width0, int width1,
                                         double* __restrict__ bid) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (s >= count) return;
    tpgx_obs_view o;
    o.f64[0] = f0 ? f0 + s * space0 : nullptr;
    o.f64[1] = f1 ? f1 + s * space1 : nullptr;
    o.i32[0] = i0 ? i0 + s * space0 : nullptr;
    o.i32[1] = i1 ? i1 + s * space1 : nullptr;
    o.width[0] = width0;
    o.width[1] = width1;
    const int cb = prog_offset[p];
    double r = tpgx_run_program(code + cb, prog_offset[p + 1] - cb, constants + (size_t)p * nb_constants, o);
    if (r != r) r = __longlong_as_double(0xfff0000000000000LL);
    bid[(size_t)p * count + s] = r;
}

cudaError_t tpgx_launch_exec_generic(const uint32_t* code, const int32_t* prog_offset, const double* constants, int nb_constants,
                                     int nb_programs, long long count, const double* f0, const double* f1, const int32_t* i0,
                                     const int32_t* i1, int space0, int space1, int width0, int width1, double* bid,
                                     cudaStream_t st) {
    dim3 grid((unsigned)((count + 127) / 128), nb_programs);
    tpgx_exec_generic_kernel<<<grid, 128, 0, st>>>(code, prog_offset, constants, nb_constants, nb_programs, count, f0, f1, i0,
                                                   i1, space0, space1, width0, width1, bid);
    return cudaGetLastError();
}

/* ================================================================================================
 * FP64 issue-rate microbenchmark (roofline denominator): 8 independent chains per thread of
 * DADD+DMUL (mode 0, what -fmad=false code issues) or DFMA (mode 1).                              */
__global__ void __launch_bounds__(256) tpgx_fp64_peak_kernel(int mode, int iters, double* sink) {
    double a0 = threadIdx.x * 1e-3 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6,
           a7 = a0 + 0.7;
    const double m = 1.0000001, c = 1e-9;
    if (mode == 0) {
        for (int i = 0; i < iters; i++) {
            a0 = a0 * m; a1 = a1 * m; a2 = a2 * m; a3 = a3 * m; a4 = a4 * m; a5 = a5 * m; a6 = a6 * m; a7 = a7 * m;
            a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c; a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c;
        }
    } else {
        for (int i = 0; i < iters; i++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;
}

cudaError_t tpgx_launch_fp64_peak(int mode, int blocks, int iters, double* sink, cudaStream_t st) {
    tpgx_fp64_peak_kernel<<<blocks, 256, 0, st>>>(mode, iters, sink);
    return cudaGetLastError();
}
