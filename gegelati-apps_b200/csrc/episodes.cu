/*
 * episodes.cu — batched lock-step restatements of the apps' sequential LearningEnvironments and
 * the episode loop of [UP] LearningAgent::evaluateJob / AdversarialLearningAgent::evaluateJob.
 *
 * warp = up to 4 episodes (job, iteration), each on a power-of-two group of lanes at least as wide as the largest
 * team of the graph (5 edges -> 8 lanes -> 4 episodes per warp).  Episodes are sequential (<= 1000 dependent steps)
 * and every episode follows its own path through the graph, so the first version (thread = episode) ran 32
 * different instruction streams per warp, serialised.  Now every lane of a group carries a copy of the (tiny)
 * environment state and steps it redundantly — uniform control flow — while the team evaluation of [UP]
 * evaluateTeam is spread over the lanes: lane l of the group interprets the program of the team's l-th outgoing
 * edge with the per-thread executor (dev_ops.cuh) and a reduction inside the group picks the winner with the
 * reference's order-dependent tie-break.  Consecutive episodes share a warp; the iterations of one job are
 * consecutive, start from the same root and therefore mostly run the SAME programs on different states — the
 * divergent per-lane interpreter then costs little more for four episodes than for one (the second version ran
 * one episode per warp with 5.8 of 32 lanes active).  A second kernel averages the iteration scores in
 * iteration order (the reference's sequential `result += getScore()`).
 *
 * Environment semantics, line by line: SURVEY.md Appendix B;
 *   gridworld  [REF] gridworld/src/gridworld.cpp:3-83
 *   stickgame  [REF] stickgame/src/Learn/stickGameAdversarial.cpp:38-171
 *   tictactoe  [REF] tic-tac-toe/src/Learn/TicTacToe.cpp:33-204
 *   pendulum   [REF] pendulum/src/Learn/pendulum.cpp:42-141
 */
#include <random>

#include "dev_ops.cuh"
#include "episodes.cuh"
#include "kernels.cuh"

void tpgx_rng_fill_raw(uint64_t seed, int n, uint64_t* out) {
    std::mt19937_64 g(seed);
    for (int i = 0; i < n; i++) out[i] = g();
}

namespace {

struct ep_counters { unsigned int eval, team, prog, line; };

/* candidate edge of a team evaluation: the reference scans edges in list order and replaces the best when
 * bid >= best (last-max) or bid > best (first-max), i.e. it ends on the maximal bid, ties going to the
 * LAST / FIRST such edge — an order-free rule, so a warp reduction reproduces it exactly. */
struct ep_cand { double bid; int e, dst; };
__device__ __forceinline__ bool ep_better(const ep_cand& a, const ep_cand& b, int tie_break) { /* a beats b; e < 0 = no candidate */
    if (a.e < 0) return false;
    if (b.e < 0) return true;
    if (a.bid != b.bid) return a.bid > b.bid;
    return tie_break == TPGX_TIE_LAST_MAX ? a.e > b.e : a.e < b.e;
}

/* recorded executions of the episode still to resolve (warp-uniform): next entry, end, executions so far */
struct ep_hits { int pos, end, base, step; };

__device__ void ep_record_obs(const tpgx_episode_params& p, int slot, const tpgx_obs_view& o) {
    double* out = p.rec_obs + (size_t)slot * p.obs_len;
    int q = 0;
    for (int s = 0; s < 2; s++)
        for (int i = 0; i < o.width[s] && q < p.obs_len; i++) out[q++] = o.f64[s] ? o.f64[s][i] : (double)o.i32[s][i];
    for (; q < p.obs_len; q++) out[q] = 0.0;
}

/* lane layout of a warp: `width` lanes per episode group (power of two), `sl` = lane inside its group */
struct ep_group { int width, sl; unsigned mask; };

/* [UP] executeFromRoot on one observation per lane group; warp-synchronous: every lane of the warp calls it, groups
   with walk == false only take part in the shuffles.  Every lane of a walking group returns the action id, or -1
   (error bits in err). */
__device__ int ep_traverse(const tpgx_episode_params& p, int root, const tpgx_obs_view& o, ep_counters& c, int& err, const ep_group g, bool walk, ep_hits& h) {
    int visited[TPGX_MAX_PATH];
    int depth = 0, cur = root, action = -1;
    while (__any_sync(0xffffffffu, walk)) {
        int eb = 0, ee = 0;
        if (walk) {
            if (depth >= TPGX_MAX_PATH) { err |= TPGX_DEV_PATH_TOO_LONG; walk = false; }
            else {
                visited[depth++] = cur;
                if (g.sl == 0) c.team++;
                eb = __ldg(p.team_edge_offset + cur); ee = __ldg(p.team_edge_offset + cur + 1);
            }
        }
        ep_cand best = {0.0, -1, 0};
        for (int e0 = eb; __any_sync(0xffffffffu, walk && e0 < ee); e0 += g.width) {
            ep_cand mine = {0.0, -1, 0};
            const int e = e0 + g.sl;
            int prog = -1;
            if (walk && e < ee) {
                const int dst = __ldg(p.edge_destination + e);
                bool seen = false;
                if (dst >= 0)
                    for (int v = 0; v < depth; v++) seen |= (visited[v] == dst);
                if (!seen) { /* admissible edge: run its program on this lane */
                    prog = __ldg(p.edge_program + e);
                    const int cb = __ldg(p.prog_offset + prog), n = __ldg(p.prog_offset + prog + 1) - cb;
                    double bid = tpgx_run_program(p.code + cb, n, p.constants + (size_t)prog * p.nb_constants, o);
                    if (bid != bid) bid = __longlong_as_double(0xfff0000000000000LL);
                    c.prog++;
                    c.line += (unsigned int)n;
                    mine.bid = bid; mine.e = e; mine.dst = dst;
                }
            }
            __syncwarp();
            if (p.hit_offset) { /* archive pass: the reference executes the admissible edges in list order = lane order */
                const unsigned adm = __ballot_sync(0xffffffffu, prog >= 0) & g.mask;
                const int nadm = __popc(adm), rank = __popc(adm & ((1u << (threadIdx.x & 31)) - 1u));
                while (h.pos < h.end) {
                    const int x = __ldg(p.hit_exec + h.pos) - h.base;
                    if (x >= nadm) break;
                    const int slot = __ldg(p.hit_slot + h.pos);
                    if (prog >= 0 && rank == x) { p.rec_program[slot] = prog; p.rec_bid[slot] = mine.bid; p.rec_step[slot] = h.step; }
                    if (g.sl == 0 && p.rec_obs) ep_record_obs(p, slot, o);
                    h.pos++;
                }
                h.base += nadm;
            }
            for (int off = g.width >> 1; off > 0; off >>= 1) { /* stays inside the group: off < width */
                ep_cand other;
                other.bid = __shfl_xor_sync(0xffffffffu, mine.bid, off);
                other.e = __shfl_xor_sync(0xffffffffu, mine.e, off);
                other.dst = __shfl_xor_sync(0xffffffffu, mine.dst, off);
                if (ep_better(other, mine, p.tie_break)) mine = other;
            }
            if (ep_better(mine, best, p.tie_break)) best = mine;
        }
        if (walk) {
            if (best.e < 0) { err |= TPGX_DEV_NO_EDGE; walk = false; }
            else if (best.dst < 0) { action = -1 - best.dst; walk = false; }
            else cur = best.dst;
        }
    }
    return action;
}

/* ---------------------------------------------------------------- gridworld */
struct GridWorld {
    int x, y, terminated;
    double total, st[2];
    __device__ static int tile(int yy, int xx) {
        const int g[3][4] = {{0, 0, 0, 2}, {0, 0, 3, 3}, {0, 0, 0, 1}};
        return g[yy][xx];
    }
    __device__ static bool available(int px, int py) {
        if (px == 4 || px == -1) return false;
        if (py == 3 || py == -1) return false;
        return tile(py, px) != 3;
    }
    __device__ void reset(const tpgx_episode_params&, tpgx_rng_cursor&) { x = 0; y = 0; terminated = 0; total = 0.0; st[0] = 0.0; st[1] = 0.0; }
    __device__ void view(tpgx_obs_view& o) const { o.f64[0] = st; o.f64[1] = nullptr; o.i32[0] = nullptr; o.i32[1] = nullptr; o.width[0] = 2; o.width[1] = 0; }
    __device__ bool do_action(const tpgx_episode_params&, int a, tpgx_rng_cursor&) {
        switch (a) {
        case 0: if (available(x - 1, y)) x--; break;
        case 1: if (available(x, y + 1)) y++; break;
        case 2: if (available(x + 1, y)) x++; break;
        case 3: if (available(x, y - 1)) y--; break;
        }
        double reward = -1;
        if (tile(y, x) == 1) { terminated = 1; reward = 100; }
        else if (tile(y, x) == 2) { terminated = 1; reward = -100; }
        total += reward;
        st[0] = x; st[1] = y;
        return true;
    }
    __device__ bool is_terminal() const { return terminated != 0; }
    __device__ double score(int) const { return total; }
    __device__ void summary(double* o) const { o[0] = st[0]; o[1] = st[1]; o[2] = total; o[3] = terminated ? 1.0 : 0.0; }
};

/* ---------------------------------------------------------------- stick game */
struct StickGame {
    int32_t sticks[1], hints[4];
    int first_won, forb1, forb2, turn, second_random;
    __device__ void reset(const tpgx_episode_params& p, tpgx_rng_cursor&) {
        hints[0] = 1; hints[1] = 2; hints[2] = 3; hints[3] = 4;
        sticks[0] = 21; first_won = 0; forb1 = 0; forb2 = 0; turn = 0; second_random = p.second_player_random;
    }
    __device__ void view(tpgx_obs_view& o) const { o.i32[0] = hints; o.i32[1] = sticks; o.f64[0] = nullptr; o.f64[1] = nullptr; o.width[0] = 4; o.width[1] = 1; }
    __device__ bool is_terminal() const { return sticks[0] == 0; }
    __device__ void random_play(tpgx_rng_cursor& rng) {
        if (!is_terminal()) {
            int s = sticks[0];
            /* short-circuit order matters: the second draw only happens in the else branch */
            if (tpgx_rng_u64(rng, 0, 3) != 0 && (s - 1) % 4 != 0) s -= (s - 1) % 4;
            else s -= (int)tpgx_rng_u64(rng, 1, (uint64_t)min(s, 3));
            sticks[0] = s;
            if (s == 0) first_won = ((!turn) % 2 == 0);
            turn++;
        }
    }
    __device__ bool do_action(const tpgx_episode_params&, int a, tpgx_rng_cursor& rng) {
        if (a < 0 || a >= 3) return false; /* [REF] stickGameAdversarial.cpp:61: LearningEnvironment::doAction throws past nbActions = 3 */
        if (!is_terminal()) {
            const bool first = turn % 2 == 0;
            int s = sticks[0];
            if (((double)a + 1) > s) {
                if (first) forb1 = 1; else forb2 = 1;
                sticks[0] = 0;
                first_won = !first;
                return true;
            }
            s -= (a + 1);
            sticks[0] = s;
            if (s == 0) first_won = !first;
            turn++;
            if (second_random) random_play(rng);
        }
        return true;
    }
    __device__ double score(int player) const {
        double s1, s2;
        if (first_won) { s1 = 1.0; s2 = forb2 ? -1.0 : 0.0; }
        else { s2 = 1.0; s1 = forb1 ? -1.0 : 0.0; }
        return player == 0 ? s1 : s2;
    }
    __device__ void summary(double* o) const { o[0] = sticks[0]; o[1] = turn; o[2] = first_won ? 1.0 : 0.0; o[3] = (forb1 ? 1.0 : 0.0) + (forb2 ? 2.0 : 0.0); }
};

/* ---------------------------------------------------------------- tic-tac-toe */
struct TicTacToe {
    double board[9];
    int turn, win1, win2, forb1, forb2, null_game, end, second_random;
    __device__ void reset(const tpgx_episode_params& p, tpgx_rng_cursor&) {
        for (int i = 0; i < 9; i++) board[i] = -1.0;
        turn = 0; win1 = win2 = forb1 = forb2 = null_game = end = 0; second_random = p.second_player_random;
    }
    __device__ void view(tpgx_obs_view& o) const { o.f64[0] = board; o.f64[1] = nullptr; o.i32[0] = nullptr; o.i32[1] = nullptr; o.width[0] = 9; o.width[1] = 0; }
    __device__ bool is_terminal() const { return end != 0; }
    __device__ void random_play(double sym, tpgx_rng_cursor& rng) {
        const int remaining = 9 - turn;
        int decision = (int)tpgx_rng_u64(rng, 0, (uint64_t)(remaining - 1));
        int i = -1;
        while (decision >= 0 && i < 8) { i++; if (board[i] == -1) decision--; }
        board[i] = sym;
    }
    __device__ void update_game() {
        const bool odd = turn % 2 != 0;
        const double x00 = board[0], x10 = board[1], x20 = board[2], x01 = board[3], x11 = board[4], x21 = board[5],
                     x02 = board[6], x12 = board[7], x22 = board[8];
        bool won = false;
        if (x11 != -1 && ((x01 == x11 && x11 == x21) || (x10 == x11 && x11 == x12) || (x00 == x11 && x11 == x22) || (x20 == x11 && x11 == x02))) won = true;
        else if (x22 != -1.0 && ((x02 == x12 && x12 == x22) || (x20 == x21 && x21 == x22))) won = true;
        else if (x00 != -1.0 && ((x00 == x10 && x10 == x20) || (x00 == x01 && x01 == x02))) won = true;
        if (won) { end = 1; if (odd) win1 = 1; else win2 = 1; return; }
        if (turn > 8) { null_game = 1; end = 1; }
    }
    __device__ bool do_action(const tpgx_episode_params&, int a, tpgx_rng_cursor& rng) {
        if (a < 0 || a >= 9) return false; /* [REF] TicTacToe.cpp:41,46: board.getDataAt / setDataAt(actionID) throw outside the 9 cells */
        const bool even = turn % 2 == 0;
        if (!is_terminal()) {
            const double cell = board[a];
            if (cell != -1.0) { if (even) forb1 = 1; else forb2 = 1; random_play(0, rng); }
            else board[a] = 0;
            turn++;
            update_game();
            if (is_terminal()) return true;
            if (second_random) { random_play(1, rng); turn++; update_game(); }
            else for (int i = 0; i < 9; i++) { const double c = board[i]; board[i] = c == 0 ? 1 : c == 1 ? 0 : -1; }
        }
        return true;
    }
    __device__ double score(int player) const {
        const double m1 = forb1 ? 1.0 : 0.0, m2 = forb2 ? 1.0 : 0.0;
        if (null_game) return player == 0 ? 0.5 - m1 : 0.5 - m2;
        if (win1) return player == 0 ? 1 - m1 : 0 - m2;
        return player == 0 ? 0 - m1 : 1 - m2;
    }
    __device__ void summary(double* o) const {
        double code = 0.0, w = 1.0;
        for (int i = 0; i < 9; i++) { code += (board[i] + 1.0) * w; w *= 3.0; }
        o[0] = code; o[1] = turn;
        o[2] = (win1 ? 1 : 0) + (win2 ? 2 : 0) + (forb1 ? 4 : 0) + (forb2 ? 8 : 0) + (null_game ? 16 : 0) + (end ? 32 : 0);
        o[3] = 0.0;
    }
};

/* ---------------------------------------------------------------- pendulum */
struct Pendulum {
    double* hist; /* [300] reward history of the episode, element i at hist[i * hs]: ONE copy per episode in shared memory (every thread that
                     steps the episode writes the same value) */
    int hs;
    double total, st[2];
    unsigned long long n_act;
    __device__ void reset(const tpgx_episode_params&, tpgx_rng_cursor& rng) {
        for (int i = 0; i < 300; i++) hist[i * hs] = 0.0; /* a fresh clone's history ({0.0}); never observable before 300 writes */
        st[0] = tpgx_rng_f64(rng, -TPGX_PI, TPGX_PI);
        st[1] = tpgx_rng_f64(rng, -1.0, 1.0);
        n_act = 0; total = 0.0;
    }
    __device__ void view(tpgx_obs_view& o) const { o.f64[0] = st; o.f64[1] = nullptr; o.i32[0] = nullptr; o.i32[1] = nullptr; o.width[0] = 2; o.width[1] = 0; }
    __device__ bool do_action(const tpgx_episode_params& p, int a, tpgx_rng_cursor&) {
        const int na = p.nb_pendulum_actions;
        if (a < 0 || a >= na * 2 + 1) return false;
        double cur = (a == 0) ? 0.0 : p.pendulum_actions[(a - 1) % na];
        cur = (a <= na) ? cur : -cur;
        cur *= 2.0; /* MAX_TORQUE */
        double angle = st[0], velocity = st[1];
        const double atu = fmod((angle + TPGX_PI), (2.f * TPGX_PI)) - TPGX_PI;
        const double reward = -((atu * atu) + 0.2f * (velocity * velocity) + 0.01f * (cur * cur));
        hist[(int)(n_act % 300) * hs] = reward;
        n_act++;
        total += reward;
        velocity = velocity + ((-3.0) * 9.81 / (2.0 * 1.0) * (tpgx_math::sin(angle + TPGX_PI)) + (3.f / (1.0 * 1.0 * 1.0)) * cur) * 0.05;
        velocity = fmin(fmax(velocity, -8.0), 8.0);
        angle = angle + velocity * 0.05;
        st[0] = angle; st[1] = velocity;
        return true;
    }
    __device__ bool is_terminal() const {
        bool r = false;
        if (n_act >= 300) {
            /* the reference sums all 300 rewards in index order and tests |sum / 300| < 0.1.  Every reward is <= 0
               (minus a sum of squares), so the partial sums never increase: once one is below -31 the full sum is
               too and the test is false whatever follows — same boolean, without the 300-long dependent chain */
            double acc = 0.0;
            int i = 0;
            for (; i < 300 && acc >= -31.0; i++) acc += hist[i * hs];
            r = i == 300 && fabs(acc / 300.0) < 0.1;
        }
        return r;
    }
    __device__ double score(int) const {
        if (is_terminal()) return 10.0 / tpgx_math::log(((double)n_act - 300.0 + 2.0));
        return total / (double)n_act;
    }
    __device__ void summary(double* o) const { o[0] = st[0]; o[1] = st[1]; o[2] = total; o[3] = (double)n_act; }
};

template <class Env> struct ep_scratch { __device__ static void attach(Env&, double*, int) {} };
template <> struct ep_scratch<Pendulum> { __device__ static void attach(Pendulum& e, double* s, int stride) { e.hist = s; e.hs = stride; } };

#define EP_WARPS 4
#define EP_MAX_GROUPS 4 /* episodes per warp (shared memory for the pendulum histories: EP_WARPS * EP_MAX_GROUPS * 304 doubles) */

template <class Env>
__global__ void __launch_bounds__(EP_WARPS * 32) tpgx_episode_kernel(const tpgx_episode_params p) {
    __shared__ double s_scratch[EP_WARPS][EP_MAX_GROUPS][304];
    const int NI = p.nb_iterations, K = p.roots_per_job;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = p.group_width, G = p.groups_per_warp;
    const int grp = lane / W;                                         /* lanes past the last group idle along */
    ep_group g;
    g.width = W; g.sl = lane % W;
    g.mask = (W == 32 ? 0xffffffffu : ((1u << W) - 1u)) << ((grp * W) & 31);
    const long long gid = (blockIdx.x * (long long)EP_WARPS + warp) * G + grp; /* episode = (job, iteration) */
    const long long total = (long long)p.nb_jobs * NI;
    const bool valid = grp < G && gid < total;
    const bool archive_pass = p.hit_offset != nullptr;
    ep_counters c = {0, 0, 0, 0};
    int err = 0;
    const int j = valid ? (int)(gid / NI) : 0, it = valid ? (int)(gid % NI) : 0;
    tpgx_rng_cursor rng;
    rng.raw = p.rng_raw + (size_t)it * TPGX_RNG_TABLE;
    rng.pos = 0;
    rng.exhausted = 0;
    ep_hits hits = {0, 0, 0, 0};
    if (archive_pass && valid) { hits.pos = __ldg(p.hit_offset + gid); hits.end = __ldg(p.hit_offset + gid + 1); }
    Env env;                                       /* replicated on every lane of the group, stepped identically */
    ep_scratch<Env>::attach(env, s_scratch[warp][grp < G ? grp : 0], 1);
    if (valid) env.reset(p, rng);
    __syncwarp();
    tpgx_obs_view o;
    int n = 0, turn = 0;
    bool running = valid;
    for (;;) {
        /* nothing (more) to resolve in this episode ends an archive pass early */
        running = running && !env.is_terminal() && n < p.max_actions && !(archive_pass && hits.pos >= hits.end);
        if (!__any_sync(0xffffffffu, running)) break;
        int root = 0;
        if (running) {
            env.view(o);
            root = __ldg(p.job_teams + (size_t)j * K + turn);
            hits.step = n;
        }
        const int action = ep_traverse(p, root, o, c, err, g, running, hits);
        if (running) {
            if (g.sl == 0) c.eval++;
            if (action < 0) running = false;
            else if (g.sl == 0 && !archive_pass && p.trace_steps > 0 && n < p.trace_steps) p.trace[(size_t)gid * p.trace_steps + n] = (int8_t)action;
        }
        __syncwarp();
        if (running) {
            if (!env.do_action(p, action, rng)) { err |= TPGX_DEV_BAD_ACTION; running = false; }
            else {
                turn = (turn + 1 == K) ? 0 : turn + 1;
                n++;
            }
        }
        __syncwarp();
    }
    if (valid && g.sl == 0 && !archive_pass) {
        for (int q = n; q < p.trace_steps; q++) p.trace[(size_t)gid * p.trace_steps + q] = -1;
        for (int k = 0; k < K; k++) p.episode_score[(size_t)gid * K + k] = env.score(k);
        p.episode_actions[gid] = n;
        env.summary(p.final_state + (size_t)gid * 4);
    }
    if (rng.exhausted) err |= TPGX_DEV_RNG_EXHAUSTED;
    for (int o2 = W >> 1; o2 > 0; o2 >>= 1) { /* totals of the group */
        c.eval += __shfl_xor_sync(0xffffffffu, c.eval, o2);
        c.team += __shfl_xor_sync(0xffffffffu, c.team, o2);
        c.prog += __shfl_xor_sync(0xffffffffu, c.prog, o2);
        c.line += __shfl_xor_sync(0xffffffffu, c.line, o2);
    }
    if (valid && g.sl == 0) {
        if (p.episode_progs && !archive_pass) p.episode_progs[gid] = (int32_t)c.prog;
        atomicAdd(p.counters + 0, (unsigned long long)c.eval);
        atomicAdd(p.counters + 1, (unsigned long long)c.team);
        atomicAdd(p.counters + 2, (unsigned long long)c.prog);
        atomicAdd(p.counters + 3, (unsigned long long)c.line);
    }
    if (err) atomicOr(p.status, err);
}

/* ------------------------------------------------------------------------------------------------
 * Plain evaluation pass, latency-oriented mapping: CTA = one job x up to 32 of its iterations, lane = iteration
 * (episode), warp = outgoing edge of the episode's current team.  The episodes of a job start every step at the same
 * root team, so the lanes of a warp mostly interpret the SAME program on different states (no divergence), and the
 * edges of a team run in parallel on different warps instead of one after the other on the lanes of one warp: the
 * dependent chain of a step is one program, not the sum of the team's programs.  Every warp carries a replica of the
 * episode state of its lanes and steps it identically, so all control flow — including the number of barriers — is
 * the same in every warp; the candidates (bid, edge, destination) meet in shared memory once per team evaluation.
 * Serves the plain pass and the archive resolution pass (hit_offset != nullptr); tpgx_episode_kernel above is the
 * alternative mapping (cta_warps == 0).                                                                          */
#define EP2_MAX_WARPS 16
template <class Env>
__global__ void __launch_bounds__(EP2_MAX_WARPS * 32) tpgx_episode_cta_kernel(const tpgx_episode_params p) {
    extern __shared__ double s_hist[];                     /* pendulum: [300][lanes] */
    __shared__ ep_cand s_cand[2][EP2_MAX_WARPS][32];
    __shared__ unsigned int s_prog[32], s_line[32];
    const int NI = p.nb_iterations, K = p.roots_per_job;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int j = blockIdx.x, it = blockIdx.y * 32 + lane;  /* jobs on grid.x (no 65 535 limit), chunks of 32 iterations on grid.y */
    const int lanes = min(32, NI - (int)blockIdx.y * 32);
    const bool valid = it < NI;
    const long long gid = (long long)j * NI + it;
    if (w == 0) { s_prog[lane] = 0; s_line[lane] = 0; }
    ep_counters c = {0, 0, 0, 0};
    int err = 0;
    tpgx_rng_cursor rng;
    rng.raw = p.rng_raw + (size_t)(valid ? it : 0) * TPGX_RNG_TABLE;
    rng.pos = 0;
    rng.exhausted = 0;
    const bool archive_pass = p.hit_offset != nullptr;     /* resolution pass of the archive: see tpgx_episode_params */
    ep_hits hits = {0, 0, 0, 0};                           /* replicated in every warp like the rest of the episode state */
    if (archive_pass && valid) { hits.pos = __ldg(p.hit_offset + gid); hits.end = __ldg(p.hit_offset + gid + 1); }
    Env env;                                               /* replicated in every warp, stepped identically */
    ep_scratch<Env>::attach(env, s_hist + lane, lanes);
    if (valid) env.reset(p, rng);
    __syncthreads();
    tpgx_obs_view o;
    int n = 0, turn = 0, buf = 0;
    bool running = valid;
    for (;;) {
        running = running && !env.is_terminal() && n < p.max_actions && !(archive_pass && hits.pos >= hits.end);
        if (!__any_sync(0xffffffffu, running)) break;
        int cur = 0;
        if (running) {
            env.view(o);
            cur = __ldg(p.job_teams + (size_t)j * K + turn);
            hits.step = n;
        }
        /* [UP] executeFromRoot, one team per round: warp w takes edges w, w + nw, ... of every lane's current team */
        int visited[TPGX_MAX_PATH];
        int depth = 0, action = -1;
        bool walk = running;
        while (__any_sync(0xffffffffu, walk)) {
            int eb = 0, ee = 0;
            if (walk) {
                if (depth >= TPGX_MAX_PATH) { err |= TPGX_DEV_PATH_TOO_LONG; walk = false; }
                else {
                    visited[depth++] = cur;
                    if (w == 0) c.team++;
                    eb = __ldg(p.team_edge_offset + cur); ee = __ldg(p.team_edge_offset + cur + 1);
                }
            }
            ep_cand best = {0.0, -1, 0};
            for (int e0 = eb; __any_sync(0xffffffffu, walk && e0 < ee); e0 += nw) {
                ep_cand mine = {0.0, -1, 0};
                const int e = e0 + w;
                if (walk && e < ee) {
                    const int dst = __ldg(p.edge_destination + e);
                    bool seen = false;
                    if (dst >= 0)
                        for (int v = 0; v < depth; v++) seen |= (visited[v] == dst);
                    if (!seen) { /* admissible edge */
                        const int prog = __ldg(p.edge_program + e);
                        const int cb = __ldg(p.prog_offset + prog), nl = __ldg(p.prog_offset + prog + 1) - cb;
                        double bid = tpgx_run_program(p.code + cb, nl, p.constants + (size_t)prog * p.nb_constants, o);
                        if (bid != bid) bid = __longlong_as_double(0xfff0000000000000LL);
                        c.prog++;
                        c.line += (unsigned int)nl;
                        mine.bid = bid; mine.e = e; mine.dst = dst;
                    }
                }
                s_cand[buf][w][lane] = mine;
                __syncthreads();                           /* one barrier per round: the buffers alternate */
                int nadm = 0, rank = 0;                    /* admissible edges of this round / of them, before this warp's */
                for (int ww = 0; ww < nw; ww++) {
                    const ep_cand other = s_cand[buf][ww][lane];
                    nadm += other.e >= 0;
                    rank += (other.e >= 0) & (ww < w);
                    if (ep_better(other, best, p.tie_break)) best = other;
                }
                if (archive_pass && walk) { /* the reference executes the admissible edges in list order = warp order */
                    while (hits.pos < hits.end) {
                        const int x = __ldg(p.hit_exec + hits.pos) - hits.base;
                        if (x >= nadm) break;
                        const int slot = __ldg(p.hit_slot + hits.pos);
                        if (mine.e >= 0 && rank == x) { p.rec_program[slot] = __ldg(p.edge_program + mine.e); p.rec_bid[slot] = mine.bid; p.rec_step[slot] = hits.step; }
                        if (w == 0 && p.rec_obs) ep_record_obs(p, slot, o);
                        hits.pos++;
                    }
                    hits.base += nadm;
                }
                buf ^= 1;
            }
            if (walk) {
                if (best.e < 0) { err |= TPGX_DEV_NO_EDGE; walk = false; }
                else if (best.dst < 0) { action = -1 - best.dst; walk = false; }
                else cur = best.dst;
            }
        }
        if (running) {
            if (w == 0) c.eval++;
            if (action < 0) running = false;
            else {
                if (w == 0 && !archive_pass && p.trace_steps > 0 && n < p.trace_steps) p.trace[(size_t)gid * p.trace_steps + n] = (int8_t)action;
                if (!env.do_action(p, action, rng)) { err |= TPGX_DEV_BAD_ACTION; running = false; }
                else {
                    turn = (turn + 1 == K) ? 0 : turn + 1;
                    n++;
                }
            }
        }
    }
    if (valid && w == 0 && !archive_pass) {
        for (int q = n; q < p.trace_steps; q++) p.trace[(size_t)gid * p.trace_steps + q] = -1;
        for (int k = 0; k < K; k++) p.episode_score[(size_t)gid * K + k] = env.score(k);
        p.episode_actions[gid] = n;
        env.summary(p.final_state + (size_t)gid * 4);
    }
    if (rng.exhausted) err |= TPGX_DEV_RNG_EXHAUSTED;
    if (valid && c.prog) { atomicAdd(&s_prog[lane], c.prog); atomicAdd(&s_line[lane], c.line); }
    __syncthreads();
    if (w == 0) {
        if (valid && p.episode_progs && !archive_pass) p.episode_progs[gid] = (int32_t)s_prog[lane];
        c.prog = valid ? s_prog[lane] : 0u; c.line = valid ? s_line[lane] : 0u;
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            c.eval += __shfl_xor_sync(0xffffffffu, c.eval, o2);
            c.team += __shfl_xor_sync(0xffffffffu, c.team, o2);
            c.prog += __shfl_xor_sync(0xffffffffu, c.prog, o2);
            c.line += __shfl_xor_sync(0xffffffffu, c.line, o2);
        }
        if (lane == 0) {
            atomicAdd(p.counters + 0, (unsigned long long)c.eval);
            atomicAdd(p.counters + 1, (unsigned long long)c.team);
            atomicAdd(p.counters + 2, (unsigned long long)c.prog);
            atomicAdd(p.counters + 3, (unsigned long long)c.line);
        }
    }
    if (err) atomicOr(p.status, err);
}

template <class Env>
static cudaError_t launch_cta(const tpgx_episode_params& p, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tpgx_episode_cta_kernel<Env>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dim3 grid(p.nb_jobs, (p.nb_iterations + 31) / 32);
    if (grid.y > 65535u) return cudaErrorInvalidConfiguration;
    tpgx_episode_cta_kernel<Env><<<grid, p.cta_warps * 32, smem, st>>>(p); TPGX_COUNT_LAUNCH();
    return cudaGetLastError();
}

/* [UP] evaluateJob: result += getScore() over iterations in order, then result / nbIterations */
__global__ void tpgx_job_score_kernel(const tpgx_episode_params p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = p.roots_per_job;
    if (i >= p.nb_jobs * K) return;
    const int j = i / K, k = i % K;
    double result = 0.0;
    for (int it = 0; it < p.nb_iterations; it++) result += p.episode_score[((size_t)j * p.nb_iterations + it) * K + k];
    p.job_score[i] = result / (double)p.nb_iterations;
}

} // namespace

cudaError_t tpgx_launch_episodes(const tpgx_episode_params& p, cudaStream_t st) {
    const long long total = (long long)p.nb_jobs * p.nb_iterations;
    const int threads = EP_WARPS * 32;
    const int W = p.group_width;
    if (W < 1 || W > 32 || (W & (W - 1)) || p.groups_per_warp < 1 || p.groups_per_warp > EP_MAX_GROUPS || p.groups_per_warp * W > 32) return cudaErrorInvalidValue;
    const long long per_block = (long long)EP_WARPS * p.groups_per_warp;
    const unsigned blocks = (unsigned)((total + per_block - 1) / per_block);
    if (p.cta_warps > 0) { /* one CTA per job, warp = edge, lane = iteration */
        if (p.cta_warps > EP2_MAX_WARPS) return cudaErrorInvalidValue;
        cudaError_t e;
        switch (p.env) {
        case TPGX_ENV_GRIDWORLD: e = launch_cta<GridWorld>(p, 0, st); break;
        case TPGX_ENV_STICKGAME: e = launch_cta<StickGame>(p, 0, st); break;
        case TPGX_ENV_TICTACTOE: e = launch_cta<TicTacToe>(p, 0, st); break;
        case TPGX_ENV_PENDULUM: e = launch_cta<Pendulum>(p, (size_t)300 * 8 * (p.nb_iterations < 32 ? p.nb_iterations : 32), st); break;
        default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        if (p.hit_offset) return cudaSuccess; /* archive pass: scores were produced by the counting pass */
        const int n = p.nb_jobs * p.roots_per_job;
        tpgx_job_score_kernel<<<(n + 127) / 128, 128, 0, st>>>(p); TPGX_COUNT_LAUNCH();
        return cudaGetLastError();
    }
    switch (p.env) {
    case TPGX_ENV_GRIDWORLD: tpgx_episode_kernel<GridWorld><<<blocks, threads, 0, st>>>(p); TPGX_COUNT_LAUNCH(); break;
    case TPGX_ENV_STICKGAME: tpgx_episode_kernel<StickGame><<<blocks, threads, 0, st>>>(p); TPGX_COUNT_LAUNCH(); break;
    case TPGX_ENV_TICTACTOE: tpgx_episode_kernel<TicTacToe><<<blocks, threads, 0, st>>>(p); TPGX_COUNT_LAUNCH(); break;
    case TPGX_ENV_PENDULUM: tpgx_episode_kernel<Pendulum><<<blocks, threads, 0, st>>>(p); TPGX_COUNT_LAUNCH(); break;
    default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (p.hit_offset) return cudaSuccess; /* archive pass: scores were produced by the counting pass */
    const int n = p.nb_jobs * p.roots_per_job;
    tpgx_job_score_kernel<<<(n + 127) / 128, 128, 0, st>>>(p); TPGX_COUNT_LAUNCH();
    return cudaGetLastError();
}
