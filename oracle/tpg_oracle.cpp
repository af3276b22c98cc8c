/*
 * tpg_oracle.cpp — CPU oracle: a plain C++17 restatement of GEGELATI's policy-evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY (see tpg_oracle.h).  "parity unpinned" for the upstream engine
 * semantics (library not vendored); pinned against the reference's in-tree sources through
 * oracle/_ref and tests/golden (instruction lambdas, environment dynamics, golden dot graphs).
 *
 * Every function cites what it follows:
 *   [REF] path:line  — a file of /root/reference (gegelati-apps)
 *   [UP]  name       — upstream GEGELATI v2.0.0 behaviour restated from its published sources
 *                      (SURVEY.md §3, Appendix F lists each such assumption U1..U14)
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -mfma (no -march=native, no fast-math) — the apps'
 * own flags ([REF] mnist/CMakeLists.txt:21-40) plus "never contract"; -mfma only inlines the
 * explicit fma() calls of tpgx_math.h.
 */
#include "tpg_oracle.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <mutex>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../gegelati-apps_b200/csrc/tpgx_math.h"
#include "../gegelati-apps_b200/csrc/tpgx_rng.h"

namespace {

/* ------------------------------------------------------------------ operand types per opcode */
enum OperandType { T_NONE = 0, T_D, T_I, T_C, T_W };

struct OpInfo {
    int nb_operands;
    OperandType type[2];
};

/* Operand signature of each device opcode = the template arguments of the reference's
 * LambdaInstruction<...> registrations ([REF] mnist/src/main.cpp:79-88,
 * pendulum/src/Learn/instructions.cpp:20-32, stickgame/src/Learn/instructions.cpp:18-23,
 * tic-tac-toe/src/Learn/main.cpp:26-34). */
OpInfo op_info(int op) {
    switch (op) {
    case TPGX_OP_SUB: case TPGX_OP_ADD: case TPGX_OP_MUL: case TPGX_OP_DIV: case TPGX_OP_MAX:
    case TPGX_OP_FMODG: case TPGX_OP_COND:
        return {2, {T_D, T_D}};
    case TPGX_OP_EXP: case TPGX_OP_LOG: case TPGX_OP_COS: case TPGX_OP_SIN: case TPGX_OP_TAN:
    case TPGX_OP_PI: case TPGX_OP_EQ0: case TPGX_OP_EQM1: case TPGX_OP_EQ1: case TPGX_OP_GE15:
        return {1, {T_D, T_NONE}};
    case TPGX_OP_MULC:
        return {2, {T_D, T_C}};
    case TPGX_OP_SUB_II:
        return {2, {T_I, T_I}};
    case TPGX_OP_CAST_I:
        return {1, {T_I, T_NONE}};
    case TPGX_OP_SOBEL_MAGN: case TPGX_OP_SOBEL_DIR:
        return {1, {T_W, T_NONE}};
    default:
        return {0, {T_NONE, T_NONE}};
    }
}

/* ------------------------------------------------------------------ math flavours */
struct Math {
    int flavour;
    double exp(double x) const { return flavour == ORC_MATH_STD ? std::exp(x) : tpgx_math::exp(x); }
    double log(double x) const { return flavour == ORC_MATH_STD ? std::log(x) : tpgx_math::log(x); }
    double atan(double x) const { return flavour == ORC_MATH_STD ? std::atan(x) : tpgx_math::atan(x); }
    double sin(double x) const { return flavour == ORC_MATH_STD ? std::sin(x) : tpgx_math::sin(x); }
    double cos(double x) const { return flavour == ORC_MATH_STD ? std::cos(x) : tpgx_math::cos(x); }
    double tan(double x) const { return flavour == ORC_MATH_STD ? std::tan(x) : tpgx_math::tan(x); }
};

/* Instruction semantics: one case per distinct lambda of the five apps. */
double apply_op(int op, const Math& m, double a, double b, int ia, int ib, double c,
                const double* w /* 3x3 row-major */) {
    switch (op) {
    case TPGX_OP_SUB: return a - b;                    /* [REF] mnist/src/main.cpp:47 */
    case TPGX_OP_ADD: return a + b;                    /* :48 */
    case TPGX_OP_MUL: return a * b;                    /* :49 */
    case TPGX_OP_DIV: return a / b;                    /* :50 */
    case TPGX_OP_MAX: return (a < b) ? b : a;          /* :51 std::max(a,b) */
    case TPGX_OP_EXP: return m.exp(a);                 /* :53 */
    case TPGX_OP_LOG: return m.log(a);                 /* :52 */
    case TPGX_OP_COS: return m.cos(a);                 /* [REF] pendulum/src/Learn/instructions.cpp:14 */
    case TPGX_OP_SIN: return m.sin(a);                 /* :15 */
    case TPGX_OP_TAN: return m.tan(a);                 /* :16 */
    case TPGX_OP_PI: return M_PI;                      /* :17 */
    case TPGX_OP_MULC: return a * c / 10.0;            /* [REF] mnist/src/main.cpp:54 */
    case TPGX_OP_FMODG: return b != 0.0 ? std::fmod(a, b) : DBL_MIN; /* [REF] stickgame/src/Learn/instructions.cpp:14-16 */
    case TPGX_OP_SUB_II: return (double)ia - (double)ib; /* :9 */
    case TPGX_OP_CAST_I: return (double)ia;            /* :10 */
    case TPGX_OP_EQ0: return (a == 0.0) ? 10.0 : 0.0;  /* :13; [REF] tic-tac-toe/src/Learn/main.cpp:20 */
    case TPGX_OP_EQM1: return (a == -1.0) ? 10.0 : 0.0;/* [REF] tic-tac-toe/src/Learn/main.cpp:19 */
    case TPGX_OP_EQ1: return (a == 1.0) ? 10.0 : 0.0;  /* :21 */
    case TPGX_OP_GE15: return (a >= 15.0) ? 10.0 : 0.0;/* :22 */
    case TPGX_OP_COND: return a < b ? -a : a;          /* :23 */
    case TPGX_OP_SOBEL_MAGN: {                         /* [REF] mnist/src/main.cpp:55-65 */
        double gx = -w[0] + w[2] - 2.0 * w[3] + 2.0 * w[5] - w[6] + w[8];
        double gy = -w[0] - 2.0 * w[1] - w[2] + w[6] + 2.0 * w[7] + w[8];
        return std::sqrt(gx * gx + gy * gy);
    }
    case TPGX_OP_SOBEL_DIR: {                          /* [REF] mnist/src/main.cpp:67-77 */
        double gx = -w[0] + w[2] - 2.0 * w[3] + 2.0 * w[5] - w[6] + w[8];
        double gy = -w[0] - 2.0 * w[1] - w[2] + w[6] + 2.0 * w[7] + w[8];
        return m.atan(gy / gx);
    }
    default: return std::numeric_limits<double>::quiet_NaN();
    }
}

/* ------------------------------------------------------------------ data model */
struct Program {
    std::vector<tpgx_line> lines;
    std::vector<uint8_t> intron;
    std::vector<double> constants;
    int n_eff = 0;
};

/* One observation: per environment data source a typed view of the current values. */
struct SourceView {
    const double* f64 = nullptr;
    const int32_t* i32 = nullptr;
    const uint8_t* u8 = nullptr; /* exact stand-in for integer-valued doubles (MNIST pixels) */
    double get_d(size_t i) const { return f64 ? f64[i] : (double)u8[i]; }
};

/* [UP] Archive (App. F U12): one per job in TRAINING evaluation.  addRecording is called by evaluateEdge for every
 * executed program; it records with probability p — `p == 1.0 || p >= rng.getDouble(0, 1)`, so no draw is consumed
 * when p is 1 — and keeps at most max_size recordings (oldest dropped). */
struct ArchiveRec { int32_t program; int64_t sample; double bid; int32_t job, iteration, step; double obs[16]; };
struct ArchiveRecorder {
    std::mt19937_64 eng;
    double p = 0.0;
    size_t max_size = 0;
    int64_t sample = -1;                 /* stands for the hash of the data sources: which observation is current */
    int32_t job = 0, iteration = 0, step = 0; /* episodic environments: where the current observation was seen */
    double obs[16] = {0};                /* ... and a copy of it ([UP] the archive clones the data handlers) */
    std::deque<ArchiveRec> recs;
    void add(int32_t program, double bid) {
        if (p == 1.0 || p >= std::uniform_real_distribution<double>(0.0, 1.0)(eng)) {
            ArchiveRec r{program, sample, bid, job, iteration, step, {0}};
            std::memcpy(r.obs, obs, sizeof obs);
            recs.push_back(r);
            while (recs.size() > max_size) recs.pop_front();
        }
    }
};

struct Counters {
    ArchiveRecorder* archive = nullptr;
    uint64_t evaluations = 0, team_visits = 0, program_executions = 0, lines_executed = 0;
    void add(const Counters& o) {
        evaluations += o.evaluations; team_visits += o.team_visits;
        program_executions += o.program_executions; lines_executed += o.lines_executed;
    }
};

} // namespace

struct orc_engine {
    tpgx_config cfg{};
    std::vector<int32_t> opcode;
    std::vector<tpgx_source_desc> src;
    std::vector<Program> programs;
    std::vector<int32_t> team_edge_offset, edge_program, edge_destination, root_team;
    Math math{ORC_MATH_SHARED};

    int first_env_source() const { return cfg.nb_constants > 0 ? 2 : 1; }
    int nb_sources_total() const { return first_env_source() + (int)src.size(); }

    /* [UP] DataHandler::getAddressSpace(type) for handler index `s` of
     * [registers, constants iff nbConstants > 0, data sources...]. */
    size_t address_space(int s, OperandType t) const {
        if (s == 0) return t == T_D ? (size_t)cfg.nb_registers : 0;          /* PrimitiveTypeArray<double> */
        if (s == 1 && cfg.nb_constants > 0) return t == T_C ? (size_t)cfg.nb_constants : 0;
        int k = s - first_env_source();
        if (k < 0 || k >= (int)src.size()) return 0;
        const tpgx_source_desc& d = src[k];
        if (d.elem == TPGX_ELEM_I32) return (t == T_I && !d.is_2d) ? (size_t)d.width : 0;
        if (t == T_D) return (size_t)d.height * (size_t)d.width;
        if (t == T_W && d.is_2d && d.height >= 3 && d.width >= 3)             /* Array2DWrapper: (h-2)*(w-2) */
            return (size_t)(d.height - 2) * (size_t)(d.width - 2);
        return 0;
    }

    /* [UP] Program::identifyIntrons: backward liveness from register 0 over register operands of
     * the instruction's real arity (App. F U8). */
    void identify_introns(Program& p) const {
        std::vector<char> useful(cfg.nb_registers, 0);
        useful[0] = 1;
        p.intron.assign(p.lines.size(), 1);
        p.n_eff = 0;
        for (size_t i = p.lines.size(); i-- > 0;) {
            const tpgx_line& l = p.lines[i];
            if (!useful[l.destination]) continue;
            p.intron[i] = 0;
            p.n_eff++;
            useful[l.destination] = 0;
            OpInfo oi = op_info(opcode[l.instruction]);
            for (int k = 0; k < oi.nb_operands; k++)
                if (l.src[k] == 0) useful[l.loc[k] % (uint32_t)cfg.nb_registers] = 1;
        }
    }

    /* [UP] ProgramExecutionEngine::executeProgram (+ fetchCurrentOperands): registers zeroed,
     * non-intron lines in order, operand = handler[src].getDataAt(type, loc % addressSpace(type)),
     * result = register 0 (App. F U4, U5). */
    double execute_program(const Program& p, const SourceView* obs, Counters* cnt) const {
        double reg[16];
        for (int i = 0; i < cfg.nb_registers; i++) reg[i] = 0.0;
        const int fes = first_env_source();
        for (size_t li = 0; li < p.lines.size(); li++) {
            if (p.intron[li]) continue;
            const tpgx_line& l = p.lines[li];
            const int op = opcode[l.instruction];
            const OpInfo oi = op_info(op);
            double d[2] = {0.0, 0.0};
            int iv[2] = {0, 0};
            double c = 0.0;
            double win[9];
            for (int k = 0; k < oi.nb_operands; k++) {
                const int s = (int)l.src[k];
                const OperandType t = oi.type[k];
                const size_t space = address_space(s, t);
                const size_t a = l.loc[k] % space; /* space > 0 checked at creation */
                if (s == 0) d[k] = reg[a];
                else if (s == 1 && cfg.nb_constants > 0) c = p.constants[a];
                else {
                    const SourceView& v = obs[s - fes];
                    const tpgx_source_desc& sd = src[s - fes];
                    if (t == T_D) d[k] = v.get_d(a);
                    else if (t == T_I) iv[k] = v.i32[a];
                    else { /* T_W: [UP] Array2DWrapper::getDataAt for T[3][3]: row = a / (w-2), col = a % (w-2) */
                        const size_t ww = (size_t)sd.width - 2;
                        const size_t r0 = a / ww, c0 = a % ww;
                        for (int r = 0; r < 3; r++)
                            for (int cc = 0; cc < 3; cc++)
                                win[r * 3 + cc] = v.get_d((r0 + r) * sd.width + c0 + cc);
                    }
                }
            }
            reg[l.destination] = apply_op(op, math, d[0], d[1], iv[0], iv[1], c, win);
            if (cnt) cnt->lines_executed++;
        }
        if (cnt) cnt->program_executions++;
        return reg[0];
    }

    /* [UP] TPGExecutionEngine::evaluateEdge: NaN -> -inf (App. F U2). */
    double evaluate_edge(int edge, const SourceView* obs, Counters* cnt) const {
        double r = execute_program(programs[edge_program[edge]], obs, cnt);
        if (std::isnan(r)) r = -std::numeric_limits<double>::infinity();
        if (cnt && cnt->archive) cnt->archive->add(edge_program[edge], r);   /* [UP] archive->addRecording(&prog, dataSources, r) */
        return r;
    }

    /* [UP] TPGExecutionEngine::evaluateTeam: out-edges in list order minus those leading to an
     * already visited vertex; first admissible edge is the initial best; a later edge replaces it
     * if bid >= best (last-max, U1) or bid > best (first-max knob).  Returns the edge index or -1
     * when no edge is admissible (the reference throws). */
    int evaluate_team(int team, const std::vector<int>& visited, const SourceView* obs, Counters* cnt) const {
        int best = -1;
        double best_bid = 0.0;
        if (cnt) cnt->team_visits++;
        for (int e = team_edge_offset[team]; e < team_edge_offset[team + 1]; e++) {
            const int dst = edge_destination[e];
            if (dst >= 0 && std::find(visited.begin(), visited.end(), dst) != visited.end()) continue;
            const double bid = evaluate_edge(e, obs, cnt);
            if (best < 0) { best = e; best_bid = bid; }
            else if (cfg.tie_break == TPGX_TIE_LAST_MAX ? (bid >= best_bid) : (bid > best_bid)) { best = e; best_bid = bid; }
        }
        return best;
    }

    /* [UP] TPGExecutionEngine::executeFromRoot: follow best edges until an action; returns the
     * action id, or -1 if a team had no admissible edge. */
    int execute_from_root(int root_team_id, const SourceView* obs, Counters* cnt, std::vector<int>& visited) const {
        visited.clear();
        int cur = root_team_id;
        visited.push_back(cur);
        for (;;) {
            int e = evaluate_team(cur, visited, obs, cnt);
            if (e < 0) return -1;
            int dst = edge_destination[e];
            if (dst < 0) return -1 - dst;
            cur = dst;
            visited.push_back(cur);
        }
    }
};

/* ================================================================== RNG ([UP] Mutator::RNG) */
namespace {
/* [UP] Mutator::RNG = std::mt19937_64 + std::uniform_{int,real}_distribution (App. F U10); the
 * oracle uses libstdc++'s own distributions, i.e. what a Linux build of the library computes. */
struct Rng {
    std::mt19937_64 eng;
    void set_seed(uint64_t s) { eng.seed(s); }
    uint64_t u64(uint64_t lo, uint64_t hi) { return std::uniform_int_distribution<uint64_t>(lo, hi)(eng); }
    double f64(double lo, double hi) { return std::uniform_real_distribution<double>(lo, hi)(eng); }
};

/* [UP] Data::Hash<T> = std::hash<T> on libstdc++ = identity on integral / enum values (App. F U9). */
inline uint64_t data_hash(uint64_t v) { return v; }

/* ================================================================== environments (Appendix B) */
struct Env {
    virtual ~Env() {}
    virtual void reset(uint64_t seed, int mode, bool raw = false) = 0; /* raw: `seed` is the value handed to rng.setSeed() as is */
    virtual bool do_action(double a) = 0; /* false = the reference would throw */
    virtual bool is_terminal() const = 0;
    virtual double score(int player) const = 0;
    virtual int nb_sources() const = 0;
    virtual SourceView view(int s) const = 0;
    virtual int state(double* out, int n) const = 0;
    virtual void set_state(const double* st, int n) = 0;
    /* 4-value digest of the final state written to tpgx_episode_result.final_state */
    virtual void summary(double* o) const { state(o, 4); }
};

/* [REF] gridworld/src/gridworld.{h,cpp} */
struct GridWorldEnv : Env {
    int x = 0, y = 0;
    bool terminated = false;
    double total = 0.0;
    double st[2] = {0, 0};
    static int tile(int yy, int xx) {
        static const int grid[3][4] = {{0, 0, 0, 2}, {0, 0, 3, 3}, {0, 0, 0, 1}}; /* gridworld.h:18-20 */
        return grid[yy][xx];
    }
    void reset(uint64_t, int, bool = false) override { x = 0; y = 0; terminated = false; total = 0.0; st[0] = 0; st[1] = 0; } /* gridworld.cpp:3-15 */
    static bool available(int px, int py) { /* gridworld.cpp:17-37 */
        if (px == 4 || px == -1) return false;
        if (py == 3 || py == -1) return false;
        return tile(py, px) != 3;
    }
    bool do_action(double a) override { /* gridworld.cpp:39-75 */
        switch ((uint64_t)a) {
        case 0: if (available(x - 1, y)) x--; break;
        case 1: if (available(x, y + 1)) y++; break;
        case 2: if (available(x + 1, y)) x++; break;
        case 3: if (available(x, y - 1)) y--; break;
        }
        double reward = -1;
        if (tile(y, x) == 1) { terminated = true; reward = 100; }
        else if (tile(y, x) == 2) { terminated = true; reward = -100; }
        total += reward;
        st[0] = x; st[1] = y;
        return true;
    }
    bool is_terminal() const override { return terminated; }
    double score(int) const override { return total; }
    int nb_sources() const override { return 1; }
    SourceView view(int) const override { SourceView v; v.f64 = st; return v; }
    int state(double* o, int n) const override {
        double s[4] = {st[0], st[1], total, terminated ? 1.0 : 0.0};
        for (int i = 0; i < n && i < 4; i++) o[i] = s[i];
        return 4;
    }
    void set_state(const double* s, int n) override { if (n >= 2) { x = (int)s[0]; y = (int)s[1]; st[0] = x; st[1] = y; } }
};

/* [REF] stickgame/src/Learn/stickGameAdversarial.{h,cpp} */
struct StickGameEnv : Env {
    int32_t sticks[1] = {21};
    int32_t hints[4] = {1, 2, 3, 4}; /* stickGameAdversarial.h:95-98 */
    bool first_won = false, forb1 = false, forb2 = false;
    int turn = 0;
    const int error_rate = 4;        /* .h:80 */
    bool second_random;
    Rng rng;
    explicit StickGameEnv(bool sr) : second_random(sr) { reset(0, 0); }
    void reset(uint64_t seed, int mode, bool raw = false) override { /* .cpp:105-116 */
        rng.set_seed(raw ? seed : (data_hash(seed) ^ data_hash((uint64_t)mode)));
        sticks[0] = 21; first_won = false; forb1 = false; forb2 = false; turn = 0;
    }
    void random_play() { /* .cpp:38-57 */
        if (!is_terminal()) {
            int s = sticks[0];
            if (rng.u64(0, error_rate - 1) != 0 && (s - 1) % 4 != 0) s -= (s - 1) % 4;
            else s -= (int)rng.u64(1, std::min(s, 3));
            sticks[0] = s;
            if (s == 0) first_won = ((!turn) % 2 == 0); /* `!this->currentTurn % 2 == 0` as written (.cpp:54) */
            turn++;
        }
    }
    bool do_action(double a) override { /* .cpp:59-103 */
        if (a < 0 || a >= 3.0) return false; /* :61 LearningEnvironment::doAction(actionID) throws past nbActions */
        if (!is_terminal()) {
            bool first = turn % 2 == 0;
            bool& forbidden = first ? forb1 : forb2;
            int s = sticks[0];
            if ((a + 1) > s) { forbidden = true; sticks[0] = 0; first_won = !first; return true; }
            s -= ((int)a + 1);
            sticks[0] = s;
            if (s == 0) first_won = !first;
            turn++;
            if (second_random) random_play();
        }
        return true;
    }
    bool is_terminal() const override { return sticks[0] == 0; }
    double score(int player) const override { /* .cpp:127-153; getScore() = first player's [UP] */
        double s1, s2;
        if (first_won) { s1 = 1.0; s2 = forb2 ? -1.0 : 0.0; }
        else { s2 = 1.0; s1 = forb1 ? -1.0 : 0.0; }
        return player == 0 ? s1 : s2;
    }
    int nb_sources() const override { return 2; }
    SourceView view(int s) const override { SourceView v; v.i32 = (s == 0) ? hints : sticks; return v; } /* .cpp:118-125: hints first */
    int state(double* o, int n) const override {
        double s[4] = {(double)sticks[0], (double)turn, first_won ? 1.0 : 0.0, (forb1 ? 1.0 : 0.0) + (forb2 ? 2.0 : 0.0)};
        for (int i = 0; i < n && i < 4; i++) o[i] = s[i];
        return 4;
    }
    void set_state(const double* s, int n) override { if (n >= 1) sticks[0] = (int)s[0]; if (n >= 2) turn = (int)s[1]; }
};

/* [REF] tic-tac-toe/src/Learn/TicTacToe.{h,cpp} */
struct TicTacToeEnv : Env {
    double board[9];
    int turn = 0;
    bool win1 = false, win2 = false, forb1 = false, forb2 = false, null_game = false, end = false;
    bool second_random;
    Rng rng;
    explicit TicTacToeEnv(bool sr) : second_random(sr) { reset(0, 0); }
    void reset(uint64_t seed, int mode, bool raw = false) override { /* TicTacToe.cpp:101-116 */
        rng.set_seed(raw ? seed : (data_hash(seed) ^ data_hash((uint64_t)mode)));
        for (int i = 0; i < 9; i++) board[i] = -1.0;
        turn = 0; win1 = win2 = forb1 = forb2 = null_game = end = false;
    }
    void random_play(double sym) { /* :83-99 */
        int remaining = 9 - turn;
        int decision = (int)rng.u64(0, (uint64_t)(remaining - 1));
        int i = -1;
        while (decision >= 0) { i++; if (board[i] == -1) decision--; }
        board[i] = sym;
    }
    void revert_board() { /* :76-81 */
        for (int i = 0; i < 9; i++) { double c = board[i]; board[i] = c == 0 ? 1 : c == 1 ? 0 : -1; }
    }
    void update_game() { /* :147-200 */
        bool& win = (turn % 2 != 0 ? win1 : win2);
        const double x00 = board[0], x10 = board[1], x20 = board[2], x01 = board[3], x11 = board[4],
                     x21 = board[5], x02 = board[6], x12 = board[7], x22 = board[8];
        if (x11 != -1 && ((x01 == x11 && x11 == x21) || (x10 == x11 && x11 == x12) ||
                          (x00 == x11 && x11 == x22) || (x20 == x11 && x11 == x02))) { end = true; win = true; return; }
        if (x22 != -1.0 && ((x02 == x12 && x12 == x22) || (x20 == x21 && x21 == x22))) { win = true; end = true; return; }
        if (x00 != -1.0 && ((x00 == x10 && x10 == x20) || (x00 == x01 && x01 == x02))) { win = true; end = true; return; }
        if (turn > 8) { null_game = true; end = true; }
    }
    bool do_action(double a) override { /* :33-74 */
        if (a < 0 || a >= 9.0) return false; /* :41,46 the board handler's getDataAt / setDataAt throw outside its 9 cells */
        bool& forbidden = (turn % 2 == 0 ? forb1 : forb2);
        if (!is_terminal()) {
            double cell = board[(int)a];
            if (cell != -1.0) { forbidden = true; random_play(0); }
            else board[(size_t)a] = 0;
            turn++;
            update_game();
            if (is_terminal()) return true;
            if (second_random) { random_play(1); turn++; update_game(); }
            else revert_board();
        }
        return true;
    }
    bool is_terminal() const override { return end; }
    double score(int player) const override { /* :124-145 */
        double m1 = forb1 ? 1.0 : 0.0, m2 = forb2 ? 1.0 : 0.0;
        if (null_game) return player == 0 ? 0.5 - m1 : 0.5 - m2;
        if (win1) return player == 0 ? 1 - m1 : 0 - m2;
        return player == 0 ? 0 - m1 : 1 - m2;
    }
    int nb_sources() const override { return 1; }
    SourceView view(int) const override { SourceView v; v.f64 = board; return v; }
    int state(double* o, int n) const override {
        int k = 0;
        for (int i = 0; i < 9 && k < n; i++) o[k++] = board[i];
        if (k < n) o[k++] = turn;
        if (k < n) o[k++] = (win1 ? 1 : 0) + (win2 ? 2 : 0) + (forb1 ? 4 : 0) + (forb2 ? 8 : 0) + (null_game ? 16 : 0) + (end ? 32 : 0);
        return 11;
    }
    void set_state(const double* s, int n) override { for (int i = 0; i < 9 && i < n; i++) board[i] = s[i]; if (n >= 10) turn = (int)s[9]; }
    void summary(double* o) const override { /* board as a base-3 number, turn, flag bits */
        double code = 0.0, w = 1.0;
        for (int i = 0; i < 9; i++) { code += (board[i] + 1.0) * w; w *= 3.0; }
        o[0] = code; o[1] = turn;
        o[2] = (win1 ? 1 : 0) + (win2 ? 2 : 0) + (forb1 ? 4 : 0) + (forb2 ? 8 : 0) + (null_game ? 16 : 0) + (end ? 32 : 0);
        o[3] = 0.0;
    }
};

/* [REF] pendulum/src/Learn/pendulum.{h,cpp} */
struct PendulumEnv : Env {
    static constexpr double MAX_SPEED = 8.0, MAX_TORQUE = 2.0, TIME_DELTA = 0.05, G = 9.81, MASS = 1.0,
                            LENGTH = 1.0, STABILITY_THRESHOLD = 0.1; /* pendulum.cpp:6-12 */
    static constexpr size_t HIST = 300;                               /* pendulum.h:23 */
    std::vector<double> avail;
    Rng rng;
    double hist[HIST];
    double total = 0.0;
    uint64_t n_act = 0;
    double st[2] = {0, 0};
    Math m;
    PendulumEnv(const double* a, int n, Math mm) : avail(a, a + n), m(mm) { for (auto& h : hist) h = 0.0; }
    void reset(uint64_t seed, int mode, bool raw = false) override { /* pendulum.cpp:42-55 */
        rng.set_seed(raw ? seed : (data_hash(seed) ^ data_hash((uint64_t)mode)));
        st[0] = rng.f64(-M_PI, M_PI);
        st[1] = rng.f64(-1.0, 1.0);
        n_act = 0; total = 0.0;
    }
    double action_from_id(uint64_t id) const { /* :57-61 */
        double r = (id == 0) ? 0.0 : avail.at((id - 1) % avail.size());
        return (id <= avail.size()) ? r : -r;
    }
    bool do_action(double a) override { /* :63-98, exact operation order */
        if (a < 0 || a >= avail.size() * 2 + 1) return false;
        double cur = action_from_id((uint64_t)a);
        cur *= MAX_TORQUE;
        double angle = st[0], velocity = st[1];
        double atu = std::fmod((angle + M_PI), (2.f * M_PI)) - M_PI;
        double reward = -((atu * atu) + 0.2f * (velocity * velocity) + 0.01f * (cur * cur));
        hist[n_act % HIST] = reward;
        n_act++;
        total += reward;
        velocity = velocity + ((-3.0) * G / (2.0 * LENGTH) * (m.sin(angle + M_PI)) + (3.f / (MASS * LENGTH * LENGTH)) * cur) * TIME_DELTA;
        velocity = std::fmin(std::fmax(velocity, -MAX_SPEED), MAX_SPEED);
        angle = angle + velocity * TIME_DELTA;
        st[0] = angle; st[1] = velocity;
        return true;
    }
    bool is_terminal() const override { /* :122-141 */
        bool r = false;
        if (n_act >= HIST) {
            double acc = 0.0;
            for (size_t i = 0; i < HIST; i++) acc += hist[i];
            acc /= (double)HIST;
            r = (std::fabs(acc) < STABILITY_THRESHOLD);
        }
        return r;
    }
    double score(int) const override { /* :110-120 */
        if (is_terminal()) return 10.0 / m.log(((double)n_act - (double)HIST + 2.0));
        return total / (double)n_act;
    }
    int nb_sources() const override { return 1; }
    SourceView view(int) const override { SourceView v; v.f64 = st; return v; }
    int state(double* o, int n) const override {
        double s[4] = {st[0], st[1], total, (double)n_act};
        for (int i = 0; i < n && i < 4; i++) o[i] = s[i];
        return 4;
    }
    void set_state(const double* s, int n) override { if (n >= 2) { st[0] = s[0]; st[1] = s[1]; } }
};

std::unique_ptr<Env> make_env(int kind, bool second_random, const double* pa, int npa, Math m) {
    switch (kind) {
    case TPGX_ENV_GRIDWORLD: return std::unique_ptr<Env>(new GridWorldEnv());
    case TPGX_ENV_STICKGAME: return std::unique_ptr<Env>(new StickGameEnv(second_random));
    case TPGX_ENV_TICTACTOE: return std::unique_ptr<Env>(new TicTacToeEnv(second_random));
    case TPGX_ENV_PENDULUM: return std::unique_ptr<Env>(new PendulumEnv(pa, npa, m));
    }
    return nullptr;
}

/* [UP] ParallelLearningAgent::slaveEvalJobThread: mutex-guarded job queue, results keyed by job
 * index so the output does not depend on scheduling. fn(job, thread-local counters). */
template <class Fn>
void run_jobs(int nb_jobs, int nb_threads, Counters& total, Fn fn) {
    if (nb_threads <= 1) { /* [UP] LearningAgent::evaluateAllRoots (sequential) */
        for (int j = 0; j < nb_jobs; j++) fn(j, total);
        return;
    }
    std::mutex mtx;
    int next = 0;
    std::vector<Counters> per(nb_threads);
    auto worker = [&](int t) {
        for (;;) {
            int j;
            { std::lock_guard<std::mutex> g(mtx); if (next >= nb_jobs) return; j = next++; }
            fn(j, per[t]);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nb_threads; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
    for (auto& c : per) total.add(c);
}

void export_counters(const Counters& c, tpgx_counters* o) {
    if (!o) return;
    std::memset(o, 0, sizeof(*o));
    o->evaluations = c.evaluations; o->team_visits = c.team_visits;
    o->program_executions = c.program_executions; o->lines_executed = c.lines_executed;
}

} // namespace

/* ================================================================== C API */
extern "C" {

orc_engine* orc_create(const tpgx_config* cfg, int32_t n_ops, const int32_t* opcode, int32_t n_src,
                       const tpgx_source_desc* src, const tpgx_graph* g, int32_t math_flavour,
                       char* err, size_t err_len) {
    auto fail = [&](const std::string& m) -> orc_engine* { if (err && err_len) snprintf(err, err_len, "%s", m.c_str()); return nullptr; };
    if (!cfg || !opcode || !g) return fail("null argument");
    if (cfg->nb_registers < 1 || cfg->nb_registers > 16) return fail("nb_registers out of range");
    std::unique_ptr<orc_engine> e(new orc_engine());
    e->cfg = *cfg;
    e->opcode.assign(opcode, opcode + n_ops);
    for (int op : e->opcode) if (op < 0 || op >= TPGX_OP_COUNT) return fail("unknown opcode");
    if (n_src > 0) e->src.assign(src, src + n_src);
    e->math.flavour = math_flavour;
    e->programs.resize(g->nb_programs);
    for (int p = 0; p < g->nb_programs; p++) {
        Program& pr = e->programs[p];
        pr.lines.assign(g->lines + g->program_line_offset[p], g->lines + g->program_line_offset[p + 1]);
        if (cfg->nb_constants > 0) {
            if (!g->constants) return fail("constants missing");
            pr.constants.assign(g->constants + (size_t)p * cfg->nb_constants, g->constants + (size_t)(p + 1) * cfg->nb_constants);
        }
        for (const tpgx_line& l : pr.lines) {
            if (l.instruction >= (uint32_t)n_ops) return fail("instruction index out of range");
            if (l.destination >= (uint32_t)cfg->nb_registers) return fail("destination register out of range");
            OpInfo oi = op_info(e->opcode[l.instruction]);
            for (int k = 0; k < oi.nb_operands; k++) {
                if ((int)l.src[k] >= e->nb_sources_total()) return fail("operand source index out of range");
                if (e->address_space((int)l.src[k], oi.type[k]) == 0)
                    return fail("operand source cannot provide the operand type (the reference would throw)");
            }
        }
        e->identify_introns(pr);
        if (g->intron) { /* caller-provided flags must agree with the restated algorithm */
            for (size_t i = 0; i < pr.lines.size(); i++)
                if ((g->intron[g->program_line_offset[p] + i] != 0) != (pr.intron[i] != 0)) return fail("intron flags disagree with identifyIntrons");
        }
    }
    e->team_edge_offset.assign(g->team_edge_offset, g->team_edge_offset + g->nb_teams + 1);
    int nE = g->team_edge_offset[g->nb_teams];
    e->edge_program.assign(g->edge_program, g->edge_program + nE);
    e->edge_destination.assign(g->edge_destination, g->edge_destination + nE);
    for (int i = 0; i < nE; i++) {
        if (e->edge_program[i] < 0 || e->edge_program[i] >= g->nb_programs) return fail("edge program out of range");
        if (e->edge_destination[i] >= g->nb_teams) return fail("edge destination out of range");
    }
    e->root_team.assign(g->root_team, g->root_team + g->nb_roots);
    for (int r : e->root_team) if (r < 0 || r >= g->nb_teams) return fail("root out of range");
    return e.release();
}

void orc_destroy(orc_engine* e) { delete e; }

int64_t orc_identify_introns(const orc_engine* e, uint8_t* flags) {
    int64_t n = 0, k = 0;
    for (const Program& p : e->programs)
        for (uint8_t f : p.intron) { if (flags) flags[k] = f; k++; n += f; }
    return n;
}

void orc_effective_lines(const orc_engine* e, int32_t* n_eff) {
    for (size_t p = 0; p < e->programs.size(); p++) n_eff[p] = e->programs[p].n_eff;
}

int32_t orc_execute_programs(const orc_engine* e, int64_t count, const double* const* obs_f64,
                             const int32_t* const* obs_i32, double* bid, int32_t raw) {
    const int ns = (int)e->src.size();
    std::vector<SourceView> v(ns);
    for (int64_t s = 0; s < count; s++) {
        for (int k = 0; k < ns; k++) {
            const size_t space = (size_t)e->src[k].height * e->src[k].width;
            v[k] = SourceView();
            if (e->src[k].elem == TPGX_ELEM_F64) v[k].f64 = obs_f64[k] + s * space;
            else v[k].i32 = obs_i32[k] + s * space;
        }
        for (size_t p = 0; p < e->programs.size(); p++) {
            double r = e->execute_program(e->programs[p], v.data(), nullptr);
            if (!raw && std::isnan(r)) r = -std::numeric_limits<double>::infinity();
            bid[p * count + s] = r;
        }
    }
    return 0;
}

/* [UP] ClassificationLearningAgent::evaluateJob (one iteration, as mnist/params.json:97):
 * table[class][action]++ per doAction ([UP] ClassificationLearningEnvironment::doAction, called
 * from [REF] mnist/src/mnist.cpp:53); per class F1 = 2*(P*R)/(P+R), 0 when TP == 0; general
 * score = mean of the per-class scores (App. F U11). */
int32_t orc_evaluate_classification(const orc_engine* e, const uint8_t* images, const uint8_t* labels,
                                    int64_t n_images, const tpgx_classif_request* req,
                                    tpgx_classif_result* out, int32_t nb_threads) {
    if (e->src.size() != 1) return TPGX_ERR_BAD_ARG;
    const size_t px = (size_t)e->src[0].height * e->src[0].width;
    const int C = req->nb_classes;
    const int nR = req->root_end - req->root_begin;
    const int64_t nS = req->nb_samples;
    for (int64_t s = 0; s < nS; s++) {
        int64_t idx = req->sample_index ? req->sample_index[s] : s;
        if (idx < 0 || idx >= n_images) return TPGX_ERR_BAD_ARG;
    }
    Counters total;
    std::atomic<int> status{0};
    run_jobs(nR, nb_threads, total, [&](int j, Counters& cnt) {
        const int root = e->root_team[req->root_begin + j];
        std::vector<uint64_t> table((size_t)C * C, 0);
        std::vector<int> visited;
        for (int64_t s = 0; s < nS; s++) {
            const int64_t idx = req->sample_index ? req->sample_index[s] : s;
            SourceView v;
            v.u8 = images + (size_t)idx * px;
            int action = e->execute_from_root(root, &v, &cnt, visited);
            cnt.evaluations++;
            if (action < 0 || action >= C) { status = TPGX_ERR_GRAPH_INVALID; return; }
            table[(size_t)labels[idx] * C + action]++;
            if (out->action) out->action[(size_t)j * nS + s] = (uint8_t)action;
        }
        double mean = 0.0;
        std::vector<double> f1(C);
        for (int c = 0; c < C; c++) {
            uint64_t tp = table[(size_t)c * C + c];
            uint64_t fn = std::accumulate(table.begin() + (size_t)c * C, table.begin() + (size_t)(c + 1) * C, (uint64_t)0) - tp;
            uint64_t fp = 0;
            for (int r = 0; r < C; r++) fp += table[(size_t)r * C + c];
            fp -= tp;
            double recall = (double)tp / (double)(tp + fn);
            double precision = (double)tp / (double)(tp + fp);
            double f = (tp != 0) ? 2 * (precision * recall) / (precision + recall) : 0.0;
            f1[c] = f / 1.0; /* divided by nbIterationsPerPolicyEvaluation = 1 */
        }
        mean = std::accumulate(f1.begin(), f1.end(), 0.0) / (double)C;
        if (out->score) out->score[j] = mean;
        if (out->class_f1) for (int c = 0; c < C; c++) out->class_f1[(size_t)j * C + c] = f1[c];
        if (out->confusion) std::copy(table.begin(), table.end(), out->confusion + (size_t)j * C * C);
    });
    if (status) return status;
    if (out->bid) {
        for (int64_t s = 0; s < nS; s++) {
            const int64_t idx = req->sample_index ? req->sample_index[s] : s;
            SourceView v;
            v.u8 = images + (size_t)idx * px;
            for (size_t p = 0; p < e->programs.size(); p++) {
                double r = e->execute_program(e->programs[p], &v, nullptr);
                if (std::isnan(r)) r = -std::numeric_limits<double>::infinity();
                out->bid[p * (size_t)nS + s] = r;
            }
        }
    }
    export_counters(total, out->counters);
    return 0;
}

/* [UP] LearningAgent::evaluateJob: for it in iterations: le.reset(seed_it, mode); while
 * !isTerminal && n < maxNbActionsPerEval: action = executeFromRoot; doAction; result += getScore();
 * result / nbIterations.  [UP] AdversarialLearningAgent::evaluateJob: k roots take turns, one
 * action each, every action counts towards maxNbActionsPerEval; per-player scores from getScores(). */
int32_t orc_evaluate_episodes(const orc_engine* e, const tpgx_episode_request* req,
                              tpgx_episode_result* out, int32_t nb_threads) {
    const int K = req->roots_per_job, NI = req->nb_iterations;
    if (K < 1 || K > 2 || NI < 1) return TPGX_ERR_BAD_ARG;
    Counters total;
    std::atomic<int> status{0};
    run_jobs(req->nb_jobs, nb_threads, total, [&](int j, Counters& cnt) {
        auto env = make_env(req->env, req->second_player_random != 0, req->pendulum_actions, req->nb_pendulum_actions, e->math);
        if (!env || env->nb_sources() != (int)e->src.size()) { status = TPGX_ERR_BAD_ARG; return; }
        std::vector<SourceView> v(env->nb_sources());
        std::vector<int> visited;
        std::vector<double> result(K, 0.0);
        for (int it = 0; it < NI; it++) {
            if (req->rng_seeds) env->reset(req->rng_seeds[it], req->mode, true); else env->reset(req->seeds[it], req->mode);
            int n = 0, turn = 0;
            while (!env->is_terminal() && n < req->max_actions) {
                for (int s = 0; s < env->nb_sources(); s++) v[s] = env->view(s);
                const int root = e->root_team[req->job_roots[(size_t)j * K + turn]];
                int action = e->execute_from_root(root, v.data(), &cnt, visited);
                cnt.evaluations++;
                if (action < 0) { status = TPGX_ERR_GRAPH_INVALID; return; }
                if (out->trace && n < req->trace_steps) out->trace[((size_t)j * NI + it) * req->trace_steps + n] = (int8_t)action;
                if (!env->do_action((double)action)) { status = TPGX_ERR_BAD_ARG; return; }
                turn = (turn + 1) % K;
                n++;
            }
            if (out->trace) for (int q = n; q < req->trace_steps; q++) out->trace[((size_t)j * NI + it) * req->trace_steps + q] = -1;
            for (int k = 0; k < K; k++) {
                double sc = env->score(k);
                result[k] += sc;
                if (out->episode_score) out->episode_score[((size_t)j * NI + it) * K + k] = sc;
            }
            if (out->episode_actions) out->episode_actions[(size_t)j * NI + it] = n;
            if (out->final_state) env->summary(out->final_state + ((size_t)j * NI + it) * 4);
        }
        if (out->score) for (int k = 0; k < K; k++) out->score[(size_t)j * K + k] = result[k] / (double)NI;
    });
    if (status) return status;
    export_counters(total, out->counters);
    return 0;
}

/* [UP] evaluateAllRoots in TRAINING mode on a sequential environment, archive side effect: one Archive per job (shared by
 * both roots of an adversarial job), a draw per executed program in execution order, per-job rings merged in job order. */
int32_t orc_archive_episodes(const orc_engine* e, const tpgx_episode_request* req, const uint64_t* archive_seed, double probability,
                             int32_t archive_size, int32_t* out_program, int32_t* out_job, int32_t* out_iteration, int32_t* out_step,
                             double* out_bid, double* out_obs, int32_t obs_len, int32_t* nb_records) {
    const int K = req->roots_per_job, NI = req->nb_iterations;
    if (K < 1 || K > 2 || NI < 1 || archive_size < 0) return TPGX_ERR_BAD_ARG;
    std::deque<ArchiveRec> merged;
    std::vector<int> visited;
    for (int j = 0; j < req->nb_jobs; j++) {
        ArchiveRecorder ar;
        ar.eng.seed(archive_seed[j]);
        ar.p = probability;
        ar.max_size = (size_t)archive_size;
        ar.job = j;
        Counters cnt;
        cnt.archive = &ar;
        auto env = make_env(req->env, req->second_player_random != 0, req->pendulum_actions, req->nb_pendulum_actions, e->math);
        if (!env || env->nb_sources() != (int)e->src.size()) return TPGX_ERR_BAD_ARG;
        std::vector<SourceView> v(env->nb_sources());
        for (int it = 0; it < NI; it++) {
            if (req->rng_seeds) env->reset(req->rng_seeds[it], req->mode, true); else env->reset(req->seeds[it], req->mode);
            int n = 0, turn = 0;
            while (!env->is_terminal() && n < req->max_actions) {
                int q = 0;
                for (int s = 0; s < env->nb_sources(); s++) {
                    v[s] = env->view(s);
                    for (int i = 0; i < e->src[s].width * std::max(1, e->src[s].height) && q < 16; i++)
                        ar.obs[q++] = v[s].f64 ? v[s].f64[i] : (double)v[s].i32[i];
                }
                for (; q < 16; q++) ar.obs[q] = 0.0;
                ar.iteration = it; ar.step = n;
                const int root = e->root_team[req->job_roots[(size_t)j * K + turn]];
                const int action = e->execute_from_root(root, v.data(), &cnt, visited);
                if (action < 0) return TPGX_ERR_GRAPH_INVALID;
                if (!env->do_action((double)action)) return TPGX_ERR_BAD_ARG;
                turn = (turn + 1) % K;
                n++;
            }
        }
        for (const ArchiveRec& r : ar.recs) merged.push_back(r);
        while (merged.size() > (size_t)archive_size) merged.pop_front();
    }
    *nb_records = (int32_t)merged.size();
    for (size_t i = 0; i < merged.size(); i++) {
        const ArchiveRec& r = merged[i];
        out_program[i] = r.program; out_job[i] = r.job; out_iteration[i] = r.iteration; out_step[i] = r.step; out_bid[i] = r.bid;
        if (out_obs) for (int q = 0; q < obs_len; q++) out_obs[i * (size_t)obs_len + q] = q < 16 ? r.obs[q] : 0.0;
    }
    return 0;
}

double orc_apply_op(int32_t opcode, int32_t math_flavour, double a, double b, int32_t ia, int32_t ib, double c, const double* win) {
    Math m{math_flavour};
    return apply_op(opcode, m, a, b, ia, ib, c, win);
}

void orc_math_array(int32_t fn, int32_t math_flavour, int64_t n, const double* in, double* out) {
    Math m{math_flavour};
    for (int64_t i = 0; i < n; i++) out[i] = apply_op(fn, m, in[i], 0.0, 0, 0, 0.0, nullptr);
}

/* [UP] ParallelLearningAgent::evaluateAllRoots in TRAINING mode, archive side effect only: every job evaluates its root
 * with a private Archive(archive_size, probability) seeded with the job's archiveSeed; mergeArchiveMap then keeps the
 * LAST archive_size recordings of the per-job archives concatenated in job-index order (App. F U12). */
int32_t orc_archive_classification(const orc_engine* e, const uint8_t* images, int64_t n_images, const tpgx_classif_request* req,
                                   const uint64_t* archive_seed, double probability, int32_t archive_size,
                                   int32_t* out_program, int64_t* out_sample, double* out_bid, int32_t* nb_records) {
    if (e->src.size() != 1 || archive_size < 0) return TPGX_ERR_BAD_ARG;
    const size_t px = (size_t)e->src[0].height * e->src[0].width;
    const int nR = req->root_end - req->root_begin;
    std::deque<ArchiveRec> merged;
    std::vector<int> visited;
    for (int j = 0; j < nR; j++) {
        ArchiveRecorder ar;
        ar.eng.seed(archive_seed[j]);
        ar.p = probability;
        ar.max_size = (size_t)archive_size;
        Counters cnt;
        cnt.archive = &ar;
        const int root = e->root_team[req->root_begin + j];
        for (int64_t s = 0; s < req->nb_samples; s++) {
            const int64_t idx = req->sample_index ? req->sample_index[s] : s;
            if (idx < 0 || idx >= n_images) return TPGX_ERR_BAD_ARG;
            SourceView v;
            v.u8 = images + (size_t)idx * px;
            ar.sample = idx;
            if (e->execute_from_root(root, &v, &cnt, visited) < 0) return TPGX_ERR_GRAPH_INVALID;
        }
        for (const ArchiveRec& r : ar.recs) merged.push_back(r);
        while (merged.size() > (size_t)archive_size) merged.pop_front();
    }
    *nb_records = (int32_t)merged.size();
    for (size_t i = 0; i < merged.size(); i++) { out_program[i] = merged[i].program; out_sample[i] = merged[i].sample; out_bid[i] = merged[i].bid; }
    return 0;
}

/* Host build of the primitives tpgx_math_probe runs on the device (same fn codes, include/tpgx.h). */
void orc_math_probe(int32_t fn, int64_t n, const double* a, const double* b, double* out) {
    for (int64_t i = 0; i < n; i++) {
        const double x = a[i], y = b ? b[i] : 0.0;
        switch (fn) {
        case TPGX_PROBE_DIV: case TPGX_PROBE_DIV_K: out[i] = x / y; break;
        case TPGX_PROBE_EXP: case TPGX_PROBE_EXP_K: out[i] = tpgx_math::exp(x); break;
        case TPGX_PROBE_LOG: case TPGX_PROBE_LOG_K: out[i] = tpgx_math::log(x); break;
        case TPGX_PROBE_ATAN: out[i] = tpgx_math::atan(x); break;
        case TPGX_PROBE_SQRT_NONNEG: out[i] = std::sqrt(x); break;
        case TPGX_PROBE_DIV_SMALL_INT: out[i] = (double)(int)x / (double)(int)y; break;
        case TPGX_PROBE_SIN: out[i] = tpgx_math::sin(x); break;
        case TPGX_PROBE_COS: out[i] = tpgx_math::cos(x); break;
        case TPGX_PROBE_TAN: out[i] = tpgx_math::tan(x); break;
        case TPGX_PROBE_DIV10: case TPGX_PROBE_DIV10_K: out[i] = x / 10.0; break;
        default: out[i] = 0.0; break;
        }
    }
}

struct orc_env { std::unique_ptr<Env> env; };

orc_env* orc_env_create(int32_t kind, int32_t second_player_random, int32_t npa, const double* pa, int32_t math_flavour) {
    auto env = make_env(kind, second_player_random != 0, pa, npa, Math{math_flavour});
    if (!env) return nullptr;
    orc_env* v = new orc_env();
    v->env = std::move(env);
    return v;
}
void orc_env_destroy(orc_env* v) { delete v; }
void orc_env_reset(orc_env* v, uint64_t seed, int32_t mode) { v->env->reset(seed, mode); }
void orc_env_set_state(orc_env* v, const double* state, int32_t n) { v->env->set_state(state, n); }
int32_t orc_env_do_action(orc_env* v, double action) { return v->env->do_action(action) ? 0 : -1; }
int32_t orc_env_is_terminal(const orc_env* v) { return v->env->is_terminal() ? 1 : 0; }
double orc_env_score(const orc_env* v, int32_t player) { return v->env->score(player); }
int32_t orc_env_state(const orc_env* v, double* state, int32_t n) { return v->env->state(state, n); }

void orc_rng_u64(uint64_t seed, int64_t n, uint64_t lo, uint64_t hi, uint64_t* out) {
    Rng r; r.set_seed(seed);
    for (int64_t i = 0; i < n; i++) out[i] = r.u64(lo, hi);
}
void orc_rng_f64(uint64_t seed, int64_t n, double lo, double hi, double* out) {
    Rng r; r.set_seed(seed);
    for (int64_t i = 0; i < n; i++) out[i] = r.f64(lo, hi);
}
void orc_rng_raw(uint64_t seed, int64_t n, uint64_t* out) {
    std::mt19937_64 g(seed);
    for (int64_t i = 0; i < n; i++) out[i] = g();
}

/* The product's restatement of libstdc++'s distributions (tpgx_rng.h) driven by the same engine:
 * lets CPU tests pin it against the std:: distributions used above. */
void orc_rng_restated_u64(uint64_t seed, int64_t n, uint64_t lo, uint64_t hi, uint64_t* out) {
    std::vector<uint64_t> raw(TPGX_RNG_TABLE);
    std::mt19937_64 g(seed);
    for (auto& v : raw) v = g();
    tpgx_rng_cursor c{raw.data(), 0, 0};
    for (int64_t i = 0; i < n; i++) out[i] = tpgx_rng_u64(c, lo, hi);
}
void orc_rng_restated_f64(uint64_t seed, int64_t n, double lo, double hi, double* out) {
    std::vector<uint64_t> raw(TPGX_RNG_TABLE);
    std::mt19937_64 g(seed);
    for (auto& v : raw) v = g();
    tpgx_rng_cursor c{raw.data(), 0, 0};
    for (int64_t i = 0; i < n; i++) out[i] = tpgx_rng_f64(c, lo, hi);
}

} /* extern "C" */
