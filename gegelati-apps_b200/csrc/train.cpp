/*
 * train.cpp — host generation loop around the GPU evaluator.
 *
 * Restates [UP] Learn::LearningAgent::{init, trainOneGeneration} as the apps call them
 * ([REF] gridworld/src/main.cpp:33-34,57-65): initRandomTPG; per generation populateTPG -> evaluateAllRoots(TRAINING)
 * -> decimateWorstRoots, with [UP] Mutator::TPGMutator / ProgramMutator operators driven by the params.json keys
 * ([REF] gridworld/params.json:7-38).  The upstream library is not in /root/reference: the operators follow the
 * published TPG algorithm (Kelly & Heywood 2017; GEGELATI, Desnos et al. 2021), not upstream's draw order.
 * All evaluation and all program execution happen behind tpgx_train_backend (the GPU); this file is pure host policy:
 * graph bookkeeping, mutation, selection, archive ring.
 */
#include <cmath>
#include <chrono>
#include <cstring>
#include <deque>
#include <map>
#include <unordered_map>
#include <memory>
#include <random>

#include "tpgx_internal.h"

namespace {

struct TProg {
    std::vector<tpgx_line> lines;
    double consts[TPGX_MAX_CONSTS] = {0};
    int refs = 0;
    bool alive = false;
};
struct TEdge { int prog, dest; /* dest >= 0: team id; < 0: action = -1 - dest */ };
struct TTeam {
    std::vector<TEdge> edges;
    int incoming = 0;
    bool alive = false;
};
struct ARec { int prog; double bid; std::vector<double> obs; };
/* [UP] Learn::EvaluationResult / ClassificationEvaluationResult of one root, as LearningAgent::resultsPerRoot keeps it */
struct REval {
    double score = 0.0;
    uint64_t n = 0;                      /* nbEvaluation */
    std::vector<double> class_score;     /* classification: score per class */
    std::vector<uint64_t> class_n;       /* classification: nbEvaluationPerClass */
    /* [UP] EvaluationResult::operator+= — `*this` is the NEW evaluation, `o` the stored one: evaluation-count-weighted mean,
       in upstream's operation order.  Classification: per class, then the general score is the mean over the classes and
       nbEvaluation the sum (unverifiable upstream detail; the oracle-side test util restates the
       same function, so GPU-vs-oracle evolution parity does not depend on it). */
    void combine(const REval& o) {
        if (!class_score.empty() && class_score.size() == o.class_score.size()) {
            double sum = 0.0;
            uint64_t total = 0;
            for (size_t c = 0; c < class_score.size(); c++) {
                class_score[c] = class_score[c] * (double)class_n[c] + o.class_score[c] * (double)o.class_n[c];
                class_n[c] += o.class_n[c];
                if (class_n[c] != 0) class_score[c] /= (double)class_n[c];
                sum += class_score[c];
                total += class_n[c];
            }
            score = sum / (double)class_score.size();
            n = total;
            return;
        }
        score = score * (double)n + o.score * (double)o.n;
        score /= (double)n + (double)o.n;
        n += o.n;
    }
};
/* a program created by this populate: still to be mutated / behaviour-checked against its origin */
struct NewProg {
    int id;
    std::vector<tpgx_line> parent_lines;
    double parent_consts[TPGX_MAX_CONSTS];
    /* [UP] mutateNewProgramBehaviors gives every new program its own RNG, seeded from the agent's.  The 2.5 KB engine is made
       when the program is known to survive the populate (settle_new_programs), not copied around with the record */
    std::unique_ptr<std::mt19937_64> rng_store;
    std::mt19937_64& rng_ref() { return *rng_store; }
};

} // namespace

struct tpgx_trainer {
    tpgx_config cfg{};
    std::vector<int32_t> opcode;
    std::vector<tpgx_source_desc> src;
    tpgx_train_params P{};
    std::mt19937_64 rng;
    std::vector<TProg> progs;
    std::vector<int> free_progs;
    std::vector<TTeam> teams;
    std::vector<int> free_teams;
    std::vector<int> order;          /* alive teams in creation order: the graph's vertex / root iteration order */
    std::deque<ARec> archive;
    std::map<int, REval> results_per_root; /* [UP] LearningAgent::resultsPerRoot, keyed by team id */
    int best_team = -1;                    /* [UP] bestRoot */
    double best_team_score = 0.0;
    int obs_len = 0;
    uint32_t max_loc = 1;
    int nb_src_total = 0;
    std::vector<std::vector<int>> providers; /* per operand type: the source indices that can provide it */
    std::string err;
    /* flattened view */
    std::vector<int64_t> v_off;
    std::vector<tpgx_line> v_lines;
    std::vector<double> v_consts;
    std::vector<int32_t> v_teo, v_ep, v_ed, v_roots;
    std::vector<int> v_prog_id, v_root_team;

    /* mutate_team's index of the valid pre-existing edges (see there) */
    const void* pre_built = nullptr;
    uint64_t deaths = 0, pre_deaths = 0;
    std::vector<size_t> pre_valid;
    std::vector<int> pre_pos;
    std::vector<int> pre_bp_off, pre_bp_pos, pre_bp_fill; /* valid pre-existing edges by program: CSR over program ids -> positions in pre_valid */
    bool defer_archive = false;      /* set while no program id can be reused (decimation): unref_prog only notes the deleted programs */
    std::vector<int> dead_progs;
    std::vector<int64_t> s_off;        /* settle_new_programs' scratch graph and bids, kept between generations */
    std::vector<tpgx_line> s_lines;
    std::vector<double> s_consts, s_bid, s_obs, s_rbid, s_robs;
    std::vector<int32_t> s_rprog;
    void purge_archive() {
        if (!dead_progs.empty()) {
            std::vector<char> dead(progs.size(), 0);
            for (int id : dead_progs) dead[(size_t)id] = 1;
            std::deque<ARec> kept;
            for (ARec& r : archive)
                if (!dead[(size_t)r.prog]) kept.push_back(std::move(r));
            archive.swap(kept);
            dead_progs.clear();
        }
        defer_archive = false;
    }
    std::vector<uint32_t> prog_mark; /* mutate_team: programs already on the team carry the current stamp */
    uint32_t stamp = 0;
    std::vector<size_t> cand_scratch;
    /* the engine u64() / dbl() draw from: the agent's, or a new program's own while it is being mutated — per host thread, so that
       the new programs of a generation (an engine each) mutate in parallel on the pool */
    static std::mt19937_64*& cur() { static thread_local std::mt19937_64* e = nullptr; return e; }
    uint64_t u64(uint64_t lo, uint64_t hi) { return std::uniform_int_distribution<uint64_t>(lo, hi)(cur() ? *cur() : rng); } /* [UP] RNG::getUnsignedInt64 */
    double dbl() { return std::uniform_real_distribution<double>(0.0, 1.0)(cur() ? *cur() : rng); }                          /* [UP] RNG::getDouble(0, 1) */

    /* ---- bookkeeping ---- */
    int new_prog() {
        int id;
        if (!free_progs.empty()) { id = free_progs.back(); free_progs.pop_back(); }
        else { id = (int)progs.size(); progs.emplace_back(); }
        progs[id] = TProg();
        progs[id].alive = true;
        return id;
    }
    void unref_prog(int id) {
        if (--progs[id].refs > 0) return;
        progs[id].alive = false;
        deaths++;
        progs[id].lines.clear();
        free_progs.push_back(id);
        /* [UP] the archive forgets the recordings of deleted programs */
        if (defer_archive) { dead_progs.push_back(id); return; } /* decimation: one pass over the archive for all of them */
        for (size_t i = 0; i < archive.size();)
            if (archive[i].prog == id) archive.erase(archive.begin() + (long)i); else i++;
    }
    int new_team() {
        int id;
        if (!free_teams.empty()) { id = free_teams.back(); free_teams.pop_back(); }
        else { id = (int)teams.size(); teams.emplace_back(); }
        teams[id] = TTeam();
        teams[id].alive = true;
        order.push_back(id);
        return id;
    }
    void add_edge(int team, int prog, int dest) {
        teams[team].edges.push_back({prog, dest});
        progs[prog].refs++;
        if (dest >= 0) teams[dest].incoming++;
    }
    void remove_edge(int team, size_t idx) {
        const TEdge e = teams[team].edges[idx];
        teams[team].edges.erase(teams[team].edges.begin() + (long)idx);
        if (e.dest >= 0) teams[e.dest].incoming--;
        unref_prog(e.prog);
    }
    void remove_team(int team) {
        while (!teams[team].edges.empty()) remove_edge(team, teams[team].edges.size() - 1);
        teams[team].alive = false;
        results_per_root.erase(team); /* [UP] decimateWorstRoots: keep only results of non-decimated roots */
        order.erase(std::find(order.begin(), order.end(), team));
        free_teams.push_back(team);
    }
    int nb_action_edges(int team) const {
        int n = 0;
        for (const TEdge& e : teams[team].edges) n += e.dest < 0;
        return n;
    }
    int count_roots() const {
        int n = 0;
        for (int t : order) n += teams[t].incoming == 0;
        return n;
    }
    std::vector<int> roots() const {
        std::vector<int> r;
        for (int t : order) if (teams[t].incoming == 0) r.push_back(t);
        return r;
    }

    /* ---- [UP] ProgramMutator ---- */
    tpgx_line random_line() {
        tpgx_line l{};
        l.instruction = (uint32_t)u64(0, opcode.size() - 1);
        l.destination = (uint32_t)u64(0, (uint64_t)cfg.nb_registers - 1);
        const tpgx_op_info oi = tpgx_get_op_info(opcode[l.instruction]);
        for (int k = 0; k < 2; k++) {
            const int t = k < oi.nb_operands ? oi.type[k] : TPGX_T_D;
            const std::vector<int>& pv = providers[t];
            l.src[k] = (uint32_t)pv[u64(0, pv.size() - 1)];
            l.loc[k] = (uint32_t)u64(0, max_loc - 1);
        }
        return l;
    }
    void alter_line(tpgx_line& l) {
        /* re-draw one of: instruction, destination, operand 0, operand 1; the operands are re-validated against the
           (possibly new) instruction's operand types */
        const uint64_t what = u64(0, 3);
        if (what == 0) l.instruction = (uint32_t)u64(0, opcode.size() - 1);
        else if (what == 1) l.destination = (uint32_t)u64(0, (uint64_t)cfg.nb_registers - 1);
        const tpgx_op_info oi = tpgx_get_op_info(opcode[l.instruction]);
        for (int k = 0; k < 2; k++) {
            const int t = k < oi.nb_operands ? oi.type[k] : TPGX_T_D;
            const std::vector<int>& pv = providers[t];
            const bool redraw = (what == (uint64_t)(2 + k)) || std::find(pv.begin(), pv.end(), (int)l.src[k]) == pv.end();
            if (redraw) { l.src[k] = (uint32_t)pv[u64(0, pv.size() - 1)]; l.loc[k] = (uint32_t)u64(0, max_loc - 1); }
        }
    }
    void mutate_program(TProg& p) {
        bool any = false;
        do {
            if (p.lines.size() > 1 && dbl() < P.p_delete) { p.lines.erase(p.lines.begin() + (long)u64(0, p.lines.size() - 1)); any = true; }
            if ((int)p.lines.size() < P.max_program_size && dbl() < P.p_add) { p.lines.insert(p.lines.begin() + (long)u64(0, p.lines.size()), random_line()); any = true; }
            if (!p.lines.empty() && dbl() < P.p_mutate) { alter_line(p.lines[u64(0, p.lines.size() - 1)]); any = true; }
            if (p.lines.size() > 1 && dbl() < P.p_swap) {
                const size_t a = u64(0, p.lines.size() - 1);
                size_t b = u64(0, p.lines.size() - 2);
                if (b >= a) b++;
                std::swap(p.lines[a], p.lines[b]);
                any = true;
            }
            for (int c = 0; c < cfg.nb_constants; c++)
                if (dbl() < P.p_constant_mutation) {
                    p.consts[c] = (double)((int64_t)u64(0, (uint64_t)(P.max_const_value - P.min_const_value)) + P.min_const_value);
                    any = true;
                }
        } while (!any);
    }

    /* ---- [UP] TPGMutator ---- */
    void init_random_tpg() {
        const int nA = P.nb_actions;
        std::vector<TEdge> made;
        for (int i = 0; i < nA; i++) {
            const int t = new_team();
            for (int k = 0; k < 2; k++) { /* two edges to two different actions */
                const int p = new_prog();
                progs[p].lines.push_back(random_line());
                for (int c = 0; c < cfg.nb_constants; c++)
                    progs[p].consts[c] = (double)((int64_t)u64(0, (uint64_t)(P.max_const_value - P.min_const_value)) + P.min_const_value);
                mutate_program(progs[p]);
                add_edge(t, p, -1 - ((i + k) % nA));
                made.push_back({p, -1 - ((i + k) % nA)});
            }
        }
        /* extra initial edges up to maxInitOutgoingEdges: duplicates of edges made above (same program, same action),
           so that a program always bids for one destination — what the dot format can express */
        for (int t : std::vector<int>(order)) {
            const int extra = P.max_init_outgoing_edges > 2 ? (int)u64(0, (uint64_t)P.max_init_outgoing_edges - 2) : 0;
            for (int x = 0; x < extra; x++) {
                const TEdge m = made[u64(0, made.size() - 1)];
                bool has = false;
                for (const TEdge& e : teams[t].edges) has |= e.prog == m.prog;
                if (!has) add_edge(t, m.prog, m.dest);
            }
        }
    }
    void mutate_team(int team, const std::vector<TEdge>& pre_edges, const std::vector<int>& pre_teams, std::vector<NewProg>& fresh) {
        TTeam& T = teams[team];
        while (T.edges.size() > 2 && dbl() < P.p_edge_deletion) {
            size_t idx = u64(0, T.edges.size() - 1);
            if (T.edges[idx].dest < 0 && nb_action_edges(team) == 1) { /* a team always keeps an action edge */
                size_t other = u64(0, T.edges.size() - 2);
                idx = other >= idx ? other + 1 : other;
            }
            remove_edge(team, idx);
        }
        while ((int)T.edges.size() < P.max_outgoing_edges && dbl() < P.p_edge_addition) {
            /* candidates, in pre_edges order: the still valid pre-existing edges (program alive, destination team alive)
               whose program is not on this team yet.  The valid ones are indexed once (and again after a program died,
               which is rare while populating); per addition only the few edges of the team's programs are taken out
               — scanning all pre-existing edges here used to be the bulk of a generation's host time. */
            if (pre_built != &pre_edges || pre_deaths != deaths) {
                pre_built = &pre_edges; pre_deaths = deaths;
                /* ... and grouped by program as a CSR index (offsets per program id, positions in pre_valid): an ordered map of
                   vectors here was 0.3 ms of a 1.2 ms populate */
                pre_valid.clear(); pre_pos.assign(pre_edges.size(), -1);
                pre_bp_off.assign(progs.size() + 1, 0);
                for (size_t i = 0; i < pre_edges.size(); i++) {
                    const TEdge& pe = pre_edges[i];
                    if (progs[pe.prog].alive && (pe.dest < 0 || teams[pe.dest].alive)) {
                        pre_pos[i] = (int)pre_valid.size();
                        pre_valid.push_back(i);
                        pre_bp_off[(size_t)pe.prog + 1]++;
                    }
                }
                for (size_t p = 0; p < progs.size(); p++) pre_bp_off[p + 1] += pre_bp_off[p];
                pre_bp_pos.resize(pre_valid.size());
                pre_bp_fill.assign(pre_bp_off.begin(), pre_bp_off.end() - 1);
                for (size_t v = 0; v < pre_valid.size(); v++) pre_bp_pos[(size_t)pre_bp_fill[(size_t)pre_edges[pre_valid[v]].prog]++] = (int)v;
            }
            if (prog_mark.size() < progs.size()) prog_mark.resize(progs.size() * 2, 0u);
            stamp++;
            std::vector<size_t>& out = cand_scratch; /* positions (in pre_valid) that are NOT candidates */
            out.clear();
            for (const TEdge& e : T.edges) {
                if (prog_mark[(size_t)e.prog] == stamp) continue;
                prog_mark[(size_t)e.prog] = stamp;
                if ((size_t)e.prog + 1 < pre_bp_off.size()) /* a program made after the index has no pre-existing edge */
                    for (int k = pre_bp_off[(size_t)e.prog]; k < pre_bp_off[(size_t)e.prog + 1]; k++) out.push_back((size_t)pre_bp_pos[(size_t)k]);
            }
            if (pre_valid.size() == out.size()) break;
            size_t r = u64(0, pre_valid.size() - out.size() - 1); /* r-th candidate = r-th valid position not taken out */
            std::sort(out.begin(), out.end());
            for (size_t x : out) { if (x <= r) r++; else break; }
            const TEdge& e = pre_edges[pre_valid[r]];
            add_edge(team, e.prog, e.dest);
        }
        bool any = false;
        do {
            for (size_t i = 0; i < T.edges.size(); i++) {
                if (!(dbl() < P.p_program_mutation)) continue;
                any = true;
                const int old = T.edges[i].prog;
                const int np = new_prog();
                progs[np].lines = progs[old].lines;
                std::memcpy(progs[np].consts, progs[old].consts, sizeof progs[np].consts);
                progs[np].refs = 1;
                NewProg rec;
                rec.id = np;
                rec.parent_lines = progs[old].lines;
                std::memcpy(rec.parent_consts, progs[old].consts, sizeof rec.parent_consts);
                fresh.push_back(std::move(rec));
                T.edges[i].prog = np;
                unref_prog(old);
                if (dbl() < P.p_edge_destination_change) {
                    const bool only_action = T.edges[i].dest < 0 && nb_action_edges(team) == 1;
                    int dest;
                    if (only_action || pre_teams.empty() || dbl() < P.p_edge_destination_is_action) dest = -1 - (int)u64(0, (uint64_t)P.nb_actions - 1);
                    else dest = pre_teams[u64(0, pre_teams.size() - 1)];
                    if (T.edges[i].dest >= 0) teams[T.edges[i].dest].incoming--;
                    T.edges[i].dest = dest;
                    if (dest >= 0) teams[dest].incoming++;
                }
            }
        } while (!any);
    }

    /* ---- view in the C ABI's layout ---- */
    /* skip (optional, by team id): root teams left out of the view's root list — the jobs [UP] evaluateJob would skip */
    void build_view(tpgx_graph* g, const std::vector<char>* skip = nullptr) {
        std::vector<int> pidx(progs.size(), -1);
        v_prog_id.clear();
        for (size_t p = 0; p < progs.size(); p++)
            if (progs[p].alive && progs[p].refs > 0) { pidx[p] = (int)v_prog_id.size(); v_prog_id.push_back((int)p); }
        const int nP = (int)v_prog_id.size(), nC = cfg.nb_constants;
        v_off.assign((size_t)nP + 1, 0);
        v_lines.clear();
        v_consts.assign((size_t)nP * (size_t)std::max(1, nC), 0.0);
        for (int i = 0; i < nP; i++) {
            const TProg& p = progs[v_prog_id[i]];
            v_lines.insert(v_lines.end(), p.lines.begin(), p.lines.end());
            v_off[i + 1] = (int64_t)v_lines.size();
            for (int c = 0; c < nC; c++) v_consts[(size_t)i * nC + c] = p.consts[c];
        }
        std::vector<int> tidx(teams.size(), -1);
        for (size_t i = 0; i < order.size(); i++) tidx[order[i]] = (int)i;
        v_teo.assign(1, 0);
        v_ep.clear(); v_ed.clear(); v_roots.clear(); v_root_team.clear();
        for (int t : order) {
            for (const TEdge& e : teams[t].edges) { v_ep.push_back(pidx[e.prog]); v_ed.push_back(e.dest >= 0 ? tidx[e.dest] : e.dest); }
            v_teo.push_back((int32_t)v_ep.size());
            if (teams[t].incoming == 0 && !(skip && (*skip)[(size_t)t])) { v_roots.push_back(tidx[t]); v_root_team.push_back(t); }
        }
        g->nb_programs = nP;
        g->program_line_offset = v_off.data();
        g->lines = v_lines.data();
        g->intron = nullptr;
        g->constants = nC > 0 ? v_consts.data() : nullptr;
        g->nb_teams = (int32_t)order.size();
        g->team_edge_offset = v_teo.data();
        g->edge_program = v_ep.data();
        g->edge_destination = v_ed.data();
        g->nb_roots = (int32_t)v_roots.size();
        g->root_team = v_roots.data();
    }

    /* ---- [UP] mutateNewProgramBehaviors: mutate every new program; with forceProgramBehaviorChangeOnMutation, again
       until it behaves differently from its origin and from every archived program on the archived observations ---- */
    tpgx_status settle_new_programs(std::vector<NewProg>& fresh, const tpgx_train_backend* be, int* retries) {
        /* every new program mutates with an engine of its own, so its chain of mutants does not depend on how the
           other programs fare: the behaviour test below can try several mutants of a program in ONE backend call
           and keep the first that passes — same result as trying them one call at a time */
        static const bool trace = getenv("TPGX_TRACE") != nullptr;
        auto tick = std::chrono::steady_clock::now();
        auto lap = [&](const char* what) {
            if (!trace) return;
            const auto t = std::chrono::steady_clock::now();
            fprintf(stderr, "[tpgx] settle: %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - tick).count());
            tick = t;
        };
        std::vector<uint64_t> seeds(fresh.size()); /* drawn in order from the agent's engine; the engines are made in parallel below */
        for (uint64_t& sd : seeds) sd = u64(0, UINT64_MAX);
        struct Scope { Scope(tpgx_trainer*, std::mt19937_64* e) { cur() = e; } ~Scope() { cur() = nullptr; } };
        /* every program draws from its own engine and touches its own entry of progs: independent of the others, of their order
           and of the number of threads */
        tpgx_parallel_blocks((int)fresh.size(), 16, [&](int b, int e, int) {
            for (int i = b; i < e; i++) {
                NewProg& n = fresh[(size_t)i];
                n.rng_store.reset(new std::mt19937_64(seeds[(size_t)i]));
                Scope sc(this, &n.rng_ref());
                mutate_program(progs[n.id]);
            }
        });
        lap("seed + first mutation");
        if (!P.force_program_behavior_change || archive.empty() || fresh.empty()) return TPGX_OK;
        if (!be || !be->execute) { err = "forceProgramBehaviorChangeOnMutation needs a backend to run the new programs on the archive"; return TPGX_ERR_BAD_ARG; }
        /* the distinct archived observations ([UP] the archive keys its data-handler copies by hash) and, per archived
           program, its recorded bid on each of them */
        struct ObsHash { /* keyed by the bytes of the doubles, as [UP] the archive hashes its data-handler copies (+0.0 and -0.0 are two observations) */
            size_t operator()(const std::vector<double>& v) const {
                uint64_t h = 1469598103934665603ull;
                for (double d : v) { uint64_t b; std::memcpy(&b, &d, 8); h = (h ^ b) * 1099511628211ull; h ^= h >> 29; }
                return (size_t)h;
            }
        };
        struct ObsEq {
            bool operator()(const std::vector<double>& a, const std::vector<double>& b) const {
                return a.size() == b.size() && (a.empty() || std::memcmp(a.data(), b.data(), a.size() * 8) == 0);
            }
        };
        std::unordered_map<std::vector<double>, int, ObsHash, ObsEq> obs_index;
        obs_index.reserve(archive.size() * 2);
        std::vector<double>& obs = s_obs;
        obs.clear();
        std::map<int, std::map<int, double>> by_prog;
        for (const ARec& r : archive) {
            auto it = obs_index.find(r.obs);
            if (it == obs_index.end()) {
                it = obs_index.emplace(r.obs, (int)obs_index.size()).first;
                obs.insert(obs.end(), r.obs.begin(), r.obs.end());
            }
            by_prog[r.prog][it->second] = r.bid;
        }
        const int64_t count = (int64_t)obs_index.size();
        /* archived programs recorded on every archived observation: the only ones that can make a candidate a duplicate */
        std::vector<std::vector<double>> full;
        for (const auto& kv : by_prog)
            if ((int64_t)kv.second.size() == count) {
                full.emplace_back((size_t)count);
                for (const auto& ob : kv.second) full.back()[(size_t)ob.first] = ob.second;
            }
        lap("archive index");
        const double tau = 1e-4; /* [UP] Archive::areProgramResultsUnique default tolerance */
        auto close = [&](double a, double b) { return a == b || std::fabs(a - b) <= tau; };
        const int max_attempts = 16, nC = cfg.nb_constants;
        std::vector<size_t> pending(fresh.size());
        for (size_t i = 0; i < fresh.size(); i++) pending[i] = i;
        /* attempts tried per backend call: 4 first (most programs pass within a few), then all that are left */
        for (int done = 0; done < max_attempts && !pending.empty();) {
            int S = done == 0 ? 4 : max_attempts - done;
            if (const char* v = getenv("TPGX_TRAIN_ATTEMPTS_PER_CALL")) { const int f = atoi(v); if (f >= 1) S = std::min(f, max_attempts - done); } /* tests: 1 = one mutant per call */
            const int M = (int)pending.size();
            /* candidates of pending program i: c[0] = its current mutant, c[k] = one more mutation of c[k-1] */
            std::vector<TProg> cand((size_t)M * S);
            tpgx_parallel_blocks(M, 16, [&](int b, int e, int) {
                for (int i = b; i < e; i++) {
                    NewProg& n = fresh[pending[(size_t)i]];
                    Scope sc(this, &n.rng_ref());
                    cand[(size_t)i * S] = progs[n.id];
                    for (int k = 1; k < S; k++) { cand[(size_t)i * S + k] = cand[(size_t)i * S + k - 1]; mutate_program(cand[(size_t)i * S + k]); }
                }
            });
            lap("candidates");
            /* a scratch graph: per pending program its S candidates then its origin, one team pointing at all of them */
            const int per = S + 1, NP = M * per;
            /* the big scratch arrays (lines, bids: tens of MB) live in the trainer and keep their capacity: fresh pages every call cost
               page faults, and on the virtualised GPU boxes a burst of first-touch faults now and then stalls for 0.5 - 0.8 s */
            std::vector<int64_t>& off = s_off;
            std::vector<tpgx_line>& lines = s_lines;
            std::vector<double>& consts = s_consts;
            off.assign(1, 0);
            lines.clear();
            consts.assign((size_t)NP * (size_t)std::max(1, nC), 0.0);
            std::vector<int32_t> teo = {0, NP}, ep((size_t)NP), ed((size_t)NP, -1), root = {0};
            for (int i = 0; i < M; i++) {
                const NewProg& n = fresh[pending[i]];
                for (int k = 0; k < S; k++) {
                    const TProg& c = cand[(size_t)i * S + k];
                    lines.insert(lines.end(), c.lines.begin(), c.lines.end());
                    off.push_back((int64_t)lines.size());
                    for (int q = 0; q < nC; q++) consts[(size_t)(i * per + k) * nC + q] = c.consts[q];
                }
                lines.insert(lines.end(), n.parent_lines.begin(), n.parent_lines.end());
                off.push_back((int64_t)lines.size());
                for (int q = 0; q < nC; q++) consts[(size_t)(i * per + S) * nC + q] = n.parent_consts[q];
            }
            for (int q = 0; q < NP; q++) ep[q] = q;
            tpgx_graph g{NP, off.data(), lines.data(), nullptr, nC > 0 ? consts.data() : nullptr, 1, teo.data(), ep.data(), ed.data(), 1, root.data()};
            std::vector<double>& bid = s_bid;
            bid.resize((size_t)NP * (size_t)count);
            lap("scratch graph");
            const tpgx_status st = be->execute(be->user, &g, count, obs.data(), obs_len, bid.data());
            if (st != TPGX_OK) { err = "backend execute failed"; return st; }
            lap("backend execute");
            std::vector<size_t> again;
            for (int i = 0; i < M; i++) {
                const double* bp = bid.data() + (size_t)(i * per + S) * count; /* the origin's bids */
                int pick = -1;
                for (int k = 0; k < S && pick < 0; k++) {
                    const double* bm = bid.data() + (size_t)(i * per + k) * count;
                    bool differs = false;
                    for (int64_t q = 0; q < count && !differs; q++) differs = !close(bm[q], bp[q]);
                    /* [UP] areProgramResultsUnique: an archived program with a recording on every archived observation and
                       all of them within tau of the candidate makes the candidate a duplicate */
                    bool unique = differs;
                    for (size_t f = 0; unique && f < full.size(); f++) {
                        bool same = true;
                        for (int64_t q = 0; q < count && same; q++) same = close(bm[q], full[f][(size_t)q]);
                        unique = !same;
                    }
                    if (unique) pick = k; else (*retries)++;
                }
                NewProg& n = fresh[pending[i]];
                if (pick >= 0) progs[n.id] = cand[(size_t)i * S + pick];
                else { /* all S failed: one more mutation of the last one opens the next call (a retry like the others) */
                    progs[n.id] = cand[(size_t)i * S + S - 1];
                    Scope sc(this, &n.rng_ref());
                    mutate_program(progs[n.id]);
                    again.push_back(pending[i]);
                }
            }
            pending.swap(again);
            done += S;
            lap("uniqueness checks");
        }
        return TPGX_OK;
    }
};

/* ================================================================================================ C ABI */
extern "C" {

tpgx_status tpgx_trainer_create(const tpgx_config* cfg, int32_t nb_instructions, const int32_t* opcode, int32_t nb_sources,
                                const tpgx_source_desc* sources, const tpgx_train_params* params, uint64_t seed, tpgx_trainer** out) {
    if (!cfg || !opcode || nb_instructions < 1 || !sources || nb_sources < 1 || nb_sources > 2 || !params || !out) return TPGX_ERR_BAD_ARG;
    if (cfg->nb_registers < 1 || cfg->nb_registers > TPGX_MAX_REGS || cfg->nb_constants < 0 || cfg->nb_constants > TPGX_MAX_CONSTS) return TPGX_ERR_UNSUPPORTED;
    const tpgx_train_params& P = *params;
    if (P.nb_actions < 2 || P.nb_actions > 128 || P.nb_roots < P.nb_actions || P.max_outgoing_edges < 2 || P.max_init_outgoing_edges < 2 ||
        P.max_init_outgoing_edges > P.max_outgoing_edges || P.max_program_size < 1 || P.max_const_value < P.min_const_value ||
        !(P.ratio_deleted_roots >= 0.0 && P.ratio_deleted_roots < 1.0) || P.archive_size < 0)
        return TPGX_ERR_BAD_ARG;
    for (int i = 0; i < nb_instructions; i++)
        if (opcode[i] < 0 || opcode[i] >= TPGX_OP_COUNT) return TPGX_ERR_UNSUPPORTED;
    tpgx_trainer* t = new tpgx_trainer();
    t->cfg = *cfg;
    t->opcode.assign(opcode, opcode + nb_instructions);
    t->src.assign(sources, sources + nb_sources);
    t->P = P;
    t->rng.seed(seed);
    t->nb_src_total = (cfg->nb_constants > 0 ? 2 : 1) + nb_sources;
    t->providers.assign(TPGX_T_W + 1, std::vector<int>());
    for (int ty = 0; ty <= TPGX_T_W; ty++)
        for (int s = 0; s < t->nb_src_total; s++) {
            const uint32_t sp = tpgx_source_space(t->cfg, t->src, s, ty);
            if (sp) { t->providers[ty].push_back(s); t->max_loc = std::max(t->max_loc, sp); }
        }
    for (int i = 0; i < nb_instructions; i++) { /* every instruction must be satisfiable, else random lines cannot be valid */
        const tpgx_op_info oi = tpgx_get_op_info(opcode[i]);
        for (int k = 0; k < oi.nb_operands; k++)
            if (t->providers[oi.type[k]].empty()) { delete t; return TPGX_ERR_UNSUPPORTED; }
    }
    for (const tpgx_source_desc& d : t->src) t->obs_len += d.height * d.width;
    t->init_random_tpg();
    *out = t;
    return TPGX_OK;
}

void tpgx_trainer_destroy(tpgx_trainer* t) { delete t; }
const char* tpgx_trainer_last_error(const tpgx_trainer* t) { return t ? t->err.c_str() : "null trainer"; }

tpgx_status tpgx_trainer_graph(tpgx_trainer* t, tpgx_graph* view) {
    if (!t || !view) return TPGX_ERR_BAD_ARG;
    t->build_view(view);
    return TPGX_OK;
}

static void fill_graph_stats(tpgx_trainer* t, tpgx_train_stats* s) {
    if (!s) return;
    int np = 0;
    for (const TProg& p : t->progs) np += p.alive && p.refs > 0;
    s->nb_teams = (int32_t)t->order.size();
    s->nb_programs = np;
    s->nb_roots = (int32_t)t->roots().size();
    s->archive_records = (int32_t)t->archive.size();
}

tpgx_status tpgx_trainer_populate(tpgx_trainer* t, const tpgx_train_backend* backend, tpgx_train_stats* stats) {
    if (!t) return TPGX_ERR_BAD_ARG;
    /* [UP] populateTPG: the roots present now are the parents; edges and teams present now are what mutations may reuse */
    const std::vector<int> parents = t->roots();
    if (parents.empty()) { t->err = "no root left to clone"; return TPGX_ERR_GRAPH_INVALID; }
    std::vector<TEdge> pre_edges;
    for (int tm : t->order) for (const TEdge& e : t->teams[tm].edges) pre_edges.push_back(e);
    const std::vector<int> pre_teams = t->order;
    t->pre_built = nullptr;
    std::vector<NewProg> fresh;
    fresh.reserve(1024);
    int nb_roots = (int)parents.size();
    static const bool trace = getenv("TPGX_TRACE") != nullptr;
    const auto c0 = std::chrono::steady_clock::now();
    while (nb_roots < t->P.nb_roots) {
        const int parent = parents[t->u64(0, parents.size() - 1)];
        const int child = t->new_team();
        for (const TEdge e : std::vector<TEdge>(t->teams[parent].edges)) t->add_edge(child, e.prog, e.dest);
        t->mutate_team(child, pre_edges, pre_teams, fresh);
        nb_roots = t->count_roots(); /* a destination change may have turned a parent into an inner team */
    }
    if (trace) fprintf(stderr, "[tpgx] populate: clone + mutate teams %.3f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
    /* programs created and dropped again within this populate are gone already */
    std::vector<NewProg> live;
    for (NewProg& n : fresh) if (t->progs[n.id].alive && t->progs[n.id].refs > 0) live.push_back(std::move(n));
    int retries = 0;
    const tpgx_status st = t->settle_new_programs(live, backend, &retries);
    if (st != TPGX_OK) return st;
    if (stats) { fill_graph_stats(t, stats); stats->nb_new_programs = (int32_t)live.size(); stats->behaviour_retries = retries; }
    return TPGX_OK;
}

tpgx_status tpgx_trainer_decimate(tpgx_trainer* t, const double* root_scores, tpgx_train_stats* stats) {
    if (!t || !root_scores) return TPGX_ERR_BAD_ARG;
    const std::vector<int> roots = t->roots();
    /* [UP] decimateWorstRoots: the floor(ratioDeletedRoots * nbRoots) lowest-scored roots go; equal scores keep root order */
    std::vector<int> idx(roots.size());
    for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return root_scores[a] < root_scores[b]; });
    int kill = (int)std::floor(t->P.ratio_deleted_roots * (double)t->P.nb_roots);
    kill = std::min(kill, (int)roots.size() - 1);
    static const bool trace = getenv("TPGX_TRACE") != nullptr;
    const auto c0 = std::chrono::steady_clock::now();
    t->defer_archive = true;
    for (int i = 0; i < kill; i++) t->remove_team(roots[idx[i]]);
    t->purge_archive();
    if (trace) fprintf(stderr, "[tpgx] decimate: remove %d teams %.3f ms\n", kill, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
    if (stats) {
        fill_graph_stats(t, stats);
        stats->nb_removed_roots = kill;
        double best = -INFINITY, worst = INFINITY, sum = 0.0;
        for (size_t i = 0; i < roots.size(); i++) { best = std::max(best, root_scores[i]); worst = std::min(worst, root_scores[i]); sum += root_scores[i]; }
        stats->best_score = best; stats->worst_score = worst; stats->mean_score = sum / (double)roots.size();
    }
    return TPGX_OK;
}

tpgx_status tpgx_trainer_train_generation(tpgx_trainer* t, const tpgx_train_backend* be, uint64_t generation, tpgx_train_stats* stats) {
    if (!t || !be || !be->evaluate) return TPGX_ERR_BAD_ARG;
    tpgx_train_stats s{};
    static const bool trace = getenv("TPGX_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    tpgx_status st = tpgx_trainer_populate(t, be, &s);
    if (st != TPGX_OK) return st;
    const auto t1 = now();
    /* [UP] evaluateAllRoots(generation, TRAINING): makeJobs draws one archive seed per root, in root order; evaluateJob then
       returns the stored result of a root that already has maxNbEvaluationPerPolicy evaluations without running it */
    const std::vector<int> all_roots = t->roots();
    const int R = (int)all_roots.size(), cap = t->P.archive_size, C = std::max(0, t->P.nb_classes);
    std::vector<uint64_t> aseed_all((size_t)R);
    for (int j = 0; j < R; j++) aseed_all[j] = t->u64(0, UINT64_MAX); /* [UP] Job::archiveSeed drawn from the agent's RNG */
    std::vector<char> skip(t->teams.size(), 0);
    std::vector<uint64_t> aseed;
    std::vector<int> job_root; /* index in all_roots of every job */
    for (int j = 0; j < R; j++) {
        const auto it = t->results_per_root.find(all_roots[j]);
        const bool skipped = it != t->results_per_root.end() && t->P.max_nb_evaluation_per_policy > 0 &&
                             it->second.n >= (uint64_t)t->P.max_nb_evaluation_per_policy;
        if (skipped) skip[(size_t)all_roots[j]] = 1;
        else { aseed.push_back(aseed_all[j]); job_root.push_back(j); }
    }
    const int J = (int)job_root.size();
    tpgx_graph g{};
    t->build_view(&g, &skip);
    /* the archive buffers (archive_size x observation length doubles: 3 MB for the mnist app) are the trainer's and keep their pages */
    std::vector<double> scores((size_t)std::max(1, J), 0.0);
    std::vector<double>&rbid = t->s_rbid, &robs = t->s_robs;
    std::vector<int32_t>& rprog = t->s_rprog;
    rbid.resize((size_t)std::max(1, cap));
    robs.resize((size_t)std::max(1, cap) * (size_t)std::max(1, t->obs_len));
    rprog.resize((size_t)std::max(1, cap));
    std::vector<double> cscore((size_t)std::max(1, J * C));
    std::vector<uint64_t> ccount((size_t)std::max(1, J * C));
    int32_t nrec = 0;
    if (J > 0) {
        st = be->evaluate(be->user, &g, generation, 0 /* TRAINING */, aseed.data(), t->P.archiving_probability, cap, scores.data(), C,
                          C ? cscore.data() : nullptr, C ? ccount.data() : nullptr, rprog.data(), rbid.data(), robs.data(), t->obs_len, &nrec);
        if (st != TPGX_OK) { t->err = "backend evaluate failed"; return st; }
    }
    const auto t2 = now();
    if (nrec < 0 || nrec > std::max(0, cap)) { t->err = "backend returned more archive records than archive_size"; return TPGX_ERR_STATE; }
    for (int i = 0; i < nrec; i++) {
        if (rprog[i] < 0 || rprog[i] >= (int32_t)t->v_prog_id.size()) { t->err = "backend returned an archive record for an unknown program"; return TPGX_ERR_STATE; }
        ARec r;
        r.prog = t->v_prog_id[rprog[i]];
        r.bid = rbid[i];
        r.obs.assign(robs.begin() + (long)i * t->obs_len, robs.begin() + (long)(i + 1) * t->obs_len);
        t->archive.push_back(std::move(r));
    }
    while ((int)t->archive.size() > cap) t->archive.pop_front();
    /* [UP] evaluateJob: the new result, combined with the stored one if any; [UP] updateEvaluationRecords stores it */
    for (int k = 0; k < J; k++) {
        const int team = all_roots[job_root[k]];
        REval e;
        e.score = scores[k];
        e.n = (uint64_t)std::max(1, t->P.nb_iterations_per_policy_evaluation);
        if (C) {
            e.class_score.assign(cscore.begin() + (long)k * C, cscore.begin() + (long)(k + 1) * C);
            e.class_n.assign(ccount.begin() + (long)k * C, ccount.begin() + (long)(k + 1) * C);
            e.n = 0;
            for (uint64_t v : e.class_n) e.n += v;
        }
        const auto it = t->results_per_root.find(team);
        if (it != t->results_per_root.end()) e.combine(it->second);
        t->results_per_root[team] = std::move(e);
    }
    std::vector<double> all_scores((size_t)R);
    for (int j = 0; j < R; j++) all_scores[j] = t->results_per_root[all_roots[j]].score;
    /* [UP] updateEvaluationRecords / bestRoot: the best result of this generation (the last of equal scores in the
       reference's multimap order, i.e. the latest root) replaces the stored best when it is at least as good or the stored
       best has left the graph */
    {
        int bj = 0;
        for (int j = 1; j < R; j++) if (all_scores[j] >= all_scores[bj]) bj = j;
        const bool have = t->best_team >= 0 && t->best_team < (int)t->teams.size() && t->teams[(size_t)t->best_team].alive;
        if (!have || all_scores[bj] >= t->best_team_score) { t->best_team = all_roots[bj]; t->best_team_score = all_scores[bj]; }
    }
    tpgx_train_stats d{};
    st = tpgx_trainer_decimate(t, all_scores.data(), &d);
    if (st != TPGX_OK) return st;
    const auto t3 = now();
    /* [UP] trainOneGeneration: if (params.doValidation) evaluateAllRoots(generation, VALIDATION) on what is left; the results
       go to the loggers, they are neither stored nor combined (previousEval is only looked up in TRAINING mode) */
    double vbest = NAN, vmean = NAN;
    if (t->P.do_validation) {
        tpgx_graph gv{};
        t->build_view(&gv);
        const int RV = gv.nb_roots;
        std::vector<double> vs((size_t)std::max(1, RV)), vcs((size_t)std::max(1, RV * C));
        std::vector<uint64_t> vcc((size_t)std::max(1, RV * C));
        int32_t vrec = 0;
        st = be->evaluate(be->user, &gv, generation, 1 /* VALIDATION */, nullptr, 0.0, 0, vs.data(), C, C ? vcs.data() : nullptr, C ? vcc.data() : nullptr,
                          rprog.data(), rbid.data(), robs.data(), t->obs_len, &vrec);
        if (st != TPGX_OK) { t->err = "backend evaluate (validation) failed"; return st; }
        vbest = -INFINITY; vmean = 0.0;
        for (int j = 0; j < RV; j++) { vbest = std::max(vbest, vs[j]); vmean += vs[j]; }
        vmean /= (double)std::max(1, RV);
    }
    if (trace) fprintf(stderr, "[tpgx] generation %llu: populate %.3f ms (%d new programs, %d behaviour retries), evaluate + archive %.3f ms (%d jobs, %d skipped), "
                       "archive merge + decimate %.3f ms, validation %.3f ms\n",
                       (unsigned long long)generation, ms(t0, t1), s.nb_new_programs, s.behaviour_retries, ms(t1, t2), J, R - J, ms(t2, t3), ms(t3, now()));
    if (stats) {
        *stats = d;
        stats->nb_new_programs = s.nb_new_programs;
        stats->behaviour_retries = s.behaviour_retries;
        stats->nb_roots = R; stats->nb_teams = s.nb_teams; stats->nb_programs = s.nb_programs; /* sizes as evaluated */
        stats->nb_evaluated_roots = J; stats->nb_skipped_roots = R - J;
        stats->best_validation_score = vbest; stats->mean_validation_score = vmean;
    }
    return TPGX_OK;
}

int32_t tpgx_trainer_root_results(tpgx_trainer* t, double* score, uint64_t* nb_evaluation, int32_t capacity) {
    if (!t) return 0;
    const std::vector<int> roots = t->roots();
    for (int j = 0; j < (int)roots.size() && j < capacity; j++) {
        const auto it = t->results_per_root.find(roots[j]);
        if (score) score[j] = it == t->results_per_root.end() ? NAN : it->second.score;
        if (nb_evaluation) nb_evaluation[j] = it == t->results_per_root.end() ? 0 : it->second.n;
    }
    return (int32_t)roots.size();
}

tpgx_status tpgx_trainer_best_root(tpgx_trainer* t, int32_t* root_index, double* score) {
    if (!t) return TPGX_ERR_BAD_ARG;
    if (t->best_team < 0 || t->best_team >= (int)t->teams.size() || !t->teams[(size_t)t->best_team].alive) { t->err = "no best root recorded"; return TPGX_ERR_STATE; }
    const std::vector<int> roots = t->roots();
    const auto it = std::find(roots.begin(), roots.end(), t->best_team);
    if (it == roots.end()) { t->err = "the best root is no longer a root"; return TPGX_ERR_STATE; }
    if (root_index) *root_index = (int32_t)(it - roots.begin());
    if (score) *score = t->best_team_score;
    return TPGX_OK;
}

tpgx_status tpgx_trainer_keep_best_policy(tpgx_trainer* t) {
    if (!t) return TPGX_ERR_BAD_ARG;
    if (t->best_team < 0 || t->best_team >= (int)t->teams.size() || !t->teams[(size_t)t->best_team].alive) { t->err = "no best root recorded"; return TPGX_ERR_STATE; }
    /* [UP] keepBestPolicy: remove every root that is not the best one, again and again (removals uncover new roots) */
    t->defer_archive = true;
    for (;;) {
        const std::vector<int> roots = t->roots();
        bool removed = false;
        for (int r : roots)
            if (r != t->best_team) { t->remove_team(r); removed = true; }
        if (!removed) break;
    }
    t->purge_archive();
    return TPGX_OK;
}

uint64_t tpgx_trainer_fingerprint(const tpgx_trainer* t) {
    if (!t) return 0;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { for (int i = 0; i < 8; i++) { h ^= (v >> (8 * i)) & 0xff; h *= 1099511628211ull; } };
    for (int tm : t->order) {
        mix(0xfeedull + (uint64_t)tm);
        for (const TEdge& e : t->teams[tm].edges) {
            mix((uint64_t)(int64_t)e.dest);
            const TProg& p = t->progs[e.prog];
            mix(p.lines.size());
            for (const tpgx_line& l : p.lines) { mix(l.instruction | ((uint64_t)l.destination << 32)); mix(l.src[0] | ((uint64_t)l.loc[0] << 32)); mix(l.src[1] | ((uint64_t)l.loc[1] << 32)); }
            for (int c = 0; c < t->cfg.nb_constants; c++) { uint64_t b; std::memcpy(&b, &p.consts[c], 8); mix(b); }
        }
    }
    return h;
}

} /* extern "C" */
