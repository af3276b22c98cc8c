/*
 * flatten.cpp — host flattener: reference-format programs + TPG -> packed device bytecode + CSR.
 *
 * Replaces the per-line work [UP] ProgramExecutionEngine::fetchCurrentOperands does at run time
 * (type lookup, handler lookup, location % addressSpace) and the intron skipping of
 * [UP] ProgramExecutionEngine::iterateThroughtProgram by a one-off pass per graph upload:
 *   - [UP] Program::identifyIntrons (backward liveness from register 0),
 *   - static reduction of every location,
 *   - "never written => 0.0" register reads marked so the device needs no register clear.
 */
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>

#include "k1v2_encode.h"
#include "tpgx_internal.h"

namespace {

/* Worker pool behind tpgx_pool_run: up to 15 sleeping threads, created on first use, joined at unload.
   One job at a time (callers of the library hold one context per thread; concurrent callers serialise here). */
class Pool {
    std::mutex run_mutex, m;
    std::condition_variable wake, idle;
    std::vector<std::thread> workers;
    void (*fn)(int, void*) = nullptr;
    void* arg = nullptr;
    int nt = 0, pending = 0;
    uint64_t generation = 0;
    bool stop = false;
    void loop(int id) {
        uint64_t seen = 0;
        for (;;) {
            void (*f)(int, void*);
            void* a;
            {
                std::unique_lock<std::mutex> lk(m);
                wake.wait(lk, [&] { return stop || generation != seen; });
                if (stop) return;
                seen = generation;
                if (id >= nt) continue; /* not part of this job */
                f = fn; a = arg;
            }
            f(id, a);
            std::lock_guard<std::mutex> lk(m);
            if (--pending == 0) idle.notify_one();
        }
    }
public:
    ~Pool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        wake.notify_all();
        for (auto& t : workers) t.join();
    }
    void run(int n, void (*f)(int, void*), void* a) {
        if (n <= 1) { f(0, a); return; }
        std::lock_guard<std::mutex> one(run_mutex);
        {
            std::lock_guard<std::mutex> lk(m);
            while ((int)workers.size() < n - 1) { const int id = (int)workers.size() + 1; workers.emplace_back([this, id] { loop(id); }); }
            fn = f; arg = a; nt = n; pending = n - 1;
            generation++;
        }
        wake.notify_all();
        f(0, a);
        std::unique_lock<std::mutex> lk(m);
        idle.wait(lk, [&] { return pending == 0; });
    }
};

struct SourceMap {
    const tpgx_config& cfg;
    const std::vector<tpgx_source_desc>& src;
    int first_env() const { return cfg.nb_constants > 0 ? 2 : 1; }
    int total() const { return first_env() + (int)src.size(); }
    /* address space of handler `s` for operand type `t`; 0 = the handler cannot provide it
     * ([UP] DataHandler::getAddressSpace; PrimitiveTypeArray / Array2DWrapper rules) */
    uint32_t space(int s, int t) const { return tpgx_source_space(cfg, src, s, t); }
};

std::string fmt(const char* f, long a = 0, long b = 0, long c = 0) {
    char buf[256];
    snprintf(buf, sizeof buf, f, a, b, c);
    return buf;
}

} // namespace

void tpgx_pool_run(int nt, void (*fn)(int, void*), void* arg) {
    static Pool pool;
    pool.run(std::min(nt, 16), fn, arg);
}

tpgx_status tpgx_flatten_graph(const tpgx_config& cfg, const std::vector<int32_t>& opcode,
                               const std::vector<tpgx_source_desc>& src, const tpgx_graph* g,
                               tpgx_flat_graph* out, std::string* err, std::vector<uint8_t>* keep_out, const std::vector<uint8_t>* prog_mask,
                               const int32_t* prog_ids) {
    auto fail = [&](tpgx_status st, const std::string& m) { if (err) *err = m; return st; };
    if (!g || !g->program_line_offset || !g->team_edge_offset || !g->edge_program || !g->edge_destination || !g->root_team)
        return fail(TPGX_ERR_BAD_ARG, "tpgx_set_graph: null array in graph");
    if (g->nb_programs < 0 || g->nb_teams < 1 || g->nb_roots < 1) return fail(TPGX_ERR_BAD_ARG, "tpgx_set_graph: empty graph");
    if (cfg.nb_constants > 0 && !g->constants) return fail(TPGX_ERR_BAD_ARG, "tpgx_set_graph: constants missing");
    if (src.size() > 2) return fail(TPGX_ERR_UNSUPPORTED, "device bytecode addresses at most 2 data sources");
    SourceMap sm{cfg, src};
    const int R = cfg.nb_registers;
    const bool trace = getenv("TPGX_TRACE_FLATTEN") != nullptr;
    auto t_last = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (!trace) return;
        const auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[tpgx] flatten: %s %.3f ms\n", what, std::chrono::duration<double, std::milli>(t - t_last).count());
        t_last = t;
    };

    tpgx_flat_graph& f = *out;
    f = tpgx_flat_graph();
    f.nb_programs = g->nb_programs;
    f.prog_flops.resize(g->nb_programs);
    /* constants are uploaded straight from the caller's array (tpgx_api.cu): no host copy */

    /* pass 1 (parallel over programs): validate, mark introns by backward liveness, count effective lines */
    const int P = g->nb_programs;
    /* prog_ids (a slice of a larger graph): program p of g has its lines and constants at program prog_ids[p] of the caller's
       arrays; the intron marks then live at offsets of their own (line_base), and nothing below is sized by the caller's graph */
    if (prog_ids && (keep_out || prog_mask)) return fail(TPGX_ERR_BAD_ARG, "tpgx_flatten_graph: prog_ids excludes keep_out / prog_mask");
    std::vector<int64_t> line_base;
    if (prog_ids) {
        line_base.resize((size_t)P + 1);
        line_base[0] = 0;
        for (int p = 0; p < P; p++) {
            const int64_t n = g->program_line_offset[prog_ids[p] + 1] - g->program_line_offset[prog_ids[p]];
            if (n < 0) return fail(TPGX_ERR_BAD_ARG, fmt("program %ld: negative length", prog_ids[p]));
            line_base[(size_t)p + 1] = line_base[(size_t)p] + n;
        }
    }
    const int64_t total_in = prog_ids ? line_base[(size_t)P] : P > 0 ? g->program_line_offset[P] : 0;
    if (total_in < 0) return fail(TPGX_ERR_BAD_ARG, "negative line count");
    /* 1 = effective; written for the programs that are flattened, never read for the others */
    std::unique_ptr<uint8_t[]> keep_buf(new uint8_t[(size_t)std::max<int64_t>(total_in, 1)]);
    uint8_t* const keep = keep_buf.get();
    std::vector<int32_t> n_eff((size_t)P, 0);
    struct Err { tpgx_status st = TPGX_OK; std::string msg; };
    std::vector<Err> errs(tpgx_parallel_workers());
    /* per-instruction operand signature and per-(source, type) address space, looked up per line below */
    struct OpRow { int32_t op, nb, type[2], flops; };
    std::vector<OpRow> optab(opcode.size());
    for (size_t i = 0; i < opcode.size(); i++) {
        const tpgx_op_info oi = tpgx_get_op_info(opcode[i]);
        optab[i] = {opcode[i], oi.nb_operands, {oi.type[0], oi.type[1]}, oi.flops};
    }
    const int nsrc = sm.total();
    std::vector<uint32_t> space_tab((size_t)nsrc * 8, 0u);
    std::vector<uint64_t> space_inv((size_t)nsrc * 8, 0u); /* loc % space without a division: a mod d = ((M * a mod 2^64) * d) >> 64, M = ceil(2^64 / d) */
    for (int sidx = 0; sidx < nsrc; sidx++)
        for (int t = 0; t <= TPGX_T_W; t++) {
            const uint32_t sp = sm.space(sidx, t);
            space_tab[(size_t)sidx * 8 + t] = sp;
            space_inv[(size_t)sidx * 8 + t] = sp ? UINT64_C(0xFFFFFFFFFFFFFFFF) / sp + 1 : 0;
        }
    auto reduce = [](uint32_t a, uint32_t sp, uint64_t inv) -> uint32_t {
        return a < sp ? a : (uint32_t)(((unsigned __int128)(inv * a) * sp) >> 64);
    };
    /* [UP] Array2DWrapper window address -> top-left element, per environment source (addresses are below TPGX_MAX_LOC) */
    std::vector<uint16_t> win_tab(src.size() * (size_t)TPGX_MAX_LOC, 0);
    for (size_t ks = 0; ks < src.size(); ks++) {
        const uint32_t ww = src[ks].width > 2 ? (uint32_t)src[ks].width - 2u : 1u;
        for (uint32_t a = 0; a < TPGX_MAX_LOC; a++) win_tab[ks * TPGX_MAX_LOC + a] = (uint16_t)std::min<uint32_t>(65535u, (a / ww) * (uint32_t)src[ks].width + (a % ww));
    }
    const uint32_t nI = (uint32_t)opcode.size();
    /* the programs to flatten: all, or the ones reachable from this context's root shard (the others stay empty and are never
       run).  The host threads split THIS list, so a shard of 1/N of the graph takes 1/N of the time on every thread */
    std::vector<int32_t> sel;
    if (prog_mask) {
        sel.reserve((size_t)P);
        for (int p = 0; p < P; p++) if ((*prog_mask)[(size_t)p]) sel.push_back(p);
    }
    const int nsel = prog_mask ? (int)sel.size() : P;
    lap("setup");
    tpgx_parallel_blocks(nsel, 128, [&](int pb, int pe, int w) {
        Err& er = errs[w];
        auto bad = [&](tpgx_status st, const std::string& m) { if (er.st == TPGX_OK) { er.st = st; er.msg = m; } };
        for (int q = pb; q < pe && er.st == TPGX_OK; q++) {
            const int p = prog_mask ? sel[(size_t)q] : q;
            const int gp = prog_ids ? prog_ids[p] : p; /* the caller's program id: addresses its arrays, named in messages */
            const int64_t b = g->program_line_offset[gp], e = g->program_line_offset[gp + 1];
            if (e < b || b < 0 || (!prog_ids && e > total_in)) { bad(TPGX_ERR_BAD_ARG, fmt("program %ld: negative length", gp)); break; }
            const tpgx_line* L = g->lines + b;
            uint8_t* kp = keep + (prog_ids ? line_base[(size_t)p] : b);
            const int n = (int)(e - b);
            /* [UP] Program::identifyIntrons: backward liveness from register 0, as a bit mask and without
               data-dependent branches (whether a line is an intron is unpredictable) */
            uint32_t live = 1u;
            int cnt = 0;
            for (int i = n - 1; i >= 0; i--) {
                const tpgx_line& l = L[i];
                if (l.instruction >= nI) { bad(TPGX_ERR_BAD_ARG, fmt("program %ld line %ld: instruction index %ld out of range", gp, i, l.instruction)); break; }
                if (l.destination >= (uint32_t)R) { bad(TPGX_ERR_BAD_ARG, fmt("program %ld line %ld: destination register out of range", gp, i)); break; }
                const OpRow& o = optab[l.instruction];
                uint32_t reads = 0;
                bool ok = true;
                for (int k = 0; k < o.nb; k++) {
                    const uint32_t sidx = l.src[k];
                    if (sidx >= (uint32_t)nsrc) { bad(TPGX_ERR_BAD_ARG, fmt("program %ld line %ld: operand source %ld out of range", gp, i, sidx)); ok = false; break; }
                    const uint32_t sp = space_tab[(size_t)sidx * 8 + o.type[k]];
                    if (sp == 0) {
                        bad(TPGX_ERR_GRAPH_INVALID, fmt("program %ld line %ld: source %ld cannot provide the operand type (the reference engine throws here)", gp, i, sidx));
                        ok = false;
                        break;
                    }
                    if (sidx == 0) reads |= 1u << reduce(l.loc[k], sp, space_inv[(size_t)sidx * 8 + o.type[k]]);
                }
                if (!ok) break;
                const uint32_t k1 = (live >> l.destination) & 1u;
                kp[i] = (uint8_t)k1;
                cnt += (int)k1;
                live = (live & ~(k1 << l.destination)) | (reads & (0u - k1));
            }
            n_eff[p] = cnt;
            if (g->intron && er.st == TPGX_OK && !prog_mask) {
                for (int i = 0; i < n; i++)
                    if ((g->intron[b + i] != 0) == (kp[i] != 0)) { bad(TPGX_ERR_BAD_ARG, fmt("program %ld line %ld: caller's intron flag disagrees with identifyIntrons", gp, i)); break; }
            }
        }
    });
    lap("pass 1 (introns)");
    for (const Err& er : errs) if (er.st != TPGX_OK) return fail(er.st, er.msg);
    if (keep_out) { /* host-only entry point (no mask): every program was marked */
        keep_out->assign(keep, keep + total_in);
    }
    f.prog_offset.assign((size_t)P + 1, 0);
    for (int p = 0; p < P; p++) f.prog_offset[p + 1] = f.prog_offset[p] + n_eff[p];
    f.code.assign((size_t)f.prog_offset[P], 0u);
    f.total_lines = f.prog_offset[P];
    f.total_introns = total_in - f.total_lines;

    lap("offsets");
    /* pass 2 (parallel): encode the effective lines; while a program's words are hot, the count pass of the K1 v2 stream
       encoder (entries, slots) and the opcodes present */
    f.v2_entries.assign((size_t)P, -1);
    f.v2_slots.assign((size_t)P, 0);
    std::vector<uint32_t> op_seen(tpgx_parallel_workers(), 0u);
    tpgx_parallel_blocks(nsel, 128, [&](int pb, int pe, int w) {
        Err& er = errs[w];
        uint32_t seen = 0;
        for (int q = pb; q < pe && er.st == TPGX_OK; q++) {
            const int p = prog_mask ? sel[(size_t)q] : q;
            const int gp = prog_ids ? prog_ids[p] : p;
            const int64_t b = g->program_line_offset[gp], e = g->program_line_offset[gp + 1];
            const tpgx_line* L = g->lines + b;
            const uint8_t* const kp = keep + (prog_ids ? line_base[(size_t)p] : b);
            const int n = (int)(e - b);
            uint32_t written = 0;
            int flops = 0;
            uint32_t* out_code = f.code.data() + f.prog_offset[p];
            for (int i = 0; i < n; i++) {
                if (!kp[i]) continue;
                const tpgx_line& l = L[i];
                const OpRow& o = optab[l.instruction];
                uint32_t kind[2] = {TPGX_KIND_REG, TPGX_KIND_REG}, loc[2] = {TPGX_LOC_ZERO, TPGX_LOC_ZERO};
                for (int k = 0; k < o.nb; k++) {
                    const int s = (int)l.src[k];
                    const uint32_t sp = space_tab[(size_t)s * 8 + o.type[k]];
                    uint32_t a = reduce(l.loc[k], sp, space_inv[(size_t)s * 8 + o.type[k]]);
                    if (s == 0) { kind[k] = TPGX_KIND_REG; loc[k] = ((written >> a) & 1u) ? a : TPGX_LOC_ZERO; }
                    else if (cfg.nb_constants > 0 && s == 1) { kind[k] = TPGX_KIND_CONST; loc[k] = a; }
                    else {
                        const int ks = s - sm.first_env();
                        kind[k] = ks == 0 ? TPGX_KIND_SRC0 : TPGX_KIND_SRC1;
                        if (o.type[k] == TPGX_T_W) a = a < TPGX_MAX_LOC ? win_tab[(size_t)ks * TPGX_MAX_LOC + a] : TPGX_MAX_LOC;
                        if (a >= TPGX_MAX_LOC) {
                            if (er.st == TPGX_OK) { er.st = TPGX_ERR_UNSUPPORTED; er.msg = fmt("program %ld line %ld: location %ld exceeds the 10-bit device field", gp, i, a); }
                            a = 0;
                        }
                        loc[k] = a;
                    }
                }
                *out_code++ = tpgx_pack_line((uint32_t)o.op, l.destination, kind[0], loc[0], kind[1], loc[1]);
                written |= 1u << l.destination;
                flops += o.flops;
                seen |= 1u << (o.op & 31);
            }
            f.prog_flops[p] = flops;
            const tpgx_k1v2_info info = tpgx_k1v2_count(f.code.data() + f.prog_offset[p], f.prog_offset[p + 1] - f.prog_offset[p]);
            if (info.ok) { f.v2_entries[p] = info.entries; f.v2_slots[p] = (uint8_t)info.slots; }
        }
        op_seen[w] |= seen;
    });
    for (uint32_t m : op_seen) f.op_mask |= m;
    lap("pass 2 (encode + K1 v2 counts)");
    for (const Err& er : errs) if (er.st != TPGX_OK) return fail(er.st, er.msg);

    f.nb_teams = g->nb_teams;
    f.team_edge_offset.assign(g->team_edge_offset, g->team_edge_offset + g->nb_teams + 1);
    if (f.team_edge_offset[0] != 0) return fail(TPGX_ERR_BAD_ARG, "team_edge_offset[0] != 0");
    for (int t = 0; t < g->nb_teams; t++) {
        if (f.team_edge_offset[t + 1] < f.team_edge_offset[t]) return fail(TPGX_ERR_BAD_ARG, "team_edge_offset not monotone");
        if (f.team_edge_offset[t + 1] == f.team_edge_offset[t]) return fail(TPGX_ERR_GRAPH_INVALID, fmt("team %ld has no outgoing edge", t));
    }
    f.nb_edges = f.team_edge_offset[g->nb_teams];
    f.edge_program.assign(g->edge_program, g->edge_program + f.nb_edges);
    f.edge_destination.assign(g->edge_destination, g->edge_destination + f.nb_edges);
    for (int e = 0; e < f.nb_edges; e++) {
        if (f.edge_program[e] < 0 || f.edge_program[e] >= g->nb_programs) return fail(TPGX_ERR_BAD_ARG, fmt("edge %ld: program out of range", e));
        if (f.edge_destination[e] >= g->nb_teams) return fail(TPGX_ERR_BAD_ARG, fmt("edge %ld: destination team out of range", e));
        if (f.edge_destination[e] < -128) return fail(TPGX_ERR_UNSUPPORTED, fmt("edge %ld: action id above 127", e));
    }
    f.nb_roots = g->nb_roots;
    f.root_team.assign(g->root_team, g->root_team + g->nb_roots);
    for (int r : f.root_team) if (r < 0 || r >= g->nb_teams) return fail(TPGX_ERR_BAD_ARG, "root team out of range");
    lap("graph arrays");
    return TPGX_OK;
}

extern "C" tpgx_status tpgx_flatten_programs(const tpgx_config* cfg, int32_t nb_instructions, const int32_t* opcode, int32_t nb_sources,
                                             const tpgx_source_desc* sources, const tpgx_graph* g, uint8_t* intron, int32_t* effective_lines,
                                             uint32_t* code, int64_t code_capacity, int64_t* nb_code_words, char* err, size_t err_len) {
    auto say = [&](const std::string& m) { if (err && err_len) snprintf(err, err_len, "%s", m.c_str()); };
    if (!cfg || !opcode || nb_instructions < 1 || !sources || nb_sources < 1 || !g) { say("tpgx_flatten_programs: null argument"); return TPGX_ERR_BAD_ARG; }
    if (cfg->nb_registers < 1 || cfg->nb_registers > TPGX_MAX_REGS || cfg->nb_constants < 0 || cfg->nb_constants > TPGX_MAX_CONSTS) {
        say("nb_registers must be in [1,8], nb_constants in [0,8]");
        return TPGX_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < nb_instructions; i++)
        if (opcode[i] < 0 || opcode[i] >= TPGX_OP_COUNT) { say("instruction without a device opcode"); return TPGX_ERR_UNSUPPORTED; }
    std::vector<int32_t> ops(opcode, opcode + nb_instructions);
    std::vector<tpgx_source_desc> src(sources, sources + nb_sources);
    tpgx_flat_graph f;
    std::vector<uint8_t> keep;
    std::string msg;
    const tpgx_status st = tpgx_flatten_graph(*cfg, ops, src, g, &f, &msg, &keep);
    if (st != TPGX_OK) { say(msg); return st; }
    if (intron) for (size_t i = 0; i < keep.size(); i++) intron[i] = keep[i] ? 0 : 1;
    if (effective_lines) for (int p = 0; p < f.nb_programs; p++) effective_lines[p] = f.prog_offset[p + 1] - f.prog_offset[p];
    if (nb_code_words) *nb_code_words = (int64_t)f.code.size();
    if (code) {
        if (code_capacity < (int64_t)f.code.size()) { say("code buffer too small"); return TPGX_ERR_BAD_ARG; }
        memcpy(code, f.code.data(), f.code.size() * 4);
    }
    return TPGX_OK;
}

/* K1 v2 encoder on its own (host only; the device runs the same function): entry stream, entry count and
 * shared-memory slots of one program given its packed generic words. */
extern "C" tpgx_status tpgx_k1v2_encode_program(const uint32_t* words, int32_t nb_words, int32_t width, int32_t hw, int32_t nwin,
                                                uint32_t* quads, int32_t quad_capacity, int32_t* nb_entries, int32_t* nb_slots) {
    if ((!words && nb_words > 0) || nb_words < 0 || width < 1 || hw < 1) return TPGX_ERR_BAD_ARG;
    const tpgx_k1v2_info count = tpgx_k1v2_encode(words, nb_words, 128u, (uint32_t)width, (uint32_t)hw, (uint32_t)nwin, nullptr);
    const tpgx_k1v2_info lean = tpgx_k1v2_count(words, nb_words); /* the planner's count pass must agree with the encoder */
    if (lean.ok != count.ok || (count.ok && (lean.entries != count.entries || lean.slots != count.slots))) return TPGX_ERR_STATE;
    if (nb_entries) *nb_entries = count.entries;
    if (nb_slots) *nb_slots = count.slots;
    if (!count.ok) return TPGX_ERR_UNSUPPORTED;
    if (quads) {
        if (quad_capacity < count.entries) return TPGX_ERR_BAD_ARG;
        const tpgx_k1v2_info w = tpgx_k1v2_encode(words, nb_words, 128u, (uint32_t)width, (uint32_t)hw, (uint32_t)nwin, reinterpret_cast<tpgx_k1_quad*>(quads));
        if (w.entries != count.entries || w.slots != count.slots) return TPGX_ERR_STATE;
    }
    return TPGX_OK;
}
extern "C" const char* tpgx_k1v2_handler_name(int32_t h) {
    static const char* const names[] = {
#define K1V2_NAME(n) #n,
        K1V2_HANDLERS(K1V2_NAME)
#undef K1V2_NAME
    };
    return (h >= 0 && h < K2H_COUNT) ? names[h] : nullptr;
}
