"""Host-side TPG containers in the reference's own encoding, the apps' instruction tables, a dot
importer for the reference's golden graphs (format: SURVEY.md Appendix C.1) and the seeded
synthetic generator G(R, E, L, D, mix) of SURVEY.md section 8d."""
from dataclasses import dataclass, field
import re
import numpy as np

from . import _cabi as A

LINE_DTYPE = np.dtype([("instruction", "<u4"), ("destination", "<u4"), ("src", "<u4", (2,)), ("loc", "<u4", (2,))])

# Instructions::Set of each app, in set.add() order
INSTRUCTION_SETS = {
    # [REF] mnist/src/main.cpp:79-88 (exp registered before ln)
    "mnist": [A.OP_SUB, A.OP_ADD, A.OP_MUL, A.OP_DIV, A.OP_MAX, A.OP_EXP, A.OP_LOG, A.OP_MULC, A.OP_SOBEL_MAGN, A.OP_SOBEL_DIR],
    # [REF] pendulum/src/Learn/instructions.cpp:20-32
    "pendulum": [A.OP_SUB, A.OP_ADD, A.OP_MUL, A.OP_DIV, A.OP_MAX, A.OP_EXP, A.OP_LOG, A.OP_COS, A.OP_SIN, A.OP_TAN, A.OP_MULC, A.OP_PI],
    # [REF] gridworld/src/instructions.cpp:7-27
    "gridworld": [A.OP_SUB, A.OP_ADD, A.OP_MUL, A.OP_DIV, A.OP_MAX, A.OP_EXP, A.OP_LOG, A.OP_COS, A.OP_SIN, A.OP_TAN],
    # [REF] stickgame/src/Learn/instructions.cpp:18-23
    "stickgame": [A.OP_FMODG, A.OP_SUB_II, A.OP_ADD, A.OP_CAST_I, A.OP_MAX, A.OP_EQ0],
    # [REF] tic-tac-toe/src/Learn/main.cpp:26-34
    "tictactoe": [A.OP_SUB, A.OP_ADD, A.OP_MAX, A.OP_FMODG, A.OP_EQM1, A.OP_EQ0, A.OP_EQ1, A.OP_GE15, A.OP_COND],
}
# (elem, height, width, is_2d) of getDataSources(), and params.json nbRegisters / nbProgramConstant
APP_SOURCES = {
    "mnist": [(A.ELEM_F64, 28, 28, 1)],
    "pendulum": [(A.ELEM_F64, 1, 2, 0)],
    "gridworld": [(A.ELEM_F64, 1, 2, 0)],
    "stickgame": [(A.ELEM_I32, 1, 4, 0), (A.ELEM_I32, 1, 1, 0)],   # hints first, remainingSticks second
    "tictactoe": [(A.ELEM_F64, 1, 9, 0)],
}
APP_CONSTANTS = {"mnist": 5, "pendulum": 5, "gridworld": 5, "stickgame": 0, "tictactoe": 0}
APP_ACTIONS = {"mnist": 10, "pendulum": 15, "gridworld": 4, "stickgame": 3, "tictactoe": 9}
APP_ENV = {"pendulum": A.ENV_PENDULUM, "gridworld": A.ENV_GRIDWORLD, "stickgame": A.ENV_STICKGAME, "tictactoe": A.ENV_TICTACTOE}


@dataclass
class TPGraph:
    """A TPG in reference iteration order (edge order within a team decides ties)."""
    lines: np.ndarray                 # LINE_DTYPE [nLines]
    program_line_offset: np.ndarray   # int64 [P+1]
    constants: np.ndarray             # float64 [P, nb_constants]
    team_edge_offset: np.ndarray      # int32 [T+1]
    edge_program: np.ndarray          # int32 [E]
    edge_destination: np.ndarray      # int32 [E]  >=0 team, <0 action = -1 - v
    root_team: np.ndarray             # int32 [R]
    names: dict = field(default_factory=dict)

    @property
    def nb_programs(self): return len(self.program_line_offset) - 1
    @property
    def nb_teams(self): return len(self.team_edge_offset) - 1
    @property
    def nb_roots(self): return len(self.root_team)
    @property
    def nb_edges(self): return len(self.edge_program)

    def subgraph(self, root_begin, root_end):
        """The part of the graph reachable from roots [root_begin, root_end), renumbered (teams in BFS order, programs in
        order of first use); its roots are exactly that slice.  What one rank of a multi-GPU job hands to tpgx_set_graph,
        so that the host flattener only touches the rank's own shard."""
        teo, ep, ed = self.team_edge_offset, self.edge_program, self.edge_destination
        roots = np.asarray(self.root_team[root_begin:root_end])
        order, seen = list(dict.fromkeys(int(t) for t in roots)), None
        seen = set(order)
        q = 0
        while q < len(order):
            t = order[q]; q += 1
            for d in ed[teo[t]:teo[t + 1]]:
                if d >= 0 and int(d) not in seen:
                    seen.add(int(d)); order.append(int(d))
        order = np.array(order, dtype=np.int64)
        tmap = np.full(self.nb_teams, -1, dtype=np.int32); tmap[order] = np.arange(len(order), dtype=np.int32)
        cnt = (teo[order + 1] - teo[order]).astype(np.int64)
        new_teo = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        eidx = np.concatenate([np.arange(teo[t], teo[t + 1]) for t in order]) if len(order) else np.zeros(0, dtype=np.int64)
        old_p = ep[eidx]
        uniq, first = np.unique(old_p, return_index=True)
        uniq = uniq[np.argsort(first)]                      # order of first use
        pmap = np.full(self.nb_programs, -1, dtype=np.int32); pmap[uniq] = np.arange(len(uniq), dtype=np.int32)
        plo = self.program_line_offset
        lens = (plo[uniq + 1] - plo[uniq]).astype(np.int64)
        new_plo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        lidx = np.concatenate([np.arange(plo[p], plo[p + 1]) for p in uniq]) if len(uniq) else np.zeros(0, dtype=np.int64)
        d_old = ed[eidx]
        new_ed = np.where(d_old >= 0, tmap[np.maximum(d_old, 0)], d_old).astype(np.int32)
        return TPGraph(lines=self.lines[lidx], program_line_offset=new_plo, constants=self.constants[uniq], team_edge_offset=new_teo,
                       edge_program=pmap[old_p], edge_destination=new_ed, root_team=tmap[roots].astype(np.int32))

    def c_struct(self):
        """Returns (Graph struct, keepalive list)."""
        arrs = dict(
            plo=np.ascontiguousarray(self.program_line_offset, dtype=np.int64),
            lines=np.ascontiguousarray(self.lines, dtype=LINE_DTYPE),
            consts=np.ascontiguousarray(self.constants, dtype=np.float64),
            teo=np.ascontiguousarray(self.team_edge_offset, dtype=np.int32),
            ep=np.ascontiguousarray(self.edge_program, dtype=np.int32),
            ed=np.ascontiguousarray(self.edge_destination, dtype=np.int32),
            rt=np.ascontiguousarray(self.root_team, dtype=np.int32),
        )
        g = A.Graph()
        g.nb_programs = self.nb_programs
        g.program_line_offset = arrs["plo"].ctypes.data
        g.lines = arrs["lines"].ctypes.data if len(arrs["lines"]) else None
        g.intron = None
        g.constants = arrs["consts"].ctypes.data if arrs["consts"].size else None
        g.nb_teams = self.nb_teams
        g.team_edge_offset = arrs["teo"].ctypes.data
        g.edge_program = arrs["ep"].ctypes.data
        g.edge_destination = arrs["ed"].ctypes.data
        g.nb_roots = self.nb_roots
        g.root_team = arrs["rt"].ctypes.data
        return g, arrs


def address_space(sources, nb_registers, nb_constants, src_idx, typ):
    """[UP] DataHandler::getAddressSpace for handler index src_idx of [registers, constants?, sources...]."""
    if src_idx == 0:
        return nb_registers if typ == "D" else 0
    if nb_constants > 0 and src_idx == 1:
        return nb_constants if typ == "C" else 0
    k = src_idx - (2 if nb_constants > 0 else 1)
    if k < 0 or k >= len(sources):
        return 0
    elem, h, w, is2d = sources[k]
    if elem == A.ELEM_I32:
        return w if (typ == "I" and not is2d) else 0
    if typ == "D":
        return h * w
    if typ == "W" and is2d and h >= 3 and w >= 3:
        return (h - 2) * (w - 2)
    return 0


def import_dot(path_or_text, nb_constants):
    """Parses a TPGGraphDotExporter file (GEGELATI v2.0.0 layout).
    Teams, programs and edges keep file order. Returns a TPGraph; names map ids back to T/P/A labels."""
    text = path_or_text if "\n" in str(path_or_text) else open(path_or_text).read()
    team_ids, team_edges = [], {}
    prog_lines, prog_consts, prog_dest, prog_order = {}, {}, {}, []
    action_label = {}
    roots = []
    for raw in text.splitlines():
        ln = raw.strip()
        m = re.match(r"^T(\d+) \[fillcolor", ln)
        if m:
            t = int(m.group(1)); team_ids.append(t); team_edges[t] = []; continue
        m = re.match(r"^P(\d+) \[.*\] //(.*)$", ln)
        if m:
            p = int(m.group(1)); prog_order.append(p)
            prog_consts[p] = [float(v) for v in m.group(2).split("|") if v.strip() != ""]
            continue
        m = re.match(r'^I(\d+) \[.*label="(.*)"\]', ln)
        if m:
            p = int(m.group(1)); body = m.group(2); lines = []
            for item in body.split("&#92;n"):
                if not item:
                    continue
                mm = re.match(r"^(\d+)\|(\d+)&(\d+)\|(\d+)#(\d+)\|(\d+)$", item)
                if not mm:
                    raise ValueError("bad line %r" % item)
                lines.append(tuple(int(v) for v in mm.groups()))
            prog_lines[p] = lines; continue
        m = re.match(r'^A(\d+) \[.*label="(\d+)"\]', ln)
        if m:
            action_label[int(m.group(1))] = int(m.group(2)); continue
        m = re.match(r"^T(\d+) -> P(\d+) -> ([TA])(\d+)$", ln)
        if m:
            t, p, kind, d = int(m.group(1)), int(m.group(2)), m.group(3), int(m.group(4))
            prog_dest[p] = (kind, d); team_edges[t].append(p); continue
        m = re.match(r"^T(\d+) -> P(\d+)$", ln)
        if m:
            team_edges[int(m.group(1))].append(int(m.group(2))); continue
        m = re.match(r"^\{ rank= same (.*)\}$", ln)
        if m:
            roots = [int(v[1:]) for v in m.group(1).split()]
    pid = {p: i for i, p in enumerate(prog_order)}
    tid = {t: i for i, t in enumerate(team_ids)}
    all_lines, plo = [], [0]
    for p in prog_order:
        for (ins, dst, s0, l0, s1, l1) in prog_lines.get(p, []):
            all_lines.append((ins, dst, (s0, s1), (l0, l1)))
        plo.append(len(all_lines))
    consts = np.zeros((len(prog_order), nb_constants))
    for p in prog_order:
        c = prog_consts.get(p, [])
        if nb_constants and len(c) != nb_constants:
            raise ValueError("program P%d has %d constants, expected %d" % (p, len(c), nb_constants))
        if nb_constants:
            consts[pid[p]] = c
    teo, ep, ed = [0], [], []
    for t in team_ids:
        for p in team_edges[t]:
            kind, d = prog_dest[p]
            ep.append(pid[p])
            ed.append(tid[d] if kind == "T" else -1 - action_label[d])
        teo.append(len(ep))
    return TPGraph(lines=np.array(all_lines, dtype=LINE_DTYPE), program_line_offset=np.array(plo, dtype=np.int64),
                   constants=consts, team_edge_offset=np.array(teo, dtype=np.int32), edge_program=np.array(ep, dtype=np.int32),
                   edge_destination=np.array(ed, dtype=np.int32), root_team=np.array([tid[r] for r in roots], dtype=np.int32),
                   names={"teams": team_ids, "programs": prog_order})


def export_dot(g, path=None, version="2.0.0"):
    """Writes a TPGraph in the TPGGraphDotExporter layout of GEGELATI v2.0.0 (the format of
    [REF] pendulum/src/CodeGen/Pendulum_out_best.dot), the reference's checkpoint format
    ([REF] mnist/src/main.cpp:140-143).  Lines are written as given (un-reduced locations, introns included);
    constants with 6 decimals like the reference.  A program referenced by edges with different destinations is
    written once per destination (the format ties a program to one destination).  Returns the text."""
    out = ["// File exported with GEGELATI v%s" % version, "// With the class File::TPGGraphDotExporter (tpgx.export_dot)", "digraph{",
           '\tgraph[pad = "0.212, 0.055" bgcolor = lightgray]', '\tnode[shape=circle style = filled label = ""]']
    nc = g.constants.shape[1] if g.constants.ndim == 2 else 0
    written = {}          # (program, destination) -> P id
    next_p = 0
    declared_actions = set()
    for t in range(g.nb_teams):
        out.append('\t\tT%d [fillcolor="#1199bb"]' % t)
    for t in range(g.nb_teams):
        for e in range(g.team_edge_offset[t], g.team_edge_offset[t + 1]):
            p, d = int(g.edge_program[e]), int(g.edge_destination[e])
            if (p, d) in written:
                out.append("\t\tT%d -> P%d" % (t, written[(p, d)]))
                continue
            pid = next_p; next_p += 1
            written[(p, d)] = pid
            consts = "".join("%.6f|" % c for c in g.constants[p]) if nc else ""
            out.append('\t\tP%d [fillcolor="#cccccc" shape=point label="0"] //%s' % (pid, consts))
            body = "".join("%d|%d&%d|%d#%d|%d&#92;n" % (l["instruction"], l["destination"], l["src"][0], l["loc"][0], l["src"][1], l["loc"][1])
                           for l in g.lines[g.program_line_offset[p]:g.program_line_offset[p + 1]])
            out.append('\t\tI%d [shape=box style=invis label="%s"] ' % (pid, body))
            out.append("\t\tP%d -> I%d[style=invis]" % (pid, pid))
            if d < 0:
                a = -1 - d
                if a not in declared_actions:
                    declared_actions.add(a)
                out.append('\t\tA%d [fillcolor="#ff3366" shape=box margin=0.03 width=0 height=0 label="%d"]' % (1000 + a, a))
                out.append("\t\tT%d -> P%d -> A%d" % (t, pid, 1000 + a))
            else:
                out.append("\t\tT%d -> P%d -> T%d" % (t, pid, d))
    out.append("\t\t{ rank= same %s}" % "".join("T%d " % int(r) for r in g.root_team))
    out.append("}")
    text = "\n".join(out) + "\n"
    if path is not None:
        with open(path, "w") as f:
            f.write(text)
    return text


def random_programs(rng, nb_programs, nb_lines, opcodes, sources, nb_registers, nb_constants, allowed=None,
                    intron_lines=0, const_range=(-10, 10), integer_constants=True, balance=True):
    """Programs whose EFFECTIVE lines follow a uniform instruction mix over `allowed`.

    A line that reads no register (a 3x3 window instruction) cannot be placed where it would leave no live
    register for the earlier lines, so drawing instructions uniformly under-represents those instructions
    (3.7 % instead of 10 % for the MNIST set).  With balance=True the draw weights are calibrated by a few
    deterministic fixed-point passes so that the realised mix is uniform (4.7 flop per
    line for the MNIST set)."""
    allowed_arr = np.arange(len(opcodes)) if allowed is None else np.asarray(allowed)
    base_seed = int(rng.integers(0, 2 ** 62))
    weights = np.ones(len(allowed_arr)) / len(allowed_arr)
    passes = 8 if (balance and nb_programs * nb_lines >= 2000 and len(allowed_arr) > 1) else 1
    out = None
    for t in range(passes):
        out = _random_programs(np.random.default_rng([base_seed, 0 if passes == 1 else t]), nb_programs, nb_lines, opcodes, sources,
                               nb_registers, nb_constants, allowed_arr, intron_lines, const_range, integer_constants, weights)
        if t + 1 < passes:
            ins = out[0]["instruction"].reshape(nb_programs, -1)[:, :nb_lines].ravel()
            got = np.array([(ins == a).mean() for a in allowed_arr])
            weights = weights * (1.0 / len(allowed_arr)) / np.maximum(got, 1e-4)
            weights /= weights.sum()
    return out


def _random_programs(rng, nb_programs, nb_lines, opcodes, sources, nb_registers, nb_constants, allowed,
                     intron_lines, const_range, integer_constants, weights):
    """nb_programs programs with exactly nb_lines EFFECTIVE lines each (built backward from r0 so that
    every destination is read later), plus `intron_lines` random extra lines spliced in that the
    intron pass must remove.  Instruction indices uniform over `allowed` (default: the whole set);
    operand sources uniform over the handlers that can provide the operand type, locations uniform
    over [0, largest address space) and left un-reduced, like [UP] Mutator::LineMutator."""
    P, L = nb_programs, nb_lines
    n_ops = len(opcodes)
    sigs = [A.OP_SIGNATURE[o] for o in opcodes]
    first_env = 2 if nb_constants > 0 else 1
    n_src = first_env + len(sources)
    spaces = {t: [address_space(sources, nb_registers, nb_constants, s, t) for s in range(n_src)] for t in "DICW"}
    largest = max(max(v) for v in spaces.values())
    admissible = {t: np.array([s for s in range(n_src) if spaces[t][s] > 0]) for t in "DICW"}
    has_reg_operand = np.array([("D" in s) for s in sigs])
    allowed_reg = allowed[has_reg_operand[allowed]]
    lines = np.zeros((P, L), dtype=LINE_DTYPE)
    live = np.zeros((P, nb_registers), dtype=bool)
    live[:, 0] = True
    for i in range(L - 1, -1, -1):
        score = rng.random((P, nb_registers)) * live
        dst = score.argmax(axis=1)
        live[np.arange(P), dst] = False
        ins = allowed[rng.choice(len(allowed), P, p=weights)]
        need_reg = (~live.any(axis=1)) & (i > 0)
        if need_reg.any():
            if len(allowed_reg) == 0:
                raise ValueError("instruction mix cannot chain registers")
            bad = need_reg & ~has_reg_operand[ins]
            ins[bad] = allowed_reg[rng.integers(0, len(allowed_reg), int(bad.sum()))]
        src = np.zeros((P, 2), dtype=np.uint32)
        loc = rng.integers(0, largest, (P, 2)).astype(np.uint32)
        src[:] = rng.integers(0, n_src, (P, 2))  # unused operand slots: any source (ignored by the engine)
        for o in np.unique(ins):
            sel = np.nonzero(ins == o)[0]
            for k, t in enumerate(sigs[o]):
                adm = admissible[t]
                if len(adm) == 0:
                    raise ValueError("no source can provide operand type %s" % t)
                src[sel, k] = adm[rng.integers(0, len(adm), len(sel))]
        forced = need_reg  # force operand 0 (a double operand of these instructions) to be a register
        for o in np.unique(ins[forced]) if forced.any() else []:
            k = sigs[o].index("D")
            sel = np.nonzero(forced & (ins == o))[0]
            src[sel, k] = 0
        for o in np.unique(ins):
            sel = np.nonzero(ins == o)[0]
            for k, t in enumerate(sigs[o]):
                if t == "D":
                    isreg = src[sel, k] == 0
                    regs = loc[sel, k] % nb_registers
                    live[sel[isreg], regs[isreg]] = True
        lines["instruction"][:, i] = ins
        lines["destination"][:, i] = dst
        lines["src"][:, i] = src
        lines["loc"][:, i] = loc
    if intron_lines > 0:
        # append lines after the last effective one whose destination is not r0 (dead stores), and
        # prepend lines that are overwritten before being read is harder to guarantee; dead tail suffices
        extra = np.zeros((P, intron_lines), dtype=LINE_DTYPE)
        ins = allowed[rng.integers(0, len(allowed), (P, intron_lines))]
        extra["instruction"] = ins
        extra["destination"] = rng.integers(1, nb_registers, (P, intron_lines))
        esrc = np.zeros((P, intron_lines, 2), dtype=np.uint32)
        for o in np.unique(ins):
            pi, li = np.nonzero(ins == o)
            for k in range(2):
                t = sigs[o][k] if k < len(sigs[o]) else "D"
                adm = admissible[t]
                esrc[pi, li, k] = adm[rng.integers(0, len(adm), len(pi))]
        extra["src"] = esrc
        extra["loc"] = rng.integers(0, largest, (P, intron_lines, 2))
        lines = np.concatenate([lines, extra], axis=1)
    Ltot = lines.shape[1]
    plo = np.arange(P + 1, dtype=np.int64) * Ltot
    if nb_constants > 0:
        lo, hi = const_range
        consts = rng.integers(lo, hi + 1, (P, nb_constants)).astype(np.float64) if integer_constants else rng.uniform(lo, hi, (P, nb_constants))
    else:
        consts = np.zeros((P, 0))
    return lines.reshape(-1), plo, consts


def synthetic_graph(app="mnist", roots=500, edges=5, lines=20, depth=1, mix="uniform", share_prob=0.0, seed=42,
                    intron_lines=0, nb_registers=8, line_jitter=0):
    """G(R, E, L, D, mix, seed) of SURVEY.md section 8d for the given app's instruction set and sources."""
    rng = np.random.default_rng(seed)
    opcodes = INSTRUCTION_SETS[app]
    sources = APP_SOURCES[app]
    nc = APP_CONSTANTS[app]
    nA = APP_ACTIONS[app]
    allowed = None
    if mix == "scalar":
        allowed = [i for i, o in enumerate(opcodes) if o not in (A.OP_SOBEL_MAGN, A.OP_SOBEL_DIR)]
    elif mix.startswith("ops:"):  # explicit instruction indices, e.g. "ops:0,1,2"
        allowed = [int(v) for v in mix[4:].split(",")]
    elif mix == "arith":  # only exactly-rounded IEEE operations (no libm)
        allowed = [i for i, o in enumerate(opcodes) if o not in (A.OP_EXP, A.OP_LOG, A.OP_SOBEL_DIR, A.OP_COS, A.OP_SIN, A.OP_TAN)]
    widths = [roots] + [max(1, roots // 2)] * (depth - 1)
    level_start = np.concatenate([[0], np.cumsum(widths)])
    T = int(level_start[-1])
    teo = np.arange(T + 1, dtype=np.int32) * edges
    E = T * edges
    ed = np.zeros(E, dtype=np.int32)
    for lvl in range(depth):
        t0, t1 = int(level_start[lvl]), int(level_start[lvl + 1])
        n = (t1 - t0) * edges
        act = -1 - rng.integers(0, nA, n).astype(np.int32)
        if lvl + 1 < depth:
            nxt = rng.integers(int(level_start[lvl + 1]), int(level_start[lvl + 2]), n).astype(np.int32)
            to_team = rng.random(n) < 0.5
            to_team.reshape(-1, edges)[:, 0] = False  # every team keeps >= 1 action edge
            ed[t0 * edges:t1 * edges] = np.where(to_team, nxt, act)
        else:
            ed[t0 * edges:t1 * edges] = act
    ep = np.arange(E, dtype=np.int32)
    if share_prob > 0:
        shared = rng.random(E) < share_prob
        shared[0] = False
        donors = (rng.random(E) * np.arange(E)).astype(np.int32)  # an earlier edge
        ep = np.where(shared, ep[donors], ep)
        uniq, inv = np.unique(ep, return_inverse=True)
        ep = inv.astype(np.int32)
        P = len(uniq)
    else:
        P = E
    ln, plo, consts = random_programs(rng, P, lines, opcodes, sources, nb_registers, nc, allowed, intron_lines)
    if line_jitter > 0:  # ragged programs: truncate each program's head by a random amount (keeps r0 = last line)
        Ltot = lines + intron_lines
        cut = rng.integers(0, line_jitter + 1, P)
        keep = np.ones((P, Ltot), dtype=bool)
        keep[np.arange(Ltot)[None, :] < cut[:, None]] = False
        ln = ln.reshape(P, Ltot)[keep]
        plo = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int64)
    return TPGraph(lines=ln, program_line_offset=plo, constants=consts, team_edge_offset=teo, edge_program=ep,
                   edge_destination=ed, root_team=np.arange(roots, dtype=np.int32))


def synthetic_images(n=60000, seed=1234, label_seed=1235, h=28, w=28, zero_fraction=0.5):
    """MNIST-shaped synthetic observations: uint8 uniform 0..255 with a 50 % zero mask
    (keeps log(0), x/0 and NaN paths exercised); labels uniform 0..9."""
    rng = np.random.Generator(np.random.PCG64(seed))
    img = rng.integers(0, 256, (n, h * w), dtype=np.uint8)
    img[rng.random((n, h * w)) < zero_fraction] = 0
    lab = np.random.Generator(np.random.PCG64(label_seed)).integers(0, 10, n, dtype=np.uint8)
    return img, lab
