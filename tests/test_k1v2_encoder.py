"""K1 v2 stream encoder (csrc/k1v2_encode.h, host build) without a GPU: the entry stream of a program — accumulator
forwarding, live-range slot allocation, handler specialisation, block chaining — run by a small Python simulator
must give the bits the generic line words give, on random MNIST-set programs.  The arithmetic of exp / ln is not the
subject here (any deterministic stand-in works on both sides); the transformation is."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import tpgx
from tpgx import abi as A

TS4 = 32  # packed words per pixel row of a 128-sample tile
HW, WIDTH, NWIN = 784, 28, 676
CAP = 36  # K1V2_CAP: entries staged at a time


def f_exp(x):
    with np.errstate(all="ignore"):
        return np.exp(x)


def f_ln(x):
    with np.errstate(all="ignore"):
        return np.log(x)


def f_max(a, b):
    return np.where(a < b, b, a)


def f_div(a, b):
    with np.errstate(all="ignore"):
        return a / b


def run_generic(words, consts, px, magn, dirn):
    """The packed 32-bit words of csrc/tpgx_internal.h on one observation vector (px [785] doubles with px[784] = 0)."""
    reg = np.zeros(8)
    with np.errstate(all="ignore"):
        for w in words:
            w = int(w)
            op, dst = w & 31, (w >> 5) & 7
            f = [(w >> 8) & 0xfff, w >> 20]

            def val(k):
                kind, loc = f[k] >> 10, f[k] & 1023
                if kind == 0:
                    return 0.0 if loc >= 8 else reg[loc]
                if kind == 1:
                    return consts[loc]
                return px[loc]

            def win(k):
                loc = f[k] & 1023
                return (loc // WIDTH) * (WIDTH - 2) + loc % WIDTH

            if op == A.OP_SUB: r = val(0) - val(1)
            elif op == A.OP_ADD: r = val(0) + val(1)
            elif op == A.OP_MUL: r = val(0) * val(1)
            elif op == A.OP_DIV: r = f_div(np.float64(val(0)), np.float64(val(1)))
            elif op == A.OP_MAX: r = f_max(np.float64(val(0)), np.float64(val(1)))
            elif op == A.OP_EXP: r = f_exp(np.float64(val(0)))
            elif op == A.OP_LOG: r = f_ln(np.float64(val(0)))
            elif op == A.OP_MULC: r = np.float64(val(0)) * val(1) / 10.0
            elif op == A.OP_SOBEL_MAGN: r = magn[win(0)]
            elif op == A.OP_SOBEL_DIR: r = dirn[win(0)]
            else: raise AssertionError(op)
            reg[dst] = r
    return reg[0]


def run_v2(quads, slots, consts, px, magn, dirn, names):
    """The entry stream: accumulator + `slots` shared-memory slots; pixel operands are word indices (loc * 32),
    slot operands byte offsets (slot * 1024), window operands 32-byte group indices (plane * 676 * 32 + wp * 32)."""
    acc = None
    slot = [None] * 8
    planes = np.concatenate([magn, dirn])
    nb_st = 0
    with np.errstate(all="ignore"):
        for i, (h, y, z, w) in enumerate(quads):
            name = names[h]
            if name == "END":
                assert i == len(quads) - 1
                return acc, nb_st
            if name == "NEXT":
                assert i % CAP == CAP - 1, "NEXT must be the last entry of a staged block"
                assert y == i + 1, "NEXT names the row-relative index of the next block's first entry"
                continue
            if name == "ST":
                assert w % 1024 == 0 and w // 1024 < slots
                slot[w // 1024] = acc
                nb_st += 1
                continue

            def R(v):
                assert v % 1024 == 0 and slot[v // 1024] is not None, "slot read before written"
                return slot[v // 1024]

            def P(v):
                assert v % TS4 == 0 and v // TS4 <= HW
                return px[v // TS4]

            if name in ("SOBM", "SOBD"):
                assert y % TS4 == 0 and y // TS4 < 2 * NWIN and (y // TS4 >= NWIN) == (name == "SOBD"), "window outside the tile's two Sobel planes"
                acc = planes[y // TS4]
                continue
            if name in ("LUTE", "LUTL"):
                acc = (f_exp if name == "LUTE" else f_ln)(np.float64(P(y)))
                continue
            op, kinds = name.split("_")
            get = {"A": lambda v: acc, "R": R, "P": P}
            if op in ("EXP", "LN"):
                acc = (f_exp if op == "EXP" else f_ln)(np.float64(get[kinds[0]](y)))
                continue
            if op == "MULC":
                assert z % 8 == 0 and z // 8 < 8, "constant outside the 8 doubles at the head of the code buffer"
                acc = np.float64(get[kinds[0]](y)) * consts[z // 8] / 10.0
                continue
            a, b = np.float64(get[kinds[0]](y)), np.float64(get[kinds[1]](z))
            acc = {"ADD": lambda: a + b, "MUL": lambda: a * b, "SUB": lambda: a - b, "MAX": lambda: f_max(a, b),
                   "DIV": lambda: f_div(a, b)}[op]()
    raise AssertionError("stream without END")


def programs_of(g):
    out = tpgx.flatten_programs(tpgx.INSTRUCTION_SETS["mnist"], tpgx.APP_SOURCES["mnist"], g, nb_constants=5)
    off = np.concatenate([[0], np.cumsum(out["effective_lines"])])
    return [out["code"][off[p]:off[p + 1]] for p in range(g.nb_programs)]


def check_graph(g, rng, expect_all_encodable=True):
    names = tpgx.k1v2_handler_names()
    assert names[0] == "END" and "ADD_AR" in names and "DIV_PA" in names
    px = np.concatenate([rng.integers(0, 256, HW) * (rng.random(HW) < 0.5), [0]]).astype(np.float64)
    magn, dirn = rng.normal(size=NWIN) * 100, rng.normal(size=NWIN)
    max_slots, stores, lines = 0, 0, 0
    for p, words in enumerate(programs_of(g)):
        enc = tpgx.k1v2_encode_program(words)
        if enc is None:
            assert not expect_all_encodable
            continue
        consts = np.concatenate([np.asarray(g.constants[p], dtype=np.float64), np.zeros(8)])
        want = run_generic(words, consts, px, magn, dirn)
        got, nb_st = run_v2(enc["quads"], enc["slots"], consts, px, magn, dirn, names)
        assert np.float64(want).view(np.uint64) == np.float64(got).view(np.uint64) or (np.isnan(want) and np.isnan(got)), (p, want, got)
        max_slots = max(max_slots, enc["slots"])
        stores += nb_st
        lines += len(words)
    return max_slots, stores, lines


def test_synthetic_programs_all_effective():
    """The C2 generator's programs (every line effective): identical results, at most 8 slots, most results never stored."""
    g = tpgx.synthetic_graph("mnist", roots=60, edges=5, lines=20, depth=1, seed=42)
    max_slots, stores, lines = check_graph(g, np.random.default_rng(1))
    assert max_slots <= 8
    assert stores < 0.5 * lines, (stores, lines)


@pytest.mark.parametrize("lines", [1, 2, 23, 30, 35, 36, 37, 60, 72, 120])
def test_block_chaining_and_lengths(lines):
    g = tpgx.synthetic_graph("mnist", roots=6, edges=3, lines=lines, depth=1, seed=lines)
    check_graph(g, np.random.default_rng(lines))


@settings(max_examples=30, deadline=None)
@given(seed=st.integers(0, 2 ** 31))
def test_programs_with_introns_property(seed):
    g = tpgx.synthetic_graph("mnist", roots=8, edges=3, lines=12, depth=2, share_prob=0.1, seed=seed % 100000, intron_lines=8, line_jitter=6)
    check_graph(g, np.random.default_rng(seed))


def test_empty_program_and_unsupported_sets():
    names = tpgx.k1v2_handler_names()
    enc = tpgx.k1v2_encode_program(np.zeros(0, dtype=np.uint32))
    assert [names[h] for h in enc["quads"][:, 0]] == ["ADD_PP", "END"] and enc["slots"] == 0
    assert (enc["quads"][0, 1:3] == HW * TS4).all()          # both operands read the all-zero pixel row
    # an instruction outside the MNIST set (cos) is left to K1 v1
    word = A.OP_COS | (0 << 5) | ((2 << 10 | 5) << 8)
    assert tpgx.k1v2_encode_program(np.array([word], dtype=np.uint32)) is None


@pytest.mark.parametrize("lines", [1, 5, 20, 34, 35, 36, 70, 71, 141, 256])
def test_count_pass_equals_the_encoder_on_dense_random_words(lines):
    """The planner's count pass is closed-form (stores = values read later than the next line, slots = the largest number of
    them live at once, NEXT entries from the entry count); tpgx_k1v2_encode_program returns TPGX_ERR_STATE when it disagrees
    with the encoder's own replay of the slot allocator.  Random line words straight in the device format: every register
    written before it is read, registers reused heavily (long live ranges, up to 8 values in slots), lengths around the
    block capacity and up to the 256-line limit."""
    rng = np.random.default_rng(1000 + lines)
    ops = [A.OP_SUB, A.OP_ADD, A.OP_MUL, A.OP_DIV, A.OP_MAX, A.OP_EXP, A.OP_LOG, A.OP_MULC, A.OP_SOBEL_MAGN, A.OP_SOBEL_DIR]
    seen_slots, seen_next = 0, False
    for _ in range(60):
        written, words = 0, []
        for i in range(lines):
            op = int(rng.choice(ops))
            dst = 0 if i == lines - 1 else int(rng.integers(0, 8))
            f = []
            for k in range(2):
                if rng.random() < 0.7:
                    r = int(rng.integers(0, 8))
                    f.append((0 << 10) | (r if (written >> r) & 1 else 8))     # register, or "never written" (reads 0.0)
                else:
                    f.append((2 << 10) | int(rng.integers(0, 784)))            # pixel
            if op == A.OP_MULC:
                f[1] = (1 << 10) | int(rng.integers(0, 5))                      # program constant
            if op in (A.OP_SOBEL_MAGN, A.OP_SOBEL_DIR):
                f[0] = (2 << 10) | (int(rng.integers(0, 26)) * 28 + int(rng.integers(0, 26)))
            words.append(op | (dst << 5) | (f[0] << 8) | (f[1] << 20))
            written |= 1 << dst
        enc = tpgx.k1v2_encode_program(np.array(words, dtype=np.uint32))        # raises on TPGX_ERR_STATE
        assert enc is not None
        names = [tpgx.k1v2_handler_names()[h] for h in enc["quads"][:, 0]]
        assert names[-1] == "END" and names.count("END") == 1
        seen_slots = max(seen_slots, enc["slots"])
        seen_next |= "NEXT" in names
        assert names.count("NEXT") == (len(names) - names.count("NEXT") - 2) // (CAP - 1)
    if lines >= 20:
        assert seen_slots >= 4, seen_slots
    assert seen_next or lines < CAP, "a program of a block's length or more chains blocks"
    assert not (seen_next and lines <= 5)
