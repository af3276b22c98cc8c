#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.  Builds oracle/_ref/libtpgref.so from the reference's
# own, unmodified sources where they lie under /root/reference (never copied into this repo):
# the four sequential LearningEnvironments and the five instruction sets.  The GEGELATI library
# those sources include is not vendored, so they are compiled against oracle/shim/gegelati.h.
# The reference's own build system (cmake + find_package(GEGELATI)) is not run.
# Flags follow the apps' CMakeLists ([REF] mnist/CMakeLists.txt:21-40: C++17, extensions off,
# Release, no -march=native, no fast-math, no -mfma): plain SSE2 IEEE-754 double arithmetic.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${TPGX_REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
GEN="$OUT/gen"
CXX="${CXX:-g++}"
[ -d "$REF" ] || { echo "build_ref.sh: $REF not present (GPU box: the prebuilt oracle/_ref is used)"; exit 0; }
mkdir -p "$GEN"

# Slice the instruction-set blocks that live inside main(): everything after the line declaring
# `Instructions::Set set;` up to the last `set.add(` line.  Generated files stay under _ref/.
slice() { # $1 = source, $2 = output
  awk '
    /Instructions::Set[ \t]+set;/ && !start { start = NR + 1 }
    /set\.add\(/ { last = NR }
    { line[NR] = $0 }
    END { if (!start || !last) exit 1; for (i = start; i <= last; i++) print line[i] }
  ' "$1" > "$2"
}
slice "$REF/mnist/src/main.cpp" "$GEN/mnist_set.inc"
slice "$REF/tic-tac-toe/src/Learn/main.cpp" "$GEN/tictactoe_set.inc"

FLAGS="-std=c++17 -O2 -fPIC -w -I$HERE/shim -I$REF -I$OUT"
obj() { # $1 = tag, $2 = source, rest = extra flags
  local tag="$1" src="$2"; shift 2
  "$CXX" $FLAGS "$@" -c "$src" -o "$OUT/$tag.o"
}
obj pendulum_env   "$REF/pendulum/src/Learn/pendulum.cpp"
obj pendulum_ins   "$REF/pendulum/src/Learn/instructions.cpp" -DfillInstructionSet=pendulum_fillInstructionSet
obj gridworld_env  "$REF/gridworld/src/gridworld.cpp"
obj gridworld_ins  "$REF/gridworld/src/instructions.cpp" -DfillInstructionSet=gridworld_fillInstructionSet
obj stickgame_env  "$REF/stickgame/src/Learn/stickGameAdversarial.cpp"
obj stickgame_ins  "$REF/stickgame/src/Learn/instructions.cpp" -DfillInstructionSet=stickgame_fillInstructionSet
obj tictactoe_env  "$REF/tic-tac-toe/src/Learn/TicTacToe.cpp"
obj wrap           "$HERE/ref_wrap.cpp"
"$CXX" -shared -o "$OUT/libtpgref.so" "$OUT"/*.o
rm -f "$OUT/wrap.o"
# The C++ adapter demo: the same reference objects + tests/cpp/adapter_driver.cpp + include/tpgx_gegelati.hpp, linked
# against the product library (needs it built; skipped otherwise).
LIBDIR="$HERE/../gegelati-apps_b200/lib"
if [ -f "$LIBDIR/libtpgx.so" ]; then
  "$CXX" $FLAGS -I"$HERE/../include" "$HERE/../tests/cpp/adapter_driver.cpp" "$OUT"/*.o -o "$OUT/adapter_driver" \
      -L"$LIBDIR" -ltpgx -Wl,-rpath,'$ORIGIN/../../gegelati-apps_b200/lib'
  echo "built $OUT/adapter_driver"
fi
rm -f "$OUT"/*.o
echo "built $OUT/libtpgref.so"
