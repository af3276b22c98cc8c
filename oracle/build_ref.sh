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
# MNIST environment: its dataset is loaded at static-initialisation time from MNIST_DATA_LOCATION (IDX files, read by the
# reference's own mnist_reader).  No dataset ships with the reference, so a small synthetic one (3000 + 1000 MNIST-shaped
# images from tpgx.synthetic_images, fixed seeds) is generated next to the binary.
DATA="$OUT/mnist_data"
mkdir -p "$DATA"
python3 - "$DATA" "$HERE/.." <<'PY'
import os, struct, sys
sys.path.insert(0, os.path.join(sys.argv[2], "gegelati-apps_b200", "python"))
from tpgx.graph import synthetic_images
d = sys.argv[1]
for name, n, seed in (("train", 3000, 101), ("t10k", 1000, 202)):
    img, lab = synthetic_images(n, seed=seed, label_seed=seed + 1)
    open(os.path.join(d, "%s-images-idx3-ubyte" % name), "wb").write(struct.pack(">BBBBIII", 0, 0, 8, 3, n, 28, 28) + img.tobytes())
    open(os.path.join(d, "%s-labels-idx1-ubyte" % name), "wb").write(struct.pack(">BBBBI", 0, 0, 8, 1, n) + lab.tobytes())
PY
obj mnist_env      "$REF/mnist/src/mnist.cpp" -DMNIST_DATA_LOCATION="\"/root/repo/oracle/_ref/mnist_data\""
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
# The gridworld app's own main.cpp, UNMODIFIED ([REF] gridworld/src/main.cpp), against the shim + the C++ adapter: its
# Learn::ParallelLearningAgent is the GPU-backed agent (init / trainOneGeneration / getSelector()->keepBestPolicy / getBestRoot).
# ROOT_DIR "." : the test runs it in a scratch directory next to a copy of the app's params.json (kept under _ref/).
if [ -f "$LIBDIR/libtpgx.so" ]; then
  mkdir -p "$OUT/gridworld_app"
  cp "$REF/gridworld/params.json" "$OUT/gridworld_app/params.json"
  "$CXX" $FLAGS -I"$HERE/../include" -DTPGX_SHIM_WITH_AGENT -DNO_CONSOLE_CONTROL -DROOT_DIR='"."' \
      "$REF/gridworld/src/main.cpp" "$REF/gridworld/src/gridworld.cpp" "$REF/gridworld/src/instructions.cpp" -o "$OUT/gridworld_main" \
      -L"$LIBDIR" -ltpgx -Wl,-rpath,'$ORIGIN/../../gegelati-apps_b200/lib'
  echo "built $OUT/gridworld_main"
fi
rm -f "$OUT"/*.o
echo "built $OUT/libtpgref.so"
