#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference tree (run in the build container only; the GPU box
has no /root/reference).

1. Copies the reference's three golden graphs verbatim (data artefacts, not sources):
     pendulum/src/CodeGen/Pendulum_out_best.dot, stickgame/src/CodeGen/StickGame_out_best.dot,
     tic-tac-toe/src/CodeGen/TicTacToe_out_best.dot
2. If oracle/_ref/libtpgref.so exists (reference env + instruction sources compiled unmodified,
   see oracle/build_ref.sh) dumps golden vectors produced by the REAL reference code:
     ref_instructions.npz — every app lambda on a fixed operand grid
     ref_env_traces.npz   — seeded random-action traces of the four sequential environments
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"

for rel in ("pendulum/src/CodeGen/Pendulum_out_best.dot", "stickgame/src/CodeGen/StickGame_out_best.dot",
            "tic-tac-toe/src/CodeGen/TicTacToe_out_best.dot"):
    shutil.copy(os.path.join(REF, rel), HERE)

sys.path.insert(0, ROOT)
try:
    from oracle import refbridge
except Exception as e:  # noqa: BLE001
    print("oracle/_ref not available:", e)
else:
    refbridge.dump_golden(HERE)
