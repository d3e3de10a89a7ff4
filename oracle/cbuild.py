"""Compile the oracle's plain-C restatement (oracle/csrc/rulebook_ref.c) with gcc into oracle/_build/.
TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference's own implementation of this path lives in
the un-vendored spconv/cumm packages (not buildable here: needs pccm code generation + cumm), so there is
no oracle/_ref/ — DESIGN.md records that."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'rulebook_ref.c')
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'librulebook_ref.so')


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    r = subprocess.run(['gcc', '-O2', '-shared', '-fPIC', '-o', LIB, SRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('gcc failed:\n' + r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force=True))
