"""Build the checker's C twin: oracle/oracle_c.c -> oracle/_build/liboracle_c.so (gcc, pthreads).

TEST INFRASTRUCTURE ONLY -- building the checker is not using it.  The reference itself is pure Python
(+ the third-party MuJoCo engine), so there is nothing to compile into oracle/_ref (DESIGN.md "Oracle").

    python oracle/build.py [--force]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_c.so")
FLAGS = ["-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-std=gnu11"]   # no FMA contraction: plain IEEE double like NumPy


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    stamp = os.path.join(OUT_DIR, "liboracle_c.stamp")
    fp = hashlib.sha256(open(SRC, "rb").read() + " ".join(FLAGS).encode()).hexdigest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == fp:
        return LIB
    cmd = [os.environ.get("CC", "gcc"), *FLAGS, SRC, "-o", LIB, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout}")
    with open(stamp, "w") as f:
        f.write(fp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
