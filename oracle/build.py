"""Build the oracle's C restatement (oracle/_build/libmsda_oracle.so) with plain gcc.

ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/msda_oracle.c header).  Called by
__graft_entry__.build() and lazily by oracle.c_oracle; building the checker is not using it.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "msda_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libmsda_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-std=c11", "-shared", "-fPIC", SRC, "-o", OUT, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
