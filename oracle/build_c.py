"""Build the oracle's C helper (test infrastructure): oracle/_ransac_f32.so from oracle/ransac_f32.c."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ransac_f32.c")
LIB = os.path.join(HERE, "_ransac_f32.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(True))
