"""gcc build of the C oracle (oracle/sgd_oracle.c -> oracle/_build/libsgd_oracle.so).  TEST INFRASTRUCTURE."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libsgd_oracle.so")


def build_all(force=False):
    src = os.path.join(HERE, "sgd_oracle.c")
    os.makedirs(OUT, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", LIB, src, "-lm"])
    return LIB


if __name__ == "__main__":
    print(build_all(force=True))
