"""TEST INFRASTRUCTURE ONLY.  g++ build of tests/hostcore/postproc_host.cpp (the host compilation of the post-processing
kernels' per-pixel arithmetic, dvis_plus_b200/csrc/resize_core.cuh) into tests/hostcore/_build/libpostproc_host.so."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "postproc_host.cpp")
CORE = os.path.join(ROOT, "dvis_plus_b200", "csrc", "resize_core.cuh")
OUT = os.path.join(HERE, "_build", "libpostproc_host.so")


def build(force=False):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        return OUT
    # -x c++: the core header keeps its .cuh name; -ffp-contract=off: no fused multiply-add the device build would not
    # also be free to choose differently -- parity at decision boundaries is tolerance-based either way
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", f"-I{os.path.dirname(CORE)}", SRC, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
