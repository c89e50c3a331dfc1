"""Regenerates profiles/r2_sass_kernels.txt: the tensor-core / TMA / mixed-precision SASS mnemonics of every kernel in the built
library (cuobjdump -sass dvis_plus_b200/lib/libdvis_b200.so).  Runs on the build machine (no GPU needed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "dvis_plus_b200", "lib", "libdvis_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|SYNCS|FHFMA|FFMA2|HMMA|LDSM|LDGSTS)[A-Za-z0-9_.]*")
KERNELS = re.compile(r"flash_attn_kernel|mask_gemm_kernel|linear_tc_kernel|small_linear_kernel|msda_fwd_staged_kernel<__nv_bfloat16, __nv_bfloat16, (\(int\))?32")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = ["SASS mnemonics per kernel of dvis_plus_b200/lib/libdvis_b200.so (cuobjdump -sass, sm_100a), round 2; tests/perf/sass_extract.py.",
           "tcgen05 / TMEM / TMA: UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM (tcgen05.ld / tcgen05.st), UTMALDG / UTMASTG "
           "(cp.async.bulk.tensor load / store), SYNCS (mbarrier)",
           "Blackwell mixed-precision FMA: FHFMA.BF16 (fma.rn.f32.bf16);  packed fp32: FFMA2;  legacy warp MMA path: HMMA + LDSM + "
           "LDGSTS (cp.async)", ""]
    name, counts = None, None
    funcs = []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name, counts = m.group(1), collections.Counter()
            funcs.append((name, counts))
            continue
        if counts is not None:
            mm = PAT.search(line)
            if mm:
                counts[mm.group(0).rstrip(".")] += 1
    dem = subprocess.run(["cu++filt"] + [f for f, _ in funcs], capture_output=True, text=True).stdout.splitlines()
    for (mangled, c), d in zip(funcs, dem):
        d = d.replace("dvis::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
        if not KERNELS.search(d) or not c:
            continue
        out.append(d)
        out.append("    " + ", ".join(f"{k} x{v}" for k, v in c.most_common()))
    open(os.path.join(ROOT, "profiles", "r2_sass_kernels.txt"), "w").write("\n".join(out) + "\n")
    print(len(out) // 2, "kernels")


if __name__ == "__main__":
    sys.exit(main())
