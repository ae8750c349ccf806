"""Per-kernel SASS evidence for libb200krylov.so (run here, no GPU needed): counts of the mnemonics that prove
which hardware paths a kernel uses — UBLKCP (TMA 1-D bulk copy, cp.async.bulk), SYNCS (mbarrier), DMMA (FP64
tensor core, mma.sync m8n8k4), LDG/STG width mix, system-scope accesses (LD/ST .STRONG.SYS = volatile / acquire / release at
sys scope, MEMBAR.*.SYS = __threadfence_system: the NVLink peer-window protocol), ATOM/RED, BAR.  Writes profiles/<tag>_sass_summary.txt.

    python tools/sass_summary.py [tag]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "krylovkit.jl_b200", "libb200krylov.so")
PAT = collections.OrderedDict([
    ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("DMMA", r"\bDMMA"),
    ("DFMA", r"\bDFMA"), ("FFMA", r"\bFFMA"), ("LDG.128", r"\bLDG\.E\.128"), ("LDG", r"\bLDG"), ("STG.128", r"\bSTG\.E\.128"),
    ("STG", r"\bSTG"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("STRONG.SYS", r"\.STRONG\.SYS\b"), ("MEMBAR.SYS", r"MEMBAR\.\w+\.SYS|FENCE\.\w*\.?SYS"), ("ATOM/RED", r"\b(ATOMG|ATOM|RED|REDG)\b"),
    ("BAR", r"\bBAR\."), ("SHFL", r"\bSHFL"),
])


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur and "/*" in line:
            for name, pat in PAT.items():
                if re.search(pat, line):
                    kernels[cur][name] += 1
            kernels[cur]["instructions"] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.txt")
    with open(path, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): per-kernel mnemonic counts\n")
        f.write("# UBLKCP = cp.async.bulk (TMA 1-D bulk copy); SYNCS = mbarrier ops; DMMA = FP64 tensor core;\n")
        f.write("# STRONG.SYS = volatile / ld.acquire.sys / st.release.sys accesses; MEMBAR.SYS = fence.sys (NVLink peer window)\n")
        cols = list(PAT) + ["instructions"]
        f.write("kernel\t" + "\t".join(cols) + "\n")
        tot = collections.Counter()
        for (mangled, cnt), name in zip(kernels.items(), demangled):
            short = re.sub(r"\(anonymous namespace\)::", "", name)
            short = re.sub(r"\(.*", "", short)
            f.write(short + "\t" + "\t".join(str(cnt.get(c, 0)) for c in cols) + "\n")
            tot.update(cnt)
        f.write("TOTAL\t" + "\t".join(str(tot.get(c, 0)) for c in cols) + "\n")
    print("wrote", path, {c: tot[c] for c in ("UBLKCP", "SYNCS", "DMMA", "STRONG.SYS", "MEMBAR.SYS")})


if __name__ == "__main__":
    main()
