"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass): the evidence that the tensor kernels are
tcgen05 / TMEM / TMA code (UTCHMMA, LDTM / STTM, UTMALDG / UTMASTG, UTCBAR), with the MUFU / spill counts of the
epilogues.  Runs in the build container (no GPU):   python tools/sass_hist.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neucam_b200", "libneucam_b200.so")
KEEP = re.compile(r"^(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UTMAPF|UTMACCTL|LDTM|STTM|UTCATOM|MUFU|HMMA|FFMA|FMUL|F2FP|"
                  r"LDS|STS|LDG|STG|LDL|STL|ATOM|RED|SYNCS|ELECT|R2UR|BAR|USETMAXREG|UBLKPF|UBLKCP|CCTL|MEMBAR|ERRBAR)")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (sm_100a), per kernel; selected opcode families")
    print("# (total = all SASS instructions of the kernel; ATOM/RED rows appear only if a kernel still had atomics)")
    for fn, c in per.items():
        demangled = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        short = re.sub(r"\(anonymous namespace\)::", "", demangled)
        short = re.sub(r"\(.*", "", short)
        total = sum(c.values())
        fam = collections.Counter()
        for op, n in c.items():
            if KEEP.match(op):
                key = op if op.startswith(("UTC", "UTMA", "UBLK", "LDTM", "STTM", "MUFU")) else op.split(".")[0]
                fam[key] += n
        print(f"\n{short}   total {total}")
        for key, n in sorted(fam.items(), key=lambda kv: (-kv[1], kv[0])):
            print(f"    {n:6d}  {key}")


if __name__ == "__main__":
    main()
