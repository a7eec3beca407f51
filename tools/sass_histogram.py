"""Per-kernel SASS opcode histogram of the in-tree library (`cuobjdump -sass`): the evidence that the attention kernels
are tcgen05 / TMA / TMEM code (UTCHMMA, UTMALDG, UTMAREDG, LDTM / STTM) and not recompiled mma.sync.

    python tools/sass_histogram.py [round-tag]        # writes profiles/<tag>_sass_histogram.txt  (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "modaltune_b200", "libmodaltune_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UTCCP", "LDTM", "STTM", "UTCATOM",
       "SYNCS", "MUFU", "HMMA", "FFMA2", "FFMA", "FADD2", "FMUL2", "F2FP", "FMNMX3", "LDS", "STS", "LDG", "STG", "RED", "ATOMG", "BAR")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["__total__"] += 1
            if m.group(1) in ("UTMALDG", "UTMAREDG", "UTCHMMA", "MUFU"):
                kernels[cur][m.group(1) + m.group(2)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = [f"# cuobjdump -sass {os.path.relpath(SO, ROOT)} -- opcode counts per kernel (static instruction counts)",
           "# tcgen05 = UTCHMMA (MMA), UTCBAR (commit), LDTM / STTM (tcgen05.ld / .st); TMA = UTMALDG / UTMAREDG / UTMASTG", ""]
    for (name, cnt), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm)
        picks = [(k, cnt[k]) for k in KEY if cnt[k]]
        sub = [(k, v) for k, v in sorted(cnt.items()) if "." in k]
        out.append(f"{short}\n    total {cnt['__total__']}  " + "  ".join(f"{k} {v}" for k, v in picks))
        if sub:
            out.append("    " + "  ".join(f"{k} {v}" for k, v in sub))
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.txt")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print(path)


if __name__ == "__main__":
    main()
