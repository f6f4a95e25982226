"""SASS evidence of the Blackwell paths in the built library: per kernel, how often the mnemonics that prove tcgen05 / TMEM / TMA /
legacy mma / bulk copies occur (`/opt/skills/guides/B200_PROFILING.md` lists them).  Runs here (no GPU):
    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "metalchat_b200" / "libmc_cuda.so"
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "HMMA", "LDSM", "SYNCS", "UTCCP", "REDG", "ATOMG", "STG.E.STRONG.SYS", "LDG.E.STRONG.SYS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name)
            counts[name] = collections.Counter()
            continue
        if name is None or "/*" not in line:
            continue
        counts[name]["instructions"] += 1
        for k in KEYS:
            if re.search(r"\b" + re.escape(k), line):
                counts[name][k] += 1
                total[k] += 1
    print(f"{LIB.name}: {len(counts)} kernels, {sum(c['instructions'] for c in counts.values())} SASS instructions")
    print("totals: " + ", ".join(f"{k} {total[k]}" for k in KEYS if total[k]))
    print()
    for n, c in counts.items():
        hits = ", ".join(f"{k} {c[k]}" for k in KEYS if c[k])
        if hits:
            print(f"{n[:150]}\n    {c['instructions']} instr: {hits}")


if __name__ == "__main__":
    sys.exit(main())
