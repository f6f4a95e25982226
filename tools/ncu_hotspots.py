"""Summarises an `ncu --page source --csv` export (SASS view) of one kernel: stall samples by opcode and by stall reason, the hottest
instructions, and -- given an index range -- the share of one loop.  Used for profiles/r02_ncu_stream_w4_hotspots.txt:
    ncu -i stream_w4.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotspots.py src.csv [first last]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    head, data = rows[hdr], rows[hdr + 1:]
    ix = {h: i for i, h in enumerate(head)}
    S, SRC, IE = ix["# Samples"], ix["Source"], ix["Instructions Executed"]
    stalls = [h for h in head if h.startswith("stall_") and "Not Issued" not in h]
    num = lambda r, c: int(r[c] or 0) if c < len(r) else 0
    total = sum(num(r, S) for r in data)
    print(f"instructions (SASS lines): {len(data)}, samples: {total}, warp instructions executed: {sum(num(r, IE) for r in data)}")
    by_reason = collections.Counter()
    by_op = collections.Counter()
    for r in data:
        for c in stalls:
            by_reason[c[6:]] += num(r, ix[c])
        op = r[SRC].split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
        by_op[op.split(".")[0]] += num(r, S)
    print("\nsamples by stall reason:")
    for k, v in by_reason.most_common(10):
        print(f"  {k:<18} {v:>7} {100.0 * v / total:5.1f} %")
    print("\nsamples by opcode:")
    for k, v in by_op.most_common(16):
        print(f"  {k:<18} {v:>7} {100.0 * v / total:5.1f} %")
    print("\nhottest instructions (index, samples, executed, SASS, top stall reasons):")
    for i in sorted(range(len(data)), key=lambda i: -num(data[i], S))[:30]:
        r = data[i]
        st = sorted(((num(r, ix[c]), c[6:]) for c in stalls), reverse=True)[:2]
        print(f"  {i:>5} {num(r, S):>6} {num(r, IE):>9}  {r[SRC].strip()[:70]:<70} {[(n, v) for v, n in st if v]}")
    if len(sys.argv) >= 4:
        a, b = int(sys.argv[2]), int(sys.argv[3])
        sub = data[a:b + 1]
        n = sum(num(r, S) for r in sub)
        c = collections.Counter()
        for r in sub:
            for col in stalls:
                c[col[6:]] += num(r, ix[col])
        print(f"\ninstructions [{a}, {b}]: {n} samples = {100.0 * n / total:.1f} % of all, {sum(num(r, IE) for r in sub)} warp instructions")
        for k, v in c.most_common(6):
            print(f"  {k:<18} {v:>7} {100.0 * v / max(n, 1):5.1f} %")


if __name__ == "__main__":
    main()
