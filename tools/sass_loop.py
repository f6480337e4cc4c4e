"""Static look at the hottest loop of a kernel's SASS.

    python tools/sass_loop.py <lib.so> <mangled-or-substring kernel name> [--skip ADDR,ADDR,...] [--dump]

Finds the longest backward branch (the depth-plane loop), lists the forward branches inside it with the
number of instructions they jump over, and prints an opcode histogram of the loop.  With --skip, the
forward branches at the given addresses are assumed TAKEN (their bodies -- the rare re-fetch / convert
paths -- are left out), which gives the steady-state instruction mix per plane.
"""
import collections
import re
import subprocess
import sys


def main():
    lib, name = sys.argv[1], sys.argv[2]
    skip = set()
    dump = "--dump" in sys.argv
    if "--skip" in sys.argv:
        skip = {int(x, 16) for x in sys.argv[sys.argv.index("--skip") + 1].split(",")}
    auto = "--auto" in sys.argv
    syms = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", syms)
    body = None
    for b in blocks[1:]:
        fn = b.split("\n", 1)[0].strip()
        if name in fn:
            body = b
            print("kernel:", fn)
            break
    if body is None:
        raise SystemExit("kernel not found")
    ins = []
    for line in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.\S+)?\s+(?:!?\w+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            if best is None or a - tgt > best[1] - best[0]:
                best = (tgt, a)
    lo, hi = best
    loop = [(a, t) for a, t in ins if lo <= a <= hi]
    print("loop 0x%x..0x%x: %d instructions" % (lo, hi, len(loop)))
    fwd = []
    for a, t in loop:
        m = re.search(r"BRA(?:\.\S+)?\s+(?:!?\w+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt > a:
                n = sum(1 for b, _ in loop if a < b < tgt)
                fwd.append((a, tgt, n, t))
    for a, tgt, n, t in fwd:
        print("  fwd branch @0x%x -> 0x%x skips %3d   %s" % (a, tgt, n, t))
    if auto:   # assume every predicated forward branch that skips >= 12 instructions is taken
        skip |= {a for a, tgt, n, t in fwd if n >= 12 and t.startswith("@")}
    hist = collections.Counter()
    kept = 0
    until = -1
    for a, t in loop:
        if a < until:
            continue
        op = t.split()
        opn = op[1] if op[0].startswith("@") else op[0]
        opn = opn.split(".")[0]
        hist[opn] += 1
        kept += 1
        if dump:
            print("    %04x  %s" % (a, t))
        if a in skip:
            m = re.search(r"0x([0-9a-f]+)", t)
            until = int(m.group(1), 16)
    print("steady-state instructions: %d" % kept)
    for k, v in hist.most_common():
        print("  %-10s %d" % (k, v))


if __name__ == "__main__":
    main()
