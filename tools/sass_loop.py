"""Static view of the hot loop of one kernel (development aid).

    python tools/sass_loop.py <file.o|.so> <kernel name substring> [--dump]

Finds every backward branch (a loop), reports the instruction count and opcode histogram of each
loop body, and of the one holding the most FP64 instructions in detail.  Cold blocks that the
compiler moved behind the loop are not counted, which is the point: this approximates the
per-iteration dynamic instruction count without a GPU.
"""
import collections
import re
import subprocess
import sys


def kernel_sass(obj, key):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, rows = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or key not in cur:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip(), cur))
    return rows


def opcode(text):
    toks = text.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    return op.split(".")[0]


def main():
    obj, key = sys.argv[1], sys.argv[2]
    rows = kernel_sass(obj, key)
    if not rows:
        raise SystemExit("kernel not found")
    names = sorted({r[2] for r in rows})
    if len(names) > 1:
        print("kernels matched:", names, "-> using", names[0])
        rows = [r for r in rows if r[2] == names[0]]
    loops = []
    for addr, text, _ in rows:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", text)
        if m and opcode(text) == "BRA":
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                loops.append((tgt, addr))
    print(f"{names[0][:80]}: {len(rows)} instructions, {len(loops)} loops")
    best = None
    for tgt, addr in loops:
        body = [r for r in rows if tgt <= r[0] <= addr]
        h = collections.Counter(opcode(t) for _, t, _ in body)
        fp64 = h["DFMA"] + h["DADD"] + h["DMUL"]
        print(f"  loop {tgt:#x}..{addr:#x}: {len(body)} instr, fp64 {fp64}")
        if best is None or fp64 > best[0]:
            best = (fp64, tgt, addr, body, h)
    if best:
        fp64, tgt, addr, body, h = best
        print("hot loop:", " ".join(f"{op}:{c}" for op, c in h.most_common()))
        if "--dump" in sys.argv:
            for a, t, _ in body:
                print(f"{a:05x}  {t}")


if __name__ == "__main__":
    main()
