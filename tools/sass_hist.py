"""Static SASS opcode histogram of one kernel in an object file (development aid).

    python tools/sass_hist.py <file.o> <substring of the mangled kernel name> [--pipes]

Counts every instruction once (loops are not weighted), grouped by base opcode, and sums them by
the pipe they issue to on sm_100 (FP64 2 cycles/warp-instruction, XU 8, ALU 2, FMA 1).
"""
import collections
import re
import subprocess
import sys

PIPE = {
    "fp64": {"DFMA", "DADD", "DMUL", "DSETP"},
    "xu": {"MUFU", "F2F", "F2I", "I2F", "FRND", "F2FP", "I2I"},
    "fma": {"FFMA", "FMUL", "FADD", "IMAD", "FSWZADD"},
    "lsu": {"LDG", "STG", "LDS", "STS", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "LDGSTS", "LDSM", "LDL", "STL"},
    "ctl": {"BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "CALL", "RET", "NOP", "DEPBAR", "ERRBAR", "MEMBAR", "BRX", "YIELD", "LDGDEPBAR"},
}


def main():
    obj, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur = None
    hist = collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or key not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        hist[op.split(".")[0]] += 1
    total = sum(hist.values())
    pipes = collections.Counter()
    for op, c in hist.items():
        for p, s in PIPE.items():
            if op in s:
                pipes[p] += c
                break
        else:
            pipes["alu"] += c
    print(" ".join(f"{op}:{c}" for op, c in hist.most_common()))
    print("total", total, dict(pipes))


if __name__ == "__main__":
    main()
