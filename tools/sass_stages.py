"""Per-stage instruction budget of a coder kernel: every SASS instruction is attributed to the stage of
the arithmetic it belongs to (by the inline chain nvdisasm reports for it, -lineinfo build), counted
statically and -- with an ncu report of a run -- dynamically, as thread-instructions per symbol.

    python tools/sass_stages.py <object.o> <kernel substring> [<report.ncu-rep> <symbols_per_launch>] [--json out.json]

Stages (functions of csrc/flic_core.cuh; an instruction goes to the innermost function of its
inline chain that is listed here, so dadd/dfma/rcp_cubic/round24 count for their caller):
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CORE = os.path.join(ROOT, "finalproject-losslessimagecompression_b200", "csrc", "flic_core.cuh")

STAGES = [
    ("model: widen mean / scale, 1/scale, window origin", ["make_model", "lower_of", "lower_of_d"]),
    ("quotient -> arg: half-bin point, division by scale, clamp, float rounding", ["div_by_scale", "arg_from_quotient", "half_bin_point", "part1_at", "clamp_mag128"]),
    ("expf (glibc's algorithm)", ["exp_core", "exp_tab_entry", "expf_glibc"]),
    ("1/(1+e), x A, float rounding, round to int", ["part1_from_arg"]),
    ("CDF glue: part2, the (lo, hi) pair", ["cdf_at", "cdf_pair"]),
    ("symbol guess (float logit + Newton step)", ["guess_symbol", "clamp_to_window", "guess_in_window"]),
    ("verify the guess / bracket search", ["decode_symbol", "decode_symbol_model", "decode_symbol_lean", "search_begin", "search_feed", "search_symbol", "search_symbol_slow"]),
    ("pop (decoder state update)", ["rans_pop", "rans_pop32"]),
    ("push (renorm test, reciprocal, divide, state update)", ["rans_push", "rans_push_rf", "rcp_biased_low", "push_reciprocal"]),
    ("table glue: grid / window checks", ["make_table", "make_table_lean"]),
    ("parameter guard", ["guard_init", "guard_note", "guard_flags", "params_ok", "param_flags"]),
]
KERNEL_STAGE = "kernel body: staging (cp.async, shared reads), word pull / emit, loop, stores"
PIPE = {
    "fp64": {"DFMA", "DADD", "DMUL", "DSETP"},
    "xu": {"MUFU", "F2F", "F2I", "I2F", "FRND", "F2FP", "I2I", "FLO", "POPC", "BREV"},
    "fma": {"FFMA", "FMUL", "FADD", "IMAD", "FSWZADD"},
    "lsu": {"LDG", "STG", "LDS", "STS", "LDC", "LDCU", "ATOMS", "ATOMG", "RED", "LDGSTS", "LDSM", "LDL", "STL", "LDGDEPBAR"},
    "ctl": {"BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "CALL", "RET", "NOP", "DEPBAR", "ERRBAR", "MEMBAR", "BRX", "YIELD", "BREAK"},
}


def pipe_of(op):
    for p, s in PIPE.items():
        if op in s:
            return p
    return "alu"


def function_table():
    """(start line, name) of every function of flic_core.cuh."""
    lines = open(CORE).read().splitlines()
    out = []
    for i, l in enumerate(lines):
        if re.match(r"\s*(FLIC_HD|__device__|static inline|static __device__)", l) and not l.strip().startswith("#"):
            text = " ".join(lines[i:i + 4])
            text = re.sub(r"^\s*(FLIC_HD|__device__|__noinline__|__forceinline__|static|inline|\s)+", "", text)
            m = re.search(r"([A-Za-z_]\w*)\s*\(", text)
            if m:
                out.append((i + 1, m.group(1)))
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    obj, key = args[0], args[1]
    rep, nsym = (args[2], int(args[3])) if len(args) > 3 else (None, 0)
    jout = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    funcs = function_table()
    stage_of_func = {f: s for s, fs in STAGES for f in fs}

    def func_at(line):
        name = None
        for start, f in funcs:
            if start <= line:
                name = f
            else:
                break
        return name

    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    # the inline chain of an instruction is the run of "//## File" lines in front of it (innermost
    # first); instructions with none in front keep the chain of the one before
    instrs = []          # (opcode, stage, text)
    cur, chain, pending = False, [], []
    for l in dis.splitlines():
        if l.startswith(".text."):
            cur = key in l
            chain, pending = [], []
            continue
        if not cur:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            if pending:
                chain, pending = pending, []
            toks = m.group(1).split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            stage = KERNEL_STAGE
            for f, ln in chain:
                if f == "flic_core.cuh":
                    fn = func_at(ln)
                    if fn in stage_of_func:
                        stage = stage_of_func[fn]
                        break
            instrs.append((op, stage, m.group(1)))
    weights = [1.0] * len(instrs)
    dynamic = False
    if rep:
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
        for b in out.split('"Kernel Name",')[1:]:
            if key.split("ILi")[0] not in b.split("\n", 1)[0]:      # ncu prints demangled names
                continue
            rows = list(csv.reader(io.StringIO(b.split("\n", 1)[1])))
            h = rows[0]
            iex = h.index("Instructions Executed")
            rows = [r for r in rows[1:] if len(r) > iex]
            if len(rows) != len(instrs):
                print(f"warning: ncu lists {len(rows)} instructions, the object {len(instrs)}: static counts only", file=sys.stderr)
                break
            weights = [float(r[iex]) * 32.0 / nsym for r in rows]
            dynamic = True
            break
    table = collections.OrderedDict()
    for (op, stage, _), w in zip(instrs, weights):
        t = table.setdefault(stage, {"static": 0, "dynamic": 0.0, "pipes": collections.Counter()})
        t["static"] += 1
        t["dynamic"] += w
        t["pipes"][pipe_of(op)] += w
    order = [s for s, _ in STAGES] + [KERNEL_STAGE]
    unit = "thread-instructions per symbol" if dynamic else "static instructions"
    print(f"| stage | static | {unit} | fp64 | xu | alu | fma | lsu | ctl |")
    print("|---|---|---|---|---|---|---|---|---|")
    tot = collections.Counter()
    res = {}
    for s in order:
        if s not in table:
            continue
        t = table[s]
        p = t["pipes"]
        print(f"| {s} | {t['static']} | {t['dynamic']:.1f} | " + " | ".join(f"{p.get(k, 0):.1f}" for k in ("fp64", "xu", "alu", "fma", "lsu", "ctl")) + " |")
        res[s] = {"static": t["static"], "dynamic": round(t["dynamic"], 2), "pipes": {k: round(v, 2) for k, v in p.items()}}
        tot["static"] += t["static"]
        tot["dynamic"] += t["dynamic"]
        for k, v in p.items():
            tot[k] += v
    print(f"| total | {tot['static']} | {tot['dynamic']:.1f} | " + " | ".join(f"{tot.get(k, 0):.1f}" for k in ("fp64", "xu", "alu", "fma", "lsu", "ctl")) + " |")
    if jout:
        json.dump({"object": os.path.basename(obj), "kernel": key, "unit": unit, "report": os.path.basename(rep) if rep else None,
                   "symbols_per_launch": nsym, "stages": res}, open(jout, "w"), indent=1)


if __name__ == "__main__":
    main()
