"""Summarise an ncu report into profiles/: key metrics per kernel, dynamic instruction mix, and
profiles/traffic.json (DRAM bytes per symbol per kernel, read by bench.py's roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep <tag> <symbols_per_launch> "<command that was profiled>"
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_warps", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_branch_resolving"]


def source_hash(kernel: str) -> str:
    """bench.py's kernel_source_hash(kernel): the digest of the sources that kernel is built from."""
    sys.path.insert(0, ROOT)
    import bench
    return bench.kernel_source_hash(kernel)


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, tag, n_sym, cmd = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    out, traffic = {}, {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("flic::", "")
        short = name.split("<")[0]
        rec = {k: f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip() for k in KEEP if k in hdr}
        rd = to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        inst = float(r[hdr.index("smsp__inst_executed.sum")].replace(",", ""))
        rec["dram_bytes_per_symbol"] = round((rd + wr) / n_sym, 4)
        rec["warp_instructions_per_symbol"] = round(inst / n_sym, 4)
        src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{short}"]))))
        if len(src) > 2:
            h = src[1]
            ia, isrc = h.index("Instructions Executed"), h.index("Source")
            mix = collections.Counter()
            for s in src[2:]:
                if len(s) <= ia:
                    continue
                toks = s[isrc].split()
                op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
                mix[op.split(".")[0]] += int(s[ia])
            rec["thread_instructions_per_symbol_by_opcode"] = {op: round(32 * c / n_sym, 2) for op, c in mix.most_common(24)}
        out[name] = rec
        traffic[short] = {"dram_bytes_per_symbol": rec["dram_bytes_per_symbol"],
                          "warp_instructions_per_symbol": rec["warp_instructions_per_symbol"],
                          "issue_slots_busy_pct": float(rec.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "nan").split()[0]),
                          "from": f"profiles/{tag}_ncu_full_summary.json",
                          # bench.py quotes these numbers only while the kernel sources still hash to this
                          "source_hash": source_hash(short)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump({"command": cmd, "symbols_per_launch": n_sym, "kernels": out},
              open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    old.update(traffic)
    json.dump(old, open(tpath, "w"), indent=1)
    for k, v in out.items():
        print(k, v.get("gpu__time_duration.sum"), "dram B/sym", v["dram_bytes_per_symbol"], "warp-instr/sym",
              v["warp_instructions_per_symbol"], "issue", v.get("smsp__issue_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()
