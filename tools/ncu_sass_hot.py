"""Per-instruction view of one kernel in an ncu report: executed count per symbol, stall samples.

    python tools/ncu_sass_hot.py <report.ncu-rep> <kernel substring> <symbols_per_launch> [min_share]
Prints every SASS instruction whose executed count is >= min_share of a warp-row (32 symbols),
with its share and its stall samples, so loops and their costs can be read off.
"""
import csv, io, subprocess, sys
rep, key, nsym = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if key not in name:
        continue
    rows = list(csv.reader(io.StringIO(b.split("\n", 1)[1])))
    h = rows[0]
    ia, isrc, isamp, iex = h.index("Address"), h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    rowsn = nsym / 32.0
    tot = 0
    print(name[:90])
    for r in rows[1:]:
        if len(r) <= iex:
            continue
        ex = float(r[iex]); tot += ex
        top = sorted(((int(r[i]), h[i][6:]) for i in stall_cols if int(r[i]) > 0), reverse=True)[:2]
        print(f"{r[ia][-4:]} {ex / rowsn:6.2f} {int(r[isamp]):6d} {r[isrc].strip():70s} {top}")
    print("warp-instructions per warp-row:", tot / rowsn)
    break
