"""Summarise `ncu --page source --csv` output: stall reasons and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if len(r) > 2 and r[1] == "Source")
hdr = rows[h]
body = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[1] != "Source"]
i_src = hdr.index("Source"); i_s = hdr.index("# Samples")
cols = ["stall_barrier", "stall_long_sb", "stall_lg", "stall_mio", "stall_short_sb", "stall_wait", "stall_membar",
        "stall_sleep", "stall_math", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_drain"]
ci = [hdr.index(c) for c in cols]
tot = sum(int(r[i_s]) for r in body)
print("total samples", tot)
print({c: sum(int(r[i]) for r in body) for c, i in zip(cols, ci)})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(body, key=lambda r: -int(r[i_s]))[:n]:
    print(r[i_s].rjust(6), r[i_src][:90].ljust(90), {c: int(r[i]) for c, i in zip(cols, ci) if int(r[i]) > 0})
