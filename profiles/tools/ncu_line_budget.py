"""Instruction budget of one kernel from an ncu report: executed warp instructions per source line / per opcode.

    ncu -i capture.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python profiles/tools/ncu_line_budget.py src.csv --units 4194304 [--top 40]

`--units` = how many work units the launch processed (for K1: 32-particle chunks = particles / 32), so the numbers read as
"instructions per chunk".  ncu lists an inlined instruction once per line of its inline chain; rows are de-duplicated by
instruction address and attributed to the first (innermost) line, so the column sums to `smsp__inst_executed.sum`."""
import argparse
import collections
import csv
import re


def load(path):
    seen, cur, curfile = set(), None, None
    byline, byop, stall = collections.Counter(), collections.Counter(), collections.Counter()
    for r in csv.reader(open(path)):
        if len(r) >= 2 and r[0] == "File Path":
            curfile = r[1].split("/")[-1]
            continue
        if len(r) < 8:
            continue
        if r[0].isdigit():
            cur = (curfile, int(r[0]), r[1].strip()[:90])
            continue
        if not r[2].startswith("0x") or r[2] in seen:
            continue
        try:
            n = int(r[7])
        except ValueError:
            continue
        seen.add(r[2])
        op = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip()).split()[0].rstrip(";").split(".")[0]
        byline[cur] += n
        byop[op] += n
        try:
            stall[cur] += int(r[4])
        except ValueError:
            pass
    return byline, byop, stall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--units", type=float, required=True)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    byline, byop, stall = load(a.csv)
    tot = sum(byop.values())
    print(f"executed warp instructions: {tot}  = {tot / a.units:.1f} per unit")
    print("\nper opcode (per unit)")
    for k, v in byop.most_common(a.top):
        print(f"  {k:10s} {v / a.units:7.1f}")
    print("\nper source line (per unit, stall samples)")
    for k, v in byline.most_common(a.top):
        print(f"  {v / a.units:7.1f}  {stall[k]:6d}  {k[0]}:{k[1]}  {k[2]}")


if __name__ == "__main__":
    main()
