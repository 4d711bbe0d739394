"""Top stall sites of one kernel launch from an ncu source-page export.

    ncu -i prof.ncu-rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > src.csv
    python tools/ncu_hot.py src.csv [--top 40] [--dyn]

Prints total executed warp-instructions, the per-stall-reason sample totals, and the instructions with the most stall samples
(address, samples, executed count, dominant stall reasons, SASS).  --dyn adds the dynamic opcode histogram."""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ix = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    data = []
    for r in rows[h + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        f = lambda n: float(r[ix[n]] or 0)
        data.append((int(r[0], 16), r[ix["Source"]], f("# Samples"), f("Instructions Executed"), {n: f(n) for n in stall_cols},
                     f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Ideal"), f("Avg. Threads Executed")))
    tot_inst = sum(d[3] for d in data)
    tot_samp = sum(d[2] for d in data)
    print(f"executed warp-instructions {tot_inst:.4g}; samples {tot_samp:.0f}")
    agg = collections.Counter()
    for d in data:
        for n, v in d[4].items():
            agg[n] += v
    print("stall samples:", ", ".join(f"{n[6:]} {v / tot_samp * 100:.1f}%" for n, v in agg.most_common() if v))
    print(f"shared wavefronts {sum(d[5] for d in data):.4g} (ideal {sum(d[6] for d in data):.4g})")
    base = data[0][0]
    for d in sorted(data, key=lambda d: -d[2])[:top]:
        why = ", ".join(f"{n[6:]} {v:.0f}" for n, v in sorted(d[4].items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"  {d[0] - base:#07x} samp {d[2]:6.0f} ({d[2] / tot_samp * 100:4.1f}%) exec {d[3]:.3g} thr {d[7]:4.1f}  {d[1][:60]:60s} | {why}")
    if "--dyn" in sys.argv:
        c = collections.Counter()
        for d in data:
            op = re.sub(r"^@!?U?P\w+\s+", "", d[1]).split()[0].split(".")[0] if d[1] else "?"
            c[op] += d[3]
        print("dynamic opcodes:", ", ".join(f"{k} {v / tot_inst * 100:.1f}%" for k, v in c.most_common(30)))


if __name__ == "__main__":
    main()
