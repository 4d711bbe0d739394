"""Instruction count of the hot path of a kernel from its SASS (no GPU needed).

    python tools/sass_path.py <object-or-cubin> <kernel-substring> [--dump]

Finds the largest basic-block run that contains the bulk of the LDS gather (the pair / tile body), prints its opcode histogram and
the per-pipe totals the B200 roofline of this kernel cares about: issue slots, LSU instructions (LDS/STS/LDG/STG/RED/SHFL/ATOM),
FMA-pipe work (packed ops count twice), XU ops (MUFU / conversions)."""
import collections
import re
import subprocess
import sys


def sass_of(path, kernel):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, res = None, {}
    for l in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", l)
        if m:
            cur = m.group(1)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return {k: v for k, v in res.items() if kernel in k}


LSU = ("LDS", "STS", "LDG", "STG", "RED", "REDG", "ATOM", "ATOMS", "ATOMG", "SHFL", "LDL", "STL", "LD", "ST", "MATCH")
XU = ("MUFU", "F2I", "I2F", "FRND", "F2F", "POPC", "FLO", "BREV")


def main():
    path, kernel = sys.argv[1], sys.argv[2]
    for name, ins in sass_of(path, kernel).items():
        # blocks = maximal runs without a branch / barrier instruction
        blocks, cur = [], []
        for a, t in ins:
            cur.append((a, t))
            if re.search(r"\b(BRA|EXIT|RET|CALL|BSYNC|WARPSYNC)\b", t):
                blocks.append(cur); cur = []
        if cur:
            blocks.append(cur)
        best = max(blocks, key=lambda b: sum(1 for _, t in b if re.search(r"\bLDS\b", t)))
        c = collections.Counter()
        for _, t in best:
            t = re.sub(r"^@!?U?P\w+\s+", "", t)
            c[t.split()[0].split(".")[0]] += 1
        n = sum(c.values())
        lsu = sum(v for k, v in c.items() if k in LSU)
        packed = sum(v for k, v in c.items() if k in ("FADD2", "FMUL2", "FFMA2"))
        fma = sum(v for k, v in c.items() if k in ("FADD", "FMUL", "FFMA", "IMAD", "DFMA", "DADD", "DMUL")) + 2 * packed
        xu = sum(v for k, v in c.items() if k in XU)
        print(f"== {name[:100]}\n   hot block {best[0][0]:#x}-{best[-1][0]:#x}: {n} instructions; LSU {lsu}, FMA-pipe {fma} (packed {packed}), XU {xu}")
        print("   " + ", ".join(f"{k} {v}" for k, v in c.most_common()))
        if "--dump" in sys.argv:
            for a, t in best:
                print(f"   {a:05x} {t}")
        if "--all" not in sys.argv:
            break


if __name__ == "__main__":
    main()
