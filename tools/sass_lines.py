"""Static SASS budget of one kernel by source line (no GPU needed).

    python tools/sass_lines.py <object-or-cubin> <kernel-name-substring> [--file pic_pair.cuh] [--top 40] [--range 300:420]

Disassembles with `nvdisasm -g` (needs -lineinfo at compile time), attributes every instruction to the `//## File ..., line N`
marker in force, and prints instruction counts per (file, line) plus an opcode histogram.  Static counts: a line inside a loop
counts once.  Inlined functions are attributed to the line of the inlined body (nvdisasm's innermost location).
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile


def disasm(path, kernel):
    if path.endswith(".o") or path.endswith(".so"):
        d = tempfile.mkdtemp()
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(path)], cwd=d, check=True, capture_output=True)
        cubins = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")]
    else:
        cubins = [path]
    for c in cubins:
        out = subprocess.run(["nvdisasm", "-g", c], capture_output=True, text=True).stdout
        secs = re.split(r"(?m)^\s*\.section\s+\.text\.", out)
        for s in secs[1:]:
            name = s.split(",", 1)[0]
            if kernel in name:
                yield name, s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path"); ap.add_argument("kernel")
    ap.add_argument("--file", default=None); ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--range", default=None, help="only lines lo:hi of --file, listed in order")
    ap.add_argument("--first", action="store_true", help="only the first matching kernel")
    a = ap.parse_args()
    for name, sec in disasm(a.path, a.kernel):
        cur = ("?", 0)
        per_line = collections.Counter(); ops = collections.Counter(); per_file = collections.Counter()
        per_line_ops = collections.defaultdict(collections.Counter)
        for l in sec.splitlines():
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
            if not m:
                continue
            ins = re.sub(r"^@!?U?P\w+\s+", "", m.group(1)).split()[0]
            op = ins.split(".")[0]
            if op in ("NOP",):
                continue
            per_line[cur] += 1; ops[op] += 1; per_file[cur[0]] += 1; per_line_ops[cur][ins] += 1
        print(f"== {name[:110]}\n   total {sum(ops.values())} instructions; by file: {dict(per_file)}")
        print("   opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(45)))
        if a.range and a.file:
            lo, hi = (int(v) for v in a.range.split(":"))
            tot = 0
            for (f, ln), n in sorted(per_line.items()):
                if f == a.file and lo <= ln <= hi:
                    tot += n
                    print(f"   {f}:{ln:5d} {n:5d}  " + " ".join(f"{k}x{v}" for k, v in per_line_ops[(f, ln)].most_common(8)))
            print(f"   -- {tot} instructions in {a.file}:{lo}-{hi}")
        else:
            for (f, ln), n in per_line.most_common(a.top):
                if a.file and f != a.file:
                    continue
                print(f"   {f}:{ln:5d} {n:5d}  " + " ".join(f"{k}x{v}" for k, v in per_line_ops[(f, ln)].most_common(6)))
        if a.first:
            break


if __name__ == "__main__":
    main()
