"""A/B harness for kernel variants selected by compile-time macros (the PIC_K9_* / PIC_* switches in csrc/).

    # here (no GPU): build the variants next to the product library, git-ignored but shipped to the GPU box
    python tools/ab.py build base jtsync:-DPIC_K9_JT_SYNC=1 nw16:-DPIC_K9_CTAS=1,-DPIC_K9_NW=16

    # on the GPU box (one gpurun call): time each variant with the product bench, same box, back to back
    gpurun --timeout 900 -- 'python tools/ab.py run base jtsync nw16 -- --steps 20 --warmup 5'

    # parity of a variant before believing its time (GPU box): the tile-kernel parity tests with PIC_B200_LIB pointing at it
    gpurun --timeout 900 -- 'python tools/ab.py test jtsync@PIC_K9_JTILE=1 nw16'

`build` writes build_ab/<name>/libpic_b200.so (the product library is untouched); `test` runs
`pytest -m gpu tests/test_gpu_parity.py tests/test_golden.py -k "tile or variant or golden or supercell"` per variant; `run` executes
`bench.py --no-e2e --no-cpu-baseline` once per variant with PIC_B200_LIB pointing at it (extra environment as name@K=V,
e.g. jt@PIC_K9_JTILE=1), stores gpurun_out/ab_<name>.json and prints one comparison line per variant.  A variant is only
adopted after `pytest -m gpu` has passed with PIC_B200_LIB set to it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def split(spec):
    name, _, rest = spec.partition(":")
    name, *env = name.split("@")
    return name, [f for f in rest.split(",") if f], dict(e.split("=", 1) for e in env)


def lib_of(name):
    return os.path.join(ROOT, "build_ab", name, "libpic_b200.so")


def build(specs):
    from pypic3d_b200 import _lib
    for spec in specs:
        name, flags, _ = split(spec)
        os.makedirs(os.path.dirname(lib_of(name)), exist_ok=True)
        _, out = _lib.build(force=True, extra_flags=tuple(flags), out=lib_of(name), ptxas_verbose=True)
        spilled, entry, regs = [], "", {}
        for l in out.splitlines():                                   # ptxas -v: "Compiling entry function '<name>'" ... "N bytes spill stores"
            if "Compiling entry function" in l:
                entry = l.split("'")[1]
            elif "spill stores" in l and "0 bytes spill stores, 0 bytes spill loads" not in l:
                spilled.append(entry)
            elif "Used " in l and "registers" in l:
                regs[entry] = int(l.split("Used ")[1].split(" ")[0])
        hot = [e for e in spilled if "k_tile3d" in e or "k_pair3d" in e]
        hot_regs = sorted({v for e, v in regs.items() if "k_pair3d" in e})
        print(f"{name}: built {lib_of(name)} ({' '.join(flags) or 'default flags'}); spilling hot kernels: {len(hot)}, "
              f"others {len(spilled) - len(hot)} (the general-configuration kernels spill in the default build too); k_pair3d registers {hot_regs}")


def run(specs, bench_args):
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for spec in specs:
        name, _, env = split(spec)
        e = dict(os.environ, PIC_B200_LIB=lib_of(name), **env)
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-e2e", "--no-cpu-baseline", "--no-second-leg", "--no-check"] + bench_args
        r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=600)
        line = next((l for l in reversed(r.stdout.splitlines()) if l.startswith("{")), None)
        if line is None:
            print(f"{name}: FAILED rc={r.returncode} {r.stderr[-400:]}")
            continue
        with open(os.path.join(ROOT, "gpurun_out", f"ab_{name}.json"), "w") as f:
            f.write(line + "\n")
        d = json.loads(line)
        rf = d["roofline"]
        print(f"{name:12s} step {d['ms_per_step']:.3f} ms  K1 {rf['avg_launch_ms']:.3f} ms {rf.get('avg_launch_ms_by_species')}  "
              f"frac {rf['frac']:.3f}  sm {d['clocks']['sm_mhz']} MHz {d['clocks']['reasons']}")


def test(specs, pytest_args):
    for spec in specs:
        name, _, env = split(spec)
        e = dict(os.environ, PIC_B200_LIB=lib_of(name), **env)
        cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "tests/test_gpu_parity.py", "tests/test_golden.py",
               "-k", "tile or variant or golden or supercell"] + pytest_args
        r = subprocess.run(cmd, env=e, capture_output=True, text=True, timeout=1500, cwd=ROOT)
        tail = (r.stdout.strip().splitlines() or ["(no output)"])[-1]
        print(f"{name:12s} pytest rc={r.returncode}  {tail}")
        if r.returncode != 0:
            print(r.stdout[-1500:])


def main(argv):
    if len(argv) < 2 or argv[0] not in ("build", "run", "test"):
        sys.exit(__doc__)
    rest = argv[1:]
    extra = []
    if "--" in rest:
        i = rest.index("--")
        rest, extra = rest[:i], rest[i + 1:]
    {"build": build, "run": lambda s_: run(s_, extra), "test": lambda s_: test(s_, extra)}[argv[0]](rest)


if __name__ == "__main__":
    main(sys.argv[1:])
