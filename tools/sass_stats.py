"""Per-kernel SASS statistics of libvsiq.so: instruction count, local-memory (STL/LDL) accesses, registers, and the
Blackwell-specific mnemonics (256-bit global accesses, bulk async copies, mbarrier waits).  Runs without a GPU.

    python tools/sass_stats.py [--lib path] [--filter substr] [--all]
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "vsiquantization_b200", "libvsiq.so"))
    ap.add_argument("--filter", default="")
    ap.add_argument("--all", action="store_true", help="also list kernels without local-memory accesses")
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", a.lib], capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    cur, stats = None, collections.OrderedDict()
    keys = ("STL", "LDL", "LDG.E.256", "LDG.E.ENL2.256", "STG.E.256", "STG.E.ENL2.256", "UBLKCP", "SYNCS", "MUFU.RCP", "CALL")
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
            continue
        if cur is None or not re.search(r"^\s+/\*[0-9a-f]{4}\*/", line):
            continue
        st = stats[cur]
        st["n"] += 1
        if re.search(r"\bSTL\b|\bSTL\.", line): st["STL"] += 1
        if re.search(r"\bLDL\b|\bLDL\.", line): st["LDL"] += 1
        if re.search(r"\bLDG\S*\.256", line): st["LDG256"] += 1
        if re.search(r"\bLDG\S*\.128", line): st["LDG128"] += 1
        if re.search(r"\bSTG\S*\.256", line): st["STG256"] += 1
        if re.search(r"\bSTG\S*\.128", line): st["STG128"] += 1
        if "UBLKCP" in line: st["UBLKCP"] += 1
        if "SYNCS" in line: st["SYNCS"] += 1
        if "MUFU.RCP" in line: st["RCP"] += 1
        if re.search(r"\bCALL\b", line): st["CALL"] += 1
        if "ACQBULK" in line or "griddepcontrol" in line.lower() or "PREEXIT" in line: st["PDL"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
    print(f"{'insts':>6} {'regs':>4} {'stack':>5} {'STL':>4} {'LDL':>4} {'LDG256':>6} {'STG256':>6} {'LDG128':>6} {'STG128':>6} {'UBLKCP':>6} {'SYNCS':>5} {'RCP':>3} {'CALL':>4}  kernel")
    for (k, st), name in zip(stats.items(), names):
        if a.filter and a.filter not in name:
            continue
        if not a.all and not (st["STL"] or st["LDL"]):
            continue
        r, stack = regs.get(k, (-1, -1))
        name = re.sub(r"\(.*", "", name)
        print(f"{st['n']:6d} {r:4d} {stack:5d} {st['STL']:4d} {st['LDL']:4d} {st['LDG256']:6d} {st['STG256']:6d} {st['LDG128']:6d} {st['STG128']:6d} "
              f"{st['UBLKCP']:6d} {st['SYNCS']:5d} {st['RCP']:3d} {st['CALL']:4d}  {name}")
    tot = collections.Counter()
    for st in stats.values():
        tot.update(st)
    print(f"# {len(stats)} functions; totals: " + ", ".join(f"{k}={v}" for k, v in sorted(tot.items())))


if __name__ == "__main__":
    main()
