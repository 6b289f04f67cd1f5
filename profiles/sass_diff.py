#!/usr/bin/env python
"""Per-kernel SASS comparison of two builds of libd3h_tets.so (cuobjdump -sass, encodings and addresses stripped):

    git worktree add /tmp/wt <commit> && (cd /tmp/wt && python d3human-code_b200/build.py --force)
    python profiles/sass_diff.py /tmp/wt/d3human-code_b200/lib/libd3h_tets.so d3human-code_b200/lib/libd3h_tets.so

Used when kernels are edited without a GPU at hand: shows which kernels of the measured build are byte-identical, which
differ (and in which instructions) and which are new.  `old=new` pairs a renamed kernel: --pair 'd3h::poly_cut_kernel=void d3h::poly_cut_kernel<false>'."""
import argparse
import re
import subprocess


def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            funcs[cur] = []
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            ins = re.sub(r"/\*[0-9a-f]{4}\*/", "", line, count=1)
            ins = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ins).strip()
            if ins:
                funcs[cur].append(ins)
    return funcs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("old")
    ap.add_argument("new")
    ap.add_argument("--pair", action="append", default=[], help="OLDNAME=NEWNAME for a renamed kernel")
    ap.add_argument("--show", type=int, default=4, help="differing lines to print per kernel")
    args = ap.parse_args()
    old, new = sass(args.old), sass(args.new)
    renamed = dict(p.split("=", 1) for p in args.pair)
    for name in sorted(set(old) | set(new)):
        target = renamed.get(name, name)
        a, b = old.get(name), new.get(target)
        if name in renamed.values() and name not in old:
            continue
        if a is None:
            print(f"NEW      {name}  ({len(b)} instr)")
        elif b is None:
            print(f"GONE     {name}  ({len(a)} instr)")
        elif a == b:
            print(f"SAME     {name}{' -> ' + target if target != name else ''}  ({len(a)} instr)")
        else:
            diffs = [(x, y) for x, y in zip(a, b) if x != y]
            print(f"CHANGED  {name}{' -> ' + target if target != name else ''}  ({len(a)} -> {len(b)} instr, {len(diffs)} lines differ)")
            for x, y in diffs[:args.show]:
                print("    -", x[:110])
                print("    +", y[:110])


if __name__ == "__main__":
    main()
