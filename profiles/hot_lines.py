#!/usr/bin/env python
"""Dynamic instruction counts of one profiled kernel attributed to source lines, with inlining resolved.

Joins (by instruction index — same cubin, same order)
  * `ncu -i X.ncu-rep --page source --csv --print-source sass`   (per-SASS-instruction executed counts and stall samples)
  * `nvdisasm -gi X.sm_100a.cubin`                               (per-instruction source line + inlined-at chain)
and prints, per line of the chosen file at the chosen inline level, warp instructions executed, their share, the active
lanes per instruction and the stall samples.  Runs here, no GPU needed.

    python profiles/hot_lines.py sass.csv disi.txt <kernel-substring> <file.cu> [top]
"""
import collections
import csv
import re
import sys


def load_disasm(path, needle):
    frames, cur, on, out = [], [], False, []
    for line in open(path):
        if line.startswith("//---") and ".text." in line:
            on = needle in line
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if not cur or cur[-1][2]:   # a new chain starts after a terminal frame
                cur = []
            cur.append((m.group(1).split("/")[-1], int(m.group(2)), "inlined at" not in line))
        elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            out.append([(f, l) for f, l, _ in cur])
    return out


def main():
    sass_csv, dis, needle, fname = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    # the csv holds one section per profiled launch: a "Kernel Name" row, the column header, then one row per instruction;
    # SPCU_SECTION (default 0) picks the launch
    import os
    rows = list(csv.reader(open(sass_csv)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = int(os.environ.get("SPCU_SECTION", "0"))
    lo = starts[sec]
    hi = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
    print("#", rows[lo][1][:120])
    hdr = rows[lo + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
    chains = load_disasm(dis, needle)
    assert len(chains) == len(data), (len(chains), len(data))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    total = 0
    for r, ch in zip(data, chains):
        ie, te = int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]])
        ns = int(r[ix["# Samples"]])
        noinst = int(r[ix["stall_no_inst"]])
        total += ie
        key = None
        for f, l in reversed(ch):      # outermost frame first: the deepest line that is still inside `fname`
            if f == fname:
                key = (f, l)
                break                  # outermost occurrence = the kernel body / top-level stage line
        # prefer the innermost frame that lies in fname (stage functions live there)
        for f, l in ch:
            if f == fname:
                key = (f, l)
                break
        a = agg[key]
        a[0] += ie; a[1] += te; a[2] += ns; a[3] += noinst
    print(f"total warp instructions {total}")
    print(f"{'line':>28s} {'Minst':>9s} {'share%':>7s} {'lanes':>6s} {'samples':>8s} {'no_inst':>8s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{str(key):>28s} {a[0]/1e6:9.1f} {100*a[0]/total:7.2f} {a[1]/max(a[0],1):6.1f} {a[2]:8d} {a[3]:8d}")


if __name__ == "__main__":
    main()
