#!/usr/bin/env python
"""Static SASS size of one kernel attributed to source lines (nvdisasm --print-line-info of the cubin; runs here, no GPU):

    cuobjdump -xelf all build/obj/csrc/X.o && nvdisasm --print-line-info X.sm_100a.cubin > dis.txt
    python profiles/sass_by_line.py dis.txt <kernel-name-substring> [top]
"""
import collections
import re
import sys


def main() -> None:
    path, needle = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    cnt, cur, on = collections.Counter(), None, False
    for line in open(path):
        if line.startswith("//---") and ".text." in line:
            on = needle in line
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
        elif re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
            cnt[cur] += 1
    total = sum(cnt.values())
    by_file = collections.Counter()
    for (f, _), c in cnt.items():
        by_file[f] += c
    print("instructions", total, "=", total * 16 // 1024, "KB")
    print(by_file.most_common())
    for k, c in cnt.most_common(top):
        print(k, c)


if __name__ == "__main__":
    main()
