#!/usr/bin/env python3
"""SASS instruction mix of selected functions of a built object (cuobjdump -sass), as a markdown table.

usage: python tools/sass_mix.py mina_bridge_b200/build/msm_fq.cu.o k_accumulate mul_call sqr_call > profiles/<name>.md
"""
import collections
import re
import subprocess
import sys


def main():
    obj, wanted = sys.argv[1], sys.argv[2:]
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    fn, mix = None, collections.OrderedDict()
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and fn:
            for w in wanted:
                if w in fn:
                    op = m.group(1) + (".WIDE" if ".WIDE" in m.group(2) else "") + (".HI" if ".HI" in m.group(2) else "")
                    mix.setdefault(fn, collections.Counter())[op] += 1
    print("# SASS instruction mix (cuobjdump -sass %s)\n" % obj)
    for fn, c in mix.items():
        total = sum(c.values())
        short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        print("## `%s` -- %d instructions\n" % (short[:140], total))
        print("| mnemonic | count | share |\n|---|---:|---:|")
        for op, n in c.most_common(14):
            print("| %s | %d | %.1f %% |" % (op, n, 100.0 * n / total))
        print()


if __name__ == "__main__":
    main()
