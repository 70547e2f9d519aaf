"""Instruction mix of a source-line range of one kernel, from `nvdisasm -g -c <cubin>` (needs -lineinfo).

    python scripts/sass_mix.py <cubin> <kernel> <first line> <last line> [<source file, to print per-line counts>]
"""
import collections
import re
import subprocess
import sys

cubin, kernel, l0, l1 = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
inside, ln = False, 0
mix, per_line = collections.Counter(), collections.Counter()
for row in out:
    m = re.match(r"\s*\.section\s+\.text\.([A-Za-z0-9_$]+)", row)
    if m:
        inside = m.group(1) == kernel
        continue
    m = re.search(r'//## File ".*", line (\d+)', row)
    if m:
        ln = int(m.group(1))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", row)
    if m and inside and l0 <= ln <= l1:
        mix[m.group(2)] += 1
        per_line[ln] += 1
print(sum(mix.values()), "instructions of", kernel, "in lines", l0, "..", l1)
print(", ".join(f"{k}:{v}" for k, v in mix.most_common(25)))
if len(sys.argv) > 5:
    src = open(sys.argv[5]).read().split("\n")
    for l, c in sorted(per_line.items()):
        print(f"{c:4d}  L{l}: {src[l-1].strip()[:110]}")
