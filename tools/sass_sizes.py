"""Instruction counts of a kernel's main body and of each out-of-line device function inside it
(from `nvdisasm -c` labels).  Usage: python tools/sass_sizes.py <object-or-so> <kernel-substring>"""
import glob
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
        for cubin in sorted(glob.glob(os.path.join(td, "*.cubin"))):
            txt = subprocess.run(["nvdisasm", "-c", cubin], capture_output=True, text=True).stdout
            cur, counts, order = None, {}, []
            for line in txt.splitlines():
                m = re.match(r"^\s*(\.text\.(\S+)|(\$\S+)):\s*$", line)
                if m:
                    cur = m.group(2) or m.group(3)
                    if cur not in counts:
                        counts[cur] = 0
                        order.append(cur)
                    continue
                if re.match(r"^\s*\.section", line):
                    cur = None
                if cur and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
                    counts[cur] += 1
            tot = 0
            for name in order:
                if pat in name:
                    short = name.split("$")[-1] if name.startswith("$") else "<kernel body> " + name
                    print(f"{counts[name]:6d}  {short}")
                    tot += counts[name]
            print(f"{tot:6d}  total ({tot * 16 / 1024:.1f} KB)")


if __name__ == "__main__":
    main()
