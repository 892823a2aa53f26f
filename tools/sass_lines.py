"""Instructions per source-line bucket inside one function label of an `nvdisasm -c -g` dump.
Usage: python tools/sass_lines.py <dump> <label-substring> [bucket]"""
import collections
import re
import sys


def main():
    dump, pat = sys.argv[1], sys.argv[2]
    bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    cur = None
    line_no, fname = 0, ""
    cnt = collections.Counter()
    inl = False
    for line in open(dump):
        m = re.match(r"^\s*(\.text\.(\S+)|(\$\S+)):\s*$", line)
        if m:
            cur = m.group(2) or m.group(3)
            continue
        m = re.match(r'^\s*//## File "([^"]+)", line (\d+)(.*)$', line)
        if m:
            fname, line_no = m.group(1).split("/")[-1], int(m.group(2))
            inl = "inlined at" in m.group(3)
            # for inlined code attribute to the outermost call site when it is given on the same line
            m2 = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            if m2:
                fname, line_no = m2[-1][0].split("/")[-1], int(m2[-1][1])
            continue
        if cur and pat in cur and re.match(r"^\s+/\*[0-9a-f]{4,6}\*/\s+\S", line):
            cnt[(cur.split("$")[-1][:40], fname, line_no // bucket * bucket)] += 1
    for k in sorted(cnt):
        print(f"{cnt[k]:5d}  {k[0]:40s} {k[1]}:{k[2]}")
    print(sum(cnt.values()))


if __name__ == "__main__":
    main()
