#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY.  Build-time extraction of kernel definitions from reference translation units that cannot
be compiled as a whole here (they drag in the GUI / Testbed).  Output goes to oracle/_ref/ (git-ignored, never committed);
nothing is copied into the repository.

    extract_ref.py <reference file> <output .inc> "<start regex> => <end regex>" ...

Each range starts at the first line matching <start regex> and ends just BEFORE the first later line matching
<end regex> (searching resumes where the previous range ended, so ranges must be given in file order)."""
import re
import sys


def main():
    src, dst, specs = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = open(src).read().split("\n")
    out, pos = [], 0
    for spec in specs:
        a, b = spec.split(" => ")
        ra, rb = re.compile(a), re.compile(b)
        i = next((k for k in range(pos, len(lines)) if ra.search(lines[k])), None)
        if i is None:
            sys.exit("extract_ref: start anchor not found: " + a)
        j = next((k for k in range(i + 1, len(lines)) if rb.search(lines[k])), None)
        if j is None:
            sys.exit("extract_ref: end anchor not found: " + b)
        out.append("// ---- %s:%d-%d ----" % (src, i + 1, j))
        out.extend(lines[i:j])
        pos = j
    open(dst, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
