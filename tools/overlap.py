"""Normalised line overlap between a file of this repo and a reference file (whitespace collapsed, comments and blank lines
dropped, lines of <= 3 characters ignored): the check VERDICT.md applies to spot transliterations."""
import re, sys

def lines(path):
    txt = open(path, errors="replace").read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = []
    for l in txt.splitlines():
        l = re.sub(r"//.*", "", l)
        l = re.sub(r"\s+", "", l)
        if len(l) > 3:
            out.append(l)
    return out

mine, ref = lines(sys.argv[1]), set(lines(sys.argv[2]))
same = [l for l in mine if l in ref]
print("%s vs %s: %d of %d substantive lines identical (%.0f %%)" % (sys.argv[1], sys.argv[2], len(same), len(mine), 100.0 * len(same) / max(len(mine), 1)))
if "-v" in sys.argv:
    for l in same: print("   ", l)
