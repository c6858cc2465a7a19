"""Developer helper: print (or summarise) the SASS of the kernels of a library whose mangled name matches a regex.
    python scripts/sass_of.py LIB REGEX [--stat]     # --stat: instruction count and opcode histogram only"""
import collections
import re
import subprocess
import sys

lib, pat = sys.argv[1], re.compile(sys.argv[2])
stat = "--stat" in sys.argv
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, body = None, []


def flush():
    if name is None or not pat.search(name):
        return
    ins = [l for l in body if re.match(r"\s*/\*[0-9a-f]{4}\*/", l)]
    print(f"=== {name}: {len(ins)} instructions")
    if stat:
        ops = collections.Counter()
        for l in ins:
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
            ops[m.group(2) if m else "?"] += 1
        print("   ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(30)))
    else:
        for l in ins:
            print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))


for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        flush()
        name, body = m.group(1), []
    else:
        body.append(line)
flush()
