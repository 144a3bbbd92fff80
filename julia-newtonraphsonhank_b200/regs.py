"""Prints registers / spills per kernel from the nvcc -Xptxas -v logs written by build.py."""
import os, re, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for f in sorted(os.listdir(os.path.join(here, "build"))):
    if not f.endswith(".log"):
        continue
    s = open(os.path.join(here, "build", f)).read()
    items = re.findall(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", s)
    for name, stack, ss, sl, regs in items:
        dn = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dn = re.sub(r"\(.*", "", dn).replace("void hank::", "")
        if len(sys.argv) > 1 and sys.argv[1] not in dn:
            continue
        print(f"{dn:56s} regs={regs:4s} stack={stack:5s} spill_st={ss:5s} spill_ld={sl}")
