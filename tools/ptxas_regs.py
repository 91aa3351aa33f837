"""Registers / spills per decoder kernel variant, from the ptxas log the Makefile keeps."""
import os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
log = open(os.path.join(root, "build/csrc", (sys.argv[1] if len(sys.argv) > 1 else "decoder_tc") + ".ptxas.log")).read()
ents = re.findall(r"Compiling entry function '([^']+)'.*?\n(.*?)ptxas info\s+: Used (\d+) registers", log, re.S)
for name, mid, regs in ents:
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"glue::\(anonymous namespace\)::|\(glue.*", "", dem).replace("void ", "")
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", mid)
    print(f"{dem:60s} regs {regs:>3s}  stack/spill {sp.groups() if sp else None}")
