"""Instruction mix of the epilogue group loop (the loop around the LDTM.x8) of one tc_main_kernel variant.
usage: sass_loop.py "<0, 0, true, true, true>" [dump]"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(root, "build/csrc/decoder_tc.o")
want = sys.argv[1]
# mangled template args: Li<FAM>ELi<PHASE>ELb<GRAD>ELb<NB1>ELb<FAST>E
vals = [v.strip() for v in want.strip("<>").split(",")]
enc = "".join(("Lb%dE" % (v == "true")) if v in ("true", "false") else "Li%sE" % v for v in vals)
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s+Function : ", sass)
body = [f for f in funcs if f.startswith("_ZN") and "tc_main_kernelI" + enc + "EEvNS" in f.split("\n")[0]]
assert len(body) == 1, len(body)
ins = []
for l in body[0].splitlines():
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
idx = [i for i, (a, s) in enumerate(ins) if "LDTM.x8" in s][0]
a0 = ins[idx][0]
top = end = None
for i in range(idx, len(ins)):
    m = re.search(r"BRA (0x[0-9a-f]+)", ins[i][1])
    if m and int(m.group(1), 16) <= a0 and "U.ANY" not in ins[i][1]:
        top, end = int(m.group(1), 16), ins[i][0]
        break
loop = [(a, s) for a, s in ins if top <= a <= end]
print(f"kernel {want}: {len(ins)} instructions, group loop {top:#x}..{end:#x} = {len(loop)} instructions")
c = collections.Counter()
for a, i in loop:
    t = i.split()
    op = t[1] if t[0].startswith("@") else t[0]
    c[op.split(".")[0]] += 1
print(", ".join(f"{k} {v}" for k, v in c.most_common()))
if len(sys.argv) > 2:
    print("\n".join(f"{a:05x} {i}" for a, i in loop))
