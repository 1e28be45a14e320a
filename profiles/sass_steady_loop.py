#!/usr/bin/env python
"""SASS excerpt of a streaming kernel's steady loop: dumps the function with cuobjdump, finds the
smallest backward-branch body that holds a whole round (>= 200) of FMUL2 / FFMA2 instructions, prints the body's opcode
histogram and its first / last lines.
usage: python profiles/sass_steady_loop.py <object> <mangled function substring> [lines]"""
import collections, re, subprocess, sys

obj, pat = sys.argv[1], sys.argv[2]
nshow = int(sys.argv[3]) if len(sys.argv) > 3 else 60
names = [l.split("Function : ")[1].strip() for l in
         subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines() if "Function : " in l]
fn = [n for n in names if pat in n][0]
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, obj], capture_output=True, text=True).stdout
ins = []
for l in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
best = None
for i, (a, s) in enumerate(ins):
    m = re.search(r"BRA\s+0x([0-9a-f]+)", s)
    if m and not s.startswith("BRA.DIV"):
        t = int(m.group(1), 16)
        if t < a and t in addr:
            body = ins[addr[t]:i + 1]
            fp = sum(1 for _, x in body if re.search(r"\b(FMUL2|FFMA2)\b", x))
            # the innermost loop that holds a whole round of packed arithmetic
            if fp >= 200 and (best is None or len(body) < len(best[3])):
                best = (fp, t, a, body)
fp, t, a, body = best
ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0] for _, x in body)
print("function:", fn)
print("steady loop: 0x%x .. 0x%x, %d instructions (%d bytes), %d packed FP32" % (t, a, len(body), len(body) * 16, fp))
print("opcode histogram:", ", ".join("%s %d" % kv for kv in ops.most_common()))
print("---- first %d instructions" % nshow)
for ad, x in body[:nshow]:
    print("  /*%04x*/ %s" % (ad, x))
print("---- last 12 instructions")
for ad, x in body[-12:]:
    print("  /*%04x*/ %s" % (ad, x))
