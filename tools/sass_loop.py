"""Diagnostic: instruction mix of the per-chunk loop of a scan kernel, from the SASS of the built library.
    python tools/sass_loop.py k_accumulate_v4ILb1ELb0ELb1 [--list]
"""
import collections, re, subprocess, sys, os
lib = os.environ.get("A3R_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "align3r_b200", "lib", "liba3r_ga.so")
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
ins, on = [], False
for l in out:
    if "Function :" in l:
        on = pat in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for a, t in ins:
    m = re.search(r"BRA(?:\.U)?(?:\.ANY)? .*?0x([0-9a-f]+)", t) or re.search(r"BRA 0x([0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        if tgt < a:
            loops.append((tgt, a))
def body(lo, hi): return [(a, t) for a, t in ins if lo <= a <= hi]
cands = [(hi - lo, lo, hi) for lo, hi in loops if any("MUFU.RSQ" in t or "MUFU.RCP" in t for _, t in body(lo, hi)) and any(("UBLKCP" in t or "LDGSTS" in t) for _, t in body(lo, hi))]
cands.sort()
_, lo, hi = cands[0]
b = body(lo, hi)
op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
c = collections.Counter(op(t) for _, t in b)
print(f"chunk loop {lo:#x}..{hi:#x}: {len(b)} instructions")
print("  ".join(f"{k}:{v}" for k, v in c.most_common()))
math = {"FFMA", "FADD", "FMUL", "FSEL", "FSETP", "FMNMX", "MUFU"}
print("math", sum(v for k, v in c.items() if k in math), "other", sum(v for k, v in c.items() if k not in math))
if "--list" in sys.argv:
    run = 0
    for a, t in b:
        if op(t) in math:
            run += 1
            continue
        if run:
            print(f"      ... {run} math")
            run = 0
        print(f"{a:#06x} {t[:110]}")
    if run:
        print(f"      ... {run} math")
