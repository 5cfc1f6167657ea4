#!/usr/bin/env python3
"""Dynamic opcode counts of poseidon_permute_kernel from its SASS listing, using the known loop trip
counts (S-box loop x3 per full round, full rounds x8, initial-matrix loop x11, partial rounds x22).
Usage: cuobjdump -sass libsvb200.so | tools/sass_dyn.py   (prints issue-slot and fmaheavy estimates
with the B200 rates measured by tools/microbench/pipes.cu: IMAD.WIDE = 2 issue slots, IMAD.HI = 4
fmaheavy cycles, other IMAD = 2 fmaheavy cycles)."""
import re, sys, collections
lines = []
on = False
for l in sys.stdin:
    if "Function :" in l:
        on = "poseidon_permute_kernelILi0E" in l
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
addr_to_idx = {a: i for i, (a, _) in enumerate(lines)}
# backward branches define loops
loops = []
for i, (a, t) in enumerate(lines):
    m = re.search(r"BRA(?:\.U)?\s+(?:U?P\d,\s*|!U?P\d,\s*|UP\d,\s*)?(0x[0-9a-f]+)", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        if tgt in addr_to_idx and addr_to_idx[tgt] <= i and tgt != a:
            loops.append((addr_to_idx[tgt], i))
loops.sort()
# expected nesting: outer (x8) contains sbox (x3); then init (x11), partial (x22) inside the f==3 branch
trip = {}
if len(loops) == 1:
    # naive-round form (SVB_PARTIAL_NAIVE): ONE loop over the 30 rounds; a forward branch inside it skips the
    # S-boxes of lanes 1..11 in the 22 partial rounds, so that region runs 8 times and the rest 30 times
    lo, hi_ = loops[0]
    mult = [1.0] * len(lines)
    for i in range(lo, hi_ + 1):
        mult[i] = 30.0
    for i in range(lo, hi_):
        t = lines[i][1]
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt in addr_to_idx and i < addr_to_idx[tgt] <= hi_:
                for j in range(i + 1, addr_to_idx[tgt]):
                    mult[j] = 8.0
                break
    cnt = collections.Counter()
    for (a, t), mm in zip(lines, mult):
        op = t.split()[0] if not t.startswith("@") else t.split()[1]
        cnt[op] += mm
    tot = sum(cnt.values())
    wide = sum(v for k, v in cnt.items() if k.startswith("IMAD.WIDE"))
    hi = sum(v for k, v in cnt.items() if k.startswith("IMAD.HI"))
    imad_other = sum(v for k, v in cnt.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI"))
    print(f"dynamic instructions per permutation: {tot:.0f}")
    for k, v in cnt.most_common(16):
        print(f"  {k:22s} {v:8.0f}")
    print(f"instructions: {tot:.0f};  fmaheavy cycles (WIDE,HI=4, other IMAD=2): {4 * (wide + hi) + 2 * imad_other:.0f}  (wide {wide:.0f}, hi {hi:.0f}, other IMAD {imad_other:.0f})")
    sys.exit(0)
def report(mult):
    cnt = collections.Counter()
    for (a, t), mm in zip(lines, mult):
        op = t.split()[0] if not t.startswith("@") else t.split()[1]
        cnt[op] += mm
    tot = sum(cnt.values())
    wide = sum(v for k, v in cnt.items() if k.startswith("IMAD.WIDE"))
    hi = sum(v for k, v in cnt.items() if k.startswith("IMAD.HI"))
    imad_other = sum(v for k, v in cnt.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI"))
    fp64 = sum(v for k, v in cnt.items() if k.startswith("DFMA") or k.startswith("DADD") or k.startswith("DMUL"))
    print(f"dynamic instructions per permutation: {tot:.0f}")
    for k, v in cnt.most_common(16):
        print(f"  {k:22s} {v:8.0f}")
    print(f"instructions: {tot:.0f};  IMAD.WIDE {wide:.0f} (x4.24 = {4.24 * wide:.0f} fmaheavy cycles), FP64 {fp64:.0f} (x2.18 = {2.18 * fp64:.0f}), other IMAD {imad_other:.0f}")
    sys.exit(0)

if len(loops) == 3 and sorted(loops)[1][1] - sorted(loops)[1][0] > 500:
    # naive-round form with double layers (SVB_PARTIAL_NAIVE): phase loop (x2) containing the full-round loop
    # (x4 per phase) and, in phase 0 only, the double-layer loop (x11)
    outer = max(loops, key=lambda x: x[1] - x[0])
    inner = sorted(l for l in loops if l != outer)
    mult = [1.0] * len(lines)
    for i in range(outer[0], outer[1] + 1):
        mult[i] = 2.0
    for i in range(inner[0][0], inner[0][1] + 1):
        mult[i] = 8.0
    for i in range(inner[1][0], inner[1][1] + 1):
        mult[i] = 11.0
    report(mult)
if len(loops) == 3:
    # fully unrolled S-box layer: outer full-round loop (x8) containing the initial-matrix (x11) and partial (x22) loops
    outer = max(loops, key=lambda x: x[1] - x[0])
    inner = sorted(l for l in loops if l != outer)
    mult = [1.0] * len(lines)
    for i in range(outer[0], outer[1] + 1):
        mult[i] = 8.0
    # the f == 3 block (forward branch over it) runs once: everything from the branch before inner[0] to the end of inner[1]
    start = inner[0][0]
    for i in range(inner[0][0] - 1, outer[0], -1):
        if "BRA" in lines[i][1]:
            start = i + 1
            break
    m = None
    for i in range(start - 1, start):
        m = re.search(r"(0x[0-9a-f]+)", lines[i][1].split("BRA")[1]) if "BRA" in lines[i][1] else None
    end = addr_to_idx[int(m.group(1), 16)] if m and int(m.group(1), 16) in addr_to_idx else inner[1][1] + 1
    for i in range(start, end):
        mult[i] = 1.0
    for i in range(inner[0][0], inner[0][1] + 1):
        mult[i] = 11.0
    for i in range(inner[1][0], inner[1][1] + 1):
        mult[i] = 22.0
    cnt = collections.Counter()
    for (a, t), mm in zip(lines, mult):
        op = t.split()[0] if not t.startswith("@") else t.split()[1]
        cnt[op] += mm
    tot = sum(cnt.values())
    wide = sum(v for k, v in cnt.items() if k.startswith("IMAD.WIDE"))
    hi = sum(v for k, v in cnt.items() if k.startswith("IMAD.HI"))
    imad_other = sum(v for k, v in cnt.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI"))
    print(f"dynamic instructions per permutation: {tot:.0f}")
    for k, v in cnt.most_common(16):
        print(f"  {k:22s} {v:8.0f}")
    print(f"instructions: {tot:.0f};  fmaheavy cycles (WIDE,HI=4, other IMAD=2): {4 * (wide + hi) + 2 * imad_other:.0f}  (wide {wide:.0f}, hi {hi:.0f}, other IMAD {imad_other:.0f})")
    sys.exit(0)
if len(loops) == 4:
    loops_sorted = sorted(loops, key=lambda x: x[1] - x[0])
    # identify: largest = outer
    outer = max(loops, key=lambda x: x[1] - x[0])
    inner = [l for l in loops if l != outer]
    inner.sort()
    trip[inner[0]] = 3; trip[inner[1]] = 11; trip[inner[2]] = 22; trip[outer] = 8
else:
    print("unexpected loop structure", loops); sys.exit(1)
mult = [1.0] * len(lines)
outer = max(loops, key=lambda x: x[1] - x[0])
for (b, e), t in trip.items():
    for i in range(b, e + 1):
        if (b, e) == outer:
            mult[i] *= 8
        elif (b, e) == sorted([l for l in loops if l != outer])[0]:
            mult[i] *= 3
# init/partial loops execute once (inside f==3): undo the outer x8 for everything between the forward branch and its target
inner = sorted([l for l in loops if l != outer])
fwd_start = inner[0][1] + 1
# find forward branch after MDS
for i in range(inner[0][1] + 1, inner[1][0]):
    if "BRA" in lines[i][1]:
        m = re.search(r"(0x[0-9a-f]+)", lines[i][1].split("BRA")[1])
        tgt = addr_to_idx[int(m.group(1), 16)]
        for j in range(i + 1, tgt):
            mult[j] /= 8
        break
for (b, e) in inner[1:]:
    for i in range(b, e + 1):
        mult[i] *= trip[(b, e)]
cnt = collections.Counter()
for (a, t), m in zip(lines, mult):
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    cnt[op] += m
tot = sum(cnt.values())
wide = sum(v for k, v in cnt.items() if k.startswith("IMAD.WIDE"))
hi = sum(v for k, v in cnt.items() if k.startswith("IMAD.HI"))
imad_other = sum(v for k, v in cnt.items() if k.startswith("IMAD") and not k.startswith("IMAD.WIDE") and not k.startswith("IMAD.HI"))
print(f"dynamic instructions per permutation: {tot:.0f}")
for k, v in cnt.most_common(16):
    print(f"  {k:22s} {v:8.0f}")
print(f"instructions: {tot:.0f};  fmaheavy cycles (WIDE,HI=4, other IMAD=2): {4 * (wide + hi) + 2 * imad_other:.0f}  (wide {wide:.0f}, hi {hi:.0f}, other IMAD {imad_other:.0f})")
