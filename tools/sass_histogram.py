"""SASS opcode histogram of a kernel's hot loop, from the built object (cuobjdump, no GPU needed).
usage: python tools/sass_histogram.py <object> <mangled-name-substring> [out.txt]
The hot loop = the longest backward-branch region (steady-state trip of the column loop)."""
import collections, re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
sel = [f for f in funcs if pat in f.split("\n")[0]]
assert sel, "no function matches " + pat
text = sel[0]
name = text.split("\n")[0]
ins = []
for line in text.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
# loops = backward branches; the hot loop is the INNERMOST loop (no other loop inside) with the most instructions
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt in addr and addr[tgt] < i:
            loops.append((addr[tgt], i))
inner = [(lo, hi) for lo, hi in loops if not any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in loops)]
lines = ["kernel: " + name, "object: " + obj, "innermost loops of more than 100 instructions (the column loop has a head version, which also holds the",
         "rarely taken segment-restart block, and a steady-state version):", ""]
for lo, hi in sorted(inner):
    body = ins[lo:hi + 1]
    if len(body) <= 100:
        continue
    hist = collections.Counter()
    for _, t in body:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        hist[t.split()[0]] += 1
    tot = sum(hist.values())
    dpx = sum(c for op, c in hist.items() if op.startswith(("VIADDMNMX", "VIMNMX", "VIADD.")))
    lines.append("loop at instructions %d..%d of %d (%d instructions, addresses 0x%x..0x%x)" % (lo, hi, len(ins), len(body), body[0][0], body[-1][0]))
    for op, c in hist.most_common():
        lines.append("    %-28s %6d  %5.1f%%" % (op, c, 100.0 * c / tot))
    lines += ["    recurrence instructions (VIADDMNMX* + VIMNMX* + VIADD.16x2): %d of %d = %.1f%%; everything else (shuffles, LDS, column "
              "words, line load/store, loop control): %d = %.1f%%" % (dpx, tot, 100.0 * dpx / tot, tot - dpx, 100.0 - 100.0 * dpx / tot), ""]
txt = "\n".join(lines) + "\n"
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt)
print(txt)
