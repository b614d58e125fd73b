"""Dynamic instruction mix of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda --kernel-name regex:K`:
executed warp instructions per SASS opcode, and per CUDA source line (file:line) when asked.
usage: python tools/ncu_opmix.py src.csv [pixels] [--lines N]"""
import csv, sys, collections, re
path = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 0
rows = list(csv.reader(open(path, errors="replace")))
ops = collections.Counter(); lines = collections.Counter(); stall = collections.Counter(); text = {}
hdr = None; cur_file = ""
seen_addr = set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); continue
    if hdr is None or len(r) != len(hdr): continue
    toi = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
    n = toi(r[ie]); s = toi(r[sm])
    if r[0] == "":            # SASS row: columns Address, Source (second 'Source')
        addr = r[2]
        if addr in seen_addr: continue     # the same SASS row is listed under every file section it maps to
        seen_addr.add(addr)
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        op = m.group(1) if m else r[3][:12]
        ops[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG")) and "." in op else "")] += n
        stall[op.split(".")[0]] += s
    else:
        key = "%s:%s" % (cur_file, r[0]); lines[key] += n; text[key] = r[1].strip()[:110]
tot = sum(ops.values())
print("total executed warp instructions %d%s" % (tot, "  = %.2f per px (%.1f thread-instr/px)" % (tot / px, 32 * tot / px) if px else ""))
for op, n in ops.most_common(28):
    print("  %-12s %6.2f%%  %s  stall-samples %5.1f%%" % (op, 100.0 * n / tot, ("%.3f/px" % (n / px)) if px else "", 100.0 * stall[op] / max(1, sum(stall.values()))))
if nlines:
    lt = sum(lines.values())
    for k, n in lines.most_common(nlines):
        print("  %5.1f%%  %-28s %s" % (100.0 * n / lt, k, text[k]))
