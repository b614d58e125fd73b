"""Summarise an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line:
instructions executed, shared wavefronts (ideal/excess), stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
per = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] == "":  # SASS row
        continue
    d = dict(zip(hdr, r))
    per.append(d)
tot = sum(int(d["Instructions Executed"] or 0) for d in per)
tots = sum(int(d["# Samples"] or 0) for d in per)
print("total warp instr", tot, "samples", tots)
per.sort(key=lambda d: -int(d["Instructions Executed"] or 0))
for d in per[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    src = [v for k, v in zip(hdr, [d[h] for h in hdr]) if k == "Source"]
    print("%5s inst %5.1f%% samp %5.1f%% shWF %9s ideal %9s | %s" % (d["Line No"], 100.0 * int(d["Instructions Executed"] or 0) / tot,
          100.0 * int(d["# Samples"] or 0) / max(tots, 1), d["L1 Wavefronts Shared"], d["L1 Wavefronts Shared Ideal"], list(csv.reader([",".join([])]))and r and d.get("Source","")[:100]))
