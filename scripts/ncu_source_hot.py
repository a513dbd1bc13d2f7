"""Reads `ncu -i <rep> --page source --csv --print-source sass` (stdin or file) and prints, per kernel, the stall-reason mix and the
hottest SASS instructions with their dominant stall reasons.  Usage: ncu -i x.ncu-rep --page source --csv --print-source sass | python scripts/ncu_source_hot.py [top_n]"""
import csv, sys, collections
top_n = int(sys.argv[1]) if len(sys.argv) > 1 else 18
rows = list(csv.reader(sys.stdin))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1][:70]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(rows[j])
            j += 1
        col = {h: k for k, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") or h.lower().startswith("warp stall")]
        # stall reason columns are named like 'stall_long_sb' in newer ncu; fall back to everything after a marker
        reason_cols = [h for h in hdr if h.startswith("stall_")]
        samp = col.get("# Samples")
        tot = sum(int(r[samp] or 0) for r in body)
        inst = col.get("Instructions Executed")
        tot_inst = sum(int(r[inst] or 0) for r in body)
        print("=====", name, "| %d SASS instructions, %d samples, %d warp instructions executed" % (len(body), tot, tot_inst))
        mix = collections.Counter()
        for r in body:
            for h in reason_cols:
                v = r[col[h]]
                if v and v != "0":
                    mix[h] += int(v)
        s = sum(mix.values()) or 1
        print("  stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in mix.most_common(10)))
        body.sort(key=lambda r: -int(r[samp] or 0))
        for r in body[:top_n]:
            rs = sorted(((int(r[col[h]] or 0), h[6:]) for h in reason_cols), reverse=True)[:2]
            print("  %5.1f%%  %-58s x%-7s %s" % (100.0 * int(r[samp] or 0) / max(tot, 1), r[col["Source"]].strip()[:58], r[inst], [(n, c) for c, n in rs if c]))
        i = j
    else:
        i += 1
