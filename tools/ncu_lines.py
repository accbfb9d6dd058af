"""Executed warp instructions and stall samples per SOURCE LINE of one kernel in an .ncu-rep (needs -lineinfo and
--import-source on):   python tools/ncu_lines.py file.ncu-rep kernel_regex [top_n]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
cur, agg, hdr = None, [], None
for r in csv.reader(raw.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0].isdigit():
        i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[i_inst]), int(r[i_samp])))
        except ValueError:
            pass
tot, tsamp = sum(a[3] for a in agg), sum(a[4] for a in agg)
print(f"{kern}: {tot} warp instructions, {tsamp} stall samples (first launch of the report that matches)")
for a in sorted(agg, key=lambda a: -a[3])[:top]:
    print(f"{a[0]:20s} {a[1]:5d} {a[3]:>10d} {100 * a[3] / max(tot, 1):5.1f}%  samples {100 * a[4] / max(tsamp, 1):5.1f}% | {a[2]}")
