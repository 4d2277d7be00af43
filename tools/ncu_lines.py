"""Hot source lines of a kernel in an .ncu-rep (needs -lineinfo and --import-source on):
python tools/ncu_lines.py file.ncu-rep [top_n]  ->  per line: share of stall samples, of executed instructions, top stall reasons"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines = []; fname = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue           # SASS rows have an empty line number
    lines.append((fname, r))
def I(x):
    try: return int(x)
    except ValueError: return 0
isamp = hdr.index("# Samples"); iinst = hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(I(r[isamp]) for _, r in lines) or 1; toti = sum(I(r[iinst]) for _, r in lines) or 1
print(f"samples {tot}  warp instructions {toti:.3e}")
for f, r in sorted(lines, key=lambda fr: -I(fr[1][isamp]))[:topn]:
    st = sorted(((I(r[i]), h[6:]) for i, h in stall), reverse=True)[:3]
    ss = " ".join(f"{h}:{c * 100 // max(I(r[isamp]), 1)}%" for c, h in st if c)
    print(f"{f}:{r[0]:>4} samp {I(r[isamp]) / tot * 100:5.1f}% inst {I(r[iinst]) / toti * 100:5.1f}%  [{ss}]  {r[1].strip()[:110]}")
