"""Per-region instruction/stall breakdown of an .ncu-rep SASS page: python tools/sass_hot.py rep pairs [chunk]"""
import collections, csv, subprocess, sys
rep, pairs = sys.argv[1], float(sys.argv[2])
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data = rows[1], rows[2:]
ia, it, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); totT = sum(int(r[it]) for r in data); totS = sum(int(r[isamp]) for r in data)
print(f"warp inst {tot:.3e} thread inst {totT:.3e} avg threads {totT/tot:.1f} thread-inst/pair {totT/pairs:.0f} warp-inst/pair-slot {tot*32/pairs:.0f} samples {totS}")
def opc(r):
    s = r[isrc].split()
    return (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
op = collections.Counter(); ops = collections.Counter()
for r in data:
    op[opc(r)] += int(r[ia]); ops[opc(r)] += int(r[isamp])
print("  ".join(f"{o}:{c/tot*100:.1f}%/{ops[o]/totS*100:.1f}%s" for o, c in op.most_common(18)))
for i in range(0, len(data), chunk):
    blk = data[i:i + chunk]
    a = sum(int(r[ia]) for r in blk); t = sum(int(r[it]) for r in blk); s = sum(int(r[isamp]) for r in blk)
    if a / tot < 0.012 and s / totS < 0.012: continue
    o2 = collections.Counter(opc(r) for r in blk)
    print(f"{i:5d} inst {a/tot*100:5.1f}% samples {s/totS*100:5.1f}% avgthr {t/max(a,1):5.1f}  {dict(o2.most_common(6))}")
