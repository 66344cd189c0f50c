"""Instruction mix and hottest non-FFMA basic blocks from an `ncu --page source --csv` export (gzip ok).
    python profiles/sass_mix.py gpurun_out/<tag>_source.csv.gz [nblocks]"""
import collections, csv, gzip, sys
path = sys.argv[1]; nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rd = csv.reader(f); next(rd)
rows = list(rd); hdr = rows[0]; data = rows[1:]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
def opc(s):
    t = s.split()
    if not t: return "?"
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
tot = 0; byop = collections.Counter(); samp = collections.Counter(); allsamp = 0
for r in data:
    if len(r) <= max(iS, iE, iSm): continue
    try: n = int(r[iE])
    except ValueError: continue
    byop[opc(r[iS])] += n; tot += n
    try: samp[opc(r[iS])] += int(r[iSm]); allsamp += int(r[iSm])
    except ValueError: pass
print("total warp instructions %d, samples %d" % (tot, allsamp))
for k, v in byop.most_common(14):
    print("%-10s %12d %5.1f%%   samples %5.1f%%" % (k, v, 100.0 * v / tot, 100.0 * samp[k] / max(allsamp, 1)))
blocks, cur, last = [], [], None
for i, r in enumerate(data):
    if len(r) <= max(iS, iE, iSm): continue
    try: n = int(r[iE])
    except ValueError: n = 0
    if last is not None and n != last and cur:
        blocks.append(cur); cur = []
    cur.append((i, n, r[iS], r[iSm])); last = n
blocks.append(cur)
st = []
for b in blocks:
    n = b[0][1]; ff = sum(1 for x in b if opc(x[2]) == "FFMA")
    sm = sum(int(x[3]) for x in b if x[3].isdigit())
    st.append((sm, n * len(b), n, len(b), ff, b[0][0]))
st.sort(reverse=True)
print("hottest blocks by stall samples: samples, executed, trip count, length, FFMAs, first row")
for s in st[:nb]: print("  %6d %12d %9d %4d %4d  @%d" % s)
