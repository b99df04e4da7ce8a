"""Top stall sites of the first kernel in an ncu report's SASS source page (read on the CPU box):
python scripts/ncu_top_stalls.py rep.ncu-rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; data = []; kernels = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        kernels += 1
        if kernels > 1: break
        print(r[1]); continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
i_s = hdr.index("# Samples"); i_src = hdr.index("Source")
tot = sum(int(r[i_s] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[hdr.index(h)] or 0) for r in data) for h in stalls}
print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for idx, r in sorted(enumerate(data), key=lambda t: -int(t[1][i_s] or 0))[:n]:
    st = {h: int(r[hdr.index(h)] or 0) for h in stalls}
    st = sorted(((v, k) for k, v in st.items() if v), reverse=True)[:3]
    print(str(idx).rjust(5), r[i_s].rjust(6), r[i_src][:80].ljust(80), st)
