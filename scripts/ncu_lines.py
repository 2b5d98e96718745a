"""Top CUDA source lines by warp-stall samples for one kernel of an .ncu-rep (needs -lineinfo).
usage: python scripts/ncu_lines.py <rep> <kernel-name-regex> [launch-skip] [top]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, data, hdr = None, [], None
for r in rows:
    if not r: continue
    if r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "" : continue
    try:
        s = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    stalls = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v) > 0}
    data.append((s, fname, r[0], r[1].strip()[:90], stalls))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, f, ln, src, st in sorted(data, key=lambda d: -d[0])[:top]:
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*s/max(tot,1):5.1f}% {f}:{ln:>4} {src}  {main}")
