"""Read an ncu report's source page (SASS) and aggregate the warp-stall samples by code region (regions end at
TMEM / TMA / barrier instructions), plus the hottest instructions: where do the epilogue warps wait?
usage: ncu_regions.py report.ncu-rep [launch index]"""
import csv, re, subprocess, sys
rep, li = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
data = data[:len(data) // 2] if len(data) > 2 and data[0][isrc] == data[len(data) // 2][isrc] else data
tot = sum(int(r[isamp]) for r in data)
print("instructions", len(data), "samples", tot)
marks = re.compile(r"LDTM|STTM|UTMASTG|UTMALDG|BAR\.SYNC|UCGABAR|UTCHMMA|SYNCS\.PHASECHK|DEPBAR|USETMAXREG|ACQBULK|EXIT|MEMBAR|FENCE")
acc, start = 0, 0
for i, r in enumerate(data):
    acc += int(r[isamp])
    if marks.search(r[isrc]):
        if acc >= max(3, tot // 400):
            print("%5d..%5d %6d  | %s (x%s)" % (start, i, acc, r[isrc].strip()[:70], r[iex]))
        acc, start = 0, i + 1
top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:25]
print("hottest:")
for i in sorted(top):
    r = data[i]
    st = {h[6:]: int(r[j]) for j, h in enumerate(hdr) if h.startswith("stall_") and "(Not" not in h and j < len(r) and r[j] not in ("", "0")}
    print(i, r[isamp], r[isrc].strip()[:60], dict(sorted(st.items(), key=lambda kv: -kv[1])[:2]))
