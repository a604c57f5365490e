"""Top stall sites of an ncu report's source page (SASS view): python tools/ncu_top_stalls.py <rep> [n]."""
import csv, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
# first kernel only: header row starts with "Address"
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
body = [r for r in rows[start + 1:end] if len(r) == len(hdr)]
si = hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
for r in sorted(body, key=lambda r: -int(r[si] or 0))[:n]:
    why = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print("%6d %5.1f%%  %-70s %s" % (int(r[si] or 0), 100.0 * int(r[si] or 0) / max(tot, 1), r[1].strip()[:70], why))
