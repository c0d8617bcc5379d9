"""Summaries of one `ncu --set full --import-source on` report for profiles/: raw page (csv), details page (text, first
200 lines) and the source page aggregated into stall-reason shares and the opcode mix.
Usage: python scripts/ncu_summarize.py report.ncu-rep out_prefix"""
import collections, csv, io, subprocess, sys
rep, pre = sys.argv[1], sys.argv[2]
run = lambda *a: subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout
open(pre + "_raw.csv", "w").write(run("--page", "raw", "--csv"))
open(pre + "_details.txt", "w").write("\n".join(run("--page", "details").splitlines()[:200]) + "\n")
src = run("--page", "source", "--csv", "--print-source", "sass")
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Address")
name = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "?")
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
col = {h: i for i, h in enumerate(hdr)}
stalls = collections.Counter()
for h, i in col.items():
    if h.startswith("stall_") and "Not Issued" not in h:
        stalls[h] = sum(int(float(r[i] or 0)) for r in data)
ops = collections.Counter()
ie = col.get("Instructions Executed")
for r in data:
    op = r[col["Source"]].split()
    op = [t for t in op if not t.startswith("@")]
    if op and ie is not None:
        ops[op[0]] += int(float(r[ie] or 0))
with open(pre + "_stalls.txt", "w") as f:
    f.write("%s (ncu --set full, source page aggregated)\n" % name)
    tot = sum(stalls.values()) or 1
    for k, v in stalls.most_common():
        f.write("%-28s %10d  %.3f\n" % (k, v, v / tot))
    f.write("\nwarp-instructions executed by opcode (top 14)\n")
    tot = sum(ops.values()) or 1
    for k, v in ops.most_common(14):
        f.write("%-24s %14d  %.4f\n" % (k, v, v / tot))
print(open(pre + "_stalls.txt").read())
