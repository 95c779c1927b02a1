"""Warp-stall samples of one kernel launch in an `ncu --set full --import-source on` report, summed per
CUDA source line: the SASS page of the report is joined with `nvdisasm -g` line info of the cubin.
usage: sass_hot_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING LAUNCH_SKIP SOURCE.cu [TOP]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kname, skip, srcfile = sys.argv[1:6]
top = int(sys.argv[6]) if len(sys.argv) > 6 else 25
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null" % (tmp, os.path.abspath(obj)), shell=True, check=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run("nvdisasm -g -c %s/%s" % (tmp, cubin), shell=True, capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l][0]
cur, off2line = None, {}
for l in dis[start + 1:]:
    if (l.startswith("\t.section") or l.startswith("//-----")) and off2line:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = int(m.group(2)) if m.group(1).endswith(os.path.basename(srcfile)) else -1
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
out = subprocess.run("ncu -i %s --page source --csv --print-source sass --launch-skip %s --launch-count 1" % (rep, skip),
                     shell=True, capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if r and r[0].startswith("0x")]
base = int(data[0][0], 16)
samp, inst = collections.Counter(), collections.Counter()
stall = collections.defaultdict(collections.Counter)
stall_cols = [h for h in hdr if h.startswith("stall_")]
for r in data:
    ln = off2line.get(int(r[0], 16) - base, -2)
    samp[ln] += int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    inst[ln] += int(r[ix["Instructions Executed"]] or 0)
    for h in stall_cols:
        v = int(r[ix[h]] or 0)
        if v:
            stall[ln][h] += v
tot = sum(samp.values())
src = open(srcfile).read().splitlines()
print("kernel", rows[0][1][:80] if rows and len(rows[0]) > 1 else kname, "| total samples", tot)
for ln, s in samp.most_common(top):
    why = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in stall[ln].most_common(3))
    print("%5.1f%% line %4d  warp-inst %7d | %-72s | %s" % (100.0 * s / max(tot, 1), ln, inst[ln],
                                                          src[ln - 1].strip()[:72] if ln > 0 else "?", why))
