"""Static per-kernel resource table of the shipped library (no GPU needed):
cuobjdump --dump-resource-usage libslam_b200.so -> registers, stack (spills), shared memory per kernel,
demangled.  Writes profiles/r01_kernel_resources.md."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(ROOT, "opendlv-logic-cfsd18-sensation-slam_b200", "libslam_b200.so")
out = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
rows = []
src = "?"
name = None
for ln in out.splitlines():
    m = re.match(r"identifier = .*/csrc/(\S+)", ln)
    if m:
        src = m.group(1)
    m = re.match(r"\s*Function (\S+):", ln)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and name:
        rows.append((src, name, *[int(v) for v in m.groups()]))
        name = None
names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    return re.sub(r"\(.*", "", n)   # drop the argument list, keep template arguments


with open(os.path.join(ROOT, "profiles", "r01_kernel_resources.md"), "w") as f:
    f.write("# Static resource usage of every kernel in libslam_b200.so (sm_100a)\n\n"
            "`python profiles/tools/resource_usage.py` (cuobjdump --dump-resource-usage). STACK > 0 means local-memory\n"
            "frames (spills or indexed local arrays); SHARED is static shared memory only (the front kernels take their\n"
            "front as dynamic shared memory on top).\n\n| source | kernel | registers | stack B | static smem B |\n|---|---|---:|---:|---:|\n")
    for (s, _, reg, stack, sh, loc), n in sorted(zip(rows, names), key=lambda t: (t[0][0], t[1])):
        f.write(f"| {s} | `{short(n)}` | {reg} | {stack} | {sh} |\n")
print(len(rows), "kernels;", sum(1 for r in rows if r[3] > 0), "with a stack frame")
