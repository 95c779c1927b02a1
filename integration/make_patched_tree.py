#!/usr/bin/env python3
"""integration/make_patched_tree.py -- puts the B200 back end into a SCRATCH copy of the reference tree.

    make_patched_tree.py REFERENCE_DIR OUT_DIR [--write-patch integration/slam_b200.patch]

Copies REFERENCE_DIR/src/{slam.hpp,slam.cpp} to OUT_DIR/src, swaps in the sections of
integration/slam_b200_bodies.cpp.in (header include, header member, glue, method bodies) and, with --write-patch,
writes the unified diff a maintainer applies (`patch -p1` in the reference checkout).  The edits are found by
NAME (the `Slam::<method>(` definition and its brace-matched body; the g2o include lines; the m_optimizer member), so the
script carries no reference text; the patch it writes does, like every patch.  Nothing is written outside OUT_DIR
and the patch path.  The public section of class Slam (slam.hpp:52-62) is asserted to be unchanged."""
import difflib
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def sections(path):
    out, cur = {}, None
    for line in open(path).read().splitlines(keepends=True):
        m = re.match(r"//@@ (.+?)\s*$", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(line)
    return {k: "".join(v).strip("\n") + "\n" for k, v in out.items()}


def find_body(src, qualified):
    """(start, end) of the brace-matched body of the definition `... qualified(...) {`; skips strings and comments."""
    m = re.search(r"\b" + re.escape(qualified) + r"\s*\(", src)
    if not m:
        raise SystemExit("definition of %s not found" % qualified)
    i = src.index("{", m.end())
    depth, j, n = 0, i, len(src)
    while j < n:
        c = src[j]
        if src.startswith("//", j):
            j = src.index("\n", j)
            continue
        if src.startswith("/*", j):
            j = src.index("*/", j) + 2
            continue
        if c == '"':
            j += 1
            while src[j] != '"':
                j += 2 if src[j] == "\\" else 1
        elif c == "'":
            j += 1
            while src[j] != "'":
                j += 2 if src[j] == "\\" else 1
        elif c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i, j + 1
        j += 1
    raise SystemExit("unbalanced braces in %s" % qualified)


def public_section(hdr):
    a = hdr.index("public:")
    b = hdr.index("private:", a)
    return hdr[a:b]


def patch_header(hdr, sec):
    lines = hdr.splitlines(keepends=True)
    out, placed = [], False
    for l in lines:
        if re.match(r'\s*#include\s+"g2o/', l):
            if not placed:
                out.append(sec["header-include"])
                placed = True
            continue
        if re.match(r"\s*g2o::SparseOptimizer\s+m_optimizer\s*;", l):
            out.append(sec["header-member"])
            continue
        if re.match(r"\s*bool\s+m_loopClosingComplete\s*=\s*false\s*;", l):
            # read without a lock by drawCurrentPose() on the viewer thread while a frame thread sets it
            out.append(re.sub(r"bool\s+m_loopClosingComplete\s*=\s*false\s*;", "std::atomic<bool> m_loopClosingComplete{false};", l))
            continue
        out.append(l)
    new = "".join(out)
    if not placed or "m_ctx" not in new or "std::atomic<bool> m_loopClosingComplete" not in new:
        raise SystemExit("slam.hpp does not look like the reference header")
    if public_section(new) != public_section(hdr):
        raise SystemExit("the public section of class Slam changed")
    return new


def patch_source(src, sec):
    new = re.sub(r",\s*m_optimizer\(\)", ", m_ctx()", src, count=1)
    if new == src:
        raise SystemExit("constructor initialiser of m_optimizer not found")
    m = re.search(r'#include\s+"WGS84toCartesian.hpp"\s*\n', new)
    new = new[:m.end()] + "\n" + sec["glue"] + new[m.end():]
    for key, body in sec.items():
        if not key.startswith("body "):
            continue
        a, b = find_body(new, key[5:])
        new = new[:a] + body.rstrip("\n") + new[b:]
    if "g2o::" in new or re.search(r"\bm_optimizer\b", new):
        raise SystemExit("g2o is still referenced after the swap")
    return new


def main():
    if len(sys.argv) < 3:
        raise SystemExit(__doc__)
    ref, out = sys.argv[1], sys.argv[2]
    patch_path = sys.argv[4] if len(sys.argv) > 4 and sys.argv[3] == "--write-patch" else None
    sec = sections(os.path.join(HERE, "slam_b200_bodies.cpp.in"))
    os.makedirs(os.path.join(out, "src"), exist_ok=True)
    diff = []
    for name, fn in (("slam.hpp", patch_header), ("slam.cpp", patch_source)):
        old = open(os.path.join(ref, "src", name)).read()
        new = fn(old, sec)
        open(os.path.join(out, "src", name), "w").write(new)
        diff += difflib.unified_diff(old.splitlines(keepends=True), new.splitlines(keepends=True),
                                     "a/src/" + name, "b/src/" + name, n=2)
    if patch_path:
        open(patch_path, "w").write("".join(diff))
    print("patched tree in %s (%d diff lines)" % (out, len(diff)))


if __name__ == "__main__":
    main()
