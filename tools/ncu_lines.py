#!/usr/bin/env python
"""Per-source-line instruction / stall-sample profile from an .ncu-rep (built with -lineinfo).

ncu's CSV source page is SASS-level; this joins it (by instruction order) with `nvdisasm -g` line markers of the
cubin extracted from the built library, and prints the hottest source lines.
usage: tools/ncu_lines.py <report.ncu-rep> <kernel substring> [library.so] [top N]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(so, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    out = []
    for cub in sorted(glob.glob(os.path.join(d, "*.cubin"))):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
        if kernel not in txt:
            continue
        in_k, cur = False, ("?", 0)
        for line in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                in_k = kernel in m.group(1)
                continue
            if not in_k:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
                out.append(cur)
        if out:
            break
    return out


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    so = os.path.abspath(sys.argv[3]) if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sylt2d_b200", "libsylt2d_b200.so")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csvtxt.splitlines()))
    # sections start with a "Kernel Name" row followed by a header row
    secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    sec = next(i for i in secs if kernel in rows[i][1])
    h = rows[sec + 1]
    end = next((j for j in secs if j > sec), len(rows))
    ci, ss = h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [k for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    body = [r for r in rows[sec + 2:end] if len(r) > ci]
    lines = sass_lines(so, kernel)
    if len(lines) != len(body):
        print(f"warning: {len(body)} profiled instructions vs {len(lines)} disassembled; joining by order up to the shorter", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    ti = ts = 0.0
    for r, ln in zip(body, lines):
        try:
            v, s = float(r[ci]), float(r[ss])
        except ValueError:
            continue
        a = agg[ln]
        a[0] += v
        a[1] += s
        for k in stall_cols:
            try:
                a[2][h[k]] += float(r[k])
            except ValueError:
                pass
        ti += v
        ts += s
    print(f"kernel {rows[sec][1][:60]}  warp-instructions {ti:.0f}  samples {ts:.0f}")
    src_cache = {}
    for (f, n), (v, s, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        if f not in src_cache:
            p = glob.glob(os.path.join(os.path.dirname(so), "csrc", f)) + glob.glob(os.path.join(os.path.dirname(so), "**", f), recursive=True)
            src_cache[f] = open(p[0]).read().splitlines() if p else []
        text = src_cache[f][n - 1].strip()[:90] if 0 < n <= len(src_cache[f]) else ""
        top_st = ",".join(f"{k[6:]}={int(c)}" for k, c in st.most_common(2))
        print(f"{s / max(ts, 1):6.2%} samp {v / max(ti, 1):6.2%} inst  {f}:{n:<4d} [{top_st}] {text}")


if __name__ == "__main__":
    main()
