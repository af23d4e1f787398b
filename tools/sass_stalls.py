"""Static issue-time estimate of a SASS range from the scheduling control bits (Volta+ 128-bit encoding):
   stall = bits 105..108, yield = 109, write-barrier index = 110..112, read-barrier = 113..115, wait mask = 116..121.
   usage: python tools/sass_stalls.py <cubin-or-so> <function-substring> [start_hex end_hex]
Prints every instruction with its stall count and barrier use, and the sum of stall counts (issue cycles of a lone warp,
excluding scoreboard waits on variable-latency instructions)."""
import re
import subprocess
import sys

path, fn = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
on = False
rows = []
cur = None
for line in txt.splitlines():
    if "Function :" in line:
        on = fn in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", line)
    if m:
        cur = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]
        rows.append(cur)
        continue
    m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", line)
    if m and cur is not None and cur[3] is None:
        cur[3] = int(m.group(1), 16)
total = 0
for addr, text, w0, w1 in rows:
    if addr < lo or addr > hi or w1 is None:
        continue
    ctrl = w1 >> 41  # bits 105.. of the 128-bit word
    stall = ctrl & 0xf
    yld = (ctrl >> 4) & 1
    wb = (ctrl >> 5) & 7
    rb = (ctrl >> 8) & 7
    wait = (ctrl >> 11) & 0x3f
    total += stall
    print(f"{addr:06x} stall {stall:2d} {'Y' if yld else ' '} wr {wb if wb != 7 else '-'} rd {rb if rb != 7 else '-'} wait {wait:06b}  {text}")
print("sum of stall counts:", total)
