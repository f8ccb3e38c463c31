"""Developer tool: decode the scheduling control bits (stall, write/read scoreboard, wait mask) of a cuobjdump -sass listing.
usage: cuobjdump -sass lib.so | python tools/sass_ctrl.py <first_line> <last_line> [all]   (line numbers of the listing)
Prints loads / texture fetches with the scoreboard they signal and every instruction that waits on a scoreboard."""
import re, sys
lines = sys.stdin.read().split("\n")
lo, hi = int(sys.argv[1]), int(sys.argv[2])
show_all = len(sys.argv) > 3
pat = re.compile(r"/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/")
hexpat = re.compile(r"/\* (0x[0-9a-f]+) \*/")
i = lo - 1
while i < min(hi, len(lines) - 1):
    m = pat.search(lines[i])
    if not m:
        i += 1
        continue
    m2 = hexpat.search(lines[i + 1])
    hiw = int(m2.group(1), 16)
    stall = (hiw >> 41) & 0xf
    wr = (hiw >> 46) & 7
    rd = (hiw >> 49) & 7
    wait = (hiw >> 52) & 0x3f
    txt = m.group(2).strip()
    if show_all or wr != 7 or wait or "DEPBAR" in txt:
        print("%6d %s st=%2d wr=%s rd=%s wait=%s  %s" % (i + 1, m.group(1), stall, wr if wr != 7 else "-", rd if rd != 7 else "-",
              "".join(str(b) for b in range(6) if wait >> b & 1) or "-", txt[:100]))
    i += 2
