"""Decode the scheduling control fields of a SASS loop (cuobjdump -sass text): per-instruction stall count, yield,
write/read barrier, wait mask.  usage: sass_loop_stalls.py <file.sass> <start_addr_hex> <end_addr_hex>
Development aid (r2): the static stall total of the lag loop is the per-warp 'private' time that other warps must cover."""
import re, sys
txt = open(sys.argv[1]).read().splitlines()
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
pat = re.compile(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/")
pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]+) \*/")
i = 0
tot = 0
n = 0
while i < len(txt):
    m = pat.search(txt[i])
    if m and i + 1 < len(txt):
        addr = int(m.group(1), 16)
        m2 = pat2.match(txt[i + 1])
        if m2 and lo <= addr <= hi:
            hw = int(m2.group(1), 16)
            stall = (hw >> 41) & 0xf
            yld = (hw >> 45) & 1
            wbar = (hw >> 46) & 7
            rbar = (hw >> 49) & 7
            wait = (hw >> 52) & 0x3f
            tot += max(stall, 1)
            n += 1
            print("%05x  st=%2d y=%d w=%d r=%d wait=%02x  %s" % (addr, stall, yld, wbar, rbar, wait, m.group(2)[:70]))
        i += 2
    else:
        i += 1
print("instructions %d, sum of stall counts %d" % (n, tot))
