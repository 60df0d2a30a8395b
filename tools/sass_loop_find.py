"""Find the lag loop(s) of a forward kernel in the built library and print them with their decoded scheduling fields.
usage: sass_loop_find.py <lib.so> <substring of the mangled kernel name> [number of DMMA in the loop body, default 8]
(r2 development aid; the committed excerpt under profiles/ was produced by it.)"""
import os, re, subprocess, sys, tempfile
lib, key = sys.argv[1], sys.argv[2]
want = int(sys.argv[3]) if len(sys.argv) > 3 else 8
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout.splitlines()
out, on = [], False
for l in sass:
    if "Function :" in l:
        if on:
            break
        on = key in l
    if on:
        out.append(l)
assert out, "kernel not found"
print(out[0].strip())
tmp = tempfile.NamedTemporaryFile("w", suffix=".sass", delete=False)
tmp.write("\n".join(out)); tmp.close()
ins = []
for l in out:
    m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
here = os.path.dirname(os.path.abspath(__file__))
for a, t in ins:
    m = re.search(r"BRA(\.U)?\s+(!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(3), 16) < a:
        tgt = int(m.group(3), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        if sum(1 for x in body if "DMMA" in x[1]) == want:
            print("\nloop 0x%x .. 0x%x" % (tgt, a))
            print(subprocess.run([sys.executable, os.path.join(here, "sass_loop_stalls.py"), tmp.name, "%x" % tgt, "%x" % a],
                                 capture_output=True, text=True).stdout)
os.unlink(tmp.name)
