"""SASS instruction count / size of every step kernel in a built library (instruction-cache budget check).
Usage: python tools/sass_size.py [lib.so]"""
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "solorl_b200/libsolo_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, counts = None, {}
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1)
        counts[name] = 0
    elif name and re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", ln):
        counts[name] += 1
for n, c in sorted(counts.items(), key=lambda kv: kv[1]):
    if "step_kernel" in n or "actuator" in n:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
        print(f"{c:7d} instr {c * 16 / 1024:7.1f} KB  {d}")
