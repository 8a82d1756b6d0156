"""Static SASS instruction count of one kernel, attributed to the innermost source function (via nvdisasm -gi
inline chains) and to the top-level call site in the kernel body.  Usage: python tools/sass_by_function.py lib.so mangled-hint"""
import os, re, subprocess, sys, tempfile
from collections import Counter
lib, hint = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
re_sec = re.compile(r"^\s*\.section\s+\.text\.(\S+?),")
re_file = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
re_ins = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*);")
active, inner, outer = False, Counter(), Counter()
cur_inner = cur_outer = None
pending = []
for ln in txt.splitlines():
    m = re_sec.match(ln)
    if m:
        active = hint in m.group(1); pending = []; continue
    if not active:
        continue
    m = re_file.search(ln)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re_ins.match(ln)
    if m:
        if pending:
            cur_inner, cur_outer = pending[0], pending[-1]
            pending = []
        inner[cur_inner] += 1; outer[cur_outer] += 1
src = {}
def text(f, l):
    if f not in src:
        try: src[f] = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "solorl_b200", "csrc", f)).read().splitlines()
        except Exception: src[f] = []
    return src[f][l - 1].strip()[:80] if 0 < l <= len(src[f]) else ""
print("total", sum(outer.values()))
print("== by outermost call site")
for (f, l), c in outer.most_common(45):
    print(f"{c:6d} {f}:{l}  {text(f, l)}")
