"""Attribute an ncu capture of step_kernel to the phases of the substep.

ncu's source page gives executed-instruction counts and warp-stall samples per SASS address;
`nvdisasm -gi` of the same cubin gives, per SASS address, the chain of inlined call sites.  Joining
the two by address offset and keeping, for every instruction, the call-site line inside
group_substep / step_kernel (solo_kernels.cu) yields a per-phase table: which call of the substep
executes how many warp instructions and where the stall samples fall.

Usage: python tools/ncu_phases.py REPORT.ncu-rep LIB.so [kernel-regex] [launch-index]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

STALLS = ["stall_selected", "stall_wait", "stall_short_sb", "stall_no_inst", "stall_long_sb", "stall_branch_resolving",
          "stall_dispatch", "stall_math", "stall_not_selected", "stall_lg", "stall_mio", "stall_barrier", "stall_misc"]


def sass_page(rep, kernel, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          f"regex:{kernel}", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "?")
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    return name, hdr, body


def line_chains(lib, mangled_hint):
    """address offset -> list of (file, line) from innermost to outermost, for the function whose
    section name contains mangled_hint."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
    chains, cur, pending, active = {}, [], [], False
    re_sec = re.compile(r"^\s*\.section\s+\.text\.(\S+?),")
    re_file = re.compile(r'//## File "([^"]+)", line (\d+)')
    re_ins = re.compile(r"^\s+/\*([0-9a-f]+)\*/\s+(.*);")
    for ln in txt.splitlines():
        m = re_sec.match(ln)
        if m:
            active = mangled_hint in m.group(1)
            cur, pending = [], []
            continue
        if not active:
            continue
        m = re_file.search(ln)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re_ins.match(ln)
        if m:
            if pending:
                cur, pending = pending, []
            chains[int(m.group(1), 16)] = cur
    return chains


SCOPES = [("solo_wide.cuh", "void wide_link("), ("solo_wide.cuh", "void wide_row("),
          ("solo_wide.cuh", "void wide_leg_substep("), ("solo_wide.cuh", "void wide_helper_substep("),
          ("solo_kernels.cu", "void contact_solve("), ("solo_kernels.cu", "void group_substep("),
          ("solo_kernels.cu", "void step_load("), ("solo_kernels.cu", "void step_finish("),
          ("solo_kernels.cu", ") step_kernel("), ("solo_kernels.cu", ") wide_step_kernel(")]


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else "step_kernel"
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    name, hdr, body = sass_page(rep, kernel, skip)
    m = re.search(r"(wide_)?step_kernel<\(int\)(\d)(?:, \(int\)(\d))?", name)
    if m and m.group(1):
        hint = f"wide_step_kernelILi{m.group(2)}E"
    elif m:
        hint = f"step_kernelILi{m.group(2)}ELi{m.group(3)}E"
    else:
        hint = "step_kernel"
    chains = line_chains(lib, hint)
    csrc = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "solorl_b200", "csrc")
    files = {f: open(os.path.join(csrc, f)).read().splitlines() for f in ("solo_kernels.cu", "solo_wide.cuh")}

    def body_range(f, sig):
        src = files[f]
        s = next((i for i, l in enumerate(src) if sig in l), None)
        if s is None:
            return None
        s += 1
        e = next(i for i in range(s, len(src)) if src[i].startswith("}")) + 1
        return f, s, e
    scopes = [r for r in (body_range(f, sig) for f, sig in SCOPES) if r]
    base = int(body[0][0], 16)
    i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    si = {s: hdr.index(s) for s in STALLS if s in hdr}
    agg = defaultdict(lambda: defaultdict(float))
    tot_i = tot_s = 0
    for r in body:
        off = int(r[0], 16) - base
        ch = chains.get(off, [])
        key = None
        for f, l in ch:                       # innermost -> outermost: first frame inside a phase-level function wins
            if any(f == sf and s0 <= l <= s1 for sf, s0, s1 in scopes):
                key = (f, l)
                break
        if key is None:
            key = ("", -1)
        a = agg[key]
        a["inst"] += int(r[i_inst]); a["samp"] += int(r[i_samp]); a["sass"] += 1
        for s, j in si.items():
            a[s] += int(r[j])
        tot_i += int(r[i_inst]); tot_s += int(r[i_samp])
    tot_nb = tot_s - sum(a.get("stall_barrier", 0) for a in agg.values())
    print(f"# {name}: {len(body)} SASS instructions ({len(body) * 16 / 1024:.0f} KiB), "
          f"{tot_i} warp instructions executed, {tot_s} stall samples ({tot_nb:.0f} outside barriers)")
    print(f"# samp% = share of all samples; work% = share of the samples that are not barrier waits")
    print(f"# {'file:line':>20} {'sass':>6} {'inst%':>6} {'samp%':>6} {'work%':>6}  " + " ".join(f"{s[6:10]:>5}" for s in si) + "  source")
    for key in sorted(agg, key=lambda k: -agg[k]["samp"]):
        a = agg[key]
        if a["samp"] < 0.002 * tot_s and a["inst"] < 0.002 * tot_i:
            continue
        f, l = key
        text = files[f][l - 1].strip()[:64] if l > 0 else "(no line info)"
        st = " ".join(f"{100 * a[s] / max(tot_s, 1):5.1f}" for s in si)
        work = a["samp"] - a.get("stall_barrier", 0)
        print(f"  {f[5:12] + ':' + str(l):>20} {int(a['sass']):6d} {100 * a['inst'] / tot_i:6.1f} {100 * a['samp'] / max(tot_s, 1):6.1f} "
              f"{100 * work / max(tot_nb, 1):6.1f}  {st}  {text}")
    st = " ".join(f"{100 * sum(agg[k][s] for k in agg) / max(tot_s, 1):5.1f}" for s in si)
    print(f"  {'total':>20} {len(body):6d} {100.0:6.1f} {100.0:6.1f} {100.0:6.1f}  {st}")


if __name__ == "__main__":
    main()
