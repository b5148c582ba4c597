"""Static evidence of the built library, no GPU needed: profiles/r02_ptxas_registers.txt (nvcc -Xptxas -v: registers,
static shared memory, stack and spills per kernel) and profiles/r02_sass_opcode_histogram.txt (cuobjdump -sass: opcode
histogram per kernel, with the TMA / tcgen05 mnemonics counted).  Run from the repo root after the library is built."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "php_aho_corasick_b200", "csrc")
LIB = os.path.join(ROOT, "php_aho_corasick_b200", "libacb200.so")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def ptxas_table(path):
    r = subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xptxas", "-v",
                        "-I", os.path.join(ROOT, "include"), "-c", os.path.join(CSRC, "engine.cu"), "-o", os.devnull],
                       capture_output=True, text=True)
    text = r.stderr + r.stdout
    rows, cur = [], None
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"name": m.group(1), "stack": 0, "st": 0, "ld": 0, "regs": 0, "smem": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["st"], cur["ld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            cur["smem"] = int(s.group(1)) if s else 0
    names = demangle([r_["name"] for r_ in rows])
    with open(path, "w") as f:
        f.write("# nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v engine.cu  (scripts/static_summaries.py, final tree of round 2): "
                "registers / static smem / stack+spills per kernel\n")
        f.write("# (dynamic shared memory: ac_scan_kernel / ac_scan_tma_kernel up to 225 KB table window (+ 64 KB text ring), ac_filter_kernel 220 KB bitmap)\n")
        for r_ in rows:
            f.write(f"{r_['regs']:3d} regs {r_['smem']:5d} B smem  stack {r_['stack']:3d} B spill st/ld {r_['st']}/{r_['ld']}   {names[r_['name']]}\n")
    return len(rows)


def sass_histogram(path):
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    names = demangle(list(kernels))
    total = collections.Counter()
    for c in kernels.values():
        total.update(c)
    special = {k: v for k, v in total.items() if k.startswith(("UTMA", "UTC", "LDTM", "STTM", "LDGSTS", "SYNCS", "UBLKCP"))}
    with open(path, "w") as f:
        f.write("# cuobjdump -sass php_aho_corasick_b200/libacb200.so (sm_100a) — opcode histogram per kernel (static instruction counts), "
                "scripts/static_summaries.py, final tree of round 2\n")
        f.write("# TMA / mbarrier / tcgen05 mnemonics library-wide: " + (", ".join(f"{k} {v}" for k, v in sorted(special.items())) or "none") + "\n")
        f.write("# (UTMALDG + SYNCS: ac_scan_tma_kernel stages the haystack text of the full walk as 32-byte x 32-slice boxes, cp.async.bulk.tensor + mbarrier ring;\n")
        f.write("#  tcgen05 (UTC*MMA, LDTM, STTM): none — the hot path is integer table lookups and bit tests.  The prefilter's haystack loads are LDG.E.128\n")
        f.write("#  (ld.global.nc.L1::no_allocate), see DESIGN.md 3.2 / 3.3.)\n\n")
        f.write(f"{len(kernels)} kernels, {sum(total.values())} instructions; library-wide: " + ", ".join(f"{k} {v}" for k, v in total.most_common(24)) + "\n\n")
        for k, c in kernels.items():
            f.write(f"{names[k]}\n    {sum(c.values())} instr: " + ", ".join(f"{o} {v}" for o, v in c.most_common(14)) + "\n")
    return len(kernels)


if __name__ == "__main__":
    n1 = ptxas_table(os.path.join(ROOT, "profiles", "r02_ptxas_registers.txt"))
    n2 = sass_histogram(os.path.join(ROOT, "profiles", "r02_sass_opcode_histogram.txt"))
    print(f"{n1} kernels in the ptxas table, {n2} in the SASS histogram")
