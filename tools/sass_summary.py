"""Per-kernel SASS evidence of the Blackwell-native paths in the built library (no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
Counts, per kernel, the mnemonics that prove tcgen05 / TMEM / TMA / programmatic dependent launch (guide:
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG/UTMASTG)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "aa-clip-iqm_b200", "libaaclip_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
def strip_args(sig: str) -> str:
    """'void f<(int)3>(A, B)' -> 'void f<(int)3>' (drops the trailing parameter list only)."""
    depth = 0
    for i in range(len(sig) - 1, -1, -1):
        if sig[i] == ")":
            depth += 1
        elif sig[i] == "(":
            depth -= 1
            if depth == 0:
                return sig[:i]
    return sig


PAT = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
    ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("MUFU.EX2", r"\bMUFU\.EX2"),
    ("HMMA", r"\bHMMA"), ("PDL wait/launch", r"\b(ACQBULK|PREEXIT|GRIDDEPCONTROL|DEPBAR\.LE\.PDL)"),
])
cur, rows = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        rows[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    rows[cur]["instr"] += bool(re.match(r"\s+/\*[0-9a-f]{4}\*/", line))
    for k, p in PAT.items():
        if re.search(p, line):
            rows[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} ({os.path.getsize(lib)} bytes); instruction counts per kernel (static)")
print("# " + "  ".join(f"{k}" for k in PAT))
tot = collections.Counter()
for fn, c in rows.items():
    if not any(c[k] for k in PAT):
        continue
    name = strip_args(demangle(fn)).replace("void ", "").replace("aaclip::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\((?:int|bool)\)", "", name)
    print(f"{name[:60]:60s} instr={c['instr']:6d} " + " ".join(f"{k}={c[k]}" for k in PAT if c[k]))
    tot.update(c)
print("TOTAL " + " ".join(f"{k}={tot[k]}" for k in PAT))
print(f"kernels in the library: {len(rows)}; with tcgen05 MMAs: {sum(1 for c in rows.values() if c['UTCHMMA'])}; "
      f"with legacy HMMA (mma.sync): {sum(1 for c in rows.values() if c['HMMA'])}")
