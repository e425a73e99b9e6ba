"""profiles/r02_sass_excerpt.md: mnemonic counts of the tcgen05 / TMEM / TMA / mbarrier instructions in the shipped library
(`cuobjdump -sass`), in total and for the conv / wgrad kernels the bench and the PWC-Net path launch."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fisr_b200", "lib", "libfisr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda names: dict(zip(names, subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()))
WANT = re.compile(r"\b(UTC[A-Z]+MA(?:\.[A-Z0-9_]+)*|UTCBAR(?:\.[A-Z0-9_]+)*|UTCATOMSWS(?:\.[A-Z0-9_]+)*|LDTM|STTM|UTMALDG(?:\.[A-Z0-9_]+)*|UTMASTG(?:\.[A-Z0-9_]+)*|SYNCS(?:\.[A-Z0-9_]+)*|HMMA(?:\.[A-Z0-9_]+)*)\b")
total, per, cur = collections.Counter(), collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur:
        w = WANT.match(m.group(1))
        if w:
            key = re.sub(r"\.(x\d+|16dp\d+bit|32dp\d+bit|\d+x\d+b?)", "", w.group(1))
            total[key] += 1
            per[cur][key.split(".")[0] if not key.startswith("UTMA") else key.split(".")[0]] += 1
names = demangle(list(per))
short = lambda n: re.sub(r"\(.*", "", names[n].replace("(anonymous namespace)::", "")).replace("void ", "").replace("fisr::convk::", "").replace("fisr::", "")
out = ["# r02 -- SASS evidence that the shipped library is tcgen05 / TMEM / TMA code", "",
       "`python tools/sass_excerpt.py` = `cuobjdump -sass fisr_b200/lib/libfisr_b200.so`, library built by `__graft_entry__.build()` with `nvcc -gencode "
       "arch=compute_100a,code=sm_100a -O3 -lineinfo` (CUDA 12.9).  Mnemonics per `/opt/skills/guides/B200_PROFILING.md`: `UTCHMMA` / `UTCQMMA` = "
       "tcgen05.mma kind::f16 / kind::f8f6f4, `LDTM` = tcgen05.ld, `UTMALDG` / `UTMASTG` = TMA tensor loads / stores, `UTCBAR` = tcgen05.commit, "
       "`SYNCS` = mbarrier operations.", "", "## Totals over all kernels in the library", "", "| mnemonic | count |", "|---|---:|"]
out += [f"| `{k}` | {v} |" for k, v in total.most_common() if not k.startswith("HMMA")]
cols = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR"]
pick = [n for n in per if re.search(r"conv3x3_umma_kernel<(16|32|64|96|128), 2, (2|3), (0|1|2|3)>|wgrad3x3_umma_kernel", names[n])]
out += ["", "## Per kernel: the `conv3x3_umma_kernel<NT, CHUNKS = 2, PLANES, EPI>` instantiations the forward paths launch (PLANES 3 = f16f8 headline, "
        "PLANES 2 = f16x3 / PWC-Net incl. the N = 96 tile) and the wgrad kernel", "", "| kernel | " + " | ".join(cols) + " |", "|---|" + "---:|" * len(cols)]
for n in sorted(pick, key=short):
    out.append(f"| `{short(n)}` | " + " | ".join(str(per[n][c]) for c in cols) + " |")
mma_kernels = sum(1 for n in per if per[n]["UTCHMMA"] + per[n]["UTCQMMA"] > 0)
legacy = sum(v for k, v in total.items() if k.startswith("HMMA"))
out += ["", f"{mma_kernels} kernels in the library issue tcgen05 MMAs; legacy `HMMA` (mma.sync) instructions in the same listing: {legacy}."]
open(os.path.join(ROOT, "profiles", "r02_sass_excerpt.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:14]))
