"""Build-container tool: instruction mix of every kernel in libstpyb.so (cuobjdump -sass) and the main-loop
listing of the dominant kernel (trailing SYRK: gemm_nt_kernel<128x64, BK=32 x 2 stages, EpiAccum>).
Writes profiles/sass_mix_r02.txt and profiles/sass_syrk_mainloop_r02.txt."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "stpy_b200", "libstpyb.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append(m.group(2).strip())
key = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDGSTS", "LDG", "STG", "LDS", "STS", "UTMALDG", "BAR", "ATOM", "RED"]
out = ["# instruction mix per kernel of stpy_b200/libstpyb.so (static SASS counts, cuobjdump -sass, sm_100a)",
       "# total  " + "  ".join(key) + "   kernel"]
for name, ins in funcs.items():
    ops = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", i).split()[0].split(".")[0] for i in ins)
    out.append("%6d  " % len(ins) + "  ".join("%4d" % ops.get(k, 0) for k in key) + "   " + demangle(name)[:150])
open(os.path.join(ROOT, "profiles", "sass_mix_r02.txt"), "w").write("\n".join(out) + "\n")
# main loop of the update kernel: the backward-branch loop that contains the most DMMAs
for name, ins in funcs.items():
    d = demangle(name)
    if "gemm_nt_kernel" in d and "32>" in d and "EpiAccum" in d and "tma" not in d:
        idx = [i for i, s in enumerate(ins) if "DMMA" in s]
        lo, hi = idx[0], idx[-1]
        body = ins[max(0, lo - 8): hi + 8]
        open(os.path.join(ROOT, "profiles", "sass_syrk_mainloop_r02.txt"), "w").write(
            "# %s\n# main loop region (first to last DMMA), %d instructions, %d DMMA.8x8x4, %d LDS, %d LDGSTS\n" % (
                d, len(body), sum("DMMA" in s for s in body), sum(s.startswith("LDS") or " LDS" in s for s in body),
                sum("LDGSTS" in s for s in body)) + "\n".join(body) + "\n")
        break
print("\n".join(out[:6]))
