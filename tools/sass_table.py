#!/usr/bin/env python
"""Static evidence from the built library (no GPU needed): per kernel, the count of the SASS mnemonics that prove
tcgen05 / TMEM / TMA / programmatic dependent launch (names per B200_PROFILING.md) and the resource usage ptxas
settled on (registers, stack = spill space, static shared memory).

    python tools/sass_table.py > profiles/sass_r02.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "advshadow-camouflaged-adversarial-attacks-via-conditional-diffusion-model-generated-shadows_b200"
LIB = os.path.join(ROOT, PKG, "lib", "libadvshadow_b200.so")
COLS = [("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), (".2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UTCBAR", r"\bUTCBAR"), ("ACQBULK", r"\bACQBULK"), ("PREEXIT", r"\bPREEXIT"),
        ("HMMA", r"\bHMMA\.")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for d in out:
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(.*$", "", d)
        d = d.replace("advs::", "").replace("(anonymous namespace)::", "").replace("__nv_bfloat16", "bf16").replace("__half", "f16")
        d = re.sub(r"\((bool)\)([01])", lambda m: "true" if m.group(2) == "1" else "false", d)
        d = re.sub(r"\(int\)", "", d)
        short.append(d)
    return short


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
    kernels = {}
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = [0] * len(COLS)
            continue
        if cur is None or "/*" not in ln:
            continue
        for i, (_, pat) in enumerate(COLS):
            if re.search(pat, ln):
                kernels[cur][i] += 1
    names = list(kernels)
    short = dict(zip(names, demangle(names)))
    print("Static evidence for lib/libadvshadow_b200.so (tools/sass_table.py; `cuobjdump -sass` instruction counts and")
    print("`cuobjdump -res-usage`).  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st,")
    print("UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit, ACQBULK / PREEXIT = griddepcontrol.wait /")
    print("launch_dependents (programmatic dependent launch), HMMA. = legacy mma.sync (none).  STACK > 0 = spill / local space.")
    print()
    hdr = f"{'kernel':58s}" + "".join(f"{c:>8s}" for c, _ in COLS) + f"{'REG':>6s}{'STACK':>6s}{'SMEM':>7s}"
    print(hdr)
    tot = [0] * len(COLS)
    rows = []
    for k in names:
        c = kernels[k]
        tot = [a + b for a, b in zip(tot, c)]
        rows.append((short[k], c, usage.get(k, (0, 0, 0))))
    tensor = [r for r in rows if r[1][0] or r[1][1]]
    other = [r for r in rows if not (r[1][0] or r[1][1])]
    for name, c, u in sorted(tensor, key=lambda r: r[0]):
        print(f"{name[:57]:58s}" + "".join(f"{v:8d}" for v in c) + f"{u[0]:6d}{u[1]:6d}{u[2]:7d}")
    print()
    print(f"{'whole library (%d kernels)' % len(rows):58s}" + "".join(f"{v:8d}" for v in tot))
    print()
    print("Other kernels (SIMT / bandwidth-bound; PDL columns and resources only):")
    print(f"{'kernel':58s}{'ACQBULK':>8s}{'PREEXIT':>8s}{'REG':>6s}{'STACK':>6s}{'SMEM':>7s}")
    ia, ip = [c for c, _ in COLS].index("ACQBULK"), [c for c, _ in COLS].index("PREEXIT")
    for name, c, u in sorted(other, key=lambda r: r[0]):
        print(f"{name[:57]:58s}{c[ia]:8d}{c[ip]:8d}{u[0]:6d}{u[1]:6d}{u[2]:7d}")
    print()
    print("No UTMASTG: outputs leave through 256-bit st.global from registers (an smem-staged TMA store measured slower for")
    print("the level-0 convs in round 1: it cost a weight stage and L2 bandwidth the main loop needed).")


if __name__ == "__main__":
    sys.exit(main())
