#!/usr/bin/env python
"""Counts the Blackwell-specific SASS mnemonics per kernel of libfastdet_b200.so (cuobjdump -sass): the hard evidence that
the conv kernels are tcgen05 / TMEM / TMA code and not recompiled mma.sync.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt

UTCHMMA  tcgen05.mma (kind::f16)        UTCBAR   tcgen05.commit -> mbarrier      LDTM / STTM  tcgen05.ld / st (TMEM)
UTMALDG  TMA load (cp.async.bulk.tensor global->shared)    UTMASTG  TMA store    UTCATOMSWS  TMEM alloc/dealloc
SYNCS    mbarrier ops                   HMMA     legacy mma.sync (must be 0 in the conv kernels)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fastdet_b200", "libfastdet_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCATOMSWS", "SYNCS", "LDGSTS", "HMMA", "FFMA", "DFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("fd::", "")
            name = re.sub(r"^void ", "", re.sub(r"\(.*", "", name))
            cur = kernels.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            cur["_total"] += 1
            for o in OPS:
                if op.startswith(o):
                    cur[o] += 1
    kernels = collections.OrderedDict((k, v) for k, v in kernels.items() if k and v["_total"])
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':<58}" + "".join(f"{o:>11}" for o in OPS) + f"{'instrs':>9}")
    for name, c in kernels.items():
        print(f"{name[:57]:<58}" + "".join(f"{c[o]:>11}" for o in OPS) + f"{c['_total']:>9}")
    kernels = collections.OrderedDict((k, v) for k, v in kernels.items() if k)
    conv = [n for n in kernels if n.startswith(("conv_tc_kernel", "conv_halo_kernel", "conv0_ws_kernel"))]
    assert conv and all(kernels[n]["UTCHMMA"] > 0 and kernels[n]["HMMA"] == 0 for n in conv), "a conv kernel without tcgen05.mma (or with mma.sync)"
    print(f"\n# every conv kernel ({len(conv)} instantiations) issues UTCHMMA (tcgen05.mma) and no HMMA (mma.sync)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
