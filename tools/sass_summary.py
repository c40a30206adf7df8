#!/usr/bin/env python
"""Opcode summary of the headline kernel's SASS for profiles/ (CPU only: cuobjdump on the object the library is linked from):
python tools/sass_summary.py build/gpu/fused_inst_lean4.o > profiles/r2_fused_step_sass.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict

WATCH = ["BAR.SYNC.DEFER_BLOCKING", "SHFL.DOWN", "SHFL.UP", "STG.E.EF.64", "SYNCS.ARRIVE.TRANS64", "SYNCS.ARRIVE.TRANS64.A1T0",
         "SYNCS.CCTL.IVALL", "SYNCS.PHASECHK.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "UBLKCP.S.G", "UBLKPF.L2", "DADD", "DMUL", "DFMA",
         "MUFU", "LDS.64", "LDS", "STS", "LDG", "LDC", "IMAD", "ELECT", "R2UR"]
PREFIX = {"MUFU", "LDS", "STS", "LDG", "LDC", "IMAD", "R2UR"}  # counted by prefix (every variant)
TENSOR = ("UTCMMA", "UTCLD", "UTCST", "HMMA", "DMMA", "UTCHMMA", "UTCQMMA")


def main(obj):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    kernels, cur = OrderedDict(), None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m and cur is not None:
            cur.append((m.group(1), ln))
    print("# SASS of the headline kernel, from the object the library is linked from (cuobjdump -sass %s; nvcc 12.9," % obj)
    print("# -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false; tools/sass_summary.py).  The fused step is an HBM-bound FP64 stencil: no")
    print("# tensor-core (tcgen05 / UTCMMA) instructions by design (BASELINE.json north_star); what to look for is the TMA bulk-copy staging")
    print("# (UBLKCP.S.G = cp.async.bulk global->shared), the mbarrier traffic (SYNCS.*), warp shuffles, FP64 arithmetic without contraction")
    print("# (DADD/DMUL; DFMA only inside division / sqrt sequences) and the evict-first 64-bit stores (STG.E.EF.64).")
    for name, ins in kernels.items():
        if "k_fused_step" not in name:
            continue
        ops = Counter(o for o, _ in ins)
        print("\nkernel %s\n  instructions: %d" % (name, len(ins)))
        for w in WATCH:
            n = sum(c for o, c in ops.items() if (o == w or (w in PREFIX and o.split(".")[0] == w)))
            if w == "LDS":
                n = sum(c for o, c in ops.items() if o.split(".")[0] == "LDS")
            if n:
                print("  %-40s %d" % (w, n))
        print("  tensor-core / TMEM instructions (%s): %d" % (", ".join(TENSOR[:5]), sum(c for o, c in ops.items() if o.split(".")[0] in TENSOR)))
    # excerpt: the refill of a row in the first kernel -- from an mbarrier arrive with a transaction count to the bulk copy
    for name, ins in kernels.items():
        if "k_fused_step" not in name:
            continue
        idx = [i for i, (o, _) in enumerate(ins) if o == "UBLKCP.S.G"]
        if len(idx) >= 3:
            k = idx[2]
            print("\n# excerpt (%s): a row's refill -- wait for the layer's column groups, arm the mbarrier with the byte count, issue the bulk copy" % name[:60])
            for _, ln in ins[max(0, k - 28):k + 3]:
                print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", ln))
        break


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "build/gpu/fused_inst_lean4.o")
