"""Opcode evidence for the Blackwell-native claim: per compiled object of libslas_cuda.so (slas_b200/lib/obj/*.o, built
by `make -C slas_b200/csrc` for sm_100a), the number of SASS instructions of the families that matter, from
`cuobjdump -sass`; and per source file the inline-PTX mnemonics it contains.  Runs on the CPU box (no GPU needed).
Usage: python profiles/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = [
    ("UTCHMMA", r"\bUTC[A-Z]*MMA"),            # tcgen05.mma (UTCHMMA, .2CTA for cta_group::2)
    ("UTCBAR", r"\bUTCBAR"),                   # tcgen05.commit -> mbarrier
    ("LDTM", r"\bLDTM"),                       # tcgen05.ld (TMEM -> registers)
    ("UTMALDG", r"\bUTMALDG"),                 # cp.async.bulk.tensor (TMA tiled load)
    ("UBLKCP", r"\bUBLKCP"),                   # cp.async.bulk (1-D bulk TMA), both directions
    ("SYNCS", r"\bSYNCS"),                     # mbarrier arrive / try_wait
    ("UCGABAR", r"\bUCGABAR"),                 # barrier.cluster
    ("ACQBULK/PDL", r"\bACQBULK|\bPREEXIT|\bDEPBAR\.DEP|\bUGETNEXTWORKID"),  # griddepcontrol.* lowers to these on sm_100
    ("FFMA", r"\bFFMA"),
    ("DFMA", r"\bDFMA"),
    ("HMMA/IMMA/mma.sync", r"\b[HI]MMA|\bDMMA"),  # legacy warp-level MMA: expected 0
    ("LDG.E.128", r"\bLDG\.E[A-Z0-9.]*\.128\b"),
    ("LDG.E.256", r"\bLDG\.E[A-Z0-9.]*\.256\b"),
    ("STG.E.128", r"\bSTG\.E[A-Z0-9.]*\.128\b"),
    ("STG.E.256", r"\bSTG\.E[A-Z0-9.]*\.256\b"),
    ("LDS.128", r"\bLDS[A-Z.]*\.128"),
    ("MEMBAR", r"\bMEMBAR"),
    ("ATOMG/RED", r"\bATOMG|\bRED\."),
]
PTX = ["tcgen05.mma", "tcgen05.ld", "tcgen05.alloc", "tcgen05.commit", "cp.async.bulk.tensor", "cp.async.bulk.shared", "cp.async.bulk.global",
       "mbarrier.", "barrier.cluster", "griddepcontrol.", "st.release.sys", "ld.acquire.sys", "ld.global.L1::no_allocate.v8", "st.global.v8",
       "cvt.rna.tf32"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "slas_b200", "lib", "obj", "*.o")))
    print("# SASS opcode counts per object (cuobjdump -sass, sm_100a); kernels = number of device functions")
    hdr = ["object", "kernels"] + [f for f, _ in FAMILIES]
    print("\t".join(hdr))
    for o in objs:
        r = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True)
        if r.returncode != 0 or "Function :" not in r.stdout:
            continue
        text = r.stdout
        counts = collections.OrderedDict((f, len(re.findall(rx, text))) for f, rx in FAMILIES)
        print("\t".join([os.path.basename(o), str(text.count("Function :"))] + [str(v) for v in counts.values()]))
    print()
    print("# inline-PTX mnemonics per source file (occurrences in slas_b200/csrc)")
    srcs = sorted(glob.glob(os.path.join(ROOT, "slas_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "slas_b200", "csrc", "*.cuh")))
    print("\t".join(["file"] + PTX))
    for sfile in srcs:
        t = open(sfile).read()
        c = [t.count(m) for m in PTX]
        if any(c):
            print("\t".join([os.path.basename(sfile)] + [str(x) for x in c]))
    print()
    print("# which SASS spelling the dependent-launch instructions get (first matches in reduce.o)")
    r = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "slas_b200", "lib", "obj", "reduce.o")], capture_output=True, text=True)
    seen = set()
    for line in r.stdout.splitlines():
        m = re.search(r"\b(ACQBULK|PREEXIT|DEPBAR[.A-Z]*|UGETNEXTWORKID[.A-Z]*|ERRBAR|CCTL[.A-Z]*)\b", line)
        if m and m.group(1) not in seen:
            seen.add(m.group(1))
            print("  ", line.strip()[:120])


if __name__ == "__main__":
    main()
