"""SASS opcode histogram of the in-tree library (cuobjdump -sass), per kernel family: the committed evidence that the hot
path is tcgen05 / TMA / TMEM code (UTCHMMA, UTMALDG, UTMAREDG, LDTM, STTM, FFMA2 ...) -- the .so itself is not in the history.

    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "pytorch-simclr_b200", "lib", "libsimclr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fam = None
hist = collections.defaultdict(collections.Counter)
nfun = collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        if "contrastive_tile_kernel" in name:
            args = re.search(r"contrastive_tile_kernelILi(\d+)ELi(\d)ELb(\d)ELb(\d)ELi(\d)ELb(\d)E", name)
            fam = f"contrastive_tile_kernel {'backward' if args.group(3) == '1' else 'forward'}" if args else "contrastive_tile_kernel"
        else:
            k = re.search(r"simclr\d+(\w+?_kernel)", name)
            fam = k.group(1) if k else "other"
        nfun[fam] += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and fam:
        hist[fam][m.group(1).split(".")[0] if not m.group(1).startswith(("UTC", "UTMA", "LDTM", "STTM", "MUFU", "SYNCS", "BAR", "UBLK", "RED", "ATOM", "MULTIMEM")) else m.group(1)] += 1
stamp = open(LIB + ".stamp").read().strip() if os.path.isfile(LIB + ".stamp") else "?"
print(f"SASS opcode histogram of pytorch-simclr_b200/lib/libsimclr_b200.so (library stamp {stamp}); cuobjdump -sass, all instantiations of a family added")
KEY = ("UTC", "UTMA", "LDTM", "STTM", "UBLKCP", "MUFU", "FFMA2", "FADD2", "FMUL2", "SYNCS", "BAR", "RED", "MULTIMEM", "HMMA", "HGMMA")
for f in sorted(hist):
    h = hist[f]
    print(f"\n== {f}: {nfun[f]} instantiation(s), {sum(h.values())} instructions")
    key = {k: v for k, v in h.items() if k.startswith(KEY)}
    print("   Blackwell / async-proxy / packed-fp32 opcodes: " + ", ".join(f"{k} {v}" for k, v in sorted(key.items(), key=lambda kv: -kv[1])))
    print("   top 12: " + ", ".join(f"{k} {v}" for k, v in h.most_common(12)))
    legacy = sum(v for k, v in h.items() if k.startswith(("HMMA", "HGMMA", "IMMA")))
    print(f"   legacy tensor opcodes (HMMA / HGMMA / IMMA): {legacy}")
