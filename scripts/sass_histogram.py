#!/usr/bin/env python
"""Opcode histogram per kernel from `cuobjdump -sass libwaxvs_cuda.so` (stdin): the SASS evidence that the kernels are
Blackwell-native -- UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), UBLKCP (TMA bulk copies), LDTM/STTM (TMEM
loads/stores), UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Prints the named opcodes per kernel, then the top opcodes.
usage: cuobjdump -sass wax_b200/libwaxvs_cuda.so | python scripts/sass_histogram.py > profiles/sass_opcodes_rNN.txt"""
import collections
import re
import subprocess
import sys

NAMED = ("UTCHMMA", "UTCQMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "FMNMX3", "HMMA", "LDGSTS")
kernels = collections.OrderedDict()
cur = None
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = collections.Counter()
        kernels[m.group(1)] = cur
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:  # noqa: BLE001
        return name


total = collections.Counter()
print("# opcode histogram of wax_b200/libwaxvs_cuda.so (cuobjdump -sass), named Blackwell opcodes by prefix")
for name, cnt in kernels.items():
    total.update(cnt)
    named = {}
    for op, c in cnt.items():
        for key in NAMED:
            if op == key or op.startswith(key + "."):
                base = key if not (key == "UTCHMMA" and ".2CTA" in op) else "UTCHMMA.2CTA"
                named[base] = named.get(base, 0) + c
    if named:
        short = demangle(name)
        short = short if len(short) < 150 else short[:147] + "..."
        print(f"{short}\n    instructions {sum(cnt.values())}: " + ", ".join(f"{k} x{v}" for k, v in sorted(named.items())))
print("\n# library totals (named opcodes, any suffix)")
for key in NAMED:
    c = sum(v for op, v in total.items() if op == key or op.startswith(key + "."))
    if c:
        print(f"{key:10s} {c}")
print("\n# top 25 opcodes overall")
for op, c in total.most_common(25):
    print(f"{op:28s} {c}")
