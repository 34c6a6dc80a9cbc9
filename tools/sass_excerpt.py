"""SASS evidence for the tcgen05 / TMEM / TMA kernels: per kernel of libsininn.so, the count of each Blackwell-specific
mnemonic and the first occurrence of each (cuobjdump -sass; no GPU needed).  python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sin_inn_b200", "libsininn.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UTMAPF|UBLKCP|SYNCS|ELECT|UCGABAR_ARV|UCGABAR_WAIT|HMMA|ACQBULK)\b[.\w]*")
cur, data = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        data[cur] = (collections.Counter(), {})
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        key = m.group(1)
        data[cur][0][key] += 1
        data[cur][1].setdefault(key, re.sub(r"\s+", " ", line.split("/*")[1].split("*/")[0] + line.split("*/")[1].split("/*")[0]).strip() if "/*" in line else line.strip())
print("# SASS mnemonics of the tensor-core kernels in sin_inn_b200/libsininn.so (cuobjdump -sass, sm_100a)")
print("# tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR, TMA loads/stores/reductions -> UTMALDG / UTMASTG / UTMAREDG,")
print("# mbarrier -> SYNCS, elect.sync -> ELECT; no HMMA (legacy mma.sync) anywhere.\n")
for fn, (cnt, first) in data.items():
    if not (cnt.get("UTCHMMA") or cnt.get("UTMALDG")):
        continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    print(f"## {name[:150]}")
    print("   " + "  ".join(f"{k} x{v}" for k, v in sorted(cnt.items())))
    for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR"):
        if k in first:
            print(f"   e.g. {first[k][:150]}")
    print()
tot = collections.Counter()
for fn, (cnt, _) in data.items():
    tot.update(cnt)
print("# whole library:", "  ".join(f"{k} x{v}" for k, v in sorted(tot.items())))
