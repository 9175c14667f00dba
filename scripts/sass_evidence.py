"""Count the Blackwell-specific SASS mnemonics per kernel of libcar_b200.so (cuobjdump -sass) -> profiles/."""
import collections, os, re, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = os.path.join(ROOT, "cross_attention_renderer_b200", "libcar_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCBAR|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|LDGSTS|HMMA|SYNCS|ELECT)\b")
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        if cur in cnt:
            cur = None
        else:
            cnt[cur] = collections.Counter()
        continue
    if cur:
        for k in pat.findall(line):
            cnt[cur][k] += 1
rows = []
for f, c in cnt.items():
    mang = f[f.index("_ZN"):] if "_ZN" in f else f
    name = subprocess.run(["c++filt", mang], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"car::\(anonymous namespace\)::", "", name)
    name = re.sub(r"\(CUtensorMap_st.*", "", name).split("(car")[0]
    if any(c.get(k, 0) for k in ("UTCHMMA", "LDTM", "UTMALDG", "LDGSTS", "UBLKCP")):
        rows.append((name, c))
rows.sort(key=lambda x: x[0])
keys = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "LDGSTS", "SYNCS", "ELECT", "HMMA"]
print("# cuobjdump -sass libcar_b200.so (sm_100a): Blackwell-specific SASS per kernel, instruction counts in the binary")
print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load,")
print("# LDGSTS = cp.async, SYNCS = mbarrier operations, ELECT = elect.sync")
for n, c in rows:
    print(f"{n[:64]:64s} " + " ".join(f"{k}={c.get(k, 0)}" for k in keys if c.get(k, 0) or k == "HMMA"))
print(f"# UTCHMMA.2CTA instructions: {len(re.findall(r'UTCHMMA.2CTA', sass)) // 2}; HMMA (legacy mma.sync) instructions in the whole library: {sum(c.get('HMMA', 0) for c in cnt.values())}")
