"""profiles/sass_summary.txt: static instruction counts per kernel of the built library (cuobjdump -sass | c++filt).
python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "speech_enhancement_by_s3prl_b200", "libse_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
keys = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "FADD2", "FFMA2", "FMUL2", "SHFL", "MUFU", "ATOMG", "REDG"]
print("cuobjdump -sass speech_enhancement_by_s3prl_b200/libse_b200.so (sm_100a), instruction counts per kernel (static).")
print("UTC*MMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA loads / bulk copies, LDGSTS = cp.async, "
      "FFMA2 / FADD2 / FMUL2 = packed f32x2.\n")
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for k in keys:
            if op.startswith(k):
                counts[cur][k] += 1
names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
tot = collections.Counter()
for name, (mangled, c) in zip(names, counts.items()):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    if not c:
        continue
    tot.update(c)
    print(name)
    print("    " + ", ".join(f"{k} {v}" for k, v in c.most_common()))
print("\nwhole library: " + ", ".join(f"{k} {v}" for k, v in tot.most_common()))
