"""Per-kernel census of the Blackwell-specific SASS in libgcm_b200.so (cuobjdump -sass): tcgen05 MMAs (UTCHMMA / UTCQMMA...),
TMEM loads / stores (LDTM / STTM), TMEM alloc, bulk copies (UBLKCP), tensor-map TMA (UTMALDG), cp.async (LDGSTS), mbarrier
ops (SYNCS), tcgen05 commit barriers (UTCBAR).  Usage: python tools/sass_census.py > profiles/sass_census_r2.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "graph-conv-memory_b200", "gcm", "_lib", "libgcm_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCATOMSWS", "UTCBAR", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "MUFU", "ACQBULK"]
rows, cur, cnt, n_inst = [], None, None, 0
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt, n_inst))
        cur, cnt, n_inst = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        n_inst += 1
        op = m.group(1).split(".")[0]
        if op in WANT:
            cnt[op] += 1
if cur:
    rows.append((cur, cnt, n_inst))
dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("# SASS census of `libgcm_b200.so` (sm_100a), round 2\n")
print("`cuobjdump -sass graph-conv-memory_b200/gcm/_lib/libgcm_b200.so`, instructions counted per kernel (static counts; every template "
      "instantiation is its own row).  UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st, "
      "UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc / dealloc, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG = tensor-map TMA "
      "(not used: every tile here is a set of 128-byte rows or one contiguous block), LDGSTS = cp.async, SYNCS = mbarrier ops.\n")
cols = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "MUFU"]
print("| kernel | SASS instr | " + " | ".join(cols) + " |")
print("|---|---|" + "---|" * len(cols))
tot = collections.Counter()
for (name, cnt, n), d in sorted(zip(rows, dem), key=lambda t: -t[0][1]["UTCHMMA"] * 100000 - t[0][2]):
    short = re.sub(r"\(anonymous namespace\)::", "", d)
    short = short.replace("(int)", "").replace("(bool)", "")
    short = re.sub(r"\(.*", "", short).replace("void ", "")
    for c in cols:
        tot[c] += cnt[c]
    if n < 40:
        continue
    print(f"| `{short}` | {n} | " + " | ".join(str(cnt[c]) if cnt[c] else "" for c in cols) + " |")
print(f"| **total** | | " + " | ".join(str(tot[c]) for c in cols) + " |")
