"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass): the evidence that the hot kernels are
Blackwell-native -- UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA), LDTM / STTM (TMEM), FFMA2 (packed fp32); HMMA (legacy
mma.sync) must only appear in the warp-level cross-check kernels of attention.cu.

    python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slamkit_b200", "libslamkit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "FFMA2", "FMUL2", "FADD2", "MUFU", "SYNCS"]
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(anonymous namespace\)::", "", cur)
        cur = re.sub(r"\(.*", "", cur).replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        for w in want:
            if op.startswith(w):
                cnt[cur][w] += 1
print(f"{'kernel':70s} " + " ".join(f"{w:>8s}" for w in want))
for k in sorted(cnt):
    if any(cnt[k][w] for w in want[:9]):
        print(f"{k[:70]:70s} " + " ".join(f"{cnt[k][w]:8d}" for w in want))
