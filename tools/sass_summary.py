"""per-kernel counts of the SASS mnemonics that tell a Blackwell-native kernel from a legacy one (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA loads / stores, HMMA = mma.sync.

    python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "latentsync_b200", "_C.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "LDGSTS", "MUFU"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
        counts[cur]["total"] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass latentsync_b200/_C.so ({os.path.getsize(so)} bytes): instructions per kernel by mnemonic")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA load/store (cp.async.bulk.tensor), HMMA = mma.sync")
print(f"{'kernel':78s} " + " ".join(f"{k:>8s}" for k in KEYS) + f" {'total':>8s}")
for (name, c), d in zip(counts.items(), dem):
    d = d[: d.rfind("(")].replace("void ls::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
    if not any(c[k] for k in ("UTCHMMA", "HMMA", "UTMALDG", "LDTM")):
        continue
    print(f"{d[:78]:78s} " + " ".join(f"{c[k]:8d}" for k in KEYS) + f" {c['total']:8d}")
print("#\n# Which attention kernel a plan launch takes (csrc/attention_tc.cu attention_tc_try): >= 128 queries and >= 2 key tiles ->")
print("# attn_tc_kernel / attn_tc2_kernel (tcgen05).  One-key-tile problems (audio cross-attention, 8x8 / 4x4 levels, temporal")
print("# attention) -> attn_fwd_kernel / attn_short_kernel (mma.sync) by default because they are FASTER there; their tcgen05")
print("# implementations attn_one_kernel<D, PACK> (LS_ATTN_ONE=1) and attn_tc_kernel<D, true> (LS_ATTN_TC_ALL=1) are built, tested")
print("# (tests/test_ops_gpu.py::test_attention_every_shape_on_the_other_kernels) and measured (profiles/r2g_*, r2d_*).")
