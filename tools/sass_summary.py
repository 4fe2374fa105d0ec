"""profiles/sass_summary.txt: per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md: tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UBLKCP), from the in-tree libb2h.so.
No GPU needed:  python tools/sass_summary.py > profiles/sass_summary.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hand_pose_sl_b200", "libb2h.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
PAT = [("UTCHMMA", r"\bUTCHMMA\b"), ("UTCBAR", r"\bUTCBAR\b"), ("LDTM", r"\bLDTM\b"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP\b"),
       ("SYNCS", r"\bSYNCS\b"), ("STG.STRONG.SYS", r"\bSTG\.E\.64\.STRONG\.SYS\b"), ("ST.STRONG.SYS", r"\bST\.E\.64\.STRONG\.SYS\b"),
       ("HMMA", r"\bHMMA\b"), ("FFMA", r"\bFFMA\b")]
print("# cuobjdump -sass hand_pose_sl_b200/libb2h.so (sm_100a), mnemonic counts per kernel")
print("# UTCHMMA = tcgen05.mma kind::f16, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA tensor map),")
print("# UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops; HMMA (legacy mma.sync) must be 0 everywhere.")
print("# Peer / multicast exchange of the fused train kernel: ST.E.64.STRONG.SYS = unicast 8-byte pushes to the peers' buffers,")
print("# STG.E.64.STRONG.SYS = multimem.st.relaxed.sys.global.b64 (the multicast-ness is in the address, not the opcode).")
print(f"{'kernel':58s} " + " ".join(f"{k:>9s}" for k, _ in PAT) + "   instrs")
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void ", "").replace("b2h::", "")
    cnt = [len(re.findall(p, f)) for _, p in PAT]
    n = len(re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", f, flags=re.M))
    print(f"{dem:58s} " + " ".join(f"{c:9d}" for c in cnt) + f"   {n}")
