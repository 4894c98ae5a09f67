"""Count the SASS mnemonics that prove tcgen05 / TMEM / TMA use per kernel of libcvflow.so (run here, no GPU needed):
python profiles/sass_mnemonics.py > profiles/r01_sass_mnemonics.txt"""
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cosyvoice_lora_finetune_framework_b200",
                   "libcvflow.so")
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS",
        "MUFU.EX2", "MUFU.TANH", "FFMA2", "FMNMX3", "HMMA", "ACQBULK", "CCTL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, counts, order = None, {}, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[cur] = dict.fromkeys(WANT, 0)
        counts[cur]["_instr"] = 0
        order.append(cur)
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        counts[cur]["_instr"] += 1
        for w in WANT:
            if re.search(r"(?<![A-Z])" + re.escape(w), line):
                counts[cur][w] += 1
print("SASS mnemonic counts per kernel of libcvflow.so (sm_100a); UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,")
print("UTMALDG/UTMASTG = TMA tensor loads/stores, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0)\n")
for k in order:
    c = counts[k]
    hits = "  ".join("%s=%d" % (w, c[w]) for w in WANT if c[w])
    print("%-52s instr=%-6d %s" % (k[:52], c["_instr"], hits))
