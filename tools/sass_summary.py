"""Per-kernel SASS evidence for the tensor-core / TMA path: counts of the Blackwell mnemonics in libseldq.so.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTCHMMA = tcgen05.mma (fp16 / bf16 kind), UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk,
LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sound-event-localization-and-detection_b200", "libseldq.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "HMMA",
             "FFMA", "MUFU", "ATOM", "RED"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    out = subprocess.check_output(["cuobjdump", "-sass", lib], stderr=subprocess.STDOUT).decode("utf-8", "replace")
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), {k: 0 for k in MNEMONICS})
            cur["_n"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        cur["_n"] += 1
        op = m.group(1).split(".")[0]
        if op in cur:
            cur[op] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# %s -- cuobjdump -sass, instruction counts per kernel (sm_100a)" % os.path.relpath(lib, ROOT))
    cols = [k for k in MNEMONICS if any(v[k] for v in kernels.values())]
    print("%-78s %6s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    for (name, v), dn in sorted(zip(kernels.items(), demangle), key=lambda kv: kv[1]):
        dn = re.sub(r"\(.*", "", dn).replace("seldq::", "")
        dn = re.sub(r"^void ", "", dn)
        print("%-78s %6d " % (dn[:78], v["_n"]) + " ".join("%7d" % v[c] for c in cols))


if __name__ == "__main__":
    main()
