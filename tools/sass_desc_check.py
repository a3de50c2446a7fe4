"""Heuristic SASS check for the ptxas hazard found in round 2 (fc_wgrad_umma_kernel): a tcgen05.mma (UTCHMMA) operand
`gdesc[URn]` is the 64-bit uniform register pair URn:URn+1; in one build ptxas dropped the instruction that writes the high
word of a descriptor inside an unrolled single-thread loop, so the MMA read a stale kernel-parameter word as its stride field.
For every UTCHMMA of every kernel in the given objects this script checks that BOTH registers of each descriptor pair are
written by some arithmetic / move instruction of that function (not only by a kernel-parameter load).
Usage: python tools/sass_desc_check.py downgan_b200/csrc/*.o     (exit code 1 if a suspicious pair is found)"""
import re, subprocess, sys

def written_regs(line):
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?(\S+)\s+(UR\d+)", line)
    if not m:
        return None, []
    op, dst = m.group(1), int(m.group(2)[2:])
    wide = ".64" in op or ".WIDE" in op
    return op, [dst, dst + 1] if wide else [dst]

bad = 0
for obj in sys.argv[1:]:
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, lines = None, {}
    for l in sass.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            fn = m.group(1); lines[fn] = []
        elif fn and "/*" in l and not l.strip().startswith("/* 0x"):
            lines[fn].append(l)
    for fn, ls in lines.items():
        writes = [written_regs(l) for l in ls]
        for i, l in enumerate(ls):
            if "UTCHMMA" not in l:
                continue
            for m in re.finditer(r"gdesc\[UR(\d+)\]", l):
                lo = int(m.group(1))
                for r in (lo, lo + 1):
                    last = None  # nearest earlier instruction (in address order) that writes URr
                    for j in range(i - 1, -1, -1):
                        op, regs = writes[j]
                        if op and not op.startswith("UTC") and r in regs:
                            last = op
                            break
                    if last is None or last.startswith("LDCU"):
                        bad += 1
                        print(f"{obj}: {fn[:90]}: UR{r} of gdesc[UR{lo}] was last written by {last or 'nothing'} before\n    {l.strip()[:140]}")
print("descriptor pairs suspicious:", bad)
sys.exit(1 if bad else 0)
