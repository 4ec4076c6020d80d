"""Static in-order issue model of a SASS loop body: cycles per iteration for ONE warp, given result latencies.

The rollout kernel runs ~3.5 warps per scheduler, each stalled most of the time on fixed-latency FP64 dependencies
(ncu: "wait"), so what matters besides the FP64 instruction count is how many dependent instructions ptxas placed
back to back. This tool replays the hot loop (found as the innermost backward branch containing N MUFU.RCP64H) through
a scoreboard: an instruction issues one cycle after its predecessor at the earliest and not before its source
registers / predicates are ready. FP64 latencies are the dependent-issue latencies measured on B200 by
tools/fp64_mix_bench.cu (DFMA 23, DMUL/DADD 25 cycles; the numbers quoted in DESIGN.md / profiles for round 1 were
produced with an earlier calibration of 18). It prints cycles per iteration and the stall histogram.

    cuobjdump -sass file.o > f.sass; python tools/sass_sched.py f.sass [n_rcp_in_loop=2]
"""
import re
import sys
from collections import Counter

LAT = {"DFMA": 23, "DMUL": 25, "DADD": 25, "DSETP": 23, "MUFU": 30, "LDC": 30, "LDCU": 30, "LDS": 30, "LDG": 300}
DEFAULT_LAT = 5
FP64 = ("DFMA", "DMUL", "DADD", "DSETP")


def parse(path):
    ops = []
    for line in open(path):
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+((?:@!?U?P\d+\s+)?)([A-Z0-9_.]+)\s*(.*?);", line)
        if m:
            ops.append((int(m.group(1), 16), m.group(2).strip(), m.group(3), m.group(4)))
    return ops


def regs_of(tok, wide):
    out = []
    for m in re.finditer(r"\b(U?R)(\d+)\b", tok):
        n = int(m.group(2))
        out.append((m.group(1), n))
        if wide:
            out.append((m.group(1), n + 1))
    for m in re.finditer(r"\b(U?P)(\d+)\b", tok):
        out.append((m.group(1), int(m.group(2))))
    return out


def loop_body(ops, n_rcp):
    best = None
    for addr, guard, op, rest in ops:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", rest)
            if m and int(m.group(1), 16) < addr:
                body = [o for o in ops if int(m.group(1), 16) <= o[0] <= addr]
                if sum(o[2].startswith("MUFU.RCP64H") for o in body) == n_rcp and (best is None or len(body) < len(best)):
                    best = body
    return best


def simulate(body, iters=6, in_order=True):
    ready = {}
    t = 0
    fp64_free = 0
    marks = []
    stalls = Counter()
    for it in range(iters):
        for addr, guard, op, rest in body:
            base = op.split(".")[0]
            wide = base in ("DFMA", "DMUL", "DADD", "DSETP", "MUFU") or ".64" in op or "WIDE" in op
            toks = [x.strip() for x in rest.split(",")]
            ndst = 2 if base in ("DSETP", "ISETP", "FSETP", "UISETP", "PLOP3") else 1
            if base in ("BRA", "BSSY", "BSYNC", "STG", "STS", "STL", "EXIT", "BAR", "NOP"):
                ndst = 0
            dst = [r for tk in toks[:ndst] for r in regs_of(tk, wide and base != "DSETP")]
            src = [r for tk in toks[ndst:] for r in regs_of(tk, wide and base not in ("MUFU",))]
            if base == "MUFU":  # RCP64H reads / writes the high word only
                dst = regs_of(toks[0], False)
                src = regs_of(toks[1], False)
            if guard:
                src += regs_of(guard, False)
            earliest = t + 1 if in_order else 0
            need = max([ready.get(r, 0) for r in src] + [0])
            if base in FP64 and in_order:
                need = max(need, fp64_free)
            issue = max(earliest, need)
            stalls[base] += issue - earliest
            t = issue if in_order else max(t, issue)
            if base in FP64:
                fp64_free = issue + 2
            lat = LAT.get(base, DEFAULT_LAT)
            for r in dst:
                ready[r] = issue + lat
        marks.append(t)
    return marks, stalls


if __name__ == "__main__":
    ops = parse(sys.argv[1])
    n_rcp = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    body = loop_body(ops, n_rcp)
    c = Counter(o[2].split(".")[0] for o in body)
    marks, stalls = simulate(body)
    per_iter = (marks[-1] - marks[1]) / (len(marks) - 2)
    nf = sum(c[k] for k in FP64)
    print(f"loop {hex(body[0][0])}..{hex(body[-1][0])}: {len(body)} instr, {nf} FP64-pipe; one warp in order: "
          f"{per_iter:.0f} cycles/iteration ({per_iter / len(body):.2f} cycles/instr); FP64 pipe floor {2 * nf} cycles")
    print("stall cycles by opcode (6 iterations):", dict(stalls.most_common(8)))
    marks2, _ = simulate(body, 12, in_order=False)
    print(f"dataflow limit (recurrence through the loop-carried registers, unlimited issue): "
          f"{(marks2[-1] - marks2[3]) / (len(marks2) - 4):.0f} cycles/iteration")
