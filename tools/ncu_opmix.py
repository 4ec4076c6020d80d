"""Per-opcode dynamic instruction mix of a kernel from an ncu report (source page)."""
import collections, csv, subprocess, sys
rep, regex, rollout_steps = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{regex}:1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, samp, totI = collections.Counter(), collections.Counter(), 0
for r in data:
    op = r[iS].strip().split()
    o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    n = int(r[iE]); totI += n; tot[o] += n; samp[o] += int(r[iSamp])
per = rollout_steps / 32
print(f"total warp instr {totI}  thread-instr per rollout-step {totI/per:.1f}  static {len(data)}")
fp64 = sum(tot[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
print(f"FP64-pipe instr per rollout-step {fp64/per:.1f} ({100*fp64/totI:.1f}%)")
for o, n in tot.most_common(int(sys.argv[4]) if len(sys.argv) > 4 else 25):
    print(f"{o:10s} {n/per:8.1f}/rollout-step {100*n/totI:5.1f}%  stall-samples {samp[o]}")
