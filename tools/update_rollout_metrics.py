"""Stamp profiles/rollout_kernel_metrics.json from an `ncu --set full --import-source on` capture of a rollout kernel.

    python tools/update_rollout_metrics.py <capture.ncu-rep> <kernel-regex> <variant> <K> [T=50]

bench.py's `roofline_fp64` multiplies FP64-pipe instructions per rollout-step (measured here, per kernel variant) by
the live launch rate. The file is stamped with a hash of the rollout kernel sources (bench.rollout_source_hash): when a
kernel changes and the capture is not redone, bench.py reports the roofline as stale instead of silently wrong."""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import rollout_source_hash  # noqa: E402

rep, regex, variant, K = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
T = int(sys.argv[5]) if len(sys.argv) > 5 else 50
rollout_steps = K * T

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
sel = [r for r in rows[2:] if len(r) == len(hdr) and __import__("re").search(regex, r[hdr.index("Kernel Name")])]
assert sel, f"no kernel matching {regex} in {rep}"
r0 = sel[0]


def metric(name, default=None):
    return float(r0[hdr.index(name)].replace(",", "")) if name in hdr and r0[hdr.index(name)] not in ("", "n/a") else default


def to_bytes(name):
    v, unit = metric(name), rows[1][hdr.index(name)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(name):
    v, unit = metric(name), rows[1][hdr.index(name)]
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)


src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{regex}:1"],
                     capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
shdr = next(r for r in srows if "Source" in r and "Instructions Executed" in r)
data = [r for r in srows if len(r) == len(shdr) and r[0].startswith("0x")]
iS, iE = shdr.index("Source"), shdr.index("Instructions Executed")
tot = collections.Counter()
for r in data:
    op = r[iS].strip().split()
    tot[(op[1] if op[0].startswith("@") else op[0]).split(".")[0]] += int(r[iE])
warp_instr = sum(tot.values())
fp64 = sum(tot[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
per = rollout_steps / 32.0

entry = {
    "kernel": r0[hdr.index("Kernel Name")][:120],
    "capture": Path(rep).name,
    "K": K, "T": T,
    "gpu_time_us": to_us("gpu__time_duration.sum"),
    "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
    "algorithmic_bytes_per_launch": (16 + 8.0 / T) * rollout_steps,
    "thread_instr_per_rollout_step": warp_instr / per,
    "fp64_thread_instr_per_rollout_step": fp64 / per,
    "fp64_pipe_pct_of_peak_active": metric("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": metric("launch__registers_per_thread"),
    "warps_active_pct": metric("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "opcode_mix_per_rollout_step": {o: round(n / per, 1) for o, n in tot.most_common(14)},
}
out = ROOT / "profiles" / "rollout_kernel_metrics.json"
doc = json.loads(out.read_text()) if out.exists() else {}
if "variants" not in doc or doc.get("rollout_source_sha256_16") != rollout_source_hash():
    doc = {"rollout_source_sha256_16": rollout_source_hash(), "variants": {}}  # other variants' entries are stale too
doc["variants"][str(variant)] = entry
out.write_text(json.dumps(doc, indent=1) + "\n")
print(json.dumps(entry, indent=1))
