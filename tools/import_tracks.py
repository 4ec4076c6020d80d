"""Imports the centre-line geometry of the reference's track CSVs (x,y per row, no header;
src/envs/car_racing_tracks/*.csv) into mpopis_b200/data/tracks.json so that the default
`Track()` of the Python host mirror works without a checkout of the reference (the GPU box has
none). Run in the build container: python tools/import_tracks.py [/root/reference]
Only numeric track DATA is imported; no reference source code is copied."""
import json
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
src = ref / "src" / "envs" / "car_racing_tracks"
out = Path(__file__).resolve().parents[1] / "mpopis_b200" / "data" / "tracks.json"
tracks = {}
for csv in sorted(src.glob("*.csv")):
    xs, ys = [], []
    for line in csv.read_text().splitlines():
        line = line.strip()
        if not line:
            continue
        a, b = line.split(",")
        xs.append(float(a))
        ys.append(float(b))
    tracks[csv.stem] = {"x": xs, "y": ys}
out.write_text(json.dumps(tracks, separators=(",", ":")))
print({k: len(v["x"]) for k, v in tracks.items()}, "->", out, out.stat().st_size, "bytes")
