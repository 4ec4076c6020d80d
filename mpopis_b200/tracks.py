"""Track geometry — host mirror of src/envs/car_racing_tracks/car_racing_tracks.jl (TRK).

Only what sits either side of the hot path lives here: loading a centre line (CSV with two
columns, no header: TRK:14-16, or one of the bundled tracks) and sub-sampling it by
`sample_factor` (TRK:21-23). The nearest-point search `within_track` itself (TRK:68-92) runs
on the GPU (csrc/rollout.cu).
"""
from __future__ import annotations

import json
from functools import lru_cache
from pathlib import Path

import numpy as np

_DATA = Path(__file__).resolve().parent / "data" / "tracks.json"


@lru_cache(maxsize=1)
def _bundled() -> dict:
    return json.loads(_DATA.read_text())


def bundled_track_names() -> list[str]:
    return sorted(_bundled())


class Track:
    """struct Track (TRK:2-12): x, y, lane_width and their sub-sampled x′, y′, lane_width′."""

    def __init__(self, infile: str = "curve", width=15.0, sample_factor: int = 20):
        name = Path(str(infile)).stem
        if Path(str(infile)).suffix == ".csv" and Path(infile).exists():
            rows = [ln.split(",") for ln in Path(infile).read_text().splitlines() if ln.strip()]
            if any(len(r) != 2 for r in rows):
                raise ValueError("Can only have 2 columns for a track file")  # TRK:16
            x = np.array([float(r[0]) for r in rows])
            y = np.array([float(r[1]) for r in rows])
        elif name in _bundled():
            x = np.array(_bundled()[name]["x"], dtype=np.float64)
            y = np.array(_bundled()[name]["y"], dtype=np.float64)
        else:
            raise FileNotFoundError(f"track {infile!r} is neither a CSV file nor one of {bundled_track_names()}")
        if np.isscalar(width):
            lane_width = np.ones(x.size) * float(width)  # TRK:30-33
        else:
            lane_width = np.asarray(width, dtype=np.float64)
            if lane_width.size != x.size:
                raise ValueError("Supplied width vector does not match length of track file")  # TRK:17
        self.x, self.y, self.lane_width = x, y, lane_width
        self.sample_factor = int(sample_factor)
        self.xs = np.ascontiguousarray(x[:: self.sample_factor])  # x′ = x[1:sample_factor:end], TRK:21
        self.ys = np.ascontiguousarray(y[:: self.sample_factor])
        self.ws = np.ascontiguousarray(lane_width[:: self.sample_factor])
        self.name = name
