"""Side outputs of the reference's process() functions: the per-frame state CSV and the run statistics.

CSV layouts:
  standard  frame_idx,time_sec,level_dbfs,state          src/process_tomatis.py:305,408-409
  xfade     + alpha; level '%.2f', alpha '%.3f'          src/process_tomatis_xfade.py:180,293-295
  adaptive  frame_idx=i+1, time '%.6f', level '%.4f', alpha '%.4f'   src/process_tomatis_adaptive.py:355-362
Rows are produced as lists of str exactly as csv.writer would render the reference's values, so they can
be compared verbatim with the reference's files.
"""
from __future__ import annotations

import csv
from typing import List

import numpy as np

from . import tables as tb


def _state_name(s) -> str:
    return "C1" if int(s) == 1 else "C2"


def state_csv_rows(mode: str, res: dict) -> List[List[str]]:
    """res: one result dict of engine.run_streaming / engine.run_adaptive."""
    states = res["states"]
    levels = res["levels"]
    if mode == "adaptive":
        alpha = tb.alpha_follow_exact(states, res["xfade_frames"], start_at_target=True)
        rows = [["frame_idx", "time_sec", "level_dbfs", "state", "alpha"]]
        for i in range(len(states)):
            rows.append([str(i + 1), f"{res['times'][i]:.6f}", f"{float(levels[i]):.4f}", _state_name(states[i]),
                         f"{alpha[i]:.4f}"])
        return rows
    sr = res["sr"]
    starts, mask = res["frame_starts"], res["csv_mask"]
    if mode == "standard":
        rows = [["frame_idx", "time_sec", "level_dbfs", "state"]]
        for k in np.nonzero(mask)[0]:
            rows.append([str(int(k)), str(int(starts[k]) / sr), str(float(levels[k])), _state_name(states[k])])
        return rows
    alpha = tb.alpha_follow_exact(states, res["xfade_frames"], start_at_target=False)
    rows = [["frame_idx", "time_sec", "level_dbfs", "state", "alpha"]]
    for k in np.nonzero(mask)[0]:
        rows.append([str(int(k)), str(int(starts[k]) / sr), f"{float(levels[k]):.2f}", _state_name(states[k]),
                     f"{alpha[k]:.3f}"])
    return rows


def write_state_csv(path: str, mode: str, res: dict) -> None:
    with open(path, "w", newline="", encoding="utf-8") as f:
        csv.writer(f).writerows(state_csv_rows(mode, res))


def gate_statistics(states: np.ndarray, total_samples: int, sr: int, min_hold_frames: int = 0) -> dict:
    """Frame counts, switch rate and short-run ratio as the reference prints them
    (src/process_tomatis.py:461-469, src/process_tomatis_adaptive.py:228-249)."""
    st = np.asarray(states)
    n = int(st.size)
    c2 = int((st == 2).sum())
    sw = np.nonzero(st[1:] != st[:-1])[0] + 1 if n > 1 else np.zeros(0, dtype=np.int64)
    bounds = np.concatenate([[0], sw, [n]]) if n else np.zeros(1, dtype=np.int64)
    runs = np.diff(bounds)
    minutes = total_samples / sr / 60.0
    return dict(frames=n, c1_frames=n - c2, c2_frames=c2, c2_ratio=(c2 / n if n else 0.0),
                switches=int(sw.size), switches_per_min=(sw.size / minutes if minutes > 0 else 0.0),
                short_run_ratio=(float((runs < min_hold_frames).sum()) / runs.size if runs.size else 0.0))
