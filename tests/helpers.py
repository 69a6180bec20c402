"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import json
import os

import numpy as np

from tomatis_audio_processor_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Fixtures of the Tomatis path (standard / xfade / adaptive)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
                  if not n.startswith(("eq_", "chan_", "val_", "cal_")))


def eq_golden_names():
    """Fixtures of the static-EQ processor (src/layer2_apply_eq.py)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "eq_*.npz"))))


def load_eq_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return dict(x=synth.pcm16_to_float(z["pcm16"]), out=z["out"], out_gp=(z["out_gp"] if meta["has_gp"] else None),
                gain_bins=z["gain_bins"], eq_freqs=z["eq_freqs"], eq_db=z["eq_db"], **meta)


def chan_golden_names():
    """Fixtures of the per-channel state analyser (src/analyze_stereo_state.py)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "chan_*.npz"))))


def load_chan_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return dict(x=synth.pcm16_to_float(z["pcm16"]), **meta)


def val_golden_names():
    """Fixtures of the validator kernels (src/validate_layer1.py, src/verify_tomatis_15db_v2.py)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "val_*.npz"))))


def load_val_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return dict(x=synth.pcm16_to_float(z["pcm16_x"]), y=synth.pcm16_to_float(z["pcm16_y"]), levels=z["levels"],
                states=["C1" if s == 1 else "C2" for s in z["states"]], c1_db=z["c1_db"], c2_db=z["c2_db"],
                v2_c1_db=z["v2_c1_db"], v2_c2_db=z["v2_c2_db"], **meta)


def cal_golden_names():
    """Fixtures of the calibration front end (src/calibrate_to_baseline_v2.py)."""
    return sorted(n for n in (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "cal_*.npz"))))


def load_cal_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return dict(orig=synth.pcm16_to_float(z["pcm16_orig"]), base=synth.pcm16_to_float(z["pcm16_base"]), mo_ds=z["mo_ds"],
                mb_ds=z["mb_ds"], corr=z["corr"], orig_level=z["orig_level"], base_level=z["base_level"], tilts=z["tilts"], **meta)


def cal_kwargs(words):
    """Command-line words of a calibration fixture -> keyword arguments of calibrate()."""
    kw, key = {}, None
    for w in words:
        if isinstance(w, str) and w.startswith("--"):
            key = w[2:]
            kw[key] = []
        else:
            kw[key].append(w)
    return {k: (v if k in ("hyst_list", "delay_list_ms", "tilt_lo", "tilt_hi") else v[0]) for k, v in kw.items()}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    x = synth.pcm16_to_float(z["pcm16"])
    return dict(x=x, out=z["out"], chunk_lengths=[int(v) for v in z["chunk_lengths"]], **meta)


def csv_states(rows):
    return [r[3] for r in rows[1:]]
