"""TEST INFRASTRUCTURE ONLY -- executes the UNMODIFIED reference in this container.

The reference scripts (`/root/reference/src/process_tomatis*.py`) `import soundfile`, which
is not installed here (no libsndfile).  All arithmetic on the path is NumPy's, soundfile only
moves samples, so this harness puts a tiny in-memory module named ``soundfile`` in front of
the reference and runs its ``process()`` functions as they are.  It is used ONLY

* by ``oracle/make_golden.py`` to generate the fixtures in ``tests/golden/``, and
* by ``tests/test_oracle_vs_reference.py`` (skipped when ``/root/reference`` is absent)
  to pin the NumPy restatement in ``oracle/tomatis_oracle.py`` to the reference itself.

Nothing on the product path and nothing that runs on the GPU box imports this file:
``/root/reference`` does not exist there.

The only permitted deviation from "unmodified" (SURVEY.md section 8c): for sample rates other
than 48 kHz the four guard lines ``src/process_tomatis.py:234-237`` /
``src/process_tomatis_xfade.py:106-109`` (``raise ValueError`` on sr != 48000 / ch != 2) are
dropped from the source text before ``exec``; the files on disk are never touched.
"""
from __future__ import annotations

import contextlib
import csv
import io
import os
import sys
import tempfile
import types

import numpy as np

REFERENCE_SRC = os.environ.get("TOMATIS_REFERENCE_SRC", "/root/reference/src")

MODULES = {
    "standard": "process_tomatis",
    "adaptive": "process_tomatis_adaptive",
    "xfade": "process_tomatis_xfade",
    "eq": "layer2_apply_eq",
    "stereo_state": "analyze_stereo_state",
    "validate": "validate_layer1",
    "verify_v2": "verify_tomatis_15db_v2",
    "calibrate": "calibrate_to_baseline_v2",
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "process_tomatis.py"))


class _Store:
    """In-memory 'file system': path -> (float32 array [N, ch], samplerate)."""

    def __init__(self):
        self.inputs = {}
        self.outputs = {}      # path -> dict(chunks=[arrays], sr=, channels=, format=, subtype=)


def make_soundfile_standin(store: _Store) -> types.ModuleType:
    """A module object that quacks like the parts of `soundfile` the reference calls
    (`src/process_tomatis.py:225,243,357,434`, `src/process_tomatis_adaptive.py:179,351`)."""
    mod = types.ModuleType("soundfile")

    class SoundFile:
        def __init__(self, path, mode="r", samplerate=None, channels=None, format=None, subtype=None):
            self.path, self.mode = path, mode
            if mode == "r":
                if path not in store.inputs and path in store.outputs:       # a file the reference wrote itself
                    rec = store.outputs[path]
                    store.inputs[path] = (_as_written(np.concatenate(rec["chunks"], axis=0), rec["subtype"]), rec["sr"])
                data, sr = store.inputs[path]
                self._data = data
                self.samplerate = sr
                self.channels = data.shape[1]
                self.frames = data.shape[0]
                self._pos = 0
            else:
                self.samplerate, self.channels = samplerate, channels
                store.outputs[path] = dict(chunks=[], sr=samplerate, channels=channels,
                                           format=format, subtype=subtype)

        def read(self, frames=-1, dtype="float64", always_2d=False):
            n = self.frames - self._pos if frames < 0 else min(frames, self.frames - self._pos)
            out = np.array(self._data[self._pos:self._pos + n], dtype=dtype, copy=True)
            self._pos += n
            return out

        def seek(self, frames):
            self._pos = max(0, min(int(frames), self.frames))
            return self._pos

        def write(self, data):
            store.outputs[self.path]["chunks"].append(np.array(data, copy=True))

        def close(self):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    def read(path, dtype="float64", always_2d=False):
        data, sr = store.inputs[path]
        out = np.array(data, dtype=dtype, copy=True)
        if out.shape[1] == 1 and not always_2d:
            out = out[:, 0]
        return out, sr

    def write(path, data, samplerate, subtype=None, format=None):
        store.outputs[path] = dict(chunks=[np.array(data, copy=True)], sr=samplerate,
                                   channels=(data.shape[1] if data.ndim > 1 else 1),
                                   format=format, subtype=subtype)

    mod.SoundFile = SoundFile
    mod.read = read
    mod.write = write
    mod.__version__ = "standin"
    return mod


def _as_written(y, subtype):
    """Samples as a later read(dtype='float32') returns them: libsndfile's FLAC PCM_24 rule with clipping on
    (lrint(y * 2^23) pinned to 24 bits), see audio_io.quantise_pcm24."""
    if subtype == "PCM_24":
        q = np.clip(np.rint(np.asarray(y, dtype=np.float64) * 8388608.0), -8388608, 8388607)
        return (q / 8388608.0).astype(np.float32)
    return np.asarray(y, dtype=np.float32)


_GUARD_MARKERS = ("if sr != 48000:", "if ch != 2:")


def _strip_guards(src: str) -> str:
    """Drop `if sr != 48000: raise ...` / `if ch != 2: raise ...` (2 x 2 lines)."""
    lines = src.split("\n")
    out, skip = [], 0
    for ln in lines:
        if skip:
            skip -= 1
            continue
        if ln.strip() in _GUARD_MARKERS:
            skip = 1            # the `raise ValueError(...)` line that follows
            continue
        out.append(ln)
    return "\n".join(out)


def load_reference_module(mode: str, store: _Store, skip_guard: bool = False) -> types.ModuleType:
    name = MODULES[mode]
    path = os.path.join(REFERENCE_SRC, name + ".py")
    with open(path, "r", encoding="utf-8") as f:
        src = f.read()
    if skip_guard:
        src = _strip_guards(src)
    mod = types.ModuleType("_ref_" + name)
    mod.__file__ = path
    standin = make_soundfile_standin(store)
    saved = sys.modules.get("soundfile")
    sys.modules["soundfile"] = standin
    try:
        exec(compile(src, path, "exec"), mod.__dict__)
    finally:
        if saved is None:
            sys.modules.pop("soundfile", None)
        else:
            sys.modules["soundfile"] = saved
    return mod


def run_reference(mode: str, x: np.ndarray, sr: int, want_csv: bool = True, **params) -> dict:
    """Run the reference `process()` of `mode` on float32 array x [N, ch].

    Returns dict(out=[N,ch] float array as handed to soundfile.write (pre-PCM quantisation),
    chunk_lengths=[...], csv=[rows as lists of str] or None, stdout=str, guard_skipped=bool).
    """
    assert reference_available(), "reference sources not present"
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[:, None]
    store = _Store()
    store.inputs["in.flac"] = (x, sr)
    skip_guard = mode in ("standard", "xfade") and (sr != 48000 or x.shape[1] != 2)
    mod = load_reference_module(mode, store, skip_guard=skip_guard)
    tmp = None
    if want_csv:
        fd, tmp = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        params = dict(params, state_csv_path=tmp)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            mod.process("in.flac", "out.flac", **params)
        rows = None
        if want_csv:
            with open(tmp, "r", encoding="utf-8", newline="") as f:
                rows = list(csv.reader(f))
    finally:
        if tmp and os.path.exists(tmp):
            os.unlink(tmp)
    rec = store.outputs["out.flac"]
    chunks = rec["chunks"]
    out = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, x.shape[1]), np.float32)
    return dict(out=out, chunk_lengths=[len(c) for c in chunks], csv=rows, stdout=buf.getvalue(),
                guard_skipped=skip_guard, subtype=rec["subtype"], format=rec["format"])


def run_reference_eq(x: np.ndarray, sr: int, eq_freqs, eq_db, **params) -> dict:
    """Run the reference `apply_eq_stft` (src/layer2_apply_eq.py:66) on float32 x [N, 2] with the EQ curve given as
    (freq_hz, delta_db) points.  Returns dict(out = what it wrote to the output file (float, pre-quantisation),
    out_gp = what it wrote to the gain-protected "_gp" file or None, stdout)."""
    assert reference_available(), "reference sources not present"
    x = np.ascontiguousarray(x, dtype=np.float32)
    store = _Store()
    store.inputs["in.flac"] = (x, sr)
    mod = load_reference_module("eq", store, skip_guard=False)
    fd, tmp = tempfile.mkstemp(suffix=".csv")
    os.close(fd)
    try:
        with open(tmp, "w", newline="", encoding="utf-8") as f:
            w = csv.writer(f)
            w.writerow(["freq_hz", "delta_db"])
            for a, b in zip(eq_freqs, eq_db):
                w.writerow([repr(float(a)), repr(float(b))])
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            mod.apply_eq_stft("in.flac", "out.flac", tmp, **params)
            freqs, dbs = mod.load_eq_csv(tmp)
            gain = mod.build_gain_per_bin(sr, params.get("n_fft", 4096), freqs, dbs)
    finally:
        os.unlink(tmp)
    out = np.concatenate(store.outputs["out.flac"]["chunks"], axis=0)
    gp = store.outputs.get("out_gp.flac")
    return dict(out=out, out_gp=(np.concatenate(gp["chunks"], axis=0) if gp else None), stdout=buf.getvalue(),
                gain_bins=gain, eq_freqs=freqs, eq_db=dbs)


def run_reference_stereo_state(x: np.ndarray, sr: int, **params) -> dict:
    """Run the reference `analyze()` (src/analyze_stereo_state.py:79) on float32 x [N, 2].  Returns dict(rc, rows = the
    CSV it wrote (lists of str), stdout)."""
    assert reference_available(), "reference sources not present"
    x = np.ascontiguousarray(x, dtype=np.float32)
    store = _Store()
    store.inputs["in.flac"] = (x, sr)
    mod = load_reference_module("stereo_state", store, skip_guard=False)
    fd, tmp = tempfile.mkstemp(suffix=".csv")
    os.close(fd)
    try:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            rc = mod.analyze("in.flac", tmp, **params)
        with open(tmp, "r", encoding="utf-8", newline="") as f:
            rows = list(csv.reader(f))
    finally:
        os.unlink(tmp)
    return dict(rc=rc, rows=rows, stdout=buf.getvalue())


def run_reference_validators(x: np.ndarray, y: np.ndarray, sr: int, threshold_dbfs: float, hyst_db: float, up_delay_ms: float,
                             n_fft: int = 4096, hop: int = 2048, level_threshold: float = -60, level_percentile: float = 10,
                             anchor_band=(900, 1100)) -> dict:
    """Run the reference's validator kernels on an input / output pair: simulate_gate + compute_conditional_spectrum
    (src/validate_layer1.py:110-163,261-389) and compute_conditional_spectrum_v2 (src/verify_tomatis_15db_v2.py:270-369),
    the latter on the states and levels of the former."""
    assert reference_available(), "reference sources not present"
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    store = _Store()
    v1 = load_reference_module("validate", store)
    v2 = load_reference_module("verify_v2", store)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        states, levels = v1.simulate_gate(x, sr, n_fft, hop, threshold_dbfs, hyst_db, up_delay_ms)
        freqs, c1_db, c2_db, n1, n2 = v1.compute_conditional_spectrum(x, y, sr, states, n_fft, hop, level_threshold)
        f2, a1_db, a2_db, m1, m2 = v2.compute_conditional_spectrum_v2(x, y, sr, states, np.array(levels), n_fft, hop,
                                                                     level_percentile, anchor_band)
    return dict(states=states, levels=np.array(levels), freqs=freqs, c1_db=np.asarray(c1_db), c2_db=np.asarray(c2_db),
                n_c1=n1, n_c2=n2, v2_c1_db=np.asarray(a1_db), v2_c2_db=np.asarray(a2_db), v2_n_c1=m1, v2_n_c2=m2)


def run_reference_calibration(orig: np.ndarray, base: np.ndarray, sr: int, args=()) -> dict:
    """Run the reference's calibration front end `main()` (src/calibrate_to_baseline_v2.py:130-313) on an original /
    baseline-recording pair.  `args`: extra command-line words.  Returns dict(json = the file it saved, stdout, delay)."""
    import json
    assert reference_available(), "reference sources not present"
    store = _Store()
    store.inputs["orig.flac"] = (np.ascontiguousarray(orig, dtype=np.float32), sr)
    store.inputs["base.flac"] = (np.ascontiguousarray(base, dtype=np.float32), sr)
    mod = load_reference_module("calibrate", store)
    fd, tmp = tempfile.mkstemp(suffix=".json")
    os.close(fd)
    argv = sys.argv
    buf = io.StringIO()
    try:
        sys.argv = ["calibrate_to_baseline_v2.py", "--orig", "orig.flac", "--base", "base.flac", "--sr", str(sr),
                    "--out_json", tmp] + [str(a) for a in args]
        with contextlib.redirect_stdout(buf):
            mod.main()
        with open(tmp, "r", encoding="utf-8") as f:
            out = json.load(f)
    finally:
        sys.argv = argv
        os.unlink(tmp)
    return dict(json=out, stdout=buf.getvalue(), module=mod)


def run_reference_calibration_parts(orig: np.ndarray, base: np.ndarray, sr: int, delay: int, max_minutes: float = 6.0,
                                    lo=(200, 1000), hi=(2000, 8000), ds_sr: int = 2000, chunk_sec: float = 25) -> dict:
    """The intermediate arrays main() never returns, obtained by calling the reference's own functions the way main() does
    (src/calibrate_to_baseline_v2.py:44-86 for the two decimated envelopes and their correlation, :166-196 for the frame
    levels and tilts)."""
    from scipy.signal import fftconvolve, resample_poly
    assert reference_available(), "reference sources not present"
    orig = np.ascontiguousarray(orig, dtype=np.float32)
    base = np.ascontiguousarray(base, dtype=np.float32)
    mod = load_reference_module("calibrate", _Store())
    mid, half = int(0.5 * len(base)), int(0.5 * chunk_sec * sr)
    s, e = max(0, mid - half), min(len(base), mid + half)
    mb_ds = resample_poly(mod.power_mono(base[s:e]), ds_sr, sr).astype(np.float32)
    mb_ds = mb_ds - np.mean(mb_ds)
    mo_ds = resample_poly(mod.power_mono(orig).astype(np.float32), ds_sr, sr).astype(np.float32)
    mo_ds = mo_ds - np.mean(mo_ds)
    corr = fftconvolve(mo_ds, mb_ds[::-1], mode="valid")
    base_start, orig_start = max(0, -delay), max(0, delay)
    avail = min(len(base) - base_start, len(orig) - orig_start, int(max_minutes * 60 * sr))
    xb, xo = base[base_start:base_start + avail], orig[orig_start:orig_start + avail]
    n_frames = 1 + (avail - 4096) // 2048
    ol, bl, tl = (np.zeros(n_frames, np.float32) for _ in range(3))
    for i in range(n_frames):
        st = i * 2048
        ol[i] = mod.rms_dbfs_from_mono(mod.power_mono(xo[st:st + 4096, :]))
        bl[i] = mod.rms_dbfs_from_mono(mod.power_mono(xb[st:st + 4096, :]))
        tl[i] = mod.stft_band_tilt(xb[st:st + 4096, :], sr, 4096, lo=tuple(lo), hi=tuple(hi))
    return dict(mo_ds=mo_ds, mb_ds=mb_ds, corr=corr, k=int(np.argmax(corr)), orig_level=ol, base_level=bl, tilts=tl)
