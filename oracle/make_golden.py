"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (container only).

    python oracle/make_golden.py            # needs /root/reference

Each fixture holds: the input (int16 PCM grid, so float32-exact), the reference's output float
array exactly as it was handed to soundfile.write (pre-PCM-24 quantisation), the chunk lengths
of those writes, the state-CSV rows the reference wrote, the kwargs, and the NumPy version.
The reference has no golden vectors of its own; these are "outputs of the
reference itself run here".  `guard_skipped` records the one permitted deviation (sr != 48 kHz
guard lines dropped for standard/xfade).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh                      # noqa: E402
from tomatis_audio_processor_b200 import synth            # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")


def _q(x):
    return synth.pcm16_to_float(synth.quantise_pcm16(x))


def cases():
    """(name, mode, sr, input float32 [N,2] on the int16 grid, kwargs)"""
    c = []
    # two limiter chunks (boundary at 239 616), default parameters, loud enough to engage the limiter
    x = synth.recipe_gated_pink(262144 / 48000, 48000, 11, env_hz=1.0, hi_dbfs=-22.0)
    c.append(("std_48k_two_chunks", "standard", 48000, _q(x), dict(gate_ui=50)))
    # linear gate map + output gain + short up-delay
    x = synth.recipe_gated_pink(1.5, 48000, 12, env_hz=2.0, lo_dbfs=-60.0, hi_dbfs=-30.0)
    c.append(("std_48k_linear_gain", "standard", 48000, _q(x),
              dict(gate_ui=58, gate_mode="linear", output_gain_db=-3.0, up_delay_ms=100.0, hysteresis_db=2.0)))
    # xfade with threshold-straddling ramps (linear map: gate_ui 50 -> T=-50)
    x = synth.recipe_threshold_ramps(2.0, 48000, 13, t_on=-48.5, t_off=-51.5, period_s=1.0)
    c.append(("xfade_48k_ramps", "xfade", 48000, _q(x), dict(gate_ui=50, xfade_ms=200.0, up_delay_ms=60.0)))
    # xfade CLI default: xfade_ms = 0 -> hard switching
    x = synth.recipe_gated_pink(1.5, 48000, 14, env_hz=2.0, lo_dbfs=-65.0, hi_dbfs=-35.0, bursts=False)
    c.append(("xfade_48k_hard", "xfade", 48000, _q(x), dict(gate_ui=50, up_delay_ms=50.0)))
    # adaptive, float32 branch (pre-attenuation active)
    x = synth.recipe_swept_pink(2.0, 48000, 15, period_s=0.9, peak=0.5)
    c.append(("adaptive_48k_f32", "adaptive", 48000, _q(x), dict(min_hold_ms=100.0, xfade_ms=200.0)))
    # adaptive, float64 branch (input peak <= 0.141 -> atten_db is int 0)
    x = synth.recipe_swept_pink(1.5, 48000, 16, period_s=0.7, peak=0.1)
    c.append(("adaptive_48k_f64", "adaptive", 48000, _q(x), dict()))
    # 44.1 kHz standard (guard skipped), ragged length
    x = synth.recipe_gated_pink(1.5, 44100, 17, env_hz=2.0, hi_dbfs=-28.0)
    c.append(("std_44k1", "standard", 44100, _q(x), dict(gate_ui=50, up_delay_ms=120.0)))
    # 96 kHz xfade, total % hop == 0 -> pad_end == 0 (degenerate tail)
    x = synth.recipe_threshold_ramps(2048 * 40 / 96000, 96000, 18, t_on=-38.5, t_off=-41.5, period_s=0.4)
    c.append(("xfade_96k_padend0", "xfade", 96000, _q(x)[:2048 * 40], dict(gate_ui=60, xfade_ms=100.0, up_delay_ms=40.0)))
    # adaptive 44.1 kHz, total % hop == hop-1 (tail reaches the window end)
    x = synth.recipe_swept_pink(2.0, 44100, 19, period_s=0.8, peak=0.4)
    c.append(("adaptive_44k1_tail", "adaptive", 44100, _q(x)[:2048 * 30 + 2047], dict(min_hold_ms=90.0, xfade_ms=150.0)))
    c.extend(multichannel_cases())
    c.extend(small_frame_cases())
    return c


def small_frame_cases():
    """--n_fft 2048 --hop 1024, the documentation's faster setting (docs/Tomatis技术说明.md:253-258), which the fused kernels serve
    through the second build of the library: the reference itself at these sizes, all three modes."""
    sz = dict(n_fft=2048, hop=1024)
    c = []
    # two limiter chunks at hop 1024 (first flush after 237 frames = 240 640 samples)
    x = synth.recipe_gated_pink(5.4, 48000, 51, env_hz=1.2, hi_dbfs=-21.0)
    c.append(("std_48k_n2048_two_chunks", "standard", 48000, _q(x), dict(gate_ui=50, up_delay_ms=120.0, **sz)))
    x = synth.recipe_threshold_ramps(2.0, 44100, 52, t_on=-48.5, t_off=-51.5, period_s=0.8)
    c.append(("xfade_44k1_n2048", "xfade", 44100, _q(x)[:44100 * 2 - 333], dict(gate_ui=50, xfade_ms=120.0, up_delay_ms=60.0, **sz)))
    x = synth.recipe_swept_pink(2.0, 48000, 53, period_s=0.6, peak=0.5)
    c.append(("adaptive_48k_n2048", "adaptive", 48000, _q(x), dict(min_hold_ms=100.0, xfade_ms=200.0, **sz)))
    return c


def multichannel_cases():
    """adaptive mode on files with more than two channels (the reference loops over channels, _adaptive.py:307-313)."""
    c = []
    # three channels (odd: the last one has no partner), float32 branch; the channels differ in level so that the all-channel
    # level, the input peak and the output peak each come from a different place
    a = synth.recipe_swept_pink(1.6, 48000, 31, period_s=0.7, peak=0.5)
    b = synth.recipe_swept_pink(1.6, 48000, 32, period_s=0.9, peak=0.3)
    c.append(("adaptive_48k_3ch", "adaptive", 48000, _q(np.concatenate([a, b[:, :1]], axis=1)), dict(min_hold_ms=100.0, xfade_ms=200.0)))
    # 5.1 layout, float64 branch (input peak <= 0.141), 44.1 kHz, ragged length
    parts = [synth.recipe_swept_pink(1.4, 44100, 33 + k, period_s=0.5 + 0.1 * k, peak=0.1 - 0.02 * k) for k in range(3)]
    c.append(("adaptive_44k1_6ch_f64", "adaptive", 44100, _q(np.concatenate(parts, axis=1))[:2048 * 25 + 777], dict(xfade_ms=300.0)))
    # nine channels: NumPy's mean over the channel axis switches to its 8-accumulator pairwise order from 8 channels on
    parts = [synth.recipe_swept_pink(1.2, 48000, 40 + k, period_s=0.45 + 0.05 * k, peak=0.45 - 0.05 * k) for k in range(5)]
    c.append(("adaptive_48k_9ch", "adaptive", 48000, _q(np.concatenate(parts, axis=1)[:, :9]), dict(min_hold_ms=80.0, xfade_ms=150.0)))
    return c


EQ_CURVE = ([20.0, 100.0, 500.0, 1000.0, 4000.0, 12000.0, 20000.0], [6.0, 4.0, 0.0, -2.0, 3.0, 8.0, 10.0])


def eq_cases():
    """(name, input, kwargs) for the static-EQ processor src/layer2_apply_eq.py."""
    x = _q(synth.recipe_gated_pink(1.0, 48000, 71, env_hz=2.0, hi_dbfs=-14.0))
    return [("eq_48k_pad_gainprotect", x, dict()),
            ("eq_48k_nopad_gain", x[:40000], dict(pad=False, global_gain_db=-3.0, auto_gain_protect=False))]


def make_eq():
    for name, x, kw in eq_cases():
        r = rh.run_reference_eq(x, 48000, EQ_CURVE[0], EQ_CURVE[1], **kw)
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(
            path, pcm16=synth.quantise_pcm16(x), out=r["out"],
            out_gp=(r["out_gp"] if r["out_gp"] is not None else np.zeros((0, 2), np.float32)),
            gain_bins=r["gain_bins"], eq_freqs=r["eq_freqs"], eq_db=r["eq_db"],
            meta=np.array(json.dumps(dict(name=name, mode="eq", sr=48000, kwargs=kw, has_gp=r["out_gp"] is not None,
                                          numpy=np.__version__, reference_files=["layer2_apply_eq.py"]))))
        print(f"{name:24s} eq        N={len(x)} out={r['out'].shape} peak={np.abs(r['out']).max():.4f} "
              f"gp={'yes' if r['out_gp'] is not None else 'no'} {os.path.getsize(path)/1e6:.2f} MB")


def chan_cases():
    """(name, sr, input, kwargs) for the per-channel state analyser src/analyze_stereo_state.py.
    The right channel is a delayed, quieter copy so that the two channels get different thresholds and states."""
    def two(x, shift, gain):
        y = x.copy()
        y[:, 1] = np.roll(x[:, 1], shift) * gain
        return _q(y)
    return [("chan_48k_default", 48000, two(synth.recipe_swept_pink(3.0, 48000, 31, period_s=0.9, peak=0.5), 5000, 0.6), dict()),
            ("chan_44k1_target30", 44100, two(synth.recipe_gated_pink(2.5, 44100, 32, env_hz=1.5, hi_dbfs=-28.0), 7000, 0.5),
             dict(target_c2=0.3, hyst_db=2.0, min_hold_ms=100.0)),
            ("chan_96k_ragged", 96000, two(synth.recipe_swept_pink(2.0, 96000, 33, period_s=0.5, peak=0.1), 9000, 0.7)[:2048 * 60 + 777],
             dict(min_hold_ms=60.0))]


def make_chan():
    for name, sr, x, kw in chan_cases():
        r = rh.run_reference_stereo_state(x, sr, **kw)
        assert r["rc"] == 0
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(path, pcm16=synth.quantise_pcm16(x), meta=np.array(json.dumps(dict(
            name=name, mode="stereo_state", sr=sr, kwargs=kw, csv=r["rows"], stdout=r["stdout"], numpy=np.__version__,
            reference_files=["analyze_stereo_state.py"]))))
        st = [row[4] + row[6] for row in r["rows"][1:]]
        print(f"{name:24s} channels  sr={sr} N={len(x)} frames={len(st)} L-C2={sum(s[:2] == 'C2' for s in st)} "
              f"R-C2={sum(s[2:] == 'C2' for s in st)} {os.path.getsize(path)/1e6:.2f} MB")


def val_cases():
    """(name, sr, x, y, kwargs) for the validator kernels.  y is the oracle's standard-mode output of x
    on the int16 grid: to the validators it is just a second input file."""
    from oracle import tomatis_oracle as orc
    out = []
    for name, sr, seed, secs, kw in (("val_48k_default", 48000, 51, 5.0, dict(threshold_dbfs=-40.0, hyst_db=3.0, up_delay_ms=250.0)),
                                     ("val_44k1_anchor", 44100, 52, 4.0, dict(threshold_dbfs=-40.0, hyst_db=2.0, up_delay_ms=100.0,
                                                                             level_percentile=20, anchor_band=(800, 1250)))):
        x = _q(synth.recipe_gated_pink(secs, sr, seed, env_hz=0.6, hi_dbfs=-24.0))
        y = _q(orc.run("standard", x, sr, gate_ui=50, hysteresis_db=kw["hyst_db"], up_delay_ms=kw["up_delay_ms"])["out"])
        out.append((name, sr, x, y, kw))
    return out


def make_val():
    for name, sr, x, y, kw in val_cases():
        r = rh.run_reference_validators(x, y, sr, **kw)
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(
            path, pcm16_x=synth.quantise_pcm16(x), pcm16_y=synth.quantise_pcm16(y), levels=r["levels"],
            states=np.array([1 if s == "C1" else 2 for s in r["states"]], dtype=np.uint8),
            c1_db=r["c1_db"], c2_db=r["c2_db"], v2_c1_db=r["v2_c1_db"], v2_c2_db=r["v2_c2_db"],
            meta=np.array(json.dumps(dict(name=name, mode="validators", sr=sr, kwargs=kw, n_c1=r["n_c1"], n_c2=r["n_c2"],
                                          v2_n_c1=r["v2_n_c1"], v2_n_c2=r["v2_n_c2"], numpy=np.__version__,
                                          reference_files=["validate_layer1.py", "verify_tomatis_15db_v2.py"]))))
        print(f"{name:24s} validate  sr={sr} N={len(x)} frames={len(r['states'])} stable C1/C2 used {r['n_c1']}/{r['n_c2']} "
              f"(v2 {r['v2_n_c1']}/{r['v2_n_c2']}) {os.path.getsize(path)/1e6:.2f} MB")


def cal_cases():
    """(name, sr, orig, base, true delay, command-line words) for the calibration front end
    src/calibrate_to_baseline_v2.py.  The baseline stands in for the hardware recording: the oracle's
    standard-mode output of a delayed stretch of the original, attenuated, with a little noise, on the int16 grid."""
    from oracle import tomatis_oracle as orc
    out = []
    for name, sr, seed, secs_o, secs_b, d, proc, words in (
            ("cal_48k_default", 48000, 91, 11.0, 8.0, 33210, dict(gate_ui=50, up_delay_ms=100.0, hysteresis_db=2.0),
             ["--hyst_list", 0, 2, 4, "--delay_list_ms", 0, 100, 200]),
            ("cal_44k1_bands", 44100, 92, 9.0, 6.5, 21007, dict(gate_ui=52, up_delay_ms=50.0, hysteresis_db=4.0),
             ["--hyst_list", 2, 4, 6, "--delay_list_ms", 50, 150, "--tilt_lo", 150, 800, "--tilt_hi", 2500, 9000,
              "--tilt_medfilt", 3, "--music_dbfs", -60, "--gain_step_db", 1.0, "--T_step_db", 0.5])):
        x = _q(synth.recipe_level_steps(secs_o, sr, seed, min_s=0.25, max_s=0.9))
        seg = x[d:d + int(secs_b * sr)]
        y = orc.run("standard", seg, sr, **proc)["out"].astype(np.float32)
        rng = np.random.default_rng(seed + 1000)
        base = _q(0.8 * y + 1e-4 * rng.standard_normal(y.shape).astype(np.float32))
        out.append((name, sr, x, base, d, words))
    return out


def make_cal():
    for name, sr, x, base, d, words in cal_cases():
        r = rh.run_reference_calibration(x, base, sr, words)
        lo = tuple(words[words.index("--tilt_lo") + 1:words.index("--tilt_lo") + 3]) if "--tilt_lo" in words else (200, 1000)
        hi = tuple(words[words.index("--tilt_hi") + 1:words.index("--tilt_hi") + 3]) if "--tilt_hi" in words else (2000, 8000)
        parts = rh.run_reference_calibration_parts(x, base, sr, r["json"]["delay_samples_orig_minus_base"], lo=lo, hi=hi)
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(
            path, pcm16_orig=synth.quantise_pcm16(x), pcm16_base=synth.quantise_pcm16(base), mo_ds=parts["mo_ds"],
            mb_ds=parts["mb_ds"], corr=parts["corr"].astype(np.float32), orig_level=parts["orig_level"],
            base_level=parts["base_level"], tilts=parts["tilts"],
            meta=np.array(json.dumps(dict(name=name, mode="calibrate", sr=sr, words=words, true_delay=d, k=parts["k"],
                                          json={k: v for k, v in r["json"].items() if k not in ("orig", "base")},
                                          stdout=r["stdout"], tilt_lo=lo, tilt_hi=hi, numpy=np.__version__,
                                          scipy=__import__("scipy").__version__,
                                          reference_files=["calibrate_to_baseline_v2.py"]))))
        j = r["json"]
        print(f"{name:24s} calibrate sr={sr} N={len(x)}/{len(base)} delay={j['delay_samples_orig_minus_base']} (true {d}) "
              f"T_raw={j['T_raw_dbfs']:.2f} hyst={j['hyst_db']} up={j['up_delay_ms']} mismatch={j['mismatch']:.4f} "
              f"{os.path.getsize(path)/1e6:.2f} MB")


def main():
    assert rh.reference_available(), "run in the build container (needs /root/reference)"
    os.makedirs(OUT_DIR, exist_ok=True)
    if "--only-cal" in sys.argv:
        return make_cal()
    if "--only-chan" in sys.argv:
        return make_chan()
    if "--only-val" in sys.argv:
        return make_val()
    if "--only-eq" in sys.argv:
        return make_eq()
    only_multi = "--only-multichannel" in sys.argv
    only_small = "--only-small-frames" in sys.argv
    if not only_multi and not only_small:
        make_chan()
        make_val()
        make_cal()
        make_eq()
    for name, mode, sr, x, kw in (multichannel_cases() if only_multi else small_frame_cases() if only_small else cases()):
        r = rh.run_reference(mode, x, sr, **kw)
        q = synth.quantise_pcm16(x)
        assert np.array_equal(synth.pcm16_to_float(q), x)
        path = os.path.join(OUT_DIR, name + ".npz")
        np.savez_compressed(
            path,
            pcm16=q,
            out=r["out"],
            chunk_lengths=np.array(r["chunk_lengths"], dtype=np.int64),
            meta=np.array(json.dumps(dict(
                name=name, mode=mode, sr=sr, kwargs=kw, csv=r["csv"], guard_skipped=r["guard_skipped"],
                out_dtype=str(r["out"].dtype), numpy=np.__version__,
                reference_files=[rh.MODULES[mode] + ".py"]))),
        )
        states = [row[3] for row in r["csv"][1:]]
        print(f"{name:24s} {mode:9s} sr={sr} N={len(x)} out={r['out'].dtype} chunks={r['chunk_lengths']} "
              f"frames={len(states)} C2={states.count('C2')}/{len(states)} peak={np.abs(r['out']).max():.4f} "
              f"{os.path.getsize(path)/1e6:.2f} MB")


if __name__ == "__main__":
    main()
