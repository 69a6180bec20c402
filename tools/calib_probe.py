"""GPU box only: stage timings of the calibration front end on a synthetic 6.5 min / 6 min pair with the
reference's default grid (13 gains x 6 delays x 6 hystereses x 81 thresholds).  Prints one JSON line."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tomatis_audio_processor_b200 import calibrate_to_baseline_v2 as cal, engine, synth  # noqa: E402

sr, d = 48000, 40123
x = synth.recipe_level_steps(390.0, sr, 7)
seg = np.ascontiguousarray(x[d:d + 360 * sr])
y = engine.run_streaming("standard", [seg], sr, gate_ui=50, up_delay_ms=200.0)[0]["out"]
base = (0.8 * y).astype(np.float32)
xo, xb = engine.to_device([x, base])


def timed(f, reps=3):
    best, out = 1e9, None
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best * 1e3, out


t_delay, dl = timed(lambda: cal.find_delay(xo, xb, sr))
t_env, _ = timed(lambda: engine.calib_envelope(xo, 0, len(x), 2000, sr))
t_xc, _ = timed(lambda: engine.calib_xcorr(dl["mo_ds"], dl["mb_ds"]))
delay = dl["delay"]
xo2, xb2 = xo[max(0, delay):max(0, delay) + 360 * sr], xb[max(0, -delay):]
n = min(len(xo2), len(xb2))
t_feat, feat = timed(lambda: cal.frame_features(xo2[:n], xb2[:n], sr))
t_all, r = timed(lambda: cal.calibrate(xo, xb, sr), reps=2)
print(json.dumps(dict(what="calibration front end, 390 s original / 360 s baseline @ 48 kHz, default grid (37 908 gate simulations)",
                      delay_found=delay, delay_true=d, ms_find_delay=round(t_delay, 2), ms_envelope_orig=round(t_env, 2),
                      ms_xcorr=round(t_xc, 2), xcorr_lags=int(dl["corr"].size), xcorr_taps=int(dl["n_base_ds"]),
                      ms_frame_features=round(t_feat, 2), frames=int(len(feat[0])), ms_calibrate_total=round(t_all, 2),
                      result=r["json"])))
