"""The reporting tools run on the CPU as well: the parity report with the oracle in place of the CUDA path (the checker
against the reference's fixtures: every difference is zero, every bar met), and the tensor-core feasibility model."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable] + list(args), cwd=ROOT, capture_output=True, text=True, timeout=600)


def test_parity_report_on_the_oracle():
    p = _run("tools/parity_report.py", "--impl", "oracle")
    assert p.returncode == 0, p.stdout + p.stderr
    rows = [line for line in p.stdout.splitlines() if line.startswith("| ") and "fixture" not in line and "---" not in line]
    assert len(rows) >= 9 and "ALL MET" in p.stdout
    for line in rows:
        cells = [c.strip() for c in line.strip("|").split("|")]
        assert cells[4] == "0" and cells[5] == "0" and cells[6] == "equal"          # mismatches, near-threshold frames, chunks
        assert float(cells[7]) == 0.0 and float(cells[10]) == 0.0 and cells[11] == "-inf"


def test_tensor_core_feasibility_model():
    p = _run("tools/experiments/tc_fft_feasibility.py")
    assert p.returncode == 0, p.stdout + p.stderr
    verdict = {line[:58].strip(): line.rstrip().split()[-1] for line in p.stdout.splitlines()[1:] if line.strip()}
    assert verdict["tf32 x1 (plain TF32 GEMM)"] == "FAILS" and verdict["bf16 x3"] == "FAILS"
    assert verdict["tf32 x3 (Fh*xh + Fh*xl + Fl*xh)"] == "ok" and verdict["fp16 x3, per-frame scale"] == "ok"
    assert verdict["float32 butterflies (what stft_kernel does now)"] == "ok"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys, timing
    the oracle port on the host cores -- runs without a GPU."""
    import json
    p = _run("bench.py", "--impl", "reference", "--steps", "1", "--warmup", "3", "--cpu-tracks", "2", "--cpu-workers", "2")
    assert p.returncode == 0, p.stdout + p.stderr
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio_seconds_per_second" and line["unit"] == "audio-s/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 1 and line["warmup"] >= 3 and line["n_gpus"] == 1
    assert "workload" in line["config"] and "model" not in line["config"]
    cb, e2e = line["cpu_baseline"], line["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["unit"] == line["unit"] and cb["sample"]
    assert e2e == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
