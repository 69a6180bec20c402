"""Static instruction counts of stft_kernel's frame loop from the built library (here, no GPU): python tools/sass_count.py [--json out].

The frame loop is the innermost backward branch around at least 400 packed FP32 instructions (since round 2's steady-state
specialisation: the steady loop, which has no clipped-output variant).  Packed FP32 instructions (FFMA2 / FMUL2 / FADD2,
two lanes each) inside it, minus the duplicated output variant (whole block / clipped block: 8 FFMA2 each, one executes), give
the per-thread, per-frame count the bench's secondary (FP32) roofline uses."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tomatis_audio_processor_b200", "csrc", "libtomatis_b200.so")


def kernel_sass(name="stft_kernel"):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    lines, on = [], False
    for ln in out.splitlines():
        if "Function :" in ln:
            on = name in ln
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if on and m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
    return lines


def main():
    sass = kernel_sass()
    addr = [a for a, _ in sass]
    packed = [a for a, t in sass if re.search(r"\b(FFMA2|FMUL2|FADD2)\b", t)]
    loops = []
    for a, t in sass:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            npk = sum(tgt <= w <= a for w in packed)
            if tgt < a and npk >= 400:                                         # a backward branch around (at least) a whole frame
                loops.append((npk / sum(tgt <= x <= a for x in addr), tgt, a, npk))
    # the steady-state loop is the one with the highest share of packed FP32 instructions (the general loop carries the clipped
    # and bounds-checked variants); it may hold several frames per iteration
    _, lo, hi, npk = max(loops)
    loop = (lo, hi)
    frames = max(1, round(npk / 607.0))
    body = [t for a, t in sass if loop[0] <= a <= loop[1]]
    def cnt(pat, where):
        return sum(1 for t in where if re.search(pat, t))
    res = {"kernel_instructions": len(sass), "loop_instructions": len(body), "frames_per_iteration": frames}
    for k, pat in (("FFMA2", r"\bFFMA2\b"), ("FMUL2", r"\bFMUL2\b"), ("FADD2", r"\bFADD2\b"), ("FFMA", r"\bFFMA\b"), ("FMUL", r"\bFMUL\b"),
                   ("FADD", r"\bFADD\b"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("LDG", r"\bLDG"),
                   ("STG", r"\bSTG"), ("LDL", r"\bLDL"), ("STL", r"\bSTL"), ("MOV", r"\bMOV\b"), ("SYNCS", r"\bSYNCS"), ("BAR", r"\bBAR\b")):
        res[k] = cnt(pat, body)
    dup_out = 8 if res["STG"] > 8 * frames else 0       # the clipped-block variant of the output FFMA2s (absent from the steady-state loop)
    ffma2 = (res["FFMA2"] - dup_out) / frames
    res["instructions_per_thread_frame"] = len(body) / frames
    res["packed_fp32_per_thread_frame"] = ffma2 + (res["FMUL2"] + res["FADD2"]) / frames
    res["flop_per_thread_frame"] = 4 * ffma2 + (2 * (res["FMUL2"] + res["FADD2"]) + 2 * res["FFMA"] + res["FMUL"] + res["FADD"]) / frames
    res["flop_per_sample_frame"] = res["flop_per_thread_frame"] * 256 / 2048
    res["lane_ops_per_sample_frame"] = 2 * res["packed_fp32_per_thread_frame"] * 256 / 2048
    res["source"] = ("static SASS count of stft_kernel's frame loop (tools/sass_count.py on the built library): packed FP32 "
                     "instructions per thread and frame, 256 threads per frame, 2048 new sample-frames per frame")
    print(json.dumps(res, indent=1))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
