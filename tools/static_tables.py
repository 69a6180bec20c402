"""Static evidence of a build of the library (here, no GPU): python tools/static_tables.py [4096|2048] > profiles/rNN/static_nXXXX.md
ptxas -v table (registers, spills, static shared memory per kernel) and SASS mnemonic counts of stft_kernel and of the whole library."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "tomatis_audio_processor_b200", "csrc")


def main():
    n_fft = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    out = "/tmp/static_tables_%d.so" % n_fft
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
           "-Xptxas", "-v", f"-DTMT_NFFT={n_fft}", "-o", out, "tomatis_b200.cu"]
    log = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True).stderr
    print(f"# Static evidence of the n_fft = {n_fft} / hop = {n_fft // 2} build (sm_100a, `-DTMT_NFFT={n_fft}`)\n")
    print("## ptxas -v\n\n| kernel | registers | spill st/ld (B) | static smem (B) |\n|---|---|---|---|")
    name = None
    for ln in log.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)[:100]
            spill = "0/0"
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", ln)
        if m:
            spill = f"{m.group(1)}/{m.group(2)}"
        m = re.search(r"Used (\d+) registers(?:.*?(\d+) bytes smem)?", ln)
        if m and name:
            print(f"| {name} | {m.group(1)} | {spill} | {m.group(2) or 0} |")
            name = None
    sass = subprocess.run(["cuobjdump", "-sass", out], capture_output=True, text=True).stdout
    allc, kc, on = collections.Counter(), collections.Counter(), False
    for ln in sass.splitlines():
        if "Function :" in ln:
            on = "stft_kernel" in ln
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Za-z0-9_]+)*)", ln)
        if m:
            mn = m.group(1).split(".")[0]
            allc[mn] += 1
            if on:
                kc[mn] += 1
    for title, c in (("stft_kernel", kc), ("whole library", allc)):
        print(f"\n## SASS mnemonic counts, {title} (top 30; LDTM / STTM = tcgen05.ld / st, FFMA2 / FMUL2 / FADD2 = packed FP32x2)\n\n```")
        for k, v in c.most_common(30):
            print(f"{v:7d} {k}")
        print("```")


if __name__ == "__main__":
    main()
