"""Dev probe: what PCIe gives on this box -- H2D alone, D2H alone, both at once (pinned memory, 847 MB chunks like the
host-buffer pipeline's waves)."""
import torch
n = 847 * 1024 * 1024 // 4
h_in = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(4)]
h_out = [torch.empty(n, dtype=torch.float32, pin_memory=True) for _ in range(4)]
d_in = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
d_out = [torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run(h2d, d2h):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    for i in range(4):
        if h2d:
            with torch.cuda.stream(s1):
                d_in[i].copy_(h_in[i], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i].copy_(d_out[i], non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    b.record(); torch.cuda.synchronize()
    return 4 * n * 4 / (a.elapsed_time(b) * 1e-3) / 1e9

for _ in range(2):
    print("H2D alone %.1f GB/s | D2H alone %.1f GB/s | both: %.1f GB/s each direction" % (run(True, False), run(False, True), run(True, True)))
