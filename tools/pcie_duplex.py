"""Do a host->device and a device->host copy of pinned buffers overlap on this box?  (bench.py's pipelined e2e loop relies on it.)"""
import json, time, torch
n = 4 << 30
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down):
    torch.cuda.synchronize(); t = time.perf_counter()
    if up:
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    if down:
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
for _ in range(2): run(True, True)
r = {"h2d_GBps": n / run(True, False) / 1e9, "d2h_GBps": n / run(False, True) / 1e9}
t = run(True, True); r["both_s"] = t; r["both_GBps_each"] = n / t / 1e9
# the same with many smaller chunks per direction
def run_chunks(k):
    torch.cuda.synchronize(); t = time.perf_counter(); c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_in[i * c:(i + 1) * c].copy_(h_in[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2): h_out[i * c:(i + 1) * c].copy_(d_out[i * c:(i + 1) * c], non_blocking=True)
    torch.cuda.synchronize(); return time.perf_counter() - t
r["both_64_chunks_GBps_each"] = n / run_chunks(64) / 1e9
print(json.dumps(r))
