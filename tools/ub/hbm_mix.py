"""HBM bandwidth by read/write mix on this GPU (torch ops over 2 GiB buffers, CUDA events, best of 10): pure write (fill_), pure
read (sum), 1:1 copy, and the 1:3 read:write mix of the QKV projection (one fp16 read, three fp16 writes via three copies of a
quarter-size source).  Used to put write-heavy streaming kernels (gemm_stream) against the right roof."""
import torch

def best(fn, n=10):
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)

N = 1 << 30  # fp16 elements = 2 GiB
a = torch.empty(N, dtype=torch.float16, device="cuda").normal_()
b = torch.empty(N, dtype=torch.float16, device="cuda")
for _ in range(2): b.copy_(a); b.fill_(1.0); a.view(torch.int32).sum()
t = best(lambda: b.fill_(1.0)); print(f"write only (fill_)      {2*N/t/1e9:8.1f} GB/s")
t = best(lambda: a.view(torch.int32).sum()); print(f"read only (sum)         {2*N/t/1e9:8.1f} GB/s")
t = best(lambda: b.copy_(a)); print(f"copy 1:1                {4*N/t/1e9:8.1f} GB/s")
q = N // 4
def mix():
    src = a[:q]
    b[:q].copy_(src); b[q:2*q].copy_(src); b[2*q:3*q].copy_(src)
t = best(mix); print(f"3 writes per read (src re-read from L2/HBM) {(2*q*3 + 2*q)/t/1e9:8.1f} GB/s algorithmic")
