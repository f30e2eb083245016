"""Pipeline probe: the copy/compute pattern of bf_eval (chunked H2D -> kernel -> D2H on three streams) replayed with torch,
to see what overlap this box's PCIe path gives for a 14.4 MB up / 8.0 MB down call."""
import time
import torch

up, down = 14_400_000, 8_000_000
h_in = torch.empty(up, dtype=torch.uint8).pin_memory()
h_out = torch.empty(down, dtype=torch.uint8).pin_memory()
d_in = torch.empty(up, dtype=torch.uint8, device="cuda")
d_out = torch.empty(down, dtype=torch.uint8, device="cuda")
work = torch.empty(8 << 20, dtype=torch.uint8, device="cuda")
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def call(chunks):
    for k in range(chunks):
        a, b = up * k // chunks, up * (k + 1) // chunks
        c, d = down * k // chunks, down * (k + 1) // chunks
        with torch.cuda.stream(s_in):
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
            e1 = torch.cuda.Event(); e1.record(s_in)
        with torch.cuda.stream(s_k):
            s_k.wait_event(e1)
            work[: (8 << 20) // chunks].fill_(1)          # stands in for the pass (~15 us per quarter)
            e2 = torch.cuda.Event(); e2.record(s_k)
        with torch.cuda.stream(s_out):
            s_out.wait_event(e2)
            h_out[c:d].copy_(d_out[c:d], non_blocking=True)
    torch.cuda.synchronize()


for chunks in (1, 2, 4, 8):
    for _ in range(5):
        call(chunks)
    t0 = time.perf_counter()
    for _ in range(40):
        call(chunks)
    print("chunks %d: %.3f ms per call" % (chunks, (time.perf_counter() - t0) / 40 * 1e3))
