"""Pinned-memory PCIe bandwidth of the box (the e2e ceiling): H2D alone, D2H alone, both at once, as a function of the copy
size and of the number of streams per direction. torch is plumbing here.  python tools/pcie_bw.py"""
import sys, time
import torch

TOT = 512 << 20  # bytes moved per direction per measurement


def run(h2d_mb, h2d_streams, d2h_mb, d2h_streams):
    res = []
    bufs = []
    for mb, ns, up in ((h2d_mb, h2d_streams, True), (d2h_mb, d2h_streams, False)):
        if mb <= 0:
            bufs.append(None)
            continue
        n = int(mb * (1 << 20))
        cnt = max(1, TOT // n)
        h = torch.empty(min(cnt, 64) * n, dtype=torch.uint8).pin_memory()
        d = torch.empty(min(cnt, 64) * n, dtype=torch.uint8, device="cuda")
        bufs.append((n, cnt, h, d, [torch.cuda.Stream() for _ in range(ns)], up))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ends = []
    maxc = max(b[1] for b in bufs if b)
    for i in range(maxc):
        for b in bufs:
            if not b or i >= b[1]:
                continue
            n, cnt, h, d, streams, up = b
            j = i % 64
            with torch.cuda.stream(streams[i % len(streams)]):
                if up:
                    d[j * n:(j + 1) * n].copy_(h[j * n:(j + 1) * n], non_blocking=True)
                else:
                    h[j * n:(j + 1) * n].copy_(d[j * n:(j + 1) * n], non_blocking=True)
    out = []
    for b in bufs:
        if not b:
            out.append(0.0)
            continue
        for st in b[4]:
            st.synchronize()
        out.append(b[0] * b[1] / (time.perf_counter() - t0) / 1e9)
    torch.cuda.synchronize()
    return out


run(5, 1, 5, 1)
print("H2D copy MB x streams | D2H copy MB x streams ->  H2D GB/s, D2H GB/s")
for cfg in [(64, 1, 0, 1), (0, 1, 64, 1), (64, 1, 64, 1), (1.8, 1, 0, 1), (0, 1, 5.4, 1), (1.8, 1, 5.4, 1), (1.8, 2, 5.4, 2), (1.8, 4, 5.4, 4), (1.8, 1, 5.4, 8),
            (232, 1, 5.4, 4), (232, 1, 5.4, 1), (16, 1, 5.4, 4), (1.8, 4, 9.6, 4), (1.8, 8, 5.4, 16)]:
    r = run(*cfg)
    print("%6.1f x %d | %6.1f x %2d -> %5.1f  %5.1f" % (cfg[0], cfg[1], cfg[2], cfg[3], r[0], r[1]))
