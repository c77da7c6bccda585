"""Does running the two halves of the batch as two concurrent streams beat one 256-tile launch chain (tile quantisation)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
enc = bench.build_encoder().to(dev)
x, m = bench.make_inputs(1000, 1)[0]
x = x.to(torch.bfloat16).to(dev); m = m.to(dev)
B = x.shape[0]

def fwd_split(n):
    cur = torch.cuda.current_stream()
    outs = []
    step = B // n
    for i in range(n):
        s = streams[i]
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append(enc(x[i * step:(i + 1) * step], src_key_padding_mask=m[i * step:(i + 1) * step])[0])
    for s in streams[:n]:
        cur.wait_stream(s)
    return outs

streams = [torch.cuda.Stream(dev) for _ in range(4)]
with torch.no_grad():
    y1 = enc(x, src_key_padding_mask=m)[0]
    for n in [int(a) for a in sys.argv[1:]] or (1, 2, 4):
        for _ in range(3):
            fwd_split(n)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs = fwd_split(n)
        g.replay(); torch.cuda.synchronize()
        y = torch.cat(outs, 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"splits={n}: {e0.elapsed_time(e1) / 20:.3f} ms/step, equal to unsplit: {bool(torch.equal(y, y1))}")
