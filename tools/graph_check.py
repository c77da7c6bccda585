"""Does a CUDA-graph replay of the encoder forward beat eager launches (i.e. is the eager loop host-bound)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
enc = bench.build_encoder().to(dev)
host = bench.make_inputs(1000, 4)
xs = [x.to(torch.bfloat16).to(dev) for x, _ in host]
ms = [m.to(dev) for _, m in host]
with torch.no_grad():
    for i in range(5):
        y = enc(xs[i % 4], src_key_padding_mask=ms[i % 4])[0]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        y = enc(xs[i % 4], src_key_padding_mask=ms[i % 4])[0]
    e1.record(); torch.cuda.synchronize()
    print("eager  ms/step", e0.elapsed_time(e1) / 20)
    y_eager = enc(xs[0], src_key_padding_mask=ms[0])[0].clone()
    g = torch.cuda.CUDAGraph()
    sx, sm = xs[0].clone(), ms[0].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            enc(sx, src_key_padding_mask=sm)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        yg = enc(sx, src_key_padding_mask=sm)[0]
    g.replay(); torch.cuda.synchronize()
    print("graph == eager:", bool(torch.equal(yg, y_eager)))
    e0.record()
    for i in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print("graph  ms/step", e0.elapsed_time(e1) / 20)
