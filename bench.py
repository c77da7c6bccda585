"""bench.py — encoder-forward frames/s of the SummaryMixing-Conformer hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libsmx through the module surface)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU; utterances sharded, no collective

Workload (BASELINE.json configs[1]): 12-layer SummaryMixing-Conformer encoder, d_model 256,
d_ffn 1024, nhead 4, kernel 31, Swish, mode "SummaryMixing"; per GPU one padded batch of B=32 utterances x
T=1000 frames x D=256 features (bf16), prefix padding masks with lengths in [500,1000].  A "step" is one
encoder forward over one batch.  Frames are counted as B*T (padded frames are computed, like the reference).

One JSON line on stdout (rank 0).  Keys follow the driver contract; see DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, T, D, FFN, HEADS, LAYERS, KSIZE = 32, 1000, 256, 1024, 4, 12, 31
METRIC = "encoder-fwd frames/sec (B=32,T=1000,D=256)"
N_ROTATE = 12  # distinct input batches cycled through the timed loop: 12 x 16.4 MB (+ outputs) > 126 MB of L2
WORKLOAD = ("cfg2: 12-layer SummaryMixing-Conformer encoder forward, D=256 d_ffn=1024 h=4 k=31, "
            "B=32 x T=1000 padded batch per GPU")


def config(world: int):
    """The `config` object, identical for both arms (--impl smx / reference): what is measured, not how."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "seq_len": T, "d_model": D, "layers": LAYERS, "n_gpus": world}


def ksm_traffic():
    """DRAM bytes of one K-SM call from the committed ncu capture (profiles/r02_ksm_traffic.json, written by
    tools/ncu_traffic.py from `ncu --set full` of tools/ncu_cell_capture.py; includes the trailing evict pass that forces the
    output's dirty lines out of L2).  None when the file is absent."""
    p = os.path.join(ROOT, "profiles", "r02_ksm_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 100 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        sm.sort()
        # "under load": the upper half of the samples (the sampler also sees the idle edges of the region)
        load = sm[len(sm) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
def make_inputs(seed: int, n: int):
    """n seeded synthetic batches: x ~ N(0,1) (B,T,D) and prefix masks with lens in [500,1000], lens[0]=T."""
    import torch

    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        x = torch.randn(B, T, D, generator=g)
        lens = torch.randint(500, T + 1, (B,), generator=g)
        lens[0] = T
        mask = torch.arange(T)[None] < lens[:, None]
        out.append((x, mask))
    return out


def build_encoder(seed: int = 0):
    import torch

    import summarymixing_b200 as S

    torch.manual_seed(seed)
    return S.ConformerEncoder(LAYERS, D, FFN, HEADS, KSIZE, attention_type="SummaryMixing", local_proj_hid_dim=[D],
                              local_proj_out_dim=D, summary_hid_dim=[D], mode="SummaryMixing").eval()


def cpu_reference_throughput(state_dict, budget_s: float, steps: int, warmup: int):
    """The reference's algorithm (oracle port, fp32 torch ops — what the reference's nn.Modules execute) on
    the host cores: frames/s over `steps` forwards of a bounded sample of the workload (first Bs utterances)."""
    import torch

    from oracle import smx_oracle as O  # checker / CPU baseline only

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: v.float().cpu() for k, v in state_dict.items()}
    x, mask = make_inputs(0, 1)[0]

    def run(bs):
        with torch.no_grad():
            t0 = time.perf_counter()
            O.conformer_encoder(x[:bs], sd, LAYERS, act="swish", src_key_padding_mask=mask[:bs])
            return time.perf_counter() - t0

    run(1)  # page in
    t2 = run(2)
    per_utt = t2 / 2
    n = max(1, steps + warmup)
    bs = int(max(1, min(B, budget_s / n / max(per_utt, 1e-6))))
    for _ in range(warmup):
        run(bs)
    times = [run(bs) for _ in range(steps)]
    tot = sum(times)
    return {"value": bs * T * steps / tot, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} forwards of the first {bs} of {B} utterances (T={T}, all {LAYERS} layers, fp32, "
                      f"{torch.get_num_threads()} torch threads); oracle/smx_oracle.py restates the reference modules",
            "ms_per_step": 1e3 * tot / steps, "frames_per_step": bs * T}


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    enc = build_encoder()
    r = cpu_reference_throughput(enc.state_dict(), budget_s=150.0, steps=args.steps, warmup=min(args.warmup, 2))
    line = {"metric": METRIC, "value": r["value"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": config(args.gpus), "details": {"frames_per_step": r["frames_per_step"], "device": "host CPU"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def time_module_calls(enc, x, mask, iters: int):
    """Per-module device time through the C ABI (CUDA events on the launching stream) for the roofline section."""
    import ctypes as C

    import torch

    from summarymixing_b200 import _host as H
    from summarymixing_b200 import _lib as L

    lib = L.lib()
    dev = x.device
    lw = enc._wv.struct[0]  # layer 0 weights (filled by the warm-up forwards)
    m8 = mask.to(torch.uint8).contiguous()
    st = H.stream_ptr(dev)
    rows = B * T
    nb = max(lib.smx_conformer_layer_workspace_bytes(C.byref(lw), L.BF16, B, T, 0),
             lib.smx_mixing_block_workspace_bytes(C.byref(lw.cell), L.BF16, B, T, 0),
             lib.smx_ffn_workspace_bytes(C.byref(lw.ffn1), L.BF16, rows),
             lib.smx_conv_module_workspace_bytes(C.byref(lw.conv), L.BF16, B, T), 1 << 20)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    xs = [torch.randn(B, T, D, device=dev).to(torch.bfloat16) for _ in range(N_ROTATE)]
    ys = [torch.empty_like(xs[0]) for _ in range(N_ROTATE)]

    # K-SM as it runs inside the layer: y = x + SummaryMixing(norm1(x)) (Conformer.py:520-541), `iters` batches back to back
    # through smx_mixing_block_fwd_batch (kernels chained by programmatic launch, one counter reset ahead of the chain -- the
    # way smx_conformer_encoder_fwd runs its layers)
    xs_p = (C.c_void_p * iters)(*[xs[i % N_ROTATE].data_ptr() for i in range(iters)])
    ys_p = (C.c_void_p * iters)(*[ys[i % N_ROTATE].data_ptr() for i in range(iters)])
    ws_cell = torch.empty(nb + iters * (4096 + 8 * B), dtype=torch.uint8, device=dev)

    def cell_batch(n):
        L.check(lib.smx_mixing_block_fwd_batch(C.byref(lw.cell), lw.norm1_w, lw.norm1_b, L.BF16, B, T, n, xs_p, m8.data_ptr(), ys_p,
                                               ws_cell.data_ptr(), ws_cell.numel(), st))

    def ffn(i):
        L.check(lib.smx_ffn_fwd(C.byref(lw.ffn1), lw.act, L.BF16, rows, xs[i].data_ptr(), None, None, 0.0,
                                ys[i].data_ptr(), ws.data_ptr(), ws.numel(), st))

    def conv(i):
        L.check(lib.smx_conv_module_fwd(C.byref(lw.conv), lw.act, L.BF16, B, T, 0, xs[i].data_ptr(), m8.data_ptr(),
                                        xs[i].data_ptr(), ys[i].data_ptr(), ws.data_ptr(), ws.numel(), st))

    out = {}
    cell_batch(3)
    torch.cuda.synchronize()
    n0 = lib.smx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cell_batch(iters)
    e1.record()
    torch.cuda.synchronize()
    out["cell"] = {"us": 1e3 * e0.elapsed_time(e1) / iters, "launches": int(lib.smx_launch_count() - n0) // iters}
    for name, fn in (("ffn", ffn), ("conv", conv)):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        n0 = lib.smx_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters):
            fn(i % N_ROTATE)
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"us": 1e3 * e0.elapsed_time(e1) / iters, "launches": int(lib.smx_launch_count() - n0) // iters}
    return out


def other_configs(dev, steps: int = 5):
    """Device-resident forward throughput of BASELINE.json configs[2] and [3] at their real model dims, one GPU:
    cfg3 = conformer_large (conformer_summarymixing.yaml:113-125: 12 layers, D=512, h=8, d_ffn=2048), B=32 x T=1000;
    cfg4 = Branchformer SummaryMixing-lite (branchformer_summarymixing.yaml:112-127: 18 layers, D=512, csgu 3072), variable-length
    padded batch, B=16, lengths uniform in [200,3000].  Frames = B*T padded (cfg4 also reports valid frames/s)."""
    import torch

    import summarymixing_b200 as S
    from summarymixing_b200 import _lib as L

    out = {}
    lib = L.lib()

    def timed(model, x, mask, valid_frames):
        with torch.no_grad():
            t0 = lib.smx_tc_launch_count()
            for _ in range(2):
                model(x, src_key_padding_mask=mask)
            torch.cuda.synchronize()
            tc = int(lib.smx_tc_launch_count() - t0) // 2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                model(x, src_key_padding_mask=mask)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        Bx, Tx = x.shape[0], x.shape[1]
        return {"ms_per_step": ms, "frames_per_s": Bx * Tx / ms * 1e3, "valid_frames_per_s": valid_frames / ms * 1e3,
                "batch": Bx, "T": Tx, "tcgen05_launches_per_step": tc, "launch": "eager", "io": "bf16"}

    def with_roofline(r, mflop_per_frame_layer, layers):
        # whole-step GEMM FLOPs (2 per multiply-add; padded frames are computed) against the sustained bf16 tensor peak
        pk = peaks()
        fl = mflop_per_frame_layer * 1e6 * layers * r["batch"] * r["T"]
        r["roofline_step"] = {"bound": "tensor", "flops_per_step": fl, "mflop_per_frame_layer": mflop_per_frame_layer,
                              "achieved": fl / (r["ms_per_step"] * 1e-3) / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                              "frac": fl / (r["ms_per_step"] * 1e-3) / 1e12 / pk["bf16_tflops_sustained"], "peak_source": pk["source"]}
        return r

    g = torch.Generator().manual_seed(7)
    try:
        torch.manual_seed(3)
        enc = S.ConformerEncoder(12, 512, 2048, 8, 31, attention_type="SummaryMixing", local_proj_hid_dim=[512], local_proj_out_dim=512,
                                 summary_hid_dim=[512], mode="SummaryMixing").eval().to(dev)
        x = torch.randn(32, 1000, 512, generator=g).to(torch.bfloat16).to(dev)
        lens = torch.randint(500, 1001, (32,), generator=g)
        lens[0] = 1000
        mask = (torch.arange(1000)[None] < lens[:, None]).to(dev)
        # MFLOP per frame and layer: 2 FFNs of 2 linears 512 x 2048 (8.39), cell with h=8 (four block-diagonal 512 x 512 / 8 linears
        # 0.26, combiner's local half 0.52), conv module (pointwise 512 -> 1024 1.05, depthwise k=31 0.03, linear 0.52)
        out["cfg3_conformer_large_D512"] = with_roofline(timed(enc, x, mask, int(lens.sum())), 8.39 + 0.26 + 0.52 + 1.05 + 0.03 + 0.52, 12)
        del enc, x
    except Exception as exc:  # a secondary line must never take the headline down
        out["cfg3_conformer_large_D512"] = {"error": repr(exc)[:200]}
    try:
        torch.manual_seed(4)
        enc = S.BranchformerEncoder(18, 512, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[512], local_proj_out_dim=512,
                                    summary_hid_dim=[512], summary_out_dim=512, mode="SummaryMixing-lite").eval().to(dev)
        lens = torch.randint(200, 3001, (16,), generator=g)
        Tm = int(lens.max())
        x = torch.randn(16, Tm, 512, generator=g).to(torch.bfloat16).to(dev)
        mask = (torch.arange(Tm)[None] < lens[:, None]).to(dev)
        # MFLOP per frame and layer: pre-projection 512 -> 3072 (3.15), CSGU depthwise over 1536 channels (0.10), post-projection
        # 1536 -> 512 (1.57), lite cell's summary MLP 2 x 512 x 512 (1.05), merge_proj 1024 -> 512 -> 512 (1.57)
        out["cfg4_branchformer_lite_D512"] = with_roofline(timed(enc, x, mask, int(lens.sum())), 3.15 + 0.10 + 1.57 + 1.05 + 1.57, 18)
        del enc, x
    except Exception as exc:
        out["cfg4_branchformer_lite_D512"] = {"error": repr(exc)[:200]}
    torch.cuda.empty_cache()
    return out


def branchformer_train_line(dev, steps: int = 3):
    """Training step of the Branchformer SummaryMixing-lite encoder at the recipe's model dims and dropout
    (branchformer_summarymixing.yaml: 18 layers, D=512, csgu 3072, dropout 0.1), B=8 x T=1000 per GPU, no gradient exchange:
    forward through the layers' autograd chains (smx_conv_branch_train_fwd, the cell, merge_proj, counter-based dropout masks),
    backward through smx_conv_branch_train_bwd / smx_summary_mixing_bwd / smx_vanilla_nn_bwd, SGD."""
    import torch

    import summarymixing_b200 as S

    try:
        torch.manual_seed(5)
        Bb, Tb, Db = 8, 1000, 512
        enc = S.BranchformerEncoder(18, Db, 1, 31, csgu_linear_units=3072, local_proj_hid_dim=[512], local_proj_out_dim=512,
                                    summary_hid_dim=[512], summary_out_dim=512, mode="SummaryMixing-lite", dropout=0.1).to(dev).train()
        opt = torch.optim.SGD(enc.parameters(), lr=0.01)
        g = torch.Generator().manual_seed(300)
        x = torch.randn(Bb, Tb, Db, generator=g).to(dev).to(torch.bfloat16)
        lens = torch.randint(Tb // 2, Tb + 1, (Bb,), generator=g)
        mask = (torch.arange(Tb)[None] < lens[:, None]).to(dev)
        target = torch.randn(Bb, Tb, Db, generator=g).to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            y = enc(x, src_key_padding_mask=mask)[0]
            loss = ((y.float() - target) * mask[..., None]).pow(2).mean()
            loss.backward()
            opt.step()
            return loss

        l0 = float(step())
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time as _time

        h0 = _time.perf_counter()
        e0.record()
        for _ in range(steps):
            l1 = step()
        e1.record()
        host_ms = (_time.perf_counter() - h0) * 1e3 / steps   # host time to enqueue a step (no read-back inside the loop)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"ms_per_step": ms, "host_enqueue_ms_per_step": host_ms, "frames_per_s": Bb * Tb / ms * 1e3, "batch": Bb, "T": Tb, "layers": 18, "dropout": 0.1,
               "loss_first": l0, "loss_last": float(l1), "io": "bf16 activations, fp32 gradients"}
        del enc, opt
    except Exception as exc:  # a secondary line must never take the headline down
        out = {"error": repr(exc)[:200]}
    torch.cuda.empty_cache()
    return out


def train_line(dev, world: int, rank: int, steps: int = 3):
    """One data-parallel TRAINING step of the same encoder (forward + backward through smx_*_bwd + gradient exchange + SGD), every
    rank on its own B x T batch.  The gradient all-reduce (NCCL over NVLink, flat fp32 buckets) is launched from gradient-ready
    hooks during backward (summarymixing_b200.parallel.GradientBucketer); `exposed_comm_ms` = step time minus the time of the
    same step without any exchange.  Called by ALL ranks (collectives inside); times are the max over ranks."""
    import torch
    import torch.distributed as dist

    import summarymixing_b200 as S
    from summarymixing_b200 import parallel as P

    torch.manual_seed(0)
    enc = S.ConformerEncoder(LAYERS, D, FFN, HEADS, KSIZE, attention_type="SummaryMixing", local_proj_hid_dim=[D], local_proj_out_dim=D,
                             summary_hid_dim=[D], mode="SummaryMixing", dropout=0.0).to(dev).train()
    params = list(enc.parameters())
    opt = torch.optim.SGD(params, lr=0.01)
    g = torch.Generator().manual_seed(200 + rank)
    x = torch.randn(B, T, D, generator=g).to(dev).to(torch.bfloat16)
    lens = torch.randint(T // 2, T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] < lens[:, None]).to(dev)
    target = torch.randn(B, T, D, generator=g).to(dev)
    bucketer = P.GradientBucketer(params) if world > 1 else None

    def step(exchange: bool):
        opt.zero_grad(set_to_none=True)
        y = enc(x, src_key_padding_mask=mask)[0]
        loss = ((y.float() - target) * mask[..., None]).pow(2).mean()
        loss.backward()
        calls = bucketer.finish() if (bucketer and exchange) else 0
        opt.step()
        return calls

    def timed(exchange: bool):
        calls = 0
        for _ in range(2):
            calls = step(exchange)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step(exchange)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        return P.max_over_ranks(e0.elapsed_time(e1) / steps, device=dev), calls

    ms_x, calls = timed(True)
    if bucketer:
        bucketer.remove()
        bucketer = None
    ms_c, _ = timed(False) if world > 1 else (ms_x, 0)
    n_par = sum(p.numel() for p in params)
    del enc, opt
    torch.cuda.empty_cache()
    branch = branchformer_train_line(dev, steps)
    return {"branchformer_lite_D512": branch, "ms_per_step": ms_x, "frames_per_s": world * B * T / ms_x * 1e3, "io": "bf16 activations, fp32 gradients",
            "backward": "smx_*_bwd: linears (recompute, dgrad, wgrad) as split-bf16 tcgen05 GEMMs, elementwise / reductions on CUDA cores",
            "allreduce_bytes_per_step": n_par * 4 if world > 1 else 0, "allreduce_calls_per_step": calls,
            "exchange": "flat fp32 buckets of 32 MB, all-reduce launched from gradient-ready hooks during backward (async NCCL)" if world > 1 else "none (1 GPU)",
            "compute_only_ms": ms_c, "exposed_comm_ms": max(0.0, ms_x - ms_c), "params": n_par, "steps": steps}


def run_smx(args):
    import torch
    import torch.distributed as dist

    from summarymixing_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path is CUDA only; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    pk = peaks()

    enc = build_encoder().to(dev)
    # utterances are sharded over ranks (weak scaling: every rank owns its own B=32 batch; no data-path collective)
    host = make_inputs(1000 + rank, N_ROTATE)
    xs = [x.to(torch.bfloat16).to(dev) for x, _ in host]
    ms = [m.to(dev) for _, m in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import summarymixing_b200 as S

    with torch.no_grad():
        # one eager forward counts the kernels of a step (a graph replay launches the same kernels without calling the library)
        enc(xs[0], src_key_padding_mask=ms[0])
        torch.cuda.synchronize()
        n0, t0 = lib.smx_launch_count(), lib.smx_tc_launch_count()
        enc(xs[1], src_key_padding_mask=ms[1])
        torch.cuda.synchronize()
        per_step, tc_per_step = int(lib.smx_launch_count() - n0), int(lib.smx_tc_launch_count() - t0)
        # the timed step: the forward captured once into a CUDA graph (summarymixing_b200.GraphedForward); every step copies
        # its input batch (one of N_ROTATE distinct device buffers) into the graph's input buffer and replays the graph
        mode = "cuda-graph replay (summarymixing_b200.GraphedForward), input copied device-to-device from a rotating batch each step"
        try:
            if args.eager:
                raise RuntimeError("eager requested")
            step_fn = S.GraphedForward(enc, xs[0], ms[0])
        except Exception as exc:  # capture is an optimisation of the host side only: fall back to eager launches
            mode = f"eager launches ({exc})"
            step_fn = lambda x, m: enc(x, src_key_padding_mask=m)[0]  # noqa: E731
        for i in range(max(args.warmup, 3)):
            step_fn(xs[i % N_ROTATE], ms[i % N_ROTATE])
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            step_fn(xs[i % N_ROTATE], ms[i % N_ROTATE])
        e1.record()
        barrier()
        launches, tc_launches = per_step * args.steps, tc_per_step * args.steps
        ms_total = e0.elapsed_time(e1)

        # ---- e2e: same metric through the package's host-buffer entry point (summarymixing_b200.HostPipeline) with HOST
        # tensors: every step's H2D (x, mask) and D2H (y) are inside the timed region, overlapped with the neighbouring steps.
        hx = [x.to(torch.bfloat16).pin_memory() for x, _ in host[:4]]
        hm = [m.pin_memory() for _, m in host[:4]]
        pipe = S.HostPipeline(enc, B, T, D, device=dev, use_graph=not args.eager)

        def e2e_loop(n):
            last = None
            for last in pipe.run(((hx[i % 4], hm[i % 4]) for i in range(n))):
                pass
            return last

        e2e_loop(3)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        e2e_loop(args.steps)
        f1.record()
        barrier()
        e2e_ms = f0.elapsed_time(f1)
        clocks = sampler.stop() if sampler else None

    train = None
    if not args.no_train:
        try:
            train = train_line(dev, world, rank)
        except Exception as exc:  # never take the headline down
            train = {"error": repr(exc)[:200]}
    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = world * B * T / (ms_step / 1e3)
    e2e_value = world * B * T / (e2e_ms / args.steps / 1e3)

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config(world),
            "details": {"frames_per_step": world * B * T,
                       "parallelism": f"utterance-sharded x{world} (no collective)",
                       "l2": f"inputs rotate over {N_ROTATE} distinct batches ({N_ROTATE * B * T * D * 2 / 1e6:.0f} MB "
                             "of x > 126 MB L2) so no step finds its input in L2",
                       "accumulate": "fp32", "io": "bf16", "launch": mode},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": pipe.h2d_bytes_per_step,
                    "d2h_bytes_per_step": pipe.d2h_bytes_per_step, "ms_per_step": e2e_ms / args.steps,
                    "api": "summarymixing_b200.HostPipeline(ConformerEncoder).run(host batches): pinned host inputs/outputs, H2D and "
                           "D2H on their own streams inside the timed region, overlapped with the previous/next step; forward = "
                           "CUDA-graph replay of ConformerEncoder.forward"},
            "gpu_launches": launches, "tc_launches": tc_launches, "clocks": clocks, "train": train}

    if rank == 0:
        with torch.no_grad():
            mod = time_module_calls(enc, xs[0], ms[0], iters=max(args.steps, 10))
        # K-SM, the kernel north_star names: algorithmic bytes = read x + write y (bf16) + 1 mask byte per frame
        # (SURVEY.md 8d); FLOPs as executed (block-diagonal f/s projections, split combiner).
        frames = B * T
        hd = D // HEADS
        bytes_cell = frames * ((D + D) * 2 + 1)
        flops_cell = frames * (4 * 2 * HEADS * hd * hd + 2 * D * D)
        us = mod["cell"]["us"]
        tr = ksm_traffic()
        traffic = tr["dram_bytes_per_call"] if tr else None
        line["roofline"] = {"kernel": f"smx_mixing_block_fwd: norm1 + SummaryMixing cell + skip ({mod['cell']['launches']} launch(es); "
                                      "the unit the fused cell kernel executes inside a layer)",
                            "bound": "hbm", "achieved": bytes_cell / us / 1e3, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": bytes_cell / us / 1e3 / pk["hbm_gbs"],
                            "traffic": traffic, "traffic_ratio": (traffic / bytes_cell) if traffic else None,
                            "traffic_source": (tr or {}).get("source"), "algorithmic_bytes": bytes_cell,
                            "us_per_call": us, "peak_source": pk["source"] + " (burst copy)",
                            "tensor_tflops": flops_cell / us / 1e6, "tensor_frac": flops_cell / us / 1e6 / pk["bf16_tflops"],
                            # what bounds this kernel in practice: one MUFU.TANH per activation, 5 activations per element,
                            # 16 MUFU/clk/SM (tools/micro/tmem_bench.cu): frames * 5 * D / (16 * SMs * clock)
                            "mufu_floor_us": frames * 5 * D / (16 * 148 * 1.965e3)}
        flops_ffn = frames * 4 * D * FFN
        flops_conv = frames * 2 * D * 3 * D
        step_us = ms_step * 1e3
        # whole step: FLOPs as executed per frame and layer (2 FFNs, cell with block-diagonal MLPs + split combiner, conv module)
        flops_step = LAYERS * (2 * flops_ffn + flops_cell + flops_conv + frames * 2 * 31 * D)
        line["roofline_step"] = {"bound": "tensor", "flops": flops_step, "achieved": flops_step / step_us / 1e6,
                                 "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                 "frac": flops_step / step_us / 1e6 / pk["bf16_tflops_sustained"],
                                 "peak_source": pk["source"] + " (sustained cuBLAS bf16)"}
        line["kernels"] = {
            "cell": {"us": us, "launches": mod["cell"]["launches"], "share_of_step": LAYERS * us / step_us},
            "ffn": {"us": mod["ffn"]["us"], "launches": mod["ffn"]["launches"],
                    "share_of_step": 2 * LAYERS * mod["ffn"]["us"] / step_us, "bound": "tensor",
                    "tflops": flops_ffn / mod["ffn"]["us"] / 1e6,
                    "frac": flops_ffn / mod["ffn"]["us"] / 1e6 / pk["bf16_tflops_sustained"]},
            "conv": {"us": mod["conv"]["us"], "launches": mod["conv"]["launches"],
                     "share_of_step": LAYERS * mod["conv"]["us"] / step_us, "bound": "tensor",
                     "tflops": flops_conv / mod["conv"]["us"] / 1e6,
                     "frac": flops_conv / mod["conv"]["us"] / 1e6 / pk["bf16_tflops_sustained"]},
        }
        if world == 1 and not args.no_others:
            # the same encoder with fp32 I/O: linears as split-bf16 tensor-core GEMMs (the arm that meets north_star's 1e-3)
            with torch.no_grad():
                xf, mf = host[0][0].to(dev), ms[0]
                for _ in range(2):
                    enc(xf, src_key_padding_mask=mf)
                torch.cuda.synchronize()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(5):
                    enc(xf, src_key_padding_mask=mf)
                g1.record()
                torch.cuda.synchronize()
            ms32 = g0.elapsed_time(g1) / 5
            line["fp32_arm"] = {"ms_per_step": ms32, "frames_per_s": B * T / ms32 * 1e3, "io": "fp32",
                                "math": "linears: split-bf16 (bf16x3) tcgen05 GEMMs with fp32 accumulation; everything else fp32 on CUDA cores",
                                "max_abs_vs_oracle": "<= 1e-3 (tests/test_fullsize_parity_gpu.py prints the measured value)"}
            line["other_configs"] = other_configs(dev)
        if world == 1 and not args.no_cpu:
            r = cpu_reference_throughput(enc.state_dict(), budget_s=20.0, steps=2, warmup=1)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="smx", choices=["smx", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step side measurement")
    ap.add_argument("--no-others", action="store_true", help="skip the cfg3 / cfg4 side measurements")
    ap.add_argument("--eager", action="store_true", help="time eager launches instead of CUDA-graph replay (ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_smx(args)


if __name__ == "__main__":
    main()
