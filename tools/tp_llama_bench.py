"""Tensor-parallel MX-linear inference + layer-sharded weight quantization at Llama-3-70B shape (BASELINE configs[4]).

One process per GPU (torchrun).  Two measurements:

 (i)  `--mode quantize`: every rank creates (random bf16, on device) and quantizes its contiguous range of the 80
      decoder layers' Linear weights to float4_e2m1 -- no collective on the data path; aggregate GB/s = sum of
      algorithmic bytes / max-over-ranks device time.
 (ii) `--mode infer`: a Llama decoder stack whose seven projections per layer are ColumnParallelMXLinear /
      RowParallelMXLinear (weights fp4_e2m1 by default, activations fp8_e4m3): q/k/v/gate/up split out_features,
      o/down split in_features and all-reduce their bf16 partial outputs over NCCL.  Prefill of 2048 tokens and decode
      at batch 32, each replayed from a CUDA graph; tokens/s = tokens / max-over-ranks device time.
      Between the projections a layer is five more launches of this library (RMSNorm with the residual add and the activation
      quantization folded in -- K5a --, rotary embedding -- K5b --, SiLU gating + quantization -- K1b) plus the KV-cache update
      and the attention over the local heads (PyTorch SDPA).  `--no-glue` runs the same stack with plain PyTorch glue.

 `--check` runs a small configuration and compares the TP output with the same stack evaluated on one rank.

    torchrun --nproc-per-node G tools/tp_llama_bench.py --mode infer [--layers L] [--wdtype float4_e2m1]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SHAPES = {
    "70b": dict(hidden=8192, inter=28672, layers=80, heads=64, kv_heads=8, vocab=128256),
    "8b": dict(hidden=4096, inter=14336, layers=32, heads=32, kv_heads=8, vocab=128256),
    "small": dict(hidden=1024, inter=2048, layers=2, heads=8, kv_heads=8, vocab=1024),
}


def rms_norm(x, w, eps=1e-5):
    return F.rms_norm(x, (x.shape[-1],), w, eps)


CACHE_IN_ROPE = os.environ.get("TP_CACHE_IN_ROPE", "1") != "0"
GLUE = True  # --no-glue: RMSNorm / rotary / SiLU as plain PyTorch ops, one launch per projection


def rope_tables(pos, d, theta=500000.0):
    """cos / sin for the positions of this forward pass, shared by every layer: [1, 1, T, d/2] fp32 (plain path) and the
    transformers-style bf16 [1, T, d] tables (cat(freqs, freqs)) the rotary kernel reads"""
    inv = 1.0 / (theta ** (torch.arange(0, d, 2, device=pos.device, dtype=torch.float32) / d))
    ang = pos.float()[:, None] * inv[None, :]
    emb = torch.cat([ang, ang], -1)[None]
    return ang.cos()[None, None], ang.sin()[None, None], emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)


def rope(x, cs):
    # x: [B, H, T, D] (query and key heads concatenated along H), rotate-half convention
    cos, sin = cs[0], cs[1]
    d = x.shape[-1]
    xf = x.float()
    x1, x2 = xf[..., : d // 2], xf[..., d // 2:]
    return torch.cat([x1 * cos - x2 * sin, x2 * cos + x1 * sin], -1).to(x.dtype)


class TPDecoderLayer(torch.nn.Module):
    def __init__(self, cfg, qc, world, rank, seed, full=False):
        super().__init__()
        from torchmx_b200.layers.tp_linear import ColumnParallelMXLinear, RowParallelMXLinear
        h, inter, nh, nkv = cfg["hidden"], cfg["inter"], cfg["heads"], cfg["kv_heads"]
        hd = h // nh
        self.hd, self.nh_local, self.nkv_local = hd, nh // world, max(nkv // world, 1)
        assert nh % world == 0 and (nkv % world == 0 or world % nkv == 0)
        g = torch.Generator(device="cuda").manual_seed(seed)
        wr = (world, rank)

        def lin(n_out, n_in, kind, kv=False):
            # the FULL weight is generated from the seed on every rank (so a 1-rank run of the same seed is the reference),
            # one projection at a time; only the shard survives
            w = torch.randn(n_out, n_in, device="cuda", dtype=torch.bfloat16, generator=g) * (n_in ** -0.5)
            m = torch.nn.Linear(n_in, n_out, bias=False, device="meta")
            m.weight = torch.nn.Parameter(w, requires_grad=False)
            if kv and world > nkv:  # more ranks than KV heads: ranks sharing a KV head hold the same (replicated) slice
                per = n_out // nkv
                head = rank * nkv // world
                m2 = torch.nn.Linear(n_in, per, bias=False, device="meta")
                m2.weight = torch.nn.Parameter(w[head * per:(head + 1) * per].contiguous(), requires_grad=False)
                return ColumnParallelMXLinear.from_float(m2, qc, world_rank=(1, 0))
            cls = ColumnParallelMXLinear if kind == "col" else RowParallelMXLinear
            return cls.from_float(m, qc, world_rank=wr)

        self.q = lin(nh * hd, h, "col")
        self.k = lin(nkv * hd, h, "col", kv=True)
        self.v = lin(nkv * hd, h, "col", kv=True)
        self.o = lin(h, nh * hd, "row")
        self.gate = lin(inter, h, "col")
        self.up = lin(inter, h, "col")
        self.down = lin(h, inter, "row")
        self.n1 = torch.ones(h, device="cuda", dtype=torch.bfloat16)
        self.n2 = torch.ones(h, device="cuda", dtype=torch.bfloat16)
        self.act = qc.activations_config.elem_dtype if qc.activations_config.block_size == 32 else None
        self.qkv = self.gate_up = None
        if GLUE:  # q/k/v and gate/up read the same activation: one launch each on row-stacked weights the three / two layers alias
            from torchmx_b200.layers.mx_llama_attention import _fuse_linears
            self.qkv = _fuse_linears([self.q, self.k, self.v])
            self.gate_up = _fuse_linears([self.gate, self.up])

    def forward(self, x, res, pos, kc, vc, cache_len, cs):
        """hidden state of the layer = x + res (res None for the first layer); returns (down-projection output, residual stream)
        -- the add is folded into the next RMSNorm launch.  x: [B, T, hidden] replicated; kc / vc: [B, nkv_local, S, hd] this
        rank's KV cache; tokens are written at pos."""
        if not GLUE or self.qkv is None or self.gate_up is None or self.act is None:
            return self.forward_plain(x if res is None else x + res, pos, kc, vc, cache_len, cs), None
        from torchmx_b200 import glue_ops, mlp_ops
        B, T, _ = x.shape
        to_mx = self.act  # (also at decode sizes: a GEMM fed codes streams packed weights faster than one that quantizes per CTA)
        r = glue_ops.rmsnorm(x, self.n1, 1e-5, residual=res, to_mx=to_mx, want_y=to_mx is None)
        assert r is not None, "the RMSNorm kernel declined a [B, T, hidden] bf16 activation"
        y, y_mx, h = r
        h = x if res is None else h
        q, k, v = self.qkv(y if to_mx is None else y_mx).split(self.qkv._split, dim=-1)
        q = q.view(B, T, self.nh_local, self.hd).transpose(1, 2)
        k = k.view(B, T, self.nkv_local, self.hd).transpose(1, 2)
        v = v.view(B, T, self.nkv_local, self.hd).transpose(1, 2)
        # rotary embedding, and the KV-cache update folded into the same launch: the rotated keys are written in place into the
        # cache slice, the value heads copied beside them (tokens go to positions p0 .. p0 + T - 1, which is what `pos` holds)
        p0 = 0 if T > 1 else cache_len
        r = glue_ops.rope(q, k, cs[2], cs[3], k_out=kc[:, :, p0:p0 + T], v=v, v_out=vc[:, :, p0:p0 + T]) if CACHE_IN_ROPE else None
        if r is not None:
            q, k = r
        else:
            r = glue_ops.rope(q, k, cs[2], cs[3])
            assert r is not None, "the rotary kernel declined the projection output"
            q, k = r
            kc.index_copy_(2, pos, k)
            vc.index_copy_(2, pos, v)
        if T > 1:  # prefill from an empty cache: causal attention over the new tokens
            a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
        else:      # decode: attend to the first cache_len + 1 cache positions
            a = F.scaled_dot_product_attention(q, kc[:, :, : cache_len + 1], vc[:, :, : cache_len + 1], enable_gqa=True)
        a_mx = glue_ops.quantize_heads(a, self.act)  # o's activation, quantized straight from the [B, H, T, D] attention output (K5c)
        o = self.o(a_mx if a_mx is not None else a.transpose(1, 2).reshape(B, T, self.nh_local * self.hd))
        r = glue_ops.rmsnorm(o, self.n2, 1e-5, residual=h, to_mx=to_mx, want_y=to_mx is None)
        assert r is not None
        y, y_mx, h = r
        gate, up = self.gate_up(y if to_mx is None else y_mx).split(self.gate_up._split, dim=-1)
        act = mlp_ops.silu_mul_to_mx(gate, up, self.act, 32)
        return self.down(act if act is not None else F.silu(gate) * up), h

    def forward_plain(self, x, pos, kc, vc, cache_len, cs):
        B, T, _ = x.shape
        y = self.q.prepare_input(rms_norm(x, self.n1))  # one activation quantization for q / k / v (prefill); bf16 for decode
        q = self.q(y).view(B, T, self.nh_local, self.hd).transpose(1, 2)
        k = self.k(y).view(B, T, self.nkv_local, self.hd).transpose(1, 2)
        v = self.v(y).view(B, T, self.nkv_local, self.hd).transpose(1, 2)
        qk = rope(torch.cat([q, k], 1), cs)
        q, k = qk[:, : self.nh_local], qk[:, self.nh_local:]
        kc.index_copy_(2, pos, k)
        vc.index_copy_(2, pos, v)
        if T > 1:
            a = F.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
        else:
            a = F.scaled_dot_product_attention(q, kc[:, :, : cache_len + 1], vc[:, :, : cache_len + 1], enable_gqa=True)
        x = x + self.o(a.transpose(1, 2).reshape(B, T, self.nh_local * self.hd))
        y = self.gate.prepare_input(rms_norm(x, self.n2))
        return x + self.down(F.silu(self.gate(y)) * self.up(y))


def time_graph(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


@torch.no_grad()
def run_infer(args, world, rank):
    import torchmx_b200  # noqa: F401
    from torchmx_b200.config import MXConfig, QLinearConfig
    global GLUE
    GLUE = not getattr(args, "no_glue", False)
    cfg = dict(SHAPES[args.model])
    if args.layers:
        cfg["layers"] = args.layers
    qc = QLinearConfig(weights_config=MXConfig(args.wdtype, 32), activations_config=MXConfig(args.adtype, 32))
    t0 = time.perf_counter()
    layers = [TPDecoderLayer(cfg, qc, world, rank, seed=1000 + i) for i in range(cfg["layers"])]
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    h = cfg["hidden"]
    pool = None
    if args.fused and world > 1:
        from torchmx_b200.layers.tp_linear import FusedAllReducePool
        pool = FusedAllReducePool(h, max(args.prefill, args.batch), None, fused_max_rows=args.fused_max_rows)
        for l in layers:
            l.o.enable_fused_allreduce(pool)
            l.down.enable_fused_allreduce(pool)
    for l in layers:
        if pool is None:
            l.o._fused_pool = l.down._fused_pool = None
    gen = torch.Generator(device="cuda").manual_seed(5)
    res = {"mode": "infer", "glue_kernels": GLUE, "fused_allreduce": bool(args.fused and world > 1), "model": args.model, "layers": cfg["layers"], "world": world, "weights": args.wdtype, "activations": args.adtype,
           "build_s": round(build_s, 2), "weight_GB_per_rank": round(torch.cuda.memory_allocated() / 1e9, 2)}

    def stack(x, pos, caches, cache_len):
        if pool is not None and layers[0].o._fused_pool is not None:
            pool.reset()
        cs = rope_tables(pos, layers[0].hd)
        res = None
        for l, (kc, vc) in zip(layers, caches):
            x, res = l(x, res, pos, kc, vc, cache_len, cs)
        return x if res is None else x + res

    def make_caches(B, S):
        return [(torch.zeros(B, l.nkv_local, S, l.hd, device="cuda", dtype=torch.bfloat16), torch.zeros(B, l.nkv_local, S, l.hd, device="cuda", dtype=torch.bfloat16))
                for l in layers]

    # prefill
    P = args.prefill
    x = torch.randn(1, P, h, device="cuda", dtype=torch.bfloat16, generator=gen)
    pos = torch.arange(P, device="cuda")
    caches = make_caches(1, P)
    ms, out = time_graph(lambda: stack(x, pos, caches, 0), args.iters)
    res["prefill_ms"], res["prefill_tok_s"] = round(ms, 3), round(P / ms * 1e3, 1)
    res["prefill_out_absmean"] = float(out.float().abs().mean())
    del caches
    # decode (fixed context length: the cache is pre-filled, every replay decodes one token per sequence at position ctx)
    B, ctx = args.batch, args.ctx
    caches = make_caches(B, ctx + 8)
    for kc, vc in caches:
        kc.normal_(generator=gen)
        vc.normal_(generator=gen)
    xd = torch.randn(B, 1, h, device="cuda", dtype=torch.bfloat16, generator=gen)
    posd = torch.tensor([ctx], device="cuda")
    ms, out = time_graph(lambda: stack(xd, posd, caches, ctx), args.iters * 4)
    res["decode_ms_per_step"], res["decode_tok_s"] = round(ms, 3), round(B / ms * 1e3, 1)
    if getattr(args, "also_fused", False) and world > 1 and pool is None:
        # the same layers once more with the reduction done in the GEMM epilogue (NVLink multicast) instead of NCCL
        from torchmx_b200.layers.tp_linear import FusedAllReducePool
        pool = FusedAllReducePool(h, max(args.batch, 128), None, fused_max_rows=args.fused_max_rows)
        for l in layers:
            l.o.enable_fused_allreduce(pool)
            l.down.enable_fused_allreduce(pool)
        ms, out = time_graph(lambda: stack(xd, posd, caches, ctx), args.iters * 4)
        res["decode_fused_ms_per_step"], res["decode_fused_tok_s"] = round(ms, 3), round(B / ms * 1e3, 1)
    res["allreduce_per_step"] = 2 * cfg["layers"] if world > 1 else 0
    res["allreduce_bytes_decode"], res["allreduce_bytes_prefill"] = B * h * 2, P * h * 2
    elems = cfg["layers"] * (2 * h * h + 2 * (h // cfg["heads"]) * cfg["kv_heads"] * h + 3 * h * cfg["inter"])
    bpe = (0.5 if args.wdtype == "float4_e2m1" else 1.0) + 1 / 32
    res["decode_weight_stream_floor_ms_per_rank"] = round(elems * bpe / world / 6.5e12 * 1e3, 3)
    from torchmx_b200 import mx_gemm
    res["gemm_stats"] = dict(mx_gemm.stats)
    if args.check:
        # every rank evaluated the same seeds; compare rank 0's TP result with a one-rank evaluation of the same stack
        ref_layers = [TPDecoderLayer(cfg, qc, 1, 0, seed=1000 + i) for i in range(cfg["layers"])]
        xr = x.clone()
        c1 = [(torch.zeros(1, l.nkv_local, P, l.hd, device="cuda", dtype=torch.bfloat16), torch.zeros(1, l.nkv_local, P, l.hd, device="cuda", dtype=torch.bfloat16))
              for l in ref_layers]
        rr = None
        for l, (kc, vc) in zip(ref_layers, c1):
            xr, rr = l(xr, rr, pos, kc, vc, 0, rope_tables(pos, l.hd))
        xr = xr if rr is None else xr + rr
        c2 = make_caches(1, P)
        xt = stack(x, pos, c2, 0).clone()
        if pool is not None:
            # the same TP stack with the NCCL all-reduce instead of the fused epilogue: only the summation differs (switch
            # vs ring order), everything else is bit-identical
            for l in layers:
                l.o._fused_pool = l.down._fused_pool = None
            xn = stack(x, pos, make_caches(1, P), 0)
            res["check_rel_err_fused_vs_nccl"] = float((xt.float() - xn.float()).norm() / xn.float().norm())
            # single row-parallel layer, decode and prefill sized: fused vs NCCL on identical inputs
            pool.fused_max_rows = pool.max_rows  # exercise the CTA-pair kernel's fused epilogue too
            for rows in (args.batch, P):
                a = torch.randn(rows, layers[0].o.in_features, device="cuda", dtype=torch.bfloat16, generator=gen)
                y_n = layers[0].o(a).clone()
                layers[0].o._fused_pool = pool
                pool.reset()
                y_f = layers[0].o(a).clone()
                layers[0].o._fused_pool = None
                res[f"check_layer_rows{rows}_max_abs_diff"] = float((y_f.float() - y_n.float()).abs().max())
                res[f"check_layer_rows{rows}_ref_absmax"] = float(y_n.float().abs().max())
                assert (y_f.float() - y_n.float()).abs().max() <= 2.0 ** -6 * y_n.float().abs().max() + 1e-3
            assert res["check_rel_err_fused_vs_nccl"] < 6e-2
        err = (xt.float() - xr.float()).norm() / xr.float().norm()
        res["check_rel_err_vs_1rank"] = float(err)
        # activations are re-quantized (fp8) between layers, so a last-bit difference of a bf16 partial sum can flip a code:
        # the stacks agree to a few percent, not to an ulp (per-layer exactness is covered by tests/test_gpu_tp.py), and on
        # random-init weights the difference grows with depth (4 layers: < 0.06, 80 layers: 0.13 at two ranks)
        assert err < (6e-2 if cfg["layers"] <= 16 else 0.25), err
    return res


@torch.no_grad()
def run_quantize(args, world, rank):
    import torchmx_b200  # noqa: F401
    from torchmx_b200.config import MXConfig, QLinearConfig
    from torchmx_b200.layers.mx_linear import MXInferenceLinear
    from torchmx_b200.sharding import layer_shard
    cfg = dict(SHAPES[args.model])
    if args.layers:
        cfg["layers"] = args.layers
    qc = QLinearConfig(weights_config=MXConfig(args.wdtype, 32), activations_config=MXConfig(args.adtype, 32))
    h, inter, hd = cfg["hidden"], cfg["inter"], cfg["hidden"] // cfg["heads"]
    shapes = [(h, h), (cfg["kv_heads"] * hd, h), (cfg["kv_heads"] * hd, h), (h, h), (inter, h), (inter, h), (h, inter)]
    lo, hi = layer_shard(list(range(cfg["layers"])), rank, world)
    gen = torch.Generator(device="cuda").manual_seed(rank)
    # one decoder layer's bf16 weights are resident at a time (a loader streams layers); the quantized layers are kept
    srcs = [torch.nn.Linear(k, n, bias=False, device="meta") for n, k in shapes]
    bufs = [torch.randn(n, k, device="cuda", dtype=torch.bfloat16, generator=gen) for n, k in shapes]
    for m, b in zip(srcs, bufs):
        m.weight = torch.nn.Parameter(b, requires_grad=False)
    kept = []
    for m in srcs:  # warm-up
        MXInferenceLinear.from_float(m, qc)
    # the codes + scales of this rank's layers stay resident; take their memory from the driver BEFORE the timed region
    # (cudaMalloc of fresh segments is synchronous, ~1 ms per GB, and is not what is being measured)
    out_bytes = (hi - lo) * sum(n * k for n, k in shapes) * ((0.5 if args.wdtype == "float4_e2m1" else 1.0) + 1 / 32)
    reserve = torch.empty(int(out_bytes * 1.15) + (256 << 20), dtype=torch.uint8, device="cuda")
    # scale tensors below 1 MiB come from the allocator's small pool (2 MiB segments): reserve those too
    small = [n * k // 32 for n, k in shapes if n * k // 32 < (1 << 20)]
    reserve_small = [torch.empty(sz, dtype=torch.uint8, device="cuda") for _ in range(hi - lo + 2) for sz in small]
    del reserve, reserve_small
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_malloc0 = torch.cuda.memory_stats().get("num_device_alloc", 0)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(lo, hi):
        for m in srcs:
            kept.append(MXInferenceLinear.from_float(m, qc))
    host_issue_s = time.perf_counter() - t0
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    n_malloc = torch.cuda.memory_stats().get("num_device_alloc", 0) - n_malloc0
    elems = (hi - lo) * sum(n * k for n, k in shapes)
    bpe = 2 + (0.5 if args.wdtype == "float4_e2m1" else 1.0) + 1 / 32
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall, float(elems)], device="cuda", dtype=torch.float64)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {"mode": "quantize", "model": args.model, "layers": cfg["layers"], "world": world, "weights": args.wdtype,
            "weight_elements": int(t[2].item()), "gpu_s_max": round(float(tmax[0]), 4), "wall_s_max": round(float(tmax[1]), 4),
            "aggregate_GBps": round(float(t[2].item()) * bpe / float(tmax[0]) / 1e9, 1),
            "linears_per_rank": (hi - lo) * len(shapes), "host_issue_s_rank0": round(host_issue_s, 4), "cudaMalloc_calls_in_timed_region_rank0": n_malloc}


def default_args(**kw):
    """argparse defaults as a namespace (bench.py calls run_infer / run_quantize in-process on its own process group)"""
    base = dict(mode="infer", model="70b", layers=None, prefill=2048, batch=32, ctx=128, iters=5, wdtype="float4_e2m1", adtype="float8_e4m3",
                check=False, fused_max_rows=128, fused=False, also_fused=False, no_glue=False)
    base.update(kw)
    return argparse.Namespace(**base)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="infer", choices=["infer", "quantize"])
    ap.add_argument("--model", default="70b", choices=list(SHAPES))
    ap.add_argument("--layers", type=int, default=None)
    ap.add_argument("--prefill", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--wdtype", default="float4_e2m1")
    ap.add_argument("--adtype", default="float8_e4m3")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--fused-max-rows", type=int, default=128, help="fused all-reduce for at most this many tokens, NCCL above")
    ap.add_argument("--no-glue", action="store_true", help="plain PyTorch RMSNorm / rotary / SiLU between the projections, one launch per projection")
    ap.add_argument("--fused", action="store_true", help="row-parallel layers reduce in the GEMM epilogue (NVLink multicast) instead of NCCL")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if "RANK" not in os.environ:
        os.environ.update(RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    world, rank = dist.get_world_size(), dist.get_rank()
    try:
        res = run_infer(args, world, rank) if args.mode == "infer" else run_quantize(args, world, rank)
    except Exception:
        import traceback
        print(f"[rank {rank}] " + traceback.format_exc()[-1500:], flush=True)
        raise
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
