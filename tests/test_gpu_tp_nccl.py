"""GPU tier, >= 2 GPUs (skipped on a single-GPU box): the tensor-parallel row-parallel layer over a real NCCL group, with the
NCCL all-reduce and with the fused GEMM + all-reduce epilogue (`FusedAllReducePool`: `multimem.red` adds of the bf16 partial tile
into every rank's symmetric buffer through the NVLink multicast address, csrc/mxq_gemm.cu / mxq_gemm_skinny.cu).

Parity (reference layer: torchmx/layers/mx_linear.py:61-95 on the unsharded weight): every rank contributes the bf16 rounding
of its partial product, so  |y - ref| <= 2^-8 |ref| + 2^-8 sum_r |partial_r| + 2^-18 sum_k |a_k b_k|  against the fp64
contraction, for both reductions.  Fused vs NCCL: with TWO ranks a sum of two bf16 values has one rounding whatever the order,
so the results are bit-identical; with more ranks the switch adds the partials in arrival order, one bf16 rounding per add --
NOT deterministic, bounded by (world - 1) bf16 ulps of the largest partial sum (asserted below, documented in DESIGN.md §6).
"""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
        import torchmx_b200  # noqa: F401
        from torchmx_b200 import dtypes, mx_gemm
        from torchmx_b200.config import MXConfig, QLinearConfig
        from torchmx_b200.layers.tp_linear import FusedAllReducePool, RowParallelMXLinear, shard_bounds
        from torchmx_b200.mx_tensor import MXTensor
        qc = QLinearConfig(weights_config=MXConfig("float4_e2m1", 32), activations_config=MXConfig("float8_e4m3", 32))
        K, N = 1024 * world, 1024
        torch.manual_seed(0)  # same full layer and inputs on every rank
        lin = torch.nn.Linear(K, N, bias=True).to(dev, torch.bfloat16)
        layer = RowParallelMXLinear.from_float(lin, qc)
        lo, hi = shard_bounds(K, world, rank, 32)
        pool = FusedAllReducePool(N, 512, fused_max_rows=512)
        report = {}
        for rows in (32, 100, 384):  # skinny kernel (<= 64: fused activation quantization; <= 128) and the CTA-pair kernel
            x = torch.randn(rows, K, device=dev, dtype=torch.bfloat16)
            xs = x[:, lo:hi].contiguous()
            layer.__dict__.pop("_fused_pool", None)
            y_nccl = layer(xs).clone()
            layer.enable_fused_allreduce(pool)
            before = mx_gemm.stats["fused_allreduce"]
            y_fused = layer(xs).clone()
            assert mx_gemm.stats["fused_allreduce"] == before + 1, "the fused epilogue was not taken"
            # reference: fp64 contraction of the unsharded quantized operands; partial magnitudes gathered from all ranks
            X = MXTensor.to_mx(x, dtypes.float8_e4m3, 32).to_dtype(torch.float32).double()
            W = MXTensor.to_mx(lin.weight.data, dtypes.float4_e2m1, 32).to_dtype(torch.float32).double()
            ref = X @ W.t() + lin.bias.double()
            S = X.abs() @ W.abs().t() + lin.bias.double().abs()
            part = (X[:, lo:hi] @ W[:, lo:hi].t()).abs()
            dist.all_reduce(part)
            tol = 2.0 ** -8 * ref.abs() + 2.0 ** -8 * part + 2.0 ** -18 * S
            ok_nccl = bool(((y_nccl.double() - ref).abs() <= tol).all())
            ok_fused = bool(((y_fused.double() - ref).abs() <= tol + (world - 2) * 2.0 ** -8 * part).all())
            same = bool(torch.equal(y_nccl, y_fused))
            close = bool(((y_nccl.double() - y_fused.double()).abs() <= (world - 1) * 2.0 ** -8 * part + 1e-30).all())
            report[rows] = (ok_nccl, ok_fused, same, close)
        torch.cuda.synchronize()
        dist.barrier()
        q.put((rank, "ok", report))
        dist.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "error", traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_parallel_nccl_and_fused_allreduce(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, status, report in results:
        assert status == "ok", f"rank {rank}: {report}"
        for rows, (ok_nccl, ok_fused, same, close) in report.items():
            assert ok_nccl, f"rank {rank} rows {rows}: NCCL result outside tolerance"
            assert ok_fused, f"rank {rank} rows {rows}: fused result outside tolerance"
            assert close, f"rank {rank} rows {rows}: fused and NCCL results differ by more than the partial-rounding bound"
            if world == 2:
                assert same, f"rank {rank} rows {rows}: two summands must give bit-identical fused / NCCL results"
