"""Multi-GPU host logic on CPU (gloo, world_size 2): each rank takes its contiguous shard of the seeded instance stream,
solves it (here with the CPU oracle standing in for the GPU solver -- test infrastructure), and ONE all-gather of the
optimal control vectors gives every rank the whole batch.  Checks the shard arithmetic of bench.py:
rank r solves instances [r*B, (r+1)*B) and the gathered result equals the single-process result bit for bit."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_local_planner_b200 import configs


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py as orc
    cfg = configs.cfg2(tol=1e-6); cfg.max_iter = 40
    data = configs.generate(2, B, first=rank * B)
    out = orc.step_batch(cfg, data, n_threads=1)
    N = cfg.n
    send = torch.from_numpy(out["u_seq"][:, : N - 1, :].reshape(-1).copy())
    recv = torch.empty(world * send.numel(), dtype=torch.float64)
    dist.all_gather_into_tensor(recv, send)
    conv = torch.tensor([float((out["status"] == 0).sum())], dtype=torch.float64)
    dist.all_reduce(conv)
    if rank == 0:
        q.put((recv.numpy().copy(), float(conv[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_allgather_two_ranks():
    B, world = 6, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, conv = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import oracle_py as orc
    cfg = configs.cfg2(tol=1e-6); cfg.max_iter = 40
    full = orc.step_batch(cfg, configs.generate(2, world * B), n_threads=1)
    ref = full["u_seq"][:, : cfg.n - 1, :].reshape(-1)
    np.testing.assert_array_equal(gathered, ref)  # instance i identical for every batch size / rank count
    assert conv == float((full["status"] == 0).sum())
