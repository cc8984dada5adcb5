"""CPU: host logic of the ensemble path (member enumeration, partition over ranks, member
initialisation, result gather over gloo with world_size 2).  No compute call into libgphm."""
import importlib
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

E = importlib.import_module("gaussian-process-slover-for-high-freq-pde_b200.ensemble")


def test_members_and_partition():
    members = E.ensemble_members()
    assert len(members) == 512 and members[0] == (0, 10) and members[8] == (1, 10) and members[-1] == (63, 80)
    for n, world in ((512, 8), (512, 3), (5, 8), (0, 2), (7, 1)):
        parts = [list(E.shard_members(n, r, world)) for r in range(world)]
        assert sum(parts, []) == list(range(n))                          # disjoint, ordered, complete
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    try:
        E.shard_members(4, 4, 4)
        assert False
    except ValueError:
        pass


class _StubModel(object):
    def __init__(self, keys, Q=6):
        self.keys, self.Q = keys, Q

    def init_params(self):
        kp = lambda: {"log-w": np.log(1 / self.Q) * np.ones(self.Q), "log-ls": np.zeros(self.Q),
                      "freq": np.linspace(0, 1, self.Q) * 20.0}
        return dict({"log_tau": 0.0, "log_v": 0.0}, **{k: kp() for k in self.keys})


def test_member_init_is_deterministic_and_scaled():
    for keys in (("kernel_paras",), ("kernel_paras_1", "kernel_paras_2")):
        m = _StubModel(keys)
        a, b, c = E.member_init(m, 3, 40), E.member_init(m, 3, 40), E.member_init(m, 4, 40)
        for k in keys:
            assert np.array_equal(a[k]["freq"], b[k]["freq"]) and not np.array_equal(a[k]["freq"], c[k]["freq"])
            assert np.abs(a[k]["freq"] - np.linspace(0, 1, m.Q) * 40).max() < 0.06      # 0.01 N(0,1) around the scaled grid
            assert np.array_equal(a[k]["log-ls"], np.zeros(m.Q))
        if len(keys) == 2:                                                             # one generator: the axes differ
            assert not np.array_equal(a[keys[0]]["freq"], a[keys[1]]["freq"])


def _worker(rank, world, port, sizes):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for n in sizes:                                       # 1: rank 1 holds no member at all
            mine = E.shard_members(n, rank, world)
            local = torch.tensor([[float(i), 10.0 * i] for i in mine], dtype=torch.float64).reshape(len(mine), 2)
            full = E.gather_results(local, n, rank, world)
            assert full.shape == (n, 2)
            assert torch.equal(full[:, 0], torch.arange(n, dtype=torch.float64)) and torch.equal(full[:, 1], 10 * full[:, 0])
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_gather_results_world2_gloo():
    mp.spawn(_worker, args=(2, _free_port(), (7, 8, 1)), nprocs=2, join=True)


def test_gather_results_world1():
    x = torch.arange(6, dtype=torch.float64).reshape(3, 2)
    assert torch.equal(E.gather_results(x, 3, 0, 1), x)
