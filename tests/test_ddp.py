"""CPU: the training config's gradient all-reduce (ccvpe_b200/ddp.py) with a real 2-process gloo group: bucketed,
hook-driven, averaged gradients equal the mean of the per-shard gradients; frozen `_fc` tensors; replicas stay in sync
through optimizer steps ("after all-reduce == mean of per-shard reference grads")."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from ccvpe_b200.ddp import GradientAllReducer, freeze_unused, ready_order


class _Enc(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 16)
        self.b = nn.Linear(16, 16)
        self._fc = nn.Linear(16, 3)            # never used in forward, like the encoders' classifier head

    def forward(self, x):
        return torch.tanh(self.b(torch.relu(self.a(x))))


class _Toy(nn.Module):
    """Same top-level parameter naming as the CVM_* models: two encoders + decoder tensors."""

    def __init__(self):
        super().__init__()
        self.grd_efficientnet = _Enc()
        self.sat_efficientnet = _Enc()
        self.deconv6 = nn.Linear(32, 8)
        self.conv1 = nn.Linear(8, 1)

    def forward(self, g, s):
        return self.conv1(torch.relu(self.deconv6(torch.cat([self.grd_efficientnet(g), self.sat_efficientnet(s)], dim=1))))


def _data(n):
    gen = torch.Generator().manual_seed(3)
    return torch.randn(n, 6, generator=gen), torch.randn(n, 6, generator=gen), torch.randn(n, 1, generator=gen)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _Toy()
    red = GradientAllReducer(model, bucket_mb=0.001, first_bucket_mb=0.0005)     # tiny buckets: several per step
    assert len(red.buckets) >= 3 and sorted(red.frozen) == ["grd_efficientnet._fc.bias", "grd_efficientnet._fc.weight",
                                                            "sat_efficientnet._fc.bias", "sat_efficientnet._fc.weight"]
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-2)
    g, s, y = _data(8)
    lo, hi = rank * 4, rank * 4 + 4
    first = None
    for step in range(3):
        red.zero_grad()
        loss = ((model(g[lo:hi], s[lo:hi]) - y[lo:hi]) ** 2).mean()
        loss.backward()
        red.finish()
        if step == 0:
            first = {k: p.grad.clone().numpy() for k, p in model.named_parameters() if p.requires_grad}
        opt.step()
    np.savez(os.path.join(out_dir, "grads%d.npz" % rank), **first)
    np.savez(os.path.join(out_dir, "params%d.npz" % rank), **{k: p.detach().numpy() for k, p in model.named_parameters()})
    dist.destroy_process_group()


def test_allreduced_gradients_equal_mean_of_shard_gradients(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-process reference: mean over the two shards' gradients == gradient of the mean of the shard losses
    torch.manual_seed(0)
    model = _Toy()
    freeze_unused(model)
    g, s, y = _data(8)
    loss = 0.5 * (((model(g[:4], s[:4]) - y[:4]) ** 2).mean() + ((model(g[4:], s[4:]) - y[4:]) ** 2).mean())
    loss.backward()
    g0 = np.load(os.path.join(str(tmp_path), "grads0.npz"))
    g1 = np.load(os.path.join(str(tmp_path), "grads1.npz"))
    for k, p in model.named_parameters():
        if not p.requires_grad:
            assert k not in g0.files
            continue
        np.testing.assert_allclose(g0[k], p.grad.numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_array_equal(g0[k], g1[k])                     # identical on every rank
    p0 = np.load(os.path.join(str(tmp_path), "params0.npz"))
    p1 = np.load(os.path.join(str(tmp_path), "params1.npz"))
    for k in p0.files:
        np.testing.assert_array_equal(p0[k], p1[k])                     # replicas stay bit-identical through Adam steps


def test_ready_order_puts_decoder_first_and_reverses_encoders():
    model = _Toy()
    freeze_unused(model)
    names = {id(p): n for n, p in model.named_parameters()}
    order = [names[id(p)] for p in ready_order(model)]
    assert order[:4] == ["deconv6.weight", "deconv6.bias", "conv1.weight", "conv1.bias"]
    assert order[4] == "sat_efficientnet.b.bias" and order[-1] == "grd_efficientnet.a.weight"
    assert not any("_fc" in n for n in order)
