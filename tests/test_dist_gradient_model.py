"""CPU model of the distributed inverse behind gpb_dist_lml_grad (csrc/dist.cu: dist_inverse_rows) -- the block-cyclic
bookkeeping of the N > 1 path, checked without a GPU for world sizes the GPU tests cannot reach.

The model restates the HOST logic of dist_inverse_rows with numpy in place of the kernels: which row blocks a rank owns,
which rows are active at panel j, where K^-1[a, b] is written (over the zero part of the row stack / into the diagonal
buffer), from which column a product may start (the block-structured GEMM_TRIK_A), what the owner packs and broadcasts.
It runs (a) with virtual ranks in one process for world = 1, 2, 3, 4, 8 including ragged last blocks, and (b) as two
real processes exchanging the panels and row blocks with gloo broadcasts.  The result must be the lower triangle of
inv(K); the traces of the gradient are then sums over disjoint row blocks, one share per rank.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def spd(n, seed):
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 1, (n, 2))
    d2 = ((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    return np.exp(-0.5 * d2 / 0.3**2) + 0.05**2 * np.eye(n)


class Rank:
    """One rank's state and steps; `bcast(buf_or_None, root)` returns the root's buffer on every rank."""

    def __init__(self, L, nbd, me, G):
        self.L, self.nbd, self.me, self.G = L, nbd, me, G
        self.n = L.shape[0]
        self.nblk = -(-self.n // nbd)
        self.na = (self.nblk - 1 - me) // G + 1 if self.nblk > me else 0          # dist.cu: na
        self.Y = np.zeros((self.na * nbd, self.n))                                 # row stack (ystack)
        self.kdiag = np.zeros((max(self.na, 1), nbd, nbd))
        for idx in range(self.na):                                                  # unit_rows_kernel
            a = me + G * idx
            for i in range(self.cols_of(a)):
                self.Y[idx * nbd + i, a * nbd + i] = 1.0

    def cols_of(self, j):
        return min(self.nbd, self.n - j * self.nbd)

    def owner(self, j):
        return j % self.G

    # ---- phase 1: one streamed panel
    def panel_to_send(self, j):                     # rows j nbd .. n of block column j (the owner's storage)
        return self.L[j * self.nbd:, j * self.nbd:j * self.nbd + self.cols_of(j)].copy()

    def forward_step(self, j, P):
        nbd, me, G = self.nbd, self.me, self.G
        cj = self.cols_of(j)
        cnt = (j - me) // G + 1 if j >= me else 0                                    # owned row blocks a <= j
        rows = cnt * nbd
        if rows == 0:
            return
        Ljj = P[:cj]
        X = np.linalg.solve(Ljj, self.Y[:rows, j * nbd:j * nbd + cj].T).T           # X L_jj^-T
        self.Y[:rows, j * nbd:j * nbd + cj] = X
        end = j * nbd + cj
        if end < self.n:
            self.Y[:rows, end:] -= X @ P[cj:].T

    # ---- phase 2: one streamed row block of Y
    def block_to_send(self, b):                     # packed from its first non-zero column
        idx = (b - self.me) // self.G
        return self.Y[idx * self.nbd:idx * self.nbd + self.cols_of(b), b * self.nbd:].copy()

    def product_step(self, b, P):
        nbd, me, G, na = self.nbd, self.me, self.G, self.na
        cb = self.cols_of(b)
        idx = 0 if b <= me else (b - me + G - 1) // G                                # first owned row block with a >= b
        if idx < na and me + G * idx == b:                                           # diagonal block, k >= b nbd
            ya = self.Y[idx * nbd:idx * nbd + cb, b * nbd:]
            self.kdiag[idx, :cb, :cb] = np.tril(ya @ P.T)
            idx += 1
        if idx < na:
            k0 = (me + G * idx) * nbd
            rows0 = idx * nbd
            out = np.zeros((self.Y.shape[0] - rows0, cb))
            for i in range(na - idx):               # block-structured GEMM_TRIK_A: row block i starts at k0 + i G nbd
                kb = k0 + i * G * nbd
                if kb >= self.n:
                    continue
                ya = self.Y[rows0 + i * nbd:rows0 + (i + 1) * nbd, kb:]
                out[i * nbd:(i + 1) * nbd] = ya @ P[:, kb - b * nbd:].T
            assert not self.Y[rows0:, b * nbd:b * nbd + cb].any()                     # written over zeros only
            assert k0 >= b * nbd + cb                                                # never over a region still read
            self.Y[rows0:, b * nbd:b * nbd + cb] = out

    def lower_rows(self):
        """(global row index, values left of and on the diagonal) for every owned row"""
        rows = {}
        for idx in range(self.na):
            a = self.me + self.G * idx
            for i in range(self.cols_of(a)):
                g = a * self.nbd + i
                v = np.zeros(g + 1)
                v[:a * self.nbd] = self.Y[idx * self.nbd + i, :a * self.nbd]
                v[a * self.nbd:] = self.kdiag[idx, i, :i + 1]
                rows[g] = v
        return rows


def run_virtual(K, nbd, G):
    L = np.linalg.cholesky(K)
    ranks = [Rank(L, nbd, r, G) for r in range(G)]
    nblk = ranks[0].nblk
    for j in range(nblk):
        P = ranks[j % G].panel_to_send(j)
        for r in ranks:
            r.forward_step(j, P)
    for b in range(nblk):
        P = ranks[b % G].block_to_send(b)
        for r in ranks:
            r.product_step(b, P)
    rows = {}
    for r in ranks:
        rows.update(r.lower_rows())
    return rows


@pytest.mark.parametrize("n,nbd,G", [(48, 8, 1), (48, 8, 2), (56, 8, 3), (64, 8, 4), (72, 8, 8), (60, 16, 2), (44, 8, 3), (24, 8, 8)])
def test_inverse_rows_model_virtual_ranks(n, nbd, G):
    K = spd(n, n + G)
    rows = run_virtual(K, nbd, G)
    Kinv = np.linalg.inv(K)
    assert sorted(rows) == list(range(n))                                            # every row owned exactly once
    err = max(np.abs(rows[g] - Kinv[g, :g + 1]).max() for g in range(n))
    assert err < 1e-9 * np.abs(Kinv).max()


def test_gradient_traces_are_sums_of_per_rank_shares():
    """1/2 sum (alpha alpha^T - K^-1) o dK over the lower triangle (strict lower counted twice) splits into disjoint
    per-rank sums over the owned rows -- the all-reduce of gpb_dist_lml_grad."""
    n, nbd, G = 40, 8, 3
    K = spd(n, 5)
    rng = np.random.default_rng(0)
    alpha = rng.normal(size=n)
    dK = rng.normal(size=(n, n))
    dK = dK + dK.T
    Q = np.outer(alpha, alpha) - np.linalg.inv(K)
    want = 0.5 * (Q * dK).sum()
    L = np.linalg.cholesky(K)
    ranks = [Rank(L, nbd, r, G) for r in range(G)]
    for j in range(ranks[0].nblk):
        P = ranks[j % G].panel_to_send(j)
        [r.forward_step(j, P) for r in ranks]
    for b in range(ranks[0].nblk):
        P = ranks[b % G].block_to_send(b)
        [r.product_step(b, P) for r in ranks]
    shares = []
    for r in ranks:
        s = 0.0
        for g, v in r.lower_rows().items():
            q = alpha[g] * alpha[:g + 1] - v
            w = np.full(g + 1, 2.0)
            w[g] = 1.0
            s += 0.5 * (w * q * dK[g, :g + 1]).sum()
        shares.append(s)
    assert abs(sum(shares) - want) < 1e-9 * abs(want)


def _gloo_worker(rank, world, port, n, nbd, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    K = spd(n, 11)
    L = np.linalg.cholesky(K)
    me = Rank(L, nbd, rank, world)

    def bcast(make, shape, root):
        t = torch.from_numpy(np.ascontiguousarray(make())) if rank == root else torch.empty(shape, dtype=torch.float64)
        dist.broadcast(t, src=root)
        return t.numpy()

    for j in range(me.nblk):
        P = bcast(lambda: me.panel_to_send(j), (n - j * nbd, me.cols_of(j)), j % world)
        me.forward_step(j, P)
    for b in range(me.nblk):
        P = bcast(lambda: me.block_to_send(b), (me.cols_of(b), n - b * nbd), b % world)
        me.product_step(b, P)
    rows = me.lower_rows()
    gathered = [None] * world
    dist.all_gather_object(gathered, {g: v.tolist() for g, v in rows.items()})
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_inverse_rows_model_two_processes_over_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    n, nbd = 52, 8                                     # 7 blocks, the last one ragged
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, nbd, out)) for r in range(2)]
    [p.start() for p in procs]
    gathered = out.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows = {}
    for part in gathered:
        rows.update({int(g): np.asarray(v) for g, v in part.items()})
    Kinv = np.linalg.inv(spd(n, 11))
    assert sorted(rows) == list(range(n))
    assert max(np.abs(rows[g] - Kinv[g, :g + 1]).max() for g in range(n)) < 1e-9 * np.abs(Kinv).max()
