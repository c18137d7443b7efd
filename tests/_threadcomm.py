"""World-size-N collectives between THREADS of one process sharing one GPU: the interface of
genomic_gnn_b200.dist.TorchComm, so that the sharded (multi-rank) code paths of the graph build -- read shards with
global positions, KF row-block shards, the score-class table reduction, the spread of the cut-off class over ranks --
run on the single-GPU test box.  It stands in for NCCL, not for the product: the real exchange is covered by
tests/test_multi_gpu.py on a multi-GPU lease and by the gloo tests on CPU.

Every rank is a thread with its own CUDA stream; a collective is `publish my tensor, barrier, read everyone's,
barrier`, with the streams synchronised at the barriers."""
import threading

import torch


class ThreadWorld:
    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world
        self.errors = []

    def run(self, fn):
        """fn(comm) on every rank; returns the per-rank results (re-raises the first failure)."""
        results = [None] * self.world

        def body(rank):
            try:
                with torch.cuda.stream(torch.cuda.Stream()):
                    results[rank] = fn(ThreadComm(self, rank))
                    torch.cuda.current_stream().synchronize()
            except BaseException as exc:  # noqa: BLE001
                self.errors.append((rank, exc))
                self.barrier.abort()

        threads = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        real = [e for e in self.errors if not isinstance(e[1], threading.BrokenBarrierError)]
        if real or self.errors:
            raise (real or self.errors)[0][1]
        return results


class ThreadComm:
    def __init__(self, tw, rank):
        self.tw, self.rank, self.world = tw, rank, tw.world
        self.group = None
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _exchange(self, value):
        """Everyone's value, in rank order (tensors: the producing stream is drained first)."""
        torch.cuda.current_stream().synchronize()
        self.tw.slots[self.rank] = value
        self.tw.barrier.wait()
        everyone = list(self.tw.slots)
        self.tw.barrier.wait()
        return everyone

    def sum_int(self, x):
        return sum(int(v) for v in self._exchange(int(x)))

    def gather_ints(self, x):
        return [int(v) for v in self._exchange(int(x))]

    def sum_tensor_(self, t):
        parts = self._exchange(t.clone())
        t.copy_(torch.stack(parts).sum(dim=0))
        torch.cuda.current_stream().synchronize()
        self.tw.barrier.wait()
        return t

    def min_tensor_(self, t):
        parts = self._exchange(t.clone())
        t.copy_(torch.stack(parts).amin(dim=0))
        torch.cuda.current_stream().synchronize()
        self.tw.barrier.wait()
        return t

    def gather_device_ints(self, t):
        return [p.cpu().tolist() for p in self._exchange(t.contiguous().clone())]

    def all_gather_into(self, out, mine):
        parts = self._exchange(mine.clone())
        out.copy_(torch.cat([p.reshape(-1) for p in parts]))
        torch.cuda.current_stream().synchronize()
        self.tw.barrier.wait()
        return out

    def all_gather_padded_(self, buf, stride):
        mine = buf[self.rank * stride: (self.rank + 1) * stride]
        return self.all_gather_into(buf, mine)

    def gather_cat(self, t, dim=0):
        return torch.cat(self._exchange(t.clone()), dim=dim).contiguous()

    def gather_packed(self, tensors, sizes, dim_list):
        everyone = self._exchange([t.clone() for t in tensors])
        return [torch.cat([everyone[r][i] for r in range(self.world)], dim=d).contiguous()
                for i, d in enumerate(dim_list)]

    def exchange_slices(self, tensors, sizes):
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        mine = [t[offs[self.rank]: offs[self.rank + 1]].clone() for t in tensors]
        everyone = self._exchange(mine)
        for r in range(self.world):
            if r == self.rank:
                continue
            for t, piece in zip(tensors, everyone[r]):
                t[offs[r]: offs[r + 1]].copy_(piece)
        torch.cuda.current_stream().synchronize()
        self.tw.barrier.wait()
        return []
