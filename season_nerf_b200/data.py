"""Device-resident ray table feeding the training step ("next" row 2 of the scope table).

The reference feeds `train_step` from a torch DataLoader over a [n,22] float32 table of rays (NN_loaders/mg_Color_Loader.py:40-104,
`Net_tool.get_data` / `data_to_dict`, mg_run_NeRF.py:122-133,229-264): 4 worker processes, CPU batches, one H2D per step.
Here the table lives in HBM once (22 floats per ray: 88 MB per million rays) and a batch is a gather with a slice of a
per-epoch random permutation - no host work per step, same dictionary keys and column layout."""
import torch as t

COLUMNS = {"Img_Pt": (0, 2), "Top": (2, 5), "Bot": (5, 8), "View_Angle": (8, 11), "Sun_Angle": (11, 14),
           "Time_Encoded": (14, 18), "Sample_Weight": (18, 19), "GT_Color": (19, None)}


def data_to_dict(data):
    """mg_run_NeRF.py:122-133: [B,22] rows -> dict of column views."""
    return {k: data[:, lo:hi] for k, (lo, hi) in COLUMNS.items()}


class RayTable:
    """Shuffled mini-batches of a [n, 22] ray table (DataLoader(shuffle=True, drop_last=False) semantics: every ray once
    per epoch, a short last batch, new permutation each epoch).  `device` may be a CUDA device (the intended use) or
    "cpu" (host-logic tests)."""

    def __init__(self, table, batch_size, device, seed=0, drop_last=False):
        self.table = t.as_tensor(table, dtype=t.float32).to(device).contiguous()
        if self.table.dim() != 2 or self.table.shape[1] < 22:
            raise ValueError("ray table must be [n, >=22] (mg_run_NeRF.py:122-133 column layout)")
        self.batch_size, self.drop_last = int(batch_size), drop_last
        self.gen = t.Generator(device=self.table.device)
        self.gen.manual_seed(seed)
        self.epoch = 0
        self._perm, self._pos = None, 0

    def __len__(self):
        n = self.table.shape[0]
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def next_batch(self):
        """-> (data_dict, is_eof) like Net_tool.get_data: is_eof is True on the first batch of a new epoch after the first."""
        n = self.table.shape[0]
        is_eof = False
        if self._perm is None or self._pos >= n or (self.drop_last and self._pos + self.batch_size > n):
            is_eof = self._perm is not None
            self._perm = t.randperm(n, device=self.table.device, generator=self.gen)
            self._pos = 0
            self.epoch += int(is_eof)
        idx = self._perm[self._pos:self._pos + self.batch_size]
        self._pos += self.batch_size
        return data_to_dict(self.table[idx]), is_eof

    @staticmethod
    def shard(table, rank, world_size):
        """rows of this rank (contiguous split of the table; every rank then shuffles its own rows)"""
        from .train import shard_range
        lo, hi = shard_range(len(table), rank, world_size)
        return table[lo:hi]
