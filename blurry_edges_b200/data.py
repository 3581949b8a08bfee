"""Host-to-device data path next to the loss.

The reference's ShapeDataset keeps every array in pageable host memory and moves ONE SAMPLE at a time in __getitem__
(data/dataset.py:40-56: up to six `.to(device)` calls and two divisions by alpha per sample); the DataLoader then stacks the
samples on the device.  At batch 8 that is 48 small pageable copies, 16 elementwise launches and 6 stack launches per step, all on
the compute stream, in front of a loss step that takes 0.7 ms.

`ShapePrefetcher` is a drop-in for `DataLoader(ShapeDataset(...), batch_size, shuffle, drop_last)`: it takes the SAME dataset object,
copies its tensors once into pinned host memory, and per batch gathers the rows into a pinned staging buffer and issues ONE
asynchronous copy per field on its own stream, one batch ahead of the consumer (double buffered), so the copy of batch k+1 overlaps
the compute of batch k.  It yields the same tuples in the same layouts:
  mode 'global'      (input_param, img_ny / alpha, img_gt / alpha, bndry_dist, deri, bndry_depth)
  mode 'local'       (img_ny / alpha, img_gt / alpha, bndry_dist, deri)
  mode 'global_pre'  img_ny / alpha"""
from __future__ import annotations

import torch

_FIELDS = {'global': ('input_param', 'img_ny', 'img_gt', 'bndry_dist', 'deri', 'bndry_depth'),
           'local': ('img_ny', 'img_gt', 'bndry_dist', 'deri'),
           'global_pre': ('img_ny',)}
_SCALED = ('img_ny', 'img_gt')


class ShapePrefetcher:
    def __init__(self, dataset, batch_size, device=None, shuffle=False, drop_last=True, generator=None, pin=None):
        self.mode = dataset.mode
        self.fields = _FIELDS[self.mode]
        self.device = torch.device(device if device is not None else dataset.device)
        self.batch_size, self.shuffle, self.drop_last, self.generator = int(batch_size), shuffle, drop_last, generator
        self.pin = (self.device.type == 'cuda') if pin is None else bool(pin)
        src = {f: getattr(dataset, f).to(torch.float32).contiguous() for f in self.fields}
        src['alpha'] = dataset.alpha.to(torch.float32).contiguous()
        self.n = src['img_ny'].shape[0]
        self.host = {k: (v.pin_memory() if self.pin else v) for k, v in src.items()}
        self.stage = [{k: self._empty((self.batch_size,) + tuple(v.shape[1:])) for k, v in self.host.items()} for _ in range(2)]
        self.stream = torch.cuda.Stream(self.device) if self.device.type == 'cuda' else None
        self.staged_free = [None, None]        # event: the H2D copies out of staging buffer i have finished

    def _empty(self, shape):
        t = torch.empty(shape, dtype=torch.float32)
        return t.pin_memory() if self.pin else t

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def _batches(self):
        order = torch.randperm(self.n, generator=self.generator) if self.shuffle else torch.arange(self.n)
        for k in range(len(self)):
            yield order[k * self.batch_size:(k + 1) * self.batch_size]

    def _issue(self, idx, slot):
        """Gather rows `idx` into staging buffer `slot` and start their copies; returns (device tensors, ready event)."""
        nb = idx.numel()
        st = self.stage[slot]
        if self.staged_free[slot] is not None:
            self.staged_free[slot].synchronize()                   # the previous user of this staging buffer has left it
        for k, v in self.host.items():
            torch.index_select(v, 0, idx, out=st[k][:nb])
        if self.stream is None:
            dev = {k: st[k][:nb].clone() for k in st}
            return dev, None
        with torch.cuda.stream(self.stream):
            dev = {k: st[k][:nb].to(self.device, non_blocking=True) for k in st}
            a = dev['alpha'].view(nb, *([1] * (dev['img_ny'].dim() - 1)))
            for k in _SCALED:
                if k in dev:
                    dev[k] = dev[k] / a                             # data/dataset.py:50,52,56
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.staged_free[slot] = ev
        return dev, ev

    def _finish(self, dev, ev):
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in dev.values():
                t.record_stream(torch.cuda.current_stream(self.device))
        else:
            a = dev['alpha'].view(-1, *([1] * (dev['img_ny'].dim() - 1)))
            for k in _SCALED:
                if k in dev:
                    dev[k] = dev[k] / a
        out = tuple(dev[f] for f in self.fields)
        return out[0] if len(out) == 1 else out

    def __iter__(self):
        it = self._batches()
        nxt = next(it, None)
        pending = self._issue(nxt, 0) if nxt is not None else None
        slot = 1
        while pending is not None:
            nxt = next(it, None)
            ahead = self._issue(nxt, slot) if nxt is not None else None     # batch k+1 is on its way while the caller works on batch k
            yield self._finish(*pending)
            pending, slot = ahead, slot ^ 1
