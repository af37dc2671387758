"""Prefetching feed for `detector[blob]` — the input side of the hot path.

The reference's loader (dataloaders/visual_genome.py:264-424 + dataloaders/blob.py:155-229) collates in DataLoader worker
processes and then copies each batch to its GPU with `.cuda(async=True)` on the COMPUTE stream right before forward, so
the 25 MB image copy of a 6-image batch (0.45 ms over PCIe 5) sits in front of every step. Here one background thread
per rank pins the next batches and issues their H2D copies on a dedicated copy stream `depth` steps ahead; the blob the
training loop receives already has its device tensors, and its `scatter()` (which `RelModel.__getitem__` calls, as the
reference's does) only makes the compute stream wait for the copy's event.

    for blob in PrefetchLoader(batches, device, depth=2):      # batches: iterable of numpy batch dicts (make_numpy_batch)
        result = detector[blob]                                # models/train_rels.py:137, unchanged

The HDF5 / JSON reader of Visual Genome itself (dataloaders/visual_genome.py:27-262) is out of scope here (no dataset,
no h5py in this environment): anything that yields the numpy batch dict can be plugged in as `batches`."""
import queue
import threading

import torch

from dataloaders.synthetic import SyntheticBlob


class _Prefetched(SyntheticBlob):
    """A SyntheticBlob whose device copies were issued on a copy stream; scatter() = wait for them."""

    def __init__(self, nb, device, is_train=True):
        super().__init__(nb, device, is_train)
        self._ready = None

    def issue(self, copy_stream):
        if self.device.type != "cuda":
            self.dev = dict(self.host)
            return
        with torch.cuda.stream(copy_stream):
            SyntheticBlob.scatter(self)
            self._ready = torch.cuda.Event()
            self._ready.record(copy_stream)

    def scatter(self):
        if self.dev is None:
            raise RuntimeError("prefetched blob used before its copies were issued")
        if self._ready is not None:
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self._ready)
            for t in self.dev.values():          # the caching allocator must not recycle these under the compute stream
                t.record_stream(cur)
            self._ready = None


class PrefetchLoader(object):
    _END = object()

    def __init__(self, batches, device, depth=2, is_train=True):
        self.batches, self.device, self.depth, self.is_train = batches, torch.device(device), max(1, int(depth)), is_train

    def __iter__(self):
        q = queue.Queue(maxsize=self.depth)
        stop = threading.Event()
        copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

        def work():
            try:
                if self.device.type == "cuda":
                    torch.cuda.set_device(self.device)
                for nb in self.batches:
                    if stop.is_set():
                        return
                    blob = _Prefetched(nb, self.device, self.is_train)        # pins the host tensors (host time, off the main thread)
                    blob.issue(copy_stream)
                    while not stop.is_set():
                        try:
                            q.put(blob, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                q.put(self._END)
            except BaseException as e:                                        # surfaces in the consumer
                q.put(e)

        th = threading.Thread(target=work, daemon=True, name="mb200-prefetch")
        th.start()
        try:
            while True:
                item = q.get()
                if item is self._END:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:
            stop.set()
            while not q.empty():
                try:
                    q.get_nowait()
                except queue.Empty:
                    break
            th.join(timeout=5.0)
