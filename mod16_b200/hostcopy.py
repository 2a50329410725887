'''
Fast staging of large pageable host arrays (plain numpy inputs of the drop-in API).

``tensor.to(device)`` from pageable memory is a synchronous bounce copy (measured on the
B200 boxes: about 10 GB/s, one array after the other).  ``upload`` instead cuts the arrays
into chunks, has a few worker threads copy each chunk into a pinned staging buffer (numpy's
``copyto`` releases the GIL) and launch its asynchronous host-to-device copy on the worker's
own stream, so the CPU copies of several chunks and the DMA of others overlap;
``download`` does the reverse for results.  The staging ring (pinned memory) is allocated once
per process and device and reused.

Only the host plumbing of the product: the numbers that reach the kernels are bit-identical
to what a plain copy would deliver.
'''
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

CHUNK_BYTES = 32 << 20
N_BUFFERS = 12
N_WORKERS = 6
THRESHOLD_BYTES = 64 << 20   # smaller inputs take the plain path

_lock = threading.Lock()
_rings = {}


class _Ring(object):
    def __init__(self, device):
        self.device = device
        self.free = queue.Queue()
        for _ in range(N_BUFFERS):
            buf = torch.empty(CHUNK_BYTES, dtype=torch.uint8).pin_memory()
            self.free.put((buf, None))
        self.pool = ThreadPoolExecutor(N_WORKERS)
        self.local = threading.local()

    def stream(self):
        s = getattr(self.local, 'stream', None)
        if s is None:
            s = torch.cuda.Stream(self.device)
            self.local.stream = s
        return s


def _ring(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    with _lock:
        r = _rings.get(key)
        if r is None:
            r = _Ring(torch.device('cuda', key))
            _rings[key] = r
    return r


def worthwhile(a):
    return isinstance(a, np.ndarray) and a.nbytes >= THRESHOLD_BYTES and a.flags.c_contiguous \
        and a.dtype in (np.float32, np.float64)


def upload(arrays, device, stream):
    '''
    Copies C-contiguous numpy ``arrays`` to new device tensors of the same shape and dtype.
    The copies are complete with respect to ``stream`` when the call returns (``stream`` waits
    on the workers' streams); the host arrays may be reused at once.
    '''
    ring = _ring(device)
    outs, jobs = [], []
    for a in arrays:
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].reshape(-1)).dtype, device=device)
        outs.append(t)
        flat_src = a.reshape(-1).view(np.uint8)
        flat_dst = t.view(-1).view(torch.uint8)
        for off in range(0, a.nbytes, CHUNK_BYTES):
            jobs.append((flat_src, flat_dst, off, min(CHUNK_BYTES, a.nbytes - off)))
    done_events = []
    ev_lock = threading.Lock()
    # The destination blocks come from the caching allocator on the CURRENT stream and may
    # still be in use by work queued there (a previous call's kernel reading its staged
    # drivers): the workers' DMA must not start before that work is done.
    allocated = torch.cuda.Event()
    allocated.record(torch.cuda.current_stream(device))

    def work(job):
        src, dst, off, n = job
        buf, ev = ring.free.get()
        try:
            if ev is not None:
                ev.synchronize()           # the DMA that last read this buffer is over
            np.copyto(buf.numpy()[:n], src[off:off + n])
            s = ring.stream()
            with torch.cuda.stream(s):
                s.wait_event(allocated)
                dst[off:off + n].copy_(buf[:n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            with ev_lock:
                done_events.append(ev)
        finally:
            ring.free.put((buf, ev))

    list(ring.pool.map(work, jobs))
    for ev in done_events:
        stream.wait_event(ev)
    return outs


def download(tensor, stream):
    '''
    Device tensor -> new numpy array, chunked through the pinned ring (D2H DMA of one chunk
    overlapping the host copy of another).  Waits for ``stream`` first.
    '''
    ring = _ring(tensor.device)
    src = tensor.contiguous()
    out = np.empty(tuple(src.shape), dtype=np.dtype(str(src.dtype).replace('torch.', '')))
    flat_src = src.view(-1).view(torch.uint8)
    flat_dst = out.reshape(-1).view(np.uint8)
    ready = torch.cuda.Event()
    ready.record(stream)
    nbytes = flat_src.numel()
    jobs = [(off, min(CHUNK_BYTES, nbytes - off)) for off in range(0, nbytes, CHUNK_BYTES)]

    def work(job):
        off, n = job
        buf, ev = ring.free.get()
        try:
            if ev is not None:
                ev.synchronize()
            s = ring.stream()
            with torch.cuda.stream(s):
                s.wait_event(ready)
                buf[:n].copy_(flat_src[off:off + n], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(s)
            ev.synchronize()
            np.copyto(flat_dst[off:off + n], buf.numpy()[:n])
        finally:
            ring.free.put((buf, None))

    list(ring.pool.map(work, jobs))
    return out
