'''
Host-resident forward runs: a (T, N) block of drivers that lives in (pinned) host memory is
streamed through the GPU in slabs of time steps, with the host->device copies, the fused
kernel and the device->host copy of the result overlapped on three CUDA streams and two
device buffer sets.  This is the end-to-end path a caller with numpy / HDF5 data uses
(a year of a MODIS tile in float64, or of the global grid on 2 GPUs, does not
fit in HBM at once).

The transpiration gate of the static path is evaluated per slab (each slab is one
``_et`` call); with real daytime data it is always open.
'''
import numpy as np
import torch

from .forward import run_forward
from .rawio import RawDriver
from .synthetic import DRIVER_NAMES, PER_PIXEL

_PER_STEP = [k for k in DRIVER_NAMES if k not in PER_PIXEL and k != 'sw_rad_night']


def bind_to_gpu_numa_node(device_index):
    '''
    Restricts this process to the CPU cores of the NUMA node the GPU hangs off (sysfs), so
    that pinned staging buffers allocated afterwards are node-local (first touch) and
    host<->device copies of several ranks do not cross the socket interconnect.  Returns the
    node, or None when the topology cannot be read (single-socket hosts: nothing to do).
    '''
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), 'pci_domain_id', 0)
        dev_id = getattr(torch.cuda.get_device_properties(device_index), 'pci_device_id', 0)
        path = '/sys/bus/pci/devices/%04x:%02x:%02x.0' % (dom, bus, dev_id)
        node = int(open(path + '/numa_node').read())
        if node < 0:
            return None
        cpus = open('/sys/devices/system/node/node%d/cpulist' % node).read().strip()
        cores = set()
        for part in cpus.split(','):
            a, _, b = part.partition('-')
            cores.update(range(int(a), int(b or a) + 1))
        allowed = cores & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


class StreamedForward(object):
    '''
    Parameters
    ----------
    N : pixels per time step
    slab_T : time steps per slab (device memory used: 2 x 12 x slab_T x N x itemsize)
    table, pft : dense 11 x 13 BPLUT table and uint8 land-cover code per pixel
    formats : optional dict name -> (numpy dtype, scale, offset, fill) of the per-step drivers
        that arrive in a storage format (``rawio.RawDriver`` entries of ``host`` in ``run``):
        their slabs cross the bus and sit in HBM as stored -- uint8 / int16 -- and the fused
        kernel converts them while it reads
    prep : loader conversions applied on the device, see ``run_forward``
    '''

    def __init__(self, N, slab_T, table, pft, device=None, dtype=np.float32, gate='global',
                 formats=None, prep=None):
        self.N, self.slab_T = int(N), int(slab_T)
        self.device = torch.device(device) if device is not None else \
            torch.device('cuda', torch.cuda.current_device())
        self.np_dtype = np.dtype(dtype)
        self.t_dtype = torch.float32 if self.np_dtype == np.float32 else torch.float64
        self.table = np.asarray(table, np.float64)
        self.gate = gate
        with torch.cuda.device(self.device):
            self.pft = torch.as_tensor(np.asarray(pft, np.uint8)).to(self.device)
            self.formats = dict(formats or {})
            self.prep = prep
            kinds = dict((k, getattr(torch, np.dtype(self.formats[k][0]).name)
                          if k in self.formats else self.t_dtype) for k in _PER_STEP)
            kinds['out'] = self.t_dtype
            self.bufs = [dict((k, torch.empty((self.slab_T, self.N), device=self.device,
                                              dtype=kinds[k])) for k in _PER_STEP + ['out'])
                         for _ in range(2)]
            self.flags = [torch.zeros(1, device=self.device, dtype=torch.int32) for _ in range(2)]
            self.s_in = torch.cuda.Stream(self.device)
            self.s_run = torch.cuda.Stream(self.device)
            self.s_out = torch.cuda.Stream(self.device)
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def run(self, host, out_host):
        '''
        host : dict name -> pinned CPU tensor; per-step drivers (T, N), ``temp_annual`` and
               ``pressure`` (N,); ``sw_rad_night`` a python scalar
        out_host : pinned CPU tensor (T, N) receiving ET [W m-2]
        Returns after the last device->host copy has completed.
        '''
        T = out_host.shape[0]
        dev = self.device
        with torch.cuda.device(dev):
            per_pixel = dict((k, host[k].to(dev, non_blocking=True)) for k in PER_PIXEL)
            self.h2d_bytes += sum(v.numel() * v.element_size() for v in per_pixel.values())
            cur = torch.cuda.current_stream(dev)
            for s in (self.s_in, self.s_run, self.s_out):
                s.wait_stream(cur)
            ev_run = [None, None]
            ev_out = [None, None]
            n_slabs = (T + self.slab_T - 1) // self.slab_T
            for i in range(n_slabs):
                b = i % 2
                t0, t1 = i * self.slab_T, min(T, (i + 1) * self.slab_T)
                rows = t1 - t0
                buf = self.bufs[b]
                with torch.cuda.stream(self.s_in):
                    if ev_run[b] is not None:
                        self.s_in.wait_event(ev_run[b])   # kernel that read this set is done
                    for k in _PER_STEP:
                        src = host[k].raw if isinstance(host[k], RawDriver) else host[k]
                        buf[k][:rows].copy_(src[t0:t1], non_blocking=True)
                        self.h2d_bytes += rows * self.N * buf[k].element_size()
                    ev_in = torch.cuda.Event()
                    ev_in.record(self.s_in)
                with torch.cuda.stream(self.s_run):
                    self.s_run.wait_event(ev_in)
                    if ev_out[b] is not None:
                        self.s_run.wait_event(ev_out[b])  # previous result left the buffer
                    drv = [(RawDriver(buf[k][:rows], *self.formats[k][1:]) if k in self.formats
                            else buf[k][:rows]) if k in buf else
                           (per_pixel[k] if k in per_pixel else host.get(k, 0))
                           for k in DRIVER_NAMES]
                    run_forward(drv, pft=self.pft, table=self.table, semantics='static',
                                out_mode='sum', gate=self.gate, dtype=self.np_dtype,
                                device=dev, stream=self.s_run, gate_flag=self.flags[b],
                                out=[buf['out'][:rows]], prep=self.prep)
                    self.launches += (2 if self.gate == 'global' else 1) * \
                        (3 if self.np_dtype == np.float32 else 1)
                    ev_run[b] = torch.cuda.Event()
                    ev_run[b].record(self.s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_run[b])
                    out_host[t0:t1].copy_(buf['out'][:rows], non_blocking=True)
                    self.d2h_bytes += rows * self.N * buf['out'].element_size()
                    ev_out[b] = torch.cuda.Event()
                    ev_out[b].record(self.s_out)
            self.s_out.synchronize()
            cur.wait_stream(self.s_run)
        return out_host
