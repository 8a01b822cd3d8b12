"""Batch sharding for the view-synthesis loss: one process per GPU, contiguous batch slices.

Every sample's warp / loss is independent: the only cross-sample coupling is the final mean.
A rank therefore runs the unchanged fused kernel on its slice with loss_scale = B_local / B_global, which makes
its gradients the rank's exact share of the global-batch gradient (a SUM all-reduce of the network gradients
then yields the global gradient), and its three loss scalars are combined with the same weights.  No
data-path collective exists; the only exchange is the 12-byte loss all-reduce below (NCCL on GPUs, gloo in the
CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous [lo, hi) of `batch` samples for `rank`; the first batch % world ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError('rank %d outside world of %d' % (rank, world))
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_snippets(d, rank, world):
    """Slice every batch-first tensor (or list of tensors) of a snippet dict to this rank's samples."""
    B = d['tgt'].shape[0]
    lo, hi = shard_range(B, rank, world)

    def cut(v):
        if isinstance(v, (list, tuple)):
            return [cut(t) for t in v]
        return v[lo:hi].contiguous()
    return {k: cut(v) for k, v in d.items()}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device, sysfs='/sys/bus/pci/devices'):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs `local_cpulist` of the GPU's PCI
    function), so that the pinned host arenas it allocates next are first-touched on that node and a rank's H2D / D2H
    copies do not cross the socket interconnect.  One rank per GPU; call it before ops.HostPipeline / any pin_memory().
    -> the set of CPUs bound to, or None when the topology is not exposed (containers, single-node hosts): then
    nothing changes."""
    import os
    try:
        props = torch.cuda.get_device_properties(device)
        bdf = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open(os.path.join(sysfs, bdf, 'local_cpulist')) as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:   # noqa: BLE001 -- no CUDA device, no sysfs, no affinity call on this platform: leave things alone
        return None


def local_loss_scale(batch_local, batch_global):
    """loss_scale for ViewSynthesisPlan so that sum-reduced rank gradients equal the global-batch gradient."""
    return float(batch_local) / float(batch_global)


def reduce_losses(local_losses, batch_local, batch_global, group=None):
    """Global-batch (pixel, smooth, exp) from per-rank means: sum over ranks of (B_local / B_global) * mean."""
    t = local_losses.detach().clone() * (float(batch_local) / float(batch_global))
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def reduce_sum_(tensors, group=None):
    """In-place SUM all-reduce of a list of gradient tensors as one flat bucket."""
    if not (dist.is_available() and dist.is_initialized()) or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        t.copy_(flat[off:off + t.numel()].reshape(t.shape))
        off += t.numel()
    return tensors


def _flat_layout(shapes):
    """Layout of a list of tensors inside one flat float32 arena: every tensor starts on a 16-byte boundary (the
    kernels that consume the views use vector accesses).  -> (shapes, sizes, offsets, numel)."""
    shapes = [tuple(int(d) for d in s) for s in shapes]
    sizes, offsets, off = [], [], 0
    for s in shapes:
        n = 1
        for d in s:
            n *= d
        sizes.append(n)
        offsets.append(off)
        off += (n + 3) // 4 * 4
    return shapes, sizes, offsets, max(off, 4)


class DataParallelAdam(object):
    """The data-parallel optimiser step of SURVEY.md 8e / 8f.3: every parameter, gradient and Adam moment lives in
    one flat float32 arena; step() all-reduces the gradient arena bucket by bucket (asynchronously: NCCL's stream)
    and applies tf.train.AdamOptimizer (train_depth_then_cam_lr.py:413) to a bucket as soon as its sum has landed,
    so the update of bucket k runs under the all-reduce of bucket k+1.  No concatenation, no copy-back.

    `params[i]` / `grads[i]` are views shaped like shapes[i]: a network writes its gradients straight into
    `grads[i]` (e.g. `p.grad = dp.grads[i]`).  Rank gradients are expected to be this rank's SHARE of the
    global-batch gradient (local_loss_scale), hence a SUM and grad_scale = 1.

    adam_fn(param, grad, m, v, step) applies one update in place to flat views; the default is the CUDA kernel
    (ops.adam_step).  The CPU gloo tests inject a restatement to exercise the bucketing logic.
    """

    def __init__(self, shapes, device, lr, beta1=0.9, beta2=0.999, eps=1e-8, bucket_bytes=32 << 20, group=None,
                 adam_fn=None, grad_scale=1.0):
        self.shapes, sizes, self.offsets, self.numel = _flat_layout(shapes)
        mk = lambda: torch.zeros(self.numel, dtype=torch.float32, device=device)
        self.param_flat, self.grad_flat, self.m_flat, self.v_flat = mk(), mk(), mk(), mk()
        cut = lambda flat: [flat[o:o + n].view(s) for o, n, s in zip(self.offsets, sizes, self.shapes)]
        self.params, self.grads = cut(self.param_flat), cut(self.grad_flat)
        per = max(int(bucket_bytes) // 4 // 4 * 4, 4)
        self.buckets = [(lo, min(lo + per, self.numel)) for lo in range(0, self.numel, per)]
        self.group, self.t = group, 0
        self.hyper = dict(lr=lr, beta1=beta1, beta2=beta2, eps=eps, grad_scale=grad_scale)
        if adam_fn is None:
            from . import ops

            def adam_fn(p, g, m, v, step):
                ops.adam_step(p, g, m, v, step, **self.hyper)
        self.adam_fn = adam_fn

    def step(self):
        """All-reduce (SUM) + Adam over every bucket; returns the step count t."""
        self.t += 1
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        works = []
        if multi:
            for lo, hi in self.buckets:
                works.append(dist.all_reduce(self.grad_flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group,
                                             async_op=True))
        for k, (lo, hi) in enumerate(self.buckets):
            if multi:
                works[k].wait()     # NCCL: the current stream waits for bucket k, the host does not
            self.adam_fn(self.param_flat[lo:hi], self.grad_flat[lo:hi], self.m_flat[lo:hi], self.v_flat[lo:hi], self.t)
        return self.t


class _DevMem(object):
    """A raw device range exposed to torch through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (ptr, False), 'version': 2}


class PeerArena(object):
    """One cudaMalloc'ed, zero-filled block per rank, mapped into every rank of the node over CUDA IPC with this
    rank's GPU as the accessor (csrc/vsl_optim.cu vsl_peer_alloc / vsl_ipc_*).  base[r] is rank r's block as a raw
    device address valid in THIS process (base[rank] is the local block).  One node only."""

    def __init__(self, nbytes, device, group=None):
        import ctypes
        from . import _lib
        self._lib, self.lib = _lib, _lib.load()
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise TypeError('PeerArena needs a CUDA device (this path has no CPU fallback)')
        multi = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if multi else 1
        self.rank = dist.get_rank(group) if multi else 0
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self._opened = []
        with torch.cuda.device(self.device):
            p = ctypes.c_void_p()
            _lib.check(self.lib.vsl_peer_alloc(self.nbytes, ctypes.byref(p)))
            self.local = p.value
            self.base = [self.local]
            if self.world > 1:
                h = ctypes.create_string_buffer(64)
                _lib.check(self.lib.vsl_ipc_get_handle(self.local, h))
                handles = [None] * self.world
                dist.all_gather_object(handles, h.raw, group=group)
                self.base = []
                for r in range(self.world):
                    if r == self.rank:
                        self.base.append(self.local)
                        continue
                    q = ctypes.c_void_p()
                    _lib.check(self.lib.vsl_ipc_open(handles[r], ctypes.byref(q)))
                    self._opened.append(q.value)
                    self.base.append(q.value)
                dist.barrier(group=group)       # every rank has opened every handle before anyone moves on
        self._group = group

    def view(self, byte_offset, n, dtype=torch.float32):
        """torch view of n elements of the LOCAL block."""
        typestr = {torch.float32: '<f4', torch.int32: '<i4'}[dtype]
        return torch.as_tensor(_DevMem(self.local + byte_offset, n, typestr), device=self.device)

    def close(self):
        """Unmap the peers' blocks, wait for every rank to have done so, free the local block."""
        if self.local is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            for q in self._opened:
                self.lib.vsl_ipc_close(q)
            self._opened = []
            if self.world > 1 and dist.is_initialized():
                dist.barrier(group=self._group)
            self.lib.vsl_peer_free(self.local)
        self.local = None


class MulticastArena(object):
    """PeerArena's interface over a symmetric allocation that is ALSO bound to an NVSwitch multicast object (NVLS):
    base[r] = rank r's block as mapped here, mc = the multicast address of the same block (a multimem.ld_reduce
    through it sums every rank's copy inside the switch, a multimem.st lands in every rank's copy).  The
    allocation, the fabric-handle exchange and the multicast binding are torch.distributed._symmetric_memory's
    (CUDA VMM + cuMulticast*: plumbing); the kernels that use the addresses are csrc/vsl_optim.cu's.
    available(device) says whether the driver / switch support it; the constructor raises if they do not."""

    @staticmethod
    def available(device):
        try:
            from torch._C._autograd import DeviceType
            from torch._C._distributed_c10d import _SymmetricMemory
            device = torch.device(device)
            idx = device.index if device.index is not None else torch.cuda.current_device()
            return bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 and
                        _SymmetricMemory.has_multicast_support(DeviceType.CUDA, idx))
        except Exception:
            return False

    def __init__(self, nbytes, device, group=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        self.lib = _lib.load()
        self.device = torch.device(device)
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        with torch.cuda.device(self.device):
            self._t = symm.empty(self.nbytes // 4, dtype=torch.float32, device=self.device)
            self._t.zero_()
            torch.cuda.synchronize()
            self._hdl = symm.rendezvous(self._t, group)
        self.base = [int(p) for p in self._hdl.buffer_ptrs]
        self.local = self.base[self.rank]
        self.mc = int(self._hdl.multicast_ptr or 0)
        if self.mc == 0:
            raise RuntimeError('the symmetric allocation has no multicast address (NVLS unavailable on this node)')
        if self.local != self._t.data_ptr():
            raise RuntimeError('symmetric memory: the local buffer pointer is not the tensor the arena was built from')
        self._group = group
        dist.barrier(group=group)

    view = None   # set below (PeerArena.view)

    def close(self):
        if self.local is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self._group)
        self.local, self._hdl, self._t = None, None, None


MulticastArena.view = PeerArena.view


class InProcessArena(object):
    """PeerArena's interface for `world` ranks that live in ONE process (one block per rank, on the given devices --
    which may all be the same GPU): the blocks are plain cudaMalloc allocations, so base[r] is valid for every rank
    and no IPC is involved.  Lets one host thread drive several ranks on separate streams -- a single-process
    multi-GPU trainer, and the way the peer kernels are exercised on a one-GPU box."""

    def __init__(self, rank, world, bases, nbytes, device):
        from . import _lib
        self.lib = _lib.load()
        self.rank, self.world, self.base, self.nbytes = rank, world, list(bases), nbytes
        self.local, self.device = bases[rank], torch.device(device)

    @classmethod
    def make(cls, nbytes, devices):
        import ctypes
        from . import _lib
        lib = _lib.load()
        nbytes = (int(nbytes) + 255) // 256 * 256
        devices = [torch.device(d) for d in devices]
        bases = []
        for d in devices:
            if d.type != 'cuda':
                raise TypeError('InProcessArena needs CUDA devices (this path has no CPU fallback)')
            with torch.cuda.device(d):
                p = ctypes.c_void_p()
                _lib.check(lib.vsl_peer_alloc(nbytes, ctypes.byref(p)))
                bases.append(p.value)
        return [cls(r, len(devices), bases, nbytes, devices[r]) for r in range(len(devices))]

    view = PeerArena.view

    def close(self):
        if self.local is None:
            return
        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            self.lib.vsl_peer_free(self.local)
        self.local = None


class PeerTimeout(RuntimeError):
    """A peer did not reach a barrier of the fused optimiser step in time; the step was NOT applied."""


class PeerDataParallelAdam(object):
    """dist.DataParallelAdam's step as ONE kernel over NVLink peer memory (csrc/vsl_optim.cu dp_adam_kernel):
    rank r sums every rank's gradient over its shard of the flat arena with peer-to-peer loads (reduce-scatter),
    applies tf.train.AdamOptimizer to the shard (the moments exist for the shard only) and stores the new
    parameters into every rank's arena (all-gather).  Two flag barriers in peer memory bracket it; nothing goes
    through NCCL after construction.  Replicas stay bit-identical (one owner per element, fixed sum order).

    The barrier epoch and Adam's step count live in device memory (`state`) and are advanced by the kernels, so
    step() passes no per-step argument: the whole step can be captured into a CUDA graph together with the networks.

    Failure behaviour: a barrier waits `timeout_s` (default 120 s) of wall-clock time for its peers.  If one does not
    arrive, the kernel raises a flag in PINNED HOST memory, every later launch of the step becomes a no-op (nothing is
    ever computed from a late peer's half-written gradients) and the next step() / check_peers() raises PeerTimeout.
    step() polls the flag without synchronising; it therefore reports a timeout one call late at the latest.

    Same `params` / `grads` views as DataParallelAdam.  One node, world <= 16, CUDA only.  Call close() before
    the process group is destroyed.
    """

    def __init__(self, shapes, device, lr, beta1=0.9, beta2=0.999, eps=1e-8, group=None, grad_scale=1.0,
                 timeout_s=120.0, arena=None, multicast=False):
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        device = torch.device(device)
        self.shapes, sizes, self.offsets, self.numel = _flat_layout(shapes)
        # local block: [flags 256 B][state: epoch, t (device ints), padded to 256 B][parameters][gradients]
        arena_bytes = 512 + 8 * self.numel
        # multicast=True: the arenas are bound to an NVSwitch multicast object and the step sums / broadcasts through
        # it (dp_adam_mc_kernel); 'auto': that, where the node supports it, else the peer-to-peer form
        # (per rank and direction the links carry 2 (N - 1) / N arenas peer to peer, 1 + 1 / N through the switch:
        # 'auto' takes the multicast form from 4 ranks up)
        if arena is None and multicast and (multicast != 'auto' or (
                MulticastArena.available(device) and dist.get_world_size(group) >= 4)):
            arena = MulticastArena(arena_bytes, device, group)
        self.arena = arena if arena is not None else PeerArena(arena_bytes, device, group)
        self.mc = int(getattr(self.arena, 'mc', 0) or 0)
        if self.arena.nbytes < arena_bytes:
            raise ValueError('arena too small: %d < %d bytes' % (self.arena.nbytes, arena_bytes))
        self.world, self.rank = self.arena.world, self.arena.rank
        if self.world > 16:
            raise ValueError('PeerDataParallelAdam supports up to 16 ranks of one node')
        per = (self.numel // 4 + self.world - 1) // self.world * 4          # shard length, a multiple of 4 floats
        self.lo = min(self.rank * per, self.numel)
        self.hi = min(self.lo + per, self.numel)
        self.param_flat = self.arena.view(512, self.numel)
        self.grad_flat = self.arena.view(512 + 4 * self.numel, self.numel)
        self.state = self.arena.view(256, 4, torch.int32)                  # [epoch, t, -, -], advanced on the device
        # the failure flag lives in pinned host memory: the barrier kernel writes it over PCIe, the host polls it
        # without a synchronisation
        self.timed_out = torch.zeros(4, dtype=torch.int32).pin_memory()
        self.timeout_ms = max(int(float(timeout_s) * 1000.0), 1)
        with torch.cuda.device(device):
            mk = lambda n: torch.zeros(max(n, 4), dtype=torch.float32, device=device)
            self.m_shard, self.v_shard = mk(self.hi - self.lo), mk(self.hi - self.lo)
        cut = lambda flat: [flat[o:o + n].view(s) for o, n, s in zip(self.offsets, sizes, self.shapes)]
        self.params, self.grads = cut(self.param_flat), cut(self.grad_flat)
        self._pf = _lib.ptr_array(list(self.arena.base))
        self._pp = _lib.ptr_array([b + 512 for b in self.arena.base])
        self._pg = _lib.ptr_array([b + 512 + 4 * self.numel for b in self.arena.base])
        self.hyper = (float(lr), float(beta1), float(beta2), float(eps))
        self.grad_scale = float(grad_scale)
        self.t = 0
        self.device = device

    @classmethod
    def in_process(cls, shapes, devices, lr, **kw):
        """`len(devices)` ranks inside this process (see InProcessArena) -> list of instances, rank order."""
        _, _, _, numel = _flat_layout(shapes)
        arenas = InProcessArena.make(512 + 8 * numel, devices)
        return [cls(shapes, a.device, lr, arena=a, **kw) for a in arenas]

    def _raise_if_timed_out(self):
        if int(self.timed_out[0]) != 0:
            raise PeerTimeout('rank %d of %d: a peer did not reach the optimiser-step barrier within %.0f s; the step '
                              'was not applied and the replicas must be considered out of sync'
                              % (self.rank, self.world, self.timeout_ms / 1000.0))

    def step(self, stream=None):
        """barrier (all gradients written) -> fused reduce-scatter + Adam + all-gather -> barrier (all parameters
        landed), one C call.  Enqueued on the current stream; the host does not wait.  Raises PeerTimeout if an
        earlier step's barrier gave up."""
        self._raise_if_timed_out()
        self.t += 1
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        if self.mc and self.world > 1:
            self._lib.check(self.lib.vsl_dp_step_mc(self._pf, self.mc + 512 + 4 * self.numel, self.mc + 512,
                                                    self.arena.local + 512, self.rank, self.world,
                                                    self.m_shard.data_ptr(), self.v_shard.data_ptr(), self.lo, self.hi,
                                                    *self.hyper, self.grad_scale, self.state.data_ptr(),
                                                    self.timed_out.data_ptr(), self.timeout_ms, st))
            return self.t
        self._lib.check(self.lib.vsl_dp_step(self._pf, self._pg, self._pp, self.rank, self.world,
                                             self.m_shard.data_ptr(), self.v_shard.data_ptr(), self.lo, self.hi,
                                             *self.hyper, self.grad_scale, self.state.data_ptr(),
                                             self.timed_out.data_ptr(), self.timeout_ms, st))
        return self.t

    def check_peers(self):
        """Synchronises this rank's device, then raises PeerTimeout if any barrier gave up.  Call it wherever the
        parameters are about to be trusted (before a checkpoint, at the end of training)."""
        torch.cuda.synchronize(self.device)
        self._raise_if_timed_out()

    def close(self):
        self.arena.close()
