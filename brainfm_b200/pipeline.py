"""Host <-> device pipelining around BaseGen.generate_batch.

The reference generator reads every volume from disk, decodes it on the host and uploads the crop for every
sample (Generator/utils.py:296-305, datasets.py:312,364), synchronously on the training stream.  Here the
volumes of batch k+1 are uploaded and the results of batch k-1 are downloaded on their own CUDA streams while
batch k is being generated, so that in steady state a step costs max(upload, generate, download) instead of
their sum (PCIe is full duplex; the copy engines run beside the SMs).

    pipe = HostPipeline(ds)
    t = pipe.submit(indices, uploads=[(path, kind, pinned_host_tensor), ...])   # asynchronous
    #   or batched: uploads=[([path0, path1, ...], kind, pinned (n, *shape) tensor)] -- one DMA per kind
    ...
    items, host_out = t.wait()      # host_out: pinned (B, 1, *size) float32 'input' volumes

Hazards are ordered with events: an upload into a cached device volume waits for the last batch that read
that volume; a download waits for the batch that produced it; a pinned output slot is reused only after its
previous download has been waited for by the caller.
"""
import torch


class _Ticket:
    def __init__(self, items, host_out, event):
        self.items, self.host_out, self._event = items, host_out, event

    def wait(self):
        self._event.synchronize()
        return self.items, self.host_out


class _DeviceTicket:
    def __init__(self, items, event, stream):
        self.items, self._event, self._stream = items, event, stream

    def wait(self, stream=None):
        """Orders the consumer's stream (default: the current one) after the batch; no host synchronisation.

        The outputs live in the lane stream's memory pool.  They are NOT record_stream()-ed (deferred frees make the
        caching allocator fall back to cudaMalloc for every batch: measured 1.38 instead of 0.72 ms per step);
        instead every submit() makes the lane wait for the submitting stream, so a block freed by the consumer is only
        rewritten after the work the consumer had enqueued on that stream.  Consume a batch on the stream you submit
        from, or keep the tensors referenced until your own stream is done with them."""
        cur = stream or torch.cuda.current_stream()
        cur.wait_event(self._event)
        return self.items

    def synchronize(self):
        self._event.synchronize()
        return self.items


class DevicePipeline:
    """Device-resident generation with `depth` batches in flight, each on its own CUDA stream with its own scratch.

    The stages of the fused chain are bound by different things (the warp gather by the L1 data pipe, GMM noise and
    the zoom back by instruction issue, the normalise pass by HBM), and every stage ends in a tail of partially
    filled SMs; with two batches in flight the block scheduler fills one batch's tails and stalls with the other
    batch's blocks.  Measured on B200, batch 8 x 160^3: 0.80 -> 0.71 ms per batch (profiles/r2_streams.json).

        pipe = DevicePipeline(ds, depth=2)
        t = pipe.submit(indices)          # asynchronous; planning happens on the calling thread
        ...
        items = t.wait()                  # the current stream now waits for that batch (no host sync)
    """

    def __init__(self, ds, depth=2):
        self.ds = ds
        self.device = ds.device
        self.lanes = [dict(stream=torch.cuda.Stream(device=self.device), ws={}) for _ in range(max(1, int(depth)))]
        self._k = 0
        # the plan arena is a ring whose slot k is reused only after batch k - len(ring) has finished on the GPU: with
        # the default 3 slots the host could not plan batch k before batch k - 3 is done, which leaves a lane idle
        # while batch k is being planned (measured: 0.93 instead of 0.72 ms per step).  2 * depth + 2 slots keep every
        # lane one batch ahead.
        want = 2 * len(self.lanes) + 2
        if len(ds.arena.slots) < want:
            from .plan import Arena
            ds.arena = Arena(ds.device, capacity=ds.arena.capacity, slots=want)

    def submit(self, indices=None, timers=None, call=None):
        """Generate `indices` (ds.generate_batch) -- or run `call()` (anything that enqueues generator work on the
        current stream with the dataset's scratch, e.g. `lambda: generate_slab(ds, idx)`) -- on the next lane."""
        ds = self.ds
        lane = self.lanes[self._k % len(self.lanes)]
        self._k += 1
        cur = torch.cuda.current_stream(self.device)
        lane["stream"].wait_stream(cur)               # volumes uploaded / refreshed on the caller's stream are visible
        saved = ds._ws
        ds._ws = lane["ws"]
        ds._alloc_stream = cur                         # output tensors: the caller's pool (see NativePlanner.run_fast)
        try:
            with torch.cuda.stream(lane["stream"]):
                if call is not None:
                    items = call()
                else:
                    items = None
                    if timers is None:
                        items = ds.generate_batch_fast(indices)          # one library call when the batch allows it
                    if items is None:
                        items = ds.generate_batch(list(indices), timers=timers)
                ev = torch.cuda.Event()
                ev.record(lane["stream"])
        finally:
            ds._ws = saved
            ds._alloc_stream = None
        return _DeviceTicket(items, ev, lane["stream"])


class HostPipeline:
    def __init__(self, ds, depth=3, key='input'):
        self.ds = ds
        self.device = ds.device
        self.key = key
        self.copy_in = torch.cuda.Stream(device=self.device)
        self.copy_out = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self._slots = []                 # pinned host output buffers (ring)
        self._slot_events = []           # download event of the last use of each slot
        self._next = 0
        self._last_reader = {}           # (path, kind) -> event of the last batch that read the volume
        self._stages = {}                # batched-upload staging buffers on the device
        self._last_done = None           # event of the most recent batch

    def _stage(self, kind, host):
        """Device staging buffer of a batched upload: a ring of two per (kind, shape), so that the DMA of step k+1 can
        run while the device-to-device copies of step k (compute stream) still read the other one."""
        key = (kind, tuple(host.shape), host.dtype)
        ring = self._stages.get(key)
        if ring is None:
            ring = {"next": 0, "slots": [{"buf": torch.empty(host.shape, dtype=host.dtype, device=self.device),
                                          "free": None} for _ in range(2)]}
            self._stages[key] = ring
        slot = ring["slots"][ring["next"]]
        ring["next"] ^= 1
        return slot["buf"], slot

    def _slot(self, shape):
        if not self._slots:
            for _ in range(self.depth):
                self._slots.append(torch.empty(shape, dtype=torch.float32).pin_memory())
                self._slot_events.append(None)
        k = self._next
        self._next = (self._next + 1) % self.depth
        if tuple(self._slots[k].shape) != tuple(shape):
            self._slots[k] = torch.empty(shape, dtype=torch.float32).pin_memory()
        ev = self._slot_events[k]
        if ev is not None:
            ev.synchronize()             # the caller is done with this slot `depth` submissions later
        return k

    def submit(self, indices, uploads=()):
        ds = self.ds
        main = torch.cuda.current_stream(self.device)
        # 1. uploads on the copy-in stream (after the last reader of each destination volume)
        flat = []                                         # (path, kind) of every uploaded volume
        direct = []                                       # ... of those copied straight into the cached tensor
        staged = []                                       # batched uploads: (paths, kind, staging buffer, slot)
        if uploads:
            with torch.cuda.stream(self.copy_in):
                for path, kind, host in uploads:
                    if isinstance(path, (list, tuple)):
                        # batched host buffer (len(paths), *volume shape), in the volumes' STORED dtype (uint8 labels,
                        # int16 / uint8 / float32 images -- widened to the cached float32 on the device by
                        # bfm_ingest_volume, so an int16 T1 crosses PCIe as 2 bytes per voxel): ONE host->device DMA into a staging buffer
                        # on the copy stream (nothing else is queued there, so the copy engine never idles between
                        # steps); the device-to-device copies into the cached volumes follow on the compute stream.
                        # With one DMA per volume the engine idles 40-50 us between copies whenever the download
                        # direction is busy too.
                        stage, slot = self._stage(kind, host)
                        if slot["free"] is not None:
                            self.copy_in.wait_event(slot["free"])     # device copies of two steps ago are done
                        stage.copy_(host, non_blocking=True)
                        staged.append((list(path), kind, stage, slot))
                        flat += [(p, kind) for p in path]
                    else:
                        # after the last batch that read this volume; a volume that was never uploaded before may have
                        # been read by ANY earlier batch (it came from cache.get): wait for the latest one
                        ev = self._last_reader.get((path, kind), self._last_done)
                        if ev is not None:
                            self.copy_in.wait_event(ev)
                        ds.cache.upload(path, kind, host)     # copies only: the copy stream never waits for an SM
                        flat.append((path, kind))
                        direct.append((path, kind))
                ev_in = torch.cuda.Event()
                ev_in.record(self.copy_in)
            main.wait_event(ev_in)
            for paths, kind, stage, slot in staged:       # compute stream: ordered after the previous batch's reads
                for p, part in zip(paths, stage.unbind(0)):
                    ds.cache.upload(p, kind, part)
                slot["free"] = torch.cuda.Event()
                slot["free"].record(main)
            for p, kind in direct:                        # (staged volumes were sanitised by bfm_ingest_volume)
                ds.cache.sanitize(p, kind)                # nan_to_num on the compute stream
        # 2. generation on the caller's stream
        items = ds.generate_batch(list(indices))
        ev_done = torch.cuda.Event()
        ev_done.record(main)
        for key in flat:
            self._last_reader[key] = ev_done
        self._last_done = ev_done
        # 3. download on the copy-out stream
        outs = [it[4][self.key] if not isinstance(it[4], list) else torch.cat([s[self.key] for s in it[4]], 0)
                for it in items]
        shape = (sum(o.shape[0] for o in outs), *outs[0].shape[1:])
        k = self._slot(shape)
        host = self._slots[k]
        whole = getattr(ds, '_last_out', None) if self.key == 'input' else None
        if whole is not None:
            n = 1
            for v in shape:
                n *= int(v)
            if whole.numel() != n or not whole.is_contiguous() or whole.data_ptr() != outs[0].data_ptr():
                whole = None
            else:
                whole = whole.view(shape)
        with torch.cuda.stream(self.copy_out):
            self.copy_out.wait_event(ev_done)
            if whole is not None:                         # the batch's volumes are one tensor: one DMA
                host.copy_(whole, non_blocking=True)
                whole.record_stream(self.copy_out)
            else:
                row = 0
                for o in outs:
                    host[row:row + o.shape[0]].copy_(o, non_blocking=True)
                    o.record_stream(self.copy_out)
                    row += o.shape[0]
            ev_out = torch.cuda.Event()
            ev_out.record(self.copy_out)
        self._slot_events[k] = ev_out
        return _Ticket(items, host, ev_out)
