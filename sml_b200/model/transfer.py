"""SML period orchestrator -- drop-in for ``meta_train`` of the reference's model/transfer.py:302-1031.

Same constructor, methods and attributes as the reference (``MFbase``, ``transfer``,
``last_user_weight`` ... ``MF_optimizer``, ``transfer_optimizer``, ``recall``/``ndcg`` lists), same
period state machine and the same quirks (SURVEY.md section 7 hard part 4), but every device
operation of the two hot loops, ``updata()`` and the evaluation is a kernel of libsml_b200.so:

  reference (stock PyTorch)                              here
  -----------------------------------------------------  ------------------------------------------
  DataLoader + per-sample __getitem__ (:438-443)         whole epoch of (u,i,j) uploaded once
  6 gathers + run_MF + backward + dense Adam (:463-511)  ops.mf_step   (sml_mf_step)
  6 gathers + run_MF + backward + Adam(theta) (:701-728) ops.tr_step   (sml_tr_step)
  transfer(all rows) + copy_ (:884-902)                  ops.transfer_forward straight into MFbase
  test_model over 1024-row batches (:445,518,685,740)    one fused gather-dot-rank launch per set

There is no CPU / eager-PyTorch fallback: without the CUDA library every step raises.
"""
from __future__ import annotations

import copy
import io
import os
import pickle
import time

import numpy as np
import torch

from .. import ops
from ..profiling import EventTimers
from ..data.batching import ReferenceStream, mf_epoch_triples, tr_epoch_triples
from ..data.dataset import offlineDataset_withsample as SampleDaset
from ..data.dataset2 import trainDataset_withPreSample as PreSampleDatast
from ..data.dataset2 import transfer_data  # noqa: F401  (re-exported like the reference)
from ..evalution.evaluation2 import DeviceTestSet, test_model
from . import MF
from .conv_transfer import ConvTransfer, ConvTransfer_com


class FusedAdam(object):
    """The slice of torch.optim.Adam's interface the reference touches (``param_groups``, ``state``,
    ``zero_grad``, ``state_dict``), backed by the fused kernels: the update itself happens inside
    sml_mf_step / sml_tr_step, which read ``param_groups[0]['lr'|'weight_decay']`` every step."""

    def __init__(self, params, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        self.param_groups = [dict(params=list(params), lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                  amsgrad=False)]
        self.state = {}
        self.adam_state = None        # device int64[4]: step counter + packed step scalars

    @property
    def step_count(self):
        return 0 if self.adam_state is None else int(self.adam_state[0].item())

    def zero_grad(self, set_to_none=True):
        pass                          # gradients are re-zeroed by the fused update

    def step(self):
        raise RuntimeError("FusedAdam.step() is fused into ops.mf_step / ops.tr_step")

    def state_dict(self):
        return dict(step=self.step_count, param_groups=[{k: v for k, v in g.items() if k != "params"} for g in self.param_groups],
                    state={k: {n: t.detach().clone() for n, t in v.items()} for k, v in self.state.items()})


class _RemapUnpickler(pickle.Unpickler):
    """``torch.load(args.pre_model)`` in the reference unpickles a whole ``model.MF.MFbasemode``
    (model/transfer.py:322-325); map that module path onto this package."""

    def find_class(self, module, name):
        if module in ("model.MF", "MF"):
            module = "sml_b200.model.MF"
        return super().find_class(module, name)


class _RemapPickle(object):
    Unpickler = _RemapUnpickler
    __name__ = "sml_b200_remap_pickle"

    @staticmethod
    def load(f, **kw):
        return _RemapUnpickler(f, **kw).load()


def load_pre_model(path, user_num, item_num, laten_dim, device):
    """Pickled module (reference format), an MFbasemode of this package, or a plain state_dict."""
    obj = torch.load(path, map_location="cpu", weights_only=False, pickle_module=_RemapPickle)
    if isinstance(obj, dict):
        with torch.random.fork_rng(devices=[]):     # do not disturb the global generator (batch-order parity)
            m = MF.MFbasemode(num_user=user_num, num_item=item_num, laten_factor=laten_dim)
        m.load_state_dict(obj)
        obj = m
    return obj.to(device)


class meta_train(object):
    """SML model (reference: model/transfer.py:302-1031)."""

    def __init__(self, args, datasets, user_num, item_num, laten_dim, device=None, batch_source=None,
                 emulate_reference_rng=True, device_sampler=False):
        """``batch_source``: optional callable (kind, stage_id, epoch, n_rows) -> (user, item, neg) int64
        numpy arrays in batch order; when given it supplies the triples of every epoch (north_star:
        "the same supplied negative-sample indices").  Otherwise triples are drawn like the reference
        does (``emulate_reference_rng``: same global-RNG consumption as ``--numworkers 0``).
        ``device_sampler=True`` (throughput runs): shuffling, pre-sampled column selection and the rejection
        sampler of the transfer step run on the GPU (Philox), nothing but the period files crosses PCIe."""
        self._require_device(device)
        user_num, item_num = int(user_num), int(item_num)
        if laten_dim != 64:
            raise ValueError("sml_b200 kernels are specialised for laten=64 (the reference default)")
        if getattr(args, "TR_with_MF_bias", False):
            # the reference concatenates the bias column into 65-wide rows (model/transfer.py:347-355,918-921) and only its unused
            # transfer2 / GRU / transfer3 modules take that width; ConvTransfer(_com)(64, 64) fails on the shape there too
            raise NotImplementedError("TR_with_MF_bias (65-wide transfer input) is not supported by conv / conv_com (nor by the "
                                      "reference with these transfer types); reference default is False")
        if getattr(args, "norm", False) and args.transfer_type == "conv":
            # ConvTransfer.run_MF divides the score by the (attached) norm of the already normalised user row
            # (model/conv_transfer.py:80-83); conv_com's BCE branch ignores ``norm`` altogether (:122-126), which is what runs below
            raise NotImplementedError("norm=True with transfer_type='conv' is not implemented (off in the reference's final version)")
        self.batch_source = batch_source
        self.emulate_reference_rng = emulate_reference_rng
        self.device_sampler = device_sampler
        self._philox_calls = 0
        if args.data_name != "yelp":
            # the reference first builds a throw-away MFbasemode on this branch (model/transfer.py:314-317), which
            # consumes the global torch generator; mirrored so that batch order stays bit-identical
            MF.MFbasemode(num_user=user_num, num_item=item_num, laten_factor=laten_dim)
        self.MFbase = load_pre_model(args.pre_model, user_num, item_num, laten_dim, self.device)

        self.transfer_type = args.transfer_type
        self.with_MF_bias = False
        self.test_in_TR_train = args.test_in_TR_Train
        self.TR_train_sampleTYpe = args.TR_sample_type
        print("with MF bias:", self.with_MF_bias)
        print("transfer type:", self.transfer_type)
        self.need_writer = args.need_writer
        self.MF_TrainDataset = None
        if args.MF_sample == "alone":
            self.MF_TrainDataset = SampleDaset
        elif args.MF_sample == "all":
            self.MF_TrainDataset = PreSampleDatast
        if args.need_writer:
            from torch.utils.tensorboard import SummaryWriter
            path = ("m-num" + str(args.multi_num) + "-MF-lr" + str(args.MF_lr) + "-l2-" + str(args.l2) + "e-" + str(args.MF_epochs)
                    + "--TR-lr" + str(args.TR_lr) + "-l2-" + str(args.TR_l2) + "-e-" + str(args.TR_epochs)
                    + str(args.TR_sample_type) + "user-norm" + str(args.norm))
            self.writer = SummaryWriter(comment=path)

        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        input_dim = laten_dim
        self.last_user_weight = torch.zeros_like(uw)                  # model/transfer.py:358-364
        self.last_item_weight = torch.zeros_like(iw)
        self.user_weight_hat = copy.deepcopy(uw)
        self.item_weight_hat = copy.deepcopy(iw)
        self.last_user_weight_hat = copy.deepcopy(self.user_weight_hat)
        self.last_item_weight_hat = copy.deepcopy(self.item_weight_hat)

        if self.transfer_type == "conv":
            self.transfer = ConvTransfer(input_dim, input_dim).to(self.device)
            self.transfer_type = "transfer2"
        elif self.transfer_type == "conv_com":
            self.transfer = ConvTransfer_com(input_dim, input_dim).to(self.device)
            self.transfer_type = "transfer2"
        elif self.transfer_type in ("transfer", "transfer2", "GRU", "transfer3", "conv_com2"):
            raise NotImplementedError("transfer_type %r is one of the reference's unused alternatives "
                                      "(model/transfer.py:1-5); only conv_com / conv are built" % self.transfer_type)
        else:
            raise TypeError("No such type transfer!!!")

        self.dataset = datasets

        self.MF_optimizer = FusedAdam(self.MFbase.parameters(), lr=args.MF_lr, weight_decay=0)
        self.transfer_optimizer = FusedAdam(self.transfer.parameters(), lr=args.TR_lr, weight_decay=args.TR_l2)
        z = torch.zeros_like
        self._mf = dict(m_user=z(uw), v_user=z(uw), m_item=z(iw), v_item=z(iw), g_user=z(uw), g_item=z(iw))
        # MF Adam is dense in the reference (model/transfer.py:392); by default the bit-identical row-lazy form of it
        # runs (ops.adam_rows: only the batch rows move per step, every row once per epoch); SML_DENSE_ADAM=1 sweeps
        self.lazy_adam = os.environ.get("SML_DENSE_ADAM", "0") != "1"
        self.MF_optimizer.adam_state = ops.new_adam_state(self.device, history=self.lazy_adam)
        self._stamps = {}
        if self.lazy_adam:
            self._stamps = dict(stamp_user=ops.new_row_stamps(uw.shape[0], self.MF_optimizer.adam_state),
                                stamp_item=ops.new_row_stamps(iw.shape[0], self.MF_optimizer.adam_state))
        self.MF_optimizer.state = {self.MFbase.user_laten.weight: dict(exp_avg=self._mf["m_user"], exp_avg_sq=self._mf["v_user"]),
                                   self.MFbase.item_laten.weight: dict(exp_avg=self._mf["m_item"], exp_avg_sq=self._mf["v_item"])}
        th = self.transfer.theta
        self._tr = dict(m=z(th), v=z(th))
        self.transfer_optimizer.adam_state = ops.new_adam_state(self.device)
        self.transfer_optimizer.state = {"theta": dict(exp_avg=self._tr["m"], exp_avg_sq=self._tr["v"])}
        self._loss = torch.zeros(2, dtype=torch.float32, device=self.device)
        self._ws = {}
        self._dev_cache = {}
        self._free_bufs = []                       # [(int64 device buffer, event)] period-file buffers waiting for reuse
        self._pin_pool = []                        # [[pinned int64 staging buffer, event of its last copy]] (sampled triples -> device)
        self._buf_sizes = []                       # capacities of all pooled buffers ever allocated
        self.dev_cache_cap = 6                     # period files kept resident on the device
        self.prefetch = os.environ.get("SML_PREFETCH", "1") != "0"   # upload the next period's files during the current one (prefetch_files)
        self._copy_stream = None
        # one CUDA graph per (loop kind, epoch length, batch size, lr, l2): an epoch is ~10^3 short kernels, and the
        # host cannot enqueue them as fast as the GPU retires them (tools/epoch_bench.py: 104 us/step of launch time
        # against 119 us/step of kernel time for the transfer loop), so epochs after the first are graph replays
        self.use_graphs = True
        self._graphs = {}
        self._graph_warm = set()
        self.graph_launches = 0                    # kernels executed through graph replays (not seen by sml_launch_count)
        self._tab_version = 0                      # bumped by every method that writes the MF tables
        self._eval_cache = None
        self.eval_passes = dict(scored=0, reused=0)   # evaluations that launched the scoring kernel / reused a kept pass
        self.events = EventTimers(False)           # CUDA-event phase timers (bench.py switches them on)
        # Nothing in a period's control flow depends on a loss or a metric value, so they are not read back one by one (60 blocking
        # reads per period, each leaving the GPU idle while the host samples the next epoch): they stay on the device, the prints /
        # writer calls / list appends that consume them are queued in order and run at the end of train_one_stage3 after ONE
        # device->host copy (flush_deferred).  The host therefore runs ahead of the GPU inside a period and its sampling work hides
        # behind the epoch graphs.  SML_DEFER=0 reads every value where the reference does (prints appear immediately).
        self.defer = os.environ.get("SML_DEFER", "1") != "0"
        self._pending = {}                         # handle -> (device tensor, post-processing): values still on the device
        self._resolved = {}                        # handle -> host value (until the end of the period)
        self._handle = 0
        self._later_q = []                         # [callable(values)] consumers, in program order
        self._stage_depth = 0                      # > 0 inside train_one_stage3 (which flushes once, at its end)

        self.recall = []
        self.ndcg = []
        self.test_num = []
        self.recall_5 = []
        self.ndcg_5 = []
        self.MF_itr = 0
        self.TR_itr = 0
        self.recall_10 = []
        self.ndcg_10 = []
        self.timers = dict(mf=0.0, tr=0.0, updata=0.0, eval=0.0)     # host wall-clock, only for reporting

    # ------------------------------------------------------------------ helpers
    def _require_device(self, device):
        ops.lib()                                            # fail loudly if the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("sml_b200.meta_train needs a CUDA device (no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)

    def _workspace(self, B):
        if B not in self._ws:
            self._ws[B] = torch.zeros(int(ops.lib().sml_step_workspace_bytes(B)), dtype=torch.uint8, device=self.device)
        return self._ws[B]

    def _to_device(self, arr):
        """numpy int array -> cached int64 device tensor (period files are re-used across stages: D_{t+1} is the validation
        file of period t and the training file of period t+1, so every file crosses PCIe once)."""
        key = id(arr)
        hit = self._dev_cache.get(key)
        if hit is not None and hit[0] is arr:
            if hit[2] is not None:                       # uploaded ahead of time on the copy stream (prefetch_files)
                torch.cuda.current_stream().wait_event(hit[2])
                self._dev_cache[key] = (arr, hit[1], None)
            return hit[1]
        src = torch.from_numpy(np.ascontiguousarray(arr))
        if src.dtype != torch.int64:
            src = src.to(torch.int64)
        t, _ = self._file_buffer(src.shape)              # (a recycled buffer was released on this same stream: already ordered)
        t.copy_(src, non_blocking=False)
        self._cache_put(key, arr, t, None)
        return t

    def _file_buffer(self, shape):
        """Device buffer for one period file, from a pool that is never returned to the allocator: a fresh 600 MB block per
        period costs up to 180 ms of cudaMalloc / allocator housekeeping (measured), a recycled one nothing.
        -> (int64 view of the requested shape, event after which the buffer may be overwritten or None)."""
        self._make_room()
        n = 1
        for d in shape:
            n *= int(d)
        pick = None
        for i, (buf, _) in enumerate(self._free_bufs):
            # smallest sufficient buffer, but never a 600 MB one for a 1 MB file (the next large file would allocate afresh)
            if n <= buf.numel() <= 4 * max(n, 1) and (pick is None or buf.numel() < self._free_bufs[pick][0].numel()):
                pick = i
        if pick is None:
            # first file of this size class: allocate every buffer the cache will ever hold for it now, while little or nothing
            # is queued on the GPU (a cudaMalloc of this size has been seen to take 50-180 ms once the host runs ahead)
            same = sum(1 for c in self._buf_sizes if n <= c <= 4 * max(n, 1))
            want = min(8, max(1, self.dev_cache_cap + 1 - same)) if n >= (1 << 17) else 1
            fresh = [torch.empty(max(n, 1), dtype=torch.int64, device=self.device) for _ in range(want)]
            ev = torch.cuda.Event()                      # the blocks may have been freed by work still queued on this stream
            ev.record(torch.cuda.current_stream())
            self._buf_sizes.extend(b.numel() for b in fresh)
            buf = fresh.pop()
            self._free_bufs.extend((b, ev) for b in fresh)
        else:
            buf, ev = self._free_bufs.pop(pick)
        t = buf[:n].view(tuple(shape))
        t._sml_buf = buf
        return t, ev

    def _cache_put(self, key, arr, t, event):
        self._make_room()
        self._dev_cache[key] = (arr, t, event)

    def _make_room(self):
        """Evict the oldest cached files (their buffers go back to the pool) -- called BEFORE a new buffer is taken."""
        while len(self._dev_cache) > self.dev_cache_cap:
            _, old, copied = self._dev_cache.pop(next(iter(self._dev_cache)))
            buf = getattr(old, "_sml_buf", None)
            if buf is not None:                          # back to the pool once everything enqueued so far has run
                if copied is not None:                   # prefetched and never read: its upload may still be in flight
                    torch.cuda.current_stream().wait_event(copied)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                self._free_bufs.append((buf, ev))

    def prefetch_files(self, arrs):
        """Start the host->device copy of period files that a LATER stage will need, on a side stream, so that the
        transfer overlaps the current period's kernels (pinned host arrays copy asynchronously; pageable ones are staged
        by the driver and still overlap the GPU work already enqueued)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        for arr in arrs:
            if arr is None or (id(arr) in self._dev_cache and self._dev_cache[id(arr)][0] is arr):
                continue
            src = torch.from_numpy(np.ascontiguousarray(arr))
            if src.dtype != torch.int64:
                src = src.to(torch.int64)
            t, free_ev = self._file_buffer(src.shape)
            with torch.cuda.stream(self._copy_stream):
                if free_ev is not None:
                    self._copy_stream.wait_event(free_ev)
                t.copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._cache_put(id(arr), arr, t, ev)

    def _sample_dataset(self, arr):
        """SampleDaset(set_tt) is stateless (its draws come from the global numpy generator), so the
        per-period user->items index is built once per file instead of once per outer phase."""
        hit = getattr(self, "_tt_ds", None)
        if hit is None or hit[0] is not arr:
            self._tt_ds = (arr, SampleDaset(arr))
        return self._tt_ds[1]

    def _test_set(self, arr):
        return DeviceTestSet(self._to_device(arr), emulate_reference_rng=self.emulate_reference_rng)

    def _device_triples(self, ds, n_rows):
        """Whole-epoch triples built on the GPU (device_sampler=True)."""
        order = torch.randperm(n_rows, device=self.device)
        if isinstance(ds, PreSampleDatast):
            d = self._to_device(ds.all_data)
            col = ds.current_column()
            ds.advance_epoch()
            return [d[order, 0].contiguous(), d[order, 1].contiguous(), d[order, col].contiguous()]
        if not hasattr(ds, "_dev"):
            T = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(self.device)
            ds._dev = (T(ds.user), T(ds.item), T(ds.item_all), T(ds._keys))
        u_all, i_all, item_all, keys = ds._dev
        u, i = u_all[order].contiguous(), i_all[order].contiguous()
        self._philox_calls += 1
        neg = ops.philox_negatives(u, item_all, keys, ds._span, seed=2002, offset=self._philox_calls)
        return [u, i, neg]

    def _triples(self, kind, ds, stage_id, epoch, n_rows):
        if self.device_sampler and self.batch_source is None:
            return self._device_triples(ds, n_rows)
        if self.batch_source is not None:
            u, i, j = self.batch_source(kind, stage_id, epoch, n_rows)
            if isinstance(ds, PreSampleDatast):
                ds.advance_epoch()
        else:
            order = ReferenceStream.shuffled_order(n_rows) if self.emulate_reference_rng else np.random.permutation(n_rows)
            if isinstance(ds, PreSampleDatast):
                u, i, j = mf_epoch_triples(ds, order)
            else:
                u, i, j = tr_epoch_triples(ds, order)
        return u, i, j

    def _upload(self, arrs):
        """Host id arrays -> device, through pinned staging buffers on the copy stream: a pageable cudaMemcpyAsync on the compute
        stream would block the host until the GPU has drained everything enqueued before it (the host runs ahead, see ``defer``)."""
        if all(isinstance(x, torch.Tensor) for x in arrs):
            return list(arrs)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        out = []
        cur = torch.cuda.current_stream()
        used = []
        for x in arrs:
            if isinstance(x, torch.Tensor):
                out.append(x)
                continue
            a = np.ascontiguousarray(x, dtype=np.int64).reshape(-1)
            stage, slot = self._pinned(a.size)
            stage[:a.size].numpy()[...] = a
            with torch.cuda.stream(self._copy_stream):
                t = stage[:a.size].to(self.device, non_blocking=True).view(np.shape(x))
            t.record_stream(cur)
            used.append(slot)
            out.append(t)
        ev = torch.cuda.Event()
        ev.record(self._copy_stream)
        for slot in used:
            slot[1] = ev                               # the staging buffer is free again once this copy has run
        cur.wait_event(ev)
        return out

    def _pinned(self, n):
        """Pinned staging buffer of >= n int64 from a ring owned by this object.  torch's pinned allocator would do, but a
        NEW pinned allocation (cudaHostAlloc) synchronises the device: with the host running a whole period ahead of the GPU one
        such call was measured to cost 60 ms (the host loses its lead and its sampling work becomes visible)."""
        BUSY = "busy"                                  # handed out by this very _upload call, event not recorded yet
        for slot in self._pin_pool:
            if slot[0].numel() >= n and slot[1] is not BUSY and (slot[1] is None or slot[1].query()):
                slot[1] = BUSY
                return slot[0], slot
        if len(self._pin_pool) >= 128:                 # everything in flight: wait for the oldest copy instead of growing further
            for slot in self._pin_pool:
                if slot[0].numel() >= n and slot[1] is not BUSY and slot[1] is not None:
                    slot[1].synchronize()
                    slot[1] = BUSY
                    return slot[0], slot
        slot = [torch.empty(max(int(n * 1.25), 1 << 16), dtype=torch.int64, pin_memory=True), BUSY]
        self._pin_pool.append(slot)
        return slot[0], slot

    # ------------------------------------------------------------------ deferred reads
    def _defer_value(self, dev_tensor, post=None):
        """Keep a small result on the device; -> handle.  After flush_deferred ``values[handle]`` = post(host float32 tensor).
        Handles stay valid until the end of the period (or of the directly called *_onestage method)."""
        self._handle += 1
        self._pending[self._handle] = (dev_tensor.detach().reshape(-1).float(), post)
        return self._handle

    def _const_value(self, value):
        self._handle += 1
        self._resolved[self._handle] = value
        return self._handle

    def _print(self, *a):
        """print, in order with the deferred consumers."""
        self._later(lambda v: print(*a))

    def _later(self, fn):
        """Queue a consumer of deferred values (print, writer call, list append); runs at once when ``defer`` is off."""
        self._later_q.append(fn)
        if not self.defer:
            self.flush_deferred(final=False)

    def flush_deferred(self, final=True):
        """One device->host copy of every pending value, then the queued consumers in program order.  ``final``: no handle
        issued so far will be used again (end of a period): the resolved values are dropped."""
        pend, q = self._pending, self._later_q
        self._pending, self._later_q = {}, []
        if pend:
            host = torch.cat([t for t, _ in pend.values()]).cpu()
            off = 0
            for h, (t, post) in pend.items():
                x = host[off:off + t.numel()]
                off += t.numel()
                self._resolved[h] = post(x) if post is not None else x
        for fn in q:
            fn(self._resolved)
        if final:
            self._resolved = {}

    def invalidate(self):
        """Call after writing the MF tables from OUTSIDE this class through a path torch cannot see
        (``MFbase.*.weight.data`` edits, raw-pointer kernels): drops every kept evaluation pass.  Writes through
        this class's own methods, ``load_state_dict`` / in-place ops on the Parameters themselves are tracked."""
        self._tab_version += 1
        self._eval_cache = None

    def _tables_key(self):
        uw, iw = self.MFbase.user_laten.weight, self.MFbase.item_laten.weight
        return (self._tab_version, uw.data_ptr(), iw.data_ptr(), uw._version, iw._version)

    def _eval(self, test_set, topK):
        """test_model(self.MFbase, test_set, topK) -> handle of the deferred (recall, ndcg) pair (see ``defer``)."""
        t0 = time.perf_counter()
        # The reference re-scores a file even when nothing changed since its last evaluation (the "before train MF" pass
        # of outer phase p+1 repeats the last pass of phase p, model/transfer.py:445 vs :740; the three K of the real
        # test, :855-868).  A scoring pass is reused ONLY when the very same device file is evaluated again and nothing
        # has written the MF tables in between; the decision is taken afresh on every call (a hit never outlives the
        # call that found it), and the reference's RNG draw per evaluation is still consumed.
        if not isinstance(test_set, DeviceTestSet):
            r = test_model(self.MFbase, test_set, topK=topK)
            self.eval_passes["scored"] += 1
            self.timers["eval"] += time.perf_counter() - t0
            return self._const_value(r)
        c = self._eval_cache
        reused = c is not None and c[0] is test_set.rows and c[1] == self._tables_key()
        test_set._rank_cache = c[2] if reused else None
        test_set.frozen = reused
        # "eval" = evaluations that launch the scoring kernel, "eval_reused" = those that only reduce a kept rank pass
        with self.events("eval_reused" if reused else "eval"):
            # evalution/evaluation2.py:8-26 on a device-resident file: one scoring pass, per-1024-row reduction, sums
            self.MFbase.eval()
            if test_set.emulate_reference_rng:
                ReferenceStream.loader_iter()          # the reference creates one DataLoader iterator per call
            n = len(test_set)
            if n == 0:
                h = self._const_value((0.0, torch.tensor(0.0)))
            else:
                h = self._defer_value(self._score_sums(test_set, topK), lambda v, n=n: (float(v[0]) / n, (v[1] / n).clone()))
        test_set.frozen = False
        if test_set._rank_cache is not None:      # the cache holds the file tensor itself: no address aliasing
            self._eval_cache = (test_set.rows, self._tables_key(), test_set._rank_cache)
        self.eval_passes["reused" if reused else "scored"] += 1
        self.timers["eval"] += time.perf_counter() - t0
        return h

    def _score_sums(self, test_set, topK):
        """[hits, sum of NDCG terms] of one evaluation as a device tensor (one scoring pass unless a kept one is reused)."""
        gt, eq = test_set.ranks(self.MFbase)
        hits, ndcg = ops.eval_reduce(gt, eq, topK, batch=test_set.batch)
        return torch.stack([hits.sum().float(), ndcg.sum()])

    def get_next_data(self, stage_id):
        set_t, set_tt, now_test, val = self.dataset.next_train(stage_id)
        return set_t, set_tt, now_test, val

    # ------------------------------------------------------------------ MF (inner) training
    def MF_train_onestage(self, args, set_t, stage_id, val=None):
        """reference: model/transfer.py:417-534 (transfer fixed, MFbase trained through it)."""
        self.transfer.eval()
        if val is not None:
            val = self._test_set(val)
        self._print("******MF (inner) training ******")
        set_t_ds = self.MF_TrainDataset(set_t)
        if val is not None:
            h = self._eval(val, args.topK)
            self._later(lambda v, h=h: print("before train MF test:recall:{:.4f} ndcg:{:.4f}".format(*v[h])))
            if self.need_writer:
                self._write_scalars([("Acc/MF-recall" + str(args.topK), h, 0), ("Acc/MF-ndcg" + str(args.topK), h, 1),
                                     ("norm/user-norm", self._user_norm(), None)], self.MF_itr)
                self.MF_itr += 1
        for epoch in range(args.MF_epochs):
            self.MFbase.train()
            self.transfer.eval()
            t0 = time.perf_counter()
            triples = self._triples("MF", set_t_ds, stage_id, epoch, len(set_t_ds))
            with self.events("mf_epoch"):
                hl = self._mf_epoch(args, triples)                                  # loss_all / MF_batch_size (:514-515)
            self.timers["mf"] += time.perf_counter() - t0
            if val is not None:
                h = self._eval(val, args.topK)
                self._later(lambda v, h=h, hl=hl, epoch=epoch: print(
                    "MF-stage:", stage_id, "epoch:", epoch, "loss:{:.5f}".format(v[hl]), "recall:{:.4f}".format(v[h][0]),
                    "ndcg:{:.4f}".format(v[h][1])))
                if self.need_writer:
                    self._write_scalars([("Acc/MF-recall" + str(args.topK), h, 0), ("Acc/MF-ndcg" + str(args.topK), h, 1),
                                         ("Loss/MF-loss", hl, None)], self.MF_itr)
            else:
                self._later(lambda v, hl=hl, epoch=epoch: print("MF-stage:", stage_id, "epoch:", epoch, "loss:", v[hl]))
            if self.need_writer:
                self._write_scalars([("norm/user-norm", self._user_norm(), None)], self.MF_itr)
                self.MF_itr += 1
            self._later(lambda v, hl=hl: setattr(self, "last_MF_loss", v[hl]))
        if self._stage_depth == 0:
            self.flush_deferred()

    def _user_norm(self):
        return self._defer_value((self.MFbase.user_laten.weight.data ** 2).sum(dim=-1).mean(), lambda v: float(v[0]))

    def _write_scalars(self, items, itr):
        """writer.add_scalar(name, value, itr) for deferred values: items = [(name, handle, index into the value or None)]."""
        def emit(v):
            for name, h, k in items:
                self.writer.add_scalar(name, v[h] if k is None else v[h][k], itr)
        self._later(emit)

    def _mf_epoch(self, args, triples):
        """HOT LOOP A (model/transfer.py:463-511) over one epoch of triples; returns the handle of the deferred epoch loss
        (mean batch loss / MF_batch_size, :514-515)."""
        user, item, neg = self._upload(triples)
        uw, iw = self.MFbase.user_laten.weight.data, self.MFbase.item_laten.weight.data
        B = int(args.MF_batch_size)
        ws = self._workspace(B)
        n = user.numel()
        self._loss.zero_()
        nb = -(-n // B)
        lr = self.MF_optimizer.param_groups[0]["lr"]
        # --need_adaptive (model/transfer.py:490-499, beta = 0.1).  --clip_grad in this loop clips the gradients of theta,
        # which no optimizer applies here (:507-511): nothing to do.  --norm reaches run_MF's BPR branch only (BCE ignores it).
        adaptive = 0.1 if getattr(args, "need_adaptive", False) else 0.0
        def build(u, i, j):
            return ops.make_step_args(user=u, item=i, neg=j, batch=B,
                                      last_user=self.last_user_weight, last_item=self.last_item_weight,
                                      hat_user=uw, hat_item=iw, theta=self.transfer.theta, variant=self.transfer.variant,
                                      loss=ops.LOSS_BCE if self.transfer.variant == ops.VARIANT_COM else ops.LOSS_BPR,
                                      adam_state=self.MF_optimizer.adam_state, lr=lr, l2=args.l2, loss_out=self._loss,
                                      adaptive_beta=adaptive, workspace=ws, **self._mf, **self._stamps)
        self._tab_version += 1
        # every pointer / scalar baked into the captured StepArgs is part of the graph key
        self._run_epoch("mf", build, (user, item, neg), n, B,
                        (lr, args.l2, self.transfer.variant, adaptive) + tuple(t.data_ptr() for t in (
                            uw, iw, self.last_user_weight, self.last_item_weight, self.transfer.theta, ws,
                            self.MF_optimizer.adam_state, *self._mf.values(), *self._stamps.values())))
        return self._defer_value(self._loss[1:2].clone(), lambda v, d=float(nb) * args.MF_batch_size: float(v[0]) / d)

    # ------------------------------------------------------------------ transfer (outer) training
    def transfer_train_onestage(self, args, set_tt, stage_id, compute_performance=False, val=None):
        """reference: model/transfer.py:644-749 (embeddings fixed, theta trained)."""
        self._print("********* this is Transfer model training stage ***********")
        self.MFbase.eval()
        now_test = None
        if self.TR_train_sampleTYpe == "alone":
            set_tt_ds = self._sample_dataset(set_tt)
            compute_performance = False
            if val is not None:
                now_test = self._test_set(val)
                compute_performance = True
        elif self.TR_train_sampleTYpe == "all":
            now_test = self._test_set(set_tt)
            set_tt_ds = PreSampleDatast(set_tt)
            compute_performance = True
        else:
            raise TypeError("no such TR sample type")
        if compute_performance:
            h = self._eval(now_test, args.topK)
            self._later(lambda v, h=h: print("before train transfer test:recall:{:.4f} ndcg:{:.4f}".format(*v[h])))
            if self.need_writer:
                self._write_scalars([("Acc/tr-TR-recall@" + str(args.topK), h, 0), ("Acc/tr-TR-ndcg@" + str(args.topK), h, 1)], self.TR_itr)
                self.TR_itr += 1
        s_time = time.time()
        for epoch in range(args.TR_epochs):
            self.transfer.train()
            t0 = time.perf_counter()
            triples = self._triples("TR", set_tt_ds, stage_id, epoch, len(set_tt_ds))
            with self.events("tr_epoch"):
                hl = self._tr_epoch(args, triples)                                  # loss_all (mean batch loss)
            self.timers["tr"] += time.perf_counter() - t0
            self._later(lambda v, dt=time.time() - s_time: print("one epcohs TR time cost:", dt))
            if self.need_writer:
                self._write_scalars([("Loss/TR-loss", self._scaled(hl, 1.0 / args.TR_batch_size), None)], self.TR_itr)
            if compute_performance:
                self.updata()
                h = self._eval(now_test, args.topK)
                self._later(lambda v, h=h, hl=hl, epoch=epoch: print("stage:{}, epcoh：{}，loss:{:.4f},*****val result  reacll:{:.4f}  ndcg:{:.4f}".format(
                    stage_id, epoch, v[hl] / args.TR_batch_size, v[h][0], v[h][1])))
                if self.need_writer:
                    self._write_scalars([("Acc/tr-TR-recall@" + str(args.topK), h, 0), ("Acc/tr-TR-ndcg@" + str(args.topK), h, 1)], self.TR_itr)
            else:
                self._later(lambda v, hl=hl, epoch=epoch: print("stage:", stage_id, "epoch:", epoch, "transfer train loss:", v[hl] / args.TR_batch_size))
            self._later(lambda v, hl=hl: setattr(self, "last_TR_loss", v[hl]))
        self._print("stage ", stage_id, " transfer trained finished!!!!")
        if self._stage_depth == 0:
            self.flush_deferred()

    def _scaled(self, h, k):
        """Handle of k * (the deferred scalar behind handle h)."""
        self._handle += 1
        hs = self._handle

        def fill(v, h=h, hs=hs, k=k):
            v[hs] = v[h] * k
        self._later_q.append(fill)
        return hs

    def _run_epoch(self, kind, build, triples, n, B, key_extra):
        """Enqueue one epoch: directly the first time a (kind, n, B, hyper-parameters) combination is seen, as a CUDA
        graph replay afterwards (the triples are copied into the graph's static id buffers first)."""
        run = ops.mf_epoch if kind == "mf" else ops.tr_epoch
        key = (kind, n, B) + tuple(key_extra)
        if not self.use_graphs or n == 0:
            run(build(*triples), n)
            return
        if key not in self._graph_warm:            # first sight: plain launches (also performs one-time kernel setup)
            self._graph_warm.add(key)
            run(build(*triples), n)
            return
        entry = self._graphs.get(key)
        if entry is None:
            bufs = [torch.empty_like(t) for t in triples]
            a = build(*bufs)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            n0 = ops.lib().sml_launch_count()
            with torch.cuda.graph(graph):
                run(a, n)
            nodes = int(ops.lib().sml_launch_count() - n0)      # captured, not executed
            self.graph_launches -= nodes
            if len(self._graphs) >= 6:
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._graphs[key] = (graph, bufs, a, nodes)
        graph, bufs, _, nodes = entry
        for b, t in zip(bufs, triples):
            b.copy_(t)
        graph.replay()
        self.graph_launches += nodes

    def _tr_epoch(self, args, triples):
        """HOT LOOP B (model/transfer.py:701-728) over one epoch of triples; returns the handle of the deferred mean batch loss."""
        user, item, neg = self._upload(triples)
        B = int(args.TR_batch_size)
        ws = self._workspace(B)
        n = user.numel()
        self._loss.zero_()
        nb = -(-n // B)
        g = self.transfer_optimizer.param_groups[0]
        clip = float(args.maxnorm_grad) if getattr(args, "clip_grad", False) else 0.0       # model/transfer.py:723-727
        def build(u, i, j):
            return ops.make_step_args(user=u, item=i, neg=j, batch=B,
                                      last_user=self.last_user_weight, last_item=self.last_item_weight,
                                      hat_user=self.user_weight_hat, hat_item=self.item_weight_hat,
                                      theta=self.transfer.theta, variant=self.transfer.variant,
                                      loss=ops.LOSS_BCE if self.transfer.variant == ops.VARIANT_COM else ops.LOSS_BPR,
                                      adam_state=self.transfer_optimizer.adam_state, lr=g["lr"], l2=g["weight_decay"],
                                      g_theta=self.transfer.theta_grad, m_theta=self._tr["m"], v_theta=self._tr["v"],
                                      loss_out=self._loss, workspace=ws, clip_max_norm=clip)
        self._run_epoch("tr", build, (user, item, neg), n, B,
                        (g["lr"], g["weight_decay"], self.transfer.variant, clip) + tuple(t.data_ptr() for t in (
                            self.transfer.theta, self.transfer.theta_grad, self._tr["m"], self._tr["v"], ws,
                            self.last_user_weight, self.last_item_weight, self.user_weight_hat, self.item_weight_hat,
                            self.transfer_optimizer.adam_state)))
        return self._defer_value(self._loss[1:2].clone(), lambda v, d=float(nb): float(v[0]) / d)

    # ------------------------------------------------------------------ one period
    def _real_test(self, now_test_arr):
        """The three test_model calls at K = 20, 10, 5 (model/transfer.py:810-823,855-868)."""
        self.test_num.append(now_test_arr.shape[0])
        now_test = self._test_set(now_test_arr)       # three K on unchanged tables: _eval keeps one scoring pass
        for K, rl, nl, tag in ((20, self.recall, self.ndcg, ""), (10, self.recall_10, self.ndcg_10, " @10"),
                               (5, self.recall_5, self.ndcg_5, " @5")):
            h = self._eval(now_test, K)

            def consume(v, h=h, rl=rl, nl=nl, tag=tag):
                recall, ndcg = v[h]
                print("test result ---------{} reacll:{:.4f}  ndcg:{:.4f}".format(tag, recall, ndcg))
                rl.append(recall)
                nl.append(ndcg.cpu().numpy())
            self._later(consume)

    def train_one_stage3(self, args, stage_id):
        """reference: model/transfer.py:753-881 -- one period, three branches.  Losses / metrics of the period are read back
        once, at its end (see ``defer``): its prints come out then, in the reference's order."""
        self._stage_depth += 1
        try:
            return self._train_one_stage3(args, stage_id)
        finally:
            self._stage_depth -= 1
            self.flush_deferred()

    def _train_one_stage3(self, args, stage_id):
        self.save_MF_weight(save_as="last")
        set_t, set_tt, now_test, val = self.get_next_data(stage_id)
        if set_t is None:
            return False
        if self.prefetch and self.device.type == "cuda" and hasattr(self.dataset, "peek_files"):
            # this stage's files first (in the order they are used), then the next stage's: their upload overlaps this period
            self.prefetch_files([set_t, val, now_test] + list(self.dataset.peek_files(stage_id + 1)))
        if now_test is None:                       # online training, no real test yet (:772-792)
            for phase in range(args.multi_num):
                self.MF_train_onestage(args, set_t, stage_id, val=val)
                self.MFbase.eval()
                self.save_MF_weight(save_as="hat")
                self.updata()
                self.transfer_train_onestage(args, set_tt, stage_id, val=val)
                if args.Load_W_hat:
                    self.load_MFbase_weight(self.user_weight_hat, self.item_weight_hat)
            self.updata()
            return True
        elif set_tt is None:                       # transfer frozen while testing (:793-825)
            s_time = time.time()
            self._print("stop train transfer while test###!!!!!")
            args.MF_epochs = 2                     # the reference mutates args here (:796)
            self.MF_train_onestage(args, set_t, stage_id, val=val)
            self.MFbase.eval()
            self.save_MF_weight(save_as="hat")
            self.updata()
            self._print("only traning time cost:", time.time() - s_time)
            self._real_test(now_test)
            self._print("include test time cost:", time.time() - s_time)
            return True
        else:                                      # test on D_{t+1}, then train the transfer on it (:826-881)
            for phase in range(args.multi_num):
                self.MF_train_onestage(args, set_t, stage_id, val=val)
                self.MFbase.eval()
                self.save_MF_weight(save_as="hat")
                self.updata()
                if phase == 0:
                    self._real_test(now_test)
                self.transfer_train_onestage(args, set_tt, stage_id, val=val)
                if args.Load_W_hat:
                    self.load_MFbase_weight(self.user_weight_hat, self.item_weight_hat)
            self.updata()
            return True

    # ------------------------------------------------------------------ table bookkeeping
    def updata(self):
        """w_t = Transfer(w_{t-1}, w_hat) for EVERY row of both tables, written straight into the
        MFbase tables (reference: model/transfer.py:884-902 + load_MFbase_weight)."""
        t0 = time.perf_counter()
        self._tab_version += 1
        self.MFbase.eval()
        self.transfer.eval()
        if self.transfer_type != "transfer2":
            raise TypeError("No such type transfer!!!")
        th = self.transfer.theta
        nu = self.transfer.variant == ops.VARIANT_CONV
        with self.events("updata"):
            ops.transfer_forward(self.last_user_weight, self.user_weight_hat, th[:ops.NET_STRIDE], variant=self.transfer.variant,
                                 normalize_out=nu, out=self.MFbase.user_laten.weight.data)
            ops.transfer_forward(self.last_item_weight, self.item_weight_hat, th[ops.NET_STRIDE:], variant=self.transfer.variant,
                                 out=self.MFbase.item_laten.weight.data)
        self.timers["updata"] += time.perf_counter() - t0

    def save_MF_weight(self, save_as="last"):
        """reference: model/transfer.py:911-943."""
        if save_as == "last":
            self.last_user_weight.copy_(self.MFbase.user_laten.weight.data)
            self.last_item_weight.copy_(self.MFbase.item_laten.weight.data)
        elif save_as == "hat":
            self.last_user_weight_hat.copy_(self.user_weight_hat.data)
            self.last_item_weight_hat.copy_(self.item_weight_hat.data)
            self.user_weight_hat.copy_(self.MFbase.user_laten.weight.data)
            self.item_weight_hat.copy_(self.MFbase.item_laten.weight.data)
        else:
            raise TypeError("save MFbase weight type is wrong")

    def load_MFbase_weight(self, user_weight, item_weight):
        """reference: model/transfer.py:945-959."""
        self._tab_version += 1
        self.MFbase.user_laten.weight.data.copy_(user_weight)
        self.MFbase.item_laten.weight.data.copy_(item_weight)

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything a mid-stream resume needs (the reference never saves Adam state):
        MF tables, the six snapshots, theta, both Adam states and step counters, metric lists, RNG states."""
        c = lambda t: t.detach().clone()
        return dict(
            MFbase={k: c(v) for k, v in self.MFbase.state_dict().items()},
            transfer={k: c(v) for k, v in self.transfer.state_dict().items()},
            snapshots={k: c(getattr(self, k)) for k in ("last_user_weight", "last_item_weight", "user_weight_hat", "item_weight_hat",
                                                        "last_user_weight_hat", "last_item_weight_hat")},
            mf_adam={k: c(v) for k, v in self._mf.items()}, mf_adam_state=c(self.MF_optimizer.adam_state),
            tr_adam={k: c(v) for k, v in self._tr.items()}, tr_adam_state=c(self.transfer_optimizer.adam_state),
            metrics={k: list(getattr(self, k)) for k in ("recall", "ndcg", "recall_10", "ndcg_10", "recall_5", "ndcg_5", "test_num")},
            dataset=dict(test_count=getattr(self.dataset, "test_count", 0)),
            rng=dict(torch=torch.get_rng_state(), numpy=np.random.get_state()), philox_calls=self._philox_calls)

    def load_state_dict(self, sd):
        self._tab_version += 1
        self.MFbase.load_state_dict(sd["MFbase"])
        self.transfer.load_state_dict(sd["transfer"])
        for k, v in sd["snapshots"].items():
            getattr(self, k).copy_(v)
        for k, v in sd["mf_adam"].items():
            self._mf[k].copy_(v)
        for k, v in sd["tr_adam"].items():
            self._tr[k].copy_(v)
        self.MF_optimizer.adam_state[:2].copy_(sd["mf_adam_state"][:2])      # step counter + step scalars
        self.transfer_optimizer.adam_state[:2].copy_(sd["tr_adam_state"][:2])
        for st in self._stamps.values():                                     # every epoch ends flushed: all rows at step t
            st.copy_(self.MF_optimizer.adam_state[0].to(torch.int32).expand_as(st))
        for k, v in sd["metrics"].items():
            setattr(self, k, list(v))
        if hasattr(self.dataset, "test_count"):
            self.dataset.test_count = sd["dataset"]["test_count"]
        torch.set_rng_state(sd["rng"]["torch"]); np.random.set_state(sd["rng"]["numpy"])
        self._philox_calls = sd.get("philox_calls", 0)

    # ------------------------------------------------------------------ whole stream
    def run(self, args):
        """reference: model/transfer.py:965-1029, including the final weighted summary with
        N3 = round(n/3) and the test slice [N3:-1] that drops the last test period."""
        pass_num = args.pass_num
        self.summary = {}
        for pass_id in range(pass_num):
            stage_id = 0
            self.dataset.reinit()
            while 1:
                flag = self.train_one_stage3(args, stage_id)
                if flag:
                    stage_id += 1
                    if pass_id < (pass_num - 1) and stage_id >= 19:
                        break
                else:
                    print(str(pass_id) + "--trained over!!!!!")
                    test_num = np.array(self.test_num)
                    N3 = round(test_num.shape[0] * 1 / 3)
                    val_num = test_num[0:N3]
                    test_num = test_num[N3:-1]
                    recall = np.array(self.recall)
                    ndcg = np.array(self.ndcg)
                    print(test_num)
                    print(recall)
                    print(ndcg)
                    print("include stage 0 of test:")
                    val_num = val_num * 1.0 / val_num.sum()
                    test_num = test_num * 1.0 / test_num.sum()
                    for K, rl, nl in ((20, self.recall, self.ndcg), (10, self.recall_10, self.ndcg_10), (5, self.recall_5, self.ndcg_5)):
                        recall = np.array(rl)
                        ndcg = np.array(nl)
                        s = dict(val_recall=(recall[0:N3] * val_num).sum(), val_ndcg=(ndcg[0:N3] * val_num).sum(),
                                 test_recall=(recall[N3:-1] * test_num).sum(), test_ndcg=(ndcg[N3:-1] * test_num).sum())
                        self.summary[K] = s
                        print("val average recall@%d:" % K, s["val_recall"])
                        print("val average ndcg@%d:" % K, s["val_ndcg"])
                        print("test average recall@%d:" % K, s["test_recall"])
                        print("test average ndcg@%d:" % K, s["test_ndcg"])
                        print("\n")
                    break
