"""Orchestration of one retrieval evaluation (codes -> ranked Hamming retrieval -> mAP@R / P@k / R@k).

The same code runs on 1 GPU and on a gallery row-sharded over G GPUs (one process per GPU, contiguous
row blocks in rank order, queries replicated).  The only data exchanged between ranks are

  * one all-gather of the per-rank key histograms  (nbins x nq u32, x2 when labels are counted),
  * (top-R mode) one all-gather of the relevant-candidate histograms,
  * one all-reduce(sum) of the per-query partial AP sums / hit counts (fp64),
  * a few scalars (flags, row counts) and, for ``remove_first_retrieved``, an nq-element max.

Stable ties "ascending gallery row index" == (rank, stripe, in-stripe prefix), which is why the gather
(not a plain sum) is needed: every rank takes the exclusive prefix over lower ranks.

All arithmetic happens in the backend (CUDA kernels behind the C-ABI); this file only sequences the
calls.  ``tests/`` substitute a numpy emulation of the backend to exercise this file on CPU with gloo.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from .codes_io import PackedCodes


class LocalComm:
    """world_size == 1: collectives are the identity."""
    world, rank = 1, 0

    def all_gather(self, t):
        return t.unsqueeze(0)

    def all_reduce_sum(self, t):
        return t

    def all_reduce_max(self, t):
        return t


class DistComm:
    """torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.timer = None      # the evaluator's event bracket (bench: time spent in collectives)

    def _t(self, kind, nbytes, fn):
        return fn() if self.timer is None else self.timer(kind, nbytes, fn)

    def all_gather(self, t):
        flat = t.contiguous().view(-1)
        out = torch.empty((self.world * flat.numel(),), dtype=t.dtype, device=t.device)
        self._t("comm_all_gather", out.numel() * out.element_size(),
                lambda: self.dist.all_gather_into_tensor(out, flat, group=self.group))
        return out.view((self.world,) + tuple(t.shape))

    def all_reduce_sum(self, t):
        self._t("comm_all_reduce", t.numel() * t.element_size(),
                lambda: self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group))
        return t

    def all_reduce_max(self, t):
        self._t("comm_all_reduce", t.numel() * t.element_size(),
                lambda: self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group))
        return t


def _as_int_list(t):
    return [int(v) for v in t.cpu().tolist()]


# Layout of the per-evaluation status block (u32 words on the device; read back ONCE, with the results):
ST_PASS = 0     # bit 0 slice overflow (select / record pass), bit 1 key outside the narrowed key range,
#                 bit 2 slot arena smaller than the capacities need, bit 3 a threshold beyond the narrowed key range
#                 (bits 2, 3: a stale hint of a speculative run -- the evaluation is repeated with host round trips)
ST_SHORT = 1    # bit 0: some query holds fewer than R candidates under its sampled threshold (-> exact path)
ST_CODES = 2    # ch_pack_sign flags: bit 0 some sign is exactly 0 (ternary keys), bit 1 NaN
ST_PACKED = 3   # group mode: 1 if this rank packed its gallery up front (then nobody streams)
# (total slots, max threshold) of the list / record allocations (ch_record_offsets_async): the sample-level list, and
# the one full-size allocation an evaluation makes (sampled pass, exact pass or full ranking)
ST_SITE = {"s1": 4, "full": 6, "exact": 8, "all": 8}
ST_QINFO, ST_GINFO = 10, 14      # ch_pack_labels statistics [max positives per row, max id + 1, rows w/o label, -]
ST_ROWS = 18    # gallery rows of every rank (group mode)
_RETRY = object()


class _TooManySlots(Exception):
    """the slices of one evaluation would need >= 2^32 slots: the query set is evaluated in chunks instead"""


class Packed:
    """Packed codes + labels of one side (query or gallery shard)."""
    __slots__ = ("n", "nbit", "bits", "nz", "ids", "masks", "info", "ncls", "i8", "i8b", "i8p", "plane", "loader", "pending")

    def __init__(self):
        self.plane = None
        self.pending = None      # labels not packed yet: (labels, nolabel, info) -- see Evaluator._ensure_labels
        self.loader = None       # native gallery loader filling ``loader.bits`` (host gallery; see Evaluator._prepare)
        self.i8b = None          # int8 plane without threshold slots (comparison in the select kernel's epilogue)
        self.i8p = None          # paired int8 plane: two gallery rows per plane row


class _TimedBackend:
    """Pass-through to the backend; while the evaluator profiles (bench), every entry point that is not already
    inside an explicit bracket gets its own CUDA-event bracket (kind = "k_<entry point>")."""
    _PLAIN = frozenset(("empty", "padded_rows", "code_words", "geometry", "tc_code_bytes", "tc_code_bytes_pair",
                        "launch_count", "begin",
                        "on_stream", "to_host", "gather_plane_words", "popc_peak"))

    def __init__(self, ev, backend):
        object.__setattr__(self, "_ev", ev)
        object.__setattr__(self, "_b", backend)

    def __getattr__(self, name):
        attr = getattr(self._b, name)
        ev = self._ev
        if not (ev.profile and ev.profile_all) or ev._bracket_open or name in self._PLAIN or not callable(attr):
            return attr
        return lambda *a, **k: ev._timed("k_" + name, 0, lambda: attr(*a, **k))

    def __setattr__(self, name, value):
        setattr(self._b, name, value)


class Evaluator:
    def __init__(self, backend, comm=None):
        self.profile = False               # bench: bracket the main kernels with CUDA events on the launch stream
        self.profile_all = False           # ... and every other entry point too (diagnostic pass: the brackets cost)
        self._bracket_open = False
        self.b = _TimedBackend(self, backend)
        self.comm = comm if comm is not None else LocalComm()
        if hasattr(self.comm, "timer"):
            self.comm.timer = self._timed
        self.stats = {}
        self.host_syncs = 0                # device -> host round trips of the current evaluation
        # Repeated evaluations of one shape (every eval_interval epochs, train_helper.py:273; the bench's steps) run
        # SPECULATIVELY: what the first evaluation had to ask the device for in mid-flight -- label form, list
        # sizes, the largest threshold -- is assumed to hold again, buffers are sized from it, the kernels verify
        # it on the device and the verdict comes back with the results in the ONE final round trip.  A failed
        # assumption repeats the evaluation the slow way; results never depend on the hint.
        self.speculate = True
        self.max_slots = 0xFFFFFFF0        # slices of one evaluation are addressed with 32-bit offsets
        self._hints = {}                   # shape key -> what the last evaluation of that shape found
        self._hint = None                  # the hint the running evaluation relies on (None: ask the device)
        self._new_hint = None
        self._status = None
        self.stripe_rows_override = None   # tests: force the stripe length (multiple of 256 on CUDA)
        self.sample_stride = 32            # top-R: 1-in-32 row sample picks the threshold (0/1 = exact two-pass)
        self.sample_two_level = True       # thresholds from the sample by a tensor-core select pass (see below)
        self.sample2_sub = 16              # ... whose own thresholds come from every 16th sample row
        self.sample2_min_rows = 2_048      # ... when every rank's sample has at least this many rows
        self.sample2_min_work = 2.0e8      # ... and histogramming it would cost >= ~0.05 ms (nq x rows x words); the
        #                                      two-level route adds one small all-reduce in group mode
        self.sample_min_rows = 200_000     # below this the two-pass path is cheap anyway
        self.sample_min_ratio = 64         # ... and the sample must still hold ~R/stride*... rows: need ndb >= ratio * R
        self.stream_host_gallery = True    # host-resident gallery: overlap its H2D copy with the select pass
        self.stream_min_rows = 200_000
        self.stream_chunks = 2             # blocks per wave-filling stripe group x 2 (fewer, larger blocks: ~0.15 ms of host work per block)
        self.stream_loader_thread = True   # pageable gallery: a loader thread packs / copies the blocks back to back
        self.stream_native_loader = True   # fp32 host gallery: packed from the first moment of the evaluation by a native
        #                                    thread (csrc/loader.cu); the blocks of the select pass only wait for their rows
        self.stream_chunks_native = 3      # ... blocks per wave-filling stripe group x 2 on that path (2: 4.97, 3: 4.65, 4: 5.0 ms on cfg4)
        self.stream_loader_labels = False  # ... 1-D label ids in pageable memory as raw-copy loader jobs (measured: no gain)
        self.stream_fused_rank = False     # ... one fused list kernel behind the last block instead of one per block
        self.stream_late_labels = False    # ... labels packed behind the first select launch (measured: no gain either)
        self.stream_cand_overlap = False   # ... list kernels beside the next block's select kernel (measured: no gain --
        #                                    they slow the select kernel by what they save)
        self._loader = None
        self._side_stream = None
        self._expand_stream = None
        self._cand_stream = None
        self.use_tensor_cores = True       # select pass on tcgen05 (int8 +-1 codes) when the shape allows it
        self.epilogue_thresholds = True    # sparse select passes: threshold comparison in the epilogue (see _bare)
        self.paired_rows = True            # select pass: two gallery rows per accumulator cell where keys fit (see _pair)
        self.debug_counts = False          # dev tools: stats["candidates"] = length of all candidate lists (costs a sync)
        # Small evaluations are bound by the host (15-30 launches of 3-50 us kernels behind ~0.4 ms of Python): the
        # speculative launch sequence of a shape -- fixed once its hint exists -- is captured into a CUDA graph and
        # replayed with ONE launch.  Inputs are copied into the graph's own buffers first (that is why only small
        # inputs qualify), the status block is checked after the replay exactly as after a speculative run, and any
        # contradiction falls back to the ordinary path.
        self.use_graphs = True
        self.graph_max_bytes = 96 << 20      # inputs up to this size are copied into graph-owned buffers
        self.graph_in_place = True           # larger ones: captured in place, replayed for the same addresses only
        # Group mode: the NCCL collectives are captured with the kernels (every rank replays or none does -- agreed in
        # the hint round trip).  Opt-in: a process group must not be destroyed while a graph that holds its collectives
        # is alive (the teardown hung when tried) -- call release_graphs() first, as bench.py does.
        self.graph_multi_gpu = os.environ.get("CH_GRAPH_MULTI_GPU", "0") == "1"
        self._graphs = {}
        self.graph_launches = 0            # kernels launched through graph replays (they bypass the C-side counter)
        self.fused_rank = True             # one rank: candidate histogram + bases + walk in one kernel (ch_cand_rank)
        self.stripe_cut = True             # sampled thresholds refined to (key, stripe) pairs (see _sample_thresholds_tc)
        self.select_dense_override = None  # tests: force the dense / sparse epilogue of the tensor-core kernel
        self.col_sub = None                # zero_mean_eval: f64[nbit] column offset of the current evaluation
        self.events = []                   # (kind, work units, start event, end event)

    def _host_ints(self, t):
        """device -> host read of a small integer tensor (ONE round trip; counted)"""
        self.host_syncs += 1
        return _as_int_list(t)

    def _timed(self, kind, units, fn):
        if not self.profile or self._bracket_open:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._bracket_open = True
        try:
            out = fn()
        finally:
            self._bracket_open = False
        e1.record()
        self.events.append((kind, units, e0, e1))
        return out

    # ------------------------------------------------------------------ packing
    def _pack_codes(self, p, codes, threshold, flags, want_nz=False):
        if isinstance(codes, PackedCodes):
            # already sign bits (codes_io): K1 is skipped; only the zero pad rows are added
            if threshold != 0 or self.col_sub is not None:
                raise ValueError("packed codes carry no magnitudes: threshold / zero_mean_eval need real-valued codes")
            bits = self.b.zeros((self.b.padded_rows(p.n), codes.bits.shape[1]), torch.int32)
            bits[:p.n] = codes.bits.to(bits.device)
            p.bits, p.nz = bits, None
            return
        pack_bytes = p.n * p.nbit * codes.element_size() + p.n * p.nbit // 8
        kind = "pack_dev" if codes.is_cuda else "pack_host"
        kw = {} if self.col_sub is None else dict(col_sub=self.col_sub)
        p.bits, p.nz = self._timed(kind, pack_bytes, lambda: self.b.pack_sign(codes, threshold, flags, want_nz, **kw))

    def _ensure_labels(self, *sides):
        """packs the labels whose packing ``_pack_side`` put off (``defer_labels``)"""
        for p in sides:
            if p.pending is not None:
                (labels, nolabel, info), p.pending = p.pending, None
                p.ids, p.masks, p.info = self.b.pack_labels(labels, nolabel, info)

    def _pack_side(self, codes, labels, threshold, flags, nolabel, want_nz=False, info=None, defer_codes=False,
                   defer_labels=False):
        p = Packed()
        p.i8 = None
        p.n, p.nbit = int(codes.shape[0]), int(codes.shape[1])
        # algorithmic bytes of K1: the real-valued codes read once + the packed bits written once
        p.bits = p.nz = None
        if not defer_codes:
            self._pack_codes(p, codes, threshold, flags, want_nz)
        p.ids = p.masks = p.info = None
        p.ncls = 0
        if labels is not None:
            if labels.shape[0] != p.n:
                raise ValueError(f"{labels.shape[0]} label rows for {p.n} code rows")
            if defer_labels:
                p.pending = (labels, nolabel, info)
            else:
                p.ids, p.masks, p.info = self.b.pack_labels(labels, nolabel, info)
            p.ncls = int(labels.shape[1]) if labels.dim() == 2 else 0
        return p

    def _column_mean(self, db_codes):
        """``db_codes.mean(dim=0)`` of the WHOLE gallery (all ranks), fp64 sums / row count, rounded to the dtype of
        the codes (what ``codes - mean`` sees in torch); experiments/train_helper.py:223-226."""
        b, comm = self.b, self.comm
        nbit = int(db_codes.shape[1])
        acc = b.zeros((nbit + 1,), torch.float64)
        if db_codes.shape[0] > 0:
            acc[:nbit] = b.column_sums(db_codes)
        acc[nbit] = float(db_codes.shape[0])
        if comm.world > 1:
            acc = comm.all_reduce_sum(acc)
        mean = acc[:nbit] / torch.clamp(acc[nbit], min=1.0)
        dt = db_codes.dtype if db_codes.dtype.is_floating_point else torch.float32
        return mean.to(dt).to(torch.float64).contiguous()

    def _prepare(self, db_codes, db_labels, q_codes, q_labels, threshold, allow_defer=False, zero_mean=False,
                 sample_stride=0):
        if q_codes.dim() != 2 or db_codes.dim() != 2:
            raise ValueError("codes must be 2-D (N, nbit)")
        if q_codes.shape[1] != db_codes.shape[1]:
            raise ValueError(f"nbit mismatch: query {q_codes.shape[1]} vs gallery {db_codes.shape[1]}")
        if q_codes.shape[0] == 0:
            raise ValueError("no queries")
        if (zero_mean or threshold != 0) and (isinstance(db_codes, PackedCodes) or isinstance(q_codes, PackedCodes)):
            raise ValueError("packed codes carry no magnitudes: threshold / zero_mean_eval need real-valued codes")
        self.col_sub = None
        if zero_mean:
            # zero_mean_eval fused into the pack kernel: the gallery's column mean is subtracted from both sets
            # on the fly.  The mean needs the whole gallery first, so a host gallery is copied once, not streamed.
            if hasattr(self.b, "device") and not db_codes.is_cuda:
                db_codes = db_codes.to(self.b.device, non_blocking=True)
            self.col_sub = self._column_mean(db_codes)
            allow_defer = False
        self._db_codes_eff = db_codes
        if (q_labels is None) != (db_labels is None):
            raise ValueError("labels must be given for both sides or neither")
        if q_labels is not None and q_labels.dim() == 2 and db_labels.dim() == 2 and \
                q_labels.shape[1] != db_labels.shape[1]:
            raise ValueError("query and gallery labels have different class counts")
        # every small status word of the evaluation lives in ONE tensor (layout: ST_* above) -> one device-to-host
        # read.  A speculative run (self._hint) does not read it here at all: the label form, the ternary flag and
        # the rows of the other ranks are taken from the hint and verified against the block after the final sync.
        st = self._status = self.b.zeros((ST_ROWS + self.comm.world,), torch.int32)
        flags = st[ST_CODES:ST_CODES + 1]
        # a large HOST gallery is not copied yet: the top-R path streams it in row blocks behind the select pass
        defer = (allow_defer and self.stream_host_gallery and not isinstance(db_codes, PackedCodes) and
                 not db_codes.is_cuda and threshold == 0 and
                 db_labels is not None and db_codes.shape[0] >= self.stream_min_rows and
                 hasattr(self.b, "hamming_select_tc") and self.b.tc_code_bytes(int(db_codes.shape[1])) > 0)
        loader = None
        if (defer and self.stream_native_loader and hasattr(self.b, "host_loader_start") and
                self.b.host_loader_ok(db_codes)):
            # fp32 rows in host memory (pageable or pinned): the host's cores are the bottleneck of this evaluation
            # (they read every code once), so they start NOW, on a native thread of their own, and run without a
            # pause to the last row (csrc/loader.cu): first the queries, then the row sample (both feed the
            # thresholds the GPU needs before it can touch the gallery), then the gallery.  This thread waits
            # ~0.3 ms for the first two (less than packing them itself took) and queues the GPU's work meanwhile.
            # Zeros / NaNs of the gallery and the sample raise ST_SHORT (the streamed pass then falls back) and come
            # back from join(); those of the queries go to ST_CODES like those of any query pack.
            b = self.b
            nbit = int(db_codes.shape[1])
            jobs, idx = [], {}
            if self.col_sub is None and b.host_loader_ok(q_codes):
                idx["q"] = len(jobs)
                jobs.append((q_codes, b.empty((b.padded_rows(int(q_codes.shape[0])), b.code_words(nbit)), torch.int32),
                             flags))
            if sample_stride > 1:
                run = max(1, 256 // nbit) if nbit in (32, 64, 128, 256) else 1
                ns, view = self._host_sample_view(db_codes, sample_stride, run)
                if ns > 0 and b.host_loader_ok(view):
                    idx["s"] = len(jobs)
                    jobs.append((view, b.empty((b.padded_rows(ns // run), b.code_words(run * nbit)), torch.int32),
                                 st[ST_SHORT:ST_SHORT + 1]))
                    idx["sample"] = (sample_stride, run, ns)
            # 1-D class ids in pageable memory (8 MB for a 1M-row gallery) travel as loader jobs too: packing them cost
            # this thread ~0.3 ms between the queries and the first sample pass
            if self.stream_loader_labels:
                for key, lab in (("ql", q_labels), ("gl", db_labels)):
                    if (isinstance(lab, torch.Tensor) and not lab.is_cuda and lab.dim() == 1 and lab.is_contiguous()
                            and lab.dtype in (torch.int64, torch.int32, torch.float32) and lab.numel() >= 4096
                            and lab.data_ptr() % 4 == 0 and not lab.is_pinned()):
                        idx[key] = len(jobs)
                        jobs.append(("copy", lab, b.empty((lab.numel(),), lab.dtype)))
            idx["g"] = len(jobs)
            jobs.append((db_codes, b.empty((b.padded_rows(int(db_codes.shape[0])), b.code_words(nbit)), torch.int32),
                         st[ST_SHORT:ST_SHORT + 1]))
            if getattr(self, "_side_stream", None) is None:
                # the loader's copy stream carries NOTHING else while a loader runs (a wait queued on it would sit in
                # front of the copies it waits for); the expansion kernels behind the copies have a stream of their
                # own (high priority: they must not queue behind the thousands of blocks of a list kernel --
                # measured 0.02 -> 0.34 ms each, and the next select waits for them)
                self._side_stream = torch.cuda.Stream(device=b.device, priority=-1)
            if getattr(self, "_expand_stream", None) is None:
                self._expand_stream = torch.cuda.Stream(device=b.device, priority=-1)
            self._side_stream.wait_stream(torch.cuda.current_stream())      # (the status block is zeroed on this stream)
            loader = self._loader = b.host_loader_start(jobs, self._side_stream)
            loader.idx = idx
        # With the loader running and a hint for the label form, the labels are not on the way to the thresholds (the
        # sample passes rank by keys alone): they are packed after those passes have been queued (0.4 ms of a 5 ms
        # evaluation sat between the query pack and the row sample otherwise) -- _ensure_labels.
        late = (self.stream_late_labels and loader is not None and self._hint is not None and self.comm.world == 1 and
                max(self._hint["mm"][ST_QINFO], self._hint["mm"][ST_GINFO]) <= 1)         # (single-label form)
        q_bits = None
        if loader is not None and "q" in loader.idx:
            q_bits = loader.bits[loader.idx["q"]]
            loader.wait(loader.idx["q"], int(q_codes.shape[0]), torch.cuda.current_stream())
        def landed(key, lab):
            """the device copy of a label array the loader carries (the current stream waits for it), or the array"""
            if loader is None or key not in loader.idx:
                return lab
            loader.wait(loader.idx[key], int(lab.numel()), torch.cuda.current_stream())
            return loader.bits[loader.idx[key]]
        q = self._pack_side(q_codes, landed("ql", q_labels), threshold, flags, L.CH_QUERY_NOLABEL,
                            info=st[ST_QINFO:ST_QINFO + 4], defer_labels=late, defer_codes=q_bits is not None)
        if q_bits is not None:
            q.bits = q_bits
        g = self._pack_side(db_codes, landed("gl", db_labels), threshold, flags, L.CH_GALLERY_NOLABEL,
                            info=st[ST_GINFO:ST_GINFO + 4], defer_codes=defer, defer_labels=late)
        g.loader = loader
        if self.comm.world > 1:
            # ranks must agree: everything by MAX (each rank fills only its own slot of the row counts)
            # (fill_: the value travels as a kernel argument -- an indexed assignment of a Python int is a pageable
            # host-to-device copy, which a CUDA graph cannot record)
            st[ST_ROWS + self.comm.rank:ST_ROWS + self.comm.rank + 1].fill_(g.n)
            if g.bits is not None:
                st[ST_PACKED:ST_PACKED + 1].fill_(1)
        if self._hint is not None:
            mm, rows, any_packed = self._hint["mm"], self._hint["rows"], self._hint["any_packed"]
        else:
            if self.comm.world > 1:
                self.comm.all_reduce_max(st)
            mm = self._host_ints(st)
            rows = mm[ST_ROWS:] if self.comm.world > 1 else [g.n]
            any_packed = bool(mm[ST_PACKED])
            mm = mm[:ST_ROWS]
        self._new_hint.update(mm=list(mm), rows=list(rows), any_packed=any_packed)
        if self.comm.world > 1 and any_packed and g.bits is None:
            if g.loader is not None:
                fl = self._loader_bits(g)
                if fl:
                    flags.bitwise_or_(fl)
                    st[ST_SHORT:ST_SHORT + 1].zero_()
            else:
                self._pack_codes(g, db_codes, threshold, flags)      # (zeros / NaN are re-checked by the caller)
        mm = [mm[ST_CODES], 0, 0, 0] + list(mm[ST_QINFO:ST_QINFO + 4]) + list(mm[ST_GINFO:ST_GINFO + 4])
        m = [mm[0] & 1, max(mm[4], mm[8]), max(mm[5], mm[9]), (mm[0] >> 1) & 1]
        if m[3]:
            raise ValueError("codes contain NaN")
        ternary = bool(m[0])
        if ternary:
            # rare: some sign is exactly 0 -> pack again, this time with the non-zero bit-plane
            if isinstance(db_codes, PackedCodes) or isinstance(q_codes, PackedCodes):
                raise ValueError("the real-valued side holds exact zeros (ternary keys): it cannot be ranked against "
                                 "packed sign bits -- pass both sides real-valued")
            kw = {} if self.col_sub is None else dict(col_sub=self.col_sub)
            _, q.nz = self.b.pack_sign(q_codes, threshold, flags, True, **kw)
            g.bits, g.nz = self.b.pack_sign(db_codes, threshold, flags, True, **kw)
        label_mode, lw, nclass = L.CH_LAB_NONE, 0, 0
        if q_labels is not None:
            if m[1] <= 1:
                label_mode, nclass = L.CH_LAB_ID, m[2]
            else:
                if q.masks is None or g.masks is None:
                    raise ValueError("multi-hot labels on one side need 2-D labels on the other side too")
                label_mode, lw = L.CH_LAB_MASK, int(q.masks.shape[1])
        return q, g, ternary, label_mode, lw, nclass, rows

    # ------------------------------------------------------------------ one histogram pass
    def _hist(self, q, g, geo, ternary, label_mode, lw, slab_all, slab_rel, thresh=None, emit=L.CH_EMIT_NONE,
              rec=None, key_limit=0):
        threads, nq_pad, nstripes, rps = geo
        lab = lambda p: None if label_mode == L.CH_LAB_NONE else (p.ids if label_mode == L.CH_LAB_ID else p.masks)
        args = dict(
            q_bits=q.bits, q_nz=q.nz, g_bits=g.bits, g_nz=g.nz, q_lab=lab(q), g_lab=lab(g),
            slab_all=slab_all, slab_rel=slab_rel if label_mode != L.CH_LAB_NONE else None, thresh=thresh,
            rec_off=rec["off"] if rec else None, rec_cap=rec["cap"] if rec else None,
            rec_cnt=rec["cnt"] if rec else None, recs=rec["recs"] if rec else None,
            err_flag=rec["err"] if rec else None, nq=q.n, nq_pad=nq_pad, ndb=g.n, nbit=q.nbit, ternary=ternary,
            label_mode=label_mode, mask_words=lw, emit_mode=emit, nstripes=nstripes, threads=threads,
            rows_per_stripe=rps, key_limit=key_limit)
        kind = "hist_select" if thresh is not None else ("hist_count_rec" if emit else "hist_count")
        if thresh is not None:
            self.stats["select_kernel"] = "popc"
        self._timed(kind, q.n * g.n, lambda: self.b.hamming_hist(**args))

    def _gathered_totals(self, slab, nstripes, nbins, nq_pad):
        tot = self.b.empty((nbins, nq_pad), torch.int32)
        self.b.slab_totals(slab, nstripes, nbins, nq_pad, tot)
        return self.comm.all_gather(tot)           # (world, nbins, nq_pad)

    def _offsets(self, cap, geo, nq, thresh, site, nbins_full):
        """Offsets of the per-(stripe, query) slices -> (off, slots to allocate, max threshold | None).

        Without a hint for ``site`` this is a host round trip (exact total, exact max threshold).  With one, the
        arena is sized from the previous evaluation's total (+ 12.5 % + 64 Ki slots), the key range from its largest
        threshold (+ 2), and ``ch_record_offsets_async`` verifies both on the device (ST_PASS bits 2 / 3)."""
        threads, nq_pad, nstripes, rps = geo
        off = self.b.empty((nstripes, nq_pad), torch.int32)
        hint = self._hint["sites"].get(site) if (self._hint is not None and site is not None) else None
        if hint is None or not hasattr(self.b, "record_offsets_async"):
            total, tmax = self.b.record_offsets(cap, nstripes, nq, nq_pad, off, thresh)
            self.host_syncs += 1
            worst = total
            if self.comm.world > 1:
                # every rank must chunk the query set (or not) together: agree on the largest local total
                worst = self._host_ints(self.comm.all_reduce_max(self.b.full((1,), total, torch.int64)))[0]
            if worst >= self.max_slots:
                raise _TooManySlots(worst)
            if site is not None:
                self._new_hint["sites"][site] = (total, tmax)
            self.stats["record_slots"] = total
            return off, max(total, 1), tmax
        ptotal, ptmax = hint
        alloc = min(int(ptotal * 1.125) + 65536, self.max_slots - 1)
        key_limit = 0
        if thresh is not None:
            key_limit = min(int(nbins_full), int(ptmax) + 3)
        w = ST_SITE[site]
        self.b.record_offsets_async(cap, nstripes, nq, nq_pad, off, thresh, alloc, key_limit,
                                    self._status[w:w + 2], self._status[ST_PASS:ST_PASS + 1])
        self._new_hint["sites"][site] = None          # filled in from the status block after the final read
        self.stats["record_slots"] = alloc
        return off, alloc, (key_limit - 1 if thresh is not None else None)

    def _alloc_records(self, cap, geo, nq, thresh=None, site=None, nbins_full=0):
        """Offsets of the per-(stripe, query) record slices and the record buffer.  ST_PASS bit 0 of the status
        block is raised by the pass when a slice overflows.  With ``thresh`` max(thresh) is returned too."""
        threads, nq_pad, nstripes, rps = geo
        off, total, tmax = self._offsets(cap, geo, nq, thresh, site, nbins_full)
        status = self._status
        rec = dict(off=off, cap=cap, cnt=self.b.zeros((nstripes, nq_pad), torch.int32),
                   recs=self.b.empty((total, 4), torch.int32), err=status[ST_PASS:ST_PASS + 1], status=status)
        return (rec, tmax) if thresh is not None else rec

    def _check_records(self, rec):
        if self._host_ints(rec["err"])[0] != 0:
            raise RuntimeError("internal error: record buffer overflow")

    # ------------------------------------------------------------------ tensor-core select pass -> candidate lists
    def _tc_ok(self, q, ternary, nq_pad):
        """the select pass can run on tcgen05: binary or ternary codes of <= 256 bits (int8 operands hold -1, 0, +1
        exactly), whole 128-query tiles"""
        return (self.use_tensor_cores and hasattr(self.b, "hamming_select_tc") and
                nq_pad % 128 == 0 and self.b.tc_code_bytes(q.nbit) > 0)

    def _alloc_cands(self, cap, geo, nq, thresh=None, site=None, nbins_full=0):
        """Candidate-list slices per (stripe, query): offsets from the capacities, row / key arrays.  ST_PASS bit 0
        of the status block is raised by the select pass when a slice overflows.  (``cnt`` needs no zeroing: the
        select kernel writes the count of every (stripe, query < nq) slice, empty stripes included.)"""
        threads, nq_pad, nstripes, rps = geo
        off, total, tmax = self._offsets(cap, geo, nq, thresh, site, nbins_full)
        status = self._status
        cand = dict(off=off, cap=cap, cnt=self.b.empty((nstripes, nq_pad), torch.int32),
                    rows=self.b.empty((total,), torch.int32), key=self.b.empty((total,), torch.int16),
                    err=status[ST_PASS:ST_PASS + 1], status=status)
        return (cand, tmax) if thresh is not None else cand

    def _dense(self, expected_per_query, ndb_total):
        """most 32-row x 32-query chunks hold a candidate: the select kernel skips its max-tree pre-filter"""
        if self.select_dense_override is not None:
            return bool(self.select_dense_override)
        return expected_per_query * 1024.0 > 0.78 * max(ndb_total, 1)

    def _bare(self, dense):
        """sparse select passes drop the threshold block of the contraction (one K block of 2-5: 64-bit codes contract
        K = 64 instead of 96) and compare in the epilogue -- a packed 16-bit max tree per 32-column chunk, the same
        instruction count as the sign-bit AND tree it replaces.  Dense passes keep the block: there every chunk
        would pay 16 packed subtractions, and the ALU pipe is their second bottleneck."""
        return bool(self.epilogue_thresholds and not dense and hasattr(self.b, "tc_code_bytes") and
                    self.b.tc_code_bytes(8, True) > 0)

    def _pair(self, q):
        """Keys up to 128 (binary codes of <= 128 bits, ternary codes of <= 64): ``thresh - key`` fits a signed byte, so
        one 16-bit accumulator value carries the comparisons of TWO gallery rows (bit 7 / bit 15) -- an accumulator
        covers 256 rows and the select kernel has half as many accumulator hand-overs, TMEM loads and barrier round
        trips per pair (``ch_tc_code_bytes_pair`` in the header).  The thresholds always ride in the contraction."""
        return bool(self.paired_rows and hasattr(self.b, "tc_code_bytes_pair") and
                    self.b.tc_code_bytes_pair(q.nbit, q.nz is not None) > 0)

    def _query_plane(self, q, nq_pad, thresh, bare=False, pair=False, scut=None, nstripes=0):
        """int8 query plane: with the thresholds in its threshold slots (made per select pass), or bare; ``scut``: one
        paired plane per stripe (threshold - 1 in the stripes beyond a query's cut)"""
        kw = {} if q.nz is None else dict(nz=q.nz)
        if pair and scut is not None:
            return self._timed("expand_i8", 0, lambda: self.b.expand_i8_query_stripes(q.bits, q.nbit, nq_pad, thresh, scut,
                                                                                      nstripes, q.n, **kw))
        if pair:
            return self._timed("expand_i8", 0, lambda: self.b.expand_i8(q.bits, q.nbit, nq_pad, thresh=thresh, nq=q.n,
                                                                        pair=True, **kw))
        if bare:
            if q.i8b is None:
                q.i8b = self._timed("expand_i8", 0, lambda: self.b.expand_i8(q.bits, q.nbit, nq_pad, bare=True,
                                                                              query=True, **kw))
            return q.i8b
        return self._timed("expand_i8", 0, lambda: self.b.expand_i8(q.bits, q.nbit, nq_pad, thresh=thresh, nq=q.n, **kw))

    def _gallery_plane(self, p, bare=False, pair=False):
        kw = {} if p.nz is None else dict(nz=p.nz)
        if pair:
            if p.i8p is None:
                p.i8p = self._timed("expand_i8", 0, lambda: self.b.expand_i8(p.bits, p.nbit, pair=True, **kw))
            return p.i8p
        if bare:
            if p.i8b is None:
                p.i8b = self._timed("expand_i8", 0, lambda: self.b.expand_i8(p.bits, p.nbit, bare=True, **kw))
            return p.i8b
        if p.i8 is None:
            p.i8 = self._timed("expand_i8", 0, lambda: self.b.expand_i8(p.bits, p.nbit, **kw))
        return p.i8

    def _select_tc(self, q, g, geo, thresh, cand, dense, kind="hist_select_tc", ndb=None, nstripes=None, rps=None,
                   bad=None, scut=None):
        """the select pass on the tensor cores over the packed shard ``g`` (or a row sample of it); ``bad`` (nq_pad):
        marked for every query one of whose slices overflows"""
        threads, nq_pad, nstripes_g, rps_g = geo
        pair = self._pair(q)
        bare = not pair and self._bare(dense)
        q_i8 = self._query_plane(q, nq_pad, thresh, bare, pair, scut if pair else None,
                                 nstripes_g if nstripes is None else nstripes)
        g_i8 = self._gallery_plane(g, bare, pair)
        kw = dict(thresh=thresh, ternary=q.nz is not None) if bare else {}
        if pair:
            kw = dict(pair=True, ternary=q.nz is not None)
        if bad is not None:
            kw["bad"] = bad
        self._timed(kind, q.n * g.n, lambda: self.b.hamming_select_tc(
            q_i8=q_i8, g_i8=g_i8, cand=cand, nq=q.n, nq_pad=nq_pad, ndb=g.n, nbit=q.nbit,
            nstripes=nstripes_g if nstripes is None else nstripes, rows_per_stripe=rps_g if rps is None else rps,
            dense=dense, **kw))
        if kind == "hist_select_tc":
            self.stats["select_kernel"] = "tcgen05"
            self.stats["select_dense"] = bool(dense)
            self.stats["select_threshold"] = "epilogue" if bare else "contraction"
            self.stats["select_rows_per_cell"] = 2 if pair else 1

    def _cand_hist(self, c, cand, nbins, tot, stripe0=0, nstripes=None, rank=None):
        """keys + label matches of the candidates of a block of stripes, accumulated into ``tot`` (2, nbins, nq_pad);
        ``rank`` (single rank, whole list): the fused kernel instead -- histogram, bases, verification and the walk
        in one launch (``ch_cand_rank``), ``rank`` = its cols / r_eff / pr_k / rmax / need / status / bad"""
        b, q, g = self.b, c["q"], c["g"]
        threads, nq_pad, nstripes_all, rps = c["geo"]
        label_mode, lw = c["label_mode"], c["lw"]
        lab = lambda p: None if label_mode == L.CH_LAB_NONE else (p.ids if label_mode == L.CH_LAB_ID else p.masks)
        kw = {}
        if q.nz is not None:
            kw = dict(q_nz=q.nz, g_nz=g.nz)            # ternary keys (doubled scale) from both planes
        elif label_mode == L.CH_LAB_ID and hasattr(b, "gather_plane"):
            # single-label gallery: [code | class id] rows, one memory sector per candidate instead of two gathers
            if nstripes is None:
                if g.plane is None:
                    g.plane = b.gather_plane(g.bits, g.ids, g.nbit)
            else:                                            # a row block of a streamed gallery
                if g.plane is None:
                    g.plane = b.empty((g.bits.shape[0], b.gather_plane_words(g.nbit)), torch.int32)
                r0 = stripe0 * rps
                r1 = min(g.bits.shape[0], (stripe0 + nstripes) * rps)
                b.gather_plane(g.bits[r0:r1], g.ids[r0:r1], g.nbit, out=g.plane[r0:r1])
            kw = dict(g_plane=g.plane)
        if rank is not None:
            self._timed("cand_rank", 0, lambda: b.cand_rank(
                cand, q_bits=q.bits, g_bits=g.bits, q_lab=lab(q), g_lab=lab(g), label_mode=label_mode, mask_words=lw,
                nq=c["nq"], nq_pad=nq_pad, nstripes=nstripes_all, nbins=nbins, nbit=q.nbit, **rank, **kw))
            return
        self._timed("cand_hist", 0, lambda: b.cand_hist(
            cand, q_bits=q.bits, g_bits=g.bits, q_lab=lab(q), g_lab=lab(g), label_mode=label_mode, mask_words=lw,
            tot_all=tot[0], tot_rel=tot[1] if label_mode != L.CH_LAB_NONE else None, nq=c["nq"], nq_pad=nq_pad,
            nstripes=nstripes_all if nstripes is None else nstripes, nbins=nbins, nbit=q.nbit, stripe0=stripe0, **kw))

    def _cand_bases(self, c, cand, nbins, need=None, tot=None, rmax=None, bad=None):
        """keys + label matches of the candidates -> per-rank key totals -> all-gather -> bases (+ verification
        that every query has >= ``need`` candidates: ST_SHORT).  ``tot``: totals already accumulated per block."""
        b, comm = self.b, self.comm
        threads, nq_pad, nstripes, rps = c["geo"]
        nq = c["nq"]
        labelled = c["label_mode"] != L.CH_LAB_NONE
        if tot is None:
            tot = b.zeros((2, nbins, nq_pad), torch.int32)
            self._cand_hist(c, cand, nbins, tot)
        if self.debug_counts:                                        # (dev tools only: a host round trip)
            self.stats["candidates"] = int(cand["cnt"][:, :nq].sum(dtype=torch.int64))
        tot = comm.all_gather(tot)                                   # (world, 2, nbins, nq_pad)
        base0_all = b.empty((nbins, nq_pad), torch.int32)
        base0_rel = b.empty((nbins, nq_pad), torch.int32) if labelled else None
        # key_max[q] = smallest key at which the (global) list holds rmax items: larger keys cannot rank below rmax
        key_max = b.empty((nq_pad,), torch.int32) if rmax is not None else None
        b.scan_bases_pair(tot, comm.world, comm.rank, nbins, nq, nq_pad, -1 if rmax is None else rmax,
                          0 if need is None else need, base0_all, base0_rel, key_max, None,
                          self._status[ST_SHORT:ST_SHORT + 1], **({} if bad is None else dict(bad=bad)))
        return base0_all, base0_rel, key_max

    # ------------------------------------------------------------------ the evaluation
    def evaluate(self, db_codes, db_labels, q_codes, q_labels, R, threshold=0.0, PRs=(),
                 remove_first_retrieved=False, return_ap=False, zero_mean=False, _raw=False):
        """Returns ``(mAPs list, recalls list, precisions list[, ap (nR, nq) tensor])``.

        ``db_codes`` / ``db_labels`` are THIS rank's contiguous gallery row block; queries are replicated.

        A repeated evaluation of a known shape runs speculatively (see ``__init__``); a query set whose candidate /
        record slices would need >= 2^32 slots is evaluated in query chunks (AP, hit counts are per-query
        quantities: the chunk means combine by a weighted mean)."""
        b, comm = self.b, self.comm
        if hasattr(b, "begin"):
            b.begin()
        self.host_syncs = 0
        self._want_raw = bool(_raw)        # internal (repair runs): return the per-query sums instead of the means
        r_list = [int(r) for r in R]
        pr_k = [int(k) for k in PRs]
        if len(r_list) == 0 and len(pr_k) == 0:
            return [], [], []
        if any(r == 0 or r < -1 for r in r_list) or any(k <= 0 for k in pr_k):
            raise ValueError("R must be -1 or positive; PRs must be positive")
        args = (db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k, remove_first_retrieved, return_ap,
                zero_mean)
        key = self._hint_key(*args)
        hint = self._hints.get(key) if self.speculate else None
        graphable = self._graphable(args, hint, _raw)
        ge = self._graphs.get(key) if graphable else None
        replay = ge is not None and ge.get("graph") is not None
        if replay and ge.get("ptrs") is not None and ge["ptrs"] != tuple(t.data_ptr() for t in args[:4]):
            replay, ge = False, None         # captured in place, for other buffers: ordinary path, graph kept
        if comm.world > 1 and self.speculate:
            # every rank must take the same road (the collectives differ): speculate only if ALL ranks hold a hint,
            # replay a captured graph only if ALL ranks hold one.  (The stream is idle here -- the previous evaluation
            # ended with a sync -- so this round trip is cheap.)
            have = b.full((1,), 2 if replay else (0 if hint is None else 1), torch.int32)
            level = -self._host_ints(comm.all_reduce_max(-have))[0]
            if level < 1:
                hint = None
            replay = replay and level == 2
            graphable = graphable and hint is not None
        if replay:
            out = self._replay(ge, args)
            if out is not None:
                return out
            ge["graph"] = None               # contradicted: the ordinary path decides (and may re-capture)
        try:
            out = self._evaluate(*args, hint=hint)
            if out is _RETRY:
                # a stale hint (the data changed under the same shape): once more, asking the device
                self._hints.pop(key, None)
                out = self._evaluate(*args, hint=None)
                self.stats["speculation"] = "retried"
            else:
                self.stats["speculation"] = "hit" if hint is not None else "none"
        except _TooManySlots as e:
            self._hints.pop(key, None)
            return self._evaluate_chunked(args, int(e.args[0]))
        if self.speculate and self._new_hint is not None and self._new_hint.get("complete"):
            self._hints[key] = self._new_hint
        self.stats["host_syncs"] = self.host_syncs
        if graphable and self.stats.get("speculation") == "hit":
            self._capture(key, args)
        return out

    # ------------------------------------------------------------------ CUDA graphs of small evaluations
    def release_graphs(self):
        """drops every captured graph (and its buffers); required before ``destroy_process_group`` in group mode"""
        self._graphs.clear()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def _graphable(self, args, hint, raw):
        db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k, rf, return_ap, zero_mean = args
        if not (self.use_graphs and self.speculate and hint is not None and not raw
                and (self.comm.world == 1 or self.graph_multi_gpu)
                and not return_ap and not zero_mean and not self.profile and hasattr(self.b, "capture_results")):
            return False
        for t in (db_codes, db_labels, q_codes, q_labels):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
                return False
        return True

    def _graph_in_place(self, args):
        """inputs too large to copy into graph-owned buffers are captured IN PLACE: such a graph is replayed only for
        calls that pass tensors at the very same addresses (a re-evaluation of the same buffers, or buffers the
        caching allocator handed out again) -- the memory read is always that of the current call's tensors"""
        return sum(t.numel() * t.element_size() for t in args[:4]) > self.graph_max_bytes

    def _capture(self, key, args):
        """Right after a successful speculative evaluation: the same evaluation once more, on copies of the inputs,
        recorded instead of executed.  The host-side decisions of the recording run are made on the results of the
        evaluation that has just finished -- same data, same hint: the same road."""
        ge = self._graphs.setdefault(key, dict(graph=None, tries=0))
        st = self.stats
        if ge["tries"] >= 2 or st.get("query_chunks") or (st.get("sample") or {}).get("fallback") or \
                (st.get("sample") or {}).get("repaired_queries"):
            return
        ge["tries"] += 1
        b = self.b
        db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k, rf, return_ap, zero_mean = args
        in_place = self._graph_in_place(args)
        if in_place and not self.graph_in_place:
            return
        if ge.get("graph") is not None:
            return                           # (an in-place graph of other buffers exists: keep it)
        static = list(args[:4]) if in_place else [t.clone() for t in (db_codes, db_labels, q_codes, q_labels)]
        saved = (dict(self.stats), self.host_syncs, self._hint, self._new_hint)
        graph = torch.cuda.CUDAGraph()
        ok = False
        l0 = b.launch_count()
        try:
            torch.cuda.synchronize()
            b.capture_results = b.last_results
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                b.begin()
                out = self._evaluate(static[0], static[1], static[2], static[3], r_list, threshold, pr_k, rf,
                                     return_ap, zero_mean, hint=self._hints.get(key))
            ok = out is not _RETRY and out is not None
        except Exception as e:                  # (a host round trip inside the sequence, an allocation the capture
            self.stats_graph_error = repr(e)    #  cannot hold, ...: this shape simply keeps the ordinary path)
            ge["tries"] = 99
        finally:
            b.capture_results = None
            b.begin()
        launches = b.launch_count() - l0
        stats = dict(self.stats)
        self.stats, self.host_syncs, self._hint, self._new_hint = saved
        if ok:
            ge.update(graph=graph, static=None if in_place else static, shape=b.last_result_shape, stats=stats,
                      launches=launches, hint=self._hints.get(key),
                      ptrs=tuple(t.data_ptr() for t in args[:4]) if in_place else None)
        else:
            torch.cuda.synchronize()

    def _replay(self, ge, args):
        b = self.b
        if ge["static"] is not None:
            for s, t in zip(ge["static"], args[:4]):
                s.copy_(t)
        ge["graph"].replay()
        b.last_result_shape = ge["shape"]
        maps, recalls, precisions, flags = b.fetch_results()
        self.host_syncs = 1 if self.comm.world == 1 else 2
        self._hint = ge["hint"]
        # what a speculative run checks after its one round trip: a contradicted hint, an overflow, a short list
        if self._hint is None or self._stale(flags) or flags[ST_PASS] or flags[ST_SHORT]:
            return None
        self.graph_launches += ge["launches"]
        self.stats = dict(ge["stats"])
        self.stats.update(speculation="graph", host_syncs=self.host_syncs)
        return maps, recalls, precisions

    def _hint_key(self, db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k, rf, return_ap, zero_mean):
        def sig(t):
            if t is None:
                return None
            dev = getattr(t, "device", None)
            return (type(t).__name__, tuple(t.shape), str(getattr(t, "dtype", "")), getattr(dev, "type", None),
                    bool(getattr(t, "is_pinned", lambda: False)()) if getattr(dev, "type", None) == "cpu" else False)
        knobs = (self.sample_stride, self.sample_two_level, self.sample2_sub, self.sample2_min_rows,
                 self.sample2_min_work, self.sample_min_rows, self.sample_min_ratio, self.stream_host_gallery,
                 self.stream_min_rows, self.stream_chunks, self.use_tensor_cores, self.select_dense_override,
                 self.stripe_rows_override, self.epilogue_thresholds, self.max_slots, self.paired_rows,
                 self.stripe_cut, self.fused_rank, self.stream_native_loader, self.stream_chunks_native,
                 self.stream_cand_overlap, self.stream_late_labels, self.stream_fused_rank, self.stream_loader_labels)
        return (sig(db_codes), sig(db_labels), sig(q_codes), sig(q_labels), tuple(r_list), float(threshold),
                tuple(pr_k), bool(rf), bool(zero_mean), self.comm.world, knobs)

    def _evaluate_chunked(self, args, total_slots):
        """The query set in chunks (each small enough for 32-bit slot offsets), combined exactly: every output is a
        mean over queries of a per-query quantity."""
        db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k, rf, return_ap, zero_mean = args
        nq = int(q_codes.shape[0])
        nchunks = min(nq, max(2, int(total_slots // max(1, (self.max_slots * 3) // 8)) + 1))
        if nq < 2:
            raise RuntimeError(f"{total_slots} candidate slots for a single query exceed the 32-bit slot index")
        bounds = [nq * i // nchunks for i in range(nchunks + 1)]
        maps, recalls, precisions = [0.0] * len(r_list), [0.0] * len(pr_k), [0.0] * len(pr_k)
        aps, syncs = [], 0
        for i in range(nchunks):
            lo, hi = bounds[i], bounds[i + 1]
            if hi == lo:
                continue
            out = self.evaluate(db_codes, db_labels, q_codes[lo:hi], q_labels[lo:hi], r_list, threshold, pr_k, rf,
                                return_ap, zero_mean)
            syncs += self.stats.get("host_syncs", 0)
            wgt = (hi - lo) / nq
            maps = [a + wgt * v for a, v in zip(maps, out[0])]
            recalls = [a + wgt * v for a, v in zip(recalls, out[1])]
            precisions = [a + wgt * v for a, v in zip(precisions, out[2])]
            if return_ap:
                aps.append(out[3])
        self.stats.update(query_chunks=nchunks, host_syncs=syncs)
        if return_ap:
            return maps, recalls, precisions, torch.cat(aps, dim=1)
        return maps, recalls, precisions

    def _loader_bits(self, g):
        """The gallery is not streamed after all: waits for the native loader, hands its bits to ``g`` (the current
        stream waits for the copies) and returns the flag bits (1: a zero sign, 2: NaN)."""
        loader, g.loader = g.loader, None
        fl = loader.join()[loader.idx["g"]]
        torch.cuda.current_stream().wait_stream(self._side_stream)
        g.bits, g.nz = loader.bits[loader.idx["g"]], None
        return fl

    def _evaluate(self, *args, **kw):
        try:
            return self._evaluate_body(*args, **kw)
        finally:
            # the loader thread reads the caller's memory and writes into the shard's arrays: nothing may outlive it
            loader, self._loader = self._loader, None
            if loader is not None:
                loader.join()

    def _evaluate_body(self, db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k, remove_first_retrieved,
                       return_ap, zero_mean, hint=None):
        b, comm = self.b, self.comm
        self._hint = hint
        self._new_hint = dict(sites={}, complete=False)
        self.stats.pop("query_chunks", None)
        # (one rank, host gallery: the row sample the top-R path will want is host-known up front -- the native loader
        # packs it right behind the queries; _pass_topr_sampled checks that it is the sample it needs)
        pre_stride = 0
        if comm.world == 1 and isinstance(db_codes, torch.Tensor) and db_codes.dim() == 2 and not db_codes.is_cuda:
            n0, rf0 = int(db_codes.shape[0]), (1 if remove_first_retrieved else 0)
            ll = max(n0 - rf0, 0)
            rmax0 = max([ll if r == -1 else min(r, ll) for r in r_list] + [min(k, ll) for k in pr_k] + [0])
            if (rmax0 * 4 < ll and self.sample_stride > 1 and n0 >= self.sample_min_rows and
                    rmax0 * self.sample_min_ratio <= n0):
                pre_stride = self.sample_stride * (2 if (rmax0 + rf0) * 5000 <= n0 else 1)
        q, g, ternary, label_mode, lw, nclass, rows = self._prepare(db_codes, db_labels, q_codes, q_labels, threshold,
                                                                    allow_defer=True, zero_mean=zero_mean,
                                                                    sample_stride=pre_stride)
        db_codes = self._db_codes_eff           # (the device copy when zero_mean moved a host gallery)
        if label_mode == L.CH_LAB_NONE:
            raise ValueError("labels are required")
        nq, nbit = q.n, q.nbit
        nbins = (2 * nbit if ternary else nbit) + 1
        ndb_total = sum(rows)
        rf = 1 if remove_first_retrieved else 0
        list_len = max(ndb_total - rf, 0)
        r_eff = [list_len if r == -1 else min(r, list_len) for r in r_list]
        rmax = max(r_eff + [min(k, list_len) for k in pr_k] + [0])
        full_ranking = rmax * 4 >= list_len          # top-R close to "all": one pass is cheaper than two
        cls_ok = label_mode == L.CH_LAB_ID and 0 < nclass <= (1 << 20)
        sampled = (not full_ranking and self.sample_stride > 1 and cls_ok and
                   ndb_total >= self.sample_min_rows and rmax * self.sample_min_ratio <= ndb_total)
        # sparser sample when the candidates are a tiny fraction of the gallery (a looser threshold then costs
        # little in the select pass and the sample pass shrinks)
        stride = self.sample_stride
        if sampled and stride > 1 and (rmax + rf) * 5000 <= ndb_total:
            stride *= 2
        geo = b.geometry(nq, g.n, nbit, ternary, label_mode, lw)
        # host-resident gallery on the sampled top-R path: it is streamed in row blocks behind the select pass
        streamed = (g.bits is None and sampled and not ternary and label_mode == L.CH_LAB_ID and geo[1] % 128 == 0
                    and self.use_tensor_cores)
        min_stripes = 0
        if streamed:
            min_stripes = self._stream_per(geo[1]) * 2 * (self.stream_chunks_native if g.loader is not None
                                                          else self.stream_chunks)
        geo = self._agree_geometry(geo, g.n, stride if sampled else 1, min_stripes)
        if (not streamed and not full_ranking and not self.stripe_rows_override and hasattr(b, "sm_count")
                and self._tc_ok(q, ternary, geo[1])):
            # the select pass will run on the tensor cores: stripes that fill its waves (one CTA per SM)
            geo = self._tc_geometry(geo, g.n, stride if sampled else 1, 0.75 if sampled else 0.1)
        threads, nq_pad, nstripes, rps = geo
        self.stats.pop("sample2", None)
        self.stats.update(dict(ternary=ternary, label_mode=label_mode, geometry=geo, nbins=nbins,
                               ndb_total=ndb_total, world=comm.world))
        ctx = dict(q=q, g=g, geo=geo, ternary=ternary, label_mode=label_mode, lw=lw, nclass=nclass, nq=nq,
                   nbins=nbins, rmax=rmax, rf=rf, pr_k=pr_k, ndb_total=ndb_total, stride=stride, rows=rows)
        ctx.update(r_eff=r_eff, r_list=r_list, return_ap=return_ap, db_codes=db_codes, threshold=threshold)
        streamed = streamed and nstripes >= 2
        if not streamed:
            self._ensure_labels(q, g)
        res = None
        if streamed:
            self.stats["mode"] = "topR-sampled-streamed"
            try:
                res = self._finish(ctx, self._pass_topr_sampled(ctx, streamed=True))
            finally:
                streamer = ctx.pop("streamer", None)
                if streamer is not None:
                    streamer.join()
            if hint is not None and self._stale(res[4]):
                return _RETRY
            if res[4][ST_PASS] or res[4][ST_SHORT]:
                # (a zero / NaN found by the native loader is not a per-query failure: no repair, the whole evaluation
                # is redone with ternary keys -- or raises)
                odd = comm.world == 1 and g.loader is not None and any(g.loader.join())
                res = self._repair(ctx, res, q_codes, q_labels, db_labels) if res[4][ST_CODES] == 0 and not odd else None
            if res is None:
                self.stats["sample"].update(fallback=True)
                sampled = False                     # the same sample would fail again: go exact
                self._status[:ST_CODES].zero_()
        if g.bits is None or res is None and streamed:
            # gallery codes were deferred but the streamed path is not applicable (or gave up): pack them now
            if g.loader is not None:
                # (the native loader has packed -- or is packing -- every row: its bits are the shard's bits)
                fl = self._loader_bits(g)
                if comm.world > 1:
                    flt = b.zeros((1,), torch.int32)
                    flt.fill_(fl)
                    fl = self._host_ints(comm.all_reduce_max(flt))[0]
            else:
                fl = b.zeros((1,), torch.int32)
                self._pack_codes(g, db_codes, threshold, fl)
                fl = self._host_ints(comm.all_reduce_max(fl) if comm.world > 1 else fl)[0]
            if fl & 2:
                raise ValueError("codes contain NaN")
            if fl & 1:
                # zeros in the gallery: ternary keys after all -> start over on the regular path
                saved, self.stream_host_gallery = self.stream_host_gallery, False
                try:
                    return self.evaluate(db_codes, db_labels, q_codes, q_labels, r_list, threshold, pr_k,
                                         remove_first_retrieved, return_ap, zero_mean)
                finally:
                    self.stream_host_gallery = saved
        if res is not None:
            pass
        elif full_ranking:
            self.stats["mode"] = "all"
            res = self._finish(ctx, self._pass_all(ctx))
        else:
            if sampled:
                self.stats["mode"] = "topR-sampled"
                res = self._finish(ctx, self._pass_topr_sampled(ctx))
                if hint is not None and self._stale(res[4]):
                    return _RETRY
                if res[4][ST_PASS] or res[4][ST_SHORT]:
                    # a slice overflowed or some query has fewer than R candidates under the sampled threshold:
                    # re-rank those queries by the exact path, or -- if they are many -- the whole evaluation
                    res = self._repair(ctx, res, q_codes, q_labels, db_labels)
                    if res is None:
                        self.stats["sample"].update(fallback=True)
                        self._status[:ST_CODES].zero_()
            if res is None:
                self.stats["mode"] = "topR"
                res = self._finish(ctx, self._pass_topr_exact(ctx))
        maps, recalls, precisions, ap, flags, raw = res
        if hint is not None and self._stale(flags):
            return _RETRY
        if flags[ST_PASS]:
            raise RuntimeError("internal error: record buffer overflow")
        self._close_hint(flags)
        if self._want_raw:
            return raw
        if return_ap:
            return maps, recalls, precisions, ap
        return maps, recalls, precisions

    def _repair(self, c, res, q_codes, q_labels, db_labels):
        """The sampled pass failed for SOME queries (marked in ``bad``: a list slice overflowed its Poisson bound, or
        the list is shorter than R).  Failures are per query -- every other query's list is complete and ranked --
        so only the marked queries are re-ranked, by the exact two-pass path of a child evaluator over the already
        packed codes, and their rows of the per-query sums are replaced before the means are taken again.  Returns
        the repaired result tuple, or None when too many queries failed (then the whole evaluation is redone)."""
        maps, recalls, precisions, ap, flags, raw = res
        b, comm = self.b, self.comm
        bad = raw.get("bad") if raw is not None else None
        q, g, nq = c["q"], c["g"], c["nq"]
        if bad is None or g.bits is None or isinstance(q_labels, PackedCodes):
            return None
        if comm.world > 1:
            bad = comm.all_reduce_max(bad)
        idx = torch.nonzero(bad[:nq]).flatten()
        self.host_syncs += 1
        n_bad = int(idx.numel())
        self.stats["sample"].update(overflow=flags[ST_PASS], short=flags[ST_SHORT], repaired_queries=n_bad)
        if n_bad == 0 or n_bad > max(64, nq // 50):
            return None
        # inputs of the repair run: the packed planes (binary codes: bits are final -- threshold and column mean are
        # already applied), the labels of the marked queries
        if c["ternary"]:
            return None
        dbc = PackedCodes(g.bits[:g.n], g.nbit)
        qc = PackedCodes(q.bits[idx].contiguous(), q.nbit)
        sel = idx.to(q_labels.device) if isinstance(q_labels, torch.Tensor) else idx.cpu().numpy()
        child = Evaluator(b._b if isinstance(b, _TimedBackend) else b, comm)
        child.speculate, child.sample_stride = False, 0           # exact two-pass path, no hints
        child.use_tensor_cores, child.stripe_rows_override = self.use_tensor_cores, self.stripe_rows_override
        sub = child.evaluate(dbc, db_labels, qc, q_labels[sel], [r for r in c["r_list"]], 0.0, c["pr_k"], bool(c["rf"]),
                             _raw=True)
        self.host_syncs += child.host_syncs
        raw["cols"][idx] = sub["cols"]
        if raw["total_rel"] is not None and sub["total_rel"] is not None:
            raw["total_rel"][idx] = sub["total_rel"][:n_bad]
        if raw["first_rel"] is not None and sub["first_rel"] is not None:
            raw["first_rel"][idx] = sub["first_rel"][:n_bad]
        status = raw["status"]
        status[:ST_CODES].zero_()                                  # the failure is repaired: clear its flags
        self.host_syncs += 1
        maps, recalls, precisions, flags = b.reduce_means(raw["cols"], raw["total_rel"], raw["first_rel"], nq,
                                                          len(c["r_eff"]), c["pr_k"], ap, status)
        return maps, recalls, precisions, ap, flags, raw

    def _stale(self, flags):
        """did the status block of a speculative run contradict its hint?  (flags = the block, MAX over ranks)"""
        h = self._hint
        if flags[ST_PASS] & 12:
            return True                                  # arena too small / threshold beyond the key range
        if flags[ST_CODES] & 2:
            raise ValueError("codes contain NaN")
        hm = h["mm"]
        if (flags[ST_CODES] & 1) and not (hm[ST_CODES] & 1):
            return True                                  # zeros showed up: ternary keys
        for a, bq in ((ST_QINFO, ST_GINFO),):
            if max(flags[a], flags[bq]) != max(hm[a], hm[bq]) or max(flags[a + 1], flags[bq + 1]) != max(hm[a + 1], hm[bq + 1]):
                return True                              # label form / class count changed
        if self.comm.world > 1 and (list(flags[ST_ROWS:]) != list(h["rows"]) or
                                     bool(flags[ST_PACKED]) != bool(h["any_packed"])):
            return True
        return False

    def _close_hint(self, flags):
        """the totals the device reported for the speculative allocations become the next hint"""
        nh = self._new_hint
        for site, v in list(nh["sites"].items()):
            if v is None:
                w = ST_SITE[site]
                nh["sites"][site] = (int(flags[w]), int(flags[w + 1]))
        nh["complete"] = True

    def _finish(self, c, st):
        """Records -> per-query sums (K4) -> all-reduce -> means.  The status words of the passes come back
        with the results in the one final host sync."""
        b, comm = self.b, self.comm
        threads, nq_pad, nstripes, rps = c["geo"]
        nq, rf, r_eff, pr_k = c["nq"], c["rf"], c["r_eff"], c["pr_k"]
        ncols = 2 * len(r_eff) + len(pr_k)
        cols = b.empty((nq, max(ncols, 1)), torch.float64)       # (every entry is written by the K4 kernels)
        if "cand" in st:
            # candidate lists (tensor-core select pass): ranks straight from the lists
            cand, nbins = st["cand"], st["nbins"]
            first_rel = None
            if st.get("fused"):
                f = st["fused"]
                self._cand_hist(c, cand, nbins, None, rank=dict(
                    cols=cols, r_eff=r_eff, pr_k=pr_k, rmax=f["rmax"], need=f["need"],
                    status=self._status[ST_SHORT:ST_SHORT + 1], bad=f["bad"]))
            else:
                kw = dict(base0_all=st["base0_all"], base0_rel=st["base0_rel"], nq=nq, nq_pad=nq_pad, nstripes=nstripes,
                          nbins=nbins, remove_first=bool(rf), key_max=st.get("key_max"))
                if rf:
                    first_rel = b.zeros((nq_pad,), torch.int32)
                    b.cand_finalize(cand, mode=1, first_rel_out=first_rel, **kw)
                    first_rel = comm.all_reduce_max(first_rel)
                self._timed("cand_finalize", 0, lambda: b.cand_finalize(cand, mode=0, first_rel=first_rel, cols=cols,
                                                                        r_eff=r_eff, pr_k=pr_k, **kw))
            cols = comm.all_reduce_sum(cols)
            status = cand["status"]
            if comm.world > 1:
                status = comm.all_reduce_max(status)
            ap = b.empty((len(r_eff), nq), torch.float64) if c["return_ap"] else None
            self.host_syncs += 1
            maps, recalls, precisions, flags = b.reduce_means(cols, st["total_rel"] if pr_k else None, first_rel, nq,
                                                              len(r_eff), pr_k, ap, status)
            raw = dict(cols=cols, total_rel=st["total_rel"] if pr_k else None, first_rel=first_rel, bad=st.get("bad"),
                       status=status)
            return maps, recalls, precisions, ap, flags, raw
        rec = st["rec"]
        f = dict(recs=rec["recs"], rec_off=rec["off"], rec_cnt=rec["cnt"], base0_all=st["base0_all"],
                 base0_rel=st["base0_rel"], sbase_all=st["sbase_all"], sbase_rel=st["sbase_rel"], first_rel=None,
                 partial=b.empty((nstripes, nq_pad, max(ncols, 1)), torch.float64), cols=cols,
                 nq=nq, nq_pad=nq_pad, nstripes=nstripes, nbins=st.get("nbins", c["nbins"]), remove_first=bool(rf),
                 r_eff=r_eff, pr_k=pr_k)
        first_rel = None
        if rf:
            first_rel = b.zeros((nq_pad,), torch.int32)
            b.first_relevant(f, first_rel)
            first_rel = comm.all_reduce_max(first_rel)
            f["first_rel"] = first_rel
        b.finalize_records(f)
        cols = comm.all_reduce_sum(cols)
        status = rec["status"]
        if comm.world > 1:
            status = comm.all_reduce_max(status)
        ap = b.empty((len(r_eff), nq), torch.float64) if c["return_ap"] else None
        self.host_syncs += 1
        maps, recalls, precisions, flags = b.reduce_means(cols, st["total_rel"] if pr_k else None, first_rel, nq,
                                                          len(r_eff), pr_k, ap, status)
        raw = dict(cols=cols, total_rel=st["total_rel"] if pr_k else None, first_rel=first_rel, bad=None, status=status)
        return maps, recalls, precisions, ap, flags, raw

    def _class_counts(self, ctx):
        """(nstripes, nclass) per-stripe class histogram of the single-label gallery shard."""
        threads, nq_pad, nstripes, rps = ctx["geo"]
        cls = self.b.zeros((nstripes, ctx["nclass"]), torch.int32)
        self.b.class_counts(ctx["g"].ids, ctx["g"].n, rps, ctx["nclass"], cls)
        return cls

    def _pass_all(self, c):
        """Full ranking: ONE pass counts every pair and records every relevant pair."""
        b, comm, q, g, geo = self.b, self.comm, c["q"], c["g"], c["geo"]
        threads, nq_pad, nstripes, rps = geo
        nbins, nq, label_mode, lw, ternary = c["nbins"], c["nq"], c["label_mode"], c["lw"], c["ternary"]
        slabs = b.zeros((2, nstripes, nbins, nq_pad), torch.int32)        # {all, relevant}: one fill, one scan
        slab_all, slab_rel = slabs[0], slabs[1]
        cap = b.empty((nstripes, nq_pad), torch.int32)
        if label_mode == L.CH_LAB_ID and 0 < c["nclass"] * nstripes <= (1 << 26):
            b.record_caps(2, self._class_counts(c), q.ids, nstripes, c["nclass"], nq, nq_pad, False, cap)
        else:
            # capacities from a counting pass (multi-hot labels, or too many classes for the table)
            self._hist(q, g, geo, ternary, label_mode, lw, slab_all, slab_rel)
            b.record_caps(1, slab_rel, None, nstripes, nbins, nq, nq_pad, False, cap)
            slabs.zero_()
        rec = self._alloc_records(cap, geo, nq, site="all")
        self._hist(q, g, geo, ternary, label_mode, lw, slab_all, slab_rel, emit=L.CH_EMIT_RELEVANT, rec=rec)
        tot = comm.all_gather(b.slab_scan(slabs, nstripes, nbins, nq_pad))      # (world, 2, nbins, nq_pad)
        base0_all = b.empty((nbins, nq_pad), torch.int32)
        base0_rel = b.empty((nbins, nq_pad), torch.int32)
        total_rel = b.empty((nq_pad,), torch.int32)
        b.scan_bases_pair(tot, comm.world, comm.rank, nbins, nq, nq_pad, -1, 0, base0_all, base0_rel, None, total_rel,
                          None)
        return dict(rec=rec, base0_all=base0_all, base0_rel=base0_rel, sbase_all=slab_all, sbase_rel=slab_rel,
                    total_rel=total_rel)

    def _pass_topr_exact(self, c):
        """Exact two-pass top-R: pass 1 counts every pair (-> per-query threshold key), pass 2 re-streams
        and only counts / matches / records the pairs with key <= threshold."""
        b, comm, q, g, geo = self.b, self.comm, c["q"], c["g"], c["geo"]
        threads, nq_pad, nstripes, rps = geo
        nbins, nq, label_mode, lw, ternary = c["nbins"], c["nq"], c["label_mode"], c["lw"], c["ternary"]
        slab_all = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        slab_rel = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        base0_all = b.empty((nbins, nq_pad), torch.int32)
        base0_rel = b.empty((nbins, nq_pad), torch.int32)
        total_rel = b.zeros((nq_pad,), torch.int32)
        need_total_rel = len(c["pr_k"]) > 0
        lm1 = label_mode if need_total_rel else L.CH_LAB_NONE
        self._hist(q, g, geo, ternary, lm1, lw, slab_all, slab_rel)
        thresh = b.empty((nq_pad,), torch.int32)
        tot_a = comm.all_gather(self._local_totals(slab_all, geo, nbins))
        b.scan_bases(tot_a, comm.world, comm.rank, nbins, nq, nq_pad, c["rmax"] + c["rf"], base0_all, thresh, None)
        if need_total_rel:
            tot_r = comm.all_gather(self._local_totals(slab_rel, geo, nbins))
            b.scan_bases(tot_r, comm.world, comm.rank, nbins, nq, nq_pad, -1, base0_rel, None, total_rel)
        cap = b.empty((nstripes, nq_pad), torch.int32)
        b.record_caps(0, slab_all, thresh, nstripes, nbins, nq, nq_pad, False, cap)
        if self._tc_ok(q, ternary, nq_pad):
            # pass 2 on the tensor cores: candidate lists (exact capacities), ranks from the lists
            del slab_rel
            cand = self._alloc_cands(cap, geo, nq, site="exact")
            self._select_tc(q, g, geo, thresh, cand, self._dense(1.3 * (c["rmax"] + c["rf"]), c["ndb_total"]))
            base0_all, base0_rel, key_max = self._cand_bases(c, cand, nbins, rmax=c["rmax"] + c["rf"])
            return dict(cand=cand, base0_all=base0_all, base0_rel=base0_rel, total_rel=total_rel, nbins=nbins,
                        key_max=key_max)
        if label_mode == L.CH_LAB_ID and 0 < c["nclass"] * nstripes <= (1 << 26):
            b.record_caps(2, self._class_counts(c), q.ids, nstripes, c["nclass"], nq, nq_pad, True, cap)
        b.slab_exscan(slab_all, nstripes, nbins, nq_pad)
        rec = self._alloc_records(cap, geo, nq, site="exact")
        scratch_all = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        slab_rel.zero_()
        self._hist(q, g, geo, ternary, label_mode, lw, scratch_all, slab_rel, thresh=thresh,
                   emit=L.CH_EMIT_RELEVANT, rec=rec)
        del scratch_all
        tot_r2 = comm.all_gather(self._local_totals(slab_rel, geo, nbins))
        b.scan_bases(tot_r2, comm.world, comm.rank, nbins, nq, nq_pad, -1, base0_rel, None, None)
        b.slab_exscan(slab_rel, nstripes, nbins, nq_pad)
        return dict(rec=rec, base0_all=base0_all, base0_rel=base0_rel, sbase_all=slab_all, sbase_rel=slab_rel,
                    total_rel=total_rel)

    @staticmethod
    def _host_sample_rows(n, stride, run):
        """rows in the host-side sample of an n-row shard (whole runs only)"""
        return (n // (run * stride)) * run if n >= run * stride else 0

    def _host_sample_view(self, codes, stride, run):
        """(sample rows, 2-D strided view (super rows, run * nbit)) over the first rows of a contiguous host tensor"""
        n, nbit = codes.shape
        ns = self._host_sample_rows(n, stride, run)
        nsup = ns // run
        flat = codes.contiguous() if not codes.is_contiguous() else codes
        view = flat[:nsup * run * stride].view(nsup, run * stride * nbit)[:, :run * nbit]
        return ns, view

    def _stream_per(self, nq_pad):
        """Stripes per streamed launch: the launch has per x ceil(nq_pad / queries per CTA) CTAs, one per SM; pick
        the smallest per <= 8 that fills its last wave best."""
        sms = getattr(self.b, "sm_count", 148)
        groups = max(1, -(-nq_pad // getattr(self.b, "tc_queries_per_cta", 512)))
        best, best_eff = 1, -1.0
        for per in range(1, 9):
            ctas = per * groups
            eff = ctas / (-(-ctas // sms) * sms)
            if eff > best_eff + 0.02:
                best, best_eff = per, eff
        return best

    class _Streamer:
        """Select pass over a HOST-resident gallery, one block of stripes at a time: the H2D copy + sign/bit-pack
        + int8 expansion of block i+1 run (copy engine / tiny kernels) while the tensor-core select kernel works
        on block i.  Outputs are exactly those of one whole-shard launch (slabs per stripe, records with shard
        row ids)."""

        def __init__(self, ev, c, flags, bare=False, pair=False):
            self.ev, self.c, self.flags, self.bare, self.pair = ev, c, flags, bare and not pair, pair
            b, q, g = ev.b, c["q"], c["g"]
            threads, nq_pad, nstripes, rps = c["geo"]
            self.rows_pad = b.padded_rows(g.n)
            self.native = g.loader           # native loader (started by _prepare): the rows arrive on their own
            g.bits = (self.native.bits[self.native.idx["g"]] if self.native is not None
                      else b.empty((self.rows_pad, q.bits.shape[1]), torch.int32))
            if pair:
                # two gallery rows per plane row: row r of the shard lives in plane row block r // 64
                self.plane8 = b.empty((self.rows_pad // 2, b.tc_code_bytes_pair(q.nbit, False)), torch.int8)
                g.i8p = self.plane8
            else:
                kb = b.tc_code_bytes(q.nbit, True) if self.bare else b.tc_code_bytes(q.nbit)
                self.plane8 = b.empty((self.rows_pad, kb), torch.int8)
                if self.bare:
                    g.i8b = self.plane8
                else:
                    g.i8 = self.plane8
            per = ev._stream_per(nq_pad)
            # loads run on a side stream so that they are not queued behind the (long) select kernels
            # one side stream per evaluator, reused by every evaluation: the caching allocator pools blocks per
            # stream, a fresh stream per call would cudaMalloc its copy blocks anew every time
            if getattr(ev, "_side_stream", None) is None:
                ev._side_stream = torch.cuda.Stream(device=g.bits.device)
            self.side = ev._side_stream
            self.pinned = bool(c["db_codes"].is_contiguous() and c["db_codes"].is_pinned())
            # Pageable memory: every block costs ~1 ms of (GIL-free) host packing.  A loader thread runs the blocks back
            # to back from the moment the sample has been packed, while this thread queues the sample-level passes, the
            # select launches and the list kernels (~0.3 ms of host time per block that used to sit between two packs).
            # (Not while the bench brackets entry points with events: the brackets are not thread-safe.)
            self.threaded = bool(ev.stream_loader_thread and not self.pinned and not ev.profile and self.native is None)
            self.thread, self.ready, self.error = None, [], None
            self.loaded = {}
            self.blocks = []
            for s0 in range(0, nstripes, per):
                s1 = min(nstripes, s0 + per)
                r0, r1 = s0 * rps, min(g.n, s1 * rps)
                if r1 > r0:
                    self.blocks.append((s0, s1, r0, r1))

        def load(self, i):
            """H2D + pack + expand of block i on the side stream.  Pageable host memory: staged copies, the host
            returns when the copy is done (kernels are queued).  Pinned host memory: one asynchronous DMA into a
            device block that the caching allocator recycles in stream order -- nothing blocks, so all blocks
            can be queued up front and the copy engine never idles."""
            ev, b, q, g = self.ev, self.ev.b, self.c["q"], self.c["g"]
            s0, s1, r0, r1 = self.blocks[i]
            db_codes = self.c["db_codes"]
            nrow8 = (self.rows_pad - r0) if r1 == g.n else (r1 - r0)
            if self.native is not None:
                # the rows of this block are (being) packed by the native loader: wait for them on the host, make the
                # evaluation's stream wait for their copy, expand them there
                # (the int8 expansion runs on a stream of its own behind the copy: beside the select kernel of the
                # previous block, not between two selects)
                xs = ev._expand_stream
                self.native.wait(self.native.idx["g"], r1, xs)
                with b.on_stream(xs):
                    ev._timed("expand_i8", 0, lambda: b.expand_i8_into(
                        g.bits[r0:r0 + nrow8], q.nbit, self.plane8[r0 // 2 if self.pair else r0:],
                        **(dict(pair=True) if self.pair else dict(bare=self.bare))))
                    done = torch.cuda.Event()
                    done.record(xs)
                torch.cuda.current_stream().wait_event(done)
                self.loaded[i] = done
                return
            blk = db_codes[r0:r1]
            if self.thread is not None:
                # the loader thread: explicit stream handles (the backend's cached stream belongs to the other thread)
                b.pack_sign(blk, 0.0, self.flags, False, out=g.bits[r0:], stream=self.side)
                b.expand_i8_into(g.bits[r0:r0 + nrow8], q.nbit, self.plane8[r0 // 2 if self.pair else r0:],
                                 stream=self.side, **(dict(pair=True) if self.pair else dict(bare=self.bare)))
                done = torch.cuda.Event()
                done.record(self.side)
                self.loaded[i] = done
                return
            with b.on_stream(self.side):
                if self.pinned:
                    def copy_and_pack():
                        dev = blk.to(g.bits.device, non_blocking=True)
                        b.pack_sign(dev, 0.0, self.flags, False, out=g.bits[r0:])
                    ev._timed("pack_host", (r1 - r0) * q.nbit * db_codes.element_size(), copy_and_pack)
                else:
                    ev._timed("pack_host", (r1 - r0) * q.nbit * db_codes.element_size(),
                              lambda: b.pack_sign(blk, 0.0, self.flags, False, out=g.bits[r0:]))
                if self.pair:
                    ev._timed("expand_i8", 0, lambda: b.expand_i8_into(g.bits[r0:r0 + nrow8], q.nbit,
                                                                        self.plane8[r0 // 2:], pair=True))
                else:
                    ev._timed("expand_i8", 0, lambda: b.expand_i8_into(g.bits[r0:r0 + nrow8], q.nbit, self.plane8[r0:],
                                                                        **(dict(bare=True) if self.bare else {})))
                done = torch.cuda.Event()
                done.record(self.side)
            self.loaded[i] = done

        def load_first(self):
            """block 0 -- or, from pinned memory, every block -- is queued while the GPU works on the sample"""
            if self.threaded or self.native is not None:
                return
            for i in range(len(self.blocks) if self.pinned else 1):
                self.load(i)

        def start(self):
            """starts the loader thread (pageable galleries): all blocks, back to back"""
            if not self.threaded or self.thread is not None:
                return
            import threading
            self.ready = [threading.Event() for _ in self.blocks]
            dev = self.ev.b.device

            def run():
                try:
                    torch.cuda.set_device(dev)
                    for i in range(len(self.blocks)):
                        self.load(i)
                        self.ready[i].set()
                except BaseException as e:      # handed to the evaluating thread (select / join)
                    self.error = e
                finally:
                    for r in self.ready:
                        r.set()
            self.thread = threading.Thread(target=run, name="ch-gallery-loader", daemon=True)
            self.thread.start()

        def join(self):
            """the loader thread has finished (it writes into the shard's packed arrays: nothing may outlive it)"""
            t, self.thread = self.thread, None
            if t is not None:
                t.join()
            self.threaded = False

        def wait_block(self, i):
            if self.thread is not None or self.ready:
                self.ready[i].wait()
                if self.error is not None:
                    raise self.error

        def select(self, i, cand, q_i8, dense, thresh, bad=None):
            ev, b, q, g = self.ev, self.ev.b, self.c["q"], self.c["g"]
            threads, nq_pad, nstripes, rps = self.c["geo"]
            s0, s1, r0, r1 = self.blocks[i]
            if self.native is not None:
                self.load(i)
            else:
                self.wait_block(i)
                torch.cuda.current_stream().wait_event(self.loaded[i])
            kw = dict(thresh=thresh, ternary=False) if self.bare else {}
            if self.pair:
                kw = dict(pair=True)
            if bad is not None:
                kw["bad"] = bad
            ev._timed("hist_select_tc", self.c["nq"] * (r1 - r0), lambda: b.hamming_select_tc(
                q_i8=q_i8, g_i8=self.plane8[r0 // 2 if self.pair else r0:], cand=cand, nq=self.c["nq"], nq_pad=nq_pad,
                ndb=r1 - r0, nbit=q.nbit,
                nstripes=s1 - s0, rows_per_stripe=rps, row_base=r0, dense=dense, stripe0=s0, **kw))

    def _pass_topr_sampled(self, c, streamed=False):
        """Top-R in ONE full pass.  A 1-in-``sample_stride`` row sample of the gallery is histogrammed first; from
        it a per-query threshold key t^ is chosen high enough that #(key <= t^) >= R with overwhelming
        probability.  The full pass then counts / matches / records only the pairs with key <= t^.  The result
        is EXACT whenever the full pass confirms #(key <= t^) >= R for every query and no record slice
        overflowed -- which it checks; otherwise ``None`` is returned and the exact two-pass path runs."""
        b, comm, q, g, geo = self.b, self.comm, c["q"], c["g"], c["geo"]
        threads, nq_pad, nstripes, rps = geo
        nbins, nq, label_mode, lw, ternary = c["nbins"], c["nq"], c["label_mode"], c["lw"], c["ternary"]
        need = min(c["rmax"] + c["rf"], c["ndb_total"])
        stride = c["stride"]
        # ---- the sample: every stride-th row of the local shard, same stripes (rps is a multiple of stride) ----
        status = self._status            # [ST_PASS] overflow bits, [ST_SHORT] verification / zeros in a streamed sample
        # (host-known: the sample size and the safety margin decide how dense the candidates of the full pass are)
        ns_host = (sum(self._host_sample_rows(r, stride, max(1, 256 // g.nbit) if g.nbit in (32, 64, 128, 256) else 1)
                       for r in c["rows"]) if streamed else sum((r + stride - 1) // stride for r in c["rows"]))
        mu_host = need * ns_host / max(c["ndb_total"], 1)
        dense = self._dense(1.25 * stride * (int(mu_host + 5.0 * mu_host ** 0.5 + 4.0) + 1), c["ndb_total"])
        sp = Packed()
        sp.i8 = None
        sp.nbit = g.nbit
        sp.nz = None
        streamer = None
        if streamed:
            # the gallery is still on the host: copy just the sample with ONE strided 2-D DMA -- runs of `run`
            # consecutive rows (1 KB segments) every run * stride rows -- and pack it
            run = max(1, 256 // g.nbit) if g.nbit in (32, 64, 128, 256) else 1
            ns, view = self._host_sample_view(c["db_codes"], stride, run)
            zflag = status[ST_SHORT:ST_SHORT + 1]     # a zero / NaN in the streamed codes sends the run to the exact path
            ld = g.loader
            if ld is not None and ld.idx.get("sample") == (stride, run, ns):
                # the native loader packs the sample right behind the queries
                packed = ld.bits[ld.idx["s"]]
                ld.wait(ld.idx["s"], ns // run, torch.cuda.current_stream())
            else:
                packed, _ = self._timed("pack_host", ns * g.nbit * view.element_size(),
                                        lambda: b.pack_sign(view, 0.0, zflag, False))
            sp.bits = packed.view(-1, q.bits.shape[1])         # (super rows, run * words) -> (rows, words)
            streamer = self._Streamer(self, c, zflag, self._bare(dense), self._pair(q))
            if streamer.native is None:
                streamer.side.wait_stream(torch.cuda.current_stream())
            else:
                # (never the loader's copy stream: a wait queued there would sit in front of the copies it waits for)
                self._expand_stream.wait_stream(torch.cuda.current_stream())
            c["streamer"] = streamer       # (the caller joins its loader thread whatever happens below)
            streamer.start()
        else:
            ns, sp.bits = b.gather_rows(g.bits, g.n, g.nbit, stride)
        sp.n = ns
        if ternary:
            _, sp.nz = b.gather_rows(g.nz, g.n, g.nbit, stride)
        sp.ids = sp.masks = sp.info = None
        geo_s = (threads, nq_pad, nstripes, rps // stride)
        if streamed:        # every rank samples the same way (streaming is agreed on by all ranks)
            ns_ranks = [self._host_sample_rows(r, stride, run) for r in c["rows"]]
        else:
            ns_ranks = [(r + stride - 1) // stride for r in c["rows"]]
        ns_total = sum(ns_ranks)
        mu = need * ns_total / max(c["ndb_total"], 1)
        m = int(mu + 5.0 * mu ** 0.5 + 4.0) + 1
        tc_pass = streamed or self._tc_ok(q, ternary, nq_pad)
        slab_s = base_tmp = None
        # block 0 of a streamed gallery travels while the GPU works on the sample: its (host-blocking) copy is
        # issued right after the first sample kernels have been queued
        first_load = streamer.load_first if streamer is not None else (lambda: None)
        # With hints for the list allocations (sites "s1" / "full") nothing below waits for the device: the (host-blocking) copy of block 0
        # is then issued LAST, after every sample-level launch and the set-up of the full pass have been queued -- the
        # GPU works through them while the host's cores pack the block (0.7 ms of a cfg4 evaluation).
        late_load = False
        if streamer is not None and self._hint is not None and hasattr(b, "record_offsets_async"):
            sites = self._hint["sites"]
            late_load = sites.get("full") is not None and sites.get("s1", True) is not None
        if late_load:
            first_load, deferred_load = (lambda: None), streamer.load_first      # (no-ops with a loader thread)
        # per-query failure marks of the candidate-list path (a slice overflowed / fewer than R candidates): the
        # evaluation then re-ranks just those queries by the exact path instead of starting over
        bad = b.zeros((nq_pad,), torch.int32) if tc_pass else None
        if (tc_pass and self.sample_two_level and min(ns_ranks) >= self.sample2_min_rows and
                float(nq) * min(ns_ranks) * int(q.bits.shape[1]) >= self.sample2_min_work):
            thresh, cap, scut = self._sample_thresholds_tc(c, sp, ns_ranks, m, status, first_load, bad)
        else:
            scut = None
            slab_s = b.zeros((nstripes, nbins, nq_pad), torch.int32)
            self._hist(q, sp, geo_s, ternary, L.CH_LAB_NONE, 0, slab_s, None)
            thresh = b.empty((nq_pad,), torch.int32)
            base_tmp = b.empty((nbins, nq_pad), torch.int32)
            tot_s = self._summed_totals(self._local_totals(slab_s, geo_s, nbins))
            b.scan_bases(tot_s, 1, 0, nbins, nq, nq_pad, m, base_tmp, thresh, None)
            # ---- capacities: scaled sample candidate counts (record path: never more than the class counts) ----
            cap = b.empty((nstripes, nq_pad), torch.int32)
            b.record_caps(0, slab_s, thresh, nstripes, nbins, nq, nq_pad, False, cap, sample_stride=stride)
            first_load()
        if not (tc_pass and streamed and not c["pr_k"]):
            self._ensure_labels(q, g)      # (put off by _prepare: the thresholds above did not need them)
        # per-stripe class counts: record capacities of the POPC path; whole-gallery relevant counts for R@k
        cls = self._class_counts(c) if (not tc_pass or c["pr_k"]) else None
        self.stats["sample"] = dict(stride=stride, rows=ns_total, m=m)
        if tc_pass:
            # ---- the one full pass on the tensor cores: candidate lists, then ranks from the lists ----
            cand, tmax = self._alloc_cands(cap, geo, nq, thresh, "full", nbins)   # (a host sync unless hinted)
            del slab_s, base_tmp
            nbins = min(nbins, tmax + 1)
            self.stats["sample"]["key_limit"] = nbins
            tot = None
            # a single rank needs no exchange of the key totals: keys, bases, verification and the in-order walk of
            # every list happen in ONE kernel, launched by _finish (a streamed gallery: when the GPU, not the host, is
            # the bottleneck of the full pass -- stream_fused_rank -- instead of list kernels per block)
            fused_late = (self.fused_rank and comm.world == 1 and not c["rf"] and hasattr(b, "cand_rank") and
                          (not streamed or (self.stream_fused_rank and streamer.native is not None)))
            if streamed:
                q_i8 = self._query_plane(q, nq_pad, thresh, streamer.bare, streamer.pair,
                                         scut if streamer.pair else None, nstripes)
                tot = b.zeros((2, nbins, nq_pad), torch.int32)
                if late_load:
                    deferred_load()
                # list kernels of block i (random L2 gathers, ~0.09 ms) on a stream of their own: they run beside the
                # select kernel of block i + 1 (tensor pipe) instead of between two selects
                side2 = None
                if self.stream_cand_overlap and streamer.native is not None and hasattr(b, "on_stream"):
                    if getattr(self, "_cand_stream", None) is None:
                        self._cand_stream = torch.cuda.Stream(device=q.bits.device)
                    side2, main = self._cand_stream, torch.cuda.current_stream()
                    if g.plane is None and q.nz is None and label_mode == L.CH_LAB_ID and hasattr(b, "gather_plane"):
                        g.plane = b.empty((streamer.rows_pad, b.gather_plane_words(g.nbit)), torch.int32)
                for i in range(len(streamer.blocks)):
                    streamer.select(i, cand, q_i8, dense, thresh, bad)
                    self._ensure_labels(q, g)      # (first needed by the list kernels: packed behind the first select)
                    if fused_late:
                        continue
                    if side2 is not None:
                        side2.wait_stream(main)
                        with b.on_stream(side2):
                            self._cand_hist(c, cand, nbins, tot, streamer.blocks[i][0],
                                            streamer.blocks[i][1] - streamer.blocks[i][0])
                        continue
                    if (not streamer.threaded and streamer.native is None and i + 1 < len(streamer.blocks)
                            and (i + 1) not in streamer.loaded):
                        streamer.load(i + 1)     # host waits for this copy while the GPU runs select(i)
                    # keys / label matches of this block's candidates while the next block is still travelling
                    s0, s1 = streamer.blocks[i][0], streamer.blocks[i][1]
                    self._cand_hist(c, cand, nbins, tot, s0, s1 - s0)
                if side2 is not None:
                    main.wait_stream(side2)
                self.stats["select_kernel"] = "tcgen05"
                self.stats["select_dense"] = bool(dense)
                self.stats["select_threshold"] = "epilogue" if streamer.bare else "contraction"
                self.stats["select_rows_per_cell"] = 2 if streamer.pair else 1
            else:
                self._select_tc(q, g, geo, thresh, cand, dense, bad=bad, scut=scut)
            self.stats["sample"]["stripe_cut"] = scut is not None
            if fused_late:
                return dict(cand=cand, fused=dict(rmax=c["rmax"] + c["rf"], need=need, bad=bad), nbins=nbins,
                            total_rel=self._total_rel_from_classes(c, cls), bad=bad)
            base0_all, base0_rel, key_max = self._cand_bases(c, cand, nbins, need, tot, rmax=c["rmax"] + c["rf"],
                                                             bad=bad)
            return dict(cand=cand, base0_all=base0_all, base0_rel=base0_rel, nbins=nbins, key_max=key_max,
                        total_rel=self._total_rel_from_classes(c, cls), bad=bad)
        b.record_caps(2, cls, q.ids, nstripes, c["nclass"], nq, nq_pad, True, cap)
        rec, tmax = self._alloc_records(cap, geo, nq, thresh, "full", nbins)   # (a host sync unless hinted)
        del slab_s, base_tmp
        # ---- the one full pass; only keys <= max threshold can occur, all slabs / bases are that narrow ----
        nbins = min(nbins, tmax + 1)
        slab_all = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        slab_rel = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        self._hist(q, g, geo, ternary, label_mode, lw, slab_all, slab_rel, thresh=thresh,
                   emit=L.CH_EMIT_RELEVANT, rec=rec, key_limit=nbins)
        base0_all = b.empty((nbins, nq_pad), torch.int32)
        base0_rel = b.empty((nbins, nq_pad), torch.int32)
        found = b.zeros((nq_pad,), torch.int32)
        tot = comm.all_gather(torch.stack([self._local_totals(slab_all, geo, nbins),
                                           self._local_totals(slab_rel, geo, nbins)]))
        b.scan_bases(tot[:, 0].contiguous(), comm.world, comm.rank, nbins, nq, nq_pad, -1, base0_all, None, found)
        b.scan_bases(tot[:, 1].contiguous(), comm.world, comm.rank, nbins, nq, nq_pad, -1, base0_rel, None, None)
        # ---- verification: ST_SHORT is raised if some query has fewer than `need` candidates; it is read back
        # together with the results (the finalisation below runs speculatively)
        b.check_counts(found, nq, need, rec["status"][ST_SHORT:ST_SHORT + 1])
        self.stats["sample"]["key_limit"] = nbins
        b.slab_exscan(slab_all, nstripes, nbins, nq_pad)
        b.slab_exscan(slab_rel, nstripes, nbins, nq_pad)
        return dict(rec=rec, base0_all=base0_all, base0_rel=base0_rel, sbase_all=slab_all, sbase_rel=slab_rel,
                    total_rel=self._total_rel_from_classes(c, cls), nbins=nbins)

    def _sample_thresholds_tc(self, c, sp, ns_ranks, m, status, after_level0=lambda: None, bad=None):
        """Per-query thresholds t^ (and the capacities of the full pass) from the row sample WITHOUT histogramming
        every (query, sample row) pair on the integer pipe:

          level 0  every ``sample2_sub``-th sample row is histogrammed (POPC kernel) -> a loose threshold t0 per
                   query, where that mini-sample holds >= mu0 + 5 sqrt(mu0) + 5 items;
          level 1  the tensor-core select kernel runs over the SAMPLE with t0; its candidate list (a few hundred
                   rows per query) is histogrammed by ``ch_cand_hist`` -> t^ = the smallest key at which the sample
                   holds >= m items -- exactly the threshold the full sample histogram would give -- or t0 if the
                   list holds fewer (then #(key <= t0) in the whole gallery is >= R with overwhelming probability:
                   the mini-sample saw >= m0 of them).
        Capacities come from the per-stripe sample candidates with key <= t^ (``ch_cand_caps``).  Nothing here
        affects exactness: the full pass verifies its counts, an overflow anywhere raises ``status``."""
        b, comm, q = self.b, self.comm, c["q"]
        threads, nq_pad, nstripes, rps = c["geo"]
        nbins, nq, stride = c["nbins"], c["nq"], c["stride"]
        need = min(c["rmax"] + c["rf"], c["ndb_total"])
        ns, sub = sp.n, self.sample2_sub
        # ---- level 0 ----
        ns0 = (ns + sub - 1) // sub
        sp0 = Packed()
        sp0.i8 = sp0.nz = sp0.ids = sp0.masks = sp0.info = None
        sp0.n, sp0.nbit = ns0, sp.nbit
        _, sp0.bits = b.gather_rows(sp.bits, ns, sp.nbit, sub)
        ternary = sp.nz is not None
        if ternary:
            _, sp0.nz = b.gather_rows(sp.nz, ns, sp.nbit, sub)
        align = getattr(b, "stripe_align", 256)
        geo0 = (threads, nq_pad, 1, max(align, (ns0 + align - 1) // align * align))
        slab0 = b.zeros((1, nbins, nq_pad), torch.int32)
        self._hist(q, sp0, geo0, ternary, L.CH_LAB_NONE, 0, slab0, None)
        ns0_total = sum((r + sub - 1) // sub for r in ns_ranks)
        mu0 = need * ns0_total / max(c["ndb_total"], 1)
        m0 = int(mu0 + 5.0 * mu0 ** 0.5 + 4.0) + 1
        thresh0 = b.empty((nq_pad,), torch.int32)
        base_tmp = b.empty((nbins, nq_pad), torch.int32)
        tot0 = self._summed_totals(slab0[0])                           # (1, nbins, nq_pad), summed over the ranks
        b.scan_bases(tot0, 1, 0, nbins, nq, nq_pad, m0, base_tmp, thresh0, None)
        # the sample select has its own stripes: just enough CTAs to fill the GPU
        tile = getattr(b, "tc_tile_rows", 1)
        groups = -(-nq_pad // getattr(b, "tc_queries_per_cta", 512))
        n1 = max(1, min(-(-getattr(b, "sm_count", 148) // groups), max(1, ns // 4096)))
        rps1 = (-(-ns // n1) + tile - 1) // tile * tile
        n1 = max(1, -(-ns // rps1))
        geo1 = (threads, nq_pad, n1, rps1)
        after_level0()              # (host-blocking work of the caller, while the GPU runs level 0)
        cap0 = b.empty((n1, nq_pad), torch.int32)
        b.record_caps(0, tot0, thresh0, n1, nbins, nq, nq_pad, False, cap0, sample_stride=sub, replicate=True)
        cand1, tmax0 = self._alloc_cands(cap0, geo1, nq, thresh0, "s1", nbins)
        nb0 = min(nbins, tmax0 + 1)
        # ---- level 1 ----
        dense = self._dense(1.25 * sub * m0, sum(ns_ranks))
        self._select_tc(q, sp, geo1, thresh0, cand1, dense, kind="sample_select_tc", bad=bad)
        tot1 = b.zeros((nb0, nq_pad), torch.int32)
        kwz = dict(q_nz=q.nz, g_nz=sp.nz) if ternary else {}
        self._timed("cand_hist", 0, lambda: b.cand_hist(
            cand1, q_bits=q.bits, g_bits=sp.bits, q_lab=None, g_lab=None, label_mode=L.CH_LAB_NONE, mask_words=0,
            tot_all=tot1, tot_rel=None, nq=nq, nq_pad=nq_pad, nstripes=n1, nbins=nb0, nbit=q.nbit, **kwz))
        thresh1 = b.empty((nq_pad,), torch.int32)
        base1 = b.empty((nb0, nq_pad), torch.int32)
        b.scan_bases(self._summed_totals(tot1), 1, 0, nb0, nq, nq_pad, m, base1, thresh1, None)
        thresh = torch.minimum(thresh1, thresh0)
        cap = b.empty((nstripes, nq_pad), torch.int32)
        # (key, stripe) refinement of the thresholds: the items with key == t^ are more than half of a key-level
        # candidate list, yet only those of the first few stripes can reach the top R (ties rank by row).  One rank:
        # the cut over (rank, stripe) would need the per-stripe counts of every rank.
        scut = None
        if self.stripe_cut and comm.world == 1 and nstripes > 1 and self._pair(q) and hasattr(b, "expand_i8_query_stripes"):
            scut = b.empty((nq_pad,), torch.int32)
        b.cand_caps(cand1, thresh, n1, rps // stride, nstripes, nq, nq_pad, stride, cap, m=m, scut=scut)
        self.stats["sample2"] = dict(sub=sub, m0=m0, key_limit0=nb0, slots=self.stats.get("record_slots"))
        return thresh, cap, scut

    def _total_rel_from_classes(self, c, cls):
        """relevant items in the whole gallery per query = class frequency of the query's class (single-label)"""
        b, comm, q, nq = self.b, self.comm, c["q"], c["nq"]
        if not c["pr_k"]:
            return None
        total_rel = b.zeros((c["geo"][1],), torch.int32)
        if c["pr_k"]:
            cls_tot = cls.sum(0, dtype=torch.int32)
            cls_tot = comm.all_reduce_sum(cls_tot) if comm.world > 1 else cls_tot
            qid = q.ids[:nq].to(torch.int64)
            ok = (qid >= 0) & (qid < c["nclass"])
            total_rel[:nq] = torch.where(ok, cls_tot[qid.clamp(0, c["nclass"] - 1)], torch.zeros_like(cls_tot[:1]))
        return total_rel

    def _summed_totals(self, tot):
        """(nbins, nq_pad) key totals summed over the ranks, shaped (1, nbins, nq_pad) for ``scan_bases`` with
        world = 1: thresholds only need the SUM (an all-reduce moves 1/world of what the all-gather would)."""
        if self.comm.world > 1:
            tot = self.comm.all_reduce_sum(tot.contiguous())
        return tot.unsqueeze(0)

    def _local_totals(self, slab, geo, nbins):
        threads, nq_pad, nstripes, rps = geo
        tot = self.b.empty((nbins, nq_pad), torch.int32)
        self.b.slab_totals(slab, nstripes, nbins, nq_pad, tot)
        return tot

    def _tc_geometry(self, geo, ndb, stride, w_tc):
        """Stripe count for an evaluation whose select pass runs on the tensor cores.  That kernel has one CTA per
        SM and ceil(nq_pad / queries per CTA) CTAs per stripe; the XOR+POPC kernel (sample / count pass, same stripes) has
        nq_pad / threads CTAs per stripe and ~3 per SM.  Score = modelled efficiency of both (longest stripe x
        waves), weighted by the share of the step each pass has (``w_tc``), with a slight preference for few
        stripes (fewer, tighter candidate slices)."""
        threads, nq_pad, nstripes, rps = geo
        sms = self.b.sm_count
        align = getattr(self.b, "stripe_align", 256) * max(stride, 1)
        groups, qtiles = -(-nq_pad // getattr(self.b, "tc_queries_per_cta", 512)), -(-nq_pad // threads)
        best, best_score = None, -1.0
        for n in range(1, 97):
            r = -(-ndb // n)
            r = (r + align - 1) // align * align
            if -(-ndb // r) != n or (n > 1 and r < 4096):
                continue

            def eff(ctas_per_stripe, slots):
                waves = -(-(ctas_per_stripe * n) // slots)
                return ndb * ctas_per_stripe / float(waves * slots * r)
            score = w_tc * eff(groups, sms) + (1.0 - w_tc) * eff(qtiles, 3 * sms) - 0.001 * n
            if score > best_score:
                best, best_score = (n, r), score
        if best is None:
            return geo
        return threads, nq_pad, best[0], best[1]

    def _agree_geometry(self, geo, ndb=None, stride=1, min_stripes=0):
        """threads / nq_pad depend only on (nq, nbins) and are identical on all ranks; the stripe layout is
        per rank (shards may differ in length), so nothing has to be exchanged.  With row sampling the stripe
        length is rounded up so that the sample of a stripe is itself a legal stripe."""
        threads, nq_pad, nstripes, rps = geo
        if self.stripe_rows_override and ndb is not None:
            rps = int(self.stripe_rows_override)
        align = getattr(self.b, "stripe_align", 256)
        if min_stripes > nstripes and ndb is not None:
            rps = max(align, -(-ndb // min_stripes))
            rps = (rps + align - 1) // align * align
        if stride > 1:
            unit = stride * align
            rps = (rps + unit - 1) // unit * unit
        if ndb is not None:
            nstripes = max(1, (ndb + rps - 1) // rps)
        return threads, nq_pad, nstripes, rps

    # ------------------------------------------------------------------ ranked retrieval
    def retrieve(self, db_codes, q_codes, R, threshold=0.0, remove_first_retrieved=False, zero_mean=False):
        """Exact ranked retrieval: ``(ids int64 (nq, L), keys int32 (nq, L), ternary)`` with
        ``L = min(R, gallery size)``, canonical order (distance, then global gallery row).
        keys = Hamming distance (binary codes) or 2 x distance (ternary)."""
        b, comm = self.b, self.comm
        if hasattr(b, "begin"):
            b.begin()
        self.host_syncs = 0
        self._hint, self._new_hint = None, dict(sites={}, complete=False)
        try:
            return self._retrieve(db_codes, q_codes, R, threshold, remove_first_retrieved, zero_mean)
        except _TooManySlots as e:
            raise RuntimeError(f"{int(e.args[0])} candidate slots exceed the 32-bit slot index: retrieve the "
                               "queries in chunks") from None

    def _retrieve(self, db_codes, q_codes, R, threshold, remove_first_retrieved, zero_mean):
        b, comm = self.b, self.comm
        q, g, ternary, _, _, _, rows = self._prepare(db_codes, None, q_codes, None, threshold, zero_mean=zero_mean)
        nq, nbit = q.n, q.nbit
        nbins = (2 * nbit if ternary else nbit) + 1
        ndb_total = sum(rows)
        rf = 1 if remove_first_retrieved else 0
        list_len = max(ndb_total - rf, 0)
        R = list_len if R == -1 else min(int(R), list_len)
        row_offset = sum(rows[:comm.rank])
        geo = self._agree_geometry(b.geometry(nq, g.n, nbit, ternary, L.CH_LAB_NONE, 0), g.n)
        threads, nq_pad, nstripes, rps = geo
        slab_all = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        self._hist(q, g, geo, ternary, L.CH_LAB_NONE, 0, slab_all, None)
        base0_all = b.empty((nbins, nq_pad), torch.int32)
        thresh = b.empty((nq_pad,), torch.int32)
        tot_a = comm.all_gather(self._local_totals(slab_all, geo, nbins))
        b.scan_bases(tot_a, comm.world, comm.rank, nbins, nq, nq_pad, R + rf, base0_all, thresh, None)
        cap = b.empty((nstripes, nq_pad), torch.int32)
        b.record_caps(0, slab_all, thresh, nstripes, nbins, nq, nq_pad, False, cap)
        ids = b.full((nq, max(R, 1)), -1, torch.int64)
        keys = b.full((nq, max(R, 1)), -1, torch.int32)
        if self._tc_ok(q, ternary, nq_pad) and R > 0:
            cand = self._alloc_cands(cap, geo, nq)
            self._select_tc(q, g, geo, thresh, cand, self._dense(1.3 * (R + rf), ndb_total))
            c = dict(q=q, g=g, geo=geo, nq=nq, label_mode=L.CH_LAB_NONE, lw=0)
            base0_c, _, key_max = self._cand_bases(c, cand, nbins, rmax=R + rf)
            b.cand_finalize(cand, mode=2, base0_all=base0_c, base0_rel=None, nq=nq, nq_pad=nq_pad, nstripes=nstripes,
                            nbins=nbins, remove_first=bool(rf), ids=ids, keys=keys, R=R, row_offset=row_offset,
                            key_max=key_max)
            # (group mode: the verdict is agreed on before anybody raises -- a rank that raised alone would leave
            # the others waiting in the collectives below)
            err = comm.all_reduce_max(cand["err"].clone()) if comm.world > 1 else cand["err"]
            if self._host_ints(err)[0] != 0:
                raise RuntimeError("internal error: candidate list overflow")
            if comm.world > 1:
                ids = comm.all_reduce_max(ids)
                keys = comm.all_reduce_max(keys)
            return ids[:, :R], keys[:, :R], ternary
        b.slab_exscan(slab_all, nstripes, nbins, nq_pad)
        rec = self._alloc_records(cap, geo, nq)
        scratch_all = b.zeros((nstripes, nbins, nq_pad), torch.int32)
        self._hist(q, g, geo, ternary, L.CH_LAB_NONE, 0, scratch_all, None, thresh=thresh,
                   emit=L.CH_EMIT_CANDIDATES, rec=rec)
        f = dict(recs=rec["recs"], rec_off=rec["off"], rec_cnt=rec["cnt"], base0_all=base0_all, base0_rel=None,
                 sbase_all=slab_all, sbase_rel=None, first_rel=None, partial=None, cols=None, nq=nq, nq_pad=nq_pad,
                 nstripes=nstripes, nbins=nbins, remove_first=bool(rf), r_eff=[], pr_k=[])
        b.scatter_ranked(f, R, row_offset, ids, keys)
        err = comm.all_reduce_max(rec["err"].clone()) if comm.world > 1 else rec["err"]
        if self._host_ints(err)[0] != 0:
            raise RuntimeError("internal error: record buffer overflow")
        if comm.world > 1:
            ids = comm.all_reduce_max(ids)
            keys = comm.all_reduce_max(keys)
        return ids[:, :R], keys[:, :R], ternary

    # ------------------------------------------------------------------ dense distances (small)
    def hamming_matrix(self, a_codes, b_codes, threshold=0.0):
        """Dense key matrix (na, nb) int16 and the ternary flag (local, no collectives)."""
        if hasattr(self.b, "begin"):
            self.b.begin()
        self.col_sub = None
        flags = self.b.zeros((1,), torch.int32)
        pa = self._pack_side(a_codes, None, threshold, flags, 0, want_nz=True)
        pb = self._pack_side(b_codes, None, 0.0, flags, 0, want_nz=True)
        fl = self._host_ints(flags)[0]
        if fl & 2:
            raise ValueError("codes contain NaN")
        ternary = bool(fl & 1)
        return self.b.hamming_matrix(pa.bits, pa.nz, pb.bits, pb.nz, pa.n, pb.n, pa.nbit, ternary), ternary
