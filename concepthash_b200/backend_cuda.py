"""Thin torch-tensor wrappers over the C-ABI (one method per exported kernel entry point).

torch is plumbing here: it owns device memory, the current stream and (in ``evaluator``) the NCCL
collectives.  All arithmetic of the hot path runs in libconcepthash_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_DTYPES = {
    torch.float32: L.CH_F32, torch.float16: L.CH_F16, torch.bfloat16: L.CH_BF16, torch.float64: L.CH_F64,
    torch.int64: L.CH_I64, torch.int32: L.CH_I32, torch.uint8: L.CH_U8, torch.int16: L.CH_I16,
    torch.int8: L.CH_I8,
}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _HostLoader:
    """Handle of a running native host loader (``CudaBackend.host_loader_start``).  Keeps the source tensors and the
    destinations alive until ``join``.  ``bits[j]`` = the packed array of job j."""

    def __init__(self, lib, handle, sources, bits):
        self.lib, self.h, self.sources, self.bits = lib, handle, sources, list(bits)
        self.flags = None

    def wait(self, job, rows, stream, block=False):
        """makes ``stream`` -- any stream but the loader's own -- wait until rows [0, rows) of ``job`` have landed on
        the device; blocks the calling thread (GIL released) until their copy has been queued."""
        assert self.h is not None, "loader already joined"
        L.check(self.lib.ch_host_loader_wait(self.h, int(job), int(rows), C.c_void_p(stream.cuda_stream), int(block)),
                "ch_host_loader_wait")

    def join(self):
        """waits for the loader thread; returns the flag bits per job (1: a zero sign, 2: NaN).  Idempotent."""
        if self.h is not None:
            h, self.h = self.h, None
            fl = (C.c_uint32 * L.CH_LOADER_MAX_JOBS)()
            rc = self.lib.ch_host_loader_join(h, fl)
            self.flags = [int(fl[j]) for j in range(len(self.bits))]
            self.sources = None
            L.check(rc, "ch_host_loader_join")
        return self.flags

    def __del__(self):
        try:
            self.join()
        except Exception:
            pass


class CudaBackend:
    """One instance per process / GPU (owns a ``ch_ws`` workspace)."""

    name = "cuda"
    stripe_align = 256     # stripe boundaries must be multiples of 256 gallery rows
    tc_tile_rows = 256     # rows per tile of the tensor-core select kernel, paired form (its stripes are whole tiles)

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("concepthash_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = L.load()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else
                                   (device.index if isinstance(device, torch.device) else int(device)))
        ws = C.c_void_p()
        L.check(self.lib.ch_workspace_create(self.device.index, C.byref(ws)), "ch_workspace_create")
        self.ws = ws
        sm, smem, l2, khz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.check(self.lib.ch_device_info(self.ws, C.byref(sm), C.byref(smem), C.byref(l2), C.byref(khz)))
        self.sm_count, self.max_smem, self.l2_bytes, self.clock_khz = sm.value, smem.value, l2.value, khz.value
        self.tc_queries_per_cta = int(self.lib.ch_tc_queries_per_cta())
        self.capture_results = None      # set while an evaluation is captured into a CUDA graph (see reduce_means)
        self.last_results = None
        self.last_result_shape = (0, 0, 0)

    def __del__(self):
        try:
            if getattr(self, "ws", None):
                self.lib.ch_workspace_destroy(self.ws)
                self.ws = None
        except Exception:
            pass

    # ---- plumbing ----
    def begin(self):
        """Called once per evaluation: caches the current torch stream handle (looked up ~20x per evaluation)."""
        self._stream_cache = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def on_stream(self, torch_stream):
        """Context manager: the entry points launch on ``torch_stream`` instead of the evaluation's stream."""
        backend = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.saved = getattr(backend, "_stream_cache", None)
                backend._stream_cache = C.c_void_p(torch_stream.cuda_stream)
                self_inner.guard = torch.cuda.stream(torch_stream)
                self_inner.guard.__enter__()

            def __exit__(self_inner, *exc):
                self_inner.guard.__exit__(*exc)
                backend._stream_cache = self_inner.saved
        return _Ctx()

    def _stream(self):
        s = getattr(self, "_stream_cache", None)
        if s is None:
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return s

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def full(self, shape, value, dtype):
        return torch.full(shape, value, dtype=dtype, device=self.device)

    def to_host(self, t):
        return t.cpu()

    def padded_rows(self, n):
        return int(self.lib.ch_padded_rows(int(n)))

    def code_words(self, nbit):
        return int(self.lib.ch_code_words(int(nbit)))

    def launch_count(self):
        return int(self.lib.ch_launch_count(self.ws))

    # ---- K1 ----
    def _src(self, t):
        """(tensor kept alive, mem kind) for a caller tensor: CUDA tensors are used in place, CPU tensors
        are handed to the library as host pointers (it stages them in pipelined chunks)."""
        if t.dtype == torch.bool:
            t = t.to(torch.uint8)
        if t.dtype not in _DTYPES:
            t = t.to(torch.float32)
        if t.is_cuda:
            if t.device != self.device:
                t = t.to(self.device)
            return t, L.CH_MEM_DEVICE
        if t.dim() == 2 and t.stride(1) != 1 or (t.dim() == 2 and t.shape[0] > 1 and t.stride(0) < t.shape[1]):
            t = t.contiguous()
        return t, L.CH_MEM_HOST

    def column_sums(self, codes):
        """fp64 column sums of DEVICE codes (n, nbit) -> f64[nbit] (deterministic)"""
        if not codes.dtype.is_floating_point:
            codes = codes.to(torch.float32)
        t, mem = self._src(codes)
        assert mem == L.CH_MEM_DEVICE
        n, nbit = t.shape
        out = self.empty((nbit,), torch.float64)
        L.check(self.lib.ch_column_sums(self.ws, _ptr(t), _DTYPES[t.dtype], n, nbit, t.stride(0) if n > 1 else nbit,
                                        t.stride(1) if nbit > 1 else 1, _ptr(out), self._stream()), "ch_column_sums")
        return out

    def pack_sign(self, codes, threshold, flags, want_nz=True, out=None, col_sub=None, stream=None):
        """codes (n, nbit) real -> (bits, nz) u32 (rows_pad, words); ``flags`` u32[1] is OR-ed.
        ``want_nz=False`` skips the non-zero plane (zeros are still detected in ``flags``) and lets
        contiguous inputs take the flat fast path."""
        n, nbit = codes.shape
        words = self.code_words(nbit)
        if words == 0:
            raise ValueError(f"nbit={nbit} unsupported (1..{L.CH_MAX_NBIT})")
        if not codes.dtype.is_floating_point:
            codes = codes.to(torch.float32)
        t, mem = self._src(codes)
        rows = self.padded_rows(n)
        if out is not None:      # a row block of a larger packed array (streamed galleries); needs `rows` rows
            assert not want_nz and out.shape[0] >= rows and out.shape[1] == words and out.is_contiguous()
            bits = out
        else:
            bits = self.empty((rows, words), torch.int32)
        nz = self.empty((rows, words), torch.int32) if want_nz else None
        thr = float(threshold)
        if thr != 0.0:   # torch compares `codes.abs() < threshold` in the dtype of codes
            thr = float(torch.tensor(thr, dtype=t.dtype))
        rs = t.stride(0) if n > 1 else nbit
        L.check(self.lib.ch_pack_sign(self.ws, _ptr(t), mem, _DTYPES[t.dtype], n, nbit, rs,
                                      t.stride(1) if nbit > 1 else 1, thr, _ptr(col_sub), _ptr(bits), _ptr(nz),
                                      _ptr(flags), self._stream() if stream is None else C.c_void_p(stream.cuda_stream)),
                "ch_pack_sign")
        return bits, nz

    def host_loader_ok(self, codes):
        """can ``host_loader_start`` take this gallery?  (a contiguous-row fp32 CPU tensor, host packing enabled)"""
        return (isinstance(codes, torch.Tensor) and not codes.is_cuda and codes.dtype == torch.float32 and
                codes.dim() == 2 and codes.shape[0] > 0 and self.code_words(int(codes.shape[1])) > 0 and
                (codes.shape[1] == 1 or codes.stride(1) == 1) and
                (codes.shape[0] == 1 or codes.stride(0) >= codes.shape[1]) and codes.data_ptr() % 4 == 0 and
                int(self.lib.ch_host_pack_threads(self.ws)) > 0)

    def host_loader_start(self, jobs, stream):
        """Starts the native host loader (csrc/loader.cu).  ``jobs`` = [(codes, bits, flags), ...] in the order they are
        needed: ``codes`` (n, nbit) fp32 on the HOST are sign/bit-packed by the host's cores on a thread of their own
        and copied chunk by chunk into ``bits`` (rows_pad, words) on ``stream`` (which must carry nothing else until
        ``join``); ``flags`` u32[1] gets the zero / NaN bits.  Returns a handle: ``wait(job, rows, stream)`` /
        ``join() -> flag bits per job``."""
        arr = (L.LoaderJob * len(jobs))()
        for k, job in enumerate(jobs):
            if job[0] == "copy":
                # ("copy", src, dst): a contiguous 1-D host array (int64 / int32 / float32 class ids) travels as it is
                _, src, dst = job
                es = src.element_size()
                assert src.dim() == 1 and src.is_contiguous() and dst.is_contiguous() and es % 4 == 0
                assert dst.numel() == src.numel() and dst.dtype == src.dtype and src.data_ptr() % 4 == 0
                arr[k].codes_host, arr[k].n, arr[k].nbit, arr[k].kind = src.data_ptr(), int(src.numel()), es, L.CH_LOADER_COPY
                arr[k].row_stride, arr[k].out_bits_dev, arr[k].flags_dev = es, dst.data_ptr(), None
                continue
            codes, bits, flags = job
            n, nbit = int(codes.shape[0]), int(codes.shape[1])
            assert bits.shape[0] >= self.padded_rows(n) and bits.shape[1] == self.code_words(nbit) and bits.is_contiguous()
            arr[k].codes_host, arr[k].n, arr[k].nbit, arr[k].kind = codes.data_ptr(), n, nbit, L.CH_LOADER_PACK
            arr[k].row_stride = codes.stride(0) if n > 1 else nbit
            arr[k].out_bits_dev, arr[k].flags_dev = bits.data_ptr(), (flags.data_ptr() if flags is not None else None)
        h = C.c_void_p()
        L.check(self.lib.ch_host_loader_start(self.ws, arr, len(jobs), C.c_void_p(stream.cuda_stream), C.byref(h)),
                "ch_host_loader_start")
        return _HostLoader(self.lib, h, [j[1] if j[0] == "copy" else j[0] for j in jobs],
                           [j[2] if j[0] == "copy" else j[1] for j in jobs])

    def pack_labels(self, labels, nolabel, info=None):
        """labels (n, C) one-/multi-hot or (n,) ids -> (ids u32 (rows_pad), masks (rows_pad, lw) | None, info u32[4]).
        ``info`` may be a zeroed u32[4] view supplied by the caller (statistics are accumulated into it)."""
        t, mem = self._src(labels)
        n = t.shape[0]
        rows = self.padded_rows(n)
        ids = self.empty((rows,), torch.int32)
        if info is None:
            info = self.zeros((4,), torch.int32)
        if t.dim() == 2:
            ncls = t.shape[1]
            lw = (ncls + 31) // 32
            masks = self.empty((rows, max(lw, 1)), torch.int32)
            rs = t.stride(0) if n > 1 else ncls
            L.check(self.lib.ch_pack_labels(self.ws, _ptr(t), mem, _DTYPES[t.dtype], n, ncls, rs,
                                            t.stride(1) if ncls > 1 else 1, nolabel, _ptr(ids), _ptr(masks),
                                            _ptr(info), self._stream()), "ch_pack_labels")
            return ids, masks, info
        if t.dtype in (torch.float16, torch.bfloat16):
            t = t.to(torch.float32)
        L.check(self.lib.ch_pack_labels(self.ws, _ptr(t), mem, _DTYPES[t.dtype], n, 0,
                                        t.stride(0) if n > 1 else 1, 1, nolabel, _ptr(ids), None, _ptr(info),
                                        self._stream()), "ch_pack_labels")
        return ids, None, info

    # ---- K2 ----
    def geometry(self, nq, ndb, nbit, ternary, label_mode, lw):
        th, st, rps = C.c_int32(), C.c_int32(), C.c_int32()
        nqp = C.c_int64()
        L.check(self.lib.ch_hist_geometry(self.ws, nq, ndb, nbit, int(ternary), label_mode, lw, C.byref(th),
                                          C.byref(nqp), C.byref(st), C.byref(rps)), "ch_hist_geometry")
        return th.value, nqp.value, st.value, rps.value

    # ---- tensor-core select pass ----
    def tc_code_bytes(self, nbit, bare=False):
        """bytes per row of an int8 operand plane; ``bare``: without threshold slots (comparison in the epilogue)"""
        return int((self.lib.ch_tc_code_bytes_bare if bare else self.lib.ch_tc_code_bytes)(int(nbit)))

    def tc_code_bytes_pair(self, nbit, ternary=False):
        """bytes per row of the PAIRED planes (two gallery rows per plane row); 0: this nbit has no paired form"""
        return int(self.lib.ch_tc_code_bytes_pair(int(nbit), int(bool(ternary))))

    def expand_i8_into(self, bits, nbit, out, bare=False, pair=False, stream=None):
        """gallery plane of a row block: ``bits`` (rows, words) -> ``out`` (rows, kb) views of larger arrays
        (``pair``: rows / 2 plane rows; the block starts on a multiple of 64 rows)"""
        n = int(bits.shape[0])
        assert n % (64 if pair else 32) == 0 and out.shape[0] >= (n // 2 if pair else n)
        L.check(self.lib.ch_expand_i8(self.ws, _ptr(bits), None, n, nbit, 0, 3 if pair else int(bool(bare)), _ptr(out),
                                      n // 2 if pair else n, None, 0,
                                      self._stream() if stream is None else C.c_void_p(stream.cuda_stream)), "ch_expand_i8")

    def expand_i8(self, bits, nbit, min_rows=0, thresh=None, nq=0, nz=None, bare=False, query=False, pair=False):
        """packed sign bits (rows_pad, words) -> {-1, 0, +1} int8 plane in the tiled operand order (rows, kb) int8,
        with the threshold slots: gallery plane when ``thresh`` is None, else the query plane of ``nq`` queries.
        ``nz``: the non-zero plane of ternary codes (thresholds are then on the doubled key scale).
        ``min_rows`` over-allocates so that whole 128-query tiles can be read.  ``pair``: the paired forms
        (``ch_tc_code_bytes_pair``): a gallery plane of rows_pad / 2 plane rows, or -- with ``thresh`` -- its query plane."""
        rows_pad = int(bits.shape[0])
        if pair:
            kb = self.tc_code_bytes_pair(nbit, nz is not None)
            gallery = thresh is None
            assert kb > 0 and (not gallery or rows_pad % 64 == 0)
            rows = rows_pad // 2 if gallery else max(rows_pad, (int(min_rows) + 31) // 32 * 32)
            out = self.empty((rows, kb), torch.int8)
            L.check(self.lib.ch_expand_i8(self.ws, _ptr(bits), _ptr(nz), rows_pad, nbit, 0 if nz is None else 1,
                                          3 if gallery else 4, _ptr(out), rows, _ptr(thresh), int(nq), self._stream()),
                    "ch_expand_i8")
            return out
        kb = self.tc_code_bytes(nbit, bare)
        rows = max(rows_pad, (int(min_rows) + 31) // 32 * 32)
        out = self.empty((rows, kb), torch.int8)
        L.check(self.lib.ch_expand_i8(self.ws, _ptr(bits), _ptr(nz), rows_pad, nbit, 0 if nz is None else 1,
                                      (2 if query else 1) if bare else 0, _ptr(out), rows,
                                      None if bare else _ptr(thresh), int(nq), self._stream()), "ch_expand_i8")
        return out

    def hamming_select_tc(self, *, q_i8, g_i8, cand, nq, nq_pad, ndb, nbit, nstripes, rows_per_stripe, row_base=0,
                          dense=False, stripe0=0, thresh=None, ternary=False, bad=None, pair=False):
        """``cand``: dict(off, cap, cnt (nstripes_total, nq_pad) u32, rows u32[], err u32[1]); ``stripe0`` = first
        stripe of this call's row block (streamed galleries).  ``thresh``: the per-query thresholds when both planes
        are ``bare`` (comparison in the sparse epilogue instead of a threshold block in the contraction).
        ``pair``: both planes are the paired forms (tiles of 256 gallery rows)."""
        a = L.SelectArgs()
        a.thresh = thresh.data_ptr() if thresh is not None else None
        a.ternary = int(bool(ternary))
        a.pair = int(bool(pair))
        a.bad = bad.data_ptr() if bad is not None else None
        a.q_i8, a.g_i8 = q_i8.data_ptr(), g_i8.data_ptr()
        if q_i8.dim() == 3:                   # one query plane per stripe (expand_i8_query_stripes)
            a.q_i8 = q_i8[stripe0:].data_ptr()
            a.q_stripe_bytes = q_i8.stride(0) * q_i8.element_size()
        a.cand_off, a.cand_cap, a.cand_cnt = (cand[k][stripe0:].data_ptr() for k in ("off", "cap", "cnt"))
        a.cand_rows, a.err_flag = cand["rows"].data_ptr(), cand["err"].data_ptr()
        a.nq, a.nq_pad, a.ndb, a.row_base = nq, nq_pad, ndb, int(row_base)
        a.nbit, a.nstripes, a.rows_per_stripe, a.dense = nbit, nstripes, rows_per_stripe, int(bool(dense))
        L.check(self.lib.ch_hamming_select_tc(self.ws, C.byref(a), self._stream()), "ch_hamming_select_tc")

    # ---- K3/K4 on candidate lists ----
    def _cand_args(self, cand, nq, nq_pad, nstripes, nbins, stripe0=0, **kw):
        a = L.CandArgs()
        a.cand_off, a.cand_cnt = cand["off"][stripe0:].data_ptr(), cand["cnt"][stripe0:].data_ptr()
        a.cand_rows, a.cand_key, a.err_flag = cand["rows"].data_ptr(), cand["key"].data_ptr(), cand["err"].data_ptr()
        a.nq, a.nq_pad, a.nstripes, a.nbins = nq, nq_pad, nstripes, nbins
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                setattr(a, k, v.data_ptr())
            elif v is not None:
                setattr(a, k, v)
        return a

    def gather_plane_words(self, nbit):
        return int(self.lib.ch_gather_plane_words(int(nbit)))

    def gather_plane(self, bits, ids, nbit, out=None):
        """[code words | class id | pad] per gallery row (single-label shards): one sector per candidate lookup"""
        rows = int(bits.shape[0])
        if out is None:
            out = self.empty((rows, self.gather_plane_words(nbit)), torch.int32)
        assert out.is_contiguous() and out.shape[0] >= rows and ids.shape[0] >= rows
        L.check(self.lib.ch_gather_plane(self.ws, _ptr(bits), _ptr(ids), rows, nbit, _ptr(out), self._stream()),
                "ch_gather_plane")
        return out

    def cand_hist(self, cand, *, q_bits, g_bits, q_lab, g_lab, label_mode, mask_words, tot_all, tot_rel, nq, nq_pad,
                  nstripes, nbins, nbit, stripe0=0, g_plane=None, q_nz=None, g_nz=None):
        """totals are accumulated; ``stripe0`` / ``nstripes`` select a block of stripes of the list; ``q_nz`` /
        ``g_nz``: non-zero planes of ternary codes (keys = 2 x distance)"""
        a = self._cand_args(cand, nq, nq_pad, nstripes, nbins, stripe0, g_plane=g_plane, q_bits=q_bits,
                            g_bits=g_bits, q_nz=q_nz, g_nz=g_nz, q_lab=q_lab, g_lab=g_lab, label_mode=label_mode,
                            mask_words=mask_words, tot_all=tot_all, tot_rel=tot_rel, nbit=nbit)
        L.check(self.lib.ch_cand_hist(self.ws, C.byref(a), self._stream()), "ch_cand_hist")

    def cand_rank(self, cand, *, q_bits, g_bits, q_lab, g_lab, label_mode, mask_words, nq, nq_pad, nstripes, nbins, nbit,
                  cols, r_eff=(), pr_k=(), rmax=-1, need=0, status=None, bad=None, g_plane=None, q_nz=None, g_nz=None):
        """``cand_hist`` + ``scan_bases_pair`` + ``cand_finalize`` (mode 0) of a single rank in ONE kernel"""
        r_eff, pr_k = list(r_eff), list(pr_k)
        if len(r_eff) > L.CH_MAX_R or len(pr_k) > L.CH_MAX_PR:
            raise ValueError(f"at most {L.CH_MAX_R} R values and {L.CH_MAX_PR} PRs cut-offs are supported")
        a = self._cand_args(cand, nq, nq_pad, nstripes, nbins, 0, g_plane=g_plane, q_bits=q_bits, g_bits=g_bits,
                            q_nz=q_nz, g_nz=g_nz, q_lab=q_lab, g_lab=g_lab, label_mode=label_mode,
                            mask_words=mask_words, nbit=nbit, mode=0, cols=cols, nR=len(r_eff), nPR=len(pr_k))
        for i, v in enumerate(r_eff):
            a.r_eff[i] = int(v)
        for i, v in enumerate(pr_k):
            a.pr_k[i] = int(v)
        L.check(self.lib.ch_cand_rank(self.ws, C.byref(a), int(rmax), int(need), _ptr(status), _ptr(bad),
                                      self._stream()), "ch_cand_rank")

    def cand_finalize(self, cand, *, mode, base0_all, base0_rel, nq, nq_pad, nstripes, nbins, remove_first=False,
                      first_rel=None, first_rel_out=None, cols=None, r_eff=(), pr_k=(), ids=None, keys=None, R=0,
                      row_offset=0, key_max=None):
        r_eff, pr_k = list(r_eff), list(pr_k)
        if len(r_eff) > L.CH_MAX_R or len(pr_k) > L.CH_MAX_PR:
            raise ValueError(f"at most {L.CH_MAX_R} R values and {L.CH_MAX_PR} PRs cut-offs are supported")
        a = self._cand_args(cand, nq, nq_pad, nstripes, nbins, mode=mode, base0_all=base0_all, base0_rel=base0_rel,
                            remove_first=int(bool(remove_first)), first_rel=first_rel, first_rel_out=first_rel_out,
                            cols=cols, ids=ids, keys=keys, R=int(R), row_offset=int(row_offset), key_max=key_max,
                            nR=len(r_eff),
                            nPR=len(pr_k))
        for i, v in enumerate(r_eff):
            a.r_eff[i] = int(v)
        for i, v in enumerate(pr_k):
            a.pr_k[i] = int(v)
        L.check(self.lib.ch_cand_finalize(self.ws, C.byref(a), self._stream()), "ch_cand_finalize")

    def cand_caps(self, cand, thresh, list_stripes, rows_per_stripe, nstripes, nq, nq_pad, sample_stride, cap,
                  m=0, scut=None):
        """``scut`` (nq_pad) out: the (key, stripe) refinement of the thresholds -- stripes >= scut[q] are selected with
        thresh[q] - 1 (``expand_i8_query_stripes``); ``m`` = the sample count the candidate prefix has to reach"""
        L.check(self.lib.ch_cand_caps(self.ws, _ptr(cand["off"]), _ptr(cand["cnt"]), _ptr(cand["rows"]),
                                      _ptr(cand["key"]), _ptr(thresh), int(list_stripes), int(rows_per_stripe),
                                      int(nstripes), nq, nq_pad, int(sample_stride), _ptr(cap), int(m), _ptr(scut),
                                      self._stream()),
                "ch_cand_caps")

    def expand_i8_query_stripes(self, bits, nbit, min_rows, thresh, scut, nstripes, nq, nz=None, stripe0=0):
        """paired query planes, one per stripe: (nstripes, rows, kb) int8 (thresh[q] - 1 for stripes >= scut[q])"""
        kb = self.tc_code_bytes_pair(nbit, nz is not None)
        rows_pad = int(bits.shape[0])
        rows = max(rows_pad, (int(min_rows) + 31) // 32 * 32)
        out = self.empty((int(nstripes), rows, kb), torch.int8)
        L.check(self.lib.ch_expand_i8_query_stripes(self.ws, _ptr(bits), _ptr(nz), rows_pad, nbit, 0 if nz is None else 1,
                                                    _ptr(out), rows, _ptr(thresh), _ptr(scut), int(stripe0),
                                                    int(nstripes), int(nq), self._stream()),
                "ch_expand_i8_query_stripes")
        return out

    def hamming_hist(self, **kw):
        a = self._hist_args(**kw)
        L.check(self.lib.ch_hamming_hist(self.ws, C.byref(a), self._stream()), "ch_hamming_hist")

    def _hist_args(self, *, q_bits, q_nz, g_bits, g_nz, q_lab, g_lab, slab_all, slab_rel, thresh, rec_off,
                   rec_cap, rec_cnt, recs, err_flag, nq, nq_pad, ndb, nbit, ternary, label_mode, mask_words,
                   emit_mode, nstripes, threads, rows_per_stripe, key_limit=0, row_base=0):
        a = L.HistArgs()
        for k, v in dict(q_bits=q_bits, q_nz=q_nz if ternary else None, g_bits=g_bits,
                         g_nz=g_nz if ternary else None, q_lab=q_lab, g_lab=g_lab, slab_all=slab_all,
                         slab_rel=slab_rel, thresh=thresh, rec_off=rec_off, rec_cap=rec_cap, rec_cnt=rec_cnt,
                         recs=recs, err_flag=err_flag).items():
            setattr(a, k, v.data_ptr() if v is not None else None)
        a.nq, a.nq_pad, a.ndb = nq, nq_pad, ndb
        a.nbit, a.ternary, a.label_mode, a.mask_words, a.emit_mode = nbit, int(ternary), label_mode, mask_words, emit_mode
        a.nstripes, a.threads, a.rows_per_stripe, a.key_limit = nstripes, threads, rows_per_stripe, int(key_limit)
        a.row_base = int(row_base)
        return a

    def slab_totals(self, slab, nstripes, nbins, nq_pad, out):
        L.check(self.lib.ch_slab_totals(self.ws, _ptr(slab), nstripes, nbins, nq_pad, _ptr(out), self._stream()),
                "ch_slab_totals")

    def slab_exscan(self, slab, nstripes, nbins, nq_pad):
        L.check(self.lib.ch_slab_exscan(self.ws, _ptr(slab), nstripes, nbins, nq_pad, self._stream()),
                "ch_slab_exscan")

    def slab_scan(self, slabs, nstripes, nbins, nq_pad):
        """``slabs`` (nslabs, nstripes, nbins, nq_pad): in-place exclusive scan over the stripes; returns the totals
        (nslabs, nbins, nq_pad)"""
        tot = self.empty((slabs.shape[0], nbins, nq_pad), torch.int32)
        L.check(self.lib.ch_slab_scan(self.ws, _ptr(slabs), int(slabs.shape[0]), nstripes, nbins, nq_pad, _ptr(tot),
                                      self._stream()), "ch_slab_scan")
        return tot

    def class_counts(self, g_ids, ndb, rows_per_stripe, nclass, cls):
        L.check(self.lib.ch_class_counts(self.ws, _ptr(g_ids), ndb, rows_per_stripe, nclass, _ptr(cls),
                                         self._stream()), "ch_class_counts")

    # ---- K3 ----
    def scan_bases(self, tot_all, world, rank, nbins, nq, nq_pad, rmax, base0, thresh, total):
        L.check(self.lib.ch_scan_bases(self.ws, _ptr(tot_all), world, rank, nbins, nq, nq_pad, rmax, _ptr(base0),
                                       _ptr(thresh), _ptr(total), self._stream()), "ch_scan_bases")

    def record_caps(self, source, a0, a1, nstripes, nb, nq, nq_pad, min_with_prev, cap, sample_stride=0,
                    replicate=False):
        """``replicate``: ``a0`` holds ONE stripe whose capacities are written to all ``nstripes`` rows of ``cap``"""
        L.check(self.lib.ch_record_caps(self.ws, source, _ptr(a0), _ptr(a1), nstripes, nb, nq, nq_pad,
                                        int(min_with_prev), int(sample_stride), 1 if replicate else nstripes,
                                        _ptr(cap), self._stream()), "ch_record_caps")

    def record_offsets_async(self, cap, nstripes, nq, nq_pad, off, thresh, limit_slots, key_limit, info, status):
        """offsets without a host round trip; ``info`` u32[2] <- (total slots, max thresh), ``status`` |= 4 / 8"""
        L.check(self.lib.ch_record_offsets_async(self.ws, _ptr(cap), nstripes, nq, nq_pad, _ptr(off), _ptr(thresh),
                                                 int(limit_slots), int(key_limit), _ptr(info), _ptr(status),
                                                 self._stream()), "ch_record_offsets_async")

    def scan_bases_pair(self, tot, world, rank, nbins, nq, nq_pad, rmax, need, base0_all, base0_rel, key_max,
                        total_rel, status, bad=None):
        """``tot`` (world, 2, nbins, nq_pad) contiguous; ``bad`` (nq_pad): 1 for every query with < ``need`` items"""
        L.check(self.lib.ch_scan_bases_pair(self.ws, _ptr(tot), world, rank, nbins, nq, nq_pad, int(rmax), int(need),
                                            _ptr(base0_all), _ptr(base0_rel), _ptr(key_max), _ptr(total_rel),
                                            _ptr(status), _ptr(bad), self._stream()), "ch_scan_bases_pair")

    def gather_rows(self, bits, n_src, nbit, stride):
        """every ``stride``-th of the first ``n_src`` rows of a packed plane -> (rows, packed plane with zero pad rows)"""
        n_out = (int(n_src) + stride - 1) // stride
        out = self.empty((self.padded_rows(n_out), bits.shape[1]), torch.int32)
        L.check(self.lib.ch_gather_rows(self.ws, _ptr(bits), int(n_src), int(nbit), int(stride), _ptr(out),
                                        int(out.shape[0]), self._stream()), "ch_gather_rows")
        return n_out, out

    def record_offsets(self, cap, nstripes, nq, nq_pad, off, thresh=None):
        """-> (total record slots, max(thresh[:nq]) or None); ONE host sync for both."""
        total, tmax = C.c_uint64(), C.c_uint32()
        L.check(self.lib.ch_record_offsets(self.ws, _ptr(cap), nstripes, nq, nq_pad, _ptr(off), C.byref(total),
                                           _ptr(thresh), C.byref(tmax), self._stream()), "ch_record_offsets")
        return int(total.value), (int(tmax.value) if thresh is not None else None)

    def check_counts(self, total, nq, need, flags):
        L.check(self.lib.ch_check_counts(self.ws, _ptr(total), nq, need, _ptr(flags), self._stream()),
                "ch_check_counts")

    # ---- K4 ----
    def _final_args(self, f):
        a = L.FinalArgs()
        for k in ("recs", "rec_off", "rec_cnt", "base0_all", "base0_rel", "sbase_all", "sbase_rel", "first_rel",
                  "partial", "cols"):
            v = f.get(k)
            setattr(a, k, v.data_ptr() if v is not None else None)
        a.nq, a.nq_pad, a.nstripes, a.nbins = f["nq"], f["nq_pad"], f["nstripes"], f["nbins"]
        a.remove_first = int(f.get("remove_first", False))
        r_eff, pr_k = list(f.get("r_eff", [])), list(f.get("pr_k", []))
        if len(r_eff) > L.CH_MAX_R or len(pr_k) > L.CH_MAX_PR:
            raise ValueError(f"at most {L.CH_MAX_R} R values and {L.CH_MAX_PR} PRs cut-offs are supported")
        a.nR, a.nPR = len(r_eff), len(pr_k)
        for i, v in enumerate(r_eff):
            a.r_eff[i] = int(v)
        for i, v in enumerate(pr_k):
            a.pr_k[i] = int(v)
        return a

    def finalize_records(self, f):
        L.check(self.lib.ch_finalize_records(self.ws, C.byref(self._final_args(f)), self._stream()),
                "ch_finalize_records")

    def first_relevant(self, f, out):
        L.check(self.lib.ch_first_relevant(self.ws, C.byref(self._final_args(f)), _ptr(out), self._stream()),
                "ch_first_relevant")

    def reduce_means(self, cols, total_rel, first_rel, nq, n_r, pr_k, ap_out=None, flags=None):
        """-> (mAPs, recalls, precisions, status words); ``flags`` (the u32 status block on the device) rides on
        the same host sync.  While an evaluation is being CAPTURED into a CUDA graph (``capture_results`` set) the
        reduction and its copies are only enqueued, and the values of the evaluation that ran just before on the same
        inputs are returned -- the replayed graph's own values are read by ``fetch_results``."""
        n_pr = len(pr_k)
        nfl = int(flags.numel()) if flags is not None else 0
        prk = (C.c_int64 * max(1, n_pr))(*[int(k) for k in pr_k])
        L.check(self.lib.ch_reduce_means_enqueue(self.ws, _ptr(cols), _ptr(total_rel), _ptr(first_rel), nq, n_r, n_pr,
                                                 prk, _ptr(ap_out), _ptr(flags), nfl, self._stream()),
                "ch_reduce_means")
        self.last_result_shape = (n_r, n_pr, nfl)
        if self.capture_results is not None:
            return self.capture_results
        self.last_results = self.fetch_results()
        return self.last_results

    def fetch_results(self):
        """waits for the stream and reads the result words of the last ``reduce_means`` (or graph replay)"""
        n_r, n_pr, nfl = self.last_result_shape
        out = (C.c_double * max(1, n_r + 2 * n_pr))()
        fl = (C.c_uint32 * max(2, nfl))()
        L.check(self.lib.ch_reduce_means_fetch(self.ws, out, n_r + 2 * n_pr, fl, nfl, self._stream()),
                "ch_reduce_means_fetch")
        vals = [float(out[i]) for i in range(n_r + 2 * n_pr)]
        return vals[:n_r], vals[n_r:n_r + n_pr], vals[n_r + n_pr:], [int(fl[i]) for i in range(max(2, nfl))]

    def scatter_ranked(self, f, R, row_offset, ids, keys):
        L.check(self.lib.ch_scatter_ranked(self.ws, C.byref(self._final_args(f)), R, row_offset, _ptr(ids),
                                           _ptr(keys), self._stream()), "ch_scatter_ranked")

    def ap_from_ranked(self, ids, nq, R, q_lab, g_lab, label_mode, mask_words, pr_k, cols):
        prk = (C.c_int64 * max(1, len(pr_k)))(*[int(k) for k in pr_k])
        L.check(self.lib.ch_ap_from_ranked(self.ws, _ptr(ids), nq, R, _ptr(q_lab), _ptr(g_lab), label_mode,
                                           mask_words, len(pr_k), prk, _ptr(cols), self._stream()),
                "ch_ap_from_ranked")

    # ---- small ----
    def hamming_matrix(self, q_bits, q_nz, g_bits, g_nz, nq, ndb, nbit, ternary):
        out = self.empty((nq, ndb), torch.int16)
        L.check(self.lib.ch_hamming_matrix(self.ws, _ptr(q_bits), _ptr(q_nz) if ternary else None, _ptr(g_bits),
                                           _ptr(g_nz) if ternary else None, nq, ndb, nbit, int(ternary), _ptr(out),
                                           self._stream()), "ch_hamming_matrix")
        return out

    def popc_peak(self):
        rate, ms = C.c_double(), C.c_double()
        L.check(self.lib.ch_popc_peak(self.ws, C.byref(rate), C.byref(ms)), "ch_popc_peak")
        return rate.value, ms.value
