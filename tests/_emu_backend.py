"""numpy emulation of the CUDA backend's entry points -- TEST INFRASTRUCTURE ONLY.

It mirrors the *semantics* of every C-ABI call (stripes, {all, relevant} slabs, in-stripe stable
prefixes, 16-byte records, candidate lists, bases, thresholds, capacities, finalisation) on CPU tensors so that
``concepthash_b200.evaluator`` -- the real orchestration code, including its torch.distributed
exchanges -- can be exercised without a GPU (single process and gloo world_size 2).  The product never
imports this module; ``concepthash_b200.hashing`` only ever builds a ``CudaBackend``.
"""
import numpy as np
import torch

CH_LAB_NONE, CH_LAB_ID, CH_LAB_MASK = 0, 1, 2
CH_EMIT_NONE, CH_EMIT_RELEVANT, CH_EMIT_CANDIDATES = 0, 1, 2


def _u32(t):
    return t.numpy().view(np.uint32)


def _unpack(bits_u32, nbit):
    """(n, words) uint32 -> (n, nbit) {0,1}"""
    b = np.ascontiguousarray(bits_u32).view(np.uint8)
    return np.unpackbits(b, axis=1, bitorder="little")[:, :nbit].astype(np.int32)


class EmuBackend:
    name = "emu"
    stripe_align = 1

    def __init__(self, rows_per_stripe=64, threads=32, tensor_cores=False):
        self.rps = rows_per_stripe
        self.threads = threads
        self.launches = 0
        self.tensor_cores = tensor_cores     # emulate the candidate-list path (needs threads % 128 == 0)
        self.paired = True                   # offer the paired plane forms (two gallery rows per plane row)

    # plumbing
    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def full(self, shape, value, dtype):
        return torch.full(shape, value, dtype=dtype)

    def padded_rows(self, n):
        return (n + 63) // 64 * 64 + 64

    def code_words(self, nbit):
        if nbit <= 0 or nbit > 256:
            return 0
        return 1 if nbit <= 32 else 2 if nbit <= 64 else 4 if nbit <= 128 else 8

    def launch_count(self):
        return self.launches

    # K1
    def column_sums(self, codes):
        self.launches += 1
        return codes.detach().cpu().double().sum(0)

    def pack_sign(self, codes, threshold, flags, want_nz=True, col_sub=None):
        n, nbit = codes.shape
        words = self.code_words(nbit)
        if words == 0:
            raise ValueError(f"nbit={nbit} unsupported")
        x = codes.detach().cpu().clone()
        if not x.dtype.is_floating_point:
            x = x.float()
        if col_sub is not None:
            x = x - col_sub.to(x.dtype)[None, :]
        if threshold != 0:
            x[x.abs() < torch.tensor(threshold, dtype=x.dtype)] = 0
        xn = x.double().numpy()
        if np.isnan(xn).any():
            flags[0] |= 2
        if (xn == 0).any():
            flags[0] |= 1
        rows = self.padded_rows(n)

        def pack(mask):
            full = np.zeros((rows, words * 32), dtype=np.uint8)
            full[:n, :nbit] = mask
            return torch.from_numpy(np.packbits(full, axis=1, bitorder="little").view(np.uint32).view(np.int32).copy())

        self.launches += 1
        return pack(xn > 0), (pack(xn != 0) if want_nz else None)

    def pack_labels(self, labels, nolabel, info=None):
        info_out = info
        lab = labels.detach().cpu()
        n = lab.shape[0]
        rows = self.padded_rows(n)
        ids = np.full((rows,), nolabel, dtype=np.uint32)
        info = np.zeros(4, dtype=np.uint32)
        self.launches += 1
        if lab.dim() == 2:
            ncls = lab.shape[1]
            lw = max((ncls + 31) // 32, 1)
            pos = (lab.double().numpy() > 0)
            full = np.zeros((rows, lw * 32), dtype=np.uint8)
            full[:n, :ncls] = pos
            masks = np.packbits(full, axis=1, bitorder="little").view(np.uint32)
            cnt = pos.sum(1)
            has = cnt > 0
            ids[:n][has] = pos[has].argmax(1)
            info[0] = cnt.max() if n else 0
            info[1] = (ids[:n][has].max() + 1) if has.any() else 0
            info[2] = (~has).sum()
            return (torch.from_numpy(ids.view(np.int32)), torch.from_numpy(masks.view(np.int32).copy()),
                    self._info(info, info_out))
        v = lab.double().numpy()
        ok = v >= 0
        ids[:n][ok] = v[ok].astype(np.uint32)
        info[0] = 1 if ok.any() else 0
        info[1] = (ids[:n][ok].max() + 1) if ok.any() else 0
        info[2] = (~ok).sum()
        return torch.from_numpy(ids.view(np.int32)), None, self._info(info, info_out)

    @staticmethod
    def _info(info, info_out):
        t = torch.from_numpy(info.view(np.int32))
        if info_out is None:
            return t
        info_out[0] = max(int(info_out[0]), int(t[0]))
        info_out[1] = max(int(info_out[1]), int(t[1]))
        info_out[2] = int(info_out[2]) + int(t[2])
        return info_out

    # K2
    def geometry(self, nq, ndb, nbit, ternary, label_mode, lw):
        nq_pad = (max(nq, 1) + self.threads - 1) // self.threads * self.threads
        nstripes = max(1, (ndb + self.rps - 1) // self.rps)
        return self.threads, nq_pad, nstripes, self.rps

    def _keys(self, q_bits, q_nz, g_bits, g_nz, nq, ndb, nbit, ternary):
        qs, gs = _unpack(_u32(q_bits)[:nq], nbit), _unpack(_u32(g_bits)[:ndb], nbit)
        if not ternary:
            return (qs[:, None, :] != gs[None, :, :]).sum(-1)
        qz, gz = _unpack(_u32(q_nz)[:nq], nbit), _unpack(_u32(g_nz)[:ndb], nbit)
        both = (qz[:, None, :] & gz[None, :, :])
        dis = ((qs[:, None, :] != gs[None, :, :]) & (both == 1)).sum(-1)
        return nbit - both.sum(-1) + 2 * dis

    def _rel(self, q_lab, g_lab, nq, ndb, label_mode, lw):
        if label_mode == CH_LAB_NONE:
            return np.zeros((nq, ndb), dtype=bool)
        if label_mode == CH_LAB_ID:
            return _u32(q_lab)[:nq, None] == _u32(g_lab)[None, :ndb]
        qm, gm = _u32(q_lab)[:nq, :lw], _u32(g_lab)[:ndb, :lw]
        return ((qm[:, None, :] & gm[None, :, :]) != 0).any(-1)

    def hamming_hist(self, *, q_bits, q_nz, g_bits, g_nz, q_lab, g_lab, slab_all, slab_rel, thresh, rec_off,
                     rec_cap, rec_cnt, recs, err_flag, nq, nq_pad, ndb, nbit, ternary, label_mode, mask_words,
                     emit_mode, nstripes, threads, rows_per_stripe, key_limit=0, row_base=0):
        self.launches += 1
        keys = self._keys(q_bits, q_nz, g_bits, g_nz, nq, ndb, nbit, ternary)
        rel = self._rel(q_lab, g_lab, nq, ndb, label_mode, mask_words)
        sa = _u32(slab_all)
        sr = _u32(slab_rel) if slab_rel is not None else None
        th = _u32(thresh) if thresh is not None else None
        full = (2 * nbit if ternary else nbit) + 1
        if th is not None and 0 < key_limit < full:
            # the kernel indexes the slabs with nbins = key_limit
            assert sa.shape[1] == key_limit, (sa.shape, key_limit)
            # the kernel clamps a threshold beyond the (hinted) key range: record_offsets_async has flagged it
            th = np.minimum(th, np.uint32(key_limit - 1))
        if emit_mode != CH_EMIT_NONE:
            off, cap, cnt, rc = _u32(rec_off), _u32(rec_cap), _u32(rec_cnt), _u32(recs)
        for s in range(nstripes):
            r0, r1 = s * rows_per_stripe, min(ndb, (s + 1) * rows_per_stripe)
            for q in range(nq):
                run_all, run_rel = {}, {}
                n_rec = 0
                for j in range(r0, r1):
                    k = int(keys[q, j])
                    if th is not None and k > th[q]:
                        continue
                    r = bool(rel[q, j])
                    pa, pr = run_all.get(k, 0), run_rel.get(k, 0)
                    run_all[k] = pa + 1
                    if r:
                        run_rel[k] = pr + 1
                    if emit_mode == CH_EMIT_CANDIDATES or (emit_mode == CH_EMIT_RELEVANT and r):
                        if n_rec < cap[s, q]:
                            rc[off[s, q] + n_rec] = (k | (0x80000000 if r else 0), pa, pr, j + row_base)
                        else:
                            _u32(err_flag)[0] |= 1
                        n_rec += 1
                for k, v in run_all.items():
                    sa[s, k, q] += v
                if sr is not None:
                    for k, v in run_rel.items():
                        sr[s, k, q] += v
                if emit_mode != CH_EMIT_NONE:
                    cnt[s, q] = min(n_rec, cap[s, q])

    # K2, tensor-core form (candidate lists).  The emulation keeps the packed bits + thresholds instead of int8 planes.
    def tc_code_bytes(self, nbit, bare=False):
        if not self.tensor_cores or nbit <= 0 or nbit > 256:
            return 0
        return (nbit + (0 if bare else 2 if nbit <= 254 else 4) + 31) // 32 * 32

    def tc_code_bytes_pair(self, nbit, ternary=False):
        # two gallery rows per plane row: thresh - key must fit a signed byte (keys <= 128)
        if not self.tensor_cores or not self.paired or nbit <= 0 or nbit > (64 if ternary else 128):
            return 0
        return (2 * nbit + 5 + 31) // 32 * 32

    def expand_i8(self, bits, nbit, min_rows=0, thresh=None, nq=0, nz=None, bare=False, query=False, pair=False):
        self.launches += 1
        assert not (bare and thresh is not None) and not (bare and pair)
        assert not pair or self.tc_code_bytes_pair(nbit, nz is not None) > 0
        return dict(bits=bits, nz=nz, nbit=nbit, nq=nq, bare=bare, pair=pair,
                    thresh=None if thresh is None else thresh.clone())

    def expand_i8_query_stripes(self, bits, nbit, min_rows, thresh, scut, nstripes, nq, nz=None, stripe0=0):
        self.launches += 1
        assert self.tc_code_bytes_pair(nbit, nz is not None) > 0 and stripe0 == 0
        return dict(bits=bits, nz=nz, nbit=nbit, nq=nq, bare=False, pair=True, thresh=thresh.clone(),
                    scut=scut.clone(), nstripes=nstripes)

    def hamming_select_tc(self, *, q_i8, g_i8, cand, nq, nq_pad, ndb, nbit, nstripes, rows_per_stripe, row_base=0,
                          dense=False, stripe0=0, thresh=None, ternary=False, bad=None, pair=False):
        self.launches += 1
        # both planes with threshold slots (thresholds inside the query plane), or both bare + explicit thresholds
        assert q_i8["bare"] == g_i8["bare"] == (thresh is not None) and not (dense and thresh is not None)
        assert q_i8["pair"] == g_i8["pair"] == bool(pair) and (not pair or rows_per_stripe % self.stripe_align == 0)
        assert ternary == (q_i8["nz"] is not None) or thresh is None
        if thresh is not None:
            q_i8 = dict(q_i8, thresh=thresh)
        assert (q_i8["nz"] is None) == (g_i8["nz"] is None)
        keys = self._keys(q_i8["bits"], q_i8["nz"], g_i8["bits"], g_i8["nz"], nq, ndb, nbit, q_i8["nz"] is not None)
        th = _u32(q_i8["thresh"])
        off, cap, cnt, rows = _u32(cand["off"]), _u32(cand["cap"]), _u32(cand["cnt"]), _u32(cand["rows"])
        for s in range(nstripes):
            r0, r1 = s * rows_per_stripe, min(ndb, (s + 1) * rows_per_stripe)
            for q in range(nq):
                tq = int(th[q])
                if q_i8.get("scut") is not None and s + stripe0 >= int(_u32(q_i8["scut"])[q]):
                    tq -= 1                  # beyond the query's stripe cut: one key less
                js = [j + row_base for j in range(r0, r1) if keys[q, j] <= tq]
                w = min(len(js), int(cap[s + stripe0, q]))
                o = int(off[s + stripe0, q])
                rows[o:o + w] = js[:w]
                cnt[s + stripe0, q] = w
                if len(js) > w:
                    _u32(cand["err"])[0] |= 1
                    if bad is not None:
                        _u32(bad)[q] = 1

    # K3/K4 on candidate lists
    def cand_hist(self, cand, *, q_bits, g_bits, q_lab, g_lab, label_mode, mask_words, tot_all, tot_rel, nq, nq_pad,
                  nstripes, nbins, nbit, stripe0=0, g_plane=None, q_nz=None, g_nz=None):
        self.launches += 1
        off, cnt, rows, key = _u32(cand["off"]), _u32(cand["cnt"]), _u32(cand["rows"]), cand["key"].numpy()
        qs, gs = _unpack(_u32(q_bits)[:nq], nbit), _unpack(_u32(g_bits), nbit)
        tern = q_nz is not None
        if tern:
            qz, gz = _unpack(_u32(q_nz)[:nq], nbit), _unpack(_u32(g_nz), nbit)
        ta = _u32(tot_all)
        tr = _u32(tot_rel) if tot_rel is not None else None
        for q in range(nq):
            for s in range(stripe0, stripe0 + nstripes):       # totals are accumulated (caller zero-initialises)
                o = int(off[s, q])
                for i in range(int(cnt[s, q])):
                    row = int(rows[o + i] & 0x7FFFFFFF)
                    if tern:
                        both = qz[q] & gz[row]
                        k = int(nbit - both.sum() + 2 * ((qs[q] != gs[row]) & (both == 1)).sum())
                    else:
                        k = int((qs[q] != gs[row]).sum())
                    if label_mode == CH_LAB_ID:
                        r = _u32(q_lab)[q] == _u32(g_lab)[row]
                    elif label_mode == CH_LAB_MASK:
                        r = bool((_u32(q_lab)[q, :mask_words] & _u32(g_lab)[row, :mask_words]).any())
                    else:
                        r = False
                    if k < nbins:
                        ta[k, q] += 1
                        if r and tr is not None:
                            tr[k, q] += 1
                    else:
                        _u32(cand["err"])[0] |= 2
                    key[o + i] = k
                    rows[o + i] = row | (0x80000000 if r else 0)

    def cand_caps(self, cand, thresh, list_stripes, rows_per_stripe, nstripes, nq, nq_pad, sample_stride, cap,
                  m=0, scut=None):
        self.launches += 1
        off, cnt, rows, key = _u32(cand["off"]), _u32(cand["cnt"]), _u32(cand["rows"]), cand["key"].numpy()
        th, c = _u32(thresh), _u32(cap)
        c[...] = 0
        raw = np.zeros((nstripes, nq_pad), dtype=np.float32)
        eq = np.zeros((nstripes, nq_pad), dtype=np.float32)
        for ls in range(list_stripes):
            for q in range(nq):
                o = int(off[ls, q])
                for j in range(int(cnt[ls, q])):
                    if key[o + j] <= th[q]:
                        s = min(int(rows[o + j] & 0x7FFFFFFF) // rows_per_stripe, nstripes - 1)
                        (eq if key[o + j] == th[q] else raw)[s, q] += 1
        if scut is not None:
            sc = _u32(scut)
            for q in range(nq_pad):
                cut = nstripes
                if q < nq:
                    have, cut = raw[:, q].sum(), 0
                    while cut < nstripes and have < m:
                        have += eq[cut, q]
                        cut += 1
                    if have < m:
                        cut = nstripes
                sc[q] = cut
                eq[cut:, q] = 0
        raw += eq
        bound = (raw + np.float32(6.0) * np.sqrt(raw + np.float32(1.0)) + np.float32(9.0)) * np.float32(sample_stride)
        c[:, :nq] = bound[:, :nq].astype(np.uint32)

    def cand_rank(self, cand, *, q_bits, g_bits, q_lab, g_lab, label_mode, mask_words, nq, nq_pad, nstripes, nbins, nbit,
                  cols, r_eff=(), pr_k=(), rmax=-1, need=0, status=None, bad=None, g_plane=None, q_nz=None, g_nz=None):
        """the fused kernel = the three steps it replaces, on one rank"""
        tot = torch.zeros((1, 2, nbins, nq_pad), dtype=torch.int32)
        self.cand_hist(cand, q_bits=q_bits, g_bits=g_bits, q_lab=q_lab, g_lab=g_lab, label_mode=label_mode,
                       mask_words=mask_words, tot_all=tot[0, 0], tot_rel=tot[0, 1] if label_mode != CH_LAB_NONE else None,
                       nq=nq, nq_pad=nq_pad, nstripes=nstripes, nbins=nbins, nbit=nbit, g_plane=g_plane, q_nz=q_nz,
                       g_nz=g_nz)
        base_all = torch.zeros((nbins, nq_pad), dtype=torch.int32)
        base_rel = torch.zeros((nbins, nq_pad), dtype=torch.int32)
        key_max = torch.zeros((nq_pad,), dtype=torch.int32)
        self.scan_bases_pair(tot, 1, 0, nbins, nq, nq_pad, rmax, need, base_all, base_rel, key_max, None, status,
                             **({} if bad is None else dict(bad=bad)))
        self.cand_finalize(cand, mode=0, base0_all=base_all, base0_rel=base_rel, nq=nq, nq_pad=nq_pad,
                           nstripes=nstripes, nbins=nbins, cols=cols, r_eff=r_eff, pr_k=pr_k, key_max=key_max)
        self.launches -= 2

    def cand_finalize(self, cand, *, mode, base0_all, base0_rel, nq, nq_pad, nstripes, nbins, remove_first=False,
                      first_rel=None, first_rel_out=None, cols=None, r_eff=(), pr_k=(), ids=None, keys=None, R=0,
                      row_offset=0, key_max=None):
        self.launches += 1
        off, cnt, rows, key = _u32(cand["off"]), _u32(cand["cnt"]), _u32(cand["rows"]), cand["key"].numpy()
        ba = _u32(base0_all)
        br = _u32(base0_rel) if base0_rel is not None else None
        r_eff, pr_k = list(r_eff), list(pr_k)
        if mode == 0:
            cols.numpy()[...] = 0
        for q in range(nq):
            run_all = {k: int(ba[k, q]) for k in range(nbins)}
            run_rel = {k: (int(br[k, q]) if br is not None else 0) for k in range(nbins)}
            for s in range(nstripes):
                o = int(off[s, q])
                for i in range(int(cnt[s, q])):
                    k = int(key[o + i])
                    if k >= nbins or (key_max is not None and k > int(_u32(key_max)[q])):
                        continue                      # flagged / cannot rank below rmax
                    rel = bool(rows[o + i] >> 31)
                    row = int(rows[o + i] & 0x7FFFFFFF)
                    rank, relrank = run_all[k], run_rel[k]
                    run_all[k] += 1
                    if rel:
                        run_rel[k] += 1
                    if mode == 1:
                        if rank == 0 and rel:
                            _u32(first_rel_out)[q] = 1
                        continue
                    if remove_first:
                        if rank == 0:
                            continue
                        rank -= 1
                        relrank -= int(_u32(first_rel)[q]) if first_rel is not None else 0
                    if mode == 2:
                        if rank < R:
                            ids[q, rank] = row_offset + row
                            if keys is not None:
                                keys[q, rank] = k
                        continue
                    if not rel:
                        continue
                    prec = (relrank + 1) / (rank + 1)
                    c = cols.numpy()
                    for j, r in enumerate(r_eff):
                        if rank < r:
                            c[q, 2 * j] += prec
                            c[q, 2 * j + 1] += 1
                    for j, kk in enumerate(pr_k):
                        if rank < kk:
                            c[q, 2 * len(r_eff) + j] += 1

    def slab_totals(self, slab, nstripes, nbins, nq_pad, out):
        self.launches += 1
        _u32(out)[...] = _u32(slab).sum(0, dtype=np.uint32)

    def slab_exscan(self, slab, nstripes, nbins, nq_pad):
        self.launches += 1
        s = _u32(slab)
        c = np.cumsum(s, axis=0, dtype=np.uint32)
        s[1:] = c[:-1]
        s[0] = 0

    def slab_scan(self, slabs, nstripes, nbins, nq_pad):
        tot = torch.zeros((slabs.shape[0], nbins, nq_pad), dtype=torch.int32)
        for i in range(slabs.shape[0]):
            self.slab_totals(slabs[i], nstripes, nbins, nq_pad, tot[i])
            self.slab_exscan(slabs[i], nstripes, nbins, nq_pad)
        self.launches -= 2 * slabs.shape[0] - 1
        return tot

    def class_counts(self, g_ids, ndb, rows_per_stripe, nclass, cls):
        self.launches += 1
        ids, c = _u32(g_ids)[:ndb], _u32(cls)
        for r in range(ndb):
            if ids[r] < nclass:
                c[r // rows_per_stripe, ids[r]] += 1

    # K3
    def scan_bases(self, tot_all, world, rank, nbins, nq, nq_pad, rmax, base0, thresh, total):
        self.launches += 1
        t = _u32(tot_all).astype(np.int64).copy()            # (world, nbins, nq_pad)
        t[:, :, nq:] = 0
        tot = t.sum(0)
        lower = t[:rank].sum(0)
        cum_incl = np.cumsum(tot, axis=0)
        _u32(base0)[...] = (cum_incl - tot + lower).astype(np.uint32)
        if thresh is not None:
            th = np.full(nq_pad, nbins - 1, dtype=np.uint32)
            if rmax >= 0:
                for q in range(nq_pad):
                    hit = np.nonzero(cum_incl[:, q] >= rmax)[0]
                    if len(hit):
                        th[q] = hit[0]
            _u32(thresh)[...] = th
        if total is not None:
            _u32(total)[...] = cum_incl[-1].astype(np.uint32)

    def record_caps(self, source, a0, a1, nstripes, nb, nq, nq_pad, min_with_prev, cap, sample_stride=0,
                    replicate=False):
        self.launches += 1
        c = np.zeros((nstripes, nq_pad), dtype=np.uint32)
        if source in (0, 1):
            s = _u32(a0)
            if replicate:
                s = np.broadcast_to(s[:1], (nstripes,) + s.shape[1:])
            for q in range(nq):
                last = int(_u32(a1)[q]) if source == 0 else nb - 1
                c[:, q] = s[:, :min(last, nb - 1) + 1, q].sum(1)
        else:
            cls, ids = _u32(a0), _u32(a1)
            for q in range(nq):
                if ids[q] < nb:
                    c[:, q] = cls[:, ids[q]]
        if sample_stride > 1:
            k = c.astype(np.float32)
            c = ((k + np.float32(6.0) * np.sqrt(k + np.float32(1.0)) + np.float32(9.0)) * np.float32(sample_stride)).astype(np.uint32)
        if min_with_prev:
            c = np.minimum(c, _u32(cap))
        _u32(cap)[...] = c

    def record_offsets(self, cap, nstripes, nq, nq_pad, off, thresh=None):
        self.launches += 1
        c = _u32(cap).astype(np.int64)
        rowtot = c.sum(0)
        start = np.cumsum(rowtot) - rowtot
        o = start[None, :] + np.cumsum(c, axis=0) - c
        _u32(off)[...] = o.astype(np.uint32)
        return int(rowtot.sum()), (int(_u32(thresh)[:nq].max()) if thresh is not None else None)

    def record_offsets_async(self, cap, nstripes, nq, nq_pad, off, thresh, limit_slots, key_limit, info, status):
        self.launches += 1
        c = _u32(cap)
        o = _u32(off)
        run, total = 0, int(c.astype(np.int64).sum())
        for q in range(nq_pad):
            for s in range(nstripes):
                if run + int(c[s, q]) > limit_slots:
                    c[s, q] = 0
                    o[s, q] = 0
                else:
                    o[s, q] = run
                    run += int(c[s, q])
        tmax = int(_u32(thresh)[:nq].max()) if thresh is not None else 0
        if info is not None:
            _u32(info)[0], _u32(info)[1] = min(total, 0xFFFFFFFF), tmax
        if total > limit_slots:
            _u32(status)[0] |= 4
        if thresh is not None and key_limit and tmax >= key_limit:
            _u32(status)[0] |= 8

    def scan_bases_pair(self, tot, world, rank, nbins, nq, nq_pad, rmax, need, base0_all, base0_rel, key_max,
                        total_rel, status, bad=None):
        found = torch.zeros(nq_pad, dtype=torch.int32)
        self.scan_bases(tot[:, 0].contiguous(), world, rank, nbins, nq, nq_pad, rmax, base0_all, key_max, found)
        if base0_rel is not None:
            self.scan_bases(tot[:, 1].contiguous(), world, rank, nbins, nq, nq_pad, -1, base0_rel, None, total_rel)
        self.launches -= 1 if base0_rel is not None else 0
        if need > 0 and (_u32(found)[:nq] < need).any():
            _u32(status)[0] |= 1
            if bad is not None:
                _u32(bad)[:nq][_u32(found)[:nq] < need] = 1

    def gather_rows(self, bits, n_src, nbit, stride):
        self.launches += 1
        n_out = (n_src + stride - 1) // stride
        out = torch.zeros((self.padded_rows(n_out), bits.shape[1]), dtype=torch.int32)
        out[:n_out] = bits[:n_src][::stride]
        return n_out, out

    def check_counts(self, total, nq, need, flags):
        self.launches += 1
        if (_u32(total)[:nq] < need).any():
            _u32(flags)[0] |= 1

    # K4
    def _ranks(self, f, rec, s, q, need_rel):
        key = int(rec[0] & 0x7FFFFFFF)
        rank = int(_u32(f["base0_all"])[key, q]) + int(_u32(f["sbase_all"])[s, key, q]) + int(rec[1])
        relrank = 0
        if need_rel:
            relrank = int(_u32(f["base0_rel"])[key, q]) + int(_u32(f["sbase_rel"])[s, key, q]) + int(rec[2])
        return bool(rec[0] >> 31), rank, relrank

    def finalize_records(self, f):
        self.launches += 1
        nq, nstripes = f["nq"], f["nstripes"]
        r_eff, pr_k = f["r_eff"], f["pr_k"]
        ncols = 2 * len(r_eff) + len(pr_k)
        cols = f["cols"].numpy()
        cols[...] = 0
        rc, off, cnt = _u32(f["recs"]), _u32(f["rec_off"]), _u32(f["rec_cnt"])
        rf = f.get("remove_first", False)
        first = _u32(f["first_rel"]) if (rf and f.get("first_rel") is not None) else None
        for s in range(nstripes):
            for q in range(nq):
                for k in range(cnt[s, q]):
                    is_rel, rank, relrank = self._ranks(f, rc[off[s, q] + k], s, q, True)
                    if not is_rel:
                        continue
                    if rf:
                        if rank == 0:
                            continue
                        rank -= 1
                        relrank -= int(first[q]) if first is not None else 0
                    prec = (relrank + 1) / (rank + 1)
                    for j, r in enumerate(r_eff):
                        if rank < r:
                            cols[q, 2 * j] += prec
                            cols[q, 2 * j + 1] += 1
                    for j, kk in enumerate(pr_k):
                        if rank < kk:
                            cols[q, 2 * len(r_eff) + j] += 1
        assert ncols <= cols.shape[1] or ncols == 0

    def first_relevant(self, f, out):
        self.launches += 1
        rc, off, cnt = _u32(f["recs"]), _u32(f["rec_off"]), _u32(f["rec_cnt"])
        o = _u32(out)
        for s in range(f["nstripes"]):
            for q in range(f["nq"]):
                for k in range(cnt[s, q]):
                    is_rel, rank, _ = self._ranks(f, rc[off[s, q] + k], s, q, False)
                    if is_rel and rank == 0:
                        o[q] = 1

    def reduce_means(self, cols, total_rel, first_rel, nq, n_r, pr_k, ap_out=None, flags=None):
        self.launches += 1
        c = cols.numpy()
        maps, recalls, precisions = [], [], []
        for j in range(n_r):
            ap = np.where(c[:, 2 * j + 1] > 0, c[:, 2 * j] / np.maximum(c[:, 2 * j + 1], 1), 0.0)
            if ap_out is not None:
                ap_out.numpy()[j] = ap
            maps.append(float(ap.mean()))
        for j, k in enumerate(pr_k):
            hits = c[:, 2 * n_r + j]
            tr = _u32(total_rel)[:nq].astype(np.float64)
            if first_rel is not None:
                tr = tr - _u32(first_rel)[:nq]
            recalls.append(float((hits / np.maximum(tr, 1.0)).mean()))
            precisions.append(float((hits / k).mean()))
        fl = [int(v) for v in _u32(flags)] if flags is not None else [0, 0]
        return maps, recalls, precisions, fl

    def scatter_ranked(self, f, R, row_offset, ids, keys):
        self.launches += 1
        rc, off, cnt = _u32(f["recs"]), _u32(f["rec_off"]), _u32(f["rec_cnt"])
        rf = f.get("remove_first", False)
        for s in range(f["nstripes"]):
            for q in range(f["nq"]):
                for k in range(cnt[s, q]):
                    rec = rc[off[s, q] + k]
                    _, rank, _ = self._ranks(f, rec, s, q, False)
                    if rf:
                        if rank == 0:
                            continue
                        rank -= 1
                    if rank < R:
                        ids[q, rank] = row_offset + int(rec[3])
                        keys[q, rank] = int(rec[0] & 0x7FFFFFFF)

    def ap_from_ranked(self, ids, nq, R, q_lab, g_lab, label_mode, mask_words, pr_k, cols):
        self.launches += 1
        c = cols.numpy()
        for q in range(nq):
            run, s = 0, 0.0
            hits = [0] * len(pr_k)
            for k in range(R):
                j = int(ids[q, k])
                if j < 0:
                    continue
                if label_mode == CH_LAB_ID:
                    r = _u32(q_lab)[q] == _u32(g_lab)[j]
                else:
                    r = bool((_u32(q_lab)[q, :mask_words] & _u32(g_lab)[j, :mask_words]).any())
                if r:
                    s += (run + 1) / (k + 1)
                    run += 1
                    for i, kk in enumerate(pr_k):
                        if k < kk:
                            hits[i] += 1
            c[q, 0], c[q, 1] = s, run
            for i in range(len(pr_k)):
                c[q, 2 + i] = hits[i]

    def hamming_matrix(self, q_bits, q_nz, g_bits, g_nz, nq, ndb, nbit, ternary):
        self.launches += 1
        return torch.from_numpy(self._keys(q_bits, q_nz, g_bits, g_nz, nq, ndb, nbit, ternary).astype(np.int16))
