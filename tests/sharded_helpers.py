"""numpy/oracle stand-ins for the three kernel entry points of mergerec_b200/merger/sharded.py, so that the distributed
host logic (radix-pass all-reduces, tie scan in rank order, per-rank cut keys, slice all-gather) can run on CPU tensors
with the gloo backend.  TEST INFRASTRUCTURE ONLY -- the product path always uses `CudaKernels`."""
import ctypes as C

import numpy as np
import torch

from mergerec_b200 import _lib
from oracle import oracle as orc


class OracleKernels:
    @staticmethod
    def rows(models):
        return list(models.unbind(0)) if isinstance(models, torch.Tensor) else list(models)

    @staticmethod
    def _bits(base, row, w, k):
        u = (row.numpy() - base.numpy()).astype(np.float32)
        if w is not None:
            u = (u * np.float32(w.reshape(-1)[k].item())).astype(np.float32)
        return (u.view(np.uint32) & np.uint32(0x7FFFFFFF)).astype(np.int64)

    @staticmethod
    def kth_largest_bits(base, rows, k, w):
        out = []
        for i, r in enumerate(rows):
            bits = OracleKernels._bits(base, r, w, i)
            out.append(int(np.partition(bits, bits.size - k)[bits.size - k]))
        return torch.tensor(out, dtype=torch.int64), None

    @staticmethod
    def mag_hist(base, rows, w, lo, shift, hist, above, cand=None, cand_count=None):
        for k, r in enumerate(rows):
            bits = OracleKernels._bits(base, r, w, k)
            l, s = int(lo[k]) & 0xFFFFFFFF, int(shift[k])
            idx = np.nonzero(bits >= l)[0]
            bins = (bits[idx] - l) >> s
            inside = bins < 2048
            hist[k] += torch.from_numpy(np.bincount(bins[inside], minlength=2048).astype(np.int64))
            above[k] += int((~inside).sum())
            if cand is not None and s == 0:
                cap = cand.shape[1]
                order = np.random.default_rng(k).permutation(int(inside.sum()))     # the kernel appends in no particular order
                b, j = bins[inside][order], idx[inside][order]
                n = min(cap, b.size)
                cand[k, :n, 0] = torch.from_numpy(b[:n].astype(np.int32))
                cand[k, :n, 1] = torch.from_numpy(j[:n].astype(np.int32))
                cand_count[k] += b.size

    @staticmethod
    def ties_build(base, rows, cut, mode, w=None, out=None, ldo=0):
        b = np.ascontiguousarray(base.numpy())
        ms = [np.ascontiguousarray(r.numpy()) for r in rows]
        K, d = len(ms), b.size
        c = np.ascontiguousarray(cut.numpy().view(np.uint64))
        if mode == _lib.MR_TIES_VECTORS:
            That = np.empty((K, d), np.float32)
            orc.lib().orc_ties_vectors(orc._ptr(b), orc._ptr_array(ms), C.c_int(K), C.c_int64(d), orc._ptr(c), orc._ptr(That), None, None)
            out[:, :d] = torch.from_numpy(That)
        elif mode == _lib.MR_TIES_TRIMSUM:
            res = np.empty_like(b)
            wv = np.ascontiguousarray(w.numpy().astype(np.float32))
            orc.lib().orc_merge_ties(orc._ptr(b), orc._ptr_array(ms), C.c_int(K), C.c_int64(d), orc._ptr(wv), orc._ptr(c), orc._ptr(res))
            out.copy_(torch.from_numpy(res))
        else:
            raise NotImplementedError(mode)

    @staticmethod
    def merge(base, rows, w, order, src_is_model):
        ws = [float(x) for x in w.reshape(-1).tolist()]
        ms = [r.numpy() for r in rows]
        if order == _lib.MR_ORDER_BASE_FIRST and src_is_model:
            return torch.from_numpy(orc.merge_task_vector(base.numpy(), ms, ws))
        if order == _lib.MR_ORDER_LINEAR:
            return torch.from_numpy(orc.merge_linear(ms, ws))
        raise NotImplementedError(order)

    @staticmethod
    def alloc_rows(K, d, device):
        return torch.empty((K, d), dtype=torch.float32)
