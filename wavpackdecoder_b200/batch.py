"""Batch front end over the C ABI: index many in-memory .wv files, decode all their blocks in one
device pass, hand back per-file PCM / int32 and per-file error counts.

This is plumbing only (buffers, offsets, ctypes); all decode work happens in libwvb.so on the GPU.
"""
import ctypes as C

import numpy as np

from . import _native as N


class WvbError(RuntimeError):
    pass


def _check(lib, rc, what):
    if rc != N.OK:
        msg = lib.wvb_last_error()
        raise WvbError("%s failed: %d %s" % (what, rc, msg.decode() if msg else ""))


class Corpus:
    """A set of .wv files packed into one contiguous host slab plus its block table.

    slab: uint8 numpy array (ideally pinned), offsets/sizes: per-file byte ranges.
    """

    def __init__(self, slab, offsets, sizes, open_flags=0, chunk_samples=4096, out_format=N.OUT_PCM, threads=0, cap_hint=0, index=True):
        lib = N.load()
        self.lib = lib
        self.slab = slab
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.sizes = np.ascontiguousarray(sizes, dtype=np.uint64)
        self.nfiles = int(self.offsets.size)
        self.out_format = out_format
        self.open_flags = open_flags
        self.chunk_samples = chunk_samples
        self.infos = (N.FileInfo * max(self.nfiles, 1))()
        self.first = np.zeros(self.nfiles, dtype=np.uint64)
        self.count = np.zeros(self.nfiles, dtype=np.uint64)
        self.file_out_offset = np.zeros(self.nfiles, dtype=np.uint64)
        self.nblocks = 0
        self.out_bytes = 0
        if not index:  # BatchDecoder.decode_slab fills the table while it decodes
            return
        nblocks = C.c_size_t()
        out_bytes = C.c_uint64()
        args = (slab.ctypes.data, self.offsets.ctypes.data, self.sizes.ctypes.data, self.nfiles, open_flags, chunk_samples,
                out_format, threads, self.infos)
        tail = (self.first.ctypes.data, self.count.ctypes.data, self.file_out_offset.ctypes.data, C.byref(nblocks), C.byref(out_bytes))

        def table(cap):
            # the table is filled by the library: backing it with uninitialised numpy memory skips ctypes' zero fill (10 ms at 200 000 blocks)
            self._descs_mem = np.empty(max(cap, 1) * C.sizeof(N.BlockDesc), dtype=np.uint8)
            self.descs = (N.BlockDesc * max(cap, 1)).from_buffer(self._descs_mem)
            return lib.wvb_index_many(*args, self.descs, cap, *tail)

        # cap_hint (e.g. the block count of the previous, similar batch) saves the counting walk; a hint that turns out too
        # small costs one more walk with the exact size
        rc = table(int(cap_hint)) if cap_hint else N.E_CAPACITY
        if rc == N.E_CAPACITY:
            if not cap_hint:
                _check(lib, lib.wvb_index_many(*args, None, 0, *tail), "wvb_index_many(count)")
            rc = table(nblocks.value)
        _check(lib, rc, "wvb_index_many")
        self.nblocks = nblocks.value
        self.out_bytes = int(out_bytes.value)

    @classmethod
    def from_files(cls, files, **kw):
        offs, sizes, pos = [], [], 0
        for f in files:
            offs.append(pos)
            sizes.append(len(f))
            pos += (len(f) + 63) & ~63
        slab = np.zeros(pos + 64, dtype=np.uint8)
        for f, o in zip(files, offs):
            slab[o:o + len(f)] = np.frombuffer(f, dtype=np.uint8)
        return cls(slab, offs, sizes, **kw)

    @property
    def total_samples(self):
        return sum(int(self.infos[i].indexed_samples) for i in range(self.nfiles))

    def file_channels(self, i):
        info = self.infos[i]
        if self.open_flags & N.OPEN_ALL_CHANNELS:
            return info.num_channels
        return info.reduced_channels or info.num_channels

    def file_output(self, out, i):
        """Slice of the output slab holding file i (uint8 view for PCM, int32 view for INT32)."""
        info = self.infos[i]
        unit = 4 if self.out_format == N.OUT_INT32 else info.bytes_per_sample
        nbytes = int(info.indexed_samples) * unit * self.file_channels(i)
        o = int(self.file_out_offset[i])
        v = out[o:o + nbytes]
        return v.view(np.int32) if self.out_format == N.OUT_INT32 else v


class BatchDecoder:
    """One wvb_batch (one CUDA device / stream)."""

    def __init__(self, device=0):
        self.lib = N.load()
        h = C.c_void_p()
        _check(self.lib, self.lib.wvb_batch_create(device, C.byref(h)), "wvb_batch_create")
        self.h = h
        self.device = device

    def close(self):
        if self.h:
            self.lib.wvb_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, in_ptr, in_bytes, descs, nblocks, out_ptr, out_bytes, out_format, mem_flags=0, results=None):
        rptr = None
        if results is not None:
            rptr = results if isinstance(results, int) else C.addressof(results)
        rc = self.lib.wvb_batch_decode(self.h, in_ptr, in_bytes, descs, nblocks, out_ptr, out_bytes, out_format, mem_flags, rptr)
        _check(self.lib, rc, "wvb_batch_decode")

    def prepare(self, descs, nblocks, out_format):
        _check(self.lib, self.lib.wvb_batch_prepare(self.h, descs, nblocks, out_format), "wvb_batch_prepare")

    def wait(self):
        _check(self.lib, self.lib.wvb_batch_wait(self.h), "wvb_batch_wait")

    def timing(self):
        k, h, d, n = C.c_float(), C.c_float(), C.c_float(), C.c_int()
        _check(self.lib, self.lib.wvb_batch_timing(self.h, C.byref(k), C.byref(h), C.byref(d), C.byref(n)), "wvb_batch_timing")
        return dict(kernel_ms=k.value, h2d_ms=h.value, d2h_ms=d.value, launches=n.value)

    def md5_ranges(self, offsets, lengths, out_bytes, device_out=None):
        """MD5 of byte ranges of the decoded output, computed on the device (wvb_batch_md5).  device_out: device pointer of
        a WVB_OUT_DEVICE decode, or None for the batch's own copy of the last host-buffer decode.  Returns (n, 16) uint8."""
        offs = np.ascontiguousarray(offsets, dtype=np.uint64)
        lens = np.ascontiguousarray(lengths, dtype=np.uint64)
        dig = np.zeros((offs.size, 16), dtype=np.uint8)
        rc = self.lib.wvb_batch_md5(self.h, device_out, out_bytes, offs.ctypes.data, lens.ctypes.data, offs.size, dig.ctypes.data)
        _check(self.lib, rc, "wvb_batch_md5")
        return dig

    def decode_slab(self, slab, offsets, sizes, out, cap_hint, open_flags=0, chunk_samples=4096, out_format=N.OUT_PCM, threads=0,
                    mem_flags=0, out_ptr=None, out_cap=None):
        """Index and decode a host slab of files in ONE library call (wvb_batch_decode_files): the index pass overlaps the
        upload and decode of the files already indexed.  offsets must ascend.  `out`: a uint8 numpy array (or, with
        N.OUT_DEVICE in mem_flags, pass out_ptr / out_cap of a device buffer).  cap_hint: block-table capacity (the block
        count of a previous, similar batch; a table or output that turns out too small raises WvbError with the sizes needed
        in .needed).  Returns (Corpus with the table filled in, results array)."""
        corpus = Corpus(slab, offsets, sizes, open_flags=open_flags, chunk_samples=chunk_samples, out_format=out_format, index=False)
        cap = max(int(cap_hint), 1)
        corpus._descs_mem = np.empty(cap * C.sizeof(N.BlockDesc), dtype=np.uint8)
        corpus.descs = (N.BlockDesc * cap).from_buffer(corpus._descs_mem)
        results = (N.BlockResult * cap)()
        nblocks, out_bytes = C.c_size_t(), C.c_uint64()
        optr = out.ctypes.data if out_ptr is None else out_ptr
        ocap = int(out.size) if out_cap is None else int(out_cap)
        rc = self.lib.wvb_batch_decode_files(self.h, slab.ctypes.data, slab.size, corpus.offsets.ctypes.data, corpus.sizes.ctypes.data, corpus.nfiles,
                                             open_flags, chunk_samples, out_format, threads, corpus.infos, corpus.descs, cap,
                                             corpus.first.ctypes.data, corpus.count.ctypes.data, corpus.file_out_offset.ctypes.data,
                                             C.byref(nblocks), C.byref(out_bytes), optr, ocap, mem_flags, C.addressof(results))
        corpus.nblocks, corpus.out_bytes = int(nblocks.value), int(out_bytes.value)
        if rc == N.E_CAPACITY:
            e = WvbError("wvb_batch_decode_files: table or output too small (need %d blocks, %d bytes)" % (corpus.nblocks, corpus.out_bytes))
            e.needed = (corpus.nblocks, corpus.out_bytes)
            raise e
        _check(self.lib, rc, "wvb_batch_decode_files")
        return corpus, results

    def decode_corpus(self, corpus, out=None):
        """Host-buffer decode of a whole Corpus.  Returns (out uint8 array, results array)."""
        if out is None:
            out = np.zeros(corpus.out_bytes + 64, dtype=np.uint8)
        results = (N.BlockResult * max(corpus.nblocks, 1))()
        self.decode(corpus.slab.ctypes.data, corpus.slab.size, corpus.descs, corpus.nblocks, out.ctypes.data, corpus.out_bytes,
                    corpus.out_format, 0, results)
        return out, results


def _file_results(corpus, out, results):
    t = N.result_table(results, corpus.nblocks)
    crc_err = np.concatenate([[0], np.cumsum((t["rflags"] & N.RF_CRC_ERROR) != 0)])
    res = []
    for i in range(corpus.nfiles):
        f, c = int(corpus.first[i]), int(corpus.count[i])
        res.append((corpus.file_output(out, i).copy(), int(crc_err[f + c] - crc_err[f]), corpus.infos[i], [results[k] for k in range(f, f + c)]))
    return res


def decode_files(files, open_flags=0, chunk_samples=4096, out_format=N.OUT_INT32, device=0, devices=None):
    """Decode a list of .wv byte strings; returns a list of (numpy array, crc_errors, info, block results) in file order.

    devices: a list of CUDA device ordinals shards ONE batch by file over several GPUs of the box: the files are
    indexed once, cut into contiguous ranges of equal decode cost (sharding.shard_contiguous_by_cost over
    sum(block_samples x passes)), and every device decodes its range on its own host thread and wvb_batch; shards share
    nothing, the only gather is the per-file results on the host.  No collective is involved."""
    if devices is None or len(devices) <= 1:
        corpus = Corpus.from_files(files, open_flags=open_flags, chunk_samples=chunk_samples, out_format=out_format)
        dec = BatchDecoder(devices[0] if devices else device)
        try:
            out, results = dec.decode_corpus(corpus)
        finally:
            dec.close()
        return _file_results(corpus, out, results)
    import threading
    from .sharding import file_costs, shard_contiguous_by_cost
    whole = Corpus.from_files(files, open_flags=open_flags, chunk_samples=chunk_samples, out_format=out_format)
    ranges = shard_contiguous_by_cost(file_costs(whole), len(devices))
    parts, errors = [None] * len(devices), []

    def work(k):
        lo, hi = ranges[k]
        if hi <= lo:
            parts[k] = []
            return
        try:
            # the shard's files are adjacent in the slab: a view of it, offsets rebased, is its own corpus (no copy)
            base = int(whole.offsets[lo])
            end = int(whole.offsets[hi - 1]) + int(whole.sizes[hi - 1])
            sub = Corpus(whole.slab[base:min(whole.slab.size, end + 64)], whole.offsets[lo:hi] - np.uint64(base), whole.sizes[lo:hi],
                         open_flags=open_flags, chunk_samples=chunk_samples, out_format=out_format)
            dec = BatchDecoder(devices[k])
            try:
                out, results = dec.decode_corpus(sub)
            finally:
                dec.close()
            parts[k] = _file_results(sub, out, results)
        except Exception as e:  # surfaced on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [r for p in parts for r in p]


def stored_md5(data):
    """The MD5 a .wv file stores for its source audio (ID_MD5_CHECKSUM), or None."""
    lib = N.load()
    buf = np.frombuffer(data, dtype=np.uint8)
    md5 = np.zeros(16, dtype=np.uint8)
    return md5.tobytes() if lib.wvb_stored_md5(buf.ctypes.data, buf.size, md5.ctypes.data) else None


def verify_files(files, device=0):
    """Decode a list of .wv byte strings on the GPU and check each file against its stored MD5 without bringing the PCM
    back: the output slab stays in device memory, only 16 bytes per file and the per-block results return.
    Returns a list of dicts: md5 (hex of the decoded PCM), stored (hex or None), match (True/False/None when the file
    stores no MD5), crc_errors, block_checksum_errors (blocks whose WavPack 5 ID_BLOCK_CHECKSUM does not match their bytes;
    None when the file carries no such checksums), error (open error message or None).
    The stored digest covers the source file's audio bytes, so `match` is meaningful for lossless integer PCM; float
    sources (decoded to 24-bit integers here, as by the reference), hybrid-lossy and DSD files (stored digest over the DSD
    bytes; compare with an OUT_DSD_RAW decode instead) legitimately differ."""
    import torch
    corpus = Corpus.from_files(files, open_flags=0, chunk_samples=4096, out_format=N.OUT_PCM)
    dev = torch.device("cuda", device)
    d_out = torch.empty(corpus.out_bytes + 64, dtype=torch.uint8, device=dev)
    results = (N.BlockResult * max(corpus.nblocks, 1))()
    dec = BatchDecoder(device)
    try:
        if corpus.nblocks:
            dec.decode(corpus.slab.ctypes.data, corpus.slab.size, corpus.descs, corpus.nblocks, d_out.data_ptr(), corpus.out_bytes,
                       N.OUT_PCM, N.OUT_DEVICE, results)
        lens = np.array([int(corpus.infos[i].indexed_samples) * corpus.file_channels(i) * int(corpus.infos[i].bytes_per_sample)
                         for i in range(corpus.nfiles)], dtype=np.uint64)
        dig = dec.md5_ranges(corpus.file_out_offset, lens, corpus.out_bytes, d_out.data_ptr())
    finally:
        dec.close()
    res = []
    table = N.desc_table(corpus.descs, corpus.nblocks) if corpus.nblocks else None
    for i in range(corpus.nfiles):
        info = corpus.infos[i]
        msg = bytes(info.error_message).split(b"\0", 1)[0].decode() or None
        f, c = int(corpus.first[i]), int(corpus.count[i])
        st = stored_md5(files[i])
        got = dig[i].tobytes()
        has_ck = bool(c) and bool((table["bflags"][f:f + c] & N.BF_BLOCK_CHECKSUM).any())
        res.append(dict(md5=got.hex(), stored=st.hex() if st else None, match=None if (st is None or msg) else st == got,
                        crc_errors=sum(1 for k in range(f, f + c) if results[k].rflags & N.RF_CRC_ERROR),
                        block_checksum_errors=sum(1 for k in range(f, f + c) if results[k].rflags & N.RF_BLOCK_CHECKSUM) if has_ck else None,
                        error=msg))
    return res
