"""WvDemo (WvDemo.cs:15-174) on the batch decoder: .wv files in, the bytes the demo writes to <name>.<ext> out.

SURVEY.md section 8f row 3 (container writer): the stored RIFF/alt header is passed through (or the 44-byte WAV header
of WvDemo.cs:78-105 is synthesised), the PCM is decoded by the GPU *in place* between header and trailer -- every
file's blocks are rebased so that the kernels write straight into the finished container -- and the stored trailer is
appended.  Host work is the index pass and a few dozen header bytes per file; no PCM is touched on the host.

This is plumbing over libwvb.so; like the rest of the package it has no CPU decode path.
"""
import struct

import numpy as np

from . import _native as N
from . import containers
from . import wavpack_utils as W
from .batch import BatchDecoder, Corpus

SAMPLE_BUFFER_SIZE = W.SAMPLE_BUFFER_SIZE  # Defines.cs:18: WvDemo unpacks in chunks of this many samples


def wave_header(total_samples, num_channels, sample_rate, bits, byteps):
    """WvDemo.cs:78-105 with RiffChunkHeader.cs / ChunkHeader.cs / WaveHeader.cs: 12 + 8 + 16 + 8 bytes, fields
    truncated to their C# types (uint / ushort casts)."""
    block_align = byteps * num_channels
    data_bytes = (total_samples * block_align) & 0xffffffff
    riff_size = ((data_bytes + 2 * 8 + 16) & 0xffffffff) + 4 & 0xffffffff  # RiffChunkHeader.cs:16 adds the 4 of "WAVE"
    return (b"RIFF" + struct.pack("<I", riff_size) + b"WAVE" + b"fmt " + struct.pack("<I", 16) +
            struct.pack("<HHIIHH", 1, num_channels & 0xffff, sample_rate & 0xffffffff, (sample_rate * block_align) & 0xffffffff,
                        block_align & 0xffff, bits & 0xffff) +
            b"data" + struct.pack("<I", data_bytes))


def _context(corpus, i):
    """A WavpackContext over file i of an indexed corpus, for the getters of wavpack_utils (no second index pass)."""
    wpc = W.WavpackContext()
    o, n = int(corpus.offsets[i]), int(corpus.sizes[i])
    wpc.data = corpus.slab[o:o + n]
    wpc.info = corpus.infos[i]
    msg = bytes(wpc.info.error_message).split(b"\0", 1)[0]
    wpc.error_message = msg.decode() if msg else None
    return wpc


def native_container(wpc, total_samples, nch, byteps):
    """(header, trailer) of the container WavpackGetFileFormat names, for a file that stores no header of its own.
    An extension over the reference, which writes RIFF/WAVE whatever the format (WvDemo.cs:76-105)."""
    fmt = W.WavpackGetFileFormat(wpc)
    rate, bits = W.WavpackGetSampleRate(wpc), W.WavpackGetBitsPerSample(wpc)
    dsd = bool(wpc.info.first_flags & 0x80000000)  # DSD_FLAG of the first audio block
    if fmt == containers.WP_FORMAT_WAV and not dsd:
        return wave_header(total_samples, nch, rate, bits, byteps), b""
    if fmt == containers.WP_FORMAT_W64 and not dsd:
        return (containers.w64_header(total_samples, nch, rate, bits, byteps, bool(W.WavpackGetIsFloat(wpc))),
                containers.w64_trailer(total_samples, nch, byteps))
    if fmt == containers.WP_FORMAT_CAF and not dsd:
        return containers.caf_header(total_samples, nch, rate, bits, byteps, bool(W.WavpackGetIsFloat(wpc))), b""
    if fmt == containers.WP_FORMAT_DFF and dsd:
        return containers.dff_header(total_samples, nch, rate), containers.dff_trailer(total_samples, nch)  # (the getter already reports the one-bit rate)
    if fmt == containers.WP_FORMAT_DSF and dsd:
        return containers.dsf_header(total_samples, nch, rate), DSF_RELAYOUT  # the data follow from the device re-layout pass
    raise NotImplementedError("no header synthesis for file format %d with %s audio" % (fmt, "DSD" if dsd else "PCM"))


DSF_RELAYOUT = object()  # marker returned in place of a trailer: the file's DSD bytes go through wvb_batch_dsd_to_dsf


def _dsf_relayout(dec, out, jobs):
    """jobs: [(offset of the file's raw DSD bytes in `out`, byte-times, channels)].  Uploads those ranges, re-lays them out on
    the device and returns one uint8 array per job (ceil(frames / 4096) * 4096 * channels bytes)."""
    import ctypes as C
    import torch
    dev = torch.device("cuda", dec.device)
    src_off, dst_off, spos, dpos = [], [], 0, 0
    for _o, frames, ch in jobs:
        src_off.append(spos); dst_off.append(dpos)
        spos += (frames * ch + 15) & ~15
        dpos += containers.dsf_data_bytes(frames, ch)
    src = np.zeros(max(spos, 16), dtype=np.uint8)
    for (o, frames, ch), so in zip(jobs, src_off):
        src[so:so + frames * ch] = out[o:o + frames * ch]
    d_src = torch.from_numpy(src).to(dev)
    d_dst = torch.empty(max(dpos, 16), dtype=torch.uint8, device=dev)
    a_src, a_dst = np.array(src_off, dtype=np.uint64), np.array(dst_off, dtype=np.uint64)
    a_fr, a_ch = np.array([j[1] for j in jobs], dtype=np.uint64), np.array([j[2] for j in jobs], dtype=np.uint32)
    rc = dec.lib.wvb_batch_dsd_to_dsf(dec.h, d_src.data_ptr(), d_src.numel(), d_dst.data_ptr(), d_dst.numel(), a_src.ctypes.data, a_dst.ctypes.data,
                                      a_fr.ctypes.data, a_ch.ctypes.data, len(jobs))
    if rc != N.OK:
        raise RuntimeError("wvb_batch_dsd_to_dsf failed: %d %s" % (rc, (dec.lib.wvb_last_error() or b"").decode()))
    host = d_dst.cpu().numpy()
    return [host[do:do + containers.dsf_data_bytes(fr, ch)] for (_o, fr, ch), do in zip(jobs, dst_off)]


def unpack_files(files, device=0, reference_quirks=True, container="reference"):
    """Decode a list of .wv byte strings the way WvDemo.Main does.  Returns a list of (output file bytes, exit code).

    container: "reference" writes what the demo writes (stored header, else RIFF/WAVE; DSD as offset-binary bytes);
    "native" is an extension: a stored header always passes through, otherwise the header of the format the file names is
    synthesised (WAV, W64, CAF, DFF, DSF: containers.py), DSD bytes stay raw, and the demo's short-file quirk does not apply.

    reference_quirks: files with fewer than 100 * SAMPLE_BUFFER_SIZE (409 600) samples, or of unknown length, make the reference demo throw
    DivideByZeroException at its progress print (`total_unpacked_samples % loop_samples`, WvDemo.cs:113,136) after the
    first chunk has been written: it leaves header + first chunk on disk and exits 1.  True reproduces that; False
    writes the complete file."""
    if container not in ("reference", "native"):
        raise ValueError("container must be 'reference' or 'native'")
    native = container == "native"
    if native:
        reference_quirks = False
    out_format = N.OUT_DSD_RAW if native else N.OUT_PCM
    corpus = Corpus.from_files(files, open_flags=0, chunk_samples=SAMPLE_BUFFER_SIZE, out_format=out_format)
    nfiles = corpus.nfiles
    ctxs = [_context(corpus, i) for i in range(nfiles)]
    heads, tails, pcm_bytes, ok = [], [], [], []
    dsf_files = {}  # file index -> (byte-times, channels): DSD data to be re-laid-out for the DSF container
    for i, wpc in enumerate(ctxs):
        if W.WavpackGetErrorMessage(wpc):  # WvDemo.cs:41-46: no output file at all
            heads.append(b""); tails.append(b""); pcm_bytes.append(0); ok.append(False)
            continue
        nch = W.WavpackGetReducedChannels(wpc)
        byteps = W.WavpackGetBytesPerSample(wpc)
        header = W.WavpackGetHeader(wpc)
        trailer = W.WavpackGetTrailer(wpc)  # WvDemo.cs:143-145
        pad = b""
        if header is not None and (native or not W.WavpackGetIsFloat(wpc)):  # WvDemo.cs:76-77
            heads.append(header)
        elif native:
            h, pad = native_container(wpc, int(wpc.info.indexed_samples), nch, byteps)
            heads.append(h)
            if pad is DSF_RELAYOUT:
                dsf_files[i] = (int(wpc.info.indexed_samples), nch)
                pad = b""
        else:
            heads.append(wave_header(W.WavpackGetNumSamples(wpc, True), nch, W.WavpackGetSampleRate(wpc),
                                     W.WavpackGetBitsPerSample(wpc), byteps))
        tails.append(pad + (trailer if trailer is not None else b""))
        pcm_bytes.append(int(wpc.info.indexed_samples) * nch * byteps)
        ok.append(True)

    # container layout: every file's PCM starts 64-byte aligned (the 16-bit stereo kernel stores aligned words), its header
    # immediately before and its trailer immediately after
    pcm_start = np.zeros(nfiles, dtype=np.uint64)
    cursor = 0
    for i in range(nfiles):
        start = (cursor + len(heads[i]) + 63) & ~63
        pcm_start[i] = start
        cursor = start + pcm_bytes[i] + len(tails[i])
    total = cursor
    if corpus.nblocks:
        delta = (pcm_start - corpus.file_out_offset).astype(np.uint64)  # modular: a negative shift wraps correctly
        table = np.frombuffer(corpus.descs, dtype=np.uint64).reshape(-1, C_DESC_WORDS)
        table[:corpus.nblocks, 1] += np.repeat(delta, corpus.count.astype(np.int64))  # out_offset is the second field

    out = np.zeros(total + 64, dtype=np.uint8)
    results = (N.BlockResult * max(corpus.nblocks, 1))()
    dsf_data = {}
    if corpus.nblocks:
        dec = BatchDecoder(device)
        try:
            dec.decode(corpus.slab.ctypes.data, corpus.slab.size, corpus.descs, corpus.nblocks, out.ctypes.data, total, out_format, 0, results)
            if dsf_files:
                order = sorted(dsf_files)
                arrays = _dsf_relayout(dec, out, [(int(pcm_start[i]),) + dsf_files[i] for i in order])
                dsf_data = dict(zip(order, arrays))
        finally:
            dec.close()

    res = []
    for i, wpc in enumerate(ctxs):
        if not ok[i]:
            res.append((b"", 1))
            continue
        s = int(pcm_start[i])
        h, t = heads[i], tails[i]
        out[s - len(h):s] = np.frombuffer(h, dtype=np.uint8)
        out[s + pcm_bytes[i]:s + pcm_bytes[i] + len(t)] = np.frombuffer(t, dtype=np.uint8)
        f, c = int(corpus.first[i]), int(corpus.count[i])
        crc_errors = sum(1 for k in range(f, f + c) if results[k].rflags & N.RF_CRC_ERROR)
        total_native = W.WavpackGetNumSamples(wpc, True)
        loop_samples = int(int(total_native / 100) / SAMPLE_BUFFER_SIZE) * SAMPLE_BUFFER_SIZE  # WvDemo.cs:113, C# division truncates (-1: unknown length)
        unpacked = int(wpc.info.indexed_samples)
        if reference_quirks and loop_samples == 0:
            block_align = W.WavpackGetReducedChannels(wpc) * W.WavpackGetBytesPerSample(wpc)
            first_chunk = min(unpacked, SAMPLE_BUFFER_SIZE) * block_align
            res.append((out[s - len(h):s + first_chunk].tobytes(), 1))
            continue
        num_samples = W.WavpackGetNumSamples(wpc)
        code = 1 if (num_samples != -1 and unpacked != num_samples) or crc_errors > 0 else 0  # WvDemo.cs:157-169
        if i in dsf_data:
            res.append((h + dsf_data[i].tobytes() + t, code))
            continue
        res.append((out[s - len(h):s + pcm_bytes[i] + len(t)].tobytes(), code))
    return res


C_DESC_WORDS = 160 // 8


def main(argv=None):
    """`python -m wavpackdecoder_b200.wvdemo a.wv [b.wv ...]`: WvDemo's command line over a batch (WvDemo.cs:15-174).
    Every input is decoded in one device pass and written next to it with the extension the file recommends
    (WavpackGetFileExtension, "wav" unless the file says otherwise).  Prints the summary lines of WvDemo.cs:57-69 per
    file; the exit code is 1 if any file's would be (the demo's per-file codes, including its short-file quirk)."""
    import os
    import sys
    args = list(sys.argv[1:] if argv is None else argv)
    if not args:
        args = ["input.wv"]  # WvDemo.cs:22
    blobs, names = [], []
    for path in args:
        try:
            with open(path, "rb") as f:
                blobs.append(f.read())
            names.append(path)
        except OSError:
            print("Input file '%s' not found" % path, file=sys.stderr)  # WvDemo.cs:30-39
            return 1
    results = unpack_files(blobs)
    corpus = Corpus.from_files(blobs, open_flags=0, chunk_samples=SAMPLE_BUFFER_SIZE, out_format=N.OUT_PCM)  # getters only
    worst = 0
    for i, (path, (data, code)) in enumerate(zip(names, results)):
        wpc = _context(corpus, i)
        err = W.WavpackGetErrorMessage(wpc)
        if err:
            print("Error: " + err, file=sys.stderr)  # WvDemo.cs:41-46
            worst = 1
            continue
        version = W.WavpackGetVersion(wpc)
        print("The WavPack %s (%d.%d) file '%s' has:" % ("5" if W.WavpackGetIsFive(wpc) else "4", version >> 8, version & 0xff, os.path.basename(path)))
        print("%d channels, %d bits per sample, %d samples/s, %d total samples, %s decoding" % (
            W.WavpackGetReducedChannels(wpc), W.WavpackGetBitsPerSample(wpc), W.WavpackGetSampleRate(wpc),
            W.WavpackGetNumSamples(wpc, True), "Lossy" if W.WavpackLossy(wpc) else "Lossless"))
        out_path = os.path.splitext(path)[0] + "." + W.WavpackGetFileExtension(wpc)
        with open(out_path, "wb") as f:
            f.write(data)
        if code:
            print("%s: exit code %d (sample count mismatch, CRC errors, or the demo's short-file exception)" % (path, code), file=sys.stderr)
        worst = max(worst, code)
    return worst


if __name__ == "__main__":
    raise SystemExit(main())
