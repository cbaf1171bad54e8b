"""GPU parity proper: the CUDA path through the C ABI (libwvb.so) against the oracle, bit exact."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from _harness import format_samples, make_file, oracle_decode
from cases import DSD_CASES, INEXACT_BY_DESIGN, PCM_CASES, corrupt_cases

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gpu():
    from wavpackdecoder_b200 import _native as N
    lib = N.load()
    assert lib.wvb_device_count() > 0, "no CUDA device visible"
    return N


def _decode(files, flags, chunk, fmt):
    from wavpackdecoder_b200.batch import decode_files
    return decode_files(files, open_flags=flags, chunk_samples=chunk, out_format=fmt)


@pytest.mark.parametrize("name,flags,chunk,kw", PCM_CASES + DSD_CASES, ids=[c[0] for c in PCM_CASES + DSD_CASES])
def test_synthetic_case_matches_oracle(gpu, name, flags, chunk, kw):
    cfg, src, data = make_file(**kw)
    ref, errs, status, info = oracle_decode(data, flags, chunk)
    assert status == 0
    (out, gerrs, ginfo, results), = _decode([data], flags, chunk, gpu.OUT_INT32)
    assert out.size == ref.size
    assert np.array_equal(out, ref)
    assert gerrs == errs
    assert not any(r.rflags & (gpu.RF_INEXACT | gpu.RF_BAD_BLOCK) for r in results)
    (pcm, _, _, _), = _decode([data], flags, chunk, gpu.OUT_PCM)
    assert np.array_equal(pcm, format_samples(ref, info["bytes_per_sample"]))


def test_golden_vectors(gpu):
    with open(os.path.join(GOLD, "manifest.json")) as f:
        manifest = json.load(f)
    names = sorted(manifest)
    for flags in sorted({manifest[n]["open_flags"] for n in names}):
        group = [n for n in names if manifest[n]["open_flags"] == flags]
        files = [open(os.path.join(GOLD, manifest[n]["file"]), "rb").read() for n in group]
        res = _decode(files, flags, 4096, gpu.OUT_INT32)
        for n, (out, errs, info, results) in zip(group, res):
            e = manifest[n]
            assert errs == 0, n
            assert out.size == e["samples"] * e["reduced_channels"], n
            assert hashlib.md5(np.ascontiguousarray(out, dtype="<i4").tobytes()).hexdigest() == e["int32_md5"], n


def test_mixed_batch_one_launch_set(gpu):
    """Many files of different kinds in ONE batch: exercises the planner's grouping and per-file output offsets."""
    files, refs = [], []
    for i, (name, flags, chunk, kw) in enumerate(PCM_CASES):
        if flags != 0 or chunk != 4096:
            continue
        kw = dict(kw)
        kw.setdefault("seconds", 0.3)
        cfg, src, data = make_file(seed=0x5EED0000 + i, **kw)
        ref, errs, status, info = oracle_decode(data, 0, 4096)
        files.append(data)
        refs.append((ref, errs, info))
    res = _decode(files, 0, 4096, gpu.OUT_PCM)
    for (ref, errs, info), (pcm, gerrs, ginfo, results) in zip(refs, res):
        assert gerrs == errs
        assert np.array_equal(pcm, format_samples(ref, info["bytes_per_sample"]))


def test_damaged_streams_follow_the_oracle(gpu):
    """Corrupt / truncated / gapped streams in one batch: muting, CRC error counts and output length as the oracle."""
    cases = corrupt_cases()
    for chunk in sorted({c[3] for c in cases}):
        group = [c for c in cases if c[3] == chunk]
        res = _decode([c[1] for c in group], 0, chunk, gpu.OUT_INT32)
        for (name, data, flags, _), (out, gerrs, info, results) in zip(group, res):
            ref, errs, status, rinfo = oracle_decode(data, flags, chunk)
            assert status == 0, name
            assert out.size == ref.size, name
            if name in INEXACT_BY_DESIGN:  # inherited decoder state: flagged, not reproduced
                assert any(r.rflags & gpu.RF_INEXACT for r in results), name
            else:
                assert gerrs == errs, name
                assert np.array_equal(out, ref), name


def test_dsd_raw_output_format(gpu):
    """WVB_OUT_DSD_RAW == WavpackFormatSamples(..., dsd: true): bytes copied without the +128 of the 8-bit PCM path."""
    cfg, src, data = make_file(kind=3, dsd_mode=3, seconds=0.1, block_samples=8192)
    ref, errs, status, info = oracle_decode(data)
    (raw, _, _, _), = _decode([data], 0, 4096, gpu.OUT_DSD_RAW)
    assert np.array_equal(raw, format_samples(ref, 1, dsd=True))
    (pcm, _, _, _), = _decode([data], 0, 4096, gpu.OUT_PCM)
    assert np.array_equal(pcm, format_samples(ref, 1, dsd=False))


def test_wvdemo_loop_through_the_api_mirror(gpu):
    """WvDemo.cs:110-135 written against the Python mirror: same samples per call, sample index, error count and PCM as the oracle."""
    from _harness import OracleFile
    from wavpackdecoder_b200 import wavpack_utils as W
    cases = corrupt_cases()
    streams = [make_file(seconds=1.3)[2], make_file(seconds=0.7, channels=1, bits=24)[2], cases[0][1], cases[3][1]]
    for data in streams:
        o = OracleFile(data)
        wpc = W.WavpackOpenFileInput(data)
        nch = W.WavpackGetReducedChannels(wpc)
        bps = W.WavpackGetBytesPerSample(wpc)
        buf = np.zeros(4096 * nch, dtype=np.int32)
        pcm = np.zeros(4096 * nch * bps, dtype=np.uint8)
        while True:
            n = W.WavpackUnpackSamples(wpc, buf, 4096)
            rn, rbuf = o.unpack(4096, nch)
            assert n == rn
            if n == 0:
                break
            assert np.array_equal(buf[: n * nch], rbuf[: n * nch])
            assert W.WavpackFormatSamples(buf, n * nch, bps, pcm)
            assert np.array_equal(pcm[: n * nch * bps], format_samples(rbuf[: n * nch], bps))
            assert W.WavpackGetSampleIndex(wpc) == o.lib.rd_get_sample_index(o.ctx)
            assert W.WavpackGetNumErrors(wpc) == o.lib.rd_get_num_errors(o.ctx)
        o.close()


def test_large_mixed_batch_properties(gpu):
    """A batch the oracle would take minutes to decode serially: size-independent checks.  Every block's CRC (written by the
    encoder from the SOURCE samples) must verify on the device, the total sample count must match, and a checksum of
    per-file checksums must equal the one computed from the encoder-side reconstruction."""
    import hashlib
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus
    kinds = [dict(), dict(channels=1), dict(bits=24), dict(bits=8), dict(terms=[17, 2, -1, 5]), dict(bits=32, int32_sent_bits=8),
             dict(joint_stereo=0, terms=[18, 18, 2, 3, -2], deltas=[3, 3, 2, 2, 1]), dict(shift=3)]
    files, expect = [], hashlib.md5()
    total = 0
    for i in range(256):
        cfg, src, data, recon = make_file(seed=0xABC000 + i, seconds=0.5 + 0.25 * (i % 3), want_recon=True, **kinds[i % len(kinds)])
        files.append(data)
        expect.update(hashlib.md5(np.ascontiguousarray(recon, dtype="<i4").tobytes()).digest())
        total += recon.size // cfg.channels
    cp = Corpus.from_files(files, out_format=gpu.OUT_INT32)
    dec = BatchDecoder(0)
    out, results = dec.decode_corpus(cp)
    dec.close()
    assert cp.total_samples == total
    assert all(results[k].rflags == 0 for k in range(cp.nblocks))
    got = hashlib.md5()
    for i in range(cp.nfiles):
        got.update(hashlib.md5(np.ascontiguousarray(cp.file_output(out, i), dtype="<i4").tobytes()).digest())
    assert got.hexdigest() == expect.hexdigest()


def test_fuzzed_batch_on_device(gpu):
    """150 randomly damaged files decoded in ONE device batch (the CUDA build of the decode functions, real planner and kernels)
    versus the oracle; differences only where the block results carry WVB_RF_INEXACT or the index stopped early."""
    import random
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity as F
    from _harness import KIND_DSD
    rng = random.Random(11)
    bases = []
    for i, kw in enumerate(F.BASES):
        kw = dict(kw)
        secs = 0.05 if kw.get("kind") == KIND_DSD else 0.4
        kw.setdefault("block_samples", 5000)
        bases.append(make_file(seed=900 + i, seconds=secs, **kw)[2])
    files, refs = [], []
    while len(files) < 150:
        data = bytearray(rng.choice(bases))
        k = rng.random()
        if k < 0.7:
            for _ in range(rng.choice([1, 1, 2, 5])):
                data[rng.randrange(len(data))] ^= 1 << rng.randrange(8)
        elif k < 0.9:
            data = data[: rng.randrange(40, len(data))]
        else:
            a = rng.randrange(len(data))
            del data[a:a + rng.randrange(1, 2000)]
        data = bytes(data)
        try:
            ref, errs, status, info = oracle_decode(data, 0, 4096)
        except RuntimeError:
            continue  # the reference refuses to open it; open errors are covered by tests/test_api_getters.py
        if status != 0:
            continue
        from wavpackdecoder_b200 import wavpack_utils as W
        if W.WavpackGetErrorMessage(W.WavpackOpenFileInput(data)) is not None:
            continue
        files.append(data)
        refs.append((ref, errs))
    res = _decode(files, 0, 4096, gpu.OUT_INT32)
    exact = 0
    for (ref, errs), (out, gerrs, info, results) in zip(refs, res):
        inexact = any(r.rflags & gpu.RF_INEXACT for r in results) or info.stopped_early
        same = out.size == ref.size and np.array_equal(out, ref) and gerrs == errs
        assert same or inexact
        exact += bool(same)
    assert exact >= 120


@pytest.mark.parametrize("kw", [dict(), dict(bits=24), dict(bits=8, channels=1), dict(channels=1)], ids=["s16", "s24", "m8", "m16"])
def test_packed_pcm_at_unaligned_output_offsets(gpu, kw):
    """A caller-rebased table may put packed PCM at any byte offset of the output slab.  16-bit stereo then leaves its
    one-word-per-frame kernel (wvb_plan.h block_is_fast16) for the packed writer, whose first and last words are written
    bytewise; nothing outside the file's byte range may be touched."""
    import torch
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus
    cfg, src, data = make_file(seconds=0.3, block_samples=1001, **kw)
    ref, errs, status, info = oracle_decode(data, 0, 4096)
    want = format_samples(ref, info["bytes_per_sample"])
    dec = BatchDecoder(0)
    try:
        for shift in (0, 1, 2, 3, 5):
            corpus = Corpus.from_files([bytes(data)], out_format=gpu.OUT_PCM)
            table = np.frombuffer(corpus.descs, dtype=np.uint64).reshape(-1, C.sizeof(gpu.BlockDesc) // 8)
            table[:corpus.nblocks, 1] += shift  # out_offset
            total = corpus.out_bytes + shift + 7
            d_out = torch.full((total + 64,), 0xEE, dtype=torch.uint8, device="cuda")
            results = (gpu.BlockResult * corpus.nblocks)()
            dec.decode(corpus.slab.ctypes.data, corpus.slab.size, corpus.descs, corpus.nblocks, d_out.data_ptr(), total, gpu.OUT_PCM,
                       gpu.OUT_DEVICE, results)
            got = d_out.cpu().numpy()
            assert np.array_equal(got[shift:shift + want.size], want), shift
            assert (got[:shift] == 0xEE).all() and (got[shift + want.size:] == 0xEE).all(), shift
            assert not any(r.rflags for r in results)
    finally:
        dec.close()


@pytest.mark.parametrize("terms", [None, [18, 2, 18, 3, -2], [18, 17]], ids=["stock_in_register", "generic", "ff_fastest_in_register"])
def test_staged_16bit_stereo_output_writes_only_its_own_bytes(gpu, terms):
    """16-bit stereo PCM leaves the device through the shared-memory staging ring (wvb_pcm.cuh Stage16): full 32-byte
    sectors where a block owns them, single words where a sector is shared with the neighbouring block, a gap or the slab
    end.  Files of every length modulo 8 and tiny blocks put block starts at all eight sector phases inside one warp;
    damaged files stop early (fault, short get_words); the bytes between the files' outputs must stay untouched."""
    import torch
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus
    kw = dict() if terms is None else dict(terms=terms, deltas=[2] * len(terms))
    files, rng = [], np.random.default_rng(3)
    for k, n in enumerate([1, 2, 7, 8, 9, 15, 16, 17, 23, 24, 25, 31, 33, 100, 1001, 2049, 4099, 5000, 7777, 12345]):
        bs = [37, 1000, 4096, 22050][k % 4]
        files.append(bytes(make_file(seed=0xA0 + k, nsamples=n, block_samples=bs, **kw)[2]))
    for k in range(6):  # damaged: bit flips in the audio and a truncation
        data = bytearray(make_file(seed=0xB0 + k, nsamples=9000 + 11 * k, block_samples=1500, **kw)[2])
        if k < 5:
            data[int(rng.integers(200, len(data) - 8))] ^= 1 << int(rng.integers(0, 8))
        else:
            del data[len(data) - 700:]
        files.append(bytes(data))
    corpus = Corpus.from_files(files, out_format=gpu.OUT_PCM)
    # spread the files' outputs over the slab with odd gaps (multiples of 4 keep the one-word-per-frame kernel)
    table = np.frombuffer(corpus.descs, dtype=np.uint64).reshape(-1, C.sizeof(gpu.BlockDesc) // 8)
    shifts = np.cumsum(4 * rng.integers(1, 12, size=corpus.nfiles)).astype(np.uint64)
    table[:corpus.nblocks, 1] += np.repeat(shifts, corpus.count.astype(np.int64))
    total = int(corpus.out_bytes + shifts[-1] + 64)
    d_out = torch.full((total + 64,), 0xEE, dtype=torch.uint8, device="cuda")
    results = (gpu.BlockResult * corpus.nblocks)()
    dec = BatchDecoder(0)
    try:
        dec.decode(corpus.slab.ctypes.data, corpus.slab.size, corpus.descs, corpus.nblocks, d_out.data_ptr(), total, gpu.OUT_PCM, gpu.OUT_DEVICE, results)
    finally:
        dec.close()
    got = d_out.cpu().numpy()
    untouched = np.ones(got.size, dtype=bool)
    for i, data in enumerate(files):
        try:
            ref, errs, status, info = oracle_decode(data, 0, 4096)
        except RuntimeError:  # a flip that makes the file unopenable: no blocks, no output
            assert int(corpus.count[i]) == 0
            continue
        want = format_samples(ref, info["bytes_per_sample"])
        lo = int(corpus.file_out_offset[i] + shifts[i])
        f, c = int(corpus.first[i]), int(corpus.count[i])
        if not any(results[k].rflags & gpu.RF_INEXACT for k in range(f, f + c)):
            assert np.array_equal(got[lo:lo + want.size], want), i
        untouched[lo:lo + want.size] = False
    assert (got[untouched] == 0xEE).all()


def test_decode_slab_overlapped_index_matches_the_two_call_path(gpu):
    """wvb_batch_decode_files (index pass overlapped with upload and decode) returns the same table, output bytes and
    block results as wvb_index_many + wvb_batch_decode, for a mixed slab including a damaged file, an unopenable one,
    DSD (whose scratch tables are sized on the way) and enough files for several segments' worth of launches."""
    from wavpackdecoder_b200.batch import BatchDecoder, Corpus, WvbError
    kinds = [dict(), dict(bits=24), dict(channels=1), dict(kind=1), dict(bits=32, int32_sent_bits=8), dict(kind=3, dsd_mode=1, block_samples=4000),
             dict(kind=3, dsd_mode=3, block_samples=4000), dict(terms=[18, 2, 18, 3, -2], deltas=[2] * 5), dict(bits=8)]
    files = []
    for k in range(90):
        kw = dict(kinds[k % len(kinds)])
        secs = 0.02 if kw.get("kind") == 3 else 0.05 + 0.01 * (k % 7)
        files.append(bytearray(make_file(seed=0xD00 + k, seconds=secs, **kw)[2]))
    files[13][len(files[13]) // 2] ^= 0x20
    files[40][0:4] = b"nope"
    two = Corpus.from_files([bytes(f) for f in files], out_format=gpu.OUT_PCM)
    dec = BatchDecoder(0)
    try:
        want, want_res = dec.decode_corpus(two)
        out = np.full(two.out_bytes + 64, 0xEE, dtype=np.uint8)
        one, res = dec.decode_slab(two.slab, two.offsets, two.sizes, out, cap_hint=two.nblocks + 7, out_format=gpu.OUT_PCM, threads=4)
        assert one.nblocks == two.nblocks and one.out_bytes == two.out_bytes
        assert np.array_equal(one.first, two.first) and np.array_equal(one.count, two.count) and np.array_equal(one.file_out_offset, two.file_out_offset)
        assert bytes(one._descs_mem[:two.nblocks * C.sizeof(gpu.BlockDesc)]) == bytes(two._descs_mem[:two.nblocks * C.sizeof(gpu.BlockDesc)])
        for i in range(two.nfiles):
            assert np.array_equal(one.file_output(out, i), two.file_output(want, i)), i
            assert one.infos[i].status == two.infos[i].status and one.infos[i].indexed_samples == two.infos[i].indexed_samples
        for k in range(two.nblocks):
            assert (res[k].crc, res[k].rflags, res[k].mute_from, res[k].crc_x) == (want_res[k].crc, want_res[k].rflags, want_res[k].mute_from, want_res[k].crc_x), k
        # too small a table / output: nothing decoded, the sizes needed come back
        for cap, ocap in ((two.nblocks - 1, out.size), (two.nblocks, two.out_bytes - 17)):  # (out_bytes ends with up to 15 bytes of padding)
            with pytest.raises(WvbError) as e:
                dec.decode_slab(two.slab, two.offsets, two.sizes, out, cap_hint=cap, out_format=gpu.OUT_PCM, out_cap=ocap)
            assert e.value.needed == (two.nblocks, two.out_bytes)
        # files out of slab order are refused
        with pytest.raises(WvbError):
            dec.decode_slab(two.slab, two.offsets[::-1].copy(), two.sizes[::-1].copy(), out, cap_hint=two.nblocks, out_format=gpu.OUT_PCM)
    finally:
        dec.close()
