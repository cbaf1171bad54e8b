"""Container writer: wavpackdecoder_b200.wvdemo against the oracle's restatement of WvDemo.Main
(oracle/refdec.c rd_wvdemo, WvDemo.cs:15-174): output file bytes and exit code."""
import numpy as np
import pytest

from _harness import KIND_DSD, KIND_FLOAT, KIND_HYBRID, make_file, oracle_wvdemo

X_RIFF, X_CONFIG, X_MD5_TRAILER = 1, 2, 16

LONG = 9.5  # seconds: >= 100 * SAMPLE_BUFFER_SIZE (409 600) samples, below which WvDemo.cs:113,136 divides by zero

CASES = [
    ("no_stored_header", dict(extras=0, seconds=LONG)),
    ("stored_riff_header", dict(extras=X_RIFF | X_CONFIG, seconds=LONG)),
    ("stored_header_and_trailer", dict(extras=X_RIFF | X_CONFIG | X_MD5_TRAILER, seconds=LONG)),
    ("mono_8bit", dict(extras=0, bits=8, channels=1, seconds=LONG)),
    ("mono_16bit_odd_blocks", dict(extras=0, channels=1, block_samples=10001, seconds=LONG)),
    ("stereo_24bit", dict(extras=X_RIFF, bits=24, seconds=LONG)),
    ("float_ignores_stored_header", dict(kind=KIND_FLOAT, bits=32, extras=X_RIFF | X_CONFIG, seconds=LONG)),
    ("int32", dict(extras=0, bits=32, int32_sent_bits=8, seconds=LONG)),
    ("hybrid", dict(kind=KIND_HYBRID, extras=X_RIFF, terms=[18, 18, 2, 3], seconds=LONG)),
    ("unknown_length", dict(extras=0, unknown_length=1, seconds=LONG)),  # total_samples -1: loop_samples 0 again
    ("short_file_demo_quirk", dict(extras=X_RIFF, seconds=1.0)),
    ("short_mono_24bit_demo_quirk", dict(extras=0, bits=24, channels=1, seconds=0.05)),  # shorter than one chunk
    ("six_channels_without_2ch_flag", dict(extras=0, channels=6, bits=24, sample_rate=48000, block_samples=24000, seconds=1.2)),
    ("dsd_raw", dict(kind=KIND_DSD, dsd_mode=0, seconds=0.2, block_samples=8192)),
    ("dsd_fast", dict(kind=KIND_DSD, dsd_mode=1, seconds=0.2, block_samples=8192)),
    ("dsd_high", dict(kind=KIND_DSD, dsd_mode=3, seconds=0.2, block_samples=8192)),
    ("dsd_short_demo_quirk", dict(kind=KIND_DSD, dsd_mode=1, seconds=0.1, block_samples=8192)),
]


def _files():
    out = []
    for name, kw in CASES:
        _cfg, _src, data = make_file(**kw)
        out.append((name, bytes(data)))
    # a damaged stream: CRC errors make the demo exit 1 with the muted audio on disk
    _cfg, _src, data = make_file(extras=X_RIFF, seconds=LONG)
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 0x5a
    out.append(("damaged_mid_file", bytes(bad)))
    return out


def test_wave_header_matches_the_reference_layout():
    """WvDemo.cs:78-105 + WaveHeader.cs: the synthesised header is what the oracle's WvDemo writes first."""
    from wavpackdecoder_b200 import wvdemo
    _cfg, _src, data = make_file(extras=0, seconds=LONG)
    ref, code = oracle_wvdemo(bytes(data))
    n = int(44100 * LONG)
    assert code == 0
    assert ref[:44] == wvdemo.wave_header(n, 2, 44100, 16, 2)
    assert len(ref) == 44 + n * 4


def test_oracle_wvdemo_passes_stored_header_and_trailer_through():
    _cfg, _src, data = make_file(extras=X_RIFF | X_CONFIG | X_MD5_TRAILER, seconds=LONG)
    ref, code = oracle_wvdemo(bytes(data))
    assert code == 0 and ref[:4] == b"RIFF" and ref[8:12] == b"WAVE"
    short, code2 = oracle_wvdemo(bytes(make_file(extras=0, seconds=1.0)[2]))
    assert code2 == 1  # DivideByZeroException at the progress print, after the first 4096-sample chunk
    assert len(short) == 44 + 4096 * 4


@pytest.mark.gpu
def test_containers_match_wvdemo():
    from wavpackdecoder_b200 import wvdemo
    files = _files()
    got = wvdemo.unpack_files([f for _n, f in files])
    for (name, f), (data, code) in zip(files, got):
        ref, ref_code = oracle_wvdemo(f)
        assert code == ref_code, name
        assert len(data) == len(ref), (name, len(data), len(ref))
        if data != ref:
            a, b = np.frombuffer(data, dtype=np.uint8), np.frombuffer(ref, dtype=np.uint8)
            assert False, (name, int(np.nonzero(a != b)[0][0]))


@pytest.mark.gpu
def test_complete_short_files_when_quirks_are_off():
    from wavpackdecoder_b200 import wvdemo
    _cfg, _src, data = make_file(extras=0, seconds=1.0)
    (out, code), = wvdemo.unpack_files([bytes(data)], reference_quirks=False)
    n = 44100
    assert code == 0 and len(out) == 44 + n * 4 and out[:44] == wvdemo.wave_header(n, 2, 44100, 16, 2)


def test_container_host_logic_with_the_device_code_compiled_for_the_host(monkeypatch):
    """The host side of the container writer (layout, table rebase, header/trailer bytes, exit codes, demo quirks) on a
    box without a GPU: the batch decoder is replaced by tests/emul (the device decode function compiled for the host, a
    test harness) and the result compared with the oracle's WvDemo.  The GPU test above runs the real thing."""
    import ctypes as C
    from _harness import emul
    from wavpackdecoder_b200 import wvdemo

    class EmulDecoder:
        def __init__(self, device=0):
            self.lib = emul()

        def decode(self, in_ptr, in_bytes, descs, nblocks, out_ptr, out_bytes, out_format, mem_flags=0, results=None):
            self.lib.emul_decode(C.c_void_p(in_ptr), descs, C.c_size_t(nblocks), C.c_void_p(out_ptr), C.c_int(out_format), results)

        def close(self):
            pass

    monkeypatch.setattr(wvdemo, "BatchDecoder", EmulDecoder)
    files = _files()
    got = wvdemo.unpack_files([f for _n, f in files])
    for (name, f), (data, code) in zip(files, got):
        ref, ref_code = oracle_wvdemo(f)
        assert code == ref_code, name
        assert data == ref, name


def test_command_line_writes_the_containers(monkeypatch, tmp_path, capsys):
    import ctypes as C
    from _harness import emul
    from wavpackdecoder_b200 import wvdemo

    class EmulDecoder:  # tests/emul stands in for the GPU, as above
        def __init__(self, device=0):
            self.lib = emul()

        def decode(self, in_ptr, in_bytes, descs, nblocks, out_ptr, out_bytes, out_format, mem_flags=0, results=None):
            self.lib.emul_decode(C.c_void_p(in_ptr), descs, C.c_size_t(nblocks), C.c_void_p(out_ptr), C.c_int(out_format), results)

        def close(self):
            pass

    monkeypatch.setattr(wvdemo, "BatchDecoder", EmulDecoder)
    a = bytes(make_file(extras=X_RIFF | X_CONFIG | X_MD5_TRAILER, seconds=LONG)[2])
    b = bytes(make_file(extras=0, bits=24, channels=1, seconds=LONG)[2])
    (tmp_path / "a.wv").write_bytes(a)
    (tmp_path / "b.wv").write_bytes(b)
    code = wvdemo.main([str(tmp_path / "a.wv"), str(tmp_path / "b.wv")])
    assert code == 0
    assert (tmp_path / "a.wav").read_bytes() == oracle_wvdemo(a)[0]
    assert (tmp_path / "b.wav").read_bytes() == oracle_wvdemo(b)[0]
    out = capsys.readouterr().out
    assert "2 channels, 16 bits per sample, 44100 samples/s" in out and "1 channels, 24 bits per sample" in out
    assert wvdemo.main([str(tmp_path / "missing.wv")]) == 1
