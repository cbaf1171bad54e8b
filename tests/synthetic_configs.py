"""The synthetic streams whose hashes are committed in tests/golden/synthetic_manifest.json: one per BASELINE.json config
(regenerated from seeds at test time, never stored) plus small DSD fixtures that ARE stored because an independent decoder
(FFmpeg) vouched for them at mint time.  Shared by tools/make_golden_synthetic.py and the tests."""
from _harness import KIND_DSD, KIND_FLOAT, KIND_HYBRID

T16 = [18, 18, 2, 3, -2, 18, 2, 4, 7, 5, 3, 6, 8, -1, 18, 2]

CONFIGS = {
    # BASELINE configs[0]: one 60 s 16-bit stereo 44.1 kHz default-mode file, 120 blocks of 22 050 samples (WvDemo's case)
    "config1_60s_s16_stereo": dict(baseline_config=1, seed=0x5EED0000, seconds=60.0, kw=dict()),
    # configs[1]: a file of the batch workload (10 s)
    "config2_10s_s16_stereo": dict(baseline_config=2, seed=0x5EED0001, seconds=10.0, kw=dict()),
    # configs[2]: 24-bit 48 kHz 5.1, 16 decorrelation terms; the reference decodes FL/FR through OPEN_2CH_MAX
    "config3_24bit_51_16terms": dict(baseline_config=3, seed=0x5EED0002, seconds=2.0, open_flags=8,
                                     kw=dict(bits=24, channels=6, sample_rate=48000, block_samples=24000, terms=T16, deltas=[2] * 16)),
    # configs[3]: float-flagged, INT32 with and without WVX, hybrid lossy
    "config4a_float": dict(baseline_config=4, seed=0x5EED0003, seconds=2.0, kw=dict(kind=KIND_FLOAT, bits=32)),
    "config4b_int32_wvx": dict(baseline_config=4, seed=0x5EED0004, seconds=2.0, kw=dict(bits=32, int32_sent_bits=8)),
    "config4b_int32_nowvx": dict(baseline_config=4, seed=0x5EED0005, seconds=2.0, kw=dict(bits=32, int32_sent_bits=8, int32_wvx=0)),
    "config4c_hybrid_stereo": dict(baseline_config=4, seed=0x5EED0006, seconds=2.0, kw=dict(kind=KIND_HYBRID)),
    "config4c_hybrid_balance": dict(baseline_config=4, seed=0x5EED0007, seconds=2.0, kw=dict(kind=KIND_HYBRID, hybrid_balance=1, terms=[18, 2])),
    "config4c_hybrid_mono": dict(baseline_config=4, seed=0x5EED0008, seconds=2.0, kw=dict(kind=KIND_HYBRID, channels=1, terms=[18, 18, 2, 3])),
    # configs[4]: DSD64 stereo, the three modes
    "config5_dsd64_raw": dict(baseline_config=5, seed=0x5EED0009, seconds=1.0, kw=dict(kind=KIND_DSD, dsd_mode=0, block_samples=22050)),
    "config5_dsd64_fast": dict(baseline_config=5, seed=0x5EED000A, seconds=1.0, kw=dict(kind=KIND_DSD, dsd_mode=1, block_samples=22050)),
    "config5_dsd64_high": dict(baseline_config=5, seed=0x5EED000B, seconds=1.0, kw=dict(kind=KIND_DSD, dsd_mode=3, block_samples=22050)),
}

DSD_FIXTURES = {
    "dsd_raw_stereo": dict(seed=0xD5D0, seconds=0.08, kw=dict(kind=KIND_DSD, dsd_mode=0, block_samples=8192)),
    "dsd_fast_stereo": dict(seed=0xD5D1, seconds=0.08, kw=dict(kind=KIND_DSD, dsd_mode=1, block_samples=8192)),
    "dsd_fast_mono_h5": dict(seed=0xD5D2, seconds=0.08, kw=dict(kind=KIND_DSD, dsd_mode=1, channels=1, dsd_history_bits=5, block_samples=8192)),
    "dsd_high_stereo": dict(seed=0xD5D3, seconds=0.08, kw=dict(kind=KIND_DSD, dsd_mode=3, block_samples=8192)),
    "dsd_high_mono_rate200": dict(seed=0xD5D4, seconds=0.08, kw=dict(kind=KIND_DSD, dsd_mode=3, channels=1, dsd_rate_i=200, block_samples=8192)),
}
