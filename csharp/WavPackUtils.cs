// csharp/WavPackUtils.cs -- drop-in replacement for the decode path of WavPack.WavPackUtils (reference:
// WavPackUtils.cs:36-512) on top of libwvb.so via P/Invoke.  SOURCE ONLY: the build image has no .NET/mono, so this file
// is not compiled or tested here; the Python mirror (wavpackdecoder_b200/wavpack_utils.py) exercises the same C ABI calls
// in the same order and is what the test-suite runs.  Same public names and signatures as the reference so callers such
// as WvDemo.Main (WvDemo.cs:28,119,125) compile unchanged.
using System;
using System.IO;
using System.Runtime.InteropServices;

namespace WavPack
{
    public static class Defines
    {
        public const int SAMPLE_BUFFER_SIZE = 4096; // Defines.cs:18
        public const byte OPEN_2CH_MAX = 0x8;       // Defines.cs:26
    }

    public enum eFileFormat { WAV = 0, W64 = 1, CAF = 2, DFF = 3, DSF = 4, AIF = 5 } // Defines.cs:148-156

    internal static class Native
    {
        const string Lib = "wvb"; // libwvb.so

        [StructLayout(LayoutKind.Sequential, Pack = 8)]
        internal unsafe struct BlockDesc // wvb_block_desc, 144 bytes
        {
            public ulong in_offset, out_offset;
            public uint in_bytes, block_samples, flags;
            public int crc;
            public long block_index;
            public fixed uint sub_off[8];
            public fixed uint sub_len[8];
            public fixed byte int32_info[4];
            public fixed byte float_info[4];
            public uint bflags;
            public ushort version;
            public byte out_channels, out_stride, out_ch_offset, out_bps;
            public ushort smem_words;
            public uint chunk_first, chunk_samples, file_id, gap_before, terms_sig;
        }

        [StructLayout(LayoutKind.Sequential)]
        internal struct BlockResult { public int crc; public uint rflags; public uint mute_from; public int crc_x; }

        [StructLayout(LayoutKind.Sequential, Pack = 8)]
        internal unsafe struct FileInfo // wvb_file_info
        {
            public int status;
            public fixed byte error_message[64];
            public long total_samples, sample_rate, config_flags, channel_mask;
            public int num_channels, reduced_channels, bits_per_sample, bytes_per_sample, float_norm_exp, xmode, version, five,
                       file_format, lossy_blocks;
            public uint dsd_multiplier, first_flags;
            public long header_off, header_len, trailer_off, trailer_len;
            public fixed byte file_extension[16];
            public long num_blocks, indexed_samples;
            public int stopped_early, reserved;
        }

        [DllImport(Lib)] internal static extern int wvb_abi_version();
        [DllImport(Lib)] internal static extern IntPtr wvb_last_error();
        [DllImport(Lib)] internal static extern int wvb_device_count();
        [DllImport(Lib)] internal static extern unsafe int wvb_index(byte* file, UIntPtr len, uint open_flags, uint chunk_samples,
            FileInfo* info, BlockDesc* blocks, UIntPtr cap, UIntPtr* nblocks);
        [DllImport(Lib)] internal static extern unsafe void wvb_rebase(BlockDesc* blocks, UIntPtr n, ulong in_base, ulong out_base, int out_format, uint file_id);
        [DllImport(Lib)] internal static extern int wvb_batch_create(int device, out IntPtr batch);
        [DllImport(Lib)] internal static extern void wvb_batch_destroy(IntPtr batch);
        [DllImport(Lib)] internal static extern unsafe int wvb_batch_decode(IntPtr batch, byte* input, UIntPtr in_bytes, BlockDesc* descs, UIntPtr nblocks,
            void* output, UIntPtr out_bytes, int out_format, uint mem_flags, BlockResult* results);
        [DllImport(Lib)] internal static extern IntPtr wvb_host_alloc(UIntPtr bytes); // pinned host memory
        [DllImport(Lib)] internal static extern void wvb_host_free(IntPtr p);
        // integrity (beyond the reference, which ignores ID_MD5_CHECKSUM): MD5 of decoded ranges on the device, stored digest lookup
        [DllImport(Lib)] internal static extern unsafe int wvb_batch_md5(IntPtr batch, void* device_out, UIntPtr out_bytes, ulong* offsets, ulong* lengths,
            UIntPtr n, byte* digests);
        [DllImport(Lib)] internal static extern unsafe int wvb_stored_md5(byte* file, UIntPtr len, byte* md5);
        internal const int WVB_OUT_INT32 = 0, WVB_OUT_PCM = 1;
        internal const uint WVB_RF_CRC_ERROR = 1;
    }

    public unsafe class WavpackContext // WavpackContext.cs:13-36 (opaque to callers)
    {
        internal byte[] data;
        internal Native.FileInfo info;
        internal string error_message;
        internal long crc_errors, sample_index;
        internal uint open_flags;
        internal int[] decoded;           // whole file, interleaved, right-justified int32
        internal long[] block_ends;
        internal bool[] block_crc_error;
        internal int channels;
    }

    public static unsafe class WavPackUtils
    {
        // WavPackUtils.cs:36 -- the stream is read to the end once: the batch decoder ships the whole file to the GPU.
        public static WavpackContext WavpackOpenFileInput(BinaryReader infile, uint flags = 0)
        {
            var wpc = new WavpackContext();
            using (var ms = new MemoryStream()) { infile.BaseStream.CopyTo(ms); wpc.data = ms.ToArray(); }
            wpc.open_flags = flags;
            UIntPtr n;
            fixed (byte* p = wpc.data)
            fixed (Native.FileInfo* fi = &wpc.info)
                Native.wvb_index(p, (UIntPtr)wpc.data.Length, flags, Defines.SAMPLE_BUFFER_SIZE, fi, null, UIntPtr.Zero, &n);
            if (wpc.info.status != 0)
                fixed (byte* m = wpc.info.error_message) wpc.error_message = Marshal.PtrToStringAnsi((IntPtr)m);
            wpc.channels = wpc.info.reduced_channels != 0 ? wpc.info.reduced_channels : wpc.info.num_channels;
            return wpc;
        }

        static void DecodeAll(WavpackContext wpc, uint chunk)
        {
            UIntPtr n;
            Native.FileInfo info;
            fixed (byte* p = wpc.data)
            {
                Native.wvb_index(p, (UIntPtr)wpc.data.Length, wpc.open_flags, chunk, &info, null, UIntPtr.Zero, &n);
                var descs = new Native.BlockDesc[Math.Max((int)n, 1)];
                var results = new Native.BlockResult[Math.Max((int)n, 1)];
                wpc.decoded = new int[info.indexed_samples * wpc.channels + 16];
                fixed (Native.BlockDesc* d = descs)
                fixed (Native.BlockResult* r = results)
                fixed (int* o = wpc.decoded)
                {
                    Native.wvb_index(p, (UIntPtr)wpc.data.Length, wpc.open_flags, chunk, &info, d, n, &n);
                    Native.wvb_rebase(d, n, 0, 0, Native.WVB_OUT_INT32, 0);
                    IntPtr batch;
                    if (Native.wvb_batch_create(0, out batch) != 0) // no CUDA device: there is no CPU fallback
                        throw new InvalidOperationException(Marshal.PtrToStringAnsi(Native.wvb_last_error()));
                    try
                    {
                        // the slab needs 3 readable bytes after the last block; wpc.data is copied with padding in production code
                        int rc = Native.wvb_batch_decode(batch, p, (UIntPtr)wpc.data.Length, d, n, o,
                            (UIntPtr)(info.indexed_samples * wpc.channels * 4), Native.WVB_OUT_INT32, 0, r);
                        if (rc != 0) throw new InvalidOperationException(Marshal.PtrToStringAnsi(Native.wvb_last_error()));
                    }
                    finally { Native.wvb_batch_destroy(batch); }
                }
                wpc.block_ends = new long[(int)n];
                wpc.block_crc_error = new bool[(int)n];
                for (int i = 0; i < (int)n; i++)
                {
                    wpc.block_ends[i] = (long)(descs[i].out_offset / (ulong)(4 * wpc.channels)) + descs[i].block_samples;
                    wpc.block_crc_error[i] = (results[i].rflags & Native.WVB_RF_CRC_ERROR) != 0;
                }
                wpc.info.lossy_blocks = info.lossy_blocks;
            }
        }

        // WavPackUtils.cs:200
        public static long WavpackUnpackSamples(WavpackContext wpc, int[] buffer, long samples)
        {
            if (wpc.error_message != null) return 0;
            if (wpc.decoded == null) DecodeAll(wpc, (uint)samples);
            long total = (wpc.decoded.Length - 16) / wpc.channels;
            long n = Math.Min(samples, total - wpc.sample_index);
            if (wpc.info.total_samples >= 0 && wpc.sample_index < wpc.info.total_samples)
                n = Math.Min(n, wpc.info.total_samples - wpc.sample_index);
            if (n <= 0) return 0;
            Array.Copy(wpc.decoded, wpc.sample_index * wpc.channels, buffer, 0, n * wpc.channels);
            long ni = wpc.sample_index + n;
            for (int i = 0; i < wpc.block_ends.Length; i++)
                if (wpc.sample_index < wpc.block_ends[i] && wpc.block_ends[i] <= ni && wpc.block_crc_error[i]) wpc.crc_errors++;
            wpc.sample_index = ni;
            return n;
        }

        // WavPackUtils.cs:288 (unchanged semantics; the batch API can also produce packed PCM on the device with WVB_OUT_PCM)
        public static bool WavpackFormatSamples(int[] src, long samcnt, int bps, byte[] pcm, int offset = 0, bool dsd = false)
        {
            if (pcm == null || pcm.Length < samcnt * bps + offset) return false;
            int c = offset;
            for (long i = 0; i < samcnt; i++)
            {
                int t = src[i];
                if (bps == 1) { pcm[c++] = dsd ? (byte)t : (byte)(0xFF & (t + 128)); continue; }
                pcm[c++] = (byte)t; pcm[c++] = (byte)(t >> 8);
                if (bps >= 3) pcm[c++] = (byte)(t >> 16);
                if (bps == 4) pcm[c++] = (byte)(t >> 24);
            }
            return true;
        }

        // getters, WavPackUtils.cs:346-499
        public static long WavpackGetNumSamples(WavpackContext w, bool native = false) => native && w.info.dsd_multiplier > 0 ? w.info.total_samples * 8 : w.info.total_samples;
        public static long WavpackGetSampleIndex(WavpackContext w) => w.sample_index;
        public static long WavpackGetNumErrors(WavpackContext w) => w.crc_errors;
        public static bool WavpackLossy(WavpackContext w) => w.info.lossy_blocks != 0 || (w.info.config_flags & 8) != 0;
        public static long WavpackGetSampleRate(WavpackContext w) => w.info.sample_rate != 0 ? (w.info.dsd_multiplier > 0 ? w.info.dsd_multiplier * w.info.sample_rate * 8 : w.info.sample_rate) : 44100;
        public static int WavpackGetNumChannels(WavpackContext w) => w.info.num_channels != 0 ? w.info.num_channels : 2;
        public static int WavpackGetBitsPerSample(WavpackContext w) => w.info.bits_per_sample != 0 ? (w.info.dsd_multiplier > 0 ? w.info.bits_per_sample / 8 : w.info.bits_per_sample) : 16;
        public static int WavpackGetBytesPerSample(WavpackContext w) => w.info.bytes_per_sample != 0 ? w.info.bytes_per_sample : 2;
        public static int WavpackGetReducedChannels(WavpackContext w) => w.info.reduced_channels != 0 ? w.info.reduced_channels : WavpackGetNumChannels(w);
        public static eFileFormat WavpackGetFileFormat(WavpackContext w) => (eFileFormat)w.info.file_format;
        public static string WavpackGetErrorMessage(WavpackContext w) => w.error_message;
        public static bool WavpackGetIsFive(WavpackContext w) => w.info.five != 0;
        public static short WavpackGetVersion(WavpackContext w) => (short)w.info.version;
        public static bool WavpackGetIsFloat(WavpackContext w) => (w.info.config_flags & 0x80) > 0;
        public static byte[] WavpackGetHeader(WavpackContext w) => Slice(w, w.info.header_off, w.info.header_len);
        public static byte[] WavpackGetTrailer(WavpackContext w) => Slice(w, w.info.trailer_off, w.info.trailer_len);
        static byte[] Slice(WavpackContext w, long off, long len) { if (len < 0) return null; var b = new byte[len]; Array.Copy(w.data, off, b, 0, len); return b; }
        // WavpackGetMode / WavpackGetCompressionLevel / WavpackGetFileExtension: as in wavpackdecoder_b200/wavpack_utils.py

        // WavPackUtils.cs:504-512: seeking is repositioning over the decoded file (block index, SURVEY 8f-1)
        public static bool SetSample(WavpackContext w, long sample) { if (w.info.total_samples >= 0 && sample >= w.info.total_samples) return false; w.sample_index = Math.Max(0, sample); return true; }
        public static bool SetTime(WavpackContext w, long ms) => SetSample(w, ms / 1000 * w.info.sample_rate);
    }
}
