// csharp/WavPackUtils.cs -- drop-in replacement for the decode path of WavPack.WavPackUtils (reference:
// WavPackUtils.cs:36-512) on top of libwvb.so via P/Invoke.  SOURCE ONLY: the build image has no .NET/mono, so this file
// is not compiled here.  What is checked here instead (tests/test_csharp_shim.py): every public static method of the
// reference's WavPackUtils exists with the same return type and parameter list, every P/Invoke prototype names a symbol
// include/wvb.h declares with the same number of arguments, and the [StructLayout] mirrors below have the field order,
// sizes and offsets the library was compiled with (wvb_abi_layout).  The Python mirror
// (wavpackdecoder_b200/wavpack_utils.py) makes the same C ABI calls in the same order and is what the test-suite runs.
// Same public names and signatures as the reference, so callers such as WvDemo.Main (WvDemo.cs:28-160) compile unchanged.
using System;
using System.IO;
using System.Runtime.InteropServices;

namespace WavPack
{
    public static class Defines
    {
        public const int SAMPLE_BUFFER_SIZE = 4096; // Defines.cs:18
        public const byte OPEN_2CH_MAX = 0x8;       // Defines.cs:26
        // Defines.cs:112-147, the subset WavpackGetMode / WavpackGetCompressionLevel read
        public const int CONFIG_HYBRID_FLAG = 8, CONFIG_FLOAT_DATA = 0x80, CONFIG_FAST_FLAG = 0x200, CONFIG_HIGH_FLAG = 0x800,
                         CONFIG_VERY_HIGH_FLAG = 0x1000, CONFIG_LOSSY_MODE = 0x1000000, CONFIG_EXTRA_MODE = 0x2000000;
        public const int MODE_WVC = 0x1, MODE_LOSSLESS = 0x2, MODE_HYBRID = 0x4, MODE_FLOAT = 0x8, MODE_VALID_TAG = 0x10, MODE_HIGH = 0x20,
                         MODE_FAST = 0x40, MODE_EXTRA = 0x80, MODE_VERY_HIGH = 0x400, MODE_XMODE = 0x7000, MODE_DSD = 0x10000;
    }

    public enum eFileFormat { WAV = 0, W64 = 1, CAF = 2, DFF = 3, DSF = 4, AIF = 5 } // Defines.cs:148-156

    internal static class Native
    {
        const string Lib = "wvb"; // libwvb.so

        [StructLayout(LayoutKind.Sequential, Pack = 8)]
        internal unsafe struct BlockDesc // wvb_block_desc, 160 bytes
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
            public uint chunk_first, chunk_samples, file_id, gap_before, terms_sig, skip_samples, skip_chunk, avg_block_size, checksum_off;
        }

        [StructLayout(LayoutKind.Sequential)]
        internal struct BlockResult // wvb_block_result, 16 bytes
        {
            public int crc;
            public uint rflags, mute_from;
            public int crc_x;
        }

        [StructLayout(LayoutKind.Sequential, Pack = 8)]
        internal unsafe struct FileInfo // wvb_file_info, 224 bytes
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

        [StructLayout(LayoutKind.Sequential, Pack = 8)]
        internal struct SeekState // wvb_seek_state, 40 bytes
        {
            public long hdr_pos, block_index, avg_block_size, file_pos;
            public uint block_samples, ck_size;
        }

        [DllImport(Lib)] internal static extern int wvb_abi_version();
        [DllImport(Lib)] internal static extern IntPtr wvb_abi_layout();
        [DllImport(Lib)] internal static extern IntPtr wvb_last_error();
        [DllImport(Lib)] internal static extern int wvb_device_count();
        [DllImport(Lib)] internal static extern unsafe int wvb_index(byte* file, UIntPtr len, uint open_flags, uint chunk_samples,
            FileInfo* info, BlockDesc* blocks, UIntPtr cap, UIntPtr* nblocks);
        [DllImport(Lib)] internal static extern unsafe int wvb_index_seek(byte* file, UIntPtr len, uint open_flags, SeekState* from, long target,
            uint skip_chunk, uint chunk_samples, UIntPtr max_blocks, FileInfo* info, BlockDesc* blocks, UIntPtr cap, UIntPtr* nblocks,
            long* window_first_sample, long* landed_sample);
        [DllImport(Lib)] internal static extern unsafe void wvb_rebase(BlockDesc* blocks, UIntPtr n, ulong in_base, ulong out_base, int out_format, uint file_id);
        [DllImport(Lib)] internal static extern int wvb_batch_create(int device, out IntPtr batch);
        [DllImport(Lib)] internal static extern void wvb_batch_destroy(IntPtr batch);
        [DllImport(Lib)] internal static extern unsafe int wvb_batch_decode(IntPtr batch, byte* input, UIntPtr in_bytes, BlockDesc* descs, UIntPtr nblocks,
            void* output, UIntPtr out_bytes, int out_format, uint mem_flags, BlockResult* results);
        // index + decode of a slab of files in one call: the index pass overlaps the upload / decode / download (batch hosts)
        [DllImport(Lib)] internal static extern unsafe int wvb_batch_decode_files(IntPtr batch, byte* slab, UIntPtr slab_bytes, ulong* offsets, ulong* sizes,
            UIntPtr nfiles, uint open_flags, uint chunk_samples, int out_format, int threads, FileInfo* infos, BlockDesc* blocks, UIntPtr cap,
            ulong* first, ulong* count, ulong* file_out_offset, UIntPtr* nblocks, ulong* out_bytes, void* output, UIntPtr out_cap, uint mem_flags,
            BlockResult* results);
        [DllImport(Lib)] internal static extern IntPtr wvb_host_alloc(UIntPtr bytes); // pinned host memory
        [DllImport(Lib)] internal static extern void wvb_host_free(IntPtr p);
        // integrity (beyond the reference, which ignores ID_MD5_CHECKSUM): MD5 of decoded ranges on the device, stored digest lookup
        [DllImport(Lib)] internal static extern unsafe int wvb_batch_md5(IntPtr batch, void* device_out, UIntPtr out_bytes, ulong* offsets, ulong* lengths,
            UIntPtr n, byte* digests);
        [DllImport(Lib)] internal static extern unsafe int wvb_stored_md5(byte* file, UIntPtr len, byte* md5);
        // WavPack 5 block checksum (ID_BLOCK_CHECKSUM, which the reference only notes: MetadataUtils.cs:183): 1 matching, 0 wrong, -1 none.
        // wvb_batch_decode checks every block that has one on the device and reports WVB_RF_BLOCK_CHECKSUM in its result flags
        [DllImport(Lib)] internal static extern unsafe int wvb_block_checksum_ok(byte* block, UIntPtr len);
        // container writer, DSD: decoded DSD bytes (DSDIFF layout) re-laid-out for Sony DSF on the device
        [DllImport(Lib)] internal static extern unsafe int wvb_batch_dsd_to_dsf(IntPtr batch, void* device_src, UIntPtr src_bytes, void* device_dst, UIntPtr dst_bytes,
            ulong* src_off, ulong* dst_off, ulong* frames, uint* channels, UIntPtr nfiles);
        internal const int WVB_ABI_VERSION = 3;
        internal const int WVB_OUT_INT32 = 0, WVB_OUT_PCM = 1;
        internal const uint WVB_RF_CRC_ERROR = 1, WVB_RF_BLOCK_CHECKSUM = 32;
        internal const int WVB_E_CAPACITY = -4;
    }

    // One decoded window: the block decoding (re)starts at and the ones after it (wavpack_utils.py: _Table)
    internal unsafe class Window
    {
        internal Native.BlockDesc[] descs;
        internal int nblocks;
        internal long first, landed, nsamples, next_pos;
        internal uint chunk;
        internal int origin;              // 0 open, 1 seek, 2 regrid (call size changed, no seek)
        internal bool has_state;          // origin == 1: the reader state the seek started from
        internal Native.SeekState state;
        internal long target;
        internal uint prev_chunk;
        internal long[] starts, ends;
        internal bool[] crc_error;
        internal int[] decoded;           // interleaved, right-justified int32; null until decoded
        internal int lossy_blocks;
    }

    public unsafe class WavpackContext // WavpackContext.cs:13-36 (opaque to callers)
    {
        internal byte[] data;             // the whole .wv stream, padded with 64 zero bytes (the kernels read whole words ahead)
        internal int data_len;
        internal Native.FileInfo info;
        internal string error_message;
        internal long crc_errors, sample_index;
        internal uint open_flags;
        internal int channels;
        internal Window win, pending;
        internal IntPtr batch = IntPtr.Zero;
        public int lookahead_blocks = 0;  // blocks decoded per device pass from the current position; 0 = to the end of the file
        public int decode_passes = 0;

        ~WavpackContext() { if (batch != IntPtr.Zero) { Native.wvb_batch_destroy(batch); batch = IntPtr.Zero; } }
    }

    public static unsafe class WavPackUtils
    {
        // WavPackUtils.cs:36 -- the stream is read to the end once: the batch decoder ships the whole file to the GPU.
        public static WavpackContext WavpackOpenFileInput(System.IO.BinaryReader infile, uint flags = 0)
        {
            if (Native.wvb_abi_version() != Native.WVB_ABI_VERSION) throw new InvalidOperationException("libwvb ABI version mismatch");
            var wpc = new WavpackContext();
            using (var ms = new MemoryStream())
            {
                infile.BaseStream.CopyTo(ms);
                wpc.data_len = (int)ms.Length;
                wpc.data = new byte[wpc.data_len + 64];
                Array.Copy(ms.GetBuffer(), wpc.data, wpc.data_len);
            }
            wpc.open_flags = flags;
            UIntPtr n;
            fixed (byte* p = wpc.data)
            fixed (Native.FileInfo* fi = &wpc.info)
                Native.wvb_index(p, (UIntPtr)wpc.data_len, flags, Defines.SAMPLE_BUFFER_SIZE, fi, null, UIntPtr.Zero, &n);
            if (wpc.info.status != 0)
                fixed (byte* m = wpc.info.error_message) wpc.error_message = Marshal.PtrToStringAnsi((IntPtr)m);
            wpc.channels = wpc.info.reduced_channels != 0 ? wpc.info.reduced_channels : wpc.info.num_channels;
            return wpc;
        }

        // Index pass for one window (wavpack_utils.py: _make_table).  origin 0: the table WavpackOpenFileInput + sequential reads
        // produce; 1: after the reference's seek() to `target` (state: the reader state it starts from, or none = header hop);
        // 2: the caller went on reading at `target` with another call size.
        static Window MakeTable(WavpackContext wpc, uint chunk, int origin, bool has_state, Native.SeekState state, long target, uint prev_chunk, int max_blocks)
        {
            var t = new Window { chunk = chunk, origin = origin, has_state = has_state, state = state, target = target, prev_chunk = prev_chunk };
            Native.FileInfo info;
            UIntPtr n = UIntPtr.Zero;
            long first = 0, landed = 0;
            fixed (byte* p = wpc.data)
            {
                for (int pass = 0; pass < 2; pass++)
                {
                    int cap = pass == 0 ? 0 : t.nblocks;
                    if (pass == 0 && origin == 0 && max_blocks > 0) { t.nblocks = max_blocks; continue; }
                    if (pass == 1) t.descs = new Native.BlockDesc[Math.Max(t.nblocks, 1)];
                    int rc;
                    fixed (Native.BlockDesc* d = t.descs)
                    {
                        Native.BlockDesc* dp = pass == 0 ? null : d;
                        if (origin == 0)
                        {
                            rc = Native.wvb_index(p, (UIntPtr)wpc.data_len, wpc.open_flags, chunk, &info, dp, (UIntPtr)cap, &n);
                            if (rc == Native.WVB_E_CAPACITY && max_blocks > 0) rc = 0; // only the head of the table was asked for
                        }
                        else
                        {
                            Native.SeekState st = state;
                            rc = Native.wvb_index_seek(p, (UIntPtr)wpc.data_len, wpc.open_flags, (origin == 1 && has_state) ? &st : null, target,
                                origin == 2 ? prev_chunk : 0u, chunk, (UIntPtr)max_blocks, &info, dp, (UIntPtr)cap, &n, &first, &landed);
                        }
                    }
                    if (rc != 0) throw new InvalidOperationException("block index failed: " + rc);
                    t.nblocks = pass == 0 ? (int)n : Math.Min(t.nblocks, (int)n);
                    if (t.nblocks == 0) break;
                }
            }
            t.first = first; t.landed = landed;
            t.nsamples = t.nblocks > 0 ? info.indexed_samples : 0;
            t.lossy_blocks = info.lossy_blocks;
            t.starts = new long[t.nblocks]; t.ends = new long[t.nblocks];
            for (int i = 0; i < t.nblocks; i++)
            {
                t.starts[i] = first + (long)t.descs[i].out_offset;
                t.ends[i] = t.starts[i] + t.descs[i].block_samples;
            }
            if (origin == 0 && max_blocks > 0 && t.nblocks > 0) t.nsamples = t.ends[t.nblocks - 1] - first;
            t.next_pos = origin != 0 ? landed : 0;
            return t;
        }

        // What the reference's reader holds when seek() starts (wavpack_utils.py: _reader_state): the header it read last and
        // the file position behind that block.
        static bool ReaderState(WavpackContext wpc, out Native.SeekState st)
        {
            st = new Native.SeekState();
            Window tab = wpc.pending ?? wpc.win;
            if (tab == null) tab = MakeTable(wpc, Defines.SAMPLE_BUFFER_SIZE, 0, false, st, 0, 0, 1);
            if (tab.nblocks == 0) return false;
            int j = 0;
            for (int i = 0; i < tab.nblocks; i++) if (tab.starts[i] < wpc.sample_index) j = i;
            // (in_offset is file-relative: windows are rebased with in_base 0)
            Native.BlockDesc b = tab.descs[j];
            st.hdr_pos = (long)b.in_offset; st.block_index = b.block_index; st.avg_block_size = b.avg_block_size;
            st.file_pos = (long)b.in_offset + b.in_bytes; st.block_samples = b.block_samples; st.ck_size = b.in_bytes >= 8 ? b.in_bytes - 8 : 0;
            return true;
        }

        // One device pass over a window's blocks, all in parallel (wavpack_utils.py: _decode_table)
        static void DecodeTable(WavpackContext wpc, Window t)
        {
            t.decoded = new int[t.nsamples * wpc.channels + 16];
            t.crc_error = new bool[t.nblocks];
            if (t.nblocks > 0)
            {
                // out_offset (samples) must survive the rebase for starts/ends: they were taken in MakeTable
                var results = new Native.BlockResult[t.nblocks];
                if (wpc.batch == IntPtr.Zero && Native.wvb_batch_create(0, out wpc.batch) != 0) // no CUDA device: there is no CPU fallback
                    throw new InvalidOperationException(Marshal.PtrToStringAnsi(Native.wvb_last_error()));
                fixed (byte* p = wpc.data)
                fixed (Native.BlockDesc* d = t.descs)
                fixed (Native.BlockResult* r = results)
                fixed (int* o = t.decoded)
                {
                    Native.wvb_rebase(d, (UIntPtr)t.nblocks, 0, 0, Native.WVB_OUT_INT32, 0);
                    int rc = Native.wvb_batch_decode(wpc.batch, p, (UIntPtr)wpc.data.Length, d, (UIntPtr)t.nblocks, o,
                        (UIntPtr)(t.nsamples * wpc.channels * 4), Native.WVB_OUT_INT32, 0, r);
                    if (rc != 0) throw new InvalidOperationException(Marshal.PtrToStringAnsi(Native.wvb_last_error()));
                }
                wpc.decode_passes++;
                for (int i = 0; i < t.nblocks; i++) t.crc_error[i] = (results[i].rflags & Native.WVB_RF_CRC_ERROR) != 0;
                wpc.info.lossy_blocks |= t.lossy_blocks;
            }
            wpc.win = t;
            wpc.pending = null;
        }

        static void NewWindow(WavpackContext wpc, uint chunk)
        {
            Window win = wpc.win, pend = wpc.pending, t;
            var none = new Native.SeekState();
            if (pend != null) // after SetSample / SetTime
                t = (pend.chunk == chunk && wpc.lookahead_blocks == 0) ? pend
                    : MakeTable(wpc, chunk, 1, pend.has_state, pend.state, pend.target, 0, wpc.lookahead_blocks);
            else if (win == null && wpc.sample_index == 0)
                t = MakeTable(wpc, chunk, 0, false, none, 0, 0, wpc.lookahead_blocks);
            else if (wpc.info.total_samples < 0) // unknown length: no seek (WavPackUtils.cs:527); re-read from the start with the new call size
                t = MakeTable(wpc, chunk, 0, false, none, 0, 0, 0);
            else
            {
                uint prev = (win != null && win.next_pos == wpc.sample_index) ? win.chunk : 0;
                t = prev != 0 ? MakeTable(wpc, chunk, 2, false, none, wpc.sample_index, prev, wpc.lookahead_blocks)
                              : MakeTable(wpc, chunk, 1, false, none, wpc.sample_index, 0, wpc.lookahead_blocks);
            }
            DecodeTable(wpc, t);
            // a seek that starts by decoding an earlier block counts that block's CRC verdict when the skip loop finishes it
            if (t.origin == 1)
                for (int i = 0; i < t.nblocks; i++)
                    if (t.crc_error[i] && t.ends[i] <= wpc.sample_index) wpc.crc_errors++;
            t.next_pos = wpc.sample_index;
        }

        // WavPackUtils.cs:200
        public static long WavpackUnpackSamples(WavpackContext wpc, int[] buffer, long samples)
        {
            if (wpc.error_message != null || samples <= 0) return 0;
            Window win = wpc.win;
            if (win == null || wpc.pending != null || win.chunk != (uint)samples || win.next_pos != wpc.sample_index)
            {
                NewWindow(wpc, (uint)samples);
                win = wpc.win;
            }
            long done = 0;
            int nch = wpc.channels;
            while (done < samples)
            {
                long n = Math.Min(samples - done, win.first + win.nsamples - wpc.sample_index);
                if (wpc.info.total_samples >= 0 && wpc.sample_index < wpc.info.total_samples)
                    n = Math.Min(n, wpc.info.total_samples - wpc.sample_index); // the call returns at total_samples (WavPackUtils.cs:277)
                if (n <= 0)
                {
                    if (wpc.lookahead_blocks > 0 && win.nblocks >= wpc.lookahead_blocks) // a full window: the stream may go on
                    {
                        NewWindow(wpc, (uint)samples);
                        win = wpc.win;
                        if (win.nsamples > 0 && wpc.sample_index < win.first + win.nsamples) continue;
                    }
                    break;
                }
                Array.Copy(win.decoded, (wpc.sample_index - win.first) * nch, buffer, done * nch, n * nch);
                long ni = wpc.sample_index + n;
                // crc_errors becomes visible when the block's last sample has been handed out (WavPackUtils.cs:273-275)
                for (int i = 0; i < win.nblocks; i++)
                    if (win.crc_error[i] && win.ends[i] > wpc.sample_index && win.ends[i] <= ni) wpc.crc_errors++;
                wpc.sample_index = ni;
                win.next_pos = ni;
                done += n;
                if (ni == wpc.info.total_samples) break;
            }
            return done;
        }

        // WavPackUtils.cs:288 (unchanged semantics; the batch API can also produce packed PCM on the device with WVB_OUT_PCM)
        public static bool WavpackFormatSamples(int[] src, long samcnt, int bps, byte[] pcm_buffer, int offset = 0, bool dsd = false)
        {
            if (pcm_buffer == null || pcm_buffer.Length < samcnt * bps + offset) return false;
            int c = offset;
            for (long i = 0; i < samcnt; i++)
            {
                int t = src[i];
                if (bps == 1) { pcm_buffer[c++] = dsd ? (byte)t : (byte)(0xFF & (t + 128)); continue; }
                pcm_buffer[c++] = (byte)t; pcm_buffer[c++] = (byte)(t >> 8);
                if (bps >= 3) pcm_buffer[c++] = (byte)(t >> 16);
                if (bps == 4) pcm_buffer[c++] = (byte)(t >> 24);
            }
            return true;
        }

        // WavPackUtils.cs:133-167
        public static int WavpackGetMode(WavpackContext wpc)
        {
            long f = wpc.info.config_flags;
            int mode = 0;
            if ((f & Defines.CONFIG_HYBRID_FLAG) != 0) mode |= Defines.MODE_HYBRID;
            else if ((f & Defines.CONFIG_LOSSY_MODE) == 0) mode |= Defines.MODE_LOSSLESS;
            if (wpc.info.lossy_blocks != 0) mode &= ~Defines.MODE_LOSSLESS;
            if ((f & Defines.CONFIG_FLOAT_DATA) != 0) mode |= Defines.MODE_FLOAT;
            if ((f & Defines.CONFIG_HIGH_FLAG) != 0)
            {
                mode |= Defines.MODE_HIGH;
                if ((f & Defines.CONFIG_VERY_HIGH_FLAG) > 0 || wpc.info.version < 0x405) mode |= Defines.MODE_VERY_HIGH;
            }
            if ((f & Defines.CONFIG_FAST_FLAG) != 0) mode |= Defines.MODE_FAST;
            if ((f & Defines.CONFIG_EXTRA_MODE) != 0) mode |= Defines.MODE_EXTRA | ((wpc.info.xmode << 12) & Defines.MODE_XMODE);
            if (wpc.info.dsd_multiplier > 0) mode |= Defines.MODE_DSD;
            return mode;
        }

        // WavPackUtils.cs:169-187
        public static string WavpackGetCompressionLevel(WavpackContext wpc)
        {
            string result = null;
            int mode = WavpackGetMode(wpc);
            if ((mode & Defines.MODE_FAST) > 0) result = "Fast";
            else if ((mode & Defines.MODE_VERY_HIGH) > 0) result = "Very High";
            else if ((mode & Defines.MODE_HIGH) > 0) result = "High";
            if ((mode & Defines.MODE_EXTRA) > 0)
                result = (result ?? "Default") + ", Extra-" + ((mode & Defines.MODE_XMODE) >> 12);
            return result;
        }

        // getters, WavPackUtils.cs:346-499
        public static long WavpackGetNumSamples(WavpackContext wpc, bool native = false) { return native && wpc.info.dsd_multiplier > 0 ? wpc.info.total_samples * 8 : wpc.info.total_samples; }
        public static long WavpackGetSampleIndex(WavpackContext wpc) { return wpc.sample_index; }
        public static long WavpackGetNumErrors(WavpackContext wpc) { return wpc.crc_errors; }
        public static bool WavpackLossy(WavpackContext wpc) { return wpc.info.lossy_blocks != 0 || (wpc.info.config_flags & Defines.CONFIG_HYBRID_FLAG) != 0; }
        public static long WavpackGetSampleRate(WavpackContext wpc)
        {
            if (wpc.info.sample_rate == 0) return 44100;
            return wpc.info.dsd_multiplier > 0 ? wpc.info.dsd_multiplier * wpc.info.sample_rate * 8 : wpc.info.sample_rate;
        }
        public static int WavpackGetNumChannels(WavpackContext wpc) { return wpc.info.num_channels != 0 ? wpc.info.num_channels : 2; }
        public static int WavpackGetBitsPerSample(WavpackContext wpc)
        {
            if (wpc.info.bits_per_sample == 0) return 16;
            return wpc.info.dsd_multiplier > 0 ? wpc.info.bits_per_sample / 8 : wpc.info.bits_per_sample;
        }
        public static int WavpackGetBytesPerSample(WavpackContext wpc) { return wpc.info.bytes_per_sample != 0 ? wpc.info.bytes_per_sample : 2; }
        public static int WavpackGetReducedChannels(WavpackContext wpc) { return wpc.info.reduced_channels != 0 ? wpc.info.reduced_channels : WavpackGetNumChannels(wpc); }
        public static eFileFormat WavpackGetFileFormat(WavpackContext wpc) { return (eFileFormat)wpc.info.file_format; }
        public static string WavpackGetFileExtension(WavpackContext wpc) // WavPackUtils.cs:463-469
        {
            string e;
            fixed (byte* p = wpc.info.file_extension) e = Marshal.PtrToStringAnsi((IntPtr)p);
            return string.IsNullOrEmpty(e) ? "wav" : e;
        }
        public static string WavpackGetErrorMessage(WavpackContext wpc) { return wpc.error_message; }
        public static byte[] WavpackGetHeader(WavpackContext wpc) { return Slice(wpc, wpc.info.header_off, wpc.info.header_len); }
        public static byte[] WavpackGetTrailer(WavpackContext wpc) { return Slice(wpc, wpc.info.trailer_off, wpc.info.trailer_len); }
        public static bool WavpackGetIsFive(WavpackContext wpc) { return wpc.info.five != 0; }
        public static short WavpackGetVersion(WavpackContext wpc) { return (short)wpc.info.version; }
        public static bool WavpackGetIsFloat(WavpackContext wpc) { return (wpc.info.config_flags & Defines.CONFIG_FLOAT_DATA) > 0; }
        static byte[] Slice(WavpackContext wpc, long off, long len) { if (len < 0) return null; var b = new byte[len]; Array.Copy(wpc.data, off, b, 0, len); return b; }

        // WavPackUtils.cs:504-507
        public static bool SetTime(WavpackContext wpc, long milliseconds)
        {
            return SetSample(wpc, milliseconds / 1000 * wpc.info.sample_rate);
        }

        // WavPackUtils.cs:509-594: the index pass replays the reference's probe sequence on the headers (host work, no
        // decoding); the next WavpackUnpackSamples decodes from the block it ends on, the discarded head included.
        public static bool SetSample(WavpackContext wpc, long sample)
        {
            if (sample >= wpc.info.total_samples || wpc.error_message != null) return false;
            Native.SeekState st;
            bool has = ReaderState(wpc, out st);
            Window t = MakeTable(wpc, wpc.win != null ? wpc.win.chunk : (uint)Defines.SAMPLE_BUFFER_SIZE, 1, has, st, sample, 0, wpc.lookahead_blocks);
            if (t.nblocks == 0) return false;
            wpc.pending = t;
            wpc.sample_index = t.landed;
            return true;
        }
    }
}
