//! flac-codec's public writer/reader types, backed by the B200 frame engine.
//!
//! Drop-in for the hot path of `flac_codec::encode::{FlacByteWriter, FlacSampleWriter, FlacChannelWriter, FlacStreamWriter}`
//! (src/encode.rs:103-1290) and `flac_codec::decode::{FlacByteReader, FlacSampleReader, FlacChannelReader, FlacStreamReader,
//! FlacSampleIterator, verify, verify_reader}` (src/decode.rs:103-1309): the crate's constructor signatures (no engine
//! argument: a per-thread engine is created on first use), the same `write`/`finalize`/`read`/`fill_buf`/`consume`/`seek`
//! signatures, the same `Error` ordinals (the C ABI returns the 1-based ordinal of `flac_codec::Error`).  All frame work
//! happens in `libflacb200.so`; this file only moves bytes between the caller's `W: Write + Seek` / `R: Read` and the
//! handles.  Readers never hold the file: bytes are fed to the handle in chunks as its decode windows ask for them, and a
//! seekable source is repositioned when the handle asks (FLACB200_NEED_SEEK).
//!
//! NOT BUILT in the development image (no cargo/rustc there) -- kept thin and mechanical on purpose; the same handle
//! protocol, call for call, is exercised end to end through ctypes by `flac_codec_b200/stream.py` in
//! tests/test_gpu_stream.py and tests/test_gpu_batch.py.
mod ffi;

use std::cell::RefCell;
use std::io::{BufRead, Read, Seek, SeekFrom, Write};
use std::marker::PhantomData;
use std::ptr;

/// `flac_codec::Error` as surfaced by the engine: the ordinal (see `flacb200_strerror`) or an I/O error.
#[derive(Debug)]
pub enum Error {
    Io(std::io::Error),
    /// 1-based ordinal of the matching `flac_codec::Error` variant (src/lib.rs:57-193); negative: CUDA / engine
    Codec(i32),
}

impl From<std::io::Error> for Error {
    fn from(e: std::io::Error) -> Self {
        Error::Io(e)
    }
}

impl From<Error> for std::io::Error {
    // flac_codec maps every non-Io error to InvalidData (src/lib.rs:303-311)
    fn from(e: Error) -> Self {
        match e {
            Error::Io(e) => e,
            Error::Codec(c) => std::io::Error::new(std::io::ErrorKind::InvalidData, format!("flac error {c}")),
        }
    }
}

fn ck(rc: i32) -> Result<(), Error> {
    match rc {
        0 => Ok(()),
        1 => Err(Error::Io(std::io::ErrorKind::UnexpectedEof.into())), // Error::Io: the stream ended inside a frame
        c => Err(Error::Codec(c)),
    }
}

// ---- the engine: one per thread and device, created on first use (handles are `&mut self` objects, as in the crate) ----
struct Engine(*mut ffi::flacb200_engine);

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::flacb200_engine_destroy(self.0) }
    }
}

thread_local! {
    static ENGINE: RefCell<Option<Engine>> = const { RefCell::new(None) };
}

/// CUDA device the calling thread's engine lives on: `FLACB200_DEVICE` (default 0).
fn engine() -> Result<*mut ffi::flacb200_engine, Error> {
    ENGINE.with(|slot| {
        let mut slot = slot.borrow_mut();
        if slot.is_none() {
            let device = std::env::var("FLACB200_DEVICE").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
            let mut e = ptr::null_mut();
            ck(unsafe { ffi::flacb200_engine_create(device, &mut e) })?;
            *slot = Some(Engine(e));
        }
        Ok(slot.as_ref().unwrap().0)
    })
}

/// `flac_codec::encode::Options` (src/encode.rs:1363-1672): the fields the engine and the container use.
#[derive(Clone)]
pub struct Options(ffi::flacb200_writer_options);

impl Default for Options {
    fn default() -> Self {
        let mut o = ffi::flacb200_writer_options::default();
        o.frame = ffi::flacb200_options { block_size: 4096, max_lpc_order: 8, max_partition_order: 5, mid_side: 1,
            exhaustive_channel_correlation: 1, window_kind: 2, reserved0: 0, tukey_p: 0.5 };
        o.padding = 4096;
        o.seektable_kind = 1;
        o.seektable_n = 10;
        Options(o)
    }
}

impl Options {
    pub fn fast() -> Self {
        let mut o = Self::default();
        o.0.frame.block_size = 1152;
        o.0.frame.mid_side = 0;
        o.0.frame.max_partition_order = 3;
        o.0.frame.max_lpc_order = 0;
        o.0.frame.exhaustive_channel_correlation = 0;
        o
    }
    pub fn best() -> Self {
        let mut o = Self::default();
        o.0.frame.max_partition_order = 6;
        o.0.frame.max_lpc_order = 12;
        o
    }
    pub fn block_size(mut self, n: u16) -> Result<Self, &'static str> {
        if n < 16 { return Err("block size must be >= 16"); }
        self.0.frame.block_size = n;
        Ok(self)
    }
    pub fn max_lpc_order(mut self, n: Option<u8>) -> Result<Self, &'static str> {
        match n { Some(v) if v == 0 || v > 32 => return Err("maximum LPC order must be <= 32"), _ => {} }
        self.0.frame.max_lpc_order = n.unwrap_or(0);
        Ok(self)
    }
    pub fn max_partition_order(mut self, n: u32) -> Result<Self, &'static str> {
        if n > 15 { return Err("max partition order must be <= 15"); }
        self.0.frame.max_partition_order = n as u8;
        Ok(self)
    }
    pub fn mid_side(mut self, on: bool) -> Self { self.0.frame.mid_side = on as u8; self }
    pub fn fast_channel_correlation(mut self, fast: bool) -> Self { self.0.frame.exhaustive_channel_correlation = (!fast) as u8; self }
    pub fn padding(mut self, size: u32) -> Self { self.0.padding = if size == 0 { -1 } else { size as i32 }; self }
    pub fn no_padding(mut self) -> Self { self.0.padding = -1; self }
    pub fn seektable_seconds(mut self, s: u8) -> Self { self.0.seektable_kind = (s != 0) as u32; self.0.seektable_n = s as u32; self }
    pub fn seektable_frames(mut self, n: usize) -> Self { self.0.seektable_kind = if n != 0 { 2 } else { 0 }; self.0.seektable_n = n as u32; self }
    pub fn no_seektable(mut self) -> Self { self.0.seektable_kind = 0; self }
    /// blocks encoded per GPU launch (engine extension; the reference encodes one frame per call)
    pub fn launch_frames(mut self, n: u32) -> Self { self.0.launch_frames = n; self }
}

// =====================================================================================================================
// writers
// =====================================================================================================================
/// `Encoder<W>` (src/encode.rs:1860-2110): owns the sink, forwards completed frames, rewrites the metadata.
struct Encoder<W: Write + Seek> {
    w: W,
    h: *mut ffi::flacb200_writer,
    start: u64,
    finalized: bool,
}

impl<W: Write + Seek> Encoder<W> {
    fn new(mut w: W, o: &Options, rate: u32, bps: u32, channels: u8, total_pcm_frames: u64) -> Result<Self, Error> {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::flacb200_writer_open(engine()?, &o.0, rate, bps, channels as u32, total_pcm_frames, &mut h) })?;
        let start = w.stream_position()?;
        let mut e = Encoder { w, h, start, finalized: false };
        e.put_header()?;
        Ok(e)
    }
    fn put_header(&mut self) -> Result<(), Error> {
        let (mut p, mut n) = (ptr::null(), 0usize);
        ck(unsafe { ffi::flacb200_writer_header(self.h, &mut p, &mut n) })?;
        self.w.write_all(unsafe { std::slice::from_raw_parts(p, n) })?;
        Ok(())
    }
    fn drain(&mut self) -> Result<(), Error> {
        let (mut p, mut n) = (ptr::null(), 0usize);
        ck(unsafe { ffi::flacb200_writer_drain(self.h, &mut p, &mut n) })?;
        if n != 0 {
            self.w.write_all(unsafe { std::slice::from_raw_parts(p, n) })?;
        }
        Ok(())
    }
    fn finalize_inner(&mut self) -> Result<(), Error> {
        if std::mem::replace(&mut self.finalized, true) {
            return Ok(());
        }
        ck(unsafe { ffi::flacb200_writer_finalize(self.h) })?;
        self.drain()?;
        let end = self.w.stream_position()?;
        self.w.seek(SeekFrom::Start(self.start))?;
        self.put_header()?;
        self.w.seek(SeekFrom::Start(end))?;
        Ok(())
    }
}

impl<W: Write + Seek> Drop for Encoder<W> {
    fn drop(&mut self) {
        let _ = self.finalize_inner(); // the reference's Drop finalises and swallows errors (:399-405)
        unsafe { ffi::flacb200_writer_close(self.h) }
    }
}

/// `FlacSampleWriter<W>` (src/encode.rs:431-628)
pub struct FlacSampleWriter<W: Write + Seek>(Encoder<W>);

impl<W: Write + Seek> FlacSampleWriter<W> {
    pub fn new(writer: W, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8, total_samples: Option<u64>) -> Result<Self, Error> {
        let mut total = 0u64;
        if let Some(t) = total_samples {
            ck(unsafe { ffi::flacb200_total_from_samples(t, channels as u32, &mut total) })?;
        }
        Ok(Self(Encoder::new(writer, &options, sample_rate, bits_per_sample, channels, total)?))
    }
    pub fn new_cdda(writer: W, options: Options, total_samples: Option<u64>) -> Result<Self, Error> {
        Self::new(writer, options, 44100, 16, 2, total_samples)
    }
    pub fn write(&mut self, samples: &[i32]) -> Result<(), Error> {
        ck(unsafe { ffi::flacb200_writer_write_samples(self.0.h, samples.as_ptr(), samples.len()) })?;
        self.0.drain()
    }
    pub fn finalize(mut self) -> Result<(), Error> {
        self.0.finalize_inner()
    }
}

/// Byte order marker types, as `flac_codec::byteorder::{LittleEndian, BigEndian}` (src/byteorder.rs:48-186).
pub trait Endianness { const BIG: bool; }
pub struct LittleEndian;
pub struct BigEndian;
impl Endianness for LittleEndian { const BIG: bool = false; }
impl Endianness for BigEndian { const BIG: bool = true; }

/// `FlacByteWriter<W, E>` (src/encode.rs:103-405)
pub struct FlacByteWriter<W: Write + Seek, E: Endianness>(Encoder<W>, PhantomData<E>);

impl<W: Write + Seek, E: Endianness> FlacByteWriter<W, E> {
    pub fn new(writer: W, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8, total_bytes: Option<u64>) -> Result<Self, Error> {
        let mut total = 0u64;
        if let Some(t) = total_bytes {
            ck(unsafe { ffi::flacb200_total_from_bytes(t, bits_per_sample, channels as u32, &mut total) })?;
        }
        Ok(Self(Encoder::new(writer, &options, sample_rate, bits_per_sample, channels, total)?, PhantomData))
    }
    pub fn endian(writer: W, _endianness: E, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8,
                  total_bytes: Option<u64>) -> Result<Self, Error> {
        Self::new(writer, options, sample_rate, bits_per_sample, channels, total_bytes)
    }
    pub fn new_cdda(writer: W, options: Options, total_bytes: Option<u64>) -> Result<Self, Error> {
        Self::new(writer, options, 44100, 16, 2, total_bytes)
    }
    pub fn finalize(mut self) -> Result<(), Error> {
        self.0.finalize_inner()
    }
}

impl<W: Write + Seek, E: Endianness> Write for FlacByteWriter<W, E> {
    fn write(&mut self, buf: &[u8]) -> std::io::Result<usize> {
        ck(unsafe { ffi::flacb200_writer_write_bytes(self.0.h, buf.as_ptr(), buf.len(), E::BIG as i32) })?;
        self.0.drain()?;
        Ok(buf.len()) // the whole slice is always consumed (:387)
    }
    fn flush(&mut self) -> std::io::Result<()> {
        // never emits a partial block (:391-395); completed blocks still buffered for the next launch are encoded now
        ck(unsafe { ffi::flacb200_writer_flush(self.0.h) })?;
        self.0.drain()?;
        self.0.w.flush()
    }
}

/// `FlacChannelWriter<W>` (src/encode.rs:713-893)
pub struct FlacChannelWriter<W: Write + Seek>(Encoder<W>);

impl<W: Write + Seek> FlacChannelWriter<W> {
    pub fn new(writer: W, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8, total_samples: Option<u64>) -> Result<Self, Error> {
        if total_samples == Some(0) {
            return Err(Error::Codec(63)); // InvalidTotalSamples
        }
        Ok(Self(Encoder::new(writer, &options, sample_rate, bits_per_sample, channels, total_samples.unwrap_or(0))?))
    }
    pub fn write<C: AsRef<[S]>, S: AsRef<[i32]>>(&mut self, channels: C) -> Result<(), Error> {
        let chans = channels.as_ref();
        let n = chans.first().map(|c| c.as_ref().len()).unwrap_or(0);
        if chans.iter().any(|c| c.as_ref().len() != n) {
            return Err(Error::Codec(65)); // ChannelLengthMismatch (:851)
        }
        let ptrs: Vec<*const i32> = chans.iter().map(|c| c.as_ref().as_ptr()).collect();
        ck(unsafe { ffi::flacb200_writer_write_channels(self.0.h, ptrs.as_ptr(), ptrs.len() as u32, n) })?; // ChannelCountMismatch inside
        self.0.drain()
    }
    pub fn finalize(mut self) -> Result<(), Error> {
        self.0.finalize_inner()
    }
}

/// `FlacStreamWriter<W>` (src/encode.rs:1063-1290): subset frames, no metadata, parameters per call.
pub struct FlacStreamWriter<W: Write> {
    w: W,
    options: Options,
    frame_number: u64,
    buf: Vec<u8>,
}

impl<W: Write> FlacStreamWriter<W> {
    pub fn new(writer: W, options: Options) -> Self {
        Self { w: writer, options, frame_number: 0, buf: Vec::new() }
    }
    pub fn write(&mut self, sample_rate: u32, channels: u8, bits_per_sample: u32, samples: &[i32]) -> Result<(), Error> {
        self.buf.resize(samples.len() * 5 + 1024, 0);
        let mut n = 0usize;
        ck(unsafe {
            ffi::flacb200_stream_write(engine()?, &self.options.0.frame, sample_rate, channels as u32, bits_per_sample, samples.as_ptr(),
                                       samples.len(), self.frame_number, self.buf.as_mut_ptr(), self.buf.len(), &mut n)
        })?;
        if n != 0 {
            self.frame_number += 1;
            self.w.write_all(&self.buf[..n])?;
        }
        Ok(())
    }
    pub fn write_cdda(&mut self, samples: &[i32]) -> Result<(), Error> {
        self.write(44100, 2, 16, samples)
    }
}

// =====================================================================================================================
// readers
// =====================================================================================================================
/// `Decoder<R>` (src/decode.rs:1311-1491) behind a fed `flacb200_reader`: the handle decodes window by window and asks
/// for bytes (NEED_DATA) or for the source to be repositioned (NEED_SEEK); the file is never held in memory.
struct Decoder<R> {
    r: R,
    h: *mut ffi::flacb200_reader,
    info: ffi::flacb200_streaminfo,
    eof: bool,
    chunk: Vec<u8>,
    reposition: Option<fn(&mut R, u64) -> std::io::Result<()>>,
}

impl<R: Read> Decoder<R> {
    fn open(r: R, reposition: Option<fn(&mut R, u64) -> std::io::Result<()>>) -> Result<Self, Error> {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::flacb200_reader_open_stream(engine()?, &mut h) })?;
        if reposition.is_some() {
            ck(unsafe { ffi::flacb200_reader_set_seekable(h, 1) })?;
        }
        let mut d = Decoder { r, h, info: Default::default(), eof: false, chunk: vec![0u8; 1 << 20], reposition };
        let mut info = ffi::flacb200_streaminfo::default();
        d.call(|h| unsafe { ffi::flacb200_reader_info(h, &mut info) })?; // BlockList::read
        d.info = info;
        Ok(d)
    }
    /// One handle call; NEED_DATA is answered with the next chunk of the source, NEED_SEEK by repositioning it.
    fn call(&mut self, mut f: impl FnMut(*mut ffi::flacb200_reader) -> i32) -> Result<(), Error> {
        loop {
            match f(self.h) {
                ffi::FLACB200_NEED_DATA => {
                    if self.eof {
                        return Err(Error::Io(std::io::ErrorKind::UnexpectedEof.into()));
                    }
                    let n = self.r.read(&mut self.chunk)?;
                    self.eof = n == 0;
                    ck(unsafe { ffi::flacb200_reader_feed(self.h, self.chunk.as_ptr(), n, self.eof as i32) })?;
                }
                ffi::FLACB200_NEED_SEEK => {
                    let mut off = 0u64;
                    ck(unsafe { ffi::flacb200_reader_wanted_offset(self.h, &mut off) })?;
                    (self.reposition.expect("seekable"))(&mut self.r, off)?;
                    self.eof = false;
                }
                rc => return ck(rc),
            }
        }
    }
    fn seek(&mut self, sample: u64) -> Result<(), Error> {
        if self.reposition.is_none() {
            return Err(Error::Io(std::io::ErrorKind::Unsupported.into())); // frames_start: None -> NotSeekable
        }
        self.call(|h| unsafe { ffi::flacb200_reader_seek(h, sample) })
    }
}

impl<R> Drop for Decoder<R> {
    fn drop(&mut self) {
        unsafe { ffi::flacb200_reader_close(self.h) }
    }
}

fn reposition<R: Seek>(r: &mut R, off: u64) -> std::io::Result<()> {
    r.seek(SeekFrom::Start(off)).map(|_| ())
}

/// `flac_codec::metadata::Metadata` (src/metadata/mod.rs:48-105) for the reader types
macro_rules! metadata_methods {
    () => {
        pub fn channel_count(&self) -> u8 { self.d.info.channels as u8 }
        pub fn sample_rate(&self) -> u32 { self.d.info.sample_rate }
        pub fn bits_per_sample(&self) -> u32 { self.d.info.bits_per_sample }
        pub fn total_samples(&self) -> Option<u64> { (self.d.info.total_samples != 0).then_some(self.d.info.total_samples) }
        pub fn md5(&self) -> Option<&[u8; 16]> { self.d.info.md5.iter().any(|b| *b != 0).then_some(&self.d.info.md5) }
        pub fn decoded_len(&self) -> Option<u64> {
            self.total_samples().map(|t| t * self.d.info.channels as u64 * self.d.info.bits_per_sample.div_ceil(8) as u64)
        }
    };
}

/// `FlacSampleReader<R>` (src/decode.rs:384-620)
pub struct FlacSampleReader<R> { d: Decoder<R> }

impl<R: Read> FlacSampleReader<R> {
    pub fn new(reader: R) -> Result<Self, Error> { Ok(Self { d: Decoder::open(reader, None)? }) }
    metadata_methods!();
    pub fn read(&mut self, samples: &mut [i32]) -> Result<usize, Error> {
        let mut n = 0usize;
        self.d.call(|h| unsafe { ffi::flacb200_reader_read(h, samples.as_mut_ptr().cast(), samples.len(), ffi::FLACB200_PCM_I32_INTERLEAVED, &mut n) })?;
        Ok(n)
    }
    pub fn read_to_end(&mut self, buf: &mut Vec<i32>) -> Result<usize, Error> {
        let mut total = 0;
        loop {
            let got = self.fill_buf()?.to_vec();
            if got.is_empty() { return Ok(total); }
            total += got.len();
            self.consume(got.len());
            buf.extend_from_slice(&got);
        }
    }
    /// the unconsumed interleaved samples of the current frame (:466)
    pub fn fill_buf(&mut self) -> Result<&[i32], Error> {
        let (mut p, mut n) = (ptr::null(), 0usize);
        self.d.call(|h| unsafe { ffi::flacb200_reader_fill_buf(h, &mut p, &mut n) })?;
        Ok(if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(p, n) } })
    }
    pub fn consume(&mut self, amt: usize) {
        unsafe { ffi::flacb200_reader_consume(self.d.h, amt) };
    }
}

impl<R: Read + Seek> FlacSampleReader<R> {
    pub fn new_seekable(reader: R) -> Result<Self, Error> { Ok(Self { d: Decoder::open(reader, Some(reposition::<R>))? }) }
    /// `FlacSampleReader::seek` (:823-860): channel-independent sample from the start of the stream
    pub fn seek(&mut self, sample: u64) -> Result<(), Error> { self.d.seek(sample) }
}

impl FlacSampleReader<std::io::BufReader<std::fs::File>> {
    pub fn open<P: AsRef<std::path::Path>>(path: P) -> Result<Self, Error> {
        Self::new_seekable(std::io::BufReader::new(std::fs::File::open(path)?))
    }
}

/// `FlacSampleIterator<R>` (src/decode.rs:667-712)
pub struct FlacSampleIterator<R> { reader: FlacSampleReader<R>, buf: std::collections::VecDeque<i32> }

impl<R: Read> IntoIterator for FlacSampleReader<R> {
    type Item = Result<i32, Error>;
    type IntoIter = FlacSampleIterator<R>;
    fn into_iter(self) -> FlacSampleIterator<R> { FlacSampleIterator { reader: self, buf: Default::default() } }
}

impl<R: Read> Iterator for FlacSampleIterator<R> {
    type Item = Result<i32, Error>;
    fn next(&mut self) -> Option<Self::Item> {
        if let Some(s) = self.buf.pop_front() { return Some(Ok(s)); }
        match self.reader.fill_buf() {
            Ok([]) => None,
            Ok(frame) => {
                let n = frame.len();
                self.buf.extend(frame.iter().copied());
                self.reader.consume(n);
                self.buf.pop_front().map(Ok)
            }
            Err(e) => Some(Err(e)),
        }
    }
}

/// `FlacByteReader<R, E>` (src/decode.rs:103-371): `Read` + `BufRead` + `Seek` over the decoded PCM bytes
pub struct FlacByteReader<R, E: Endianness> { d: Decoder<R>, buf: Vec<u8>, pos: usize, byte_pos: u64, _e: PhantomData<E> }

impl<R: Read, E: Endianness> FlacByteReader<R, E> {
    pub fn new(reader: R) -> Result<Self, Error> {
        Ok(Self { d: Decoder::open(reader, None)?, buf: Vec::new(), pos: 0, byte_pos: 0, _e: PhantomData })
    }
    pub fn endian(reader: R, _endianness: E) -> Result<Self, Error> { Self::new(reader) }
    metadata_methods!();
    fn refill(&mut self) -> Result<(), Error> {
        // one frame at a time, like Decoder::read_frame + Frame::to_buf (src/audio.rs:110)
        let frame_bytes = self.d.info.max_block_size.max(16) as usize * self.d.info.channels as usize * self.d.info.bits_per_sample.div_ceil(8) as usize;
        self.buf.resize(frame_bytes, 0);
        let (mut n, kind) = (0usize, if E::BIG { ffi::FLACB200_PCM_BYTES_BE } else { ffi::FLACB200_PCM_BYTES_LE });
        let (p, cap) = (self.buf.as_mut_ptr(), self.buf.len());
        self.d.call(|h| unsafe { ffi::flacb200_reader_read(h, p.cast(), cap, kind, &mut n) })?;
        self.buf.truncate(n);
        self.pos = 0;
        Ok(())
    }
}

impl<R: Read, E: Endianness> Read for FlacByteReader<R, E> {
    fn read(&mut self, out: &mut [u8]) -> std::io::Result<usize> {
        let got = self.fill_buf()?;
        let n = got.len().min(out.len());
        out[..n].copy_from_slice(&got[..n]);
        self.consume(n);
        Ok(n)
    }
}

impl<R: Read, E: Endianness> BufRead for FlacByteReader<R, E> {
    fn fill_buf(&mut self) -> std::io::Result<&[u8]> {
        if self.pos >= self.buf.len() { self.refill()?; }
        Ok(&self.buf[self.pos..])
    }
    fn consume(&mut self, amt: usize) {
        self.pos += amt;
        self.byte_pos += amt as u64;
    }
}

impl<R: Read + Seek, E: Endianness> FlacByteReader<R, E> {
    pub fn new_seekable(reader: R) -> Result<Self, Error> {
        Ok(Self { d: Decoder::open(reader, Some(reposition::<R>))?, buf: Vec::new(), pos: 0, byte_pos: 0, _e: PhantomData })
    }
}

impl<R: Read + Seek, E: Endianness> Seek for FlacByteReader<R, E> {
    // src/decode.rs:715-820
    fn seek(&mut self, pos: SeekFrom) -> std::io::Result<u64> {
        let bpf = (self.d.info.bits_per_sample.div_ceil(8) * self.d.info.channels) as u64;
        let invalid = |m: &'static str| std::io::Error::new(std::io::ErrorKind::InvalidInput, m);
        let desired = match pos {
            SeekFrom::Start(p) => p,
            SeekFrom::Current(0) => return Ok(self.byte_pos),
            SeekFrom::Current(d) if d < 0 => self.byte_pos.checked_sub(d.unsigned_abs()).ok_or_else(|| invalid("cannot seek below byte 0"))?,
            SeekFrom::Current(d) => self.byte_pos.checked_add(d as u64).ok_or_else(|| invalid("seek offset too large"))?,
            SeekFrom::End(d) => {
                // (the crate takes total_samples -- PCM frames, not bytes -- as the end position, :766-768; kept as it is)
                let max = self.d.info.total_samples;
                if max == 0 { return Err(std::io::Error::new(std::io::ErrorKind::Unsupported, "total samples not known")); }
                if d > 0 { return Err(invalid("cannot seek beyond end of file")); }
                max.checked_sub(d.unsigned_abs()).ok_or_else(|| invalid("cannot seek below byte 0"))?
            }
        };
        self.d.seek(desired / bpf)?;
        self.buf.clear();
        self.pos = 0;
        self.byte_pos = desired / bpf * bpf;
        while self.byte_pos < desired {   // a position inside a PCM frame: skip its leading bytes
            let have = self.fill_buf()?.len();
            if have == 0 { return Err(std::io::Error::new(std::io::ErrorKind::UnexpectedEof, "stream exhausted before sample reached")); }
            self.consume(have.min((desired - self.byte_pos) as usize));
        }
        Ok(desired)
    }
}

/// `FlacChannelReader<R>` (src/decode.rs:880-1065)
pub struct FlacChannelReader<R> { d: Decoder<R> }

impl<R: Read> FlacChannelReader<R> {
    pub fn new(reader: R) -> Result<Self, Error> { Ok(Self { d: Decoder::open(reader, None)? }) }
    metadata_methods!();
    /// one slice per channel over the unconsumed part of the current frame (:917)
    pub fn fill_buf(&mut self) -> Result<Vec<&[i32]>, Error> {
        let (mut pp, mut n) = (ptr::null(), 0usize);
        self.d.call(|h| unsafe { ffi::flacb200_reader_fill_channels(h, &mut pp, &mut n) })?;
        let ch = self.d.info.channels as usize;
        Ok((0..ch).map(|c| if n == 0 { &[][..] } else { unsafe { std::slice::from_raw_parts(*pp.add(c), n) } }).collect())
    }
    pub fn consume(&mut self, amt: usize) {
        unsafe { ffi::flacb200_reader_consume_channels(self.d.h, amt) };
    }
}

impl<R: Read + Seek> FlacChannelReader<R> {
    pub fn new_seekable(reader: R) -> Result<Self, Error> { Ok(Self { d: Decoder::open(reader, Some(reposition::<R>))? }) }
    pub fn seek(&mut self, sample: u64) -> Result<(), Error> { self.d.seek(sample) }
}

/// `FrameBuf` (src/decode.rs:1253-1268)
#[derive(Copy, Clone, Debug, Eq, PartialEq)]
pub struct FrameBuf<'s> {
    pub samples: &'s [i32],
    pub sample_rate: u32,
    pub channels: u8,
    pub bits_per_sample: u32,
}

/// `FlacStreamReader<R>` (src/decode.rs:1149-1240): subset frames, parameters from every frame header
pub struct FlacStreamReader<R> { r: R, h: *mut ffi::flacb200_stream_reader, eof: bool }

impl<R: BufRead> FlacStreamReader<R> {
    pub fn new(reader: R) -> Self {
        let mut h = ptr::null_mut();
        // engine errors surface at the first read (the reference's constructor is infallible)
        if let Ok(e) = engine() { unsafe { ffi::flacb200_stream_reader_open(e, &mut h) }; }
        Self { r: reader, h, eof: false }
    }
    pub fn read(&mut self) -> Result<FrameBuf<'_>, Error> {
        if self.h.is_null() { engine()?; }
        let mut fb = std::mem::MaybeUninit::<ffi::flacb200_framebuf>::zeroed();
        loop {
            match unsafe { ffi::flacb200_stream_reader_read(self.h, fb.as_mut_ptr()) } {
                ffi::FLACB200_NEED_DATA => {
                    if self.eof { return Err(Error::Io(std::io::Error::new(std::io::ErrorKind::UnexpectedEof, "eof looking for frame sync"))); }
                    let chunk = self.r.fill_buf()?;
                    let n = chunk.len();
                    self.eof = n == 0;
                    ck(unsafe { ffi::flacb200_stream_reader_feed(self.h, chunk.as_ptr(), n, self.eof as i32) })?;
                    self.r.consume(n);
                }
                rc => { ck(rc)?; break; }
            }
        }
        let fb = unsafe { fb.assume_init() };
        Ok(FrameBuf { samples: unsafe { std::slice::from_raw_parts(fb.samples, fb.n_samples) }, sample_rate: fb.sample_rate,
                      channels: fb.channels as u8, bits_per_sample: fb.bits_per_sample })
    }
}

impl<R> Drop for FlacStreamReader<R> {
    fn drop(&mut self) {
        if !self.h.is_null() { unsafe { ffi::flacb200_stream_reader_close(self.h) } }
    }
}

/// `flac_codec::decode::Verified` (src/decode.rs:1271-1280)
#[derive(Debug, Copy, Clone, PartialEq, Eq, Hash, PartialOrd, Ord)]
pub enum Verified { MD5Match, MD5Mismatch, NoMD5 }

/// `flac_codec::decode::verify_reader` (src/decode.rs:1291-1309)
pub fn verify_reader<R: Read>(reader: R) -> Result<Verified, Error> {
    let mut d = Decoder::open(reader, None)?;
    let mut res = 0i32;
    d.call(|h| unsafe { ffi::flacb200_reader_verify(h, &mut res, ptr::null_mut()) })?;
    Ok(match res { 0 => Verified::MD5Match, 1 => Verified::MD5Mismatch, _ => Verified::NoMD5 })
}

/// `flac_codec::decode::verify` (src/decode.rs:1282-1286)
pub fn verify<P: AsRef<std::path::Path>>(p: P) -> Result<Verified, Error> {
    verify_reader(std::io::BufReader::new(std::fs::File::open(p)?))
}

// =====================================================================================================================
// whole-file batches: what the reference's examples do with rayon (examples/flac2wav.rs:31-38, flac-split.rs:84-87)
// =====================================================================================================================
/// Encodes many tracks (interleaved little-endian PCM bytes) into complete `.flac` images on the given CUDA devices.
pub fn encode_files(tracks: &[(&[u8], u32, u32, u8)], options: &Options, devices: &[i32]) -> Vec<Result<Vec<u8>, Error>> {
    let t: Vec<ffi::flacb200_track> = tracks.iter().map(|(pcm, rate, bps, ch)| ffi::flacb200_track {
        pcm: pcm.as_ptr().cast(), n_pcm_frames: (pcm.len() / (*ch as usize * bps.div_ceil(8) as usize)) as u64, sample_rate: *rate,
        bits_per_sample: *bps, channels: *ch as u32, pcm_kind: ffi::FLACB200_PCM_BYTES_LE }).collect();
    let mut files = vec![ffi::flacb200_file { data: ptr::null_mut(), capacity: 0, len: 0, status: 0, frames: 0, md5: [0; 16] }; t.len()];
    unsafe { ffi::flacb200_encode_batch(t.as_ptr(), t.len(), &options.0, devices.as_ptr(), devices.len() as i32, files.as_mut_ptr()) };
    let out = files.iter().map(|f| if f.status == 0 { Ok(unsafe { std::slice::from_raw_parts(f.data, f.len) }.to_vec()) } else { Err(Error::Codec(f.status)) }).collect();
    unsafe { ffi::flacb200_files_free(files.as_mut_ptr(), files.len()) };
    out
}

/// Decodes many `.flac` images to little-endian PCM bytes, optionally verifying each MD5.
pub fn decode_files(flacs: &[&[u8]], verify: bool, devices: &[i32]) -> Vec<Result<(Vec<u8>, Option<Verified>), Error>> {
    let ptrs: Vec<*const u8> = flacs.iter().map(|f| f.as_ptr()).collect();
    let lens: Vec<usize> = flacs.iter().map(|f| f.len()).collect();
    let mut out: Vec<ffi::flacb200_pcm> = (0..flacs.len()).map(|_| unsafe { std::mem::zeroed() }).collect();
    unsafe { ffi::flacb200_decode_batch(ptrs.as_ptr(), lens.as_ptr(), flacs.len(), ffi::FLACB200_PCM_BYTES_LE, verify as i32, devices.as_ptr(),
                                        devices.len() as i32, out.as_mut_ptr()) };
    let res = out.iter().map(|o| if o.status != 0 { Err(Error::Codec(o.status)) } else {
        Ok((unsafe { std::slice::from_raw_parts(o.data as *const u8, o.len) }.to_vec(),
            match o.verified { 0 => Some(Verified::MD5Match), 1 => Some(Verified::MD5Mismatch), 2 => Some(Verified::NoMD5), _ => None }))
    }).collect();
    unsafe { ffi::flacb200_pcm_free(out.as_mut_ptr(), out.len()) };
    res
}
