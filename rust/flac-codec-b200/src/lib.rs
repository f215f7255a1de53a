//! flac-codec's public writer/reader types, backed by the B200 frame engine.
//!
//! Drop-in for the hot path of `flac_codec::encode::{FlacByteWriter, FlacSampleWriter, FlacChannelWriter}` and
//! `flac_codec::decode::{FlacByteReader, FlacSampleReader}`: same constructors, same `write`/`finalize`/`read`
//! signatures, same `Error` ordinals (the C ABI returns the 1-based ordinal of `flac_codec::Error`).  All frame
//! work happens in `libflacb200.so`; this file only moves bytes between the caller's `W: Write + Seek` /
//! `R: Read` and the handles.  NOT BUILT in the development image (no cargo/rustc there) -- kept thin and
//! mechanical on purpose; the same handle protocol is exercised end to end by `flac_codec_b200/stream.py`
//! (ctypes) in tests/test_gpu_stream.py.
mod ffi;

use std::io::{Read, Seek, SeekFrom, Write};
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
    if rc == 0 { Ok(()) } else { Err(Error::Codec(rc)) }
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

/// One engine per GPU; cheap to share between handles used from one thread.
pub struct Engine(*mut ffi::flacb200_engine);

impl Engine {
    pub fn new(device: i32) -> Result<Self, Error> {
        let mut e = ptr::null_mut();
        ck(unsafe { ffi::flacb200_engine_create(device, &mut e) })?;
        Ok(Engine(e))
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::flacb200_engine_destroy(self.0) }
    }
}

/// `Encoder<W>` (src/encode.rs:1860-2110): owns the sink, forwards completed frames, rewrites the metadata.
struct Encoder<W: Write + Seek> {
    w: W,
    h: *mut ffi::flacb200_writer,
    start: u64,
    finalized: bool,
}

impl<W: Write + Seek> Encoder<W> {
    fn new(mut w: W, engine: &Engine, o: &Options, rate: u32, bps: u32, channels: u8, total_pcm_frames: u64) -> Result<Self, Error> {
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::flacb200_writer_open(engine.0, &o.0, rate, bps, channels as u32, total_pcm_frames, &mut h) })?;
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
        let _ = self.finalize_inner(); // the reference's Drop finalises and swallows errors (:2113)
        unsafe { ffi::flacb200_writer_close(self.h) }
    }
}

/// `FlacSampleWriter<W>` (src/encode.rs:431-628)
pub struct FlacSampleWriter<W: Write + Seek>(Encoder<W>);

impl<W: Write + Seek> FlacSampleWriter<W> {
    pub fn new(writer: W, engine: &Engine, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8,
               total_samples: Option<u64>) -> Result<Self, Error> {
        let mut total = 0u64;
        if let Some(t) = total_samples {
            ck(unsafe { ffi::flacb200_total_from_samples(t, channels as u32, &mut total) })?;
        }
        Ok(Self(Encoder::new(writer, engine, &options, sample_rate, bits_per_sample, channels, total)?))
    }
    pub fn new_cdda(writer: W, engine: &Engine, options: Options, total_samples: Option<u64>) -> Result<Self, Error> {
        Self::new(writer, engine, options, 44100, 16, 2, total_samples)
    }
    pub fn write(&mut self, samples: &[i32]) -> Result<(), Error> {
        ck(unsafe { ffi::flacb200_writer_write_samples(self.0.h, samples.as_ptr(), samples.len()) })?;
        self.0.drain()
    }
    pub fn finalize(mut self) -> Result<(), Error> {
        self.0.finalize_inner()
    }
}

/// `FlacByteWriter<W, E>` (src/encode.rs:103-405); `BIG` selects the byte order of the input samples.
pub struct FlacByteWriter<W: Write + Seek, const BIG: bool = false>(Encoder<W>);

impl<W: Write + Seek, const BIG: bool> FlacByteWriter<W, BIG> {
    pub fn new(writer: W, engine: &Engine, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8,
               total_bytes: Option<u64>) -> Result<Self, Error> {
        let mut total = 0u64;
        if let Some(t) = total_bytes {
            ck(unsafe { ffi::flacb200_total_from_bytes(t, bits_per_sample, channels as u32, &mut total) })?;
        }
        Ok(Self(Encoder::new(writer, engine, &options, sample_rate, bits_per_sample, channels, total)?))
    }
    pub fn finalize(mut self) -> Result<(), Error> {
        self.0.finalize_inner()
    }
}

impl<W: Write + Seek, const BIG: bool> Write for FlacByteWriter<W, BIG> {
    fn write(&mut self, buf: &[u8]) -> std::io::Result<usize> {
        ck(unsafe { ffi::flacb200_writer_write_bytes(self.0.h, buf.as_ptr(), buf.len(), BIG as i32) })?;
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
    pub fn new(writer: W, engine: &Engine, options: Options, sample_rate: u32, bits_per_sample: u32, channels: u8,
               total_samples: Option<u64>) -> Result<Self, Error> {
        if total_samples == Some(0) {
            return Err(Error::Codec(63)); // InvalidTotalSamples
        }
        Ok(Self(Encoder::new(writer, engine, &options, sample_rate, bits_per_sample, channels, total_samples.unwrap_or(0))?))
    }
    pub fn write<C: AsRef<[S]>, S: AsRef<[i32]>>(&mut self, channels: C) -> Result<(), Error> {
        let chans = channels.as_ref();
        let n = chans.first().map(|c| c.as_ref().len()).unwrap_or(0);
        if chans.iter().any(|c| c.as_ref().len() != n) {
            return Err(Error::Codec(65)); // ChannelLengthMismatch (:851)
        }
        let ptrs: Vec<*const i32> = chans.iter().map(|c| c.as_ref().as_ptr()).collect();
        ck(unsafe { ffi::flacb200_writer_write_channels(self.0.h, ptrs.as_ptr(), ptrs.len() as u32, n) })?; // count checked inside
        self.0.drain()
    }
    pub fn finalize(mut self) -> Result<(), Error> {
        self.0.finalize_inner()
    }
}

/// `FlacSampleReader<R>` (src/decode.rs:384-620): the file image is read once, decoded in one GPU batch.
pub struct FlacSampleReader {
    h: *mut ffi::flacb200_reader,
    _image: Vec<u8>,
    info: ffi::flacb200_streaminfo,
}

impl FlacSampleReader {
    pub fn new<R: Read>(mut reader: R, engine: &Engine) -> Result<Self, Error> {
        let mut image = Vec::new();
        reader.read_to_end(&mut image)?;
        let mut h = ptr::null_mut();
        ck(unsafe { ffi::flacb200_reader_open(engine.0, image.as_ptr(), image.len(), &mut h) })?;
        let mut info = ffi::flacb200_streaminfo::default();
        ck(unsafe { ffi::flacb200_reader_info(h, &mut info) })?;
        Ok(Self { h, _image: image, info })
    }
    pub fn channel_count(&self) -> u8 { self.info.channels as u8 }
    pub fn sample_rate(&self) -> u32 { self.info.sample_rate }
    pub fn bits_per_sample(&self) -> u32 { self.info.bits_per_sample }
    pub fn total_samples(&self) -> Option<u64> { (self.info.total_samples != 0).then_some(self.info.total_samples) }
    pub fn md5(&self) -> Option<&[u8; 16]> { self.info.md5.iter().any(|b| *b != 0).then_some(&self.info.md5) }
    pub fn read(&mut self, samples: &mut [i32]) -> Result<usize, Error> {
        let mut n = 0usize;
        ck(unsafe { ffi::flacb200_reader_read(self.h, samples.as_mut_ptr().cast(), samples.len(), ffi::FLACB200_PCM_I32_INTERLEAVED, &mut n) })?;
        Ok(n)
    }
    pub fn seek(&mut self, sample: u64) -> Result<(), Error> {
        ck(unsafe { ffi::flacb200_reader_seek(self.h, sample) })
    }
}

impl Drop for FlacSampleReader {
    fn drop(&mut self) {
        unsafe { ffi::flacb200_reader_close(self.h) }
    }
}

/// `flac_codec::decode::Verified` (src/decode.rs:1271-1280)
#[derive(Debug, PartialEq, Eq)]
pub enum Verified { MD5Match, MD5Mismatch, NoMD5 }

/// `flac_codec::decode::verify_reader` (src/decode.rs:1291-1309)
pub fn verify_reader<R: Read>(reader: R, engine: &Engine) -> Result<Verified, Error> {
    let r = FlacSampleReader::new(reader, engine)?;
    let mut res = 0i32;
    ck(unsafe { ffi::flacb200_reader_verify(r.h, &mut res, ptr::null_mut()) })?;
    Ok(match res { 0 => Verified::MD5Match, 1 => Verified::MD5Mismatch, _ => Verified::NoMD5 })
}
