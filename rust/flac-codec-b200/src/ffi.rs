//! Raw bindings of include/flacb200.h and include/flacb200_stream.h (kept in sync by hand; the ABI is plain C:
//! tests/test_abi_exports.py checks that the library exports every symbol the headers declare).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_uint, c_void};

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct flacb200_options {
    pub block_size: u16,
    pub max_lpc_order: u8,
    pub max_partition_order: u8,
    pub mid_side: u8,
    pub exhaustive_channel_correlation: u8,
    pub window_kind: u8,
    pub reserved0: u8,
    pub tukey_p: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct flacb200_writer_options {
    pub frame: flacb200_options,
    pub padding: i32,
    pub seektable_kind: u32,
    pub seektable_n: u32,
    pub launch_frames: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct flacb200_streaminfo {
    pub min_block_size: u16,
    pub max_block_size: u16,
    pub min_frame_size: u32,
    pub max_frame_size: u32,
    pub sample_rate: u32,
    pub channels: u32,
    pub bits_per_sample: u32,
    pub total_samples: u64,
    pub md5: [u8; 16],
    pub frames_start: u64,
    pub n_seekpoints: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct flacb200_framebuf {
    pub samples: *const i32,
    pub n_samples: usize,
    pub sample_rate: u32,
    pub channels: u32,
    pub bits_per_sample: u32,
    pub block_size: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct flacb200_track {
    pub pcm: *const c_void,
    pub n_pcm_frames: u64,
    pub sample_rate: u32,
    pub bits_per_sample: u32,
    pub channels: u32,
    pub pcm_kind: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct flacb200_file {
    pub data: *mut u8,
    pub capacity: usize,
    pub len: usize,
    pub status: i32,
    pub frames: u32,
    pub md5: [u8; 16],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct flacb200_pcm {
    pub data: *mut c_void,
    pub capacity: usize,
    pub len: usize,
    pub status: i32,
    pub verified: i32,
    pub info: flacb200_streaminfo,
}

pub enum flacb200_engine {}
pub enum flacb200_writer {}
pub enum flacb200_reader {}
pub enum flacb200_stream_reader {}

pub const FLACB200_PCM_BYTES_LE: c_int = 0;
pub const FLACB200_PCM_BYTES_BE: c_int = 1;
pub const FLACB200_PCM_I32_INTERLEAVED: c_int = 2;
pub const FLACB200_NEED_DATA: c_int = -10;
pub const FLACB200_NEED_SEEK: c_int = -11;

unsafe extern "C" {
    pub fn flacb200_engine_create(device: c_int, out: *mut *mut flacb200_engine) -> c_int;
    pub fn flacb200_engine_destroy(e: *mut flacb200_engine);
    pub fn flacb200_strerror(code: c_int) -> *const c_char;

    // ---- writers (src/encode.rs:103-1290) ----
    pub fn flacb200_writer_open(
        e: *mut flacb200_engine, opt: *const flacb200_writer_options, sample_rate: u32, bits_per_sample: u32, channels: u32,
        total_pcm_frames: u64, out: *mut *mut flacb200_writer,
    ) -> c_int;
    pub fn flacb200_writer_close(w: *mut flacb200_writer);
    pub fn flacb200_total_from_bytes(total_bytes: u64, bps: u32, channels: u32, pcm_frames: *mut u64) -> c_int;
    pub fn flacb200_total_from_samples(total_samples: u64, channels: u32, pcm_frames: *mut u64) -> c_int;
    pub fn flacb200_writer_header(w: *mut flacb200_writer, bytes: *mut *const u8, len: *mut usize) -> c_int;
    pub fn flacb200_writer_write_bytes(w: *mut flacb200_writer, pcm: *const u8, n: usize, big_endian: c_int) -> c_int;
    pub fn flacb200_writer_write_samples(w: *mut flacb200_writer, s: *const i32, n: usize) -> c_int;
    pub fn flacb200_writer_write_channels(w: *mut flacb200_writer, ch: *const *const i32, nch: u32, n: usize) -> c_int;
    pub fn flacb200_writer_drain(w: *mut flacb200_writer, frames: *mut *const u8, len: *mut usize) -> c_int;
    pub fn flacb200_writer_flush(w: *mut flacb200_writer) -> c_int;
    pub fn flacb200_writer_finalize(w: *mut flacb200_writer) -> c_int;
    pub fn flacb200_stream_write(
        e: *mut flacb200_engine, opt: *const flacb200_options, sample_rate: u32, channels: u32, bits_per_sample: u32,
        samples: *const i32, n_samples: usize, frame_number: u64, out: *mut u8, out_capacity: usize, out_len: *mut usize,
    ) -> c_int;

    // ---- readers (src/decode.rs:103-1309) ----
    pub fn flacb200_reader_open_stream(e: *mut flacb200_engine, out: *mut *mut flacb200_reader) -> c_int;
    pub fn flacb200_reader_feed(r: *mut flacb200_reader, bytes: *const u8, len: usize, eof: c_int) -> c_int;
    pub fn flacb200_reader_set_seekable(r: *mut flacb200_reader, seekable: c_int) -> c_int;
    pub fn flacb200_reader_wanted_offset(r: *mut flacb200_reader, offset: *mut u64) -> c_int;
    pub fn flacb200_reader_close(r: *mut flacb200_reader);
    pub fn flacb200_reader_info(r: *mut flacb200_reader, si: *mut flacb200_streaminfo) -> c_int;
    pub fn flacb200_reader_read(r: *mut flacb200_reader, out: *mut c_void, capacity: usize, pcm_kind: c_int, n_out: *mut usize) -> c_int;
    pub fn flacb200_reader_fill_buf(r: *mut flacb200_reader, samples: *mut *const i32, n_samples: *mut usize) -> c_int;
    pub fn flacb200_reader_consume(r: *mut flacb200_reader, n_samples: usize) -> c_int;
    pub fn flacb200_reader_fill_channels(r: *mut flacb200_reader, channels: *mut *const *const i32, n_per_channel: *mut usize) -> c_int;
    pub fn flacb200_reader_consume_channels(r: *mut flacb200_reader, n_per_channel: usize) -> c_int;
    pub fn flacb200_reader_seek(r: *mut flacb200_reader, pcm_frame: u64) -> c_int;
    pub fn flacb200_reader_verify(r: *mut flacb200_reader, result: *mut c_int, md5_out: *mut u8) -> c_int;
    pub fn flacb200_stream_reader_open(e: *mut flacb200_engine, out: *mut *mut flacb200_stream_reader) -> c_int;
    pub fn flacb200_stream_reader_close(r: *mut flacb200_stream_reader);
    pub fn flacb200_stream_reader_feed(r: *mut flacb200_stream_reader, bytes: *const u8, len: usize, eof: c_int) -> c_int;
    pub fn flacb200_stream_reader_read(r: *mut flacb200_stream_reader, out: *mut flacb200_framebuf) -> c_int;

    // ---- whole-file batches over several GPUs (rayon fan-out of examples/flac2wav.rs:31-38) ----
    pub fn flacb200_encode_batch(
        tracks: *const flacb200_track, n_tracks: usize, opt: *const flacb200_writer_options, devices: *const c_int, n_devices: c_int,
        files: *mut flacb200_file,
    ) -> c_int;
    pub fn flacb200_files_free(files: *mut flacb200_file, n: usize);
    pub fn flacb200_decode_batch(
        flac: *const *const u8, flac_len: *const usize, n_files: usize, pcm_kind: c_int, verify: c_int, devices: *const c_int,
        n_devices: c_int, out: *mut flacb200_pcm,
    ) -> c_int;
    pub fn flacb200_pcm_free(out: *mut flacb200_pcm, n: usize);
    pub fn flacb200_md5_many(data: *const *const u8, len: *const usize, n: usize, digests: *mut u8, threads: c_uint);
}
