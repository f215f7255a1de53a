//! Raw bindings of include/flacb200.h and include/flacb200_stream.h (keep in sync by hand; the ABI is plain C).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

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
pub struct flacb200_stream_params {
    pub sample_rate: u32,
    pub bits_per_sample: u32,
    pub channels: u32,
    pub subset: u32,
    pub max_block_size: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct flacb200_segment {
    pub pcm_offset: u64,
    pub n_pcm_frames: u64,
    pub first_frame_number: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct flacb200_decode_segment {
    pub byte_offset: u64,
    pub byte_length: u64,
    pub pcm_offset: u64,
    pub n_pcm_frames: u64,
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

pub enum flacb200_engine {}
pub enum flacb200_writer {}
pub enum flacb200_reader {}

pub const FLACB200_HOST: c_int = 0;
pub const FLACB200_PCM_BYTES_LE: c_int = 0;
pub const FLACB200_PCM_BYTES_BE: c_int = 1;
pub const FLACB200_PCM_I32_INTERLEAVED: c_int = 2;

unsafe extern "C" {
    pub fn flacb200_engine_create(device: c_int, out: *mut *mut flacb200_engine) -> c_int;
    pub fn flacb200_engine_destroy(e: *mut flacb200_engine);
    pub fn flacb200_encode(
        e: *mut flacb200_engine, opt: *const flacb200_options, params: *const flacb200_stream_params, pcm: *const c_void,
        pcm_bytes: usize, pcm_kind: c_int, pcm_location: c_int, planar_stride: u64, segments: *const flacb200_segment,
        n_segments: usize, out: *mut c_void, out_capacity: usize, out_location: c_int, frame_bytes: *mut u32,
        frame_bytes_capacity: usize, n_frames: *mut u64, total_bytes: *mut u64,
    ) -> c_int;
    pub fn flacb200_encode_bound(
        opt: *const flacb200_options, params: *const flacb200_stream_params, segments: *const flacb200_segment, n: usize,
    ) -> usize;
    pub fn flacb200_decode(
        e: *mut flacb200_engine, params: *const flacb200_stream_params, frames: *const c_void, frames_bytes: usize,
        frames_location: c_int, segments: *const flacb200_decode_segment, n_segments: usize, pcm_out: *mut c_void,
        pcm_out_bytes: usize, pcm_kind: c_int, pcm_location: c_int, planar_stride: u64, n_frames: *mut u64,
        n_pcm_frames: *mut u64, bad_frame: *mut u64,
    ) -> c_int;
    pub fn flacb200_strerror(code: c_int) -> *const c_char;
    /// MD5 of many streams (update_md5, src/encode.rs:1292): one 16-byte digest per segment
    pub fn flacb200_md5_batch(
        e: *mut flacb200_engine, pcm: *const c_void, pcm_bytes: usize, pcm_kind: c_int, pcm_location: c_int, planar_stride: u64,
        channels: u32, bits_per_sample: u32, segments: *const flacb200_segment, n_segments: usize, digests: *mut u8,
    ) -> c_int;

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

    pub fn flacb200_reader_open(e: *mut flacb200_engine, flac: *const u8, len: usize, out: *mut *mut flacb200_reader) -> c_int;
    pub fn flacb200_reader_close(r: *mut flacb200_reader);
    pub fn flacb200_reader_info(r: *mut flacb200_reader, si: *mut flacb200_streaminfo) -> c_int;
    pub fn flacb200_reader_read(r: *mut flacb200_reader, out: *mut c_void, capacity: usize, pcm_kind: c_int, n_out: *mut usize) -> c_int;
    pub fn flacb200_reader_seek(r: *mut flacb200_reader, pcm_frame: u64) -> c_int;
    pub fn flacb200_reader_verify(r: *mut flacb200_reader, result: *mut c_int, md5_out: *mut u8) -> c_int;
}
