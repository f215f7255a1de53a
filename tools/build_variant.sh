#!/bin/bash
# A/B build: recompile ONE source with extra nvcc flags and link it with the regular objects into a variant library.
# usage: tools/build_variant.sh <tag> <source.cu> <extra nvcc flags...>   ->  flac_codec_b200/libflacb200_<tag>.so
# select it with FLACB200_LIB=flac_codec_b200/libflacb200_<tag>.so
set -e
tag=$1; src=$2; shift 2
cd "$(dirname "$0")/../flac_codec_b200/csrc"
obj=../../build/flacb200/${src%.*}_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 -ccbin /usr/bin/g++ "$@" -c -o $obj $src
objs=""
for s in engine.cu encode_kernels.cu decode_kernels.cu synth.cu stream.cpp encode_frame.cu encode_analyze.cu encode_lpc.cu decode_parse.cu md5.cu md5_mb.cpp batch.cpp; do
  if [ "$s" = "$src" ]; then objs="$objs $obj"; else objs="$objs ../../build/flacb200/${s%.*}.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o ../libflacb200_$tag.so $objs
echo ../libflacb200_$tag.so
