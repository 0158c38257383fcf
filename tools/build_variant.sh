#!/bin/bash
# build a tuning variant of the library: tools/build_variant.sh NAME [extra nvcc flags...]  ->  tools/_dbg/liblogmel_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o tools/_dbg/liblogmel_$name.so mlx8_ws_audio_transformer_b200/csrc/logmel_api.cu
echo built tools/_dbg/liblogmel_$name.so
