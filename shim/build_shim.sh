#!/usr/bin/env bash
# Builds shim/_build/shim_demo: the reference's Flowgraph / scheduler_synchronous / NativeCudaRuntime / CUDA
# backend + memory (compiled from /root/reference where they lie, CUDA enabled) + the reference CPU modules +
# shim/b200_modules.cc (provider "b200" -> libb200dsp.so) + shim/shim_driver.cc. Outputs only under shim/_build/
# (git-ignored; the binary travels to the GPU box). Needs oracle/build_ref.sh to have prepared the fmt headers.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
R="${JST_REFERENCE:-/root/reference}"
B="$HERE/_build"
FMT="$ROOT/oracle/_ref/build/fmt_src"
[ -d "$R/src" ] || { echo "[build_shim] $R absent; keeping prebuilt $B/shim_demo"; exit 0; }
[ -f "$FMT/jetstream/fmt/format.h" ] || bash "$ROOT/oracle/build_ref.sh"
mkdir -p "$B/obj" "$B/gen/jetstream"
sed -e '$a #define JETSTREAM_BACKEND_CUDA_AVAILABLE\n#define JETSTREAM_LOADER_CUDA_AVAILABLE' \
    "$ROOT/oracle/_ref/build/gen/jetstream/config.hh" > "$B/gen/jetstream/config.hh"
CUDA=/usr/local/cuda
INC="-I$B/gen -I$FMT -I$R/include -I$R/src -I$R/include/jetstream/render/tools -I$CUDA/include -I$ROOT/include"
CXXFLAGS="-std=c++20 -O2 -fPIC -DJST_FMT_HEADER_ONLY -w $INC"
CORE="logger memory/axis memory/buffer memory/buffer_cpu memory/buffer_cuda memory/tensor memory/token memory/types
 module module_impl module_context module_interface module_surface registry
 runtime/runtime runtime/native/cpu/impl runtime/native/cpu/context runtime/native/cuda/impl runtime/native/cuda/context
 scheduler scheduler_context scheduler_synchronous tensor_link
 parser_map parser_encode parser_decode
 backend/base backend/devices/cpu/base backend/devices/cuda/base
 platform/process platform/terminal platform/paths
 block block_impl block_context block_interface
 flowgraph flowgraph_environment flowgraph_metadata flowgraph_view"
MODS="dsp/window dsp/fft dsp/amplitude dsp/invert dsp/agc core/range core/multiply core/cast core/reshape"
SRCS=()
for c in $CORE; do SRCS+=("$R/src/$c.cc"); done
for m in $MODS; do SRCS+=("$R/src/domains/$m/module_impl.cc" "$R/src/domains/$m/module_impl_native_cpu.cc"); done
SRCS+=("$R/src/domains/dsp/spectrum_engine/block_impl.cc" "$ROOT/oracle/ref_stubs.cc"
       "$HERE/b200_modules.cc" "$HERE/shim_driver.cc")
compile_one() {
  src="$1"; obj="$B/obj/$(echo "$src" | sed -e 's#^/##' -e 's#[/.]#_#g').o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then g++ $CXXFLAGS -c "$src" -o "$obj" || { echo "FAILED: $src"; exit 1; }; fi
}
export -f compile_one; export B CXXFLAGS
printf '%s\n' "${SRCS[@]}" | xargs -P "$(nproc)" -I{} bash -c 'compile_one {}'
OBJS=(); for s in "${SRCS[@]}"; do OBJS+=("$B/obj/$(echo "$s" | sed -e 's#^/##' -e 's#[/.]#_#g').o"); done
g++ -o "$B/shim_demo" "${OBJS[@]}" -L"$ROOT/cyberether_b200" -lb200dsp -Wl,-rpath,'$ORIGIN/../../cyberether_b200' \
    -L$CUDA/lib64 -lcudart -lnvrtc -L$CUDA/lib64/stubs -lcuda -lnvidia-ml -lpthread -ldl
echo "[build_shim] built $B/shim_demo"
