#!/usr/bin/env bash
# Builds the reference-side binding of provider "b200" against the reference where it lies:
#
#   shim/_build/libjst_b200.so   the reference's Flowgraph / scheduler_synchronous / NativeCpuRuntime / CUDA backend singleton
#                                (compiled from /root/reference, CUDA enabled) + OUR CUDA tensor allocation and CUDA runtime
#                                behind the reference's interfaces (shim/b200_buffer.cc, shim/b200_runtime.cc) + the reference CPU
#                                modules and blocks + shim/b200_modules.cc + shim/b200_blocks.cc (provider "b200" ->
#                                libb200dsp.so; the latter REPLACES the spectrum_engine and filter block TUs) + the C-ABI
#                                harness shim/shim_capi.cc. Loaded by tests/ and bench.py through shim/binding.py.
#   shim/_build/shim_demo        stand-alone self-check over the same library (no Python).
#
# Outputs only under shim/_build/ (git-ignored; the binaries travel to the GPU box). Needs oracle/build_ref.sh to have
# prepared the jst::fmt headers.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
R="${JST_REFERENCE:-/root/reference}"
B="$HERE/_build"
FMT="$ROOT/oracle/_ref/build/fmt_src"
[ -d "$R/src" ] || { echo "[build_shim] $R absent; keeping prebuilt $B/libjst_b200.so"; exit 0; }
[ -f "$FMT/jetstream/fmt/format.h" ] || bash "$ROOT/oracle/build_ref.sh"
mkdir -p "$B/obj" "$B/gen/jetstream"
NEWCFG="$(sed -e '$a #define JETSTREAM_BACKEND_CUDA_AVAILABLE\n#define JETSTREAM_LOADER_CUDA_AVAILABLE' \
    "$ROOT/oracle/_ref/build/gen/jetstream/config.hh")"
if [ ! -f "$B/gen/jetstream/config.hh" ] || [ "$NEWCFG" != "$(cat "$B/gen/jetstream/config.hh")" ]; then
  printf '%s\n' "$NEWCFG" > "$B/gen/jetstream/config.hh"
fi
CUDA=/usr/local/cuda
INC="-I$B/gen -I$FMT -I$R/include -I$R/src -I$R/include/jetstream/render/tools -I$CUDA/include -I$ROOT/include -I$HERE -I$HERE/stubs"
CXXFLAGS="-std=c++20 -O2 -fPIC -DJST_FMT_HEADER_ONLY -w $INC"
# src/memory/buffer_cuda.cc, src/runtime/native/cuda/impl.cc and src/backend/devices/cuda/base.cc are NOT in this list:
# shim/b200_buffer.cc, shim/b200_runtime.cc and shim/b200_backend.cc define detail::CreateCudaBackend(),
# NativeCudaRuntimeFactory() and Backend::CUDA over the C ABI instead
CORE="logger memory/axis memory/buffer memory/buffer_cpu memory/tensor memory/token memory/types
 module module_impl module_context module_interface module_surface registry
 runtime/runtime runtime/native/cpu/impl runtime/native/cpu/context runtime/native/cuda/context
 scheduler scheduler_context scheduler_synchronous tensor_link
 parser_map parser_encode parser_decode
 backend/base backend/devices/cpu/base
 platform/process platform/terminal platform/paths
 block block_impl block_context block_interface
 flowgraph flowgraph_environment flowgraph_metadata flowgraph_view"
# reference CPU modules (provider "generic"): the comparison target inside the same Flowgraph runtime
MODS="dsp/window dsp/fft dsp/amplitude dsp/invert dsp/fm dsp/filter_taps dsp/fold dsp/overlap_add dsp/phase_correction
 dsp/agc core/range core/multiply core/cast core/reshape core/pad core/unpad core/multiply_constant core/expand_dims"
# reference blocks; spectrum_engine and filter come from shim/b200_blocks.cc (which compiles the reference TUs in place)
BLOCKS="dsp/fm dsp/agc dsp/amplitude dsp/fft core/cast core/range"
SRCS=()
for c in $CORE; do SRCS+=("$R/src/$c.cc"); done
for m in $MODS; do SRCS+=("$R/src/domains/$m/module_impl.cc" "$R/src/domains/$m/module_impl_native_cpu.cc"); done
for b in $BLOCKS; do [ -f "$R/src/domains/$b/block_impl.cc" ] && SRCS+=("$R/src/domains/$b/block_impl.cc"); done
# lineplot / waterfall: the reference's CPU compute TUs + blocks; their render halves are shim/viz_headless.cc
for v in lineplot waterfall; do
  SRCS+=("$R/src/domains/visualization/$v/module_impl_native_cpu.cc" "$R/src/domains/visualization/$v/block_impl.cc")
done
SRCS+=("$HERE/shim_stubs.cc" "$HERE/viz_headless.cc" "$HERE/b200_backend.cc" "$HERE/b200_buffer.cc" "$HERE/b200_runtime.cc" "$HERE/b200_modules.cc"
       "$HERE/b200_blocks.cc" "$HERE/shim_capi.cc")
objname() { echo "$B/obj/$(echo "$1" | sed -e 's#^/##' -e 's#[/.]#_#g').o"; }
compile_one() {
  src="$1"; obj="$B/obj/$(echo "$src" | sed -e 's#^/##' -e 's#[/.]#_#g').o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$HERE/b200_provider.hh" -nt "$obj" ] || [ "$ROOT/include/b200dsp.h" -nt "$obj" ]; then
    g++ $CXXFLAGS -c "$src" -o "$obj" || { echo "FAILED: $src"; exit 1; }
  fi
}
export -f compile_one; export B CXXFLAGS HERE ROOT
printf '%s\n' "${SRCS[@]}" | xargs -P "$(nproc)" -I{} bash -c 'compile_one {}'
OBJS=(); for s in "${SRCS[@]}"; do OBJS+=("$(objname "$s")"); done
g++ -shared -o "$B/libjst_b200.so" "${OBJS[@]}" -L"$ROOT/cyberether_b200" -lb200dsp \
    -Wl,-rpath,'$ORIGIN/../../cyberether_b200' \
    -L$CUDA/lib64 -lcudart -lnvrtc -L$CUDA/lib64/stubs -lcuda -lnvidia-ml -lpthread -ldl
g++ -std=c++17 -O2 -o "$B/shim_demo" "$HERE/shim_demo.cc" -L"$B" -ljst_b200 -Wl,-rpath,'$ORIGIN' \
    -Wl,-rpath,'$ORIGIN/../../cyberether_b200' -L"$ROOT/cyberether_b200" -lb200dsp -L$CUDA/lib64 -lcudart -lpthread -ldl -Wl,--allow-shlib-undefined
echo "[build_shim] built $B/libjst_b200.so and $B/shim_demo"
