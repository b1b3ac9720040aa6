#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds oracle/_ref/libjst_ref.so: the UNMODIFIED reference
# (luigifcruz/CyberEther v1.9.1, "Jetstream") CPU compute path, compiled from the
# sources where they lie under /root/reference with plain g++ (no meson), plus
# our own C-ABI driver (oracle/ref_driver.cc) that feeds tensors to the reference
# Registry -> Module -> Runtime -> scheduler_synchronous -> Flowgraph objects.
#
# Nothing from /root/reference is copied into this repository; outputs go only
# to oracle/_ref/ (git-ignored, NOT gpurun-ignored: the .so travels to the GPU box).
# Recipe follows SURVEY.md Appendix A.  Usage: oracle/build_ref.sh [-j N]
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
R="${JST_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
B="$OUT/build"
JOBS="${JOBS:-$(nproc)}"
OPT="${REF_OPT:--O3}"           # reference default buildtype=release (meson.build:8); native=off (meson_options.txt:1)

if [ ! -d "$R/src" ]; then
  echo "[build_ref] $R not present; keeping prebuilt $OUT/libjst_ref.so (if any)"; exit 0
fi
mkdir -p "$B/obj" "$B/gen/jetstream" "$B/fmt_src"

# 1. jst::fmt headers: torch ships fmt; the reference expects the same headers in
#    namespace jst::fmt with JST_FMT_ macros (what its fmt wrap's patch step does).
if [ ! -f "$B/fmt_src/jetstream/fmt/format.h" ]; then
  SP="$(python3 -c 'import torch,os;print(os.path.dirname(torch.__file__))')/include"
  mkdir -p "$B/fmt_src/jetstream"
  rm -rf "$B/fmt_src/jetstream/fmt"; cp -r "$SP/fmt" "$B/fmt_src/jetstream/fmt"
  find "$B/fmt_src/jetstream/fmt" -type f \( -name '*.h' -o -name '*.cc' \) -print0 |
    xargs -0 sed -i -e 's/fmt::/jst::fmt::/g' -e 's/namespace fmt/namespace jst::fmt/g' -e 's/FMT_/JST_FMT_/g'
fi

# 2. config.hh (hand-written from include/jetstream/config.hh.in)
cat > "$B/gen/jetstream/config.hh" <<'EOF'
#define JETSTREAM_VERSION_STR "1.9.1"
#define JETSTREAM_VERSION_MAJOR 1
#define JETSTREAM_VERSION_MINOR 9
#define JETSTREAM_VERSION_PATCH 1
#define JETSTREAM_VERSION_ENCODE(major, minor, patch) ((((major) & 0xffu) << 16) | (((minor) & 0xffu) << 8) | ((patch) & 0xffu))
#define JETSTREAM_VERSION_CURRENT JETSTREAM_VERSION_ENCODE(JETSTREAM_VERSION_MAJOR, JETSTREAM_VERSION_MINOR, JETSTREAM_VERSION_PATCH)
#define JETSTREAM_BUILD_TYPE "release"
#define JETSTREAM_BUILD_OPTIMIZATION "3"
#define JETSTREAM_BUILD_DEBUG "false"
#define JETSTREAM_BUILD_NATIVE "false"
#define JST_IS_STATIC
#define JST_RELEASE_MODE
#define JST_OS_LINUX
#define JETSTREAM_LOADER_FMT_AVAILABLE
#define JETSTREAM_BACKEND_CPU_AVAILABLE
#define JETSTREAM_VIEWPORT_HEADLESS_AVAILABLE
EOF

INC="-I$B/gen -I$B/fmt_src -I$R/include -I$R/src -I$R/include/jetstream/render/tools -I$ROOT/shim/stubs"
CXXFLAGS="-std=c++20 $OPT -fPIC -DJST_FMT_HEADER_ONLY -w $INC"

CORE="logger memory/axis memory/buffer memory/buffer_cpu memory/tensor memory/token memory/types
 module module_impl module_context module_interface module_surface registry
 runtime/runtime runtime/native/cpu/impl runtime/native/cpu/context
 scheduler scheduler_context scheduler_synchronous tensor_link
 parser_map parser_encode parser_decode testing
 backend/base backend/devices/cpu/base
 platform/process platform/terminal platform/paths
 block block_impl block_context block_interface
 flowgraph flowgraph_environment flowgraph_metadata flowgraph_view"

# module TUs on the hot path
MODS="dsp/window dsp/fft dsp/amplitude dsp/invert dsp/fm dsp/filter_taps dsp/fold dsp/overlap_add
 dsp/phase_correction dsp/agc
 core/range core/multiply core/cast core/reshape core/pad core/unpad core/multiply_constant
 core/expand_dims core/ones_tensor"
BLOCKS="$MODS dsp/filter dsp/filter_engine dsp/spectrum_engine"

SRCS=()
for c in $CORE; do SRCS+=("$R/src/$c.cc"); done
for m in $MODS; do
  SRCS+=("$R/src/domains/$m/module_impl.cc" "$R/src/domains/$m/module_impl_native_cpu.cc")
done
for b in $BLOCKS; do
  [ -f "$R/src/domains/$b/block_impl.cc" ] && SRCS+=("$R/src/domains/$b/block_impl.cc")
done
# the reference's lineplot / waterfall COMPUTE TUs (their render halves are shim/viz_headless.cc)
for v in lineplot waterfall; do SRCS+=("$R/src/domains/visualization/$v/module_impl_native_cpu.cc"); done
SRCS+=("$HERE/ref_stubs.cc" "$ROOT/shim/viz_headless.cc" "$HERE/ref_driver.cc")

compile_one() {
  src="$1"; obj="$B/obj/$(echo "$src" | sed -e 's#^/##' -e 's#[/.]#_#g').o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ]; then
    g++ $CXXFLAGS -c "$src" -o "$obj" || { echo "FAILED: $src"; exit 1; }
  fi
}
export -f compile_one; export B CXXFLAGS
printf '%s\n' "${SRCS[@]}" | xargs -P "$JOBS" -I{} bash -c 'compile_one {}'

OBJS=()
for s in "${SRCS[@]}"; do OBJS+=("$B/obj/$(echo "$s" | sed -e 's#^/##' -e 's#[/.]#_#g').o"); done
g++ -shared -o "$OUT/libjst_ref.so" "${OBJS[@]}" -lpthread -ldl
echo "[build_ref] built $OUT/libjst_ref.so ($(stat -c %s "$OUT/libjst_ref.so") bytes)"
